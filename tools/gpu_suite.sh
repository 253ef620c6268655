#!/bin/bash
# Runs on the GPU box (via gpurun): parity tests in isolated processes (a trapping kernel poisons its CUDA context),
# smoke, bench, and optional ncu captures.  Everything is logged under gpurun_out/.
#   tools/gpu_suite.sh [tests] [bench] [launches] [ncufull] [fpn] [fusion] [loop] [micro] [dist] [multi]
set +e
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
run() { # name, timeout, command...
  local name=$1; local to=$2; shift 2
  echo "=== $name" | tee -a $OUT/summary.txt
  timeout $to "$@" > $OUT/$name.log 2>&1
  echo "exit $? : $(tail -n 1 $OUT/$name.log | cut -c1-600)" | tee -a $OUT/summary.txt
}
: > $OUT/summary.txt
PT="python -m pytest -q --tb=short -p no:cacheprovider"
for what in "$@"; do
case $what in
tests)
  run t_corr 600 $PT tests/test_corr_gpu.py
  run t_geom 600 $PT tests/test_geom_gpu.py
  run t_update 600 $PT tests/test_update_gpu.py
  run t_step 900 $PT tests/test_step_gpu.py
  run t_net 600 $PT tests/test_net_gpu.py
  run t_fpn 900 $PT tests/test_fpn_gpu.py
  run t_edge 600 $PT tests/test_edge_cases_gpu.py
  run t_fusion 600 $PT tests/test_video_fusion_gpu.py
  run t_visual 300 $PT tests/test_visual_ops_gpu.py
  run t_config2 900 $PT tests/test_config2_gpu.py
  run t_dropin 600 $PT tests/test_dropin_gpu.py
  run t_misc 600 $PT tests/test_misc_gpu.py
  run t_modes 600 $PT tests/test_zy_graph_modes_gpu.py
  run t_loop 900 $PT tests/test_zz_slam_loop_gpu.py
  run smoke 300 python __graft_entry__.py smoke ;;
bench)
  run bench_fp16 900 python bench.py --steps 50 --warmup 5
  run bench_fp16_nograph 900 python bench.py --steps 20 --warmup 5 --no-graph --profile-only
  run bench_ref 900 python bench.py --impl reference --steps 3 --warmup 1 ;;
launches)
  run ncu_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-graph --profile-only ;;
ncufull)
  run ncu_full 1500 ncu --set full --clock-control none --import-source on -k regex:"corr_lookup_fused|conv_tc|heads_stencil|ba_pass|ba_solve|corr_volume_tc" -s 40 -c 30 -o $OUT/prof_r01 -f python bench.py --steps 2 --warmup 3 --no-graph --profile-only ;;
fpn)
  run t_fpn 900 $PT tests/test_fpn_gpu.py ;;
fusion)
  run t_fusion 600 $PT tests/test_video_fusion_gpu.py
  run bench_fusion 300 python tools/bench_fusion.py ;;
loop)      # the whole per-frame loop: decisions vs the reference golden, then the config-3 bench and its GPU-time profile
  run t_modes 600 $PT tests/test_zy_graph_modes_gpu.py
  run t_loop 900 $PT tests/test_zz_slam_loop_gpu.py tests/test_visual_ops_gpu.py
  run bench_loop 900 python tools/bench_loop.py
  run prof_loop 600 python tools/prof_loop.py ;;
micro)     # HBM-bound kernels next to their reference counterparts
  run bench_lookup 600 python tools/bench_lookup.py
  run bench_altcorr 600 python tools/bench_altcorr.py
  run diag_ba 600 env DIAG_BA_KERNELS=1 python tools/diag_ba.py 8 16 32 64
  run prof_solve 300 python tools/prof_solve.py 64        # needs tools/_prof/libpvo_ba_prof.so (python tools/prof_solve.py build, here)
  run ubench_fp64 120 python -c "import ctypes; ctypes.CDLL('tools/_prof/libubench.so').ubench_run()" ;;
dist)
  run t_dist 600 $PT tests/test_dist_gpu.py ;;
multi)
  run bench_2gpu 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 ;;
esac
done
cat $OUT/summary.txt
