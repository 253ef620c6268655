#!/bin/bash
# Runs on the GPU box (via gpurun): parity tests in isolated processes (a trapping kernel poisons its CUDA context),
# smoke, bench, and an ncu launch list.  Everything is logged under gpurun_out/.
set +e
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
run() { # name, timeout, command...
  local name=$1; local to=$2; shift 2
  echo "=== $name" | tee -a $OUT/summary.txt
  timeout $to "$@" > $OUT/$name.log 2>&1
  echo "exit $? : $(tail -n 1 $OUT/$name.log)" | tee -a $OUT/summary.txt
}
: > $OUT/summary.txt
PT="python -m pytest -q --tb=short -p no:cacheprovider"
run t_corr_simt 600 $PT tests/test_corr_gpu.py -k "not tensor_cores and not full_size"
run t_geom 600 $PT tests/test_geom_gpu.py
run t_update_fp32 600 $PT tests/test_update_gpu.py -k "fp32 or gru"
run t_step_fp32 900 $PT tests/test_step_gpu.py -k "golden or medium or bookkeeping"
run t_conv_tc 300 $PT tests/test_update_gpu.py -k "tensor_core"
run t_vol_tc 300 $PT tests/test_corr_gpu.py -k "tensor_cores or full_size"
run t_half 300 $PT tests/test_update_gpu.py tests/test_step_gpu.py -k "half"
run smoke 300 python __graft_entry__.py smoke
run bench_fp32 600 python bench.py --steps 3 --warmup 3 --precision fp32
run bench_fp16 600 python bench.py --steps 30 --warmup 5
if [ "$1" == "ncu" ]; then
  run ncu_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3
fi
cat $OUT/summary.txt
