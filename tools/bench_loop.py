"""the whole per-frame loop at the VKITTI2 shape (240x808 -> 30x101 grid, segment filter on),
pvo_b200.droid.Droid with seeded random-init weights on synthetic frames: MotionFilter + DroidFrontend per frame, then the
two global passes of DroidBackend (7 and 12 steps) and PoseTrajectoryFiller.

    python tools/bench_loop.py [--frames 60] [--no-backend]   ->  one JSON line (also gpurun_out/loop_bench.json)

Random weights do not measure motion, so the two data-dependent decisions are pinned to make the load deterministic:
filter_thresh = 0 (every frame becomes a keyframe; the motion filter's volume + lookup + update-operator call still runs) and
keyframe_thresh = 0 (no keyframe is removed): after the warm-up every frame costs 6 frame-graph update iterations over up to
48 edges, as a moving camera does in the reference's defaults (evaluation_scripts/test_vo.py:63-83)."""
import argparse
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=60)
    ap.add_argument('--height', type=int, default=240)
    ap.add_argument('--width', type=int, default=808)
    ap.add_argument('--no-backend', action='store_true')
    a = ap.parse_args()
    import torch
    from tests.golden.make_golden import droid_net_state          # seeded weights only
    from pvo_b200 import _lib
    from pvo_b200.droid import Droid
    from pvo_b200.factor_graph import FactorGraph
    H, W = a.height, a.width
    args = types.SimpleNamespace(weights=droid_net_state(), device='cuda:0', image_size=[H, W], buffer=max(128, a.frames + 32), segm_filter=True,
                                 thresh=0.8, filter_thresh=0.0, warmup=12, beta=0.6, frontend_nms=1, keyframe_thresh=0.0, frontend_window=25,
                                 frontend_thresh=12.0, frontend_radius=2, backend_thresh=15.0, backend_radius=2, backend_nms=3, disable_vis=True)
    g = torch.Generator().manual_seed(7)
    tex = torch.nn.functional.interpolate(torch.rand(1, 3, 64, 160, generator=g), size=(H + 2 * a.frames + 8, W + 4 * a.frames + 8),
                                          mode='bicubic', align_corners=True).clamp(0, 1)[0]
    intr = torch.tensor([725.0087, 725.0087, 620.5, 187.0]) * (W / 1242.0)                    # test_vo.py:21,52-54
    seg = torch.zeros(1, H // 8, W // 8, dtype=torch.int32)
    for k in range(4):
        seg[:, :, k * (W // 32):(k + 1) * (W // 32)] = (k + 1) * 10000                        # stuff bands
    for k in range(8):
        seg[:, 4 + 2 * k:10 + 2 * k, 20 + 9 * k:28 + 9 * k] = 120001 + k                      # rectangular things
    frames = [(float(k), (tex[:, 2 * k:2 * k + H, 4 * k:4 * k + W] * 255).round().to(torch.uint8).cuda(), intr.cuda(), seg.cuda())
              for k in range(a.frames)]
    droid = Droid(args)
    iters = [0]
    orig = FactorGraph.update

    def counted(self, *x, **k):
        iters[0] += 1
        return orig(self, *x, **k)
    FactorGraph.update = counted
    torch.cuda.synchronize()
    _lib.LAUNCHES[0] = 0
    per_frame = []
    t_all = time.perf_counter()
    for f in frames:
        t0 = time.perf_counter()
        droid.track(*f[:2], None, f[2], f[3])
        torch.cuda.synchronize()
        per_frame.append(time.perf_counter() - t0)
    front_s = time.perf_counter() - t_all
    front_iters, launches = iters[0], _lib.LAUNCHES[0]
    steady = per_frame[args.warmup + 1:]
    line = {'metric': 'VO loop frames/sec (config 3 shape 240x808, segment filter on, every frame a keyframe)', 'unit': 'frames/s',
            'value': round(len(steady) / sum(steady), 2), 'n_gpus': 1, 'dtype': 'f16', 'data': 'synthetic frames + panoptic ids, seeded random-init weights',
            'frames': a.frames, 'keyframes': droid.video.counter.value, 'frontend_s': round(front_s, 3), 'frontend_update_iterations': front_iters,
            'update_iterations_per_s': round(front_iters / front_s, 1), 'initialisation_s': round(per_frame[args.warmup - 1], 3),
            'steady_ms_per_frame': round(1e3 * sum(steady) / len(steady), 2), 'gpu_launches': launches,
            'config': {'workload': 'SURVEY 8(d) config 3 shape', 'warmup': args.warmup, 'max_factors': 48, 'timing': 'host clock with a device synchronize per frame'},
            'roofline_note': 'a frame = motion filter (2 feature encoders + 1 volume + 1 update-operator call) + ~5 new edge volumes + 6 window update '
                             'iterations (48 active + stored edges) + 12 dense-window BA solves; tools/prof_loop.py (torch.profiler) puts the CUDA time at '
                             '~13 ms of the frame, the rest is host work between the reference\'s own synchronisation points (.item() on the motion '
                             'magnitude and the keyframe distance, the proximity matrix to the host); the iterations inside are the config-2 kernels '
                             '(z|r conv 0.85 of the tensor peak, lookup 0.33 of HBM, bench.py)'}
    if not a.no_backend:
        t0 = time.perf_counter()
        droid.backend(7)
        droid.backend(12)
        torch.cuda.synchronize()
        line['backend_s'] = round(time.perf_counter() - t0, 3)
        t0 = time.perf_counter()
        traj = droid.traj_filler([(f[0], f[1], f[2], None) for f in frames])
        torch.cuda.synchronize()
        line['trajectory_filler_s'] = round(time.perf_counter() - t0, 3)
        line['trajectory_shape'] = list(traj.data.shape)
    print(json.dumps(line))
    if os.path.isdir(os.path.join(ROOT, 'gpurun_out')):
        json.dump(line, open(os.path.join(ROOT, 'gpurun_out', 'loop_bench.json'), 'w'))


if __name__ == '__main__':
    main()
