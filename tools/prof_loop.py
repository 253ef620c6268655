"""Where does a steady-state frame of the per-frame loop spend its time?  torch.profiler over a few frames of tools/bench_loop.py's
workload: total CUDA kernel time per frame (GPU-bound vs host-bound) and the top kernels.  python tools/prof_loop.py [--frames 30]"""
import argparse
import collections
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=30)
    a = ap.parse_args()
    import torch
    from torch.profiler import profile, ProfilerActivity
    from tests.golden.make_golden import droid_net_state
    from pvo_b200.droid import Droid
    H, W = 240, 808
    args = types.SimpleNamespace(weights=droid_net_state(), device='cuda:0', image_size=[H, W], buffer=128, segm_filter=True,
                                 thresh=0.8, filter_thresh=0.0, warmup=12, beta=0.6, frontend_nms=1, keyframe_thresh=0.0, frontend_window=25,
                                 frontend_thresh=12.0, frontend_radius=2, backend_thresh=15.0, backend_radius=2, backend_nms=3, disable_vis=True)
    g = torch.Generator().manual_seed(7)
    tex = torch.nn.functional.interpolate(torch.rand(1, 3, 64, 160, generator=g), size=(H + 2 * a.frames + 8, W + 4 * a.frames + 8),
                                          mode='bicubic', align_corners=True).clamp(0, 1)[0]
    intr = torch.tensor([725.0087, 725.0087, 620.5, 187.0]) * (W / 1242.0)
    seg = torch.zeros(1, H // 8, W // 8, dtype=torch.int32)
    for k in range(4):
        seg[:, :, k * (W // 32):(k + 1) * (W // 32)] = (k + 1) * 10000
    frames = [(float(k), (tex[:, 2 * k:2 * k + H, 4 * k:4 * k + W] * 255).round().to(torch.uint8).cuda(), intr.cuda(), seg.cuda())
              for k in range(a.frames)]
    droid = Droid(args)
    n0 = 20
    for f in frames[:n0]:
        droid.track(*f[:2], None, f[2], f[3])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for f in frames[n0:]:
            droid.track(*f[:2], None, f[2], f[3])
        torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    nfr = len(frames) - n0
    kern = collections.defaultdict(lambda: [0, 0.0])
    total = 0.0
    for e in prof.events():
        if e.device_type is not None and str(e.device_type).endswith('CUDA'):
            kern[e.name][0] += 1
            kern[e.name][1] += e.device_time if hasattr(e, 'device_time') else e.cuda_time
            total += e.device_time if hasattr(e, 'device_time') else e.cuda_time
    top = sorted(kern.items(), key=lambda kv: -kv[1][1])[:25]
    out = {'frames': nfr, 'wall_ms_per_frame_under_profiler': round(1e3 * wall / nfr, 2), 'cuda_ms_per_frame': round(total / 1e3 / nfr, 3),
           'top': [{'name': k[:90], 'calls_per_frame': round(v[0] / nfr, 1), 'ms_per_frame': round(v[1] / 1e3 / nfr, 3)} for k, v in top]}
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
