"""BA timing vs window size (2 Gauss-Newton iterations, 60x80 grid, neighbourhood radius 3): the per-rank cost of the
redundant solve when a frame graph is sharded over 1 / 2 / 4 / 8 GPUs (P = 7 / 15 / 31 / 63 free poses)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from tests.golden import cases
from pvo_b200 import droid_backends
RADIUS = int(os.environ.get('DIAG_BA_RADIUS', '3'))          # 8 = BASELINE configs[4] (banded-8 graph)
for F in ([int(a) for a in sys.argv[1:]] or [8, 16, 32, 64]):
    c = cases.graph_case(F=F, radius=RADIUS, h=60, w=80, seed=1)
    ii, jj = c['ii'].cuda(), c['jj'].cuda()
    own = ii < 8                                    # one rank's shard: the out-edges of its 8 keyframes
    ii, jj = ii[own].contiguous(), jj[own].contiguous()
    E = ii.shape[0]
    g = torch.Generator(device='cuda').manual_seed(0)
    poses, disps = c['poses'].cuda(), c['disps'].cuda()
    intr = c['intrinsics'][0].cuda().contiguous()
    tgt = torch.rand(E, 2, 60, 80, device='cuda', generator=g) * torch.tensor([80.0, 60.0], device='cuda').view(1, 2, 1, 1)
    wgt = torch.rand(E, 2, 60, 80, device='cuda', generator=g)
    kx = torch.unique(torch.cat([torch.arange(1, F, device='cuda'), ii]))
    eta = torch.full((kx.shape[0], 60, 80), 0.02, device='cuda')
    def run():
        p, d = poses.clone(), disps.clone()
        droid_backends.ba(p, d, intr, tgt, wgt, eta, ii, jj, 1, F, 2, 1e-4, 0.1, False, max_span=RADIUS)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    print('radius %d  F=%2d  P=%2d  edges on this rank=%3d   2 GN iterations: %.1f us' % (RADIUS, F, F - 1, E, e0.elapsed_time(e1) / 20 * 1e3), flush=True)
    if os.environ.get('DIAG_BA_KERNELS'):           # per-kernel GPU time of one call (torch.profiler / CUPTI)
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(5): run()
            torch.cuda.synchronize()
        for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:12]:
            print('      %-90s n=%3d  avg %.1f us' % (ev.key[:90], ev.count, ev.device_time_total / ev.count), flush=True)
