"""Timing attribution for conv_tc_kernel on FPN-shaped layers (PVO_TC_DEBUG: 1 no A loads, 2 no B loads, 4 no MMAs, 8 no stores)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from pvo_b200 import nn as pnn, _lib
shapes = [('res2.conv2 3x3 64->64 @120x160', 8, 120, 160, 64, 64, 3), ('res2.conv3 1x1 64->256', 8, 120, 160, 64, 256, 1),
          ('res3.conv2 3x3 128->128 @60x80', 8, 60, 80, 128, 128, 3), ('fpn_out2 3x3 256->256 @120x160', 8, 120, 160, 256, 256, 3),
          ('res4.conv2 3x3 256->256 @30x40', 8, 30, 40, 256, 256, 3), ('res4.conv3 1x1 256->1024 @30x40', 8, 30, 40, 256, 1024, 1)]
for name, N, H, W, ci, co, k in shapes:
    w = torch.randn(co, ci, k, k, device='cuda') * 0.05
    cv = pnn._Conv(w, torch.zeros(co, device='cuda'), tc=True, simt=False)
    x = torch.randn(N * H * W, ci, device='cuda').half()
    out = torch.empty(N * H * W, co, device='cuda', dtype=torch.float16)
    def run():
        pnn.conv2d(cv, [(x, ci, ci)], N, H, W, _lib.F16, out, co, 0, 1, use_tc=True)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    fl = 2.0 * N * H * W * ci * co * k * k
    print('%-36s dbg=%s  %8.1f us  %7.1f TF' % (name, os.environ.get('PVO_TC_DEBUG', '0'), us, fl / us / 1e6), flush=True)
