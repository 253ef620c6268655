"""Correlation lookup + volume build micro-benchmark at BASELINE config 2 (E=36, 60x80): planar (reference layout) vs the tiled
private layout, CUDA-event timed, L2 flushed between launches.  Also the warm volume build (per new edge) against its HBM-write
roofline.  Prints one JSON object; `--ncu` runs a short sequence suited to `ncu -k regex:corr_` captures.

    python tools/bench_lookup.py [--reps 20] [--granularity 32|64|128] [--ncu]
"""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def set_l2_granularity(nbytes):
    rt = ctypes.CDLL('libcudart.so')
    cudaLimitMaxL2FetchGranularity = 0x05
    rc = rt.cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, ctypes.c_size_t(nbytes))
    v = ctypes.c_size_t(0)
    rt.cudaDeviceGetLimit(ctypes.byref(v), cudaLimitMaxL2FetchGranularity)
    return rc, v.value


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reps', type=int, default=20)
    ap.add_argument('--granularity', type=int, default=0)
    ap.add_argument('--ncu', action='store_true')
    args = ap.parse_args()
    from tests.golden import cases
    from pvo_b200.depth_video import DepthVideo
    from pvo_b200.modules.corr import CorrBlock, build_pyramid, build_pyramid_tiled
    out = {}
    if args.granularity:
        out['l2_fetch_granularity'] = set_l2_granularity(args.granularity)
    c = cases.graph_case()
    F_, h, w, E = c['F'], c['h'], c['w'], c['E']
    v = DepthVideo(image_size=(8 * h, 8 * w), buffer=F_ + 3, device='cuda', store_images=False)
    v.poses[:F_] = c['poses'].cuda()
    v.disps[:F_] = c['disps'].cuda()
    v.intrinsics[:] = c['intrinsics'][0].cuda()
    ii, jj = c['ii'].cuda(), c['jj'].cuda()
    coords, _ = v.reproject(ii, jj)                                  # (1,E,h,w,2): what the graph looks up
    coords = (coords + 0.3 * torch.randn_like(coords)).contiguous()
    f = c['fmaps'].cuda()
    f1, f2 = f[ii][None], f[jj][None]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')

    def timed(fn, reps):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        return ts[len(ts) // 2]

    pm = torch.empty(E * h * w, 208, dtype=torch.float16, device='cuda')
    alg = E * h * w * (4 * 64 * 2 + 8 + 208 * 2)
    reps = 3 if args.ncu else args.reps
    for name, tiled in (('planar', False), ('tiled', True)):
        blk = CorrBlock(f1, f2, tiled=tiled)
        ms = timed(lambda: blk(coords, out=pm, layout=1, out_cstride=208), reps)
        out['lookup_' + name] = {'ms': round(ms, 4), 'algorithmic_GBps': round(alg / ms / 1e6, 1)}
        del blk
        torch.cuda.empty_cache()
    out['lookup_algorithmic_bytes'] = alg
    # volume build, warm: (E,128,h,w) x2 -> 4 levels; algorithmic bytes = the pyramid written once + the two feature maps read
    g1, g2 = f[ii].contiguous(), f[jj].contiguous()
    vol_bytes = E * h * w * sum((h >> l) * (w >> l) for l in range(4)) * 2 + 2 * E * 128 * h * w * 2
    for name, fn in (('planar', lambda: build_pyramid(g1, g2, 4, use_tc=True)), ('tiled', lambda: build_pyramid_tiled(g1, g2))):
        ms = timed(fn, max(3, reps // 2))
        out['volume_' + name] = {'ms': round(ms, 3), 'ms_per_edge': round(ms / E, 4), 'write_GBps': round(vol_bytes / ms / 1e6, 1)}
        torch.cuda.empty_cache()
    out['volume_algorithmic_bytes'] = vol_bytes
    print(json.dumps(out))


if __name__ == '__main__':
    main()
