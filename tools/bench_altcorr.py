"""On-the-fly correlation (AltCorrBlock / altcorr_forward, the global-BA path) at BASELINE config 2: this package's kernel next to
the reference's own kernel (oracle/_ref, same inputs), per pyramid level and for the whole 4-level call.  CUDA-event timed.
    python tools/bench_altcorr.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    from tests.golden import cases
    from pvo_b200 import droid_backends
    from pvo_b200.depth_video import DepthVideo
    from pvo_b200.modules.corr import AltCorrBlock
    try:
        from oracle import build_ref
        ref = build_ref.load()
    except Exception:
        ref = None
    c = cases.graph_case()
    F_, h, w, E = c['F'], c['h'], c['w'], c['E']
    v = DepthVideo(image_size=(8 * h, 8 * w), buffer=F_ + 3, device='cuda', store_images=False)
    v.poses[:F_] = c['poses'].cuda()
    v.disps[:F_] = c['disps'].cuda()
    v.intrinsics[:] = c['intrinsics'][0].cuda()
    ii, jj = c['ii'].cuda(), c['jj'].cuda()
    coords, _ = v.reproject(ii, jj)
    fm = c['fmaps'].cuda()[None]
    blk = AltCorrBlock(fm)
    out = {'edges': E, 'grid': [h, w]}
    out['altcorrblock_4_levels_ms'] = round(timed(lambda: blk(coords, ii, jj)), 3)
    per = {}
    for l in range(4):
        f1 = blk.pyramid[0][0, ii].float().contiguous()
        f2 = blk.pyramid[l][0, jj].float().contiguous()
        cl = (coords[0] / 2 ** l)[:, None].contiguous()
        ours = timed(lambda: droid_backends.altcorr_forward(f1, f2, cl, 3))
        row = {'ours_ms': round(ours, 4), 'gflop': round(2.0 * E * h * w * 64 * 128 / 1e9, 2)}
        if ref is not None:
            row['reference_kernel_ms'] = round(timed(lambda: ref.altcorr_forward(f1, f2, cl, 3)), 4)
            a, b = droid_backends.altcorr_forward(f1, f2, cl, 3)[0], ref.altcorr_forward(f1, f2, cl, 3)[0]
            row['max_abs_diff'] = float((a - b).abs().max())
        per['level%d' % l] = row
    out['per_level'] = per
    print(json.dumps(out))


if __name__ == '__main__':
    main()
