"""Video-fusion frame rate of Panoptic-FPN (PanopticFPN.inference_fusion, second and later frames of a video: backbone on the
(previous, current) pair, flow transport, fusion conv, heads) at 480x640, frames and flow resident in HBM.

    python tools/bench_fusion.py [--steps 10] [--warmup 3]   ->  one JSON line (also written to gpurun_out/ when it exists)"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--height', type=int, default=480)
    ap.add_argument('--width', type=int, default=640)
    args = ap.parse_args()
    import torch
    from oracle import panoptic_fpn as ofpn          # seeded weights only
    from pvo_b200 import _lib
    from pvo_b200.panoptic_fpn import PanopticFPN
    m = PanopticFPN(ofpn.init_params(0, fusion=True), device='cuda')
    g = torch.Generator().manual_seed(2)
    frames = [torch.randint(0, 256, (3, args.height, args.width), generator=g).float().cuda() for _ in range(2)]
    flow = (torch.rand(2, args.height, args.width, generator=g) * 8 - 2).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    m.inference_fusion([{'image': frames[0], 'image_id': '0001_0', 'flow': flow}])
    for i in range(args.warmup):
        m.inference_fusion([{'image': frames[(i + 1) % 2], 'image_id': '0001_%d' % (i + 1), 'flow': flow}])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    _lib.LAUNCHES[0] = 0
    for i in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        e0.record()
        m.inference_fusion([{'image': frames[i % 2], 'image_id': '0001_x', 'flow': flow}])
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = sum(ms) / len(ms)
    line = {'metric': 'Panoptic-FPN video-fusion frames/sec (480x640, one frame per call, fp16)', 'value': round(1e3 / t, 2), 'unit': 'frames/s',
            'ms_per_step': round(t, 3), 'steps': args.steps, 'warmup': args.warmup, 'n_gpus': 1, 'dtype': 'f16',
            'gpu_launches': _lib.LAUNCHES[0] // max(args.steps, 1), 'data': 'synthetic frames and flow, seeded random-init weights',
            'config': {'workload': 'inference_fusion, eager launches (no CUDA graph yet)', 'l2': '256 MB flush between timed steps'}}
    print(json.dumps(line))
    if os.path.isdir(os.path.join(ROOT, 'gpurun_out')):
        json.dump(line, open(os.path.join(ROOT, 'gpurun_out', 'fusion_bench.json'), 'w'))


if __name__ == '__main__':
    main()
