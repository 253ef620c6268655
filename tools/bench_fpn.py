"""Panoptic-FPN forward throughput (BASELINE.json configs[3]: R50-FPN, 480x640, batch 8, half tensor-core path).

    python tools/bench_fpn.py [--batch 8] [--steps 10] [--warmup 3] [--cpu-sample]

Prints one JSON line: images/s with the batch resident in HBM, e2e images/s with the images arriving from pinned host
memory and the panoptic id maps returning to the host, per-stage milliseconds (CUDA events), and the conv FLOP rate of
the backbone+FPN against the measured bf16 peak.  Random-init weights (oracle.panoptic_fpn.init_params: the seeded
initialisation shared with the parity tests), synthetic images.  --cpu-sample times the fp32 CPU oracle on one image.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def conv_flops(H, W):
    """MAC*2 per image of every convolution / FC that runs on the tensor cores (fixed-size part: backbone, FPN,
    sem-seg head, RPN head).  The ROI heads depend on the number of proposals and are reported separately."""
    from pvo_b200.panoptic_fpn import RES_STAGES
    f = 0
    h, w = (H + 1) // 2, (W + 1) // 2
    f += h * w * 64 * 147                                      # stem (CUDA cores)
    h, w = (h + 1) // 2, (w + 1) // 2
    cin = 64
    for si, (nb, mid, cout) in enumerate(RES_STAGES):
        if si > 0:
            h, w = (h + 1) // 2, (w + 1) // 2
        for bi in range(nb):
            f += h * w * (cin * mid + mid * mid * 9 + mid * cout + (cin * cout if bi == 0 else 0))
            cin = cout
    h4, w4 = (H + 3) // 4, (W + 3) // 4
    for lvl, c in zip(range(4), (256, 512, 1024, 2048)):
        hh, ww = h4 >> lvl, w4 >> lvl
        f += hh * ww * (c * 256 + 256 * 256 * 9)               # lateral + output
        f += hh * ww * (256 * 256 * 9 + 256 * 15)              # RPN head
        n = max(1, lvl)
        for j in range(n):                                     # sem-seg head convs at shrinking scale factors
            f += (hh << j) * (ww << j) * (256 if j == 0 else 128) * 128 * 9
    f += (h4 >> 4) * (w4 >> 4) * (256 * 256 * 9 + 256 * 15)    # RPN on p6
    f += h4 * w4 * 128 * 12
    return 2.0 * f


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--height', type=int, default=480)
    ap.add_argument('--width', type=int, default=640)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--cpu-sample', action='store_true')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--profile-only', action='store_true', help='warm up, run --steps inferences, print nothing (for ncu)')
    args = ap.parse_args()
    import torch
    from oracle import panoptic_fpn as ofpn          # seeded weights only (and the optional CPU baseline leg)
    from pvo_b200 import _lib
    from pvo_b200.panoptic_fpn import PanopticFPN
    sd = ofpn.init_params(0)
    g = torch.Generator().manual_seed(2)                    # U{0..255}, seed 2
    host = [torch.randint(0, 256, (3, args.height, args.width), generator=g).float().pin_memory() for _ in range(args.batch)]
    m = PanopticFPN(sd, device='cuda', use_cuda_graph=not args.no_graph)
    dev_imgs = [h.cuda() for h in host]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')

    stages = {}

    def staged():
        ev = []

        def mark(name):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            ev.append((name, e))
        mark('start')
        x, sizes = m.preprocess(dev_imgs); mark('preprocess')
        feats, sem, rpn = m.dense(x); mark('dense_backbone_fpn_semseg_rpn')
        props = m.proposals(rpn, sizes); mark('proposals_nms')
        cls, reg = m.box_head(m.roi_pool(feats, [p[0] for p in props], 7)); mark('roi_align7_box_head')
        dets = m.detections(cls, reg, props, sizes); mark('detections_nms')
        ml = m.mask_head(m.roi_pool(feats, [d[0] for d in dets], 14)); mark('roi_align14_mask_head')
        cls_all = torch.cat([d[2] for d in dets])
        prob = ml[torch.arange(ml.shape[0], device='cuda'), cls_all][:, None].sigmoid()
        masks = m.paste(prob, torch.cat([d[0] for d in dets]), sizes[0][0], sizes[0][1])
        per_image, o = [], 0
        for n, (b, s_, c) in enumerate(dets):
            per_image.append((s_, c, masks[o:o + b.shape[0]]))
            o += b.shape[0]
        mark('paste_masks')
        labels = sem[:, :sizes[0][0], :sizes[0][1], :12].argmax(dim=-1); mark('sem_argmax')
        m.combine(per_image, labels, boxes=[d[0] for d in dets]); mark('panoptic_merge')
        torch.cuda.synchronize()
        for (n0, e0), (n1, e1) in zip(ev[:-1], ev[1:]):
            stages.setdefault(n1, []).append(e0.elapsed_time(e1))
        return sum(p[0].shape[0] for p in props), sum(d[0].shape[0] for d in dets)

    for _ in range(args.warmup):
        m.inference(dev_imgs)
    torch.cuda.synchronize()
    if args.profile_only:
        for _ in range(args.steps):
            m.inference(dev_imgs)
        torch.cuda.synchronize()
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    _lib.LAUNCHES[0] = 0
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        e0.record()
        m.inference(dev_imgs)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    launches = _lib.LAUNCHES[0] // max(args.steps, 1)
    ms_step = sum(ms) / len(ms)
    # e2e: pinned host images in, panoptic ids out
    t = []
    for _ in range(max(3, args.steps // 2)):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = m.inference([h.cuda(non_blocking=True) for h in host])
        pans = [o['panoptic_seg'][0].cpu() for o in out]
        torch.cuda.synchronize()
        t.append(time.perf_counter() - t0)
    e2e = args.batch / (sum(t) / len(t))
    nprops = ndets = 0
    for _ in range(3):
        nprops, ndets = staged()
    st = {k: round(sum(v[1:]) / max(len(v) - 1, 1), 4) for k, v in stages.items()}
    fl = conv_flops(args.height, args.width) * args.batch
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
    peak = peaks.get('bf16_tflops_sustained', peaks.get('bf16_tflops', 1400.0))
    dense_ms = st['dense_backbone_fpn_semseg_rpn']
    line = {'metric': 'Panoptic-FPN R50 forward images/sec (480x640, batch 8, fp16)', 'value': round(args.batch / ms_step * 1e3, 2),
            'unit': 'images/s', 'n_gpus': 1, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': round(ms_step, 3),
            'higher_is_better': True, 'dtype': 'f16', 'data': 'synthetic images, seeded random-init weights',
            'config': {'workload': 'BASELINE configs[3]', 'batch': args.batch, 'image': [args.height, args.width],
                       'l2': '256 MB flush between timed steps'},
            'e2e': {'value': round(e2e, 2), 'unit': 'images/s', 'h2d_bytes_per_step': args.batch * 3 * args.height * args.width * 4,
                    'd2h_bytes_per_step': args.batch * args.height * args.width * 4},
            'gpu_launches': launches, 'stage_ms': st, 'proposals': nprops, 'detections': ndets,
            'roofline': {'bound': 'tensor', 'kernel': 'conv_tc_kernel: all launches of the captured dense graph (backbone + FPN + sem-seg + RPN heads; includes its pooling / norm / upsample kernels)',
                         'achieved': round(fl / (dense_ms * 1e-3) / 1e12, 1), 'peak': peak, 'unit': 'TFLOP/s',
                         'frac': round(fl / (dense_ms * 1e-3) / 1e12 / peak, 4), 'flops_per_step': fl, 'traffic': None}}
    if args.cpu_sample:
        torch.set_num_threads(min(32, os.cpu_count() or 1))
        t0 = time.perf_counter()
        ofpn.inference(sd, [host[0]])
        dt = time.perf_counter() - t0
        line['cpu_baseline'] = {'value': round(1.0 / dt, 4), 'unit': 'images/s', 'cores': torch.get_num_threads(), 'kind': 'port',
                                'sample': '1 image through the fp32 CPU oracle: %.2f s' % dt}
    print(json.dumps(line))


if __name__ == '__main__':
    main()
