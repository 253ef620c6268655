"""Panoptic-FPN forward on the B200 kernels vs the fp32 oracle (oracle/panoptic_fpn.py), SURVEY.md section 8 row (g).

Tensor-level gates (half operands, fp32 accumulation, ~55 layers deep): relative L2 error per stage with the discrete
decisions (top-k, NMS) taken from the ORACLE so the same boxes are pooled on both sides; the panoptic merge is integer
work and must be bit-exact on identical inputs; the composed output is compared by pixel agreement."""
import pytest
import torch

from oracle import panoptic_fpn as ofpn

pytestmark = pytest.mark.gpu

H, W = 128, 192


NOTES = {}


def rel(a, b, name=None):
    a, b = a.float().cpu(), b.float().cpu()
    e = ((a - b).norm() / b.norm().clamp(min=1e-12)).item()
    if name:
        NOTES[name] = e
    return e


@pytest.fixture(scope='module')
def setup():
    from pvo_b200.panoptic_fpn import PanopticFPN
    sd = ofpn.init_params(0)
    g = torch.Generator().manual_seed(3)
    images = [torch.randint(0, 256, (3, H, W), generator=g).float() for _ in range(2)]
    # smooth the noise so the network sees image-like statistics
    images = [torch.nn.functional.avg_pool2d(im[None], 5, 1, 2)[0] * 2 - 64 for im in images]
    dbg = {}
    ref = ofpn.inference(sd, images, debug=dbg)
    m = PanopticFPN(sd, device='cuda')
    yield sd, images, ref, dbg, m
    import json, os
    if os.path.isdir('gpurun_out'):                       # measured errors, kept with the run's evidence
        json.dump(NOTES, open('gpurun_out/fpn_parity_errors.json', 'w'), indent=1)


def nchw(t, N, h, w):
    return t.view(N, h, w, -1).permute(0, 3, 1, 2)


def test_backbone_and_heads(setup):
    sd, images, ref, dbg, m = setup
    x, sizes = m.preprocess(images)
    feats = m.backbone(x)
    for k in ('p2', 'p3', 'p4', 'p5', 'p6'):
        t, h, w = feats[k]
        e = rel(nchw(t, 2, h, w), dbg['feats'][k], k)
        assert e < 2e-2, (k, e)
    sem = m.sem_seg(feats)
    e = rel(sem[..., :12].permute(0, 3, 1, 2), dbg['sem'], 'sem_logits')
    assert e < 2e-2, e
    assert (sem[..., :12].argmax(-1).cpu() == dbg['sem'].argmax(1)).float().mean() > 0.97
    rpn = m.rpn_head(feats)
    for i, (o, rl, rd) in enumerate(zip(rpn, dbg['rpn_logits'], dbg['rpn_deltas'])):
        N, A, h, w = rl.shape
        lg, dl = o[..., :3].float(), o[..., 3:15].reshape(N, h, w, 3, 4).float()
        assert rel(lg, rl.permute(0, 2, 3, 1), 'rpn_logits_p%d' % (i + 2)) < 3e-2
        assert rel(dl, rd.view(N, A, 4, h, w).permute(0, 3, 4, 1, 2), 'rpn_deltas_p%d' % (i + 2)) < 3e-2


def test_roi_heads_on_oracle_boxes(setup):
    sd, images, ref, dbg, m = setup
    x, sizes = m.preprocess(images)
    feats = m.backbone(x)
    props = [(p[0].cuda(), p[1].cuda()) for p in dbg['proposals']]
    pooled = m.roi_pool(feats, [p[0] for p in props], 7)
    ref_pooled = ofpn.roi_pool(dbg['feats'], [p[0] for p in dbg['proposals']], 7)
    assert rel(pooled.permute(0, 3, 1, 2), ref_pooled, 'roi_align_7') < 5e-3
    cls, reg = m.box_head(pooled)
    # the class logits are a small-gain 1024-term sum: absolute error relative to the logit scale
    assert rel(cls, dbg['box_cls'], 'box_cls') < 5e-2
    assert rel(reg, dbg['box_reg'], 'box_reg') < 5e-2
    dets = [tuple(t.cuda() for t in d) for d in dbg['detections']]
    ml = m.mask_head(m.roi_pool(feats, [d[0] for d in dets], 14))
    assert ml.shape == dbg['mask_logits'].shape
    assert rel(ml, dbg['mask_logits'], 'mask_logits') < 3e-2


def test_proposals_and_detections_on_oracle_head_outputs(setup):
    """the batched decode / clip / NMS / top-k stages reproduce the oracle's proposal and detection lists when both sides
    start from the same (half-rounded) head outputs"""
    sd, images, ref, dbg, m = setup
    m.N = 2
    sizes = [(H, W), (H, W)]
    packed, lgs, dls = [], [], []
    # objectness goes through half: map the oracle's logits monotonically onto DISTINCT half values so that neither side
    # has score ties (tie order is unspecified in top-k / sort / NMS on both sides)
    halves = torch.arange(0x3000, 0x4800, dtype=torch.int32).to(torch.int16).view(torch.float16).float()     # 0.125 .. 8, all distinct
    halves = torch.cat([-halves.flip(0), halves])
    for rl, rd in zip(dbg['rpn_logits'], dbg['rpn_deltas']):
        N, A, h, w = rl.shape
        rank = rl.flatten().argsort().argsort()
        rl = halves[(rank * (halves.numel() // rank.numel())).long()].view(rl.shape)
        o = torch.cat([rl.permute(0, 2, 3, 1), rd.permute(0, 2, 3, 1), torch.zeros(N, h, w, 1)], -1).half()
        packed.append(o.cuda().contiguous())
        lgs.append(o[..., :3].float().permute(0, 3, 1, 2).contiguous())
        dls.append(o[..., 3:15].float().permute(0, 3, 1, 2).contiguous())
    want = ofpn.rpn_proposals(lgs, dls, sizes)
    got = m.proposals(packed, sizes)
    def canon(b, s_):                                                 # the same half value can occur on two levels: order ties by box
        import numpy as np
        b, s_ = b.cpu().numpy(), s_.cpu().numpy()
        o = np.lexsort((np.round(b[:, 3], 2), np.round(b[:, 2], 2), np.round(b[:, 1], 2), np.round(b[:, 0], 2), -s_))
        return torch.from_numpy(b[o]), torch.from_numpy(s_[o])
    for (wb, ws), (gb, gs) in zip(want, got):
        assert gb.shape == wb.shape
        assert torch.equal(gs.cpu(), ws)                              # same survivors, same order
        (wb, ws), (gb, gs) = canon(wb, ws), canon(gb, gs)
        assert (gb.cpu() - wb).abs().max() < 1e-4                     # decode + clip (expf differs by an ulp between CPU and GPU)
    props = [(b.cuda(), s_.cuda()) for b, s_ in dbg['proposals']]        # the rows box_cls / box_reg were computed for
    cls, reg = dbg['box_cls'].cuda(), dbg['box_reg'].cuda()
    dets = m.detections(cls, reg, props, sizes)
    o = 0
    for n, (pb, _) in enumerate(dbg['proposals']):
        R = pb.shape[0]
        boxes = ofpn.apply_deltas(dbg['box_reg'][o:o + R], pb, (10.0, 10.0, 5.0, 5.0))
        scores = torch.softmax(dbg['box_cls'][o:o + R], dim=-1)
        wb, ws, wc = ofpn.fast_rcnn_inference_single(boxes, scores, sizes[n])
        o += R
        gb, gs, gc = dets[n]
        assert gb.shape == wb.shape and torch.equal(gc.cpu(), wc)
        assert (gs.cpu() - ws).abs().max() < 1e-6 and (gb.cpu() - wb).abs().max() < 1e-4


def test_nms_groups_match_torchvision():
    """greedy NMS on chip == torchvision.ops.nms (what layers/nms.py calls) on every group, bit for bit"""
    import torchvision
    from pvo_b200.panoptic_fpn import PanopticFPN
    m = PanopticFPN.__new__(PanopticFPN)
    m.device = torch.device('cuda')
    g = torch.Generator().manual_seed(5)
    for n, G, thr in ((1000, 6, 0.7), (333, 3, 0.5), (1024, 2, 0.3), (1, 2, 0.5)):
        ctr = torch.rand(G, n, 2, generator=g) * 200
        wh = torch.rand(G, n, 2, generator=g) * 60 + 1
        boxes = torch.cat([ctr - wh / 2, ctr + wh / 2], -1).cuda()
        boxes[:, n // 2:n // 2 + 3] = boxes[:, :3]                     # exact duplicates
        scores = torch.rand(G, n, generator=g).cuda().sort(dim=1, descending=True)[0]
        valid = (torch.rand(G, n, generator=g) > 0.1).cuda()
        keep = m.nms_groups(boxes, valid, thr)
        for k in range(G):
            vi = valid[k].nonzero().reshape(-1)
            ref = torch.zeros(n, dtype=torch.bool, device='cuda')
            ref[vi[torchvision.ops.nms(boxes[k, vi], scores[k, vi], thr)]] = True
            assert torch.equal(keep[k], ref), (n, k)


def test_roi_align_fpn_matches_torchvision():
    """pixel-major ROIAlign + level assignment vs the oracle's torchvision path on fp32 features"""
    from pvo_b200.panoptic_fpn import PanopticFPN
    from pvo_b200 import _lib
    from pvo_b200._lib import check, ptr, stream
    g = torch.Generator().manual_seed(9)
    N, Hi, Wi = 2, 256, 320
    feats = {'p%d' % l: torch.randn(N, 256, Hi >> l, Wi >> l, generator=g) for l in (2, 3, 4, 5)}
    boxes = []
    for _ in range(N):
        c = torch.rand(40, 2, generator=g) * torch.tensor([Wi, Hi])
        wh = torch.exp(torch.rand(40, 2, generator=g) * 5.5) + 1
        b = torch.cat([c - wh / 2, c + wh / 2], 1)
        b[0] = torch.tensor([-20., -10., Wi + 30., Hi + 5.])            # beyond the image
        b[1] = torch.tensor([10., 10., 10., 40.])                       # zero width
        boxes.append(b)
    for S in (7, 14):
        ref = ofpn.roi_pool(feats, boxes, S)
        rois = torch.cat([torch.cat([torch.full((b.shape[0], 1), float(i)), b], 1) for i, b in enumerate(boxes)]).cuda().contiguous()
        nhwc = [feats['p%d' % l].permute(0, 2, 3, 1).contiguous().cuda() for l in (2, 3, 4, 5)]
        out = torch.empty(rois.shape[0], S, S, 256, device='cuda')
        fp = (_lib.c_void_p * 4)(*[t.data_ptr() for t in nhwc])
        check(_lib.lib().pvo_roi_align_fpn(fp, _lib.int_array([t.shape[1] for t in nhwc]), _lib.int_array([t.shape[2] for t in nhwc]), 256,
                                           ptr(rois), rois.shape[0], S, ptr(out), _lib.F32, stream()), 'roi_align_fpn')
        assert (out.permute(0, 3, 1, 2).cpu() - ref).abs().max() < 1e-4 * ref.abs().max()


def paste_masks(masks, boxes, img_h, img_w, threshold=0.5):
    """layers/mask_ops.py:17-152 restated with torch ops (the test's reference): bilinear resample of the 28x28 probabilities into
    the box via F.grid_sample, thresholded -> (R,H,W) bool"""
    R = masks.shape[0]
    x0, y0, x1, y1 = torch.split(boxes, 1, dim=1)
    img_y = (torch.arange(0, img_h, dtype=torch.float32, device=masks.device) + 0.5 - y0) / (y1 - y0) * 2 - 1
    img_x = (torch.arange(0, img_w, dtype=torch.float32, device=masks.device) + 0.5 - x0) / (x1 - x0) * 2 - 1
    grid = torch.stack([img_x[:, None, :].expand(R, img_h, img_w), img_y[:, :, None].expand(R, img_h, img_w)], dim=3)
    return torch.nn.functional.grid_sample(masks.float(), grid, align_corners=False)[:, 0] >= threshold


@pytest.mark.parametrize('w', [200, 203])
def test_paste_masks_matches_grid_sample(w):
    """the paste kernel restates mask_ops.py's grid construction + grid_sample operation for operation: same booleans (4 pixels
    per thread when 4 | W, one otherwise -- there is no torch fallback)"""
    from pvo_b200.panoptic_fpn import PanopticFPN
    m = PanopticFPN.__new__(PanopticFPN)
    m.device = torch.device('cuda')
    g = torch.Generator().manual_seed(11)
    h, R = 120, 64
    c = torch.rand(R, 2, generator=g) * torch.tensor([w, h])
    wh = torch.exp(torch.rand(R, 2, generator=g) * 4.5) + 2
    boxes = torch.cat([c - wh / 2, c + wh / 2], 1)
    boxes[0] = torch.tensor([0., 0., float(w), float(h)])
    boxes[1] = torch.tensor([-30.3, 10.2, 50.7, 300.0])
    boxes[2] = torch.tensor([10., 20., 38., 48.])               # integer box: pixel centres exactly on cell borders
    prob = torch.nn.functional.avg_pool2d(torch.rand(R, 1, 30, 30, generator=g), 3, 1).clamp(0, 1)
    prob[3] = 1.0
    ref = paste_masks(prob.cuda(), boxes.cuda(), h, w)
    out = m.paste(prob.cuda(), boxes.cuda(), h, w)
    bad = (ref != out).sum().item()
    NOTES['paste_mismatch_px'] = [bad, ref.numel()]
    assert bad == 0, bad
    # masks never extend beyond the box (what lets the merge scan only the box region)
    ys, xs = torch.arange(h, device='cuda') + 0.5, torch.arange(w, device='cuda') + 0.5
    b = boxes.cuda()
    inside = ((ys[None, :, None] >= b[:, 1, None, None] - 1) & (ys[None, :, None] <= b[:, 3, None, None] + 1) &
              (xs[None, None, :] >= b[:, 0, None, None] - 1) & (xs[None, None, :] <= b[:, 2, None, None] + 1))
    assert not (ref & ~inside).any()


def test_panoptic_merge_bit_exact(setup):
    sd, images, ref, dbg, m = setup
    per_image, sems = [], []
    for r in ref:
        inst = r['instances']
        per_image.append((inst['scores'].cuda(), inst['pred_classes'].cuda(), inst['pred_masks'].cuda()))
        sems.append(r['sem_seg'].argmax(0))
    for boxes in (None, [r['instances']['pred_boxes'].cuda() for r in ref]):     # full-image scan / box-limited scan
        pan, segs = m.combine(per_image, torch.stack(sems).cuda(), boxes=boxes)
        for n, r in enumerate(ref):
            assert torch.equal(pan[n].cpu(), r['panoptic_seg'][0])
            assert [s['id'] for s in segs[n]] == [s['id'] for s in r['panoptic_seg'][1]]
            assert [s.get('instance_id') for s in segs[n]] == [s.get('instance_id') for s in r['panoptic_seg'][1]]


def test_panoptic_merge_synthetic():
    """overlap suppression, partial overlap trimming, score threshold, stuff area threshold -- vs the oracle's loop"""
    from pvo_b200.panoptic_fpn import PanopticFPN
    m = PanopticFPN.__new__(PanopticFPN)
    m.device = torch.device('cuda')
    h, w = 96, 160
    masks = torch.zeros(5, h, w, dtype=torch.bool)
    masks[0, 10:50, 10:60] = True
    masks[1, 12:48, 12:58] = True          # inside #0 -> dropped (overlap > 0.5)
    masks[2, 40:90, 50:120] = True         # partial overlap with #0 -> trimmed
    masks[3, 0:5, 0:5] = True              # score below 0.5 -> stops the loop
    scores = torch.tensor([0.9, 0.8, 0.7, 0.3, 0.95])     # #4 empty mask -> skipped
    classes = torch.tensor([0, 1, 2, 0, 1])
    sem = torch.zeros(h, w, dtype=torch.int64)
    sem[:, 80:] = 3
    sem[60:, :30] = 7                      # 36*30 = 1080 < 4096 -> dropped
    pan_ref, segs_ref = ofpn.combine_semantic_and_instance((None, scores, classes, masks), sem)
    pan, segs = m.combine([(scores.cuda(), classes.cuda(), masks.cuda())], sem[None].cuda())
    assert torch.equal(pan[0].cpu(), pan_ref)
    assert [s['id'] for s in segs[0]] == [s['id'] for s in segs_ref]
    assert [s.get('area') for s in segs[0] if not s['isthing']] == [s.get('area') for s in segs_ref if not s['isthing']]


def test_inference_end_to_end(setup):
    sd, images, ref, dbg, m = setup
    out = m.inference([im.cuda() for im in images])
    assert len(out) == 2
    for o, r in zip(out, ref):
        assert o['sem_seg'].shape == r['sem_seg'].shape
        pan, segs = o['panoptic_seg']
        assert pan.shape == (H, W) and pan.dtype == torch.int32
        # class map agreement (ids // 10000 is the category; instance numbering may differ when fp16 flips an NMS tie)
        agree = ((pan.cpu() // 10000) == (r['panoptic_seg'][0] // 10000)).float().mean().item()
        NOTES.setdefault('panoptic_class_agreement', []).append(agree)
        NOTES.setdefault('detections', []).append([o['instances']['scores'].shape[0], r['instances']['scores'].shape[0]])
        assert agree > 0.9, agree
        assert (o['sem_seg'].argmax(0).cpu() == r['sem_seg'].argmax(0)).float().mean() > 0.97


def test_config4_resolution_480x640():
    """BASELINE configs[3] resolution (one 480x640 image; the bench runs 8): feature pyramid, semantic logits and RPN heads against
    the fp32 oracle at the stage tolerances of the 128x192 tests, then the composed output by pixel agreement"""
    from pvo_b200.panoptic_fpn import PanopticFPN
    sd = ofpn.init_params(0)
    g = torch.Generator().manual_seed(5)
    hh, ww = 480, 640
    img = torch.nn.functional.avg_pool2d(torch.randint(0, 256, (1, 3, hh, ww), generator=g).float(), 7, 1, 3)[0] * 2 - 64
    dbg = {}
    ref = ofpn.inference(sd, [img], debug=dbg)
    m = PanopticFPN(sd, device='cuda')
    x, sizes = m.preprocess([img])
    feats = m.backbone(x)
    for k in ('p2', 'p3', 'p4', 'p5', 'p6'):
        t, h, w = feats[k]
        e = rel(nchw(t, 1, h, w), dbg['feats'][k], 'cfg4_' + k)
        assert e < 2e-2, (k, e)
    sem = m.sem_seg(feats)
    assert rel(sem[..., :12].permute(0, 3, 1, 2), dbg['sem'], 'cfg4_sem_logits') < 2e-2
    assert (sem[..., :12].argmax(-1).cpu() == dbg['sem'].argmax(1)).float().mean() > 0.97
    out = m.inference([img.cuda()])
    pan, segs = out[0]['panoptic_seg']
    assert pan.shape == (hh, ww)
    agree = ((pan.cpu() // 10000) == (ref[0]['panoptic_seg'][0] // 10000)).float().mean().item()
    NOTES['cfg4_panoptic_class_agreement'] = agree
    assert agree > 0.9, agree
