"""Known-answer tests lifted from Detectron2's own suite (the literal expectations, VPS_Module/tests/...) that pin the
pieces of the Panoptic-FPN oracle that CAN be pinned without importing detectron2, plus a
smoke run of the whole restated inference path."""
import os

import numpy as np
import torch
import torchvision

from oracle import panoptic_fpn as ofpn


def test_roi_align_known_answer():
    """tests/layers/test_roi_align.py:14-45: 5x5 ramp, box [1,1,3,3], 4x4 output, aligned vs legacy."""
    x = torch.arange(25, dtype=torch.float32).reshape(1, 1, 5, 5)
    rois = torch.tensor([[0, 1, 1, 3, 3]], dtype=torch.float32)
    old = torchvision.ops.roi_align(x, rois, (4, 4), 1.0, 0, aligned=False)
    new = torchvision.ops.roi_align(x, rois, (4, 4), 1.0, 0, aligned=True)
    old_results = [[7.5, 8, 8.5, 9], [10, 10.5, 11, 11.5], [12.5, 13, 13.5, 14], [15, 15.5, 16, 16.5]]
    correct_results = [[4.5, 5.0, 5.5, 6.0], [7.0, 7.5, 8.0, 8.5], [9.5, 10.0, 10.5, 11.0], [12.0, 12.5, 13.0, 13.5]]
    assert np.allclose(old.flatten().numpy(), np.asarray(old_results).flatten())
    assert np.allclose(new.flatten().numpy(), np.asarray(correct_results).flatten())


def test_anchor_grid_known_answer():
    """tests/modeling/test_anchor_generator.py:14-42 (sizes [32,64], ratios [.25,1,4], stride 4, 1x2 feature map),
    restated with the oracle's cell_anchors / grid construction."""
    import math
    cell = []
    for size in (32, 64):
        for ratio in (0.25, 1.0, 4.0):
            w = math.sqrt(size * size / ratio)
            h = ratio * w
            cell.append([-w / 2, -h / 2, w / 2, h / 2])
    cell = torch.tensor(cell)
    shifts = torch.tensor([[0., 0., 0., 0.], [4., 0., 4., 0.]])
    anchors = (shifts.view(-1, 1, 4) + cell.view(1, -1, 4)).reshape(-1, 4)
    expected = torch.tensor([[-32.0, -8.0, 32.0, 8.0], [-16.0, -16.0, 16.0, 16.0], [-8.0, -32.0, 8.0, 32.0],
                             [-64.0, -16.0, 64.0, 16.0], [-32.0, -32.0, 32.0, 32.0], [-16.0, -64.0, 16.0, 64.0],
                             [-28.0, -8.0, 36.0, 8.0], [-12.0, -16.0, 20.0, 16.0], [-4.0, -32.0, 12.0, 32.0],
                             [-60.0, -16.0, 68.0, 16.0], [-28.0, -32.0, 36.0, 32.0], [-12.0, -64.0, 20.0, 64.0]])
    assert torch.allclose(anchors, expected)
    # the oracle's own generator, one size per level, same construction
    a = ofpn.grid_anchors(1, 2, 4, 32)
    assert a.shape == (6, 4) and torch.allclose(a[0], torch.tensor([-32 / math.sqrt(2), -8 * math.sqrt(2), 32 / math.sqrt(2), 8 * math.sqrt(2)]), atol=1e-4)
    assert torch.allclose(a[3] - a[0], torch.tensor([4., 0., 4., 0.]))


def test_box2box_round_trip():
    """tests/modeling/test_box2box_transform.py:17-31: apply_deltas(get_deltas(src, dst), src) == dst."""
    g = torch.Generator().manual_seed(0)
    weights = (5.0, 5.0, 10.0, 10.0)

    def random_boxes(n):
        b = torch.rand(n, 4, generator=g) * 100
        b[:, 2:] += b[:, :2] + 1
        return b
    src, dst = random_boxes(10), random_boxes(10)
    sw, sh = src[:, 2] - src[:, 0], src[:, 3] - src[:, 1]
    scx, scy = src[:, 0] + 0.5 * sw, src[:, 1] + 0.5 * sh
    dw, dh = dst[:, 2] - dst[:, 0], dst[:, 3] - dst[:, 1]
    dcx, dcy = dst[:, 0] + 0.5 * dw, dst[:, 1] + 0.5 * dh
    deltas = torch.stack([weights[0] * (dcx - scx) / sw, weights[1] * (dcy - scy) / sh,
                          weights[2] * torch.log(dw / sw), weights[3] * torch.log(dh / sh)], dim=1)
    assert torch.allclose(ofpn.apply_deltas(deltas, src, weights), dst, atol=1e-4)


def test_restated_inference_runs_and_is_consistent():
    sd = ofpn.init_params(0)
    g = torch.Generator().manual_seed(2)
    images = [torch.randint(0, 256, (3, 96, 160), generator=g).float(), torch.randint(0, 256, (3, 80, 128), generator=g).float()]
    dbg = {}
    with torch.no_grad():
        res = ofpn.inference(sd, images, debug=dbg)
    assert dbg['feats']['p2'].shape == (2, 256, 24, 40) and dbg['feats']['p6'].shape == (2, 256, 2, 3)
    assert dbg['sem'].shape == (2, 12, 96, 160)
    for r, im in zip(res, images):
        h, w = im.shape[-2:]
        pan, segs = r['panoptic_seg']
        assert pan.shape == (h, w) and pan.dtype == torch.int32
        ids = set(torch.unique(pan).tolist()) - {0}
        assert ids == set(s['id'] for s in segs)
        for s in segs:
            assert s['category_id'] == s['id'] // 10000
        assert r['instances']['pred_masks'].shape[1:] == (h, w)


# ---- pieces pinned against the reference's OWN code, executed on CPU by tests/golden/make_golden_fpn_parts.py ---------------
def _parts():
    return np.load(os.path.join(os.path.dirname(__file__), 'golden', 'fpn_parts.npz'))


def test_apply_deltas_matches_reference_box2box_transform():
    z = _parts()
    src = torch.from_numpy(z['d_boxes'])
    for name, wts in (('d1', (1.0, 1.0, 1.0, 1.0)), ('d3', (10.0, 10.0, 5.0, 5.0))):
        got = ofpn.apply_deltas(torch.from_numpy(z[name + '_in']), src, wts)
        assert torch.equal(got, torch.from_numpy(z[name + '_out'])), name


def test_rpn_proposals_match_reference_find_top_rpn_proposals():
    z = _parts()
    logits = [torch.from_numpy(z['rpn_logits%d' % i]) for i in range(3)]
    deltas = [torch.from_numpy(z['rpn_deltas%d' % i]) for i in range(3)]
    got = ofpn.rpn_proposals(logits, deltas, [(96, 160), (90, 150)], pre_nms_topk=300, post_nms_topk=200, nms_thresh=0.7)
    for n, (b, s_) in enumerate(got):
        assert torch.equal(s_, torch.from_numpy(z['rpn_scores%d' % n])), n
        assert torch.equal(b, torch.from_numpy(z['rpn_boxes%d' % n])), n


def test_level_assignment_matches_reference_poolers():
    z = _parts()
    assert torch.equal(ofpn.assign_levels(torch.from_numpy(z['lvl_boxes'])), torch.from_numpy(z['lvl_out']))


def test_detections_match_reference_fast_rcnn_inference():
    z = _parts()
    b, s_, c = ofpn.fast_rcnn_inference_single(torch.from_numpy(z['det_boxes_in']), torch.from_numpy(z['det_scores_in']), (200, 320))
    assert torch.equal(b, torch.from_numpy(z['det_boxes']))
    assert torch.equal(s_, torch.from_numpy(z['det_scores']))
    assert torch.equal(c, torch.from_numpy(z['det_classes']))


def test_paste_masks_matches_reference_mask_ops():
    z = _parts()
    got = ofpn.paste_masks(torch.from_numpy(z['paste_probs'])[:, None], torch.from_numpy(z['paste_boxes']), 96, 160)
    assert torch.equal(got, torch.from_numpy(z['paste_out']))


def test_panoptic_merge_matches_reference_combine():
    z = _parts()
    assert dict(z['thing_id2cat'].tolist()) == ofpn.THING_ID2CAT and dict(z['stuff_id2cat'].tolist()) == ofpn.STUFF_ID2CAT
    masks = torch.from_numpy(z['paste_out'])
    pan, segs = ofpn.combine_semantic_and_instance((None, torch.from_numpy(z['cmb_scores']), torch.from_numpy(z['cmb_classes']), masks),
                                                   torch.from_numpy(z['cmb_sem']), stuff_area_thresh=1024)
    assert torch.equal(pan, torch.from_numpy(z['cmb_pan']))
    assert [s['id'] for s in segs] == z['cmb_seg_ids'].tolist()
    assert [s.get('instance_id', -1) for s in segs] == z['cmb_seg_inst'].tolist()


def test_convolutional_half_matches_reference_modules():
    """oracle backbone / FPN / sem-seg head / RPN head / box head / mask head vs the reference's OWN module classes
    (resnet.py, fpn.py, semantic_seg.py, rpn.py, box_head.py, mask_head.py, fast_rcnn.py) loaded from the reference tree and run
    on CPU with the oracle's seeded parameters -- strict state-dict load, so the parameter inventory is pinned too
    (tests/golden/make_golden_fpn_net.py)."""
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'fpn_net.npz'))
    sd = ofpn.init_params(0)
    g = torch.Generator().manual_seed(21)
    x = torch.randn(2, 3, 64, 96, generator=g) * 40
    pooled = torch.randn(6, 256, 7, 7, generator=g)
    mp = torch.randn(4, 256, 14, 14, generator=g)

    def close(a, b, name):
        b = torch.from_numpy(b)
        assert a.shape == b.shape, name
        assert (a - b).abs().max() <= 2e-4 * b.abs().max() + 1e-5, (name, (a - b).abs().max().item(), b.abs().max().item())
    with torch.no_grad():
        feats = ofpn.backbone(sd, x)
        for k in ('p2', 'p3', 'p4', 'p5', 'p6'):
            close(feats[k], z['feat_' + k], k)
        close(ofpn.sem_seg_head(sd, feats)[:, :, ::2, ::2], z['sem'], 'sem')
        lg, dl = ofpn.rpn_head(sd, feats)
        for i in range(5):
            close(lg[i], z['rpn_logits%d' % i], 'rpn_logits%d' % i)
            close(dl[i], z['rpn_deltas%d' % i], 'rpn_deltas%d' % i)
        cls, reg = ofpn.box_head(sd, pooled)
        close(cls, z['box_cls'], 'box_cls')
        close(reg, z['box_reg'], 'box_reg')
        close(ofpn.mask_head(sd, mp), z['mask_logits'], 'mask_logits')


def test_whole_model_matches_reference_panoptic_fpn_inference():
    """the reference's own PanopticFPN.inference (every module class + preprocess_image / ImageList padding / RPN /
    StandardROIHeads / sem_seg_postprocess / detector_postprocess / combine, tests/golden/make_golden_fpn_e2e.py) on two
    images of different size against the restated whole-model inference.  float32 summation order differs between the two
    (folded norms, fused loops), so boxes agree to ~5e-4 px and a handful of mask / id pixels sit on the 0.5 threshold."""
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'fpn_e2e.npz'))
    g = torch.Generator().manual_seed(31)
    ims = [torch.randint(0, 256, s, generator=g).float() for s in ((3, 160, 224), (3, 150, 200))]
    with torch.no_grad():
        res = ofpn.inference(ofpn.init_params(0), ims)
    for n, r in enumerate(res):
        pan, segs = r['panoptic_seg']
        inst = r['instances']
        assert [s['id'] for s in segs] == z['seg_ids%d' % n].tolist()
        assert [s['category_id'] for s in segs] == z['seg_cat%d' % n].tolist()
        assert [bool(s['isthing']) for s in segs] == z['seg_isthing%d' % n].tolist()
        assert [s.get('instance_id', -1) for s in segs] == z['seg_inst%d' % n].tolist()
        assert inst['pred_boxes'].shape[0] == z['boxes%d' % n].shape[0] == 100
        assert (inst['pred_boxes'] - torch.from_numpy(z['boxes%d' % n])).abs().max() < 5e-3
        assert (inst['scores'] - torch.from_numpy(z['scores%d' % n])).abs().max() < 1e-4
        assert np.array_equal(inst['pred_classes'].numpy(), z['classes%d' % n])
        masks = np.unpackbits(z['masks%d' % n], axis=-1)[..., :pan.shape[1]].astype(bool)
        assert (inst['pred_masks'].numpy() != masks).mean() < 2e-5
        assert (pan.numpy() != z['pan%d' % n]).mean() < 1e-3
        assert (r['sem_seg'].argmax(0).numpy() != z['sem_label%d' % n]).mean() < 1e-3
        want = torch.from_numpy(z['sem_logits%d' % n])
        assert (r['sem_seg'][:, ::4, ::4] - want).abs().max() < 1e-4 * want.abs().max()


def _fusion_inputs():
    """the inputs of tests/golden/make_golden_fpn_e2e.py fusion_inputs()"""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(41)
    a = torch.randint(0, 256, (3, 128, 192), generator=g).float()
    b = torch.randint(0, 256, (3, 128, 192), generator=g).float()
    coarse = torch.rand(1, 2, 5, 7, generator=g) * 6.0
    flow = F.interpolate(coarse, size=(128, 192), mode='bilinear', align_corners=True)[0]
    flow[:, 20:50, 30:70] -= 7.5
    flow[0, 80:110, 150:] += 60.0
    return a, b, flow


def test_flow_transport_matches_reference_bit_for_bit():
    """PanopticFPN.flow_transport_feature executed from the reference (numpy uint16 cast, serial index assignment) on three
    level shapes, with negative flow (dropped), flow leaving the image and repeated targets"""
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'fpn_e2e.npz'))
    flow = _fusion_inputs()[2]
    g = torch.Generator().manual_seed(42)
    for k, (h, w) in {'p2': (32, 48), 'p4': (8, 12), 'p6': (2, 3)}.items():
        feat = torch.randn(2, 16, h, w, generator=g)
        assert np.array_equal(ofpn.flow_transport(feat[0], flow).numpy(), z['transport_' + k]), k
    assert (z['transport_p2'] != 0).mean() > 0.5 and (z['transport_p2'] == 0).mean() > 0.2


def test_video_fusion_matches_reference_inference_fusion():
    """two frames of one video through the reference's own PanopticFPN (FUSION = True): frame 1 = single-image path, frame 2 =
    wrap_in_stage2 (pair backbone, flow transport, fusion_conv1, heads)"""
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'fpn_e2e.npz'))
    a, b, flow = _fusion_inputs()
    sd = ofpn.init_params(0, fusion=True)
    with torch.no_grad():
        runs = (('f1_0', ofpn.inference(sd, [a])[0]), ('f2_0', ofpn.inference_fusion(sd, a, flow, b)[0]))
    for tag, r in runs:
        pan, segs = r['panoptic_seg']
        inst = r['instances']
        assert [s['id'] for s in segs] == z['seg_ids' + tag].tolist(), tag
        assert [s.get('instance_id', -1) for s in segs] == z['seg_inst' + tag].tolist(), tag
        assert inst['pred_boxes'].shape[0] == z['boxes' + tag].shape[0]
        assert (inst['pred_boxes'] - torch.from_numpy(z['boxes' + tag])).abs().max() < 5e-3
        assert (inst['scores'] - torch.from_numpy(z['scores' + tag])).abs().max() < 1e-4
        assert np.array_equal(inst['pred_classes'].numpy(), z['classes' + tag])
        masks = np.unpackbits(z['masks' + tag], axis=-1)[..., :pan.shape[1]].astype(bool)
        assert (inst['pred_masks'].numpy() != masks).mean() < 2e-5
        assert (pan.numpy() != z['pan' + tag]).mean() < 1e-3
        want = torch.from_numpy(z['sem_logits' + tag])
        assert (r['sem_seg'][:, ::4, ::4] - want).abs().max() < 1e-4 * want.abs().max()
    assert z['seg_idsf1_0'].tolist() != z['seg_idsf2_0'].tolist()       # the fused frame is a different result, not a replay
