"""Plain-torch CPU restatement of PVO's Panoptic-FPN *inference* path -- oracle, test infrastructure only.

Pinning status.  ``import detectron2`` fails in this container (fvcore / iopath / pycocotools / the compiled _C are absent,
SURVEY.md section 8(c)), but every FILE the inference path needs loads from the reference tree once those third-party names are
stubbed, so the reference model itself was run here, three ways (generators under tests/golden/, tests in
tests/test_fpn_oracle_cpu.py):
  WHOLE MODEL (make_golden_fpn_e2e.py -> fpn_e2e.npz): the reference's own PanopticFPN.inference -- GeneralizedRCNN.preprocess_image,
      ImageList padding, ResNet-50 + FPN, SemSegFPNHead, RPN (DefaultAnchorGenerator, find_top_rpn_proposals), StandardROIHeads
      (ROIPooler, FastRCNNOutputLayers.inference, mask head + mask_rcnn_inference), sem_seg_postprocess, detector_postprocess,
      combine_semantic_and_instance_outputs -- built with the VKITTI config's arguments, strict load_state_dict of init_params(),
      two images of different size.  This file's `inference` gives the same 100 detections per image (boxes to 5e-4 px, scores
      to 5e-6), the same segment list, the same panoptic ids on all but 6 of 65 840 pixels and the same masks on all but 18 of
      6.6 M (float32 summation order at the 0.5 threshold).
  PIECES, bit for bit (make_golden_fpn_parts.py -> fpn_parts.npz): apply_deltas, rpn_proposals (top-k / clip / NMS / order),
      assign_levels, fast_rcnn_inference_single, paste_masks, combine_semantic_and_instance and the id maps, on inputs chosen
      to hit ties, clipping, empty classes and the area / overlap thresholds.
  CONVOLUTIONAL HALF at full output (make_golden_fpn_net.py -> fpn_net.npz): features p2..p6, sem-seg logits, RPN logits /
      deltas, box head, mask head, 2e-4 relative; strict load_state_dict pins the parameter inventory to the reference's.
  Detectron2's own known-answer tests: ROIAlign (VPS_Module/tests/layers/test_roi_align.py:14-45), the default anchor grid
      (tests/modeling/test_anchor_generator.py:14-42), the Box2Box delta round trip (tests/modeling/test_box2box_transform.py:17-31).
  VIDEO-FUSION BRANCH (same generator, model built with fusion=True; `f1_*`, `f2_*`, `transport_*` in fpn_e2e.npz): the reference's
      inference_fusion on two frames of one video -- `inference_fusion` below gives the same detections / segment list /
      panoptic ids (<= 1 pixel differs) -- and its flow_transport_feature alone on three level shapes with negative,
      out-of-image and colliding flow vectors: `flow_transport` below is bit-for-bit equal.
  Not covered: training; the depth-filtered transport (flow_depth_transport, off in the shipped settings, panoptic_fpn.py:72).

Follows (file:line under VPS_Module/detectron2/):
  modeling/meta_arch/rcnn.py:224-231          preprocess: (x - pixel_mean) / pixel_std, pad to a multiple of 32
  modeling/backbone/resnet.py:100-211,330-359 BottleneckBlock (stride in the first 1x1, defaults.py:489), BasicStem
  layers/batch_norm.py:13-60                  FrozenBatchNorm2d (eps 1e-5): y = (x - mean) * weight / sqrt(var+eps) + bias
  modeling/backbone/fpn.py:113-154,175-187    lateral 1x1, top-down nearest x2 + sum, 3x3 output convs, p6 = max_pool(p5, 1, 2)
  modeling/meta_arch/semantic_seg.py:181-251  SemSegFPNHead: [conv3x3 -> GN(32) -> ReLU (-> bilinear x2)]^k per level, sum,
                                              1x1 predictor, bilinear x4
  modeling/anchor_generator.py                sizes 32..512, ratios (0.5,1,2), offset 0
  modeling/proposal_generator/rpn.py:67-180   StandardRPNHead;  proposal_utils.py:22-126 find_top_rpn_proposals
  modeling/box_regression.py                  Box2BoxTransform.apply_deltas (scale clamp log(1000/16))
  modeling/poolers.py:22-245                  ROIPooler level assignment (canonical 224 @ level 4), ROIAlignV2
  modeling/roi_heads/box_head.py, fast_rcnn.py:45-170   2 x FC1024, class-specific boxes, score 0.05, NMS 0.5, top 100
  modeling/roi_heads/mask_head.py:114-152,209-285       4 x conv3x3, deconv 2x2, 1x1; per-class sigmoid
  modeling/postprocessing.py, layers/mask_ops.py:17-152 detector_postprocess / paste_masks_in_image / sem_seg_postprocess
  modeling/meta_arch/panoptic_fpn.py:453-494,497-588    inference + combine_semantic_and_instance_outputs (VKITTI id maps)
Config: configs/COCO-PanopticSegmentation/panoptic_fpn_R_50_3x_vkitti_511.yaml (3 thing classes, 12 sem-seg classes).
"""
import math

import torch
import torch.nn.functional as F
import torchvision

PIXEL_MEAN = (103.530, 116.280, 123.675)       # defaults.py:42-46 (BGR), std 1
NUM_THING, NUM_SEM = 3, 12
THING_ID2CAT = {0: 12, 1: 13, 2: 14}            # builtin_meta.py:250-265 trainIds of Truck / Car / Van
STUFF_ID2CAT = {i: i for i in range(1, 12)}     # sem-seg label i (1..11) -> trainId i ; label 0 = "thing" placeholder
RES_STAGES = ((3, 64, 256), (4, 128, 512), (6, 256, 1024), (3, 512, 2048))   # blocks, bottleneck, out channels


# ------------------------------------------------------------------------------------------------ parameters
def param_shapes(fusion=False):
    s = {}

    def conv(name, co, ci, k, bias=True, bn=False):
        s[name + '.weight'] = (co, ci, k, k)
        if bias:
            s[name + '.bias'] = (co,)
        if bn:
            for t in ('weight', 'bias', 'running_mean', 'running_var'):
                s['%s.norm.%s' % (name, t)] = (co,)
    conv('backbone.bottom_up.stem.conv1', 64, 3, 7, bias=False, bn=True)
    cin = 64
    for si, (nb, mid, cout) in enumerate(RES_STAGES):
        for b in range(nb):
            pre = 'backbone.bottom_up.res%d.%d' % (si + 2, b)
            if b == 0:
                conv(pre + '.shortcut', cout, cin, 1, bias=False, bn=True)
            conv(pre + '.conv1', mid, cin, 1, bias=False, bn=True)
            conv(pre + '.conv2', mid, mid, 3, bias=False, bn=True)
            conv(pre + '.conv3', cout, mid, 1, bias=False, bn=True)
            cin = cout
    for lvl, ci in zip((2, 3, 4, 5), (256, 512, 1024, 2048)):
        conv('backbone.fpn_lateral%d' % lvl, 256, ci, 1)
        conv('backbone.fpn_output%d' % lvl, 256, 256, 3)
    for lvl, stride in zip((2, 3, 4, 5), (4, 8, 16, 32)):
        k = max(1, int(math.log2(stride) - 2))
        for j in range(k):
            name = 'sem_seg_head.p%d.%d' % (lvl, 2 * j if stride != 4 else j)
            s[name + '.weight'] = (128, 256 if j == 0 else 128, 3, 3)
            s[name + '.norm.weight'] = (128,)
            s[name + '.norm.bias'] = (128,)
    conv('sem_seg_head.predictor', NUM_SEM, 128, 1)
    conv('proposal_generator.rpn_head.conv', 256, 256, 3)
    conv('proposal_generator.rpn_head.objectness_logits', 3, 256, 1)
    conv('proposal_generator.rpn_head.anchor_deltas', 12, 256, 1)
    s['roi_heads.box_head.fc1.weight'] = (1024, 256 * 7 * 7)
    s['roi_heads.box_head.fc1.bias'] = (1024,)
    s['roi_heads.box_head.fc2.weight'] = (1024, 1024)
    s['roi_heads.box_head.fc2.bias'] = (1024,)
    s['roi_heads.box_predictor.cls_score.weight'] = (NUM_THING + 1, 1024)
    s['roi_heads.box_predictor.cls_score.bias'] = (NUM_THING + 1,)
    s['roi_heads.box_predictor.bbox_pred.weight'] = (NUM_THING * 4, 1024)
    s['roi_heads.box_predictor.bbox_pred.bias'] = (NUM_THING * 4,)
    for j in range(1, 5):
        conv('roi_heads.mask_head.mask_fcn%d' % j, 256, 256, 3)
    s['roi_heads.mask_head.deconv.weight'] = (256, 256, 2, 2)        # ConvTranspose2d: (in, out, kh, kw)
    s['roi_heads.mask_head.deconv.bias'] = (256,)
    conv('roi_heads.mask_head.predictor', NUM_THING, 256, 1)
    if fusion:                                                       # panoptic_fpn.py:81: the video-fusion conv (last: the other
        conv('fusion_conv1', 256, 512, 3)                            # seeded parameters do not move)
    return s


def init_params(seed=0, fusion=False):
    """Seeded random weights with per-layer gains chosen so that activations stay O(1) through all ~55 layers (raw pixel
    input has std ~60: the stem is scaled down by 64; bottleneck tails and un-rectified convs use smaller gains so the
    residual stream does not grow) -- the fp16 path must not overflow where a trained checkpoint would not -- and so that
    the RPN / box / mask heads give non-degenerate proposals, detections and masks.  FrozenBN statistics: mean 0, var 1,
    gamma 1, beta 0, each perturbed by N(0, 0.1)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in param_shapes(fusion).items():
        if name.endswith('running_var'):
            sd[name] = (1.0 + 0.1 * torch.randn(shape, generator=g)).clamp(min=0.5)
        elif '.norm.' in name and name.endswith('weight'):
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif '.norm.' in name:
            sd[name] = 0.1 * torch.randn(shape, generator=g)
        elif name.endswith('.weight'):
            fan = shape[1] * (shape[2] * shape[3] if len(shape) == 4 else 1)
            gain = 2.0 ** 0.5
            if 'stem' in name:
                gain = 2.0 ** 0.5 / 64.0
            elif 'bottom_up' in name and 'conv3' in name:
                gain = 0.5
            elif 'shortcut' in name or 'fpn_' in name or 'fusion' in name:
                gain = 1.0
            elif 'objectness' in name or 'predictor.weight' in name:
                gain = 2.0
            elif 'cls_score' in name:
                gain = 0.4
            elif 'anchor_deltas' in name or 'bbox_pred' in name:
                gain = 0.3
            sd[name] = torch.randn(shape, generator=g) * gain / fan ** 0.5
        else:
            sd[name] = 0.05 * torch.randn(shape, generator=g)
    return sd


# ------------------------------------------------------------------------------------------------ building blocks
def fold_bn(sd, name, eps=1e-5):
    """FrozenBatchNorm2d folded into the preceding bias-free conv: w' = w * s, b' = beta - mean * s."""
    s = sd[name + '.norm.weight'] * (sd[name + '.norm.running_var'] + eps).rsqrt()
    return sd[name + '.weight'] * s[:, None, None, None], sd[name + '.norm.bias'] - sd[name + '.norm.running_mean'] * s


def conv_bn(sd, name, x, stride=1, pad=0):
    w, b = fold_bn(sd, name)
    return F.conv2d(x, w, b, stride=stride, padding=pad)


def conv(sd, name, x, pad=0):
    return F.conv2d(x, sd[name + '.weight'], sd[name + '.bias'], padding=pad)


def preprocess(images):
    """list of (3,H,W) BGR float/uint8 -> (N,3,H',W') normalised, zero padded to a multiple of 32; image sizes."""
    sizes = [tuple(im.shape[-2:]) for im in images]
    H = (max(s[0] for s in sizes) + 31) // 32 * 32
    W = (max(s[1] for s in sizes) + 31) // 32 * 32
    mean = torch.tensor(PIXEL_MEAN).view(3, 1, 1)
    out = torch.zeros(len(images), 3, H, W)
    for i, im in enumerate(images):
        out[i, :, :sizes[i][0], :sizes[i][1]] = im.float() - mean
    return out, sizes


def backbone(sd, x):
    """-> dict p2..p6"""
    x = F.relu(conv_bn(sd, 'backbone.bottom_up.stem.conv1', x, stride=2, pad=3))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    feats = []
    for si, (nb, mid, cout) in enumerate(RES_STAGES):
        for b in range(nb):
            pre = 'backbone.bottom_up.res%d.%d' % (si + 2, b)
            stride = 2 if (b == 0 and si > 0) else 1
            sc = conv_bn(sd, pre + '.shortcut', x, stride=stride) if b == 0 else x
            y = F.relu(conv_bn(sd, pre + '.conv1', x, stride=stride))          # STRIDE_IN_1X1 = True
            y = F.relu(conv_bn(sd, pre + '.conv2', y, pad=1))
            y = conv_bn(sd, pre + '.conv3', y)
            x = F.relu(y + sc)
        feats.append(x)
    out = {}
    prev = None
    for lvl, f in zip((5, 4, 3, 2), feats[::-1]):
        lat = conv(sd, 'backbone.fpn_lateral%d' % lvl, f)
        if prev is not None:
            lat = lat + F.interpolate(prev, scale_factor=2.0, mode='nearest')
        prev = lat
        out['p%d' % lvl] = conv(sd, 'backbone.fpn_output%d' % lvl, lat, pad=1)
    out['p6'] = F.max_pool2d(out['p5'], kernel_size=1, stride=2, padding=0)
    return out


def sem_seg_head(sd, feats):
    """-> logits (N,12,H,W) at input resolution"""
    total = None
    for lvl, stride in zip((2, 3, 4, 5), (4, 8, 16, 32)):
        x = feats['p%d' % lvl]
        k = max(1, int(math.log2(stride) - 2))
        for j in range(k):
            name = 'sem_seg_head.p%d.%d' % (lvl, 2 * j if stride != 4 else j)
            x = F.conv2d(x, sd[name + '.weight'], None, padding=1)
            x = F.relu(F.group_norm(x, 32, sd[name + '.norm.weight'], sd[name + '.norm.bias'], eps=1e-5))
            if stride != 4:
                x = F.interpolate(x, scale_factor=2.0, mode='bilinear', align_corners=False)
        total = x if total is None else total + x
    x = conv(sd, 'sem_seg_head.predictor', total)
    return F.interpolate(x, scale_factor=4.0, mode='bilinear', align_corners=False)


# ------------------------------------------------------------------------------------------------ RPN
def cell_anchors(size):
    """anchor_generator.py generate_cell_anchors: area = size^2, w = sqrt(area / ratio), h = ratio * w, centred at 0."""
    out = []
    for ratio in (0.5, 1.0, 2.0):
        w = math.sqrt(size * size / ratio)
        h = ratio * w
        out.append([-w / 2.0, -h / 2.0, w / 2.0, h / 2.0])
    return torch.tensor(out)


def grid_anchors(h, w, stride, size):
    sx = torch.arange(0, w * stride, stride, dtype=torch.float32)
    sy = torch.arange(0, h * stride, stride, dtype=torch.float32)
    yy, xx = torch.meshgrid(sy, sx, indexing='ij')
    shifts = torch.stack([xx.reshape(-1), yy.reshape(-1), xx.reshape(-1), yy.reshape(-1)], dim=1)
    return (shifts.view(-1, 1, 4) + cell_anchors(size).view(1, -1, 4)).reshape(-1, 4)


SCALE_CLAMP = math.log(1000.0 / 16)


def apply_deltas(deltas, boxes, weights):
    """box_regression.py Box2BoxTransform.apply_deltas; deltas (N, k*4), boxes (N,4)"""
    boxes = boxes.to(deltas.dtype)
    widths = boxes[:, 2] - boxes[:, 0]
    heights = boxes[:, 3] - boxes[:, 1]
    ctr_x = boxes[:, 0] + 0.5 * widths
    ctr_y = boxes[:, 1] + 0.5 * heights
    wx, wy, ww, wh = weights
    dx = deltas[:, 0::4] / wx
    dy = deltas[:, 1::4] / wy
    dw = (deltas[:, 2::4] / ww).clamp(max=SCALE_CLAMP)
    dh = (deltas[:, 3::4] / wh).clamp(max=SCALE_CLAMP)
    pcx = dx * widths[:, None] + ctr_x[:, None]
    pcy = dy * heights[:, None] + ctr_y[:, None]
    pw = torch.exp(dw) * widths[:, None]
    ph = torch.exp(dh) * heights[:, None]
    return torch.stack([pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph], dim=-1).reshape(deltas.shape)


def clip_boxes(b, size):
    h, w = size
    b = b.clone()
    b[..., 0::2] = b[..., 0::2].clamp(min=0, max=w)
    b[..., 1::2] = b[..., 1::2].clamp(min=0, max=h)
    return b


def rpn_head(sd, feats):
    logits, deltas = [], []
    for lvl in (2, 3, 4, 5, 6):
        t = F.relu(conv(sd, 'proposal_generator.rpn_head.conv', feats['p%d' % lvl], pad=1))
        logits.append(conv(sd, 'proposal_generator.rpn_head.objectness_logits', t))
        deltas.append(conv(sd, 'proposal_generator.rpn_head.anchor_deltas', t))
    return logits, deltas


def rpn_proposals(logits, deltas, image_sizes, pre_nms_topk=1000, post_nms_topk=1000, nms_thresh=0.7):
    """proposal_utils.py:22-126 (find_top_rpn_proposals).  Returns per image (boxes (R,4), objectness logits (R,))."""
    N = logits[0].shape[0]
    per_level = []
    for li, (lg, dl, stride, size) in enumerate(zip(logits, deltas, (4, 8, 16, 32, 64), (32, 64, 128, 256, 512))):
        _, A, h, w = lg.shape
        anchors = grid_anchors(h, w, stride, size)                       # (h*w*A, 4), A fastest
        lg = lg.permute(0, 2, 3, 1).reshape(N, -1)
        dl = dl.view(N, A, 4, h, w).permute(0, 3, 4, 1, 2).reshape(N, -1, 4)
        k = min(pre_nms_topk, lg.shape[1])
        top, idx = lg.topk(k, dim=1)
        props = torch.stack([apply_deltas(dl[n, idx[n]], anchors[idx[n]], (1.0, 1.0, 1.0, 1.0)) for n in range(N)])
        per_level.append((props, top, torch.full((k,), li, dtype=torch.int64)))
    out = []
    for n in range(N):
        boxes = torch.cat([p[0][n] for p in per_level])
        scores = torch.cat([p[1][n] for p in per_level])
        lvl = torch.cat([p[2] for p in per_level])
        valid = torch.isfinite(boxes).all(dim=1) & torch.isfinite(scores)
        boxes, scores, lvl = boxes[valid], scores[valid], lvl[valid]
        boxes = clip_boxes(boxes, image_sizes[n])
        keep = ((boxes[:, 2] - boxes[:, 0]) > 0) & ((boxes[:, 3] - boxes[:, 1]) > 0)       # min_box_size = 0
        boxes, scores, lvl = boxes[keep], scores[keep], lvl[keep]
        keep = torchvision.ops.batched_nms(boxes.float(), scores.float(), lvl, nms_thresh)[:post_nms_topk]
        out.append((boxes[keep], scores[keep]))
    return out


# ------------------------------------------------------------------------------------------------ ROI heads
def assign_levels(boxes, min_level=2, max_level=5, canonical_size=224, canonical_level=4):
    sizes = torch.sqrt(((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])).clamp(min=0))
    lvl = torch.floor(canonical_level + torch.log2(sizes / canonical_size + 1e-8))
    return lvl.clamp(min=min_level, max=max_level).to(torch.int64) - min_level


def roi_pool(feats, boxes_per_image, out_size):
    """poolers.py ROIPooler with ROIAlignV2 (aligned=True, sampling_ratio=0) over p2..p5."""
    rois = torch.cat([torch.cat([torch.full((b.shape[0], 1), i, dtype=b.dtype), b], dim=1) for i, b in enumerate(boxes_per_image)])
    lv = assign_levels(rois[:, 1:])
    C = feats['p2'].shape[1]
    out = torch.zeros(rois.shape[0], C, out_size, out_size, dtype=feats['p2'].dtype)
    for l, (name, scale) in enumerate(zip(('p2', 'p3', 'p4', 'p5'), (1 / 4, 1 / 8, 1 / 16, 1 / 32))):
        idx = (lv == l).nonzero().reshape(-1)
        if idx.numel():
            out[idx] = torchvision.ops.roi_align(feats[name], rois[idx], (out_size, out_size), spatial_scale=scale,
                                                 sampling_ratio=0, aligned=True)
    return out


def box_head(sd, x):
    x = x.flatten(1)
    x = F.relu(F.linear(x, sd['roi_heads.box_head.fc1.weight'], sd['roi_heads.box_head.fc1.bias']))
    x = F.relu(F.linear(x, sd['roi_heads.box_head.fc2.weight'], sd['roi_heads.box_head.fc2.bias']))
    scores = F.linear(x, sd['roi_heads.box_predictor.cls_score.weight'], sd['roi_heads.box_predictor.cls_score.bias'])
    deltas = F.linear(x, sd['roi_heads.box_predictor.bbox_pred.weight'], sd['roi_heads.box_predictor.bbox_pred.bias'])
    return scores, deltas


def fast_rcnn_inference_single(boxes, scores, image_size, score_thresh=0.05, nms_thresh=0.5, topk=100):
    """fast_rcnn.py:117-170.  boxes (R, K*4) decoded, scores (R, K+1) softmaxed."""
    valid = torch.isfinite(boxes).all(dim=1) & torch.isfinite(scores).all(dim=1)
    boxes, scores = boxes[valid], scores[valid]
    scores = scores[:, :-1]
    K = scores.shape[1]
    boxes = clip_boxes(boxes.view(-1, K, 4).reshape(-1, 4), image_size).view(-1, K, 4)
    mask = scores > score_thresh
    inds = mask.nonzero()
    b = boxes[mask]
    s = scores[mask]
    keep = torchvision.ops.batched_nms(b.float(), s.float(), inds[:, 1], nms_thresh)[:topk]
    return b[keep], s[keep], inds[keep, 1]


def mask_head(sd, x):
    for j in range(1, 5):
        x = F.relu(conv(sd, 'roi_heads.mask_head.mask_fcn%d' % j, x, pad=1))
    x = F.relu(F.conv_transpose2d(x, sd['roi_heads.mask_head.deconv.weight'], sd['roi_heads.mask_head.deconv.bias'], stride=2))
    return conv(sd, 'roi_heads.mask_head.predictor', x)


def paste_masks(masks, boxes, img_h, img_w, threshold=0.5):
    """layers/mask_ops.py:17-152 (_do_paste_mask over the whole image + threshold); masks (R,1,28,28) probabilities."""
    R = masks.shape[0]
    if R == 0:
        return torch.zeros(0, img_h, img_w, dtype=torch.bool)
    x0, y0, x1, y1 = torch.split(boxes, 1, dim=1)
    img_y = torch.arange(0, img_h, dtype=torch.float32) + 0.5
    img_x = torch.arange(0, img_w, dtype=torch.float32) + 0.5
    img_y = (img_y - y0) / (y1 - y0) * 2 - 1
    img_x = (img_x - x0) / (x1 - x0) * 2 - 1
    gx = img_x[:, None, :].expand(R, img_h, img_w)
    gy = img_y[:, :, None].expand(R, img_h, img_w)
    grid = torch.stack([gx, gy], dim=3)
    out = F.grid_sample(masks.float(), grid, align_corners=False)
    return out[:, 0] >= threshold


def combine_semantic_and_instance(boxes_scores_classes_masks, semantic, overlap_threshold=0.5, stuff_area_thresh=4096,
                                  instances_score_thresh=0.5):
    """panoptic_fpn.py:497-588 with the VKITTI id maps: things -> trainId*10000 + running id, stuff -> trainId*10000."""
    _, scores, classes, masks = boxes_scores_classes_masks
    pan = torch.zeros_like(semantic, dtype=torch.int32)
    segs = []
    cur = 0
    for i in torch.argsort(-scores).tolist():
        sc = scores[i].item()
        if sc < instances_score_thresh:
            break
        mask = masks[i]
        area = int(mask.sum())
        if area == 0:
            continue
        inter = mask & (pan > 0)
        ia = int(inter.sum())
        if ia * 1.0 / area > overlap_threshold:
            continue
        if ia > 0:
            mask = mask & (pan == 0)
        cur += 1
        cat = THING_ID2CAT[int(classes[i])]
        pan[mask] = cat * 10000 + cur
        segs.append({'id': cat * 10000 + cur, 'isthing': True, 'score': sc, 'category_id': cat, 'instance_id': i})
    for lab in torch.unique(semantic).tolist():
        if lab == 0:
            continue
        mask = (semantic == lab) & (pan == 0)
        area = int(mask.sum())
        if area < stuff_area_thresh:
            continue
        cur += 1
        cat = STUFF_ID2CAT[lab]
        pan[mask] = cat * 10000
        segs.append({'id': cat * 10000, 'isthing': False, 'category_id': cat, 'area': area})
    return pan, segs


# ------------------------------------------------------------------------------------------------ whole model
def _roi_heads_and_merge(sd, feats, sem, logits, deltas, sizes, debug=None):
    """RPN proposals -> StandardROIHeads -> post-processing -> merge, for features of len(sizes) images"""
    props = rpn_proposals(logits, deltas, sizes)
    pooled = roi_pool(feats, [p[0] for p in props], 7)
    cls, reg = box_head(sd, pooled)
    results = []
    o = 0
    dets = []
    for n, (pb, _) in enumerate(props):
        R = pb.shape[0]
        boxes = apply_deltas(reg[o:o + R], pb, (10.0, 10.0, 5.0, 5.0))
        scores = F.softmax(cls[o:o + R], dim=-1)
        dets.append(fast_rcnn_inference_single(boxes, scores, sizes[n]))
        o += R
    mp = roi_pool(feats, [d[0] for d in dets], 14)
    ml = mask_head(sd, mp)
    o = 0
    for n, (b, s, c) in enumerate(dets):
        R = b.shape[0]
        prob = ml[o:o + R][torch.arange(R), c][:, None].sigmoid()
        o += R
        h, w = sizes[n]
        sem_r = sem[n, :, :h, :w]                                # sem_seg_postprocess with output size == image size
        masks = paste_masks(prob, b, h, w)
        pan, segs = combine_semantic_and_instance((b, s, c, masks), sem_r.argmax(dim=0))
        results.append({'sem_seg': sem_r, 'instances': {'pred_boxes': b, 'scores': s, 'pred_classes': c, 'pred_masks': masks},
                        'panoptic_seg': (pan, segs)})
    if debug is not None:
        debug.update(feats=feats, sem=sem, rpn_logits=logits, rpn_deltas=deltas, proposals=props, box_cls=cls, box_reg=reg,
                     detections=dets, mask_logits=ml)
    return results


def inference(sd, images, debug=None):
    """PanopticFPN.inference for a list of (3,H,W) BGR images.  Returns per image dict(sem_seg, instances, panoptic_seg)."""
    x, sizes = preprocess(images)
    feats = backbone(sd, x)
    sem = sem_seg_head(sd, feats)
    logits, deltas = rpn_head(sd, feats)
    return _roi_heads_and_merge(sd, feats, sem, logits, deltas, sizes, debug)


# ------------------------------------------------------------------------------------------------ video fusion branch
def flow_transport(feat_ref, flow):
    """PanopticFPN.flow_transport_feature (panoptic_fpn.py:310-344) for one level: feat_ref (C,h,w) features of the previous
    frame, flow (2,H,W) its forward optical flow (x, y) in full-resolution pixels -> (C,h,w) features carried to where the
    flow points, zero where nothing lands.  Exactly as the reference does it, oddities included: the flow is resized to
    (h,w) (bilinear, align_corners=True, :433-438) but NOT rescaled to feature pixels; it is cast float -> uint16 (:314), i.e.
    truncated toward zero and wrapped mod 2^16, so any flow component <= -1 lands out of range and the pixel is dropped
    (:330-335); several sources landing on one target: the last one in raster order stays (numpy-ordered index assignment,
    :337)."""
    import numpy as np
    C, h, w = feat_ref.shape
    f = F.interpolate(flow[None, :2].float(), size=(h, w), mode='bilinear', align_corners=True)[0]
    f = f.to(torch.int32).numpy().astype(np.int64) & 0xFFFF       # trunc toward zero, wrap to uint16
    v, u = np.mgrid[0:h, 0:w]
    u1, v1 = (u + f[0]).astype(np.int32), (v + f[1]).astype(np.int32)
    ok = (0 <= u1) & (u1 < w) & (0 <= v1) & (v1 < h)
    out = np.zeros((C, h, w), dtype=np.float32)
    src = feat_ref.numpy()
    out[:, v1[ok], u1[ok]] = src[:, v[ok], u[ok]]                  # numpy: repeated targets keep the last assignment
    return torch.from_numpy(out)


def inference_fusion(sd, ref_image, ref_flow, cur_image, debug=None, alpha=1.0):
    """PanopticFPN.inference_fusion for the second and later frames of a video (panoptic_fpn.py:254-308 -> wrap_in_stage2
    :193-240, flow_transport on, depth transport off): backbone on (ref, cur) together, per level cat(cur features,
    alpha * transported ref features) -> fusion_conv1 (512 -> 256, 3x3, :410-414), then the ordinary heads on the fused
    pyramid for the current image.  The first frame of a video runs `inference` (:256-264).  Needs init_params(fusion=True)."""
    x, sizes = preprocess([ref_image, cur_image])
    feats = backbone(sd, x)
    fused = {}
    for k, f in feats.items():
        warped = flow_transport(f[0], ref_flow)
        fused[k] = conv(sd, 'fusion_conv1', torch.cat([f[1], alpha * warped], dim=0)[None], pad=1)
    sem = sem_seg_head(sd, fused)
    logits, deltas = rpn_head(sd, fused)
    if debug is not None:
        debug['pair_feats'] = feats
    return _roi_heads_and_merge(sd, fused, sem, logits, deltas, sizes[1:], debug)
