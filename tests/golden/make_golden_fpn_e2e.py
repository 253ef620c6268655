"""Generate tests/golden/fpn_e2e.npz: the reference's WHOLE Panoptic-FPN inference, run from the reference tree unmodified on
CPU (build container only; needs /root/reference).

make_golden_fpn_net.py already loads the convolutional modules (ResNet, FPN, SemSegFPNHead, StandardRPNHead, the box / mask
heads).  This script goes on from there: it loads, whole and from where they lie, structures/{boxes,image_list,instances,masks}.py,
layers/{nms,roi_align,mask_ops}.py, modeling/{box_regression,matcher,sampling,anchor_generator,poolers,postprocessing}.py,
proposal_generator/{proposal_utils,rpn}.py, roi_heads/{fast_rcnn,roi_heads}.py and meta_arch/rcnn.py, compiles the class
PanopticFPN + combine_semantic_and_instance_outputs + the VKITTI id maps from meta_arch/panoptic_fpn.py's source text (its
module-level imports of cv2 / PIL / the AOT losses are not needed by `inference`), builds the model with the arguments
from_config would pass for configs/COCO-PanopticSegmentation/panoptic_fpn_R_50_3x_vkitti_511.yaml (FUSION off: the
single-image path), strict-loads oracle.panoptic_fpn.init_params() and calls the reference's own
`PanopticFPN.inference(batched_inputs)`: preprocess_image -> backbone -> sem_seg_head -> RPN -> StandardROIHeads ->
sem_seg_postprocess / detector_postprocess -> combine_semantic_and_instance_outputs.

Stubbed third-party names (absent from this image): fvcore.nn (weight init / losses: weights are loaded, losses never run),
pycocotools.mask (polygon masks only), detectron2.utils.{events,memory,logger,env,registry}, detectron2.config.configurable,
detectron2.data.detection_utils (visualisation only).  Only the seeds of the inputs and the outputs are stored."""
import os
import sys
import types

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import make_golden_fpn_net as net          # noqa: E402  (stubs + the convolutional modules)
from make_golden_fpn_parts import compile_names  # noqa: E402  (keeps top-level assignments too)
from oracle import panoptic_fpn as ofpn    # noqa: E402

SIZES = ((3, 160, 224), (3, 150, 200))      # two images of different size: exercises ImageList padding (to 160 x 224)
SEED = 31


def images():
    g = torch.Generator().manual_seed(SEED)
    return [torch.randint(0, 256, s, generator=g).float() for s in SIZES]


def build(fusion=False):
    backbone, sem, rpn_head, bh, pred_unused, mh = net.build()
    stub, load_file, Registry = net.stub, net.load_file, net.Registry
    noop = lambda *a, **k: None
    mod = sys.modules
    L = mod['detectron2.layers']
    wr = mod['detectron2.layers.wrappers']
    nms = load_file('detectron2.layers.nms', 'layers/nms.py')
    ra = load_file('detectron2.layers.roi_align', 'layers/roi_align.py')
    losses = load_file('detectron2.layers.losses', 'layers/losses.py')
    L.batched_nms, L.ROIAlign, L.roi_align, L.ROIAlignRotated = nms.batched_nms, ra.ROIAlign, ra.roi_align, None
    L.shapes_to_tensor, L.ciou_loss, L.diou_loss = wr.shapes_to_tensor, losses.ciou_loss, losses.diou_loss
    mask_ops = load_file('detectron2.layers.mask_ops', 'layers/mask_ops.py')
    L.paste_masks_in_image = mask_ops.paste_masks_in_image
    stub('detectron2.utils.memory', retry_if_cuda_oom=lambda f: f)
    stub('detectron2.utils.logger', log_first_n=noop)
    stub('pycocotools')
    stub('pycocotools.mask')
    S = mod['detectron2.structures']
    boxes = load_file('detectron2.structures.boxes', 'structures/boxes.py')
    il = load_file('detectron2.structures.image_list', 'structures/image_list.py')
    inst = load_file('detectron2.structures.instances', 'structures/instances.py')
    S.Boxes, S.pairwise_iou, S.ImageList, S.Instances, S.RotatedBoxes = boxes.Boxes, boxes.pairwise_iou, il.ImageList, inst.Instances, None
    masks = load_file('detectron2.structures.masks', 'structures/masks.py')
    S.ROIMasks, S.BitMasks, S.PolygonMasks = masks.ROIMasks, masks.BitMasks, masks.PolygonMasks
    S.Keypoints = S.heatmaps_to_keypoints = None
    M = 'detectron2.modeling.'
    boxreg = load_file(M + 'box_regression', 'modeling/box_regression.py')
    matcher = load_file(M + 'matcher', 'modeling/matcher.py')
    load_file(M + 'sampling', 'modeling/sampling.py')
    ag = load_file(M + 'anchor_generator', 'modeling/anchor_generator.py')
    poolers = load_file(M + 'poolers', 'modeling/poolers.py')
    post = load_file(M + 'postprocessing', 'modeling/postprocessing.py')
    stub(M + 'proposal_generator', build_proposal_generator=noop)
    stub(M + 'proposal_generator.build', PROPOSAL_GENERATOR_REGISTRY=Registry('P'))
    load_file(M + 'proposal_generator.proposal_utils', 'modeling/proposal_generator/proposal_utils.py')
    rpn = load_file(M + 'proposal_generator.rpn', 'modeling/proposal_generator/rpn.py')
    fast = load_file(M + 'roi_heads.fast_rcnn', 'modeling/roi_heads/fast_rcnn.py')
    stub(M + 'roi_heads.keypoint_head', build_keypoint_head=noop)
    rh = load_file(M + 'roi_heads.roi_heads', 'modeling/roi_heads/roi_heads.py')
    mod[M + 'roi_heads'].build_roi_heads = noop
    bb = mod[M + 'backbone']
    bb.Backbone, bb.build_backbone = mod[M + 'backbone.backbone'].Backbone, noop
    stub('detectron2.data')
    stub('detectron2.data.detection_utils', convert_image_to_rgb=noop)
    stub(M + 'meta_arch')
    stub(M + 'meta_arch.build', META_ARCH_REGISTRY=Registry('M'))
    rcnn = load_file(M + 'meta_arch.rcnn', 'modeling/meta_arch/rcnn.py')
    meta = load_file('builtin_meta_standalone', 'data/datasets/builtin_meta.py')
    from typing import Dict, List
    ns = dict(torch=torch, nn=nn, F=F, np=np, Dict=Dict, List=List, configurable=net.configurable, META_ARCH_REGISTRY=Registry('M'),
              GeneralizedRCNN=rcnn.GeneralizedRCNN, ImageList=il.ImageList, VKITTI_CATEGORIES=meta.VKITTI_CATEGORIES,
              detector_postprocess=post.detector_postprocess, sem_seg_postprocess=post.sem_seg_postprocess)
    compile_names('modeling/meta_arch/panoptic_fpn.py',
                  {'thing_ids', 'thing_id2cat', 'stuff_ids', 'stuff_id2cat', 'PanopticFPN', 'combine_semantic_and_instance_outputs'}, ns)

    # ---- the arguments from_config passes for the VKITTI config (config/defaults.py line in brackets)
    Sp = mod['detectron2.layers.shape_spec'].ShapeSpec
    anchors = ag.DefaultAnchorGenerator(sizes=[[32], [64], [128], [256], [512]], aspect_ratios=[[0.5, 1.0, 2.0]],
                                        strides=[4, 8, 16, 32, 64], offset=0.0)                                     # [193]
    proposal_generator = rpn.RPN(
        in_features=['p2', 'p3', 'p4', 'p5', 'p6'], head=rpn_head, anchor_generator=anchors,
        anchor_matcher=matcher.Matcher([0.3, 0.7], [0, -1, 1], allow_low_quality_matches=True),
        box2box_transform=boxreg.Box2BoxTransform(weights=(1.0, 1.0, 1.0, 1.0)),                                    # [226]
        batch_size_per_image=256, positive_fraction=0.5, pre_nms_topk=(2000, 1000), post_nms_topk=(1000, 1000),     # Base-RCNN-FPN.yaml
        nms_thresh=0.7, min_box_size=0.0)                                                                           # [243], [165]
    scales = tuple(1.0 / s for s in (4, 8, 16, 32))
    predictor = fast.FastRCNNOutputLayers(Sp(channels=1024), box2box_transform=boxreg.Box2BoxTransform(weights=(10.0, 10.0, 5.0, 5.0)),  # [300]
                                          num_classes=3, test_score_thresh=0.05, test_nms_thresh=0.5, test_topk_per_image=100)  # [278],[281],[598]
    roi_heads = rh.StandardROIHeads(
        box_in_features=['p2', 'p3', 'p4', 'p5'], box_pooler=poolers.ROIPooler(7, scales, 0, 'ROIAlignV2'), box_head=bh, box_predictor=predictor,
        mask_in_features=['p2', 'p3', 'p4', 'p5'], mask_pooler=poolers.ROIPooler(14, scales, 0, 'ROIAlignV2'), mask_head=mh,
        num_classes=3, batch_size_per_image=512, positive_fraction=0.25, proposal_matcher=matcher.Matcher([0.5], [0, 1], allow_low_quality_matches=False))
    model = ns['PanopticFPN'](sem_seg_head=sem, combine_overlap_thresh=0.5, combine_stuff_area_thresh=4096, combine_instances_score_thresh=0.5,
                              fusion=fusion, backbone=backbone, proposal_generator=proposal_generator, roi_heads=roi_heads,
                              pixel_mean=(103.530, 116.280, 123.675), pixel_std=(1.0, 1.0, 1.0), input_format='BGR')
    return model


def pack(res, out, tag=''):
    for n, r in enumerate(res):
        pan, segs = r['panoptic_seg']
        inst = r['instances']
        k = '%s%d' % (tag, n)
        out['pan' + k] = pan.numpy().astype(np.int32)
        out['seg_ids' + k] = np.array([s['id'] for s in segs], dtype=np.int64)
        out['seg_cat' + k] = np.array([s['category_id'] for s in segs], dtype=np.int64)
        out['seg_isthing' + k] = np.array([bool(s['isthing']) for s in segs])
        out['seg_inst' + k] = np.array([s.get('instance_id', -1) for s in segs], dtype=np.int64)
        out['boxes' + k] = inst.pred_boxes.tensor.numpy()
        out['scores' + k] = inst.scores.numpy()
        out['classes' + k] = inst.pred_classes.numpy()
        out['masks' + k] = np.packbits(inst.pred_masks.numpy(), axis=-1)
        out['sem_label' + k] = r['sem_seg'].argmax(dim=0).numpy().astype(np.uint8)
        out['sem_logits' + k] = r['sem_seg'][:, ::4, ::4].numpy()
        print(k, 'detections', len(inst), 'kept (score >= 0.5)', int((inst.scores >= 0.5).sum()), 'segments', len(segs))


FUSION_SIZE, FUSION_SEED = (3, 128, 192), 41


def fusion_inputs():
    """two frames of one 'video' + the forward flow of the first: a smooth field of a few pixels with a patch of negative
    flow (dropped by the reference's uint16 cast) and a patch that leaves the image"""
    g = torch.Generator().manual_seed(FUSION_SEED)
    a = torch.randint(0, 256, FUSION_SIZE, generator=g).float()
    b = torch.randint(0, 256, FUSION_SIZE, generator=g).float()
    _, H, W = FUSION_SIZE
    coarse = torch.rand(1, 2, 5, 7, generator=g) * 6.0
    flow = F.interpolate(coarse, size=(H, W), mode='bilinear', align_corners=True)[0]
    flow[:, 20:50, 30:70] -= 7.5
    flow[0, 80:110, 150:] += 60.0
    return a, b, flow


def main_fusion(out):
    """the video branch: frame 1 goes through the plain path and leaves (image, flow) behind; frame 2 of the same video id
    goes through wrap_in_stage2: backbone on (ref, cur), flow_transport_feature, fusion_conv1, heads"""
    # index assignment with repeated targets (`mask[:, v1, u1] = ...`, :337): serial execution keeps the last source in raster
    # order; with several intra-op threads torch's CPU index_put races at chunk boundaries (single channels keep an earlier
    # source).  The serial result is the one pinned.
    torch.set_num_threads(1)
    model = build(fusion=True)
    sd = ofpn.init_params(0, fusion=True)
    model.load_state_dict(sd, strict=True)
    model.eval()
    a, b, flow = fusion_inputs()
    r1 = model([{'image': a, 'image_id': '0001_00000', 'flow': flow}])
    r2 = model([{'image': b, 'image_id': '0001_00001', 'flow': flow}])
    pack(r1, out, 'f1_')
    pack(r2, out, 'f2_')
    # the transport alone, every level shape of a 128 x 192 frame, random features
    g = torch.Generator().manual_seed(FUSION_SEED + 1)
    for k, (h, w) in {'p2': (32, 48), 'p4': (8, 12), 'p6': (2, 3)}.items():
        feat = torch.randn(2, 16, h, w, generator=g)
        res = model.flow_transport_feature({k: feat.clone()}, flow)[k]
        assert torch.equal(res[0, :16], feat[1])
        out['transport_' + k] = res[0, 16:].numpy()
    torch.set_num_threads(min(16, os.cpu_count() or 1))


def main():
    torch.set_grad_enabled(False)
    model = build()
    sd = ofpn.init_params(0)
    own = model.state_dict()
    extra = sorted(set(own) - set(sd))
    assert extra == ['fusion_conv1.bias', 'fusion_conv1.weight'], extra     # the video-fusion conv: unused by `inference`
    assert not set(sd) - set(own)
    model.load_state_dict({**sd, **{k: own[k] for k in extra}}, strict=True)
    model.eval()
    res = model.inference([{'image': im} for im in images()])
    out = {}
    pack(res, out)
    main_fusion(out)
    np.savez_compressed(os.path.join(HERE, 'fpn_e2e.npz'), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
