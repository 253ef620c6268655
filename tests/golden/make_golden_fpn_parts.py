"""Generate tests/golden/fpn_parts.npz: the parts of Detectron2's Panoptic-FPN post-processing that CAN be executed here, run
from the reference tree unmodified (build container only; needs /root/reference).

`import detectron2` is impossible (fvcore / iopath / compiled _C are absent), but several of its files depend only on torch /
torchvision / numpy, and single functions of others can be executed from their source text with a small namespace:
    layers/mask_ops.py             (whole file)   paste_masks_in_image                  -> oracle.panoptic_fpn.paste_masks
    structures/boxes.py, instances.py, layers/nms.py, layers/wrappers.py (whole files)  Boxes, Instances, batched_nms, cat
    modeling/box_regression.py     (whole file, fvcore losses stubbed: unused here)     Box2BoxTransform.apply_deltas -> apply_deltas
    proposal_generator/proposal_utils.py (whole file)   find_top_rpn_proposals          -> rpn_proposals
    modeling/poolers.py            (function)     assign_boxes_to_levels                -> assign_levels
    roi_heads/fast_rcnn.py         (function)     fast_rcnn_inference_single_image      -> fast_rcnn_inference_single
    meta_arch/panoptic_fpn.py      (function + its two id maps)  combine_semantic_and_instance_outputs -> combine_semantic_and_instance
Nothing is copied: files are loaded / functions compiled from where they lie; only inputs and outputs are stored."""
import ast
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
D2 = '/root/reference/VPS_Module/detectron2'

from oracle import panoptic_fpn as ofpn   # noqa: E402  (anchor grid only: pinned by Detectron2's own known-answer test)


def load_file(qualname, rel):
    spec = importlib.util.spec_from_file_location(qualname, os.path.join(D2, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[qualname] = mod
    spec.loader.exec_module(mod)
    return mod


def compile_names(rel, names, ns):
    """execute the top-level definitions / assignments called `names` of a reference file inside namespace ns"""
    src = open(os.path.join(D2, rel)).read()
    tree = ast.parse(src)
    keep = []
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            keep.append(node)
        elif isinstance(node, ast.Assign) and any(isinstance(t, ast.Name) and t.id in names for t in node.targets):
            keep.append(node)
    code = compile(ast.Module(body=keep, type_ignores=[]), os.path.join(D2, rel), 'exec')
    exec(code, ns)
    return ns


def main():
    torch.set_grad_enabled(False)
    for pkg in ('detectron2', 'detectron2.layers', 'detectron2.structures', 'detectron2.modeling', 'fvcore', 'fvcore.nn'):
        m = types.ModuleType(pkg)
        m.__path__ = []
        sys.modules[pkg] = m
    sys.modules['fvcore.nn'].giou_loss = sys.modules['fvcore.nn'].smooth_l1_loss = None
    wrappers = load_file('detectron2.layers.wrappers', 'layers/wrappers.py')
    nms = load_file('detectron2.layers.nms', 'layers/nms.py')
    losses = load_file('detectron2.layers.losses', 'layers/losses.py')
    mask_ops = load_file('detectron2.layers.mask_ops', 'layers/mask_ops.py')
    L = sys.modules['detectron2.layers']
    L.cat, L.nonzero_tuple, L.batched_nms = wrappers.cat, wrappers.nonzero_tuple, nms.batched_nms
    L.ciou_loss, L.diou_loss = losses.ciou_loss, losses.diou_loss
    boxes_m = load_file('detectron2.structures.boxes', 'structures/boxes.py')
    inst_m = load_file('detectron2.structures.instances', 'structures/instances.py')
    S = sys.modules['detectron2.structures']
    S.Boxes, S.Instances = boxes_m.Boxes, inst_m.Instances
    boxreg = load_file('detectron2.modeling.box_regression', 'modeling/box_regression.py')
    putils = load_file('detectron2.modeling.proposal_generator.proposal_utils', 'modeling/proposal_generator/proposal_utils.py')
    meta = load_file('builtin_meta_standalone', 'data/datasets/builtin_meta.py')
    Boxes, Instances = boxes_m.Boxes, inst_m.Instances
    g = torch.Generator().manual_seed(77)
    out = {}

    # ---- Box2BoxTransform.apply_deltas
    def rboxes(n, W=320, H=200):
        c = torch.rand(n, 2, generator=g) * torch.tensor([W, H])
        wh = torch.exp(torch.rand(n, 2, generator=g) * 5.0) + 1
        return torch.cat([c - wh / 2, c + wh / 2], 1)
    src = rboxes(64)
    for name, wts, k in (('d1', (1.0, 1.0, 1.0, 1.0), 1), ('d3', (10.0, 10.0, 5.0, 5.0), 3)):
        d = torch.randn(64, 4 * k, generator=g) * torch.tensor([0.5, 0.5, 2.0, 2.0]).repeat(k)
        out[name + '_in'], out[name + '_out'] = d.numpy(), boxreg.Box2BoxTransform(weights=wts).apply_deltas(d, src).numpy()
    out['d_boxes'] = src.numpy()

    # ---- find_top_rpn_proposals (3 levels, 2 images, pre/post top-k small enough to bite)
    sizes = [(96, 160), (90, 150)]
    props, logits, lg_list, dl_list = [], [], [], []
    for (h, w, stride, size) in ((24, 40, 4, 32), (12, 20, 8, 64), (6, 10, 16, 128)):
        anchors = ofpn.grid_anchors(h, w, stride, size)
        lg = torch.randn(2, 3, h, w, generator=g) * 2
        dl = torch.randn(2, 12, h, w, generator=g) * 0.4
        dflat = dl.view(2, 3, 4, h, w).permute(0, 3, 4, 1, 2).reshape(2, -1, 4)
        dec = torch.stack([boxreg.Box2BoxTransform(weights=(1.0, 1.0, 1.0, 1.0)).apply_deltas(dflat[n], anchors) for n in range(2)])
        props.append(dec)
        logits.append(lg.permute(0, 2, 3, 1).reshape(2, -1))
        lg_list.append(lg)
        dl_list.append(dl)
    res = putils.find_top_rpn_proposals(props, logits, sizes, nms_thresh=0.7, pre_nms_topk=300, post_nms_topk=200, min_box_size=0.0, training=False)
    for li, (lg, dl) in enumerate(zip(lg_list, dl_list)):
        out['rpn_logits%d' % li], out['rpn_deltas%d' % li] = lg.numpy(), dl.numpy()
    for n, r in enumerate(res):
        out['rpn_boxes%d' % n], out['rpn_scores%d' % n] = r.proposal_boxes.tensor.numpy(), r.objectness_logits.numpy()

    # ---- assign_boxes_to_levels
    ns = compile_names('modeling/poolers.py', {'assign_boxes_to_levels'}, {'torch': torch, 'cat': wrappers.cat, 'List': list, 'Boxes': Boxes})
    lb = torch.cat([rboxes(100, 640, 480), torch.tensor([[0., 0., 640., 480.], [5., 5., 5., 9.], [10., 10., 12., 12.]])])
    out['lvl_boxes'] = lb.numpy()
    out['lvl_out'] = ns['assign_boxes_to_levels']([Boxes(lb[:50]), Boxes(lb[50:])], 2, 5, 224, 4).numpy()

    # ---- fast_rcnn_inference_single_image
    ns = compile_names('modeling/roi_heads/fast_rcnn.py', {'fast_rcnn_inference_single_image'},
                       {'torch': torch, 'Boxes': Boxes, 'Instances': Instances, 'batched_nms': nms.batched_nms, 'Tuple': tuple})
    R = 300
    fb = rboxes(R)[:, None, :].repeat(1, 3, 1) + torch.randn(R, 3, 4, generator=g) * 3
    fs = torch.softmax(torch.randn(R, 4, generator=g) * 2, dim=-1)
    fb[7, 1, 2] = float('inf')
    inst, _ = ns['fast_rcnn_inference_single_image'](fb.reshape(R, 12).clone(), fs.clone(), (200, 320), 0.05, 0.5, 100)
    out['det_boxes_in'], out['det_scores_in'] = fb.reshape(R, 12).numpy(), fs.numpy()
    out['det_boxes'], out['det_scores'], out['det_classes'] = inst.pred_boxes.tensor.numpy(), inst.scores.numpy(), inst.pred_classes.numpy()

    # ---- paste_masks_in_image
    pb = rboxes(24, 160, 96)
    pb[0] = torch.tensor([0., 0., 160., 96.])
    pb[1] = torch.tensor([-20.5, 10.25, 70.75, 130.0])
    pm = torch.nn.functional.avg_pool2d(torch.rand(24, 1, 30, 30, generator=g), 3, 1)[:, 0].clamp(0, 1)
    out['paste_boxes'], out['paste_probs'] = pb.numpy(), pm.numpy()
    out['paste_out'] = mask_ops.paste_masks_in_image(pm, pb, (96, 160), threshold=0.5).numpy()

    # ---- combine_semantic_and_instance_outputs (+ the VKITTI id maps it closes over)
    ns = compile_names('modeling/meta_arch/panoptic_fpn.py', {'thing_ids', 'thing_id2cat', 'stuff_ids', 'stuff_id2cat', 'combine_semantic_and_instance_outputs'},
                       {'torch': torch, 'VKITTI_CATEGORIES': meta.VKITTI_CATEGORIES})
    out['thing_id2cat'] = np.array(sorted(ns['thing_id2cat'].items()))
    out['stuff_id2cat'] = np.array(sorted(ns['stuff_id2cat'].items()))
    H, W = 96, 160
    masks = mask_ops.paste_masks_in_image(pm, pb, (H, W), threshold=0.5)
    scores = torch.rand(24, generator=g)
    classes = torch.randint(0, 3, (24,), generator=g)
    sem = torch.zeros(H, W, dtype=torch.int64)
    sem[:, 40:100] = 3
    sem[60:, :] = 7
    sem[:20, 100:] = 11
    inst = types.SimpleNamespace(scores=scores, pred_masks=masks, pred_classes=classes)
    pan, segs = ns['combine_semantic_and_instance_outputs'](inst, sem, 0.5, 1024, 0.5)
    out['cmb_scores'], out['cmb_classes'], out['cmb_sem'], out['cmb_pan'] = scores.numpy(), classes.numpy(), sem.numpy(), pan.numpy()
    out['cmb_seg_ids'] = np.array([s['id'] for s in segs])
    out['cmb_seg_inst'] = np.array([s.get('instance_id', -1) for s in segs])
    np.savez_compressed(os.path.join(HERE, 'fpn_parts.npz'), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
