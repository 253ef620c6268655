"""Generate tests/golden/fpn_net.npz: the CONVOLUTIONAL half of the reference's Panoptic-FPN, run from the reference tree
unmodified on CPU (build container only; needs /root/reference).

`import detectron2` is impossible here (fvcore / iopath / compiled _C absent).  This script stubs exactly the missing
third-party names (fvcore.nn.weight_init -- weights are loaded, never initialised --, fvcore.nn.distributed, the registries,
the `configurable` decorator, comm / env) and then loads the reference's own module FILES: layers/{wrappers,batch_norm,
shape_spec,blocks}.py, modeling/backbone/{backbone,resnet,fpn}.py, roi_heads/{box_head,mask_head}.py whole, and the classes
SemSegFPNHead, StandardRPNHead, FastRCNNOutputLayers compiled from their source text.  The modules are built with the VKITTI
config's arguments, `load_state_dict(strict=True)` of oracle.panoptic_fpn.init_params() (so the oracle's parameter inventory
IS the reference's), and run on seeded inputs.  Only inputs' seeds and outputs are stored."""
import ast
import importlib.util
import os
import sys
import types

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
D2 = '/root/reference/VPS_Module/detectron2'

from oracle import panoptic_fpn as ofpn   # noqa: E402


def stub(name, **attrs):
    m = types.ModuleType(name)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load_file(qualname, rel):
    spec = importlib.util.spec_from_file_location(qualname, os.path.join(D2, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[qualname] = mod
    spec.loader.exec_module(mod)
    return mod


def compile_names(rel, names, ns):
    tree = ast.parse(open(os.path.join(D2, rel)).read())
    keep = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    exec(compile(ast.Module(body=keep, type_ignores=[]), os.path.join(D2, rel), 'exec'), ns)
    return ns


class Registry:
    def __init__(self, name):
        self.name = name

    def register(self, obj=None):
        return obj if obj is not None else (lambda o: o)


def configurable(init_func=None, *, from_config=None):
    if init_func is not None:
        return init_func
    return lambda f: f


def build():
    noop = lambda *a, **k: None
    winit = stub('fvcore.nn.weight_init', c2_msra_fill=noop, c2_xavier_fill=noop)
    stub('fvcore')
    stub('fvcore.nn', weight_init=winit, giou_loss=None, smooth_l1_loss=None)
    stub('fvcore.nn.distributed', differentiable_all_reduce=noop)
    stub('detectron2')
    stub('detectron2.utils', comm=types.SimpleNamespace(get_world_size=lambda: 1), env=types.SimpleNamespace(TORCH_VERSION=(2, 0)))
    stub('detectron2.utils.registry', Registry=Registry)
    stub('detectron2.utils.events', get_event_storage=noop)
    stub('detectron2.config', configurable=configurable)
    L = stub('detectron2.layers')
    wr = load_file('detectron2.layers.wrappers', 'layers/wrappers.py')
    bn = load_file('detectron2.layers.batch_norm', 'layers/batch_norm.py')
    ss = load_file('detectron2.layers.shape_spec', 'layers/shape_spec.py')
    for k in ('Conv2d', 'ConvTranspose2d', 'cat', 'BatchNorm2d', 'Linear', 'nonzero_tuple', 'cross_entropy'):
        setattr(L, k, getattr(wr, k))
    L.FrozenBatchNorm2d, L.get_norm, L.ShapeSpec = bn.FrozenBatchNorm2d, bn.get_norm, ss.ShapeSpec
    L.DeformConv = L.ModulatedDeformConv = None
    bl = load_file('detectron2.layers.blocks', 'layers/blocks.py')
    L.CNNBlockBase = bl.CNNBlockBase
    stub('detectron2.structures', Instances=object, ImageList=object, Boxes=object)
    stub('detectron2.modeling')
    stub('detectron2.modeling.backbone')
    stub('detectron2.modeling.backbone.build', BACKBONE_REGISTRY=Registry('BACKBONE'))
    load_file('detectron2.modeling.backbone.backbone', 'modeling/backbone/backbone.py')
    resnet = load_file('detectron2.modeling.backbone.resnet', 'modeling/backbone/resnet.py')
    fpn = load_file('detectron2.modeling.backbone.fpn', 'modeling/backbone/fpn.py')
    stub('detectron2.modeling.roi_heads')
    box_head = load_file('detectron2.modeling.roi_heads.box_head', 'modeling/roi_heads/box_head.py')
    mask_head = load_file('detectron2.modeling.roi_heads.mask_head', 'modeling/roi_heads/mask_head.py')
    from typing import Callable, Dict, List, Optional, Tuple, Union
    ns = dict(nn=nn, F=F, torch=torch, np=np, Conv2d=wr.Conv2d, ShapeSpec=ss.ShapeSpec, get_norm=bn.get_norm, weight_init=winit,
              configurable=configurable, SEM_SEG_HEADS_REGISTRY=Registry('S'), RPN_HEAD_REGISTRY=Registry('R'), Callable=Callable, Dict=Dict,
              List=List, Optional=Optional, Tuple=Tuple, Union=Union, cat=wr.cat, cross_entropy=wr.cross_entropy, nonzero_tuple=wr.nonzero_tuple,
              Box2BoxTransform=object, Boxes=object, Instances=object, batched_nms=None, get_event_storage=noop, _dense_box_regression_loss=None,
              _log_classification_stats=None, fast_rcnn_inference=None)
    compile_names('modeling/meta_arch/semantic_seg.py', {'SemSegFPNHead'}, ns)
    compile_names('modeling/proposal_generator/rpn.py', {'StandardRPNHead'}, ns)
    compile_names('modeling/roi_heads/fast_rcnn.py', {'FastRCNNOutputLayers'}, ns)
    S = ss.ShapeSpec
    bottom_up = resnet.ResNet(resnet.BasicStem(3, 64, norm='FrozenBN'),
                              resnet.ResNet.make_default_stages(50, norm='FrozenBN', stride_in_1x1=True),
                              out_features=['res2', 'res3', 'res4', 'res5'])
    backbone = fpn.FPN(bottom_up, ['res2', 'res3', 'res4', 'res5'], 256, norm='', top_block=fpn.LastLevelMaxPool(), fuse_type='sum')
    shapes = {'p%d' % l: S(channels=256, stride=2 ** l) for l in (2, 3, 4, 5)}
    sem = ns['SemSegFPNHead'](shapes, num_classes=12, conv_dims=128, common_stride=4, norm='GN', ignore_value=255)
    rpn = ns['StandardRPNHead'](in_channels=256, num_anchors=3, box_dim=4)
    bh = box_head.FastRCNNConvFCHead(S(channels=256, height=7, width=7), conv_dims=[], fc_dims=[1024, 1024])
    pred = ns['FastRCNNOutputLayers'](S(channels=1024), box2box_transform=types.SimpleNamespace(weights=(10.0, 10.0, 5.0, 5.0)), num_classes=3)
    mh = mask_head.MaskRCNNConvUpsampleHead(S(channels=256, height=14, width=14), num_classes=3, conv_dims=[256, 256, 256, 256, 256])
    return backbone, sem, rpn, bh, pred, mh


def main():
    torch.set_grad_enabled(False)
    backbone, sem, rpn, bh, pred, mh = build()
    sd = ofpn.init_params(0)

    def load(mod, prefix):
        sub = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
        mod.load_state_dict(sub, strict=True)              # the oracle's parameter inventory must equal the reference module's
        mod.eval()
        return len(sub)
    counts = [load(backbone, 'backbone.'), load(sem, 'sem_seg_head.'), load(rpn, 'proposal_generator.rpn_head.'),
              load(bh, 'roi_heads.box_head.'), load(pred, 'roi_heads.box_predictor.'), load(mh, 'roi_heads.mask_head.')]
    assert sum(counts) == len(sd), (counts, len(sd))
    g = torch.Generator().manual_seed(21)
    x = torch.randn(2, 3, 64, 96, generator=g) * 40
    feats = backbone(x)
    out = {}                                    # inputs are regenerated from the seed by the test
    for k, v in feats.items():
        out['feat_' + k] = v.numpy()
    logits, _ = sem(feats)
    out['sem'] = logits[:, :, ::2, ::2].numpy()   # every other pixel is plenty
    lg, dl = rpn([feats[k] for k in ('p2', 'p3', 'p4', 'p5', 'p6')])
    for i, (a, b) in enumerate(zip(lg, dl)):
        out['rpn_logits%d' % i], out['rpn_deltas%d' % i] = a.numpy(), b.numpy()
    pooled = torch.randn(6, 256, 7, 7, generator=g)
    scores, deltas = pred(bh(pooled))
    out['box_cls'], out['box_reg'] = scores.numpy(), deltas.numpy()
    mp = torch.randn(4, 256, 14, 14, generator=g)
    out['mask_logits'] = mh.layers(mp).numpy()
    np.savez_compressed(os.path.join(HERE, 'fpn_net.npz'), **{k: v.astype(np.float32) for k, v in out.items()})
    print({k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
