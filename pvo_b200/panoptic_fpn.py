"""Panoptic-FPN (R50-FPN) *inference* on the B200 kernels -- the network whose panoptic ids gate the BA weights.

Mirrors PVO's ``PanopticFPN.inference`` (VPS_Module/detectron2/modeling/meta_arch/panoptic_fpn.py:453-494) and
``combine_semantic_and_instance_outputs`` (:497-588) with the VKITTI configuration
(configs/COCO-PanopticSegmentation/panoptic_fpn_R_50_3x_vkitti_511.yaml: 3 thing classes, 12 sem-seg classes).
Parameter names are Detectron2's checkpoint names (``backbone.bottom_up.res2.0.conv1.weight`` ...); FrozenBatchNorm is
folded into the conv weights at load.

Where the time goes and what runs where:
  * every convolution and FC layer (backbone, FPN, sem-seg head, RPN head, box head FCs, mask head incl. the 2x2
    deconvolution as a 1x1 conv onto 4 sub-positions) -> tcgen05 implicit GEMM (csrc/conv_tc.cu), half operands, fp32
    accumulate, bias / ReLU / bottleneck residual fused in the epilogue.  FC layers view the (R, C) matrix as a 16-wide
    image so the same kernel serves them.  The 7x7/3-channel stem runs on the CUDA-core conv.
  * max-pool, strided sampling, top-down nearest+add, GroupNorm+ReLU, bilinear upsampling -> csrc/fpn_ops.cu.
  * ROIAlign over the pyramid and the per-level / per-class NMS -> csrc/roi_ops.cu (pixel-major features, fp32 box
    arithmetic; the reference calls torchvision's NCHW ops, layers/roi_align.py:49-58, layers/nms.py:9-22); top-k, sort
    and the box decode stay torch tensor ops, batched over images.
  * panoptic merge: one CUDA block per image (csrc/fpn_ops.cu), no per-instance host syncs.
"""
import math

import torch
import torch.nn.functional as F

from . import _lib
from . import nn as pnn
from ._lib import check, ptr, stream

PIXEL_MEAN = (103.530, 116.280, 123.675)
NUM_THING, NUM_SEM = 3, 12
THING_CAT0 = 12                                   # builtin_meta.py:250-265: thing trainIds 12, 13, 14
RES_STAGES = ((3, 64, 256), (4, 128, 512), (6, 256, 1024), (3, 512, 2048))
SCALE_CLAMP = math.log(1000.0 / 16)


def _fold(sd, name, eps=1e-5):
    s = sd[name + '.norm.weight'] * (sd[name + '.norm.running_var'] + eps).rsqrt()
    return sd[name + '.weight'] * s[:, None, None, None], sd[name + '.norm.bias'] - sd[name + '.norm.running_mean'] * s


def apply_deltas(deltas, boxes, weights):
    """box_regression.py Box2BoxTransform.apply_deltas"""
    boxes = boxes.to(deltas.dtype)
    widths, heights = boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1]
    cx, cy = boxes[:, 0] + 0.5 * widths, boxes[:, 1] + 0.5 * heights
    wx, wy, ww, wh = weights
    dx, dy = deltas[:, 0::4] / wx, deltas[:, 1::4] / wy
    dw, dh = (deltas[:, 2::4] / ww).clamp(max=SCALE_CLAMP), (deltas[:, 3::4] / wh).clamp(max=SCALE_CLAMP)
    pcx, pcy = dx * widths[:, None] + cx[:, None], dy * heights[:, None] + cy[:, None]
    pw, ph = torch.exp(dw) * widths[:, None], torch.exp(dh) * heights[:, None]
    return torch.stack([pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph], dim=-1).reshape(deltas.shape)


def clip_boxes(b, size):
    h, w = size
    b = b.clone()
    b[..., 0::2] = b[..., 0::2].clamp(min=0, max=w)
    b[..., 1::2] = b[..., 1::2].clamp(min=0, max=h)
    return b


def cell_anchors(size):
    """anchor_generator.py generate_cell_anchors: 3 aspect ratios, area size^2, centred on the cell origin"""
    cell = []
    for ratio in (0.5, 1.0, 2.0):
        aw = math.sqrt(size * size / ratio)
        ah = ratio * aw
        cell.append([-aw / 2.0, -ah / 2.0, aw / 2.0, ah / 2.0])
    return cell


def grid_anchors(h, w, stride, size, device):
    """anchor_generator.py DefaultAnchorGenerator: anchor index fastest"""
    cell = torch.tensor(cell_anchors(size), device=device)
    sx = torch.arange(0, w * stride, stride, dtype=torch.float32, device=device)
    sy = torch.arange(0, h * stride, stride, dtype=torch.float32, device=device)
    yy, xx = torch.meshgrid(sy, sx, indexing='ij')
    shifts = torch.stack([xx.reshape(-1), yy.reshape(-1), xx.reshape(-1), yy.reshape(-1)], dim=1)
    return (shifts.view(-1, 1, 4) + cell.view(1, -1, 4)).reshape(-1, 4)


def segments_for_vo(pan, size=None):
    """In-process hand-off of a panoptic id map to VO_Module (replaces the PNG round trip panoptic_evaluation.py ->
    split_init_segm.py -> evaluation_scripts/test_vo.py:44-50): nearest resize to the VO input size, then nearest 1/8
    down-sampling -> (1, 1, h//8, w//8) int32, what `Droid.track(..., segments=)` / `DepthVideo.segms` take.
    Faithful to the reference as written: its crop `segm[:h1-h1%8, :w1-w1%8]` slices the batch / channel dims of a 4-D
    tensor (a no-op), so the second interpolate maps h1 -> h1//8 rows with scale h1/(h1//8) (= 8 when 8 | h1).
    Nearest interpolation picks source index floor(dst * in/out): the whole chain is two strided gathers on the device."""
    H, W = pan.shape[-2:]
    h1, w1 = size if size is not None else (H, W)

    def nearest(n_out, n_in, dev):          # at::upsample_nearest: min(floor(dst * (float)n_in / n_out), n_in - 1) in fp32
        i = (torch.arange(n_out, device=dev, dtype=torch.float32) * (float(n_in) / float(n_out))).floor().long()
        return i.clamp(max=n_in - 1)
    ys = nearest(h1, H, pan.device)[nearest(h1 // 8, h1, pan.device)]
    xs = nearest(w1, W, pan.device)[nearest(w1 // 8, w1, pan.device)]
    return pan[..., ys[:, None], xs[None, :]].to(torch.int32)[None, None]


def _cat_rows(ts):
    """torch.cat(ts) along dim 0, without the copy when the pieces are consecutive row blocks of one buffer"""
    if len(ts) == 1:
        return ts[0].contiguous()
    t0 = ts[0]
    off, ok = t0.storage_offset(), all(t.is_contiguous() for t in ts)
    for t in ts:
        ok = ok and t.untyped_storage().data_ptr() == t0.untyped_storage().data_ptr() and t.storage_offset() == off and t.shape[1:] == t0.shape[1:]
        off += t.numel()
    if ok:
        rows = sum(t.shape[0] for t in ts)
        return torch.as_strided(t0, (rows,) + tuple(t0.shape[1:]), t0.stride(), t0.storage_offset())
    return torch.cat(ts).contiguous()


def flow_transport_index(flow, h, w):
    """Index form of PanopticFPN.flow_transport_feature (panoptic_fpn.py:310-344) for one pyramid level: flow (2+,H,W)
    forward flow (x, y) of the previous frame -> (h*w,) int64, for every pixel of the level the raster index of the
    previous-frame pixel whose features land on it, -1 where nothing lands.  The reference's rules, oddities included: the
    flow is resized to (h,w) (bilinear, align_corners=True, :433-438) but not rescaled; float -> uint16 cast (:314) =
    truncate toward zero and wrap mod 2^16, so a component <= -1 lands out of range and is dropped (:330-335); when
    several sources land on one target the last in raster order stays (:337, serial index assignment) -- here a
    scatter-max over source indices, deterministic on the GPU.  Tensor ops only: runs on whatever device `flow` is on."""
    dev = flow.device
    f = F.interpolate(flow[None, :2].float(), size=(h, w), mode='bilinear', align_corners=True)[0]
    f = f.to(torch.int32).to(torch.int64) & 0xFFFF
    u1 = torch.arange(w, device=dev)[None, :] + f[0]
    v1 = torch.arange(h, device=dev)[:, None] + f[1]
    n = h * w
    tgt = torch.where((u1 < w) & (v1 < h), v1 * w + u1, torch.full_like(u1, n)).reshape(-1)     # slot n collects the dropped ones
    win = torch.full((n + 1,), -1, dtype=torch.int64, device=dev)
    win.scatter_reduce_(0, tgt, torch.arange(n, device=dev), 'amax', include_self=True)
    return win[:n]


def flow_transport(feat_ref, win):
    """feat_ref (h*w, C) pixel-major features of the previous frame, win from flow_transport_index -> (h*w, C) carried
    features, zero where nothing lands"""
    got = feat_ref[win.clamp(min=0)]
    return torch.where((win >= 0)[:, None], got, torch.zeros_like(got))


class PanopticFPN:
    def __init__(self, state_dict, device='cuda', use_cuda_graph=True):
        self.device = torch.device(device)
        self.use_cuda_graph = use_cuda_graph     # replay the fixed-shape part (backbone, FPN, sem-seg and RPN heads) as one graph
        self._graphs = {}
        sd = {k: v.detach().to(self.device).float() for k, v in state_dict.items()}
        self.cv = {}
        tc = lambda w, b: pnn._Conv(w, b, tc=True, simt=False)
        w, b = _fold(sd, 'backbone.bottom_up.stem.conv1')
        # 7x7/s2 on 3 channels -> 5x1/s1 on the 48-channel space-to-depth, x-expanded image (csrc/fpn_ops.cu stem_expand)
        w2 = torch.zeros(64, 48, 5, 1, device=self.device)
        for ty in range(5):
            for tx in range(4):
                for dy in range(2):
                    for dx in range(2):
                        ky, kx = 2 * ty + dy - 1, 2 * tx + dx - 1
                        if 0 <= ky < 7 and 0 <= kx < 7:
                            j = tx * 12 + dy * 6 + dx * 3
                            w2[:, j:j + 3, ty, 0] = w[:, :, ky, kx]
        self.cv['stem'] = tc(w2, b)
        for si, (nb, mid, cout) in enumerate(RES_STAGES):
            for bi in range(nb):
                pre = 'backbone.bottom_up.res%d.%d' % (si + 2, bi)
                for c in ('conv1', 'conv2', 'conv3') + (('shortcut',) if bi == 0 else ()):
                    self.cv[pre + '.' + c] = tc(*_fold(sd, pre + '.' + c))
        for lvl in (2, 3, 4, 5):
            for n in ('fpn_lateral', 'fpn_output'):
                self.cv['%s%d' % (n, lvl)] = tc(sd['backbone.%s%d.weight' % (n, lvl)], sd['backbone.%s%d.bias' % (n, lvl)])
        self.gn = {}
        for lvl, stride in zip((2, 3, 4, 5), (4, 8, 16, 32)):
            for j in range(max(1, int(math.log2(stride) - 2))):
                name = 'sem_seg_head.p%d.%d' % (lvl, 2 * j if stride != 4 else j)
                self.cv[name] = tc(sd[name + '.weight'], torch.zeros(128, device=self.device))
                self.gn[name] = (sd[name + '.norm.weight'].contiguous(), sd[name + '.norm.bias'].contiguous())
        self.cv['sem_pred'] = tc(sd['sem_seg_head.predictor.weight'], sd['sem_seg_head.predictor.bias'])
        self.cv['rpn_conv'] = tc(sd['proposal_generator.rpn_head.conv.weight'], sd['proposal_generator.rpn_head.conv.bias'])
        self.cv['rpn_out'] = tc(torch.cat([sd['proposal_generator.rpn_head.objectness_logits.weight'],
                                           sd['proposal_generator.rpn_head.anchor_deltas.weight']], 0),
                                torch.cat([sd['proposal_generator.rpn_head.objectness_logits.bias'],
                                           sd['proposal_generator.rpn_head.anchor_deltas.bias']], 0))
        fc1 = sd['roi_heads.box_head.fc1.weight']                                # columns (c, y, x) -> (y, x, c): pixel-major ROIAlign output
        fc1 = fc1.view(fc1.shape[0], 256, 7, 7).permute(0, 2, 3, 1).reshape(fc1.shape[0], -1)
        self.cv['fc1'] = tc(fc1[:, :, None, None].contiguous(), sd['roi_heads.box_head.fc1.bias'])
        self.cv['fc2'] = tc(sd['roi_heads.box_head.fc2.weight'][:, :, None, None], sd['roi_heads.box_head.fc2.bias'])
        self.cv['box_out'] = tc(torch.cat([sd['roi_heads.box_predictor.cls_score.weight'], sd['roi_heads.box_predictor.bbox_pred.weight']], 0)[:, :, None, None],
                                torch.cat([sd['roi_heads.box_predictor.cls_score.bias'], sd['roi_heads.box_predictor.bbox_pred.bias']], 0))
        for j in range(1, 5):
            self.cv['mask_fcn%d' % j] = tc(sd['roi_heads.mask_head.mask_fcn%d.weight' % j], sd['roi_heads.mask_head.mask_fcn%d.bias' % j])
        dw = sd['roi_heads.mask_head.deconv.weight']                             # (in, out, 2, 2) -> rows (dy, dx, out)
        self.cv['deconv'] = tc(dw.permute(2, 3, 1, 0).reshape(4 * 256, 256)[:, :, None, None].contiguous(),
                               sd['roi_heads.mask_head.deconv.bias'].repeat(4))
        self.cv['mask_pred'] = tc(sd['roi_heads.mask_head.predictor.weight'], sd['roi_heads.mask_head.predictor.bias'])
        self.fusion = 'fusion_conv1.weight' in sd          # the video branch's 512 -> 256 conv (panoptic_fpn.py:81, :410-414)
        if self.fusion:
            self.cv['fusion_conv1'] = tc(sd['fusion_conv1.weight'], sd['fusion_conv1.bias'])
        self.alpha = 1.0                                   # panoptic_fpn.py:70
        self.vid = self.ref_flow = self.ref_image = self.ref_id = None
        self.mean = torch.tensor(PIXEL_MEAN, device=self.device).view(1, 1, 1, 3)
        self._scratch = None

    # ------------------------------------------------------------------------------------------ primitives
    def _buf(self, *shape, dtype=torch.float16):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def _conv(self, name, src, N, H, W, relu=False, res=None, cpad=None):
        """TC conv on a pixel-major (N*H*W, Cin) half tensor -> (N*H*W, Cout') with Cout' = cpad or Cout."""
        cv = self.cv[name]
        co = cpad or cv.cout
        out = self._buf(N * H * W, co)
        if co != cv.cout:
            out.zero_()
        pnn.conv2d(cv, [(src, cv.cin, src.shape[1])], N, H, W, _lib.F16, out, co, 0, 1 if relu else 0, aux0=res,
                   aux_stride=res.shape[1] if res is not None else 0, use_tc=True)
        return out

    def _matmul(self, name, x, relu=False, cpad=None):
        """FC layer: x (R, K) half -> (R, Cout').  The (R, K) matrix is presented as a 16-wide image of R/16 rows."""
        R = x.shape[0]
        Rp = (R + 15) // 16 * 16
        if Rp != R:
            x = torch.cat([x, x.new_zeros(Rp - R, x.shape[1])], 0)
        return self._conv(name, x.contiguous(), 1, Rp // 16, 16, relu=relu, cpad=cpad)[:R]

    def _scr(self, n):
        if self._scratch is None or self._scratch.numel() < n:
            self._scratch = torch.empty(n, dtype=torch.float32, device=self.device)
        return self._scratch

    # ------------------------------------------------------------------------------------------ stages
    def preprocess(self, images):
        """list of (3,H,W) BGR -> pixel-major (N, H', W', 3) half, zero padded to a multiple of 32 (rcnn.py:224-231)"""
        sizes = [tuple(im.shape[-2:]) for im in images]
        H = (max(s[0] for s in sizes) + 31) // 32 * 32
        W = (max(s[1] for s in sizes) + 31) // 32 * 32
        x = torch.zeros(len(images), H, W, 3, dtype=torch.float16, device=self.device)
        for i, im in enumerate(images):
            x[i, :sizes[i][0], :sizes[i][1]] = (im.to(self.device).float().permute(1, 2, 0) - self.mean[0]).half()
        return x, sizes

    def backbone(self, x):
        """x (N,H,W,3) half -> {'p2'..'p6': ((N*h*w, 256) half, h, w)}"""
        L = _lib.lib()
        N, H, W, _ = x.shape
        H1, W1 = H // 2, W // 2
        E = self._buf(N * H1 * W1, 64)
        check(L.pvo_stem_expand(ptr(x), ptr(E), N, H, W, stream()), 'stem_expand')
        y = self._conv('stem', E, N, H1, W1, relu=True)
        H2, W2 = (H1 - 1) // 2 + 1, (W1 - 1) // 2 + 1
        t = self._buf(N * H2 * W2, 64)
        check(L.pvo_maxpool3s2(ptr(y), ptr(t), N, H1, W1, 64, _lib.F16, stream()), 'maxpool3s2')
        x, H, W = t, H2, W2
        feats = []
        for si, (nb, mid, cout) in enumerate(RES_STAGES):
            for bi in range(nb):
                pre = 'backbone.bottom_up.res%d.%d' % (si + 2, bi)
                if bi == 0 and si > 0:                       # STRIDE_IN_1X1: the strided 1x1s read every other pixel
                    Hs, Ws = (H + 1) // 2, (W + 1) // 2
                    xs = self._buf(N * Hs * Ws, x.shape[1])
                    check(L.pvo_subsample2(ptr(x), ptr(xs), N, H, W, x.shape[1], _lib.F16, stream()), 'subsample2')
                    x, H, W = xs, Hs, Ws
                sc = self._conv(pre + '.shortcut', x, N, H, W) if bi == 0 else x
                y = self._conv(pre + '.conv1', x, N, H, W, relu=True)
                y = self._conv(pre + '.conv2', y, N, H, W, relu=True)
                x = self._conv(pre + '.conv3', y, N, H, W, relu=False, res=sc)          # relu(conv3 + shortcut)
            feats.append((x, H, W))
        out = {}
        prev = None
        for lvl, (f, h, w) in zip((5, 4, 3, 2), feats[::-1]):
            lat = self._conv('fpn_lateral%d' % lvl, f, N, h, w)
            if prev is not None:
                check(L.pvo_upsample_add(ptr(lat), ptr(prev[0]), N, h, w, 256, prev[1], prev[2], _lib.F16, stream()), 'upsample_add')
            prev = (lat, h, w)
            out['p%d' % lvl] = (self._conv('fpn_output%d' % lvl, lat, N, h, w), h, w)
        p5, h, w = out['p5']
        p6 = self._buf(N * ((h + 1) // 2) * ((w + 1) // 2), 256)
        check(L.pvo_subsample2(ptr(p5), ptr(p6), N, h, w, 256, _lib.F16, stream()), 'subsample2')
        out['p6'] = (p6, (h + 1) // 2, (w + 1) // 2)
        self.N = N
        return out

    def dense(self, x):
        """backbone + FPN + sem-seg head + RPN head: ~75 launches whose shapes depend only on the batch shape, so the
        second call with a given shape captures them into a CUDA graph and later calls replay it (the per-launch host cost,
        tensor-map encoding included, would otherwise exceed the GPU time).  Outputs live in the graph's buffers until the
        next call with the same shape."""
        def run(inp):
            feats = self.backbone(inp)
            return feats, self.sem_seg(feats), self.rpn_head(feats)
        if not self.use_cuda_graph:
            return run(x)
        key = tuple(x.shape)
        e = self._graphs.get(key)
        if e is None:
            self._graphs[key] = 'warm'
            return run(x)
        if e == 'warm':
            xs = x.clone()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = run(xs)
            e = self._graphs[key] = (g, xs, out)
        g, xs, out = e
        xs.copy_(x)
        g.replay()
        self.N = x.shape[0]
        return out

    def sem_seg(self, feats):
        """-> logits (N, H, W, 16) half at input resolution (channels 12..15 are padding)"""
        L = _lib.lib()
        N = self.N
        total = None
        for lvl, stride in zip((2, 3, 4, 5), (4, 8, 16, 32)):
            x, h, w = feats['p%d' % lvl]
            k = max(1, int(math.log2(stride) - 2))
            for j in range(k):
                name = 'sem_seg_head.p%d.%d' % (lvl, 2 * j if stride != 4 else j)
                x = self._conv(name, x, N, h, w)
                g, b = self.gn[name]
                check(L.pvo_group_norm(ptr(x), ptr(x), ptr(g), ptr(b), ptr(self._scr(N * 128 * 130 + 64)), N, h * w, 128, 4, 1e-5, 1,
                                       _lib.F16, stream()), 'group_norm', launches=3)
                if stride != 4:
                    last = j == k - 1
                    up = total if (last and total is not None) else self._buf(N * 4 * h * w, 128)
                    check(L.pvo_upsample_bilinear(ptr(x), ptr(up), N, h, w, 128, 2, 1 if (last and total is not None) else 0,
                                                  _lib.F16, stream()), 'upsample_bilinear')
                    x, h, w = up, 2 * h, 2 * w
            if total is None:
                total = x
        h4, w4 = feats['p2'][1], feats['p2'][2]
        lg = self._conv('sem_pred', total, N, h4, w4, cpad=16)
        out = self._buf(N * 16 * h4 * w4, 16)
        check(L.pvo_upsample_bilinear(ptr(lg), ptr(out), N, h4, w4, 16, 4, 0, _lib.F16, stream()), 'upsample_bilinear')
        return out.view(N, 4 * h4, 4 * w4, 16)

    def rpn_head(self, feats):
        """-> per level the fused head output (N,h,w,16) half: objectness (3 anchors) | anchor deltas (3x4) | pad"""
        N = self.N
        res = []
        for lvl in (2, 3, 4, 5, 6):
            x, h, w = feats['p%d' % lvl]
            t = self._conv('rpn_conv', x, N, h, w, relu=True)
            res.append(self._conv('rpn_out', t, N, h, w, cpad=16).view(N, h, w, 16))
        return res

    def nms_groups(self, boxes, valid, thresh):
        """boxes (G,n,4) float32, each group sorted by descending score; valid (G,n) bool -> keep (G,n) bool"""
        G, n = valid.shape
        boxes = boxes.float().contiguous()
        v = valid.to(torch.uint8).contiguous()
        keep = torch.empty(G, n, dtype=torch.uint8, device=self.device)
        scratch = torch.empty(G * n * 32, dtype=torch.int32, device=self.device)
        check(_lib.lib().pvo_nms_groups(ptr(boxes), ptr(v), G, n, float(thresh), ptr(keep), ptr(scratch), stream()), 'nms_groups', launches=2)
        return keep.bool()

    def _limits(self, image_sizes):
        hw = torch.tensor(image_sizes, dtype=torch.float32, device=self.device)
        return hw[:, [1, 0, 1, 0]]                                     # per image clip limits for (x0, y0, x1, y1)

    def proposals(self, rpn_out, image_sizes, pre_nms_topk=1000, post_nms_topk=1000, nms_thresh=0.7):
        """find_top_rpn_proposals (proposal_utils.py:22-126): per-level top-k, decode, clip, drop empty, per-level NMS,
        best post_nms_topk per image -- batched over images and levels, one device->host read (the counts).
        rpn_out: per level the raw head output (N,h,w,16) half (objectness x3 | deltas 3x4 | pad)."""
        N, P = self.N, pre_nms_topk
        L = _lib.lib()
        idx, ks = [], []
        for o in rpn_out:
            lg = o[..., :3].reshape(N, -1).float()
            k = min(P, lg.shape[1])
            idx.append(lg.topk(k, dim=1)[1].contiguous())             # descending
            ks.append(k)
        G = len(rpn_out)
        boxes = torch.empty(N, G, P, 4, device=self.device)
        scores = torch.empty(N, G, P, device=self.device)
        valid = torch.empty(N, G, P, dtype=torch.uint8, device=self.device)
        cells = (_lib.c_float * 60)(*[v for size in (32, 64, 128, 256, 512) for a in cell_anchors(size) for v in a])
        check(L.pvo_rpn_decode(_lib.ptr_array(rpn_out), _lib.ptr_array(idx), _lib.int_array([o.shape[1] for o in rpn_out]),
                               _lib.int_array([o.shape[2] for o in rpn_out]), _lib.int_array(ks), _lib.int_array([4, 8, 16, 32, 64]), cells,
                               ptr(self._limits(image_sizes)), N, P, ptr(boxes), ptr(scores), ptr(valid), stream()), 'rpn_decode')
        keep = self.nms_groups(boxes.view(N * G, P, 4), valid.view(N * G, P), nms_thresh).view(N, G * P)
        sc = torch.where(keep, scores.view(N, G * P), torch.full_like(scores.view(N, G * P), float('-inf')))
        top, ti = sc.topk(min(post_nms_topk, G * P), dim=1)
        bx = torch.gather(boxes.view(N, G * P, 4), 1, ti[..., None].expand(-1, -1, 4))
        counts = keep.sum(dim=1).clamp(max=post_nms_topk).tolist()
        return [(bx[n, :c], top[n, :c]) for n, c in enumerate(counts)]

    def detections(self, cls, reg, props, image_sizes, score_thresh=0.05, nms_thresh=0.5, topk=100):
        """FastRCNNOutputLayers.inference (fast_rcnn.py:117-170): softmax, per-class decode + clip, score threshold,
        per-class NMS, best `topk` per image -> per image (boxes, scores, classes)"""
        N, K = len(props), NUM_THING
        counts = [p[0].shape[0] for p in props]
        R, P = sum(counts), max(max(counts), 1)
        offs = [0]
        for c in counts:
            offs.append(offs[-1] + c)
        offsets = torch.tensor(offs, dtype=torch.int32, device=self.device)
        pb = _cat_rows([p[0] for p in props]).float()
        sp = torch.zeros(N, K, P, device=self.device)
        bp = torch.zeros(N, K, P, 4, device=self.device)
        cls, reg = cls.float().contiguous(), reg.float().contiguous()
        check(_lib.lib().pvo_det_decode(ptr(cls), ptr(reg), ptr(pb), ptr(offsets), ptr(self._limits(image_sizes)), R, N, P, K, ptr(sp), ptr(bp),
                                        stream()), 'det_decode')
        ss, order = sp.sort(dim=2, descending=True)                                           # (N, K, P)
        bs = torch.gather(bp, 2, order[..., None].expand(-1, -1, -1, 4))
        keep = self.nms_groups(bs.view(N * K, P, 4), (ss > score_thresh).view(N * K, P), nms_thresh).view(N, K * P)
        sc = torch.where(keep, ss.view(N, K * P), torch.full_like(ss.view(N, K * P), -1.0))
        top, ti = sc.topk(min(topk, K * P), dim=1)
        bx = torch.gather(bs.view(N, K * P, 4), 1, ti[..., None].expand(-1, -1, 4))
        cl = torch.div(ti, P, rounding_mode='floor')
        kept = keep.sum(dim=1).clamp(max=topk).tolist()
        return [(bx[n, :c], top[n, :c], cl[n, :c]) for n, c in enumerate(kept)]

    # ------------------------------------------------------------------------------------------ ROI heads
    def roi_pool(self, feats, boxes_per_image, out_size):
        """ROIPooler (poolers.py:189-235): ROIAlignV2 over p2..p5, level by box size -> (R, S, S, 256) half, pixel-major"""
        rois = torch.cat([torch.cat([torch.full((b.shape[0], 1), float(i), device=self.device), b.float()], dim=1)
                          for i, b in enumerate(boxes_per_image)]).contiguous()
        R = rois.shape[0]
        out = self._buf(R, out_size, out_size, 256)
        lv = [feats['p%d' % l] for l in (2, 3, 4, 5)]
        fp = (_lib.c_void_p * 4)(*[t[0].data_ptr() for t in lv])
        check(_lib.lib().pvo_roi_align_fpn(fp, _lib.int_array([t[1] for t in lv]), _lib.int_array([t[2] for t in lv]), 256, ptr(rois), R,
                                           out_size, ptr(out), _lib.F16, stream()), 'roi_align_fpn')
        return out

    def box_head(self, pooled):
        """(R,7,7,256) -> class logits (R,4), box deltas (R,12) float32"""
        if pooled.shape[0] == 0:
            return torch.zeros(0, NUM_THING + 1, device=self.device), torch.zeros(0, 4 * NUM_THING, device=self.device)
        x = self._matmul('fc1', pooled.view(pooled.shape[0], -1), relu=True)
        x = self._matmul('fc2', x, relu=True)
        o = self._matmul('box_out', x, cpad=16).float()
        return o[:, :NUM_THING + 1], o[:, NUM_THING + 1:NUM_THING + 1 + 4 * NUM_THING]

    def mask_head(self, pooled):
        """(R,14,14,256) -> mask logits (R, 3, 28, 28) float32"""
        R = pooled.shape[0]
        if R == 0:
            return torch.zeros(0, NUM_THING, 28, 28, device=self.device)
        x = pooled.view(R * 196, 256)
        for j in range(1, 5):
            x = self._conv('mask_fcn%d' % j, x, R, 14, 14, relu=True)
        x = self._matmul('deconv', x, relu=True)                                   # (R*196, 4*256): sub-positions (dy,dx)
        x = self._matmul('mask_pred', x.reshape(R * 196 * 4, 256), cpad=16).float()[:, :NUM_THING]
        return x.view(R, 14, 14, 2, 2, NUM_THING).permute(0, 5, 1, 3, 2, 4).reshape(R, NUM_THING, 28, 28)

    # ------------------------------------------------------------------------------------------ panoptic merge
    def paste(self, prob, boxes, H, W, threshold=0.5):
        """paste_masks_in_image (mask_ops.py:74-152): prob (R,1,28,28) probabilities, boxes (R,4) -> (R,H,W) bool"""
        R = prob.shape[0]
        out = torch.empty(R, H, W, dtype=torch.uint8, device=self.device)
        prob, boxes = prob.float().contiguous(), boxes.float().contiguous()
        check(_lib.lib().pvo_paste_masks(ptr(prob), ptr(boxes), R, prob.shape[-1], H, W, float(threshold), ptr(out), stream()), 'paste_masks')
        return out.view(torch.bool)

    def combine(self, dets_masks, sem_labels, overlap_threshold=0.5, stuff_area_thresh=4096, instances_score_thresh=0.5, boxes=None):
        """combine_semantic_and_instance_outputs (panoptic_fpn.py:497-588) for images of one size, one launch.
        dets_masks: per image (scores (R,), classes (R,), masks (R,H,W) bool); sem_labels (N,H,W) int64; boxes: optional
        per-image (R,4) detection boxes that bound the masks (limits each instance's scan to its box).
        -> pan (N,H,W) int32, per-image segments_info lists."""
        N, H, W = sem_labels.shape
        scores = torch.cat([d[0] for d in dets_masks]).float().contiguous()
        classes = torch.cat([d[1] for d in dets_masks]).long().contiguous()
        masks = _cat_rows([d[2] for d in dets_masks])
        order = torch.cat([torch.argsort(-d[0].float()) for d in dets_masks]).long().contiguous()
        off = [0]
        for d in dets_masks:
            off.append(off[-1] + d[0].shape[0])
        inst_off = torch.tensor(off, dtype=torch.int32, device=self.device)
        max_info = 128
        pan = torch.empty(N, H, W, dtype=torch.int32, device=self.device)
        info = torch.zeros(N, max_info, 5, dtype=torch.int32, device=self.device)
        counts = torch.zeros(N, dtype=torch.int32, device=self.device)
        sem_labels = sem_labels.contiguous()
        bx = torch.cat(boxes).float().contiguous() if boxes is not None else None
        check(_lib.lib().pvo_panoptic_combine(ptr(masks), ptr(bx), ptr(scores), ptr(classes), ptr(order), ptr(inst_off), ptr(sem_labels), ptr(pan),
                                              ptr(info), ptr(counts), N, H, W, NUM_SEM, float(overlap_threshold), int(stuff_area_thresh),
                                              float(instances_score_thresh), THING_CAT0, max_info, stream()), 'panoptic_combine')
        info_l, counts_l, scores_l = info.tolist(), counts.tolist(), scores.tolist()       # the only device->host reads
        segs = []
        for n in range(N):
            lst = []
            for sid, cat, thing, area, inst in info_l[n][:min(counts_l[n], max_info)]:
                if thing:
                    lst.append({'id': sid, 'isthing': True, 'score': scores_l[off[n] + inst], 'category_id': cat, 'instance_id': inst})
                else:
                    lst.append({'id': sid, 'isthing': False, 'category_id': cat, 'area': area})
            segs.append(lst)
        return pan, segs

    # ------------------------------------------------------------------------------------------ whole model
    @torch.no_grad()
    def inference(self, images, debug=None):
        """PanopticFPN.inference (panoptic_fpn.py:453-494): list of (3,H,W) BGR images (or dicts with 'image') ->
        per image {'sem_seg', 'instances', 'panoptic_seg': (ids (H,W) int32, segments_info)}"""
        images = [im['image'] if isinstance(im, dict) else im for im in images]
        x, sizes = self.preprocess(images)
        feats, sem, rpn = self.dense(x)
        return self._roi_heads_and_merge(feats, sem, rpn, sizes, debug)

    def _roi_heads_and_merge(self, feats, sem, rpn, sizes, debug=None):
        """proposals -> box head -> detections -> mask head -> paste -> merge for the len(sizes) images the pyramid holds"""
        props = self.proposals(rpn, sizes)
        cls, reg = self.box_head(self.roi_pool(feats, [p[0] for p in props], 7))
        dets = self.detections(cls, reg, props, sizes)
        ml = self.mask_head(self.roi_pool(feats, [d[0] for d in dets], 14))
        results, per_image = [], []
        same = all(sz == sizes[0] for sz in sizes)
        R_all = ml.shape[0]
        cls_all = torch.cat([d[2] for d in dets]) if R_all else torch.zeros(0, dtype=torch.long, device=self.device)
        prob_all = ml[torch.arange(R_all, device=self.device), cls_all][:, None].sigmoid()      # mask of the predicted class
        if same:                                            # one paste launch for the batch; per-image views, no concatenation later
            masks_all = self.paste(prob_all, torch.cat([d[0] for d in dets]), sizes[0][0], sizes[0][1])
        o = 0
        for n, (b, s, c) in enumerate(dets):
            R = b.shape[0]
            per_image.append((s, c, masks_all[o:o + R] if same else self.paste(prob_all[o:o + R], b, sizes[n][0], sizes[n][1])))
            o += R
        if same:
            h, w = sizes[0]
            pans, segs = self.combine(per_image, sem[:, :h, :w, :NUM_SEM].argmax(dim=-1), boxes=[d[0] for d in dets])
        else:
            pans, segs = [], []
            for n, (h, w) in enumerate(sizes):
                p_, s_ = self.combine(per_image[n:n + 1], sem[n:n + 1, :h, :w, :NUM_SEM].argmax(dim=-1), boxes=[dets[n][0]])
                pans.append(p_[0])
                segs.append(s_[0])
        for n, (b, s, c) in enumerate(dets):
            h, w = sizes[n]
            results.append({'sem_seg': sem[n, :h, :w, :NUM_SEM].permute(2, 0, 1),
                            'instances': {'pred_boxes': b, 'scores': s, 'pred_classes': c, 'pred_masks': per_image[n][2]},
                            'panoptic_seg': (pans[n], segs[n])})
        if debug is not None:
            debug.update(feats=feats, sem=sem, rpn=rpn, proposals=props, box_cls=cls, box_reg=reg, detections=dets, mask_logits=ml)
        return results

    # ------------------------------------------------------------------------------------------ video fusion branch
    def fuse_pair(self, ref_image, ref_flow, cur_image):
        """wrap_in_stage2 up to the heads (panoptic_fpn.py:193-212): backbone on (ref, cur) as one batch of two; per level
        cat(cur features, alpha * ref features carried along the flow) -> fusion_conv1 on the tensor cores.
        -> fused pyramid of ONE image ({'p2'..'p6': ((h*w, 256) half, h, w)}), its sizes list"""
        if not self.fusion:
            raise RuntimeError('the state dict holds no fusion_conv1: build PanopticFPN from a FUSION checkpoint')
        x, sizes = self.preprocess([ref_image, cur_image])
        feats = self.backbone(x)
        flow = ref_flow.to(self.device)
        fused = {}
        for k, (f, h, w) in feats.items():
            f = f.view(2, h * w, 256)
            carried = flow_transport(f[0], flow_transport_index(flow, h, w))
            if self.alpha != 1.0:
                carried = carried * self.alpha
            pair = torch.cat([f[1], carried], dim=1).contiguous()               # (h*w, 512): current | carried
            fused[k] = (self._conv('fusion_conv1', pair, 1, h, w), h, w)
        self.N = 1
        return fused, sizes[1:]

    @torch.no_grad()
    def inference_fusion(self, batched_inputs, debug=None):
        """PanopticFPN.inference_fusion (panoptic_fpn.py:254-308), FUSION = True, flow transport on, depth transport off (the
        shipped settings, :71-73).  batched_inputs: [{'image': (3,H,W) BGR, 'image_id': str, optional 'flow': (2,H,W)}] -- one
        frame.  The first frame of a video (image_id[:4] changes) takes the single-image path; later frames are fused with
        the previous frame's features carried along its flow.  The frame's own image / flow are kept for the next call."""
        inp = batched_inputs[0]
        vid = inp['image_id'][:4]
        if vid != self.vid:
            self.vid = vid
            res = self.inference([inp['image']], debug)
        else:
            fused, sizes = self.fuse_pair(self.ref_image, self.ref_flow, inp['image'])
            res = self._roi_heads_and_merge(fused, self.sem_seg(fused), self.rpn_head(fused), sizes, debug)
        if 'flow' in inp:
            self.ref_flow, self.ref_image, self.ref_id = inp['flow'], inp['image'], inp['image_id']
        return res

    __call__ = inference
