"""GraphAgg / DynamicUpdateModule with the reference's parameters and call signature
(VO_Module/droid_slam/droid_net.py:64-95,167-314) on the fused CUDA update operator (pvo_b200/nn.py).

State-dict keys match the reference checkpoint (``update.corr_encoder.0.weight`` ... SURVEY.md Appendix A).
The modules are inference-only: they raise if asked to differentiate (training is out of scope).
"""
import torch
import torch.nn as nn

from . import _lib
from . import nn as pnn
from .modules.gru import ConvGRU


class GradientClip(nn.Module):
    """modules/clipping.py:19-24: identity in the forward pass."""

    def forward(self, x):
        return x


class GraphAgg(nn.Module):
    """droid_net.py:64-95 (parameters only; the arithmetic runs inside UpdateEngine)."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(128, 128, 3, padding=1)
        self.conv2 = nn.Conv2d(128, 128, 3, padding=1)
        self.relu = nn.ReLU(inplace=True)
        self.eta = nn.Sequential(nn.Conv2d(128, 1, 3, padding=1), GradientClip(), nn.Softplus())
        self.upmask_disp = nn.Sequential(nn.Conv2d(128, 8 * 8 * 9, 1, padding=0))


def _head(cout):
    return nn.Sequential(nn.Conv2d(128, 128, 3, padding=1), nn.ReLU(inplace=True),
                         nn.Conv2d(128, cout, 3, padding=1), GradientClip())


class LazyUpmask(dict):
    """upmask dict whose 'disp' entry (GraphAgg's 128->576 1x1 conv, 0.71 GFLOP per frame) is computed only when
    read: FactorGraph.update discards it (factor_graph.py:242), DroidNet.forward consumes it (droid_net.py:420)."""

    def __init__(self, module, shape):
        super().__init__(flow=None, dy_mask=None)
        self._m, self._shape = module, shape
        self._eng = module._engine
        self._gen = self._eng.generation          # the engine run whose GraphAgg features this mask belongs to

    def __getitem__(self, k):
        if k == 'disp' and not dict.__contains__(self, 'disp'):
            eng = self._eng
            if eng.generation != self._gen or self._m._engine is not eng:
                raise RuntimeError('upmask["disp"] was read after another forward() of the same update module: its GraphAgg '
                                   'features are gone -- read it before the next call (the reference returns a concrete tensor)')
            up = eng.run_upmask()
            B, K, H, W = self._shape
            nchw = eng.store_nchw(up, 576, 0, 576, K)
            dict.__setitem__(self, 'disp', nchw.view(B, K, 576, H, W))
        return dict.__getitem__(self, k)


class DynamicUpdateModule(nn.Module):
    """PVO's update head (droid_net.py:167-314).  precision: 'auto' (float16 when the state is half or autocast is
    on, as the reference's inference; else float32), 'fp32', 'fp16'."""

    def __init__(self, use_aff_bri=False, precision='auto'):
        super().__init__()
        if use_aff_bri:
            raise NotImplementedError('use_aff_bri=True is a training-time option (droid_net.py:216-221)')
        self.mask_num = 2
        self.precision = precision
        self.corr_encoder = nn.Sequential(nn.Conv2d(196, 128, 1, padding=0), nn.ReLU(inplace=True),
                                          nn.Conv2d(128, 128, 3, padding=1), nn.ReLU(inplace=True))
        self.flow_encoder = nn.Sequential(nn.Conv2d(4 + self.mask_num + 2, 128, 7, padding=3), nn.ReLU(inplace=True),
                                          nn.Conv2d(128, 64, 3, padding=1), nn.ReLU(inplace=True))
        self.weight = _head(2)
        self.delta = _head(2)
        self.delta_dy = _head(2)
        self.delta_mask = _head(self.mask_num)
        self.gru = ConvGRU(128, 128 + 128 + 64)
        self.agg = GraphAgg()
        self._engine = None
        self._engine_key = None
        self._forks = None
        self.register_load_state_dict_post_hook(lambda m, k: setattr(m, '_engine', None))

    # -------------------------------------------------------------------------------------------
    def engine(self, mode, device, owner=None):
        """The fused-kernel engine for (mode, device): weights packed once per load_state_dict.  owner=None -> the engine
        forward() runs on; owner=<a FactorGraph> -> that graph's private fork (same weights, own buffers and hidden state),
        dropped with the graph."""
        key = (mode, str(device))
        if self._engine is None or self._engine_key != key:
            sd = {k: v for k, v in self.state_dict().items()}
            self._engine = pnn.UpdateEngine(sd, mode=mode, device=device)
            self._engine_key = key
        if owner is None:
            return self._engine
        if self._forks is None:
            import weakref
            self._forks = weakref.WeakKeyDictionary()
        e = self._forks.get(owner)
        if e is None or e.base is not self._engine:
            e = self._forks[owner] = self._engine.fork()
        return e

    def _mode(self, net):
        if self.precision in ('fp32', 'fp16'):
            return self.precision
        return 'fp16' if (net.dtype == torch.float16 or torch.is_autocast_enabled()) else 'fp32'

    def forward(self, net, inp, corr, flow=None, ii=None, jj=None, use_aff_bri=False, raw_mask=None, segments=None):
        """net, inp (B,E,128,h,w); corr (B,E,196,h,w); flow (B,E,8,h,w).  Returns, like the reference,
        (net, delta (B,E,h,w,4), weight (B,E,h,w,2), eta, upmask, delta_m) or (net, delta, weight, delta_m)."""
        if use_aff_bri:
            raise NotImplementedError('use_aff_bri')
        for t in (net, inp, corr, flow):
            if t is not None and t.requires_grad:
                raise NotImplementedError('DynamicUpdateModule: inference only (no autograd through the CUDA path)')
        batch, num, ch, ht, wd = net.shape
        E = batch * num
        mode = self._mode(net)
        eng = self.engine(mode, net.device)
        K = 0
        if ii is not None:
            ii = ii.to(net.device)
            K = int(torch.unique(ii).shape[0])             # the reference syncs here too (droid_net.py:83)
        eng.reserve(E, max(K, 1), ht, wd)
        eng.load_nchw(net.reshape(E, -1, ht, wd).contiguous(), eng.X, 448, 0)
        eng.load_nchw(inp.reshape(E, -1, ht, wd).contiguous(), eng.X, 448, 128)
        eng.load_nchw(corr.reshape(E, -1, ht, wd).contiguous(), eng.corr, pnn.CORR_STRIDE, 0)
        if flow is None:
            eng.motn.zero_()
        else:
            eng.load_nchw(flow.reshape(E, -1, ht, wd).contiguous(), eng.motn, 8, 0)
        if ii is not None:
            eng.group(ii)
        eng.run(with_agg=ii is not None)
        net_out = eng.store_nchw(eng.X, 448, 0, 128, E).view(batch, num, 128, ht, wd)
        ho = eng.head_out.view(batch, num, ht, wd, 8)
        delta = ho[..., 0:4].contiguous()
        weight = ho[..., 4:6].contiguous()
        delta_m = ho[..., 6:8].contiguous()
        if ii is None:
            return net_out, delta, weight, delta_m
        eta = eng.eta.view(batch, K, ht, wd).clone()
        upmask = LazyUpmask(self, (batch, K, ht, wd))
        return net_out, delta, weight, eta, upmask, delta_m


UpdateModule = DynamicUpdateModule     # the stale reference class (droid_net.py:98-164) is never instantiated by PVO


# ------------------------------------------------------------------------------------------------------
# DroidNet: encoders + update operator + (Python-semantics) BA, pairwise / training-style entry point
# ------------------------------------------------------------------------------------------------------
def cvx_upsample(data, mask):
    """convex 8x upsampling with a softmax-9 mask (droid_net.py:23-37): one kernel (csrc/upsample.cu) instead of the reference's
    softmax + F.unfold + broadcast product + sum.  data (B,ht,wd,dim), mask (B,576,ht,wd) or anything that views to it."""
    batch, ht, wd, dim = data.shape
    if not data.is_cuda:
        raise RuntimeError('pvo_b200.cvx_upsample: CUDA tensors only (no CPU fallback)')
    d = data.float().contiguous()
    m = mask.reshape(batch, 576, ht, wd).contiguous()
    if m.dtype not in (torch.float32, torch.float16):
        m = m.float()
    out = torch.empty(batch, 8 * ht, 8 * wd, dim, dtype=torch.float32, device=data.device)
    _lib.check(_lib.lib().pvo_cvx_upsample(_lib.ptr(d), _lib.ptr(m), _lib.ptr(out), batch, ht, wd, dim, _lib.dtype_code(m), _lib.stream()),
               'cvx_upsample')
    return out.to(torch.promote_types(data.dtype, mask.dtype))


def upsample_dim_1(disp, mask):
    batch, num, ht, wd = disp.shape
    return cvx_upsample(disp.reshape(batch * num, ht, wd, 1), mask.reshape(batch * num, -1, ht, wd)).view(batch, num, 8 * ht, 8 * wd)


def upsample_dim_x(flow, mask):
    """droid_net.py:46-51"""
    batch, num, ht, wd, dim = flow.shape
    return cvx_upsample(flow.reshape(batch * num, ht, wd, dim), mask.reshape(batch * num, -1, ht, wd)).view(batch, num, 8 * ht, 8 * wd, dim)


def upsample_inter(mask):
    """bilinear x8, align_corners=True (droid_net.py:54-61; also droid.py:106, factor_graph.py:431), one kernel"""
    batch, num, ht, wd, dim = mask.shape
    if not mask.is_cuda:
        raise RuntimeError('pvo_b200.upsample_inter: CUDA tensors only (no CPU fallback)')
    m = mask.float().contiguous()
    out = torch.empty(batch, num, 8 * ht, 8 * wd, dim, dtype=torch.float32, device=mask.device)
    _lib.check(_lib.lib().pvo_upsample_inter(_lib.ptr(m), _lib.ptr(out), batch * num, ht, wd, dim, _lib.stream()), 'upsample_inter')
    return out.to(mask.dtype)


class DroidNet(nn.Module):
    """droid_net.py:317-440.  Checkpoint keys fnet.*, cnet.*, update.* as the reference.  forward() is the reference's
    pairwise inference loop (evaluation_scripts/test_vo2.py:110-113): per step one correlation lookup, one update
    operator call and two Gauss-Newton steps of geom.ba.BA -- all on the CUDA kernels; inference only."""

    def __init__(self, use_aff_bri=False):
        super().__init__()
        from .modules.extractor import BasicEncoder
        self.fnet = BasicEncoder(output_dim=128, norm_fn='instance')
        self.cnet = BasicEncoder(output_dim=256, norm_fn='none')
        self.update = DynamicUpdateModule(use_aff_bri)
        self.use_aff_bri = use_aff_bri

    def extract_features(self, images):
        """droid_net.py:325-340"""
        images = images[:, :, [2, 1, 0]] / 255.0
        mean = torch.as_tensor([0.485, 0.456, 0.406], device=images.device)
        std = torch.as_tensor([0.229, 0.224, 0.225], device=images.device)
        images = images.sub_(mean[:, None, None]).div_(std[:, None, None])
        fmaps = self.fnet(images)
        net = self.cnet(images)
        net, inp = net.split([128, 128], dim=2)
        return fmaps, torch.tanh(net), torch.relu(inp)

    @torch.no_grad()
    def forward(self, Gs, images, disps, intrinsics, graph=None, num_steps=12, fixedp=2, ret_flow=False, downsample=False,
                segments=None):
        from .geom import projective_ops as pops
        from .geom.ba import BA
        from .modules.corr import CorrBlock
        dev = images.device
        ii = torch.as_tensor([u for u in graph for _ in graph[u]], dtype=torch.long, device=dev)      # graph_to_edge_list
        jj = torch.as_tensor([v for u in graph for v in graph[u]], dtype=torch.long, device=dev)
        fmaps, net, inp = self.extract_features(images)
        net, inp = net[:, ii].contiguous(), inp[:, ii].contiguous()
        corr_fn = CorrBlock(fmaps[:, ii].contiguous(), fmaps[:, jj].contiguous(), num_levels=4, radius=3)
        ht, wd = images.shape[-2:]
        coords0 = pops.coords_grid(ht // 8, wd // 8, device=dev)
        coords1, _ = pops.projective_transform(Gs, disps, intrinsics, ii, jj)
        target_cam = coords1.clone()
        delta_dy = torch.zeros_like(coords1)
        raw_mask = torch.zeros_like(coords1)
        Gs_list, disp_list, residual_list, full_flow_list, mask_list = [], [], [], [], []
        for step in range(num_steps):
            corr = corr_fn(coords1)
            resd = target_cam - coords1
            cam_flow = coords1 - coords0
            motion = torch.cat([cam_flow, cam_flow + delta_dy, resd, raw_mask], dim=-1)
            motion = motion.permute(0, 1, 4, 2, 3).clamp(-64.0, 64.0).contiguous()
            net, delta, weight, eta, upmask, delta_m = self.update(net, inp, corr, motion, ii, jj)
            raw_mask = raw_mask + delta_m
            mask = torch.sigmoid(raw_mask)
            bin_mask = (mask >= 0.5).float()
            target_cam = coords1 + delta[..., 0:2]
            weight = torch.sigmoid(weight + (1 - bin_mask) * 10)
            for _ in range(2):
                Gs, disps = BA(target_cam, weight, eta, Gs, disps, intrinsics, ii, jj, fixedp=fixedp)
            coords1, valid_mask = pops.projective_transform(Gs, disps, intrinsics, ii, jj)
            residual = (target_cam - coords1) * valid_mask
            delta_dy = delta[..., 2:4] * (1 - bin_mask)
            target_all = coords1 + delta_dy
            Gs_list.append(Gs)
            disp_list.append(upsample_dim_1(disps, upmask['disp'].float()))
            residual_list.append(residual)
            mask_list.append(upsample_inter(mask))
            if ret_flow:
                full_flow_list.append(target_all - coords0 if downsample else upsample_inter((target_all - coords0) * 8))
        if ret_flow:
            return Gs_list, disp_list, residual_list, full_flow_list, mask_list
        return Gs_list, disp_list, residual_list, mask_list
