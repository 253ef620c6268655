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

    def __getitem__(self, k):
        if k == 'disp' and not dict.__contains__(self, 'disp'):
            eng = self._m._engine
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
        self.register_load_state_dict_post_hook(lambda m, k: setattr(m, '_engine', None))

    # -------------------------------------------------------------------------------------------
    def engine(self, mode, device):
        key = (mode, str(device))
        if self._engine is None or self._engine_key != key:
            sd = {k: v for k, v in self.state_dict().items()}
            self._engine = pnn.UpdateEngine(sd, mode=mode, device=device)
            self._engine_key = key
        return self._engine

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
