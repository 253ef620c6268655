"""Plain-torch CPU restatement of PVO's update operator -- oracle, test infrastructure only.

Functional (no nn.Module): every function takes ``sd``, a dict of tensors carrying the
reference checkpoint's parameter names, so the same weights drive
the oracle, the reference classes and the CUDA path.

Follows (reference file:line, under VO_Module/droid_slam/):
  * modules/gru.py:6-32         ConvGRU (global-context gate; channel order [net|inp|corr|flow])
  * droid_net.py:64-95          GraphAgg (scatter_mean over edges sharing ii, eta = .01*softplus)
  * droid_net.py:167-225        DynamicUpdateModule parameter shapes
  * droid_net.py:256-314        DynamicUpdateModule.forward (np.range at :295 is dead code and skipped)
  * modules/clipping.py:7-17    GradientClip is the identity in the forward pass
"""
import torch
import torch.nn.functional as F
from .geom import scatter_mean


def param_shapes():
    """name -> shape of every tensor in ``DroidNet.update`` (110-tensor checkpoint subset)."""
    s = {}

    def conv(name, co, ci, k):
        s[name + '.weight'] = (co, ci, k, k)
        s[name + '.bias'] = (co,)
    conv('corr_encoder.0', 128, 196, 1)
    conv('corr_encoder.2', 128, 128, 3)
    conv('flow_encoder.0', 128, 8, 7)
    conv('flow_encoder.2', 64, 128, 3)
    for head in ('weight', 'delta', 'delta_dy', 'delta_mask'):
        conv(head + '.0', 128, 128, 3)
        conv(head + '.2', 2, 128, 3)
    for g in ('convz', 'convr', 'convq'):
        conv('gru.' + g, 128, 448, 3)
    for g in ('w', 'convz_glo', 'convr_glo', 'convq_glo'):
        conv('gru.' + g, 128, 128, 1)
    conv('agg.conv1', 128, 128, 3)
    conv('agg.conv2', 128, 128, 3)
    conv('agg.eta.0', 1, 128, 3)
    conv('agg.upmask_disp.0', 576, 128, 1)
    return s


def init_params(seed=0, dtype=torch.float32):
    """Deterministic PyTorch-default-like init (kaiming-uniform(a=sqrt5) == U(-1/sqrt(fan_in), ..))."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in param_shapes().items():
        if name.endswith('.weight'):
            fan_in = shape[1] * shape[2] * shape[3]
        else:
            w = param_shapes()[name[:-5] + '.weight']
            fan_in = w[1] * w[2] * w[3]
        bound = 1.0 / fan_in ** 0.5
        sd[name] = ((torch.rand(shape, generator=g) * 2 - 1) * bound).to(dtype)
    return sd


def _conv(sd, name, x, pad):
    return F.conv2d(x, sd[name + '.weight'], sd[name + '.bias'], padding=pad)


def conv_gru(sd, net, inp_cat):
    """modules/gru.py:19-32."""
    b, c, h, w = net.shape
    net_inp = torch.cat([net, inp_cat], dim=1)
    glo = torch.sigmoid(_conv(sd, 'gru.w', net, 0)) * net
    glo = glo.view(b, c, h * w).mean(-1).view(b, c, 1, 1)
    z = torch.sigmoid(_conv(sd, 'gru.convz', net_inp, 1) + _conv(sd, 'gru.convz_glo', glo, 0))
    r = torch.sigmoid(_conv(sd, 'gru.convr', net_inp, 1) + _conv(sd, 'gru.convr_glo', glo, 0))
    q = torch.tanh(_conv(sd, 'gru.convq', torch.cat([r * net, inp_cat], dim=1), 1)
                   + _conv(sd, 'gru.convq_glo', glo, 0))
    return (1 - z) * net + z * q


def graph_agg(sd, net, ii):
    """droid_net.py:79-95.  net (B,E,128,h,w) -> eta (B,F',h,w), upmask (B,F',576,h,w)."""
    batch, num, ch, ht, wd = net.shape
    _, ix = torch.unique(ii, return_inverse=True)
    x = torch.relu(_conv(sd, 'agg.conv1', net.reshape(batch * num, ch, ht, wd), 1))
    x = scatter_mean(x.view(batch, num, 128, ht, wd), ix)
    x = x.view(-1, 128, ht, wd)
    x = torch.relu(_conv(sd, 'agg.conv2', x, 1))
    eta = F.softplus(_conv(sd, 'agg.eta.0', x, 1)).view(batch, -1, ht, wd)
    upmask = _conv(sd, 'agg.upmask_disp.0', x, 0).view(batch, -1, 576, ht, wd)
    return .01 * eta, upmask


def update_op(sd, net, inp, corr, flow, ii=None):
    """DynamicUpdateModule.forward (droid_net.py:256-314).

    net, inp (B,E,128,h,w); corr (B,E,196,h,w); flow (B,E,8,h,w).
    Returns net, delta (B,E,h,w,4) = [delta | delta_dy], weight (B,E,h,w,2) raw logits,
    eta, upmask, delta_m (B,E,h,w,2); eta/upmask are None when ii is None.
    """
    batch, num, ch, ht, wd = net.shape
    od = (batch, num, -1, ht, wd)
    net = net.reshape(batch * num, -1, ht, wd)
    inp = inp.reshape(batch * num, -1, ht, wd)
    corr = corr.reshape(batch * num, -1, ht, wd)
    flow = flow.reshape(batch * num, -1, ht, wd)
    corr = torch.relu(_conv(sd, 'corr_encoder.2', torch.relu(_conv(sd, 'corr_encoder.0', corr, 0)), 1))
    flow = torch.relu(_conv(sd, 'flow_encoder.2', torch.relu(_conv(sd, 'flow_encoder.0', flow, 3)), 1))
    net = conv_gru(sd, net, torch.cat([inp, corr, flow], dim=1))

    def head(name):
        y = _conv(sd, name + '.2', torch.relu(_conv(sd, name + '.0', net, 1)), 1)
        return y.view(*od).permute(0, 1, 3, 4, 2).contiguous()
    delta, delta_dy, weight, delta_m = head('delta'), head('delta_dy'), head('weight'), head('delta_mask')
    net = net.view(*od)
    delta = torch.cat([delta, delta_dy], dim=-1)
    if ii is None:
        return net, delta, weight, None, None, delta_m
    eta, upmask = graph_agg(sd, net, ii)
    return net, delta, weight, eta, upmask, delta_m
