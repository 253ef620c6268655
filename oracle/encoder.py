"""Plain-torch CPU restatement of BasicEncoder -- oracle, test infrastructure only.

Follows VO_Module/droid_slam/modules/extractor.py:6-55 (ResidualBlock) and :118-198 (BasicEncoder, DIM = 32):
conv1 7x7 s2 -> norm -> ReLU -> layer1 (2 x Res(32)) -> layer2 (Res(32->64, s2), Res(64)) -> layer3 (Res(64->128, s2),
Res(128)) -> conv2 1x1.  norm_fn 'instance' (fnet: nn.InstanceNorm2d, no affine) or 'none' (cnet).
"""
import torch
import torch.nn.functional as F

DIM = 32


def param_shapes(output_dim):
    s = {}

    def conv(name, co, ci, k):
        s[name + '.weight'] = (co, ci, k, k)
        s[name + '.bias'] = (co,)
    conv('conv1', DIM, 3, 7)
    cin = DIM
    for li, dim in ((1, DIM), (2, 2 * DIM), (3, 4 * DIM)):
        for bi in (0, 1):
            conv('layer%d.%d.conv1' % (li, bi), dim, cin if bi == 0 else dim, 3)
            conv('layer%d.%d.conv2' % (li, bi), dim, dim, 3)
            if bi == 0 and li > 1:
                conv('layer%d.0.downsample.0' % li, dim, cin, 1)
        cin = dim
    conv('conv2', output_dim, 4 * DIM, 1)
    return s


def init_params(output_dim, seed=0):
    """kaiming_normal(fan_out, relu) weights as the reference constructor (:158-165), seeded; small random biases."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in param_shapes(output_dim).items():
        if name.endswith('.weight'):
            fan_out = shape[0] * shape[2] * shape[3]
            sd[name] = torch.randn(shape, generator=g) * (2.0 / fan_out) ** 0.5
        else:
            sd[name] = 0.1 * torch.randn(shape, generator=g)
    return sd


def _norm(x, norm_fn):
    return F.instance_norm(x, eps=1e-5) if norm_fn == 'instance' else x


def _res_block(sd, pre, x, norm_fn, stride):
    y = torch.relu(_norm(F.conv2d(x, sd[pre + '.conv1.weight'], sd[pre + '.conv1.bias'], stride=stride, padding=1), norm_fn))
    y = torch.relu(_norm(F.conv2d(y, sd[pre + '.conv2.weight'], sd[pre + '.conv2.bias'], padding=1), norm_fn))
    if stride != 1:
        x = _norm(F.conv2d(x, sd[pre + '.downsample.0.weight'], sd[pre + '.downsample.0.bias'], stride=stride), norm_fn)
    return torch.relu(x + y)


def basic_encoder(sd, x, norm_fn):
    """x (b,n,3,H,W) -> (b,n,output_dim,H/8,W/8)."""
    b, n, c1, h1, w1 = x.shape
    x = x.reshape(b * n, c1, h1, w1)
    x = torch.relu(_norm(F.conv2d(x, sd['conv1.weight'], sd['conv1.bias'], stride=2, padding=3), norm_fn))
    for li, stride in ((1, 1), (2, 2), (3, 2)):
        x = _res_block(sd, 'layer%d.0' % li, x, norm_fn, stride)
        x = _res_block(sd, 'layer%d.1' % li, x, norm_fn, 1)
    x = F.conv2d(x, sd['conv2.weight'], sd['conv2.bias'])
    return x.view(b, n, x.shape[1], x.shape[2], x.shape[3])


def extract_features(sd_f, sd_c, images):
    """DroidNet.extract_features (droid_net.py:325-340): images (b,n,3,H,W) in 0..255 BGR->RGB, ImageNet normalisation."""
    images = images[:, :, [2, 1, 0]] / 255.0
    mean = torch.as_tensor([0.485, 0.456, 0.406])
    std = torch.as_tensor([0.229, 0.224, 0.225])
    images = (images - mean[:, None, None]) / std[:, None, None]
    fmaps = basic_encoder(sd_f, images, 'instance')
    net = basic_encoder(sd_c, images, 'none')
    net, inp = net.split([128, 128], dim=2)
    return fmaps, torch.tanh(net), torch.relu(inp)
