"""GPU parity: the update operator (DynamicUpdateModule / ConvGRU / GraphAgg) against the float32 CPU oracle and the
golden vectors produced by the reference's own module; the tensor-core convolution against the CUDA-core one."""
import os

import numpy as np
import pytest
import torch

from oracle import update_op as oupd
from tests.golden import cases

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), 'golden')


def _module(precision):
    from pvo_b200.droid_net import DynamicUpdateModule
    m = DynamicUpdateModule(precision=precision).cuda()
    m.load_state_dict(oupd.init_params(0))
    return m


def _run(m, inputs, ii):
    net, inp, corr, flow = [t.cuda() for t in inputs]
    with torch.no_grad():
        return m(net, inp, corr, flow, ii.cuda(), None)


def test_update_op_fp32_matches_reference_golden():
    """float32 CUDA path vs tests/golden/update_op.npz (the reference's DynamicUpdateModule run on CPU)."""
    gold = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(G, 'update_op.npz')).items()}
    out = _run(_module('fp32'), cases.update_inputs(E=3, h=12, w=16, seed=0), gold['ii'])
    net, delta, weight, eta, upmask, delta_m = out
    for name, val in (('net', net), ('delta', delta), ('weight', weight), ('eta', eta), ('delta_m', delta_m),
                      ('upmask', upmask['disp'][:, :, ::16])):
        err = (val.float().cpu() - gold[name]).abs().max().item()
        assert err < 2e-5, (name, err)


@pytest.mark.parametrize('E,h,w', [(2, 30, 40), (5, 16, 24)])
def test_update_op_fp32_matches_oracle(E, h, w):
    """BASELINE config 1 shapes (2 edges, 240x320 -> 30x40)."""
    sd = oupd.init_params(0)
    inputs = cases.update_inputs(E=E, h=h, w=w, seed=3)
    ii = torch.arange(E) // 2
    with torch.no_grad():
        ref = oupd.update_op(sd, *inputs, ii)
    out = _run(_module('fp32'), inputs, ii)
    names = ('net', 'delta', 'weight', 'eta', None, 'delta_m')
    for name, a, b in zip(names, out, ref):
        if name is None:
            continue
        err = (a.float().cpu() - b).abs().max().item()
        assert err < 3e-5, (name, err)
    assert (out[4]['disp'].float().cpu() - ref[4]).abs().max().item() < 3e-5


@pytest.mark.parametrize('precision', ['fp16'])
def test_update_op_half_close_to_float32_oracle(precision):
    """The half path (tensor cores) is not a parity path: report and bound its deviation from the fp32 oracle."""
    sd = oupd.init_params(0)
    inputs = cases.update_inputs(E=4, h=16, w=32, seed=4)
    ii = torch.tensor([0, 0, 1, 2])
    with torch.no_grad():
        ref = oupd.update_op(sd, *inputs, ii)
    out = _run(_module(precision), [t.half() for t in inputs], ii)
    for name, a, b in zip(('net', 'delta', 'weight', 'eta'), out[:4], ref[:4]):
        err = (a.float().cpu() - b).abs().max().item()
        assert err < 3e-2, (name, err)


def _rand_conv_case(seed, nimg, H, W, segs, cout, kh, kw):
    g = torch.Generator().manual_seed(seed)
    cin = sum(segs)
    w = torch.randn(cout, cin, kh, kw, generator=g) / (cin * kh * kw) ** 0.5
    b = torch.randn(cout, generator=g)
    xs = [torch.randn(nimg * H * W, c + 8, generator=g).half() for c in segs]     # stride = channels + 8
    return w, b, xs


@pytest.mark.parametrize('case', [
    dict(nimg=3, H=16, W=24, segs=[128], cout=128, kh=3, kw=3),
    dict(nimg=2, H=20, W=40, segs=[128, 320], cout=128, kh=3, kw=3),
    dict(nimg=2, H=9, W=21, segs=[448], cout=256, kh=3, kw=3),
    dict(nimg=4, H=12, W=16, segs=[208], cout=128, kh=1, kw=1),
    dict(nimg=2, H=17, W=33, segs=[128], cout=64, kh=3, kw=3),
    dict(nimg=2, H=16, W=16, segs=[128], cout=640, kh=3, kw=3),
    dict(nimg=1, H=24, W=16, segs=[64], cout=128, kh=7, kw=1),
    # enough tiles for the stacked-M-tile halo kernel (two 16x8 tiles share each weight tile), ragged in both directions
    dict(nimg=16, H=72, W=116, segs=[128], cout=128, kh=3, kw=3),
    dict(nimg=16, H=72, W=116, segs=[64, 64], cout=128, kh=3, kw=3),
])
def test_conv_tensor_core_matches_cuda_core_and_torch(case):
    """pvo_conv2d_tc (TMA + tcgen05 + TMEM) vs pvo_conv2d (CUDA cores) vs F.conv2d on the same half inputs."""
    from pvo_b200 import nn as pnn, _lib
    w, b, xs = _rand_conv_case(11, **case)
    nimg, H, W, segs, cout = case['nimg'], case['H'], case['W'], case['segs'], case['cout']
    cv = pnn._Conv(w.cuda(), b.cuda(), tc=True, segs=segs)
    xs_d = [x.cuda() for x in xs]
    seg_list = [(x, c, c + 8) for x, c in zip(xs_d, segs)]
    M = nimg * H * W
    outs = []
    for use_tc in (False, True):
        out = torch.zeros(M, cout + 8, dtype=torch.float16, device='cuda')
        pnn.conv2d(cv, seg_list, nimg, H, W, _lib.F16, out, cout + 8, 0, 1, use_tc=use_tc)
        torch.cuda.synchronize()
        outs.append(out[:, :cout].float().cpu())
        assert (out[:, cout:] == 0).all()
    x = torch.cat([x[:, :c].float() for x, c in zip(xs, segs)], 1).view(nimg, H, W, -1).permute(0, 3, 1, 2)
    ref = torch.relu(torch.nn.functional.conv2d(x, w.half().float(), b, padding=(case['kh'] // 2, case['kw'] // 2)))
    ref = ref.permute(0, 2, 3, 1).reshape(M, cout)
    assert (outs[0] - ref).abs().max() < 2e-2
    assert (outs[1] - ref).abs().max() < 2e-2
    assert (outs[1] - outs[0]).abs().max() < 4e-3      # both accumulate in fp32; only the final half rounding differs


def test_gru_module_alone():
    """modules.gru.ConvGRU.forward (reference signature: net, *inputs in NCHW)."""
    from pvo_b200.modules.gru import ConvGRU
    sd = oupd.init_params(0)
    m = ConvGRU(128, 320).cuda()
    m.load_state_dict({k[4:]: v for k, v in sd.items() if k.startswith('gru.')})
    g = torch.Generator().manual_seed(2)
    net = torch.tanh(torch.randn(3, 128, 12, 20, generator=g))
    a, b_, c = torch.randn(3, 128, 12, 20, generator=g), torch.randn(3, 128, 12, 20, generator=g), torch.randn(3, 64, 12, 20, generator=g)
    with torch.no_grad():
        ref = oupd.conv_gru(sd, net, torch.cat([a, b_, c], 1))
        out = m(net.cuda(), a.cuda(), b_.cuda(), c.cuda())
    assert (out.cpu() - ref).abs().max() < 2e-5


@pytest.mark.parametrize('sbo,row_off', [(1024, 0), (1024, 9), (2048, 3), (1280, 0), (1280, 11), (1280, 22), (1152, 5)])
def test_umma_swizzle_is_a_function_of_the_address(sbo, row_off):
    """The hardware fact the halo-reuse convolution rests on (csrc/umma_probe.cu): a SWIZZLE_128B K-major A operand whose
    descriptor starts `row_off` 128-byte rows into a TMA-written patch, with 8-row groups `sbo` bytes apart and
    base_offset = 0, multiplies exactly the rows it addresses -- the swizzle follows the shared-memory address bits."""
    from pvo_b200 import _lib
    from pvo_b200._lib import check, ptr, stream
    g = torch.Generator(device='cuda').manual_seed(sbo + row_off)
    G = torch.randn(512, 64, device='cuda', generator=g).half()
    B = torch.randn(64, 64, device='cuda', generator=g).half()
    D = torch.zeros(128, 64, device='cuda')
    idx = torch.tensor([row_off + (i // 8) * (sbo // 128) + (i % 8) for i in range(128)], device='cuda')
    check(_lib.lib().pvo_umma_probe(ptr(G), ptr(B), 512, row_off, sbo, 0, ptr(D), stream()), 'umma_probe')
    want = G[idx].float() @ B.float().t()
    assert (D - want).abs().max() < 1e-3 * want.abs().max()
