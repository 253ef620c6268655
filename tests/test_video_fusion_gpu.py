"""The video-fusion branch of Panoptic-FPN on the B200 kernels vs the fp32 oracle (oracle.panoptic_fpn.inference_fusion, itself
pinned to the reference's own inference_fusion by tests/golden/fpn_e2e.npz): flow transport (index work: bit-exact), the
512 -> 256 fusion conv on the tensor cores (relative L2 like the other conv stages), and the two-frame protocol."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import panoptic_fpn as ofpn

pytestmark = pytest.mark.gpu

H, W = 128, 192


def fusion_inputs():
    g = torch.Generator().manual_seed(41)
    a = torch.randint(0, 256, (3, H, W), generator=g).float()
    b = torch.randint(0, 256, (3, H, W), generator=g).float()
    coarse = torch.rand(1, 2, 5, 7, generator=g) * 6.0
    flow = F.interpolate(coarse, size=(H, W), mode='bilinear', align_corners=True)[0]
    flow[:, 20:50, 30:70] -= 7.5
    flow[0, 80:110, 150:] += 60.0
    return a, b, flow


@pytest.fixture(scope='module')
def setup():
    from pvo_b200.panoptic_fpn import PanopticFPN
    sd = ofpn.init_params(0, fusion=True)
    a, b, flow = fusion_inputs()
    dbg = {}
    with torch.no_grad():
        ref1 = ofpn.inference(sd, [a])
        ref2 = ofpn.inference_fusion(sd, a, flow, b, debug=dbg)
    return sd, (a, b, flow), (ref1, ref2), dbg, PanopticFPN(sd, device='cuda')


def test_flow_transport_on_the_gpu_is_bit_exact():
    from pvo_b200.panoptic_fpn import flow_transport, flow_transport_index
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'fpn_e2e.npz'))
    flow = fusion_inputs()[2]
    g = torch.Generator().manual_seed(42)
    for k, (h, w) in {'p2': (32, 48), 'p4': (8, 12), 'p6': (2, 3)}.items():
        feat = torch.randn(2, 16, h, w, generator=g)
        win = flow_transport_index(flow.cuda(), h, w)
        assert torch.equal(win.cpu(), flow_transport_index(flow, h, w)), k
        got = flow_transport(feat[0].permute(1, 2, 0).reshape(h * w, 16).cuda(), win)
        assert np.array_equal(got.view(h, w, 16).permute(2, 0, 1).cpu().numpy(), z['transport_' + k]), k


def test_fused_pyramid(setup):
    sd, (a, b, flow), _, dbg, m = setup
    with torch.no_grad():
        fused, sizes = m.fuse_pair(a.cuda(), flow.cuda(), b.cuda())
    assert sizes == [(H, W)]
    for k in ('p2', 'p3', 'p4', 'p5', 'p6'):
        t, h, w = fused[k]
        got = t.view(1, h, w, 256).permute(0, 3, 1, 2).float().cpu()
        want = dbg['feats'][k]
        e = ((got - want).norm() / want.norm()).item()
        assert e < 3e-2, (k, e)


def test_two_frame_protocol(setup):
    sd, (a, b, flow), (ref1, ref2), dbg, m = setup
    out1 = m.inference_fusion([{'image': a.cuda(), 'image_id': '0001_00000', 'flow': flow.cuda()}])
    out2 = m.inference_fusion([{'image': b.cuda(), 'image_id': '0001_00001', 'flow': flow.cuda()}])
    for o, r in ((out1, ref1), (out2, ref2)):
        pan = o[0]['panoptic_seg'][0]
        assert pan.shape == (H, W) and pan.dtype == torch.int32
        agree = ((pan.cpu() // 10000) == (r[0]['panoptic_seg'][0] // 10000)).float().mean().item()
        assert agree > 0.9, agree
        assert (o[0]['sem_seg'].argmax(0).cpu() == r[0]['sem_seg'].argmax(0)).float().mean() > 0.97
    # a new video id goes back to the single-image path: same result as frame 1
    out3 = m.inference_fusion([{'image': a.cuda(), 'image_id': '0002_00000', 'flow': flow.cuda()}])
    assert torch.equal(out3[0]['panoptic_seg'][0], out1[0]['panoptic_seg'][0])


def test_fusion_needs_the_fusion_weights():
    from pvo_b200.panoptic_fpn import PanopticFPN
    m = PanopticFPN(ofpn.init_params(0), device='cuda')
    a, b, flow = fusion_inputs()
    with pytest.raises(RuntimeError):
        m.fuse_pair(a.cuda(), flow.cuda(), b.cuda())
