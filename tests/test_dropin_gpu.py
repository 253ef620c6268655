"""The drop-in proof on the GPU: the reference's OWN Python -- factor_graph.FactorGraph,
depth_video.DepthVideo, modules/corr.CorrBlock + CorrSampler, droid_net.DynamicUpdateModule, geom/projective_ops -- staged as
bytecode from /root/reference (oracle/ref_stage.py; nothing of it is modified) runs on `pvo_b200.droid_backends` and
`pvo_b200.lietorch` in place of its compiled extensions, on the B200, and reproduces the golden vectors its CPU run produced
(tests/golden/graph_step.npz).  This is INTEGRATION.md's substitution, executed."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_stage, update_op as oupd
from tests import ref_harness as rh
from tests.golden import cases

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_stage.available(), reason='oracle/_ref/py not staged')]
G = os.path.join(os.path.dirname(__file__), 'golden')


@pytest.fixture(autouse=True)
def _no_tf32():
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b
    ref_stage.force_fp32(False)


def test_reference_factor_graph_unmodified_on_pvo_backends_matches_its_golden():
    """two FactorGraph.update steps with the panoptic segment filter on, float32 (the precision of the golden run): the
    reference's update() reaches this package through droid_backends.corr_index_forward (4 calls), droid_backends.ba and
    lietorch.SE3 (reproject)."""
    gold = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(G, 'graph_step.npz')).items()}
    c = cases.graph_case(F=5, radius=2, h=16, w=24, seed=4, with_segm=True)
    video, fg, _ = rh.reference_graph(c, oupd.init_params(0), fp32=True, segm_filter=True, thresh=0.45, buffer=8)
    import droid_backends
    import lietorch
    import pvo_b200.droid_backends
    import pvo_b200.lietorch
    assert droid_backends is pvo_b200.droid_backends and lietorch is pvo_b200.lietorch
    assert type(fg).__module__ == 'factor_graph' and type(video).__module__ == 'depth_video'
    F_ = c['F']
    with torch.no_grad():
        for it in range(2):
            fg.update(t0=None, t1=None, itrs=2, use_inactive=False)
            tag = str(it)
            dt, ang = rh.pose_err(video.poses[:F_].cpu(), gold['poses' + tag])
            assert dt < 1e-4 and ang < 1e-4, (it, dt, ang)
            drel = ((video.disps[:F_].cpu() - gold['disps' + tag]).abs() / gold['disps' + tag].abs().clamp(min=1e-3)).max().item()
            assert drel < 1e-4, (it, drel)
            epe = (fg.full_flow.cpu() - gold['full_flow' + tag]).norm(dim=-1).mean().item()
            assert epe < 1e-5, (it, epe)
            assert (fg.weight.cpu() - gold['weight' + tag]).abs().mean().item() < 1e-4
            assert (fg.damping[:F_].cpu() - gold['damping' + tag]).abs().max().item() < 1e-6


def test_reference_graph_and_this_package_agree_in_the_reference_inference_mode():
    """the reference's native inference mode (autocast) on our backends next to this package's half path, same inputs: both are
    half-precision runs of the same computation, so they agree to half-level tolerances (and the reference's half CorrBlock
    volume goes through our lookup kernel's half branch)."""
    c = cases.graph_case(F=5, radius=2, h=16, w=40, seed=12)
    sd = oupd.init_params(0)
    with torch.no_grad():
        vr, gr, _ = rh.reference_graph(c, sd, fp32=False)
        gr.update(itrs=2)
        vo, go = rh.our_graph(c, sd, 'fp16')
        go.update(itrs=2)
    assert gr.corr.corr_pyramid[0].dtype == torch.float16
    dt, ang = rh.pose_err(vo.poses[:5].cpu(), vr.poses[:5].cpu())
    assert dt < 2e-2 and ang < 2e-3, (dt, ang)
    assert (go.full_flow - gr.full_flow).norm(dim=-1).mean().item() < 5e-2


def test_reference_corrblock_and_altcorr_on_pvo_backends():
    """reference CorrBlock / AltCorrBlock objects (their own volume build on cuBLAS, their own autograd Functions) calling into
    pvo_b200.droid_backends == this package's CorrBlock / AltCorrBlock"""
    ref = ref_stage.load()
    from pvo_b200.modules.corr import CorrBlock, AltCorrBlock
    g = torch.Generator().manual_seed(7)
    F_, h, w = 4, 16, 24
    fm = torch.randn(1, F_, 128, h, w, generator=g).cuda()
    ii = torch.tensor([0, 1, 2, 3, 1]).cuda()
    jj = torch.tensor([1, 0, 3, 2, 3]).cuda()
    y, x = torch.meshgrid(torch.arange(h).float(), torch.arange(w).float(), indexing='ij')
    coords = (torch.stack([x, y], -1)[None, None] + 4 * (torch.rand(1, 5, h, w, 2, generator=g) * 2 - 1)).cuda()
    a = ref.corr.CorrBlock(fm[:, ii], fm[:, jj])(coords)
    b = CorrBlock(fm[:, ii], fm[:, jj])(coords)
    assert a.shape == b.shape and torch.allclose(a, b, atol=2e-4, rtol=1e-4)
    a2 = ref.corr.AltCorrBlock(fm)(coords, ii, jj)
    b2 = AltCorrBlock(fm)(coords, ii, jj)
    assert a2.shape == b2.shape and torch.allclose(a2.float(), b2.float(), atol=2e-4, rtol=1e-4)
    assert torch.allclose(a2.float(), a, atol=5e-4, rtol=1e-3)          # on-the-fly == materialised volume
