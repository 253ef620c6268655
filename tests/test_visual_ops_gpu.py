"""droid_backends.iproj / projmap / depth_filter (the visualiser-side entry points, tensor expressions here) against the
reference's own kernels (droid_kernels.cu:406-495, 640-829, compiled into oracle/_ref/droid_ba_ref.so by oracle/build_ref.py) and
against the per-pixel restatements of the oracle.

Passed on its first hardware run (round 2)."""
import pytest
import torch

from oracle import ba_cuda, se3

# the timeout ends the process instead of leaving the box hung should an untried path ever spin on the device
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(420, method='thread')]


def case():
    g = torch.Generator().manual_seed(4)
    F_, h, w = 9, 12, 16
    poses = se3.se3_exp(torch.arange(F_, dtype=torch.float32)[:, None] * torch.tensor([0.06, 0.01, 0.03, 0.0, 0.02, 0.004])
                        + 0.01 * torch.randn(F_, 6, generator=g))
    disps = 0.3 + 1.5 * torch.rand(F_, h, w, generator=g)
    intr = torch.tensor([w / 2.0, w / 2.0, w / 2.0, h / 2.0])
    return poses, disps, intr


def _ref():
    from oracle import build_ref
    try:
        return build_ref.load_ba()
    except Exception:
        return None


def test_against_the_pixel_loops():
    from pvo_b200 import droid_backends as db
    poses, disps, intr = case()
    P, D, K = poses.cuda(), disps.cuda(), intr.cuda()
    want = ba_cuda.iproj_loop(poses, disps, intr)
    assert (db.iproj(P, D, K).cpu() - want).abs().max() <= 1e-5 * want.abs().max()
    ii, jj = torch.tensor([0, 3, 8, 4]), torch.tensor([2, 0, 5, 4])
    c, v = db.projmap(P, D, K, ii.cuda(), jj.cuda())
    cw, vw = ba_cuda.projmap_loop(poses, disps, intr, ii, jj)
    assert torch.equal(v.cpu(), vw) and (c.cpu() - cw).abs().max() < 1e-3
    ix, th = torch.tensor([0, 2, 4, 8]), torch.tensor([0.05, 0.1, 0.2, 0.1])
    cnt = db.depth_filter(P, D, K, ix.cuda(), th.cuda())
    assert (cnt.cpu() != ba_cuda.depth_filter_loop(poses, disps, intr, ix, th)).float().mean() < 0.01


def test_against_the_reference_kernels():
    ref = _ref()
    if ref is None or not hasattr(ref, 'iproj'):
        pytest.skip('oracle/_ref/droid_ba_ref.so (with the visualiser kernels) not present')
    from pvo_b200 import droid_backends as db
    poses, disps, intr = case()
    P, D, K = poses.cuda(), disps.cuda(), intr.cuda()
    a, b = db.iproj(P, D, K), ref.iproj(P, D, K)
    assert a.shape == b.shape and (a - b).abs().max() <= 1e-5 * b.abs().max()
    ii, jj = torch.tensor([0, 3, 8, 4]).cuda(), torch.tensor([2, 0, 5, 4]).cuda()
    (c, v), (cr, vr) = db.projmap(P, D, K, ii, jj), ref.projmap(P, D, K, ii, jj)
    assert c.shape == cr.shape and torch.equal(v, vr) and (c - cr).abs().max() < 1e-3
    ix, th = torch.tensor([0, 2, 4, 8]).cuda(), torch.tensor([0.05, 0.1, 0.2, 0.1]).cuda()
    cnt, cntr = db.depth_filter(P, D, K, ix, th), ref.depth_filter(P, D, K, ix, th)
    assert cnt.shape == cntr.shape and (cnt != cntr).float().mean() < 0.01
