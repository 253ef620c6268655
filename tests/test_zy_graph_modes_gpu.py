"""FactorGraph.update in the two modes the per-frame loop uses besides the plain one -- `use_inactive=True` (with stored edges,
after a keyframe removal, with an explicit t0) and `motion_only=True` on a window of appended frames -- against the
reference's FactorGraph run on CPU (tests/golden/make_golden_graph_modes.py -> graph_modes.npz), float32 path.  Same
observables and tolerances as test_step_gpu.test_graph_update_matches_reference_golden, slightly wider for the later steps.

Passed on its first hardware run (round 2).  The file name sorts after the older test files on purpose."""
import os

import numpy as np
import pytest
import torch

from tests.golden import cases

# the timeout ends the process instead of leaving the box hung should an untried path ever spin on the device
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(420, method='thread')]
G = os.path.join(os.path.dirname(__file__), 'golden')
CASE = dict(F=7, radius=2, h=16, w=24, seed=6, with_segm=True)
EXTRA = 3


def check(tag, v, fg, gold, tol):
    from tests.test_step_gpu import pose_err
    n = gold['poses_' + tag].shape[0]
    assert np.array_equal(fg.ii.cpu().numpy(), gold['ii_' + tag]) and np.array_equal(fg.jj.cpu().numpy(), gold['jj_' + tag]), tag
    assert np.array_equal(fg.ii_inac.cpu().numpy(), gold['ii_inac_' + tag]), tag
    dt, ang = pose_err(v.poses[:n].cpu(), torch.from_numpy(gold['poses_' + tag]))
    assert dt < tol and ang < tol, (tag, dt, ang)
    d, dg = v.disps[:n].cpu().numpy(), gold['disps_' + tag]
    assert (np.abs(d - dg) / np.maximum(np.abs(dg), 1e-3)).max() < tol, tag
    assert np.abs(fg.target_cam.cpu().numpy() - gold['target_' + tag]).max() < 50 * tol, tag
    assert np.abs(fg.weight.cpu().numpy() - gold['weight_' + tag]).mean() < tol, tag


def test_inactive_edges_keyframe_removal_and_motion_only():
    from pvo_b200.factor_graph import FactorGraph
    from tests.test_step_gpu import build_graph
    gold = np.load(os.path.join(G, 'graph_modes.npz'))
    c = cases.graph_case(**CASE)
    F_ = c['F']
    v, fg = build_graph(c, 'fp32', segm_filter=True, thresh=0.45, buffer=F_ + EXTRA + 2)
    fg.update(None, None, itrs=2, use_inactive=True)
    check('a', v, fg, gold, 1e-4)
    fg.rm_factors((fg.ii <= 1) & (fg.jj <= 2), store=True)
    fg.update(None, None, itrs=2, use_inactive=True)
    check('b', v, fg, gold, 2e-4)
    fg.rm_keyframe(F_ - 2)
    v.counter.value = F_ - 1
    fg.update(None, None, itrs=2, use_inactive=True)
    check('c', v, fg, gold, 3e-4)
    fg.update(1, None, itrs=2, use_inactive=True)
    check('d', v, fg, gold, 4e-4)
    # the trajectory filler's call: frames appended behind the keyframes, two edges each, poses-only BA on [n, n + EXTRA)
    n = v.counter.value
    g = torch.Generator().manual_seed(9)
    for k in range(EXTRA):
        v.poses[n + k] = v.poses[n - 1 - k]
        v.disps[n + k] = 1.0
        v.intrinsics[n + k] = v.intrinsics[0]
        v.fmaps[n + k] = torch.randn(128, c['h'], c['w'], generator=g).half().cuda()
    v.counter.value = n + EXTRA
    mo = FactorGraph(v, fg.update_op, device='cuda', precision='fp32')
    new = torch.arange(n, n + EXTRA, device='cuda')
    mo.add_factors(torch.tensor([n - 1, n - 2, n - 3], device='cuda'), new)
    mo.add_factors(torch.tensor([n - 2, n - 3, n - 4], device='cuda'), new)
    for it in range(2):
        mo.update(n, n + EXTRA, motion_only=True)
        check('m%d' % it, v, mo, gold, 5e-4)


@pytest.mark.parametrize('precision', ['fp32', 'fp16'])
def test_use_inactive_replays_as_a_cuda_graph(precision):
    """the frontend's mode (`use_inactive=True`, droid_frontend.py:49,62,81,87) on the captured path: [inactive | active] slab built
    once, no per-iteration cat / unique / host sync.  Three iterations eager == warm + capture + replay on an identical graph; another
    BA call on the device between replays (which rewrites the shared schedule / may grow the workspace) must not disturb the replay."""
    from pvo_b200 import droid_backends
    from tests.test_step_gpu import build_graph, pose_err
    c = cases.graph_case(**(CASE if precision == 'fp32' else dict(CASE, w=40)))      # the tensor-core path wants w = 40 / 80
    F_ = c['F']
    out = []
    for graph in (False, True):
        v, fg = build_graph(c, precision, segm_filter=True, thresh=0.45, buffer=F_ + 4)
        fg.use_cuda_graph = graph
        fg.update(None, None, itrs=2, use_inactive=True)
        fg.rm_factors((fg.ii <= 1) & (fg.jj <= 2), store=True)
        for k in range(4):
            fg.update(None, None, itrs=2, use_inactive=True)
            if graph and k == 2:       # a foreign BA between two replays: different edge set, bigger problem
                big = cases.graph_case(F=9, radius=3, h=16, w=c['w'], seed=3)
                v2, fg2 = build_graph(big, precision, buffer=13)
                fg2.update(itrs=2)
        if graph:
            assert any(not isinstance(g, str) for g in fg._cache['graphs'].values()), 'no graph was captured'
        assert fg._cache['inac']['n'] > 0
        out.append((v.poses[:F_].clone(), v.disps[:F_].clone(), fg.target_cam.clone(), fg.weight.clone()))
    tol = 1e-5 if precision == 'fp32' else 2e-3
    dt, ang = pose_err(out[1][0].cpu(), out[0][0].cpu())
    assert dt < tol and ang < tol, (dt, ang)
    assert ((out[1][1] - out[0][1]).abs() / out[0][1].abs().clamp(min=1e-3)).max().item() < 10 * tol
    assert (out[1][2] - out[0][2]).abs().max().item() < 100 * tol
