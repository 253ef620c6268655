"""GPU parity of the whole hot-loop body: FactorGraph.update (lookup + update operator + post-processing with the
panoptic segment filter + 2 BA iterations) against the reference-pinned golden vectors and the CPU oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import step as ostep, update_op as oupd
from tests.golden import cases

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), 'golden')


def build_graph(c, precision, segm_filter=False, thresh=0.8, buffer=None):
    from pvo_b200.depth_video import DepthVideo
    from pvo_b200.droid_net import DynamicUpdateModule
    from pvo_b200.factor_graph import FactorGraph
    F_, h, w = c['F'], c['h'], c['w']
    v = DepthVideo(image_size=(8 * h, 8 * w), buffer=buffer or F_ + 3, device='cuda', segm_filter=segm_filter, thresh=thresh,
                   store_images=False)
    v.poses[:F_] = c['poses'].cuda()
    v.disps[:F_] = c['disps'].cuda()
    v.intrinsics[:] = c['intrinsics'][0].cuda()
    v.fmaps[:F_] = c['fmaps'].cuda()
    v.nets[:F_] = c['nets'].cuda()
    v.inps[:F_] = c['inps'].cuda()
    if 'segms' in c:
        v.segms[:F_] = c['segms'].cuda()
    v.counter.value = F_
    m = DynamicUpdateModule(precision=precision).cuda()
    m.load_state_dict(oupd.init_params(0))
    fg = FactorGraph(v, m, device='cuda', precision=precision)
    fg.add_factors(c['ii'], c['jj'])
    return v, fg


def pose_err(a, b):
    """translation relative error and rotation angle between (n,7) pose sets (SURVEY.md 8(d) parity observables)."""
    dt = (a[:, :3] - b[:, :3]).norm(dim=-1) / b[:, :3].norm(dim=-1).clamp(min=1e-2)
    qa, qb = a[:, 3:] / a[:, 3:].norm(dim=-1, keepdim=True), b[:, 3:] / b[:, 3:].norm(dim=-1, keepdim=True)
    # angle of qa^-1 qb from its vector part: 2 atan2(|v|, |w|) (acos of the dot product has a ~7e-4 rad noise floor in fp32)
    av, aw, bv, bw = -qa[:, :3], qa[:, 3:], qb[:, :3], qb[:, 3:]
    rv = aw * bv + bw * av + torch.cross(av, bv, dim=-1)
    rw = aw[:, 0] * bw[:, 0] - (av * bv).sum(-1)
    ang = 2 * torch.atan2(rv.double().norm(dim=-1), rw.double().abs())
    return dt.max().item(), ang.max().item()


def test_graph_update_matches_reference_golden():
    """Two iterations with the segment filter ON vs tests/golden/graph_step.npz (reference FactorGraph.update run on
    CPU, float32 update operator).  Tolerances: poses/depth 1e-4 relative, flow EPE 1e-5 (BASELINE.json)."""
    gold = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(G, 'graph_step.npz')).items()}
    c = cases.graph_case(F=5, radius=2, h=16, w=24, seed=4, with_segm=True)
    v, fg = build_graph(c, 'fp32', segm_filter=True, thresh=0.45)
    F_ = c['F']
    for it in range(2):
        fg.update(t0=None, t1=None, itrs=2)
        tag = str(it)
        dt, ang = pose_err(v.poses[:F_].cpu(), gold['poses' + tag])
        assert dt < 1e-4 and ang < 1e-4, (it, dt, ang)
        drel = ((v.disps[:F_].cpu() - gold['disps' + tag]).abs() / gold['disps' + tag].abs().clamp(min=1e-3)).max().item()
        assert drel < 1e-4, (it, drel)
        epe = (fg.full_flow.cpu() - gold['full_flow' + tag]).norm(dim=-1).mean().item()
        assert epe < 1e-5, (it, epe)
        # discrete decisions (bin_mask, segment filter) must agree except on measure-zero ties
        same = ((fg.delta_dy.cpu() == 0) == (gold['delta_dy' + tag] == 0)).float().mean().item()
        assert same > 0.999, (it, same)
        assert (fg.weight.cpu() - gold['weight' + tag]).abs().mean().item() < 1e-4
        assert (fg.raw_mask.cpu() - gold['raw_mask' + tag]).abs().max().item() < 1e-4
        assert (fg.net.float().cpu()[:, :, ::8] - gold['net' + tag]).abs().max().item() < 1e-4
        assert (fg.damping[:F_].cpu() - gold['damping' + tag]).abs().max().item() < 1e-6


@pytest.mark.parametrize('segm', [False, True])
def test_graph_update_matches_oracle_medium(segm):
    """A mid-size graph (6 keyframes, 30x40 -- BASELINE config 1 resolution) vs the CPU oracle, 2 iterations."""
    c = cases.graph_case(F=6, radius=2, h=30, w=40, seed=9, with_segm=segm)
    sd = oupd.init_params(0)
    st = ostep.make_state(c)
    v, fg = build_graph(c, 'fp32', segm_filter=segm, thresh=0.5)
    F_ = c['F']
    with torch.no_grad():
        for it in range(2):
            ostep.graph_update(sd, st, segm_filter=segm, thresh=0.5)
            fg.update(itrs=2)
            dt, ang = pose_err(v.poses[:F_].cpu(), st['poses'][:F_])
            assert dt < 1e-4 and ang < 1e-4, (it, dt, ang)
            drel = ((v.disps[:F_].cpu() - st['disps'][:F_]).abs() / st['disps'][:F_].abs().clamp(min=1e-3)).max().item()
            assert drel < 1e-4, (it, drel)
            epe = (fg.full_flow.cpu() - st['full_flow']).norm(dim=-1).mean().item()
            assert epe < 1e-5, (it, epe)


def test_graph_update_half_tensor_core_path_runs_and_tracks_float32():
    """The production configuration (half state, tensor-core convs): bounded deviation from the float32 oracle."""
    c = cases.graph_case(F=5, radius=2, h=16, w=40, seed=12)
    sd = oupd.init_params(0)
    st = ostep.make_state(c)
    v, fg = build_graph(c, 'fp16')
    with torch.no_grad():
        ostep.graph_update(sd, st)
        fg.update(itrs=2)
    dt, ang = pose_err(v.poses[:5].cpu(), st['poses'][:5])
    assert dt < 2e-2 and ang < 2e-3, (dt, ang)
    epe = (fg.full_flow.cpu() - st['full_flow']).norm(dim=-1).mean().item()
    assert epe < 5e-2, epe


def test_edge_bookkeeping_round_trip():
    """add_factors / rm_factors keep the pixel-major engine state and the reference-layout views in sync."""
    c = cases.graph_case(F=5, radius=2, h=16, w=24, seed=4)
    v, fg = build_graph(c, 'fp32')
    fg.update(itrs=1)
    net_before = fg.net.clone()
    E = fg.ii.shape[0]
    mask = torch.zeros(E, dtype=torch.bool, device='cuda')
    mask[::3] = True
    fg.rm_factors(mask, store=True)
    assert fg.net.shape[1] == E - int(mask.sum()) and fg.corr.num_edges == E - int(mask.sum())
    assert torch.equal(fg.net, net_before[:, ~mask])
    assert fg.ii_inac.shape[0] == int(mask.sum())
    fg.update(itrs=1)
    assert torch.isfinite(v.poses).all() and torch.isfinite(v.disps).all()


def test_add_proximity_factors_selects_the_reference_edges():
    """frame_distance on the device + the host selection == the reference's selection applied to the same distances
    (the selection itself is pinned to the reference on CPU: tests/test_host_logic_cpu.py)"""
    from pvo_b200.depth_video import DepthVideo
    from pvo_b200.droid_net import DynamicUpdateModule
    from pvo_b200.factor_graph import FactorGraph
    from pvo_b200.graph_select import select_proximity_edges
    c = cases.graph_case(F=9, radius=1, h=16, w=24, seed=8)
    v, fg = build_graph(c, 'fp32')
    t = v.counter.value
    before = torch.stack([fg.ii, fg.jj], 1).cpu().tolist()
    ii, jj = torch.meshgrid(torch.arange(0, t), torch.arange(0, t), indexing='ij')
    d = v.distance(ii.reshape(-1), jj.reshape(-1), beta=0.25).cpu().numpy()
    want = select_proximity_edges(d, before, 0, 0, t, rad=2, nms=2, thresh=1e3)
    fg.add_proximity_factors(t0=0, t1=0, rad=2, nms=2, beta=0.25, thresh=1e3)
    have = set(map(tuple, torch.stack([fg.ii, fg.jj], 1).cpu().tolist()))
    assert have == set(map(tuple, before)) | set(want)
    assert fg.corr.num_edges == fg.ii.shape[0] == fg.net.shape[1]
    fg.update(itrs=1)
    assert torch.isfinite(v.poses).all()


def test_update_lowmem_matches_oracle():
    """FactorGraph.update_lowmem (global-BA form: on-the-fly correlation, chunked update operator, one BA over the window
    with lm=1e-5 / ep=1e-2) vs the oracle's restatement of factor_graph.py:309-360.  12 keyframes -> two chunks of source
    frames; features are fp16 in the video buffers as in the reference, arithmetic float32 on both sides."""
    c = cases.graph_case(F=12, radius=2, h=16, w=24, seed=12)
    sd = oupd.init_params(0)
    st = ostep.make_state(c)
    v, fg = build_graph(c, 'fp32')
    F_ = c['F']
    with torch.no_grad():
        ostep.graph_update_lowmem(sd, st, c['fmaps'], c['inps'], F_, itrs=2, steps=2)
        fg.update_lowmem(itrs=2, steps=2)
    dt, ang = pose_err(v.poses[:F_].cpu(), st['poses'][:F_])
    assert dt < 1e-4 and ang < 1e-4, (dt, ang)
    drel = ((v.disps[:F_].cpu() - st['disps'][:F_]).abs() / st['disps'][:F_].abs().clamp(min=1e-3)).max().item()
    assert drel < 1e-4, drel
    assert (fg.target_cam.cpu() - st['target_cam']).norm(dim=-1).mean().item() < 1e-5
    assert (fg.net.float().cpu() - st['net'].float()).abs().max().item() < 2e-3      # hidden state is stored in half between steps
    # the graph keeps working through the regular path afterwards
    fg.update(itrs=1)
    assert torch.isfinite(v.poses).all()
