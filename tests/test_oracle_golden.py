"""The CPU oracle against the golden vectors produced by the REFERENCE's own code
(tests/golden/make_golden.py) -- this is what pins the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import geom, se3, step as ostep, update_op as oupd
from tests.golden import cases

G = os.path.join(os.path.dirname(__file__), 'golden')


def _load(name):
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(G, name)).items()}


def test_update_op_matches_reference_module():
    gold = _load('update_op.npz')
    sd = oupd.init_params(0)
    net, inp, corr, flow = cases.update_inputs(E=3, h=12, w=16, seed=0)
    with torch.no_grad():
        o_net, o_delta, o_weight, o_eta, o_up, o_dm = oupd.update_op(sd, net, inp, corr, flow, gold['ii'])
    for name, val in (('net', o_net), ('delta', o_delta), ('weight', o_weight), ('eta', o_eta),
                      ('upmask', o_up[:, :, ::16]), ('delta_m', o_dm)):
        assert torch.allclose(val, gold[name], atol=1e-6, rtol=1e-6), name


def test_geometry_and_python_ba_match_reference():
    gold = _load('geom.npz')
    c = cases.graph_case(F=5, radius=2, h=12, w=16, seed=3, fdtype=torch.float32)
    ii, jj = c['ii'], c['jj']
    poses, disps, intr = c['poses'][None], c['disps'][None], c['intrinsics'][None]
    x1, valid, (Ji, Jj, Jz) = geom.projective_transform(poses, disps, intr, ii, jj, jacobian=True)
    for name, val in (('coords', x1), ('valid', valid), ('Ji', Ji), ('Jj', Jj), ('Jz', Jz)):
        assert torch.allclose(val, gold[name], atol=1e-6, rtol=1e-6), name
    p1, d1 = geom.BA(gold['target'], gold['weight'], gold['eta'], poses, disps, intr, ii, jj, fixedp=2)
    p2, d2 = geom.BA(gold['target'], gold['weight'], gold['eta'], p1, d1, intr, ii, jj, fixedp=2)
    pm = geom.MoBA(gold['target'], gold['weight'], gold['eta'], poses, disps, intr, ii, jj, fixedp=1)
    assert torch.allclose(p1, gold['ba1_poses'], atol=1e-6)
    assert torch.allclose(d1, gold['ba1_disps'], atol=1e-6)
    assert torch.allclose(p2, gold['ba2_poses'], atol=1e-6)
    assert torch.allclose(d2, gold['ba2_disps'], atol=1e-6)
    assert torch.allclose(pm, gold['moba_poses'], atol=1e-6)


def test_graph_step_matches_reference_factor_graph():
    gold = _load('graph_step.npz')
    c = cases.graph_case(F=5, radius=2, h=16, w=24, seed=4, with_segm=True)
    sd = oupd.init_params(0)
    st = ostep.make_state(c)
    with torch.no_grad():
        for it in range(2):
            ostep.graph_update(sd, st, segm_filter=True, thresh=0.45)
            F_ = c['F']
            tag = str(it)
            assert torch.allclose(st['poses'][:F_], gold['poses' + tag], atol=1e-6), it
            assert torch.allclose(st['disps'][:F_], gold['disps' + tag], atol=1e-6), it
            assert torch.allclose(st['target_cam'], gold['target' + tag], atol=1e-5), it
            assert torch.allclose(st['weight'], gold['weight' + tag], atol=1e-6), it
            assert torch.allclose(st['delta_dy'], gold['delta_dy' + tag], atol=1e-6), it
            assert torch.allclose(st['raw_mask'], gold['raw_mask' + tag], atol=1e-6), it
            assert torch.allclose(st['full_flow'], gold['full_flow' + tag], atol=1e-5), it
            assert torch.allclose(st['net'][:, :, ::8], gold['net' + tag], atol=1e-6), it
            assert torch.allclose(st['damping'][:F_], gold['damping' + tag], atol=1e-7), it
    # the segment filter must actually have fired on this case, otherwise the test pins nothing
    assert (st['delta_dy'] != 0).any() and (st['weight'] > 0.999).any()


def test_encoder_oracle_matches_reference_droidnet_features():
    """oracle.encoder vs the reference DroidNet.extract_features (fnet instance-norm, cnet no-norm), droid_net.npz."""
    from oracle import encoder as oenc
    from tests.golden.make_golden import droid_net_inputs, droid_net_state
    gold = _load('droid_net.npz')
    sd = droid_net_state()
    images, *_ = droid_net_inputs()
    sf = {k[5:]: v for k, v in sd.items() if k.startswith('fnet.')}
    sc = {k[5:]: v for k, v in sd.items() if k.startswith('cnet.')}
    with torch.no_grad():
        fm, net, inp = oenc.extract_features(sf, sc, images)
    assert torch.allclose(fm[:, :, ::8], gold['fmaps'], atol=1e-5)
    assert torch.allclose(net[:, :, ::8], gold['net'], atol=1e-5)
    assert torch.allclose(inp[:, :, ::8], gold['inp'], atol=1e-5)


def test_oracle_lowmem_matches_reference_golden():
    """oracle.step.graph_update_lowmem vs the reference's FactorGraph.update_lowmem run on CPU (tests/golden/make_golden_lowmem.py):
    the global-BA control flow (chunks of source frames, per-chunk state updates, one BA per step with lm=1e-5, ep=1e-2)."""
    import os
    import numpy as np
    from oracle import step as ostep, update_op as oupd
    from tests.golden import cases
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'lowmem.npz'))
    c = cases.graph_case(F=12, radius=2, h=16, w=24, seed=12)
    st = ostep.make_state(c)
    with torch.no_grad():
        ostep.graph_update_lowmem(oupd.init_params(0), st, c['fmaps'], c['inps'], 12, itrs=2, steps=2)
    T = lambda k: torch.from_numpy(z[k])
    assert (st['poses'][:12] - T('poses')).abs().max() < 1e-5
    assert ((st['disps'][:12] - T('disps')).abs() / T('disps').abs().clamp(min=1e-3)).max() < 1e-4
    assert (st['target_cam'] - T('target')).abs().max() < 1e-3
    assert (st['weight'] - T('weight')).abs().max() < 1e-4
    assert (st['delta_dy'] - T('delta_dy')).abs().max() < 1e-3
    assert (st['raw_mask'] - T('raw_mask')).abs().max() < 1e-3
    assert (st['net'].float()[:, :, ::8] - T('net')).abs().max() < 2e-3
    assert (st['damping'][:12] - T('damping')).abs().max() < 1e-5
