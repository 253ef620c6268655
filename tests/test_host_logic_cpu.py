"""Host-side logic that needs no GPU: weight packing layouts, parameter names, oracle self-consistency
(lietorch's own property tests restated, CUDA-BA restatement vs the reference's Python BA)."""
import pytest
import torch

from oracle import ba_cuda, geom, se3, update_op as oupd, corr as ocorr
from pvo_b200 import nn as pnn
from pvo_b200.droid_net import DynamicUpdateModule
from tests.golden import cases


def test_pack_conv_weight_layouts():
    g = torch.Generator().manual_seed(0)
    w = torch.randn(6, 5, 3, 3, generator=g)
    W = pnn.pack_conv_weight(w, cin_pad=8)
    assert W.shape == (9 * 8, 8)
    for co in range(6):
        for ci in range(5):
            for kh in range(3):
                for kw in range(3):
                    assert W[(kh * 3 + kw) * 8 + ci, co] == w[co, ci, kh, kw]
    assert W[:, 6:].abs().sum() == 0 and W.reshape(9, 8, 8)[:, 5:].abs().sum() == 0
    w = torch.randn(4, 70 + 10, 1, 3, generator=g)
    Wh = pnn.pack_conv_weight_tc(w, segs=[70, 10])
    assert Wh.shape == (4, 3 * (128 + 64)) and Wh.dtype == torch.float16
    Wr = Wh.reshape(4, 3, 192)
    assert torch.equal(Wr[:, :, :70], w[:, :70].permute(0, 2, 3, 1).reshape(4, 3, 70).half())
    assert torch.equal(Wr[:, :, 128:138], w[:, 70:].permute(0, 2, 3, 1).reshape(4, 3, 10).half())
    assert Wr[:, :, 70:128].abs().sum() == 0 and Wr[:, :, 138:].abs().sum() == 0
    ws = pnn.pack_small_weight([torch.randn(2, 128, 3, 3, generator=g) for _ in range(4)])
    assert ws.shape == (4, 9, 128, 2)


def test_module_state_dict_matches_reference_names():
    """checkpoint-compatibility contract of DroidNet.update."""
    m = DynamicUpdateModule()
    shapes = oupd.param_shapes()
    sd = m.state_dict()
    assert sorted(sd) == sorted(shapes)
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(shapes[k]), k
    assert sum(v.numel() for v in sd.values()) == 2881801


# ---- lietorch/run_tests.py:16-52 restated on the oracle's SE3 (float64, atol 1e-8) ----------------------------
def test_se3_exp_log():
    a = .2 * torch.randn(2, 3, 4, 6, dtype=torch.float64)
    assert torch.allclose(se3.se3_log(se3.se3_exp(a)), a, atol=1e-8)


def test_se3_inverse():
    X = se3.se3_exp(.1 * torch.randn(2, 3, 6, dtype=torch.float64))
    a = se3.se3_log(se3.se3_mul(X, se3.se3_inv(X)))
    assert torch.allclose(a, torch.zeros_like(a), atol=1e-8)


def test_se3_adjoint():
    X = se3.se3_exp(torch.randn(2, 3, 6, dtype=torch.float64))
    a = torch.randn(2, 3, 6, dtype=torch.float64)
    Y1 = se3.se3_mul(X, se3.se3_exp(a))
    Y2 = se3.se3_mul(se3.se3_exp(se3.se3_adj(X, a)), X)
    c = se3.se3_log(se3.se3_mul(Y1, se3.se3_inv(Y2)))
    assert torch.allclose(c, torch.zeros_like(c), atol=1e-8)
    # adjT is the transpose of adj
    b = torch.randn(2, 3, 6, dtype=torch.float64)
    assert torch.allclose((se3.se3_adj(X, a) * b).sum(-1), (a * se3.se3_adjT(X, b)).sum(-1), atol=1e-8)


def test_se3_act_matches_matrix():
    X = se3.se3_exp(torch.randn(5, 6, dtype=torch.float64))
    p = torch.randn(5, 3, dtype=torch.float64)
    ph = torch.cat([p, torch.ones(5, 1, dtype=torch.float64)], -1)
    assert torch.allclose(se3.se3_act3(X, p), (se3.se3_matrix(X) @ ph[..., None])[..., :3, 0], atol=1e-8)


def test_cuda_ba_restatement_agrees_with_python_ba_where_semantics_coincide():
    """SURVEY.md 8(c) self-validation (2): same reduced system when all depths are > 0.25, eta shifted by 1e-7,
    lm = 0 (the two implementations damp different matrices, 8(a) addendum)."""
    c = cases.graph_case(F=5, radius=2, h=12, w=16, seed=7, fdtype=torch.float32)
    g = torch.Generator().manual_seed(70)
    ii, jj = c['ii'], c['jj']
    poses, disps, intr = c['poses'], c['disps'].clamp(min=0.5), c['intrinsics']
    tgt, _ = geom.projective_transform(poses[None], disps[None], intr[None], ii, jj)
    tgt = tgt + 0.5 * torch.randn(tgt.shape, generator=g)
    wgt = torch.rand(tgt.shape, generator=g)
    eta = 0.01 * torch.rand(torch.unique(ii).shape[0], 12, 16, generator=g)
    _, _, dbg = geom.BA(tgt, wgt, eta[None], poses[None], disps[None], intr[None], ii, jj, fixedp=2, return_debug=True)
    rec = []
    T = tgt[0].permute(0, 3, 1, 2).contiguous()
    W = wgt[0].permute(0, 3, 1, 2).contiguous()
    ba_cuda.ba(poses.clone(), disps.clone(), intr[0], T, W, eta + 1e-7, ii, jj, 2, 5, 1, 0.0, 0.1, False, debug=rec)
    S_py = dbg['S'][0].double()
    A_cu = rec[0]['A']
    off = (S_py - A_cu) - torch.diag(torch.diagonal(S_py - A_cu))
    assert off.abs().max() < 1e-6 * A_cu.abs().max()
    assert (dbg['v'][0, :, 0].double() - rec[0]['b']).abs().max() < 1e-5 * rec[0]['b'].abs().max()


def test_cuda_ba_restatement_jacobians_by_finite_differences():
    """SURVEY.md 8(c) self-validation (3): K5's Jj / Jz against numeric derivatives of the projection."""
    torch.manual_seed(3)
    c = cases.graph_case(F=3, radius=1, h=6, w=8, seed=9, fdtype=torch.float32)
    ii, jj = c['ii'][:1], c['jj'][:1]
    poses, disps, intr = c['poses'].double(), c['disps'].double().clamp(min=0.5), c['intrinsics'][0].double()
    z = torch.zeros(1, 2, 6, 8, dtype=torch.float64)

    def proj(poses, disps):
        b = ba_cuda.edge_blocks(poses.float(), disps.float(), intr.float(), z.float(), z.float() + 1, ii, jj)
        return -b['ru'].double(), -b['rv'].double()   # target = 0 -> r = -(projection)

    b0 = ba_cuda.edge_blocks(poses.float(), disps.float(), intr.float(), z.float(), z.float() + 1, ii, jj)
    eps = 1e-3
    d2 = disps.clone()
    d2[ii[0]] += eps
    u1, v1 = proj(poses, d2)
    u0, v0 = proj(poses, disps)
    assert torch.allclose((u1 - u0) / eps, b0['Jzu'].double(), atol=2e-2, rtol=2e-2)
    assert torch.allclose((v1 - v0) / eps, b0['Jzv'].double(), atol=2e-2, rtol=2e-2)
    for k in range(6):
        xi = torch.zeros(6, dtype=torch.float64)
        xi[k] = eps
        p2 = poses.clone()
        p2[jj[0]] = ba_cuda.retr_se3(xi.float()[None], poses[jj[0]].float()[None])[0].double()
        u1, v1 = proj(p2, disps)
        assert torch.allclose((u1 - u0) / eps, b0['Jju'][..., k].double(), atol=5e-2, rtol=5e-2), k
        assert torch.allclose((v1 - v0) / eps, b0['Jjv'][..., k].double(), atol=5e-2, rtol=5e-2), k


def test_oracle_lookup_half_matches_float_on_exact_inputs():
    """With integer coordinates the bilinear weights are exactly 0/1, so the half and float models must agree."""
    g = torch.Generator().manual_seed(1)
    vol = torch.randint(-8, 8, (2, 5, 6, 9, 11), generator=g).float()
    coords = torch.stack([torch.randint(-2, 13, (2, 5, 6), generator=g), torch.randint(-2, 11, (2, 5, 6), generator=g)], 1).float()
    a = ocorr.corr_index_forward(vol, coords, 3)
    b = ocorr.corr_index_forward(vol.half(), coords, 3)
    assert torch.equal(a, b.float())
    # window centre equals the volume value at the (in-bounds) coordinate
    for n, y, x in ((0, 1, 2), (1, 4, 5)):
        cx, cy = int(coords[n, 0, y, x]), int(coords[n, 1, y, x])
        if 0 <= cx < 11 and 0 <= cy < 9:
            assert a[n, 3, 3, y, x] == vol[n, y, x, cy, cx]


def test_proximity_edge_selection_matches_reference_golden():
    """pvo_b200.graph_select vs the reference's FactorGraph.add_proximity_factors run on prescribed distances
    (tests/golden/make_golden_proximity.py): identical edge lists, in order."""
    import os
    import numpy as np
    from pvo_b200.graph_select import select_proximity_edges
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'proximity.npz'))
    k = 0
    while 'd%d' % k in z:
        t, t0, t1, rad, nms, thresh, ne = z['cfg%d' % k]
        es = select_proximity_edges(z['d%d' % k], z['ex%d' % k].tolist(), int(t0), int(t1), int(t), rad=int(rad), nms=int(nms),
                                    thresh=float(thresh))
        assert np.array_equal(np.asarray(es, dtype=np.int64).reshape(-1, 2), z['es%d' % k]), k
        k += 1
    assert k == 6


def test_segments_for_vo_matches_the_reference_resize_chain():
    """pvo_b200.panoptic_fpn.segments_for_vo == evaluation_scripts/test_vo.py:44-50 (interpolate -> crop -> interpolate 1/8)"""
    import torch.nn.functional as F
    from pvo_b200.panoptic_fpn import segments_for_vo
    g = torch.Generator().manual_seed(0)
    for (H, W), size in (((375, 1242), (240, 808)), ((480, 640), None), ((376, 1241), (243, 811))):
        pan = torch.randint(0, 150000, (H, W), generator=g, dtype=torch.int32)
        h1, w1 = size if size is not None else (H, W)
        segm = pan[None, None].float()
        segm = F.interpolate(segm, size=(h1, w1))
        segm = segm[:h1 - h1 % 8, :w1 - w1 % 8]                # as written in the reference: slices dims 0, 1 of a 4-D tensor (no-op)
        want = F.interpolate(segm, scale_factor=1 / 8, recompute_scale_factor=True).int()
        got = segments_for_vo(pan, size)
        assert got.shape == want.shape and torch.equal(got, want), (H, W, size)


def test_flow_transport_index_matches_the_reference_on_cpu_tensors():
    """pvo_b200.panoptic_fpn.flow_transport_index / flow_transport are plain tensor ops: run them on CPU tensors against the
    golden output of the reference's flow_transport_feature (tests/golden/make_golden_fpn_e2e.py), bit for bit"""
    import os
    import numpy as np
    import torch.nn.functional as F
    from pvo_b200.panoptic_fpn import flow_transport, flow_transport_index
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'fpn_e2e.npz'))
    g = torch.Generator().manual_seed(41)
    torch.randint(0, 256, (3, 128, 192), generator=g), torch.randint(0, 256, (3, 128, 192), generator=g)    # the two frames
    coarse = torch.rand(1, 2, 5, 7, generator=g) * 6.0
    flow = F.interpolate(coarse, size=(128, 192), mode='bilinear', align_corners=True)[0]
    flow[:, 20:50, 30:70] -= 7.5
    flow[0, 80:110, 150:] += 60.0
    g = torch.Generator().manual_seed(42)
    for k, (h, w) in {'p2': (32, 48), 'p4': (8, 12), 'p6': (2, 3)}.items():
        feat = torch.randn(2, 16, h, w, generator=g)
        win = flow_transport_index(flow, h, w)
        assert win.shape == (h * w,) and win.dtype == torch.int64 and int(win.max()) < h * w
        got = flow_transport(feat[0].permute(1, 2, 0).reshape(h * w, 16), win)             # pixel-major, as on the GPU
        assert np.array_equal(got.view(h, w, 16).permute(2, 0, 1).numpy(), z['transport_' + k]), k
    # nothing lands anywhere: flow entirely negative
    assert int(flow_transport_index(torch.full((2, 16, 16), -3.0), 4, 4).max()) == -1


def test_fuse_pair_glue_with_stand_in_kernels():
    """PanopticFPN.fuse_pair's tensor plumbing (pair order, pixel-major views, channel order of the 512-channel fused input,
    per-level flow resize) with the two CUDA stages replaced by fp32 torch stand-ins IN THIS TEST ONLY: the fused pyramid
    must equal the oracle's.  (The real kernels are checked on the GPU in tests/test_video_fusion_gpu.py.)"""
    import torch.nn.functional as F
    from oracle import panoptic_fpn as ofpn
    from pvo_b200.panoptic_fpn import PanopticFPN
    sd = ofpn.init_params(0, fusion=True)
    g = torch.Generator().manual_seed(9)
    a, b = (torch.randint(0, 256, (3, 64, 96), generator=g).float() for _ in range(2))
    flow = torch.rand(2, 64, 96, generator=g) * 5 - 1
    m = PanopticFPN.__new__(PanopticFPN)
    m.device, m.fusion, m.alpha = torch.device('cpu'), True, 1.0
    m.mean = torch.tensor(ofpn.PIXEL_MEAN).view(1, 1, 1, 3)

    def backbone(x):                                        # stand-in: the oracle's backbone, handed over pixel-major
        feats = ofpn.backbone(sd, x.float().permute(0, 3, 1, 2))
        return {k: (v.permute(0, 2, 3, 1).reshape(-1, 256), v.shape[2], v.shape[3]) for k, v in feats.items()}

    def conv(name, src, N, H, W, **kw):                     # stand-in for the tensor-core conv: pixel-major in and out
        x = src.view(N, H, W, -1).permute(0, 3, 1, 2)
        return F.conv2d(x, sd[name + '.weight'], sd[name + '.bias'], padding=1).permute(0, 2, 3, 1).reshape(N * H * W, -1)
    m.backbone, m._conv = backbone, conv
    with torch.no_grad():
        fused, sizes = m.fuse_pair(a, flow, b)
        dbg = {}
        ofpn.inference_fusion(sd, a, flow, b, debug=dbg)
    assert sizes == [(64, 96)] and m.N == 1
    for k, (t, h, w) in fused.items():
        want = dbg['feats'][k]
        assert (h, w) == tuple(want.shape[-2:])
        got = t.view(1, h, w, 256).permute(0, 3, 1, 2)
        assert (got - want).abs().max() <= 2e-3 * want.abs().max(), k      # half rounding of the normalised image only


def test_product_se3_passes_lietorch_property_tests():
    """thirdparty/lietorch/run_tests.py:16-52 (exp-log, inverse, adjoint, action vs matrix; float64, atol 1e-8) on the SE3
    class the package ships (pvo_b200.lietorch), plus agreement with the oracle's SE3"""
    from pvo_b200.lietorch import SE3
    g = torch.Generator().manual_seed(0)
    a = .2 * torch.randn(2, 3, 4, 6, dtype=torch.float64, generator=g)
    assert torch.allclose(SE3.exp(a).log(), a, atol=1e-8)
    big = torch.randn(64, 6, dtype=torch.float64, generator=g) * torch.tensor([1, 1, 1, 1.7, 1.7, 1.7], dtype=torch.float64)
    keep = big[:, 3:].norm(dim=-1) < 3.0                                    # principal branch
    assert torch.allclose(SE3.exp(big[keep]).log(), big[keep], atol=1e-8)
    tiny = 1e-9 * torch.randn(8, 6, dtype=torch.float64, generator=g)
    assert torch.allclose(SE3.exp(tiny).log(), tiny, atol=1e-12)
    X = SE3.exp(.1 * torch.randn(2, 3, 6, dtype=torch.float64, generator=g))
    assert torch.allclose((X * X.inv()).log(), torch.zeros(2, 3, 6, dtype=torch.float64), atol=1e-8)
    X = SE3.exp(torch.randn(2, 3, 6, dtype=torch.float64, generator=g))
    a = torch.randn(2, 3, 6, dtype=torch.float64, generator=g)
    Y1, Y2 = X * SE3.exp(a), SE3.exp(X.adj(a)) * X
    assert torch.allclose((Y1 * Y2.inv()).log(), torch.zeros_like(a), atol=1e-8)
    b = torch.randn(2, 3, 6, dtype=torch.float64, generator=g)
    assert torch.allclose((X.adj(a) * b).sum(-1), (a * X.adjT(b)).sum(-1), atol=1e-8)
    p = torch.randn(2, 3, 3, dtype=torch.float64, generator=g)
    ph = torch.cat([p, torch.ones(2, 3, 1, dtype=torch.float64)], -1)
    assert torch.allclose(X * p, (X.matrix() @ ph[..., None])[..., :3, 0], atol=1e-8)
    assert torch.allclose((X * ph)[..., :3], X * p, atol=1e-12)
    assert torch.allclose(X.log(), se3.se3_log(X.data), atol=1e-10) and torch.allclose(X.adjT(b), se3.se3_adjT(X.data, b), atol=1e-10)
    assert torch.allclose(X.retr(a).data, se3.se3_retr(X.data, a), atol=1e-10)


def test_trajectory_filler_interpolation_on_cpu_tensors():
    """PoseTrajectoryFiller.interpolate (constant velocity between the enclosing keyframes, trajectory_filler.py:42-57) is
    tensor code: run it on a CPU video stub against the same formulas written with the oracle's SE3, frame by frame"""
    import types
    from pvo_b200.trajectory_filler import PoseTrajectoryFiller
    g = torch.Generator().manual_seed(3)
    N = 5
    poses = se3.se3_exp(0.3 * torch.randn(N, 6, generator=g))
    ts = torch.tensor([0.0, 2.0, 3.0, 7.0, 8.0])
    video = types.SimpleNamespace(counter=types.SimpleNamespace(value=N), tstamp=torch.cat([ts, torch.zeros(3)]),
                                  poses=torch.cat([poses, torch.zeros(3, 7)]))
    f = PoseTrajectoryFiller.__new__(PoseTrajectoryFiller)
    f.video, f.device = video, 'cpu'
    tstamps = [0.0, 1.0, 2.0, 2.5, 6.0, 7.0, 7.9, 8.0, 9.5]
    Gs, t0, t1 = f.interpolate(tstamps)
    assert t0.tolist() == [0, 0, 1, 1, 2, 3, 3, 4, 4] and t1.tolist() == [1, 1, 2, 2, 3, 4, 4, 4, 4]
    for k, t in enumerate(tstamps):
        a, b = int(t0[k]), int(t1[k])
        dP = se3.se3_mul(poses[b], se3.se3_inv(poses[a]))
        v = se3.se3_log(dP) / (ts[b] - ts[a] + 1e-3)
        want = se3.se3_mul(se3.se3_exp(v * (t - ts[a])), poses[a])
        assert torch.allclose(Gs.data[k], want, atol=1e-5), k
    assert torch.allclose(Gs.data[0], poses[0], atol=1e-6) and torch.allclose(Gs.data[8], poses[4], atol=1e-6)   # at / past the ends


def test_repeated_edge_filter_equals_the_reference_set_walk():
    """FactorGraph's duplicate-edge filter (one isin over packed keys) == the reference's per-element set walk
    (factor_graph.py:60-70): candidates already active or inactive go, repeats inside the candidate list stay"""
    import types
    from pvo_b200.factor_graph import FactorGraph
    g = torch.Generator().manual_seed(0)
    fg = FactorGraph.__new__(FactorGraph)
    fg.video = types.SimpleNamespace(poses=torch.zeros(40, 7))
    for trial in range(20):
        n = 12 + trial
        fg.ii, fg.jj = torch.randint(0, n, (30,), generator=g), torch.randint(0, n, (30,), generator=g)
        fg.ii_inac, fg.jj_inac = torch.randint(0, n, (10,), generator=g), torch.randint(0, n, (10,), generator=g)
        ii, jj = torch.randint(0, n, (50,), generator=g), torch.randint(0, n, (50,), generator=g)
        eset = set(zip(fg.ii.tolist(), fg.jj.tolist())) | set(zip(fg.ii_inac.tolist(), fg.jj_inac.tolist()))
        want = [(i, j) for i, j in zip(ii.tolist(), jj.tolist()) if (i, j) not in eset]
        a, b = fg._FactorGraph__filter_repeated_edges(ii, jj)
        assert list(zip(a.tolist(), b.tolist())) == want
    fg.ii = fg.jj = fg.ii_inac = fg.jj_inac = torch.zeros(0, dtype=torch.long)
    a, b = fg._FactorGraph__filter_repeated_edges(torch.tensor([1, 1]), torch.tensor([2, 2]))
    assert a.tolist() == [1, 1] and b.tolist() == [2, 2]


def test_visualiser_side_backends_match_the_pixel_loops():
    """droid_backends.iproj / projmap / depth_filter are tensor expressions; on CPU tensors they must reproduce the per-pixel
    restatements of the reference kernels (oracle.ba_cuda.*_loop, droid_kernels.cu:406-495, 640-829)"""
    from oracle import ba_cuda
    from pvo_b200 import droid_backends as db
    g = torch.Generator().manual_seed(4)
    F_, h, w = 9, 6, 8
    poses = se3.se3_exp(torch.arange(F_, dtype=torch.float32)[:, None] * torch.tensor([0.06, 0.01, 0.03, 0.0, 0.02, 0.004])
                        + 0.01 * torch.randn(F_, 6, generator=g))
    disps = 0.3 + 1.5 * torch.rand(F_, h, w, generator=g)
    disps[:, 2, 3] = 1e-4                                             # a far point: lands outside the neighbour image
    intr = torch.tensor([w / 2.0, w / 2.0, w / 2.0, h / 2.0])
    pts = db._iproj(poses, disps, intr)
    want = ba_cuda.iproj_loop(poses, disps, intr)
    assert pts.shape == (F_, h, w, 3) and (pts - want).abs().max() <= 1e-5 * want.abs().max()
    ii, jj = torch.tensor([0, 3, 8, 4]), torch.tensor([2, 0, 5, 4])
    p2 = poses.clone()
    p2[5, 2] -= 4.0                                                   # pushes the nearer points of 8 -> 5 behind the camera
    c, v = db._projmap(p2, disps, intr, ii, jj)
    cw, vw = ba_cuda.projmap_loop(p2, disps, intr, ii, jj)
    assert torch.equal(v, vw) and (v == 0).any() and (v == 1).any()
    assert (c - cw).abs().max() < 1e-3 and (c[..., 2] == 0).all()
    ix = torch.tensor([0, 2, 4, 8])
    thresh = torch.tensor([0.05, 0.1, 0.2, 0.1])
    cnt = db._depth_filter(poses, disps, intr, ix, thresh)
    cw = ba_cuda.depth_filter_loop(poses, disps, intr, ix, thresh)
    assert cnt.shape == (4, h, w) and cnt.max() >= 1 and (cnt == 0).any()
    assert (cnt != cw).float().mean() < 0.01                           # a float32 tie at a threshold / pixel edge at most
    with pytest.raises(RuntimeError):                                  # the public entry points keep the CUDA-only contract
        db.iproj(poses, disps, intr)


def test_every_factor_graph_gets_its_own_engine_state():
    """The update operator's fused engine keeps a graph's hidden state resident between iterations; MotionFilter's forward()
    call and other graphs (backend, trajectory filler) run in between.  Each owner must get a private fork: same packed
    weights, separate buffers; a weight reload invalidates the forks; a dead owner releases its fork."""
    import gc
    m = DynamicUpdateModule(precision='fp32')
    base = m.engine('fp32', 'cpu')

    class Owner:
        pass
    a, b = Owner(), Owner()
    ea, eb = m.engine('fp32', 'cpu', owner=a), m.engine('fp32', 'cpu', owner=b)
    assert ea is not eb and ea is not base and m.engine('fp32', 'cpu', owner=a) is ea and m.engine('fp32', 'cpu') is base
    assert ea.c_zr is base.c_zr and eb.glo_W is base.glo_W                  # weights shared
    ea.reserve(3, 2, 4, 6)
    eb.reserve(1, 1, 4, 6)
    base.reserve(1, 1, 4, 6)
    assert ea.X.shape[0] == 3 * 24 and eb.X.shape[0] == 24
    assert ea.X.data_ptr() != eb.X.data_ptr() != base.X.data_ptr()
    ea.X.fill_(1.0)
    base.reserve(2, 1, 4, 6)                                                # forward() with another edge count: a's state survives
    assert float(ea.X.min()) == 1.0
    m.load_state_dict(m.state_dict())                                       # weights reloaded -> new base, new forks
    assert m.engine('fp32', 'cpu', owner=a) is not ea
    n = len(m._forks)
    del a, ea
    gc.collect()
    assert len(m._forks) == n - 1


def test_geom_chol_matches_dense_algebra_and_its_gradient():
    """pvo_b200.geom.chol (reference interface geom/chol.py:5-73): solutions against torch.linalg.solve on the assembled system,
    CholeskySolver.backward against autograd through torch.linalg.solve, and the failed-factorisation contract"""
    import torch
    from pvo_b200.geom import chol
    g = torch.Generator().manual_seed(2)
    B, P, M, D, HW = 1, 3, 2, 6, 5
    J = torch.randn(B, P * D, P * D + 4, generator=g, dtype=torch.float64)
    Hfull = (J @ J.transpose(1, 2)) + 0.5 * torch.eye(P * D, dtype=torch.float64)
    H = Hfull.view(B, P, D, P, D).permute(0, 1, 3, 2, 4).contiguous()
    b = torch.randn(B, P, D, generator=g, dtype=torch.float64)
    x = chol.block_solve(H, b, ep=0.1, lm=1e-4)
    Hd = Hfull + (0.1 + 1e-4 * Hfull) * torch.eye(P * D, dtype=torch.float64)
    assert torch.allclose(x.reshape(B, -1, 1), torch.linalg.solve(Hd, b.reshape(B, -1, 1)), atol=1e-10)
    E = 0.3 * torch.randn(B, P, M, D, HW, generator=g, dtype=torch.float64)
    C = 2.0 + torch.rand(B, M, HW, generator=g, dtype=torch.float64)
    w = torch.randn(B, M, HW, generator=g, dtype=torch.float64)
    dx, dz = chol.schur_solve(H, E, C, b, w)
    Em = E.permute(0, 1, 3, 2, 4).reshape(B, P * D, M * HW)
    full = torch.cat([torch.cat([Hd, Em], 2), torch.cat([Em.transpose(1, 2), torch.diag_embed(C.view(B, -1))], 2)], 1)
    sol = torch.linalg.solve(full, torch.cat([b.reshape(B, -1, 1), w.reshape(B, -1, 1)], 1))
    assert torch.allclose(dx.reshape(B, -1, 1), sol[:, :P * D], atol=1e-9) and torch.allclose(dz.reshape(B, -1, 1), sol[:, P * D:], atol=1e-9)
    assert torch.equal(chol.schur_solve(H, E, C, b, w, sless=True), dx)
    Hr, br = Hd.clone().requires_grad_(True), b.reshape(B, -1, 1).clone().requires_grad_(True)
    gx = torch.randn(B, P * D, 1, generator=g, dtype=torch.float64)
    chol.CholeskySolver.apply(Hr, br).backward(gx)
    Ha, ba_ = Hd.clone().requires_grad_(True), b.reshape(B, -1, 1).clone().requires_grad_(True)
    torch.linalg.solve(Ha, ba_).backward(gx)
    # the reference returns dH = -x u^T, the transpose of autograd's -u x^T: identical for every symmetric H(theta), which is all BA builds
    sym = lambda t: t + t.transpose(-1, -2)
    assert torch.allclose(sym(Hr.grad), sym(Ha.grad), atol=1e-9) and torch.allclose(Hr.grad, Ha.grad.transpose(-1, -2), atol=1e-9)
    assert torch.allclose(br.grad, ba_.grad, atol=1e-9)
    bad = -torch.eye(4, dtype=torch.float64)[None]
    assert float(chol.CholeskySolver.apply(bad, torch.ones(1, 4, 1, dtype=torch.float64)).abs().max()) == 0.0


def test_launch_plan_records_and_replays_host_side():
    """_lib.LaunchPlan: calls made through the library proxy and torch_op closures while recording are re-issued verbatim by replay();
    nothing is recorded outside a recording, nested recordings restore the outer one, a failed recording is discarded"""
    import torch
    from pvo_b200 import _lib
    log = []

    class FakeLib:
        def pvo_fake(self, a, b):
            log.append((a, b))
            return 0
    proxy = _lib._Recording(FakeLib())
    proxy.pvo_fake(1, 2)                                   # not recording
    x = torch.zeros(3)
    with _lib.recording(True) as plan:
        _lib.check(proxy.pvo_fake(3, 4), 'fake')
        _lib.torch_op(lambda: x.add_(1))
        with _lib.recording(False) as inner:               # a disabled nested recording
            proxy.pvo_fake(5, 6)
        assert inner is None and _lib._REC[0] is plan
        keep = torch.ones(2)
        assert _lib.ptr(keep) == keep.data_ptr() and plan.keep[-1] is keep
    assert _lib._REC[0] is None and plan.launches == 1 and len(plan.calls) == 2
    n0 = _lib.LAUNCHES[0]
    plan.replay()
    plan.replay()
    assert log == [(1, 2), (3, 4), (5, 6), (3, 4), (3, 4)] and x.tolist() == [3.0, 3.0, 3.0] and _lib.LAUNCHES[0] == n0 + 2
    try:
        with _lib.recording(True) as bad:
            raise ValueError('boom')
    except ValueError:
        pass
    assert bad.calls is None and _lib._REC[0] is None


def test_tiled_volume_geometry():
    """tensor-core tile widths and the tiled level shapes (csrc/corr_lookup_tiled.cu): 4-row x 8-column tiles, zero padded"""
    from pvo_b200.modules.corr import pad_width, tiled_shape
    assert [pad_width(w) for w in (24, 40, 41, 80, 101, 104, 128, 129)] == [40, 40, 80, 80, 104, 104, 128, None]
    assert tiled_shape(36, 60, 80, 0) == (36, 4800, 15, 10, 32) and tiled_shape(36, 60, 80, 3) == (36, 4800, 2, 2, 32)
    assert tiled_shape(5, 30, 101, 1) == (5, 3030, 4, 7, 32) and tiled_shape(5, 30, 101, 3) == (5, 3030, 1, 2, 32)
