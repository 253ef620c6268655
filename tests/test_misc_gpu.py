"""Round-2 additions: convex / bilinear 8x upsampling kernels (droid_net.py:23-61), altcorr backward (altcorr_kernel.cu:152-286)
and the launch-plan replay of FactorGraph.update."""
import pytest
import torch

from tests.golden import cases

pytestmark = pytest.mark.gpu


def _cvx_reference(data, mask):
    """droid_net.py:23-37 restated with plain torch ops (the float32 reference of the kernel)"""
    import torch.nn.functional as F
    batch, ht, wd, dim = data.shape
    d = data.permute(0, 3, 1, 2)
    m = torch.softmax(mask.view(batch, 1, 9, 8, 8, ht, wd), dim=2)
    up = F.unfold(d, [3, 3], padding=1).view(batch, dim, 9, 1, 1, ht, wd)
    up = torch.sum(m * up, dim=2).permute(0, 4, 2, 5, 3, 1)
    return up.reshape(batch, 8 * ht, 8 * wd, dim)


@pytest.mark.parametrize('shape', [(2, 12, 16, 1), (1, 30, 101, 2), (3, 7, 5, 4)])
@pytest.mark.parametrize('mdtype', [torch.float32, torch.float16])
def test_cvx_upsample_matches_torch(shape, mdtype):
    from pvo_b200.droid_net import cvx_upsample, upsample_dim_1
    g = torch.Generator().manual_seed(3)
    N, ht, wd, dim = shape
    data = torch.randn(N, ht, wd, dim, generator=g).cuda()
    mask = (3 * torch.randn(N, 576, ht, wd, generator=g)).to(mdtype).cuda()
    ref = _cvx_reference(data, mask.float())
    out = cvx_upsample(data, mask)
    assert out.shape == ref.shape
    assert torch.allclose(out.float(), ref, atol=2e-6 if mdtype == torch.float32 else 2e-3, rtol=1e-5)
    if dim == 1:
        a = upsample_dim_1(data.view(1, N, ht, wd), mask.view(1, N, 576, ht, wd))
        assert torch.equal(a.view(N, 8 * ht, 8 * wd, 1), out)


@pytest.mark.parametrize('shape', [(1, 3, 12, 16, 2), (2, 2, 30, 101, 1), (1, 1, 5, 7, 4)])
def test_upsample_inter_matches_torch(shape):
    import torch.nn.functional as F
    from pvo_b200.droid_net import upsample_inter
    g = torch.Generator().manual_seed(4)
    x = torch.randn(*shape, generator=g).cuda()
    b, n, ht, wd, dim = shape
    ref = F.interpolate(x.permute(0, 1, 4, 2, 3).reshape(b * n, dim, ht, wd), scale_factor=8, mode='bilinear', align_corners=True,
                        recompute_scale_factor=True).permute(0, 2, 3, 1).reshape(b, n, 8 * ht, 8 * wd, dim)
    out = upsample_inter(x)
    assert out.shape == ref.shape and torch.allclose(out, ref, atol=1e-5, rtol=1e-5)


def test_altcorr_backward_vs_reference_kernel():
    """droid_backends.altcorr_backward == the reference's altcorr_backward_kernel (oracle/_ref) and == autograd of a dense torch
    restatement of the forward"""
    from oracle import build_ref
    from pvo_b200 import droid_backends
    from pvo_b200.modules.corr import CorrLayer
    g = torch.Generator().manual_seed(31)
    B, S, H, W, C = 2, 2, 12, 16, 128
    f1 = (torch.randn(B, H, W, C, generator=g) / 4).cuda()
    f2 = (torch.randn(B, H, W, C, generator=g) / 4).cuda()
    y, x = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing='ij')
    coords = (torch.stack([x, y], -1)[None, None] + 4 * (torch.rand(B, S, H, W, 2, generator=g) * 2 - 1)).contiguous()
    coords[:, :, 0] -= 40.0
    coords = coords.cuda()
    grad = torch.randn(B, S, 49, H, W, generator=g).cuda()
    g1, g2, gc = droid_backends.altcorr_backward(f1, f2, coords, grad, 3)
    assert gc.shape == coords.shape and float(gc.abs().max()) == 0.0
    try:
        ref = build_ref.load()
    except Exception:
        ref = None
    if ref is not None:
        r1, r2, rc = ref.altcorr_backward(f1, f2, coords, grad, 3)
        assert torch.allclose(g1, r1, atol=1e-4, rtol=1e-4)
        assert torch.allclose(g2, r2, atol=1e-4, rtol=1e-4)
    # autograd through CorrLayer
    a, b = f1.clone().requires_grad_(True), f2.clone().requires_grad_(True)
    out = CorrLayer.apply(a, b, coords, 3)
    out.backward(grad)
    assert torch.equal(a.grad, g1) and torch.allclose(b.grad, g2, atol=1e-5, rtol=1e-5)
    # finite-difference check of one direction (the forward is bilinear in each feature map)
    d1 = torch.randn_like(f1)
    lhs = (droid_backends.altcorr_forward(f1 + d1, f2, coords, 3)[0] - droid_backends.altcorr_forward(f1, f2, coords, 3)[0])
    assert abs(float((lhs * grad).sum()) - float((d1 * g1).sum())) < 2e-3 * max(1.0, abs(float((d1 * g1).sum())))


@pytest.mark.parametrize('use_inactive', [False, True])
def test_launch_plan_replay_equals_marshalled_launches(use_inactive):
    """FactorGraph.update re-issues the recorded launches of an edge set (pvo_b200._lib.LaunchPlan) -- same results as issuing
    every launch through the marshalling path, and a new edge set records a new plan"""
    from tests.test_step_gpu import build_graph, pose_err
    c = cases.graph_case(F=6, radius=2, h=16, w=40, seed=5, with_segm=True)
    res = []
    for plan in (False, True):
        v, fg = build_graph(c, 'fp16', segm_filter=True, thresh=0.5)
        fg.use_launch_plan = plan
        fg.update(itrs=2, use_inactive=use_inactive)
        fg.rm_factors((fg.ii == 0) & (fg.jj == 2), store=True)
        for _ in range(3):
            fg.update(itrs=2, use_inactive=use_inactive)
        if plan:
            assert len(fg._cache['plans']) == 1 and fg.launch_count > 20
        res.append((v.poses[:6].clone(), v.disps[:6].clone(), fg.target_cam.clone(), fg.age.clone()))
    dt, ang = pose_err(res[1][0].cpu(), res[0][0].cpu())
    assert dt < 2e-3 and ang < 2e-3
    assert (res[1][2] - res[0][2]).abs().max().item() < 0.2
    assert torch.equal(res[1][3], res[0][3])


def test_ba_reports_misaligned_damping_rows_when_checked():
    """eta must have one row per depth frame unique(arange(t0,t1) U ii) (droid_kernels.cu:1311-1319; the reference raises a shape error
    at `C + eta`).  The schedule kernel records the mismatch on the device; with droid_backends.CHECK_BA the call raises."""
    from pvo_b200 import droid_backends
    from tests.test_geom_gpu import _ba_case, _video
    c, tgt, wgt, g = _ba_case(6, 2, 12, 16, seed=3)
    ii, jj = c['ii'].cuda(), c['jj'].cuda()
    v = _video(c, buffer=9)
    K = torch.unique(torch.cat([torch.arange(1, 6), c['ii']])).numel()     # depth frames of the window t0 = 1, t1 = 6 and the edges
    eta_short = torch.full((K - 2, 12, 16), 0.01, device='cuda')
    old = droid_backends.CHECK_BA
    droid_backends.CHECK_BA = True
    try:
        with pytest.raises(RuntimeError, match='depth frames'):
            droid_backends.ba(v.poses, v.disps, v.intrinsics[0].contiguous(), tgt.cuda(), wgt.cuda(), eta_short, ii, jj, 1, 6, 2, 1e-4, 0.1, False)
        eta_ok = torch.full((K, 12, 16), 0.01, device='cuda')
        droid_backends.ba(v.poses, v.disps, v.intrinsics[0].contiguous(), tgt.cuda(), wgt.cuda(), eta_ok, ii, jj, 1, 6, 2, 1e-4, 0.1, False)
    finally:
        droid_backends.CHECK_BA = old
