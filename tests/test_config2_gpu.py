"""Parity AT THE BENCHMARKED SIZE (BASELINE.json configs[1]: 8 keyframes, radius-3 graph = 36 edges, 480x640 images = 60x80
grid) -- VERDICT r1 'Next round' item 1(b)-(d):

  * correlation lookup, all four levels of a real pyramid, bit for bit against the reference's own kernel (oracle/_ref);
  * dense BA against the reference's own CUDA kernels (oracle/_ref) and the CPU oracle;
  * two FactorGraph.update steps (float32 path) against the CPU oracle: poses / depth 1e-4, flow EPE 1e-5;
  * the benchmarked half / tensor-core path: its deviation from float32 is gated by the deviation of the reference's own
    inference mode (update operator under torch.autocast, cuDNN) on the same GPU;
  * corr_index_backward against the reference's backward kernel.
"""
import pytest
import torch

from oracle import ba_cuda, corr as ocorr, step as ostep, update_op as oupd
from tests import ref_harness as rh
from tests.golden import cases

pytestmark = pytest.mark.gpu
E2, H2, W2 = 36, 60, 80


def _ref_corr():
    from oracle import build_ref
    try:
        return build_ref.load()
    except Exception:
        return None


def _ref_ba():
    from oracle import build_ref
    try:
        return build_ref.load_ba()
    except Exception:
        return None


def _flow_coords(E, h, w, seed, spread):
    """coords1 as the graph produces them: the pixel grid plus a smooth displacement, plus noise; a band of pixels is pushed out of
    the image on every side so that every boundary branch of every level is visited"""
    g = torch.Generator().manual_seed(seed)
    y, x = torch.meshgrid(torch.arange(h).float(), torch.arange(w).float(), indexing='ij')
    base = torch.stack([x, y], -1)[None].repeat(E, 1, 1, 1)
    coarse = (torch.rand(E, 2, 5, 6, generator=g) * 2 - 1) * spread
    smooth = torch.nn.functional.interpolate(coarse, size=(h, w), mode='bilinear', align_corners=True).permute(0, 2, 3, 1)
    c = base + smooth + 0.75 * torch.randn(E, h, w, 2, generator=g)
    c[:, :2] -= 9.0
    c[:, -2:] += 11.0
    c[:, :, :2, 0] -= 7.5
    c[:, :, -2:, 0] += 13.25
    c[:, 5] = torch.round(c[:, 5])                      # exact integers: dx = dy = 0
    return c.contiguous()


def test_lookup_config2_bit_exact_vs_reference_kernel():
    """E=36, 60x80, the four levels of a real half pyramid built by the tensor-core volume kernel: the fused single-launch lookup
    (reference layout and the pixel-major layout the update operator consumes) == the reference's corr_index_forward_kernel
    called level by level as CorrBlock.__call__ does (modules/corr.py:40-50)."""
    ref = _ref_corr()
    if ref is None:
        pytest.skip('oracle/_ref/droid_corr_ref.so not present')
    from pvo_b200.modules.corr import CorrBlock
    g = torch.Generator().manual_seed(2)
    f1 = torch.randn(1, E2, 128, H2, W2, generator=g).half().cuda()
    f2 = torch.randn(1, E2, 128, H2, W2, generator=g).half().cuda()
    blk = CorrBlock(f1, f2)
    coords = _flow_coords(E2, H2, W2, 3, spread=6.0).cuda()[None]                      # (1,E,h,w,2)
    ours = blk(coords)                                                                # (1,E,196,h,w)
    planar = coords[0].permute(0, 3, 1, 2).contiguous()                               # (E,2,h,w)
    for i in range(4):
        vol = blk.corr_pyramid[i].reshape(E2, H2, W2, H2 >> i, W2 >> i)
        r, = ref.corr_index_forward(vol, (planar / 2 ** i).contiguous(), 3)           # (E,7,7,h,w)
        got = ours[0, :, 49 * i:49 * (i + 1)].reshape(E2, 7, 7, H2, W2)
        assert torch.equal(got.view(torch.int16), r.view(torch.int16)), 'level %d' % i
        del r
    pm = torch.full((E2 * H2 * W2, 208), 7.0, dtype=torch.float16, device='cuda')
    blk(coords, out=pm, layout=1, out_cstride=208)
    pm = pm.view(E2, H2 * W2, 208)
    assert torch.equal(pm[:, :, :196].permute(0, 2, 1).reshape(1, E2, 196, H2, W2).contiguous().view(torch.int16), ours.view(torch.int16))
    assert (pm[:, :, 196:] == 0).all()


@pytest.mark.parametrize('motion_only', [False, True])
def test_ba_config2_vs_reference_kernels(motion_only):
    """F=8, 36 edges, 60x80: pvo_ba == the reference's BA kernels (oracle/_ref, documented patches) == the CPU oracle, 1e-4."""
    ref = _ref_ba()
    if ref is None:
        pytest.skip('oracle/_ref/droid_ba_ref.so not present')
    from pvo_b200 import droid_backends
    from tests.test_geom_gpu import _ba_case, _rel, _video
    c, tgt, wgt, g = _ba_case(8, 3, H2, W2, seed=77)
    ii, jj, t0, t1 = c['ii'], c['jj'], 1, 8
    kx = torch.unique(torch.cat([torch.arange(t0, t1), ii]))
    eta = 0.02 * torch.rand(kx.shape[0], H2, W2, generator=g) + 1e-7
    pr, dr = c['poses'].clone().cuda(), c['disps'].clone().cuda()
    ref.ba(pr, dr, c['intrinsics'][0].cuda().contiguous(), tgt.cuda(), wgt.cuda(), eta.cuda(), ii.cuda(), jj.cuda(), t0, t1, 2, 1e-4, 0.1,
           motion_only)
    po, do = c['poses'].clone(), c['disps'].clone()
    ba_cuda.ba(po, do, c['intrinsics'][0], tgt, wgt, eta, ii, jj, t0, t1, 2, 1e-4, 0.1, motion_only)
    v = _video(c, buffer=11)
    droid_backends.ba(v.poses, v.disps, v.intrinsics[0].contiguous(), tgt.cuda(), wgt.cuda(), eta.cuda(), ii.cuda(), jj.cuda(), t0, t1, 2,
                      1e-4, 0.1, motion_only)
    assert _rel(po, pr.cpu()) < 1e-4, 'CPU oracle != reference kernels (poses)'
    assert _rel(v.poses[:8].cpu(), pr.cpu()) < 1e-4, 'pvo_ba != reference kernels (poses)'
    if not motion_only:
        assert _rel(do, dr.cpu()) < 1e-4, 'CPU oracle != reference kernels (depth)'
        assert _rel(v.disps[:8].cpu(), dr[:8].cpu()) < 1e-4, 'pvo_ba != reference kernels (depth)'


def test_graph_update_config2_float32_matches_oracle():
    """two update iterations of the config-2 graph, float32 path, vs the CPU oracle (2-3 s per oracle step)"""
    c = cases.graph_case()                        # defaults == config 2
    assert (c['E'], c['h'], c['w']) == (E2, H2, W2)
    sd = oupd.init_params(0)
    st = ostep.make_state(c)
    v, fg = rh.our_graph(c, sd, 'fp32')
    with torch.no_grad():
        for it in range(2):
            ostep.graph_update(sd, st)
            fg.update(itrs=2)
            dt, ang = rh.pose_err(v.poses[:8].cpu(), st['poses'][:8])
            assert dt < 1e-4 and ang < 1e-4, (it, dt, ang)
            drel = ((v.disps[:8].cpu() - st['disps'][:8]).abs() / st['disps'][:8].abs().clamp(min=1e-3)).max().item()
            assert drel < 1e-4, (it, drel)
            epe = (fg.full_flow.cpu() - st['full_flow']).norm(dim=-1).mean().item()
            assert epe < 1e-5, (it, epe)


def test_benchmarked_half_path_is_gated_by_the_reference_autocast_path():
    """The path bench.py times (half state, tcgen05 convolutions, CUDA graph) is not a 1e-4 parity path -- neither is the
    reference's inference mode, which autocasts the update operator to half (factor_graph.py:227).  Gate: after two update
    iterations at config 2, every observable of ours-fp16 deviates from the float32 result by at most 2x what the reference's
    own autocast run (its unmodified FactorGraph + DynamicUpdateModule on cuDNN, same GPU, same inputs) deviates."""
    from oracle import ref_stage
    if not ref_stage.available():
        pytest.skip('oracle/_ref/py not staged')
    r = rh.fp16_gate(steps=2)
    ours, ref = r['ours_fp16'], r['reference_autocast']
    print('fp16 gate:', r)
    floor = dict(pose_rel=1e-4, rot_rad=1e-4, depth_rel_mean=1e-4, depth_rel_max=1e-3, flow_epe=1e-3)
    for k in ours:
        assert ours[k] <= 2.0 * ref[k] + floor[k], (k, ours[k], ref[k])


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16])
def test_corr_index_backward_vs_reference_kernel(dtype):
    """droid_backends.corr_index_backward (training only; correlation_kernels.cu:73-124, :157-185).  The reference scatters with
    non-atomic read-modify-writes from one thread per SOURCE pixel into that pixel's private plane, so it is deterministic; the
    float32 result is compared to 1e-6 of the gradient scale (accumulation order within a plane may differ), half to 2 ulp."""
    ref = _ref_corr()
    if ref is None:
        pytest.skip('oracle/_ref/droid_corr_ref.so not present')
    from pvo_b200 import droid_backends
    g = torch.Generator().manual_seed(41)
    N, h1, w1, h2, w2 = 3, 20, 28, 20, 28
    vol = torch.randn(N, h1, w1, h2, w2, generator=g).to(dtype).cuda()
    grad = torch.randn(N, 7, 7, h1, w1, generator=g).to(dtype).cuda()
    y, x = torch.meshgrid(torch.arange(h1).float(), torch.arange(w1).float(), indexing='ij')
    coords = (torch.stack([x, y], 0)[None] + 5.0 * (torch.rand(N, 2, h1, w1, generator=g) * 2 - 1)).contiguous()
    coords[:, :, 0] -= 30.0                                  # a row of windows entirely outside
    coords[:, :, 1] = torch.round(coords[:, :, 1])
    coords = coords.cuda()
    r, = ref.corr_index_backward(vol, coords, grad, 3)
    o, = droid_backends.corr_index_backward(vol, coords, grad, 3)
    assert o.shape == r.shape and o.dtype == r.dtype
    scale = r.float().abs().max().item()
    tol = 1e-6 * scale if dtype == torch.float32 else 2 * torch.finfo(torch.float16).eps * scale
    assert (o.float() - r.float()).abs().max().item() <= tol
    # and through autograd, as CorrSampler.backward does (modules/corr.py:14-20)
    from pvo_b200.modules.corr import CorrSampler
    v = vol.clone().requires_grad_(True)
    out = CorrSampler.apply(v, coords, 3)
    out.backward(grad)
    assert torch.equal(v.grad, o)
