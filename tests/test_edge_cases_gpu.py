"""Edge cases of the hot path on the B200 kernels: empty and ragged inputs, degenerate graphs, error behaviour
(contiguity / device violations raise RuntimeError; a failed solve is a zero update)."""
import pytest
import torch

from oracle import corr as ocorr
from oracle import ba_cuda
from tests.golden import cases

pytestmark = pytest.mark.gpu


def test_lookup_empty_batch_and_ragged_sizes():
    from pvo_b200 import droid_backends
    from pvo_b200.modules.corr import CorrBlock
    # zero edges: a well-formed empty result, no launch error
    vol = torch.zeros(0, 12, 20, 12, 20, device='cuda')
    out, = droid_backends.corr_index_forward(vol, torch.zeros(0, 2, 12, 20, device='cuda'), 3)
    assert out.shape == (0, 7, 7, 12, 20)
    # odd target sizes (pyramid 13x21 -> 6x10 -> 3x5 -> 1x2): rows are not 16-byte aligned, windows mostly out of range
    g = torch.Generator().manual_seed(3)
    for dtype in (torch.float32, torch.float16):
        f1 = torch.randn(1, 2, 128, 13, 21, generator=g).to(dtype).cuda()
        f2 = torch.randn(1, 2, 128, 13, 21, generator=g).to(dtype).cuda()
        blk = CorrBlock(f1, f2)
        assert [tuple(p.shape[-2:]) for p in blk.corr_pyramid] == [(13, 21), (6, 10), (3, 5), (1, 2)]
        coords = (torch.rand(1, 2, 13, 21, 2, generator=g) * torch.tensor([26.0, 18.0]) - 3.0).cuda()
        got = blk(coords)
        it = torch.int16 if dtype == torch.float16 else torch.int32
        for lvl, vol in enumerate(blk.corr_pyramid):
            c = (coords[0].permute(0, 3, 1, 2) / 2 ** lvl).contiguous()
            want = ocorr.corr_index_forward(vol.cpu(), c.cpu(), 3)
            have = got[0, :, 49 * lvl:49 * (lvl + 1)].reshape(2, 7, 7, 13, 21)
            assert torch.equal(have.cpu().contiguous().view(it), want.view(it)), (dtype, lvl)


def test_lookup_all_windows_out_of_range_is_zero():
    from pvo_b200.modules.corr import CorrBlock
    f = torch.randn(1, 1, 128, 16, 24, device='cuda')
    blk = CorrBlock(f, f)
    coords = torch.full((1, 1, 16, 24, 2), -1000.0, device='cuda')
    assert blk(coords).abs().max() == 0


def test_error_behaviour_matches_the_reference_contract():
    from pvo_b200 import droid_backends
    vol = torch.zeros(1, 8, 8, 8, 8, device='cuda')
    coords = torch.zeros(1, 2, 8, 8, device='cuda')
    with pytest.raises(RuntimeError):                       # CHECK_CONTIGUOUS, droid.cpp:83-84
        droid_backends.corr_index_forward(vol.transpose(1, 2), coords, 3)
    with pytest.raises(RuntimeError):                       # CPU tensors never reach a kernel: there is no CPU path
        droid_backends.corr_index_forward(vol.cpu(), coords.cpu(), 3)
    with pytest.raises(RuntimeError):
        droid_backends.corr_index_forward(vol, coords.double(), 3)


def _video(c, buffer):
    from pvo_b200.depth_video import DepthVideo
    F_, h, w = c['F'], c['h'], c['w']
    v = DepthVideo(image_size=(8 * h, 8 * w), buffer=buffer, device='cuda', store_images=False)
    v.poses[:F_] = c['poses'].cuda()
    v.disps[:F_] = c['disps'].cuda()
    v.intrinsics[:] = c['intrinsics'][0].cuda()
    v.counter.value = F_
    return v


def test_ba_single_free_pose_and_zero_weight_edges():
    """P = 1 (one optimised pose) and edges that carry no information: finite result equal to the oracle"""
    from pvo_b200 import droid_backends
    c = cases.graph_case(F=3, radius=2, h=12, w=16, seed=21)
    ii, jj = c['ii'], c['jj']
    g = torch.Generator().manual_seed(5)
    E = ii.shape[0]
    tgt = torch.rand(E, 2, 12, 16, generator=g) * torch.tensor([16.0, 12.0]).view(1, 2, 1, 1)
    wgt = torch.rand(E, 2, 12, 16, generator=g)
    wgt[::2] = 0                                            # every other edge: zero confidence
    t0, t1 = 2, 3                                           # frames 0, 1 fixed -> P = 1
    kx = torch.unique(torch.cat([torch.arange(t0, t1), ii]))
    eta = torch.full((kx.shape[0], 12, 16), 0.02)
    p_ref, d_ref = c['poses'].clone(), c['disps'].clone()
    ba_cuda.ba(p_ref, d_ref, c['intrinsics'][0], tgt, wgt, eta, ii, jj, t0, t1, 2, 1e-4, 0.1, False)
    v = _video(c, 8)
    droid_backends.ba(v.poses, v.disps, v.intrinsics[0].contiguous(), tgt.cuda(), wgt.cuda(), eta.cuda(), ii.cuda(), jj.cuda(),
                      t0, t1, 2, 1e-4, 0.1, False)
    assert torch.isfinite(v.poses).all() and torch.isfinite(v.disps).all()
    assert (v.poses[:3].cpu() - p_ref).abs().max() < 1e-4
    assert ((v.disps[:3].cpu() - d_ref).abs() / d_ref.abs().clamp(min=1e-3)).max() < 1e-4


def test_ba_all_weights_zero_leaves_the_state_in_place():
    """no information at all: H = 0, the damped system gives dx = 0 and dz = 0 (never NaN)"""
    from pvo_b200 import droid_backends
    c = cases.graph_case(F=4, radius=1, h=12, w=16, seed=22)
    ii, jj = c['ii'], c['jj']
    E = ii.shape[0]
    tgt = torch.zeros(E, 2, 12, 16)
    wgt = torch.zeros(E, 2, 12, 16)
    kx = torch.unique(torch.cat([torch.arange(1, 4), ii]))
    eta = torch.full((kx.shape[0], 12, 16), 0.02)
    v = _video(c, 8)
    p0, d0 = v.poses.clone(), v.disps.clone()
    dx, dz = droid_backends.ba(v.poses, v.disps, v.intrinsics[0].contiguous(), tgt.cuda(), wgt.cuda(), eta.cuda(), ii.cuda(), jj.cuda(),
                               1, 4, 2, 1e-4, 0.1, False)
    assert torch.isfinite(dx).all() and torch.isfinite(dz).all()
    assert (v.poses - p0).abs().max() < 1e-6 and (v.disps - d0).abs().max() < 1e-6


def test_fpn_heads_with_no_detections():
    """score threshold above every score -> zero detections: mask head, paste and merge degrade to the stuff-only map"""
    from pvo_b200.panoptic_fpn import PanopticFPN
    m = PanopticFPN.__new__(PanopticFPN)
    m.device = torch.device('cuda')
    assert m.mask_head(torch.zeros(0, 14, 14, 256, dtype=torch.float16, device='cuda')).shape == (0, 3, 28, 28)
    assert m.paste(torch.zeros(0, 1, 28, 28, device='cuda'), torch.zeros(0, 4, device='cuda'), 64, 96).shape == (0, 64, 96)
    sem = torch.zeros(1, 64, 96, dtype=torch.int64, device='cuda')
    sem[:, :, 48:] = 5                                      # 64*48 = 3072 < 4096 -> dropped too
    sem[:, 16:, :48] = 2                                    # 48*48 = 2304 < 4096
    sem[:, :, 10:90] = 7                                    # 64*80 = 5120 -> kept
    empty = (torch.zeros(0, device='cuda'), torch.zeros(0, dtype=torch.long, device='cuda'), torch.zeros(0, 64, 96, dtype=torch.bool, device='cuda'))
    pan, segs = m.combine([empty], sem, boxes=[torch.zeros(0, 4, device='cuda')])
    assert [s['id'] for s in segs[0]] == [70000]
    assert torch.equal(pan[0] == 70000, sem[0] == 7) and int((pan[0] != 0).sum()) == 5120
