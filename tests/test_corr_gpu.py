"""GPU parity: correlation volume, pyramid and window lookup against the CPU oracle (bit-exact for the lookup)."""
import numpy as np
import pytest
import torch

from oracle import corr as ocorr
from tests.golden import cases

pytestmark = pytest.mark.gpu


def _coords(N, h, w, seed, spread=3.0, far=False):
    g = torch.Generator().manual_seed(seed)
    y, x = torch.meshgrid(torch.arange(h).float(), torch.arange(w).float(), indexing='ij')
    c = torch.stack([x, y], 0)[None].repeat(N, 1, 1, 1)
    c = c + (torch.rand(N, 2, h, w, generator=g) * 2 - 1) * spread
    if far:   # push some pixels well outside (all taps out of bounds) and onto exact integers / borders
        c[:, :, 0, :] -= 20.0
        c[:, :, -1, :] += 50.0
        c[:, :, 1, :] = torch.round(c[:, :, 1, :])
        c[:, 0, 2, :] = -3.5
        c[:, 1, 3, :] = h + 2.25
    return c.contiguous()


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16])
@pytest.mark.parametrize('shape', [(3, 12, 16, 12, 16), (2, 9, 13, 5, 7), (2, 8, 8, 2, 3)])
def test_corr_index_forward_bit_exact(dtype, shape):
    from pvo_b200 import droid_backends
    N, h1, w1, h2, w2 = shape
    g = torch.Generator().manual_seed(5)
    vol = torch.randn(N, h1, w1, h2, w2, generator=g).to(dtype)
    # coordinates live in the (h2, w2) frame of this level
    coords = _coords(N, h1, w1, 11, spread=4.0, far=True) * (w2 / w1)
    ref = ocorr.corr_index_forward(vol, coords, 3)
    out, = droid_backends.corr_index_forward(vol.cuda(), coords.cuda(), 3)
    assert out.dtype == dtype and out.shape == ref.shape
    assert torch.equal(out.cpu().view(torch.int16 if dtype == torch.float16 else torch.int32),
                       ref.view(torch.int16 if dtype == torch.float16 else torch.int32))


@pytest.mark.parametrize('radius', [0, 1, 2, 4])
def test_corr_index_forward_other_radii(radius):
    from pvo_b200 import droid_backends
    g = torch.Generator().manual_seed(6)
    vol = torch.randn(2, 6, 7, 10, 9, generator=g).half()
    coords = _coords(2, 6, 7, 12, spread=5.0, far=True)
    ref = ocorr.corr_index_forward(vol, coords, radius)
    out, = droid_backends.corr_index_forward(vol.cuda(), coords.cuda(), radius)
    assert torch.equal(out.cpu().view(torch.int16), ref.view(torch.int16))


def test_vector_and_scalar_row_loads_agree():
    """flags bit0 switches the half fast path to 2-byte loads: same bits either way."""
    from pvo_b200 import _lib
    g = torch.Generator().manual_seed(8)
    vol = torch.randn(4, 20, 24, 20, 24, generator=g).half().cuda()
    coords = _coords(4, 20, 24, 13, spread=6.0).cuda()
    outs = []
    for flags in (0, 1):
        out = torch.empty(4, 7, 7, 20, 24, dtype=torch.float16, device='cuda')
        _lib.check(_lib.lib().pvo_corr_index_forward(vol.data_ptr(), coords.data_ptr(), out.data_ptr(), 4, 20, 24, 20, 24, 3,
                                                     _lib.F16, flags, _lib.stream()), 'corr_index_forward')
        outs.append(out)
    assert torch.equal(outs[0].view(torch.int16), outs[1].view(torch.int16))


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16])
def test_corrblock_call_matches_oracle_both_layouts(dtype):
    """CorrBlock.__call__: all 4 levels + concat in one launch, reference layout and pixel-major layout."""
    from pvo_b200.modules.corr import CorrBlock
    c = cases.graph_case(F=4, radius=2, h=16, w=24, seed=21, fdtype=dtype)
    ii, jj = c['ii'], c['jj']
    E = ii.shape[0]
    f1, f2 = c['fmaps'][ii][None], c['fmaps'][jj][None]
    pyr = ocorr.corr_pyramid(f1, f2)
    coords = _coords(E, 16, 24, 31, spread=5.0, far=True).permute(0, 2, 3, 1)[None].contiguous()   # (1,E,h,w,2)
    ref = ocorr.corr_lookup(pyr, coords, 3)
    blk = CorrBlock(f1.cuda(), f2.cuda(), use_tc=False)
    # use the ORACLE's pyramid so the gather is tested on identical bits
    blk.corr_pyramid = [p.cuda() for p in pyr]
    out = blk(coords.cuda())
    it = torch.int16 if dtype == torch.float16 else torch.int32
    assert torch.equal(out.cpu().view(it), ref.view(it))
    pm = torch.full((E * 16 * 24, 208), 7.0, dtype=dtype, device='cuda')
    blk(coords.cuda(), out=pm, layout=1, out_cstride=208)
    pm = pm.cpu().view(E, 16 * 24, 208)
    assert torch.equal(pm[:, :, :196].permute(0, 2, 1).reshape(1, E, 196, 16, 24).contiguous().view(it), ref.view(it))
    assert (pm[:, :, 196:] == 0).all()


@pytest.mark.parametrize('dtype,tol', [(torch.float32, 2e-5), (torch.float16, None)])
def test_volume_and_pyramid_cuda_cores(dtype, tol):
    from pvo_b200.modules.corr import build_pyramid
    g = torch.Generator().manual_seed(3)
    f1 = torch.randn(3, 128, 16, 24, generator=g).to(dtype)
    f2 = torch.randn(3, 128, 16, 24, generator=g).to(dtype)
    ref = ocorr.corr_pyramid(f1[None], f2[None])
    out = build_pyramid(f1.cuda(), f2.cuda(), 4, use_tc=False)
    for l in range(4):
        a, b = out[l].cpu().float(), ref[l].float()
        assert a.shape == b.shape
        if dtype == torch.float32:
            assert torch.allclose(a, b, atol=tol, rtol=1e-5), l
        else:
            # summation order differs from the CPU matmul: within one half ulp of the fp32-accumulated value
            assert (a - b).abs().max() <= 2 * torch.finfo(torch.float16).eps * b.abs().max().clamp(min=1), l
    # the pooling cascade itself is exact: level l+1 == avg_pool(level l) of the kernel's OWN output
    for l in range(3):
        x = out[l].cpu()
        R = x.shape[0] * x.shape[1] * x.shape[2]
        pooled = ocorr.avg_pool_level(x.reshape(R, 1, x.shape[3], x.shape[4])).reshape(out[l + 1].shape)
        assert torch.equal(pooled, out[l + 1].cpu()), l


@pytest.mark.parametrize('hw', [(60, 80), (30, 40), (16, 40), (9, 80)])
def test_volume_tensor_cores_match_cuda_cores(hw):
    """tcgen05 volume build (fp32 accumulate in TMEM) vs the CUDA-core build and the oracle."""
    from pvo_b200.modules.corr import build_pyramid
    h, w = hw
    g = torch.Generator().manual_seed(4)
    E = 3
    f1 = torch.randn(E, 128, h, w, generator=g).half()
    f2 = torch.randn(E, 128, h, w, generator=g).half()
    tc = build_pyramid(f1.cuda(), f2.cuda(), 4, use_tc=True)
    cc = build_pyramid(f1.cuda(), f2.cuda(), 4, use_tc=False)
    ulp = torch.finfo(torch.float16).eps
    for l in range(4):
        assert tc[l].shape == cc[l].shape
        d = (tc[l].float() - cc[l].float()).abs().max().item()
        assert d <= 2 * ulp * max(1.0, cc[l].float().abs().max().item()), (l, d)
    for l in range(3):     # the fused pooling epilogue reproduces the avg_pool2d cascade bit for bit
        x = tc[l].cpu()
        R = x.shape[0] * x.shape[1] * x.shape[2]
        pooled = ocorr.avg_pool_level(x.reshape(R, 1, x.shape[3], x.shape[4])).reshape(tc[l + 1].shape)
        assert torch.equal(pooled, tc[l + 1].cpu()), l
    # level 0 against the oracle on a sample of source pixels
    ref = ocorr.corr_volume(f1[None], f2[None])[0]
    assert (tc[0].cpu().float() - ref.float()).abs().max() <= 2 * ulp * ref.float().abs().max()


def test_lookup_full_size_properties():
    """BASELINE config-2 size (E=36, 60x80): linearity in the volume and the integer-coordinate identity."""
    from pvo_b200.modules.corr import CorrBlock
    g = torch.Generator(device='cuda').manual_seed(0)
    E, h, w = 36, 60, 80
    f1 = torch.randn(1, E, 128, h, w, device='cuda', generator=g).half()
    f2 = torch.randn(1, E, 128, h, w, device='cuda', generator=g).half()
    blk = CorrBlock(f1, f2)
    yy, xx = torch.meshgrid(torch.arange(h, device='cuda').float(), torch.arange(w, device='cuda').float(), indexing='ij')
    grid = torch.stack([xx, yy], -1)[None, None].repeat(1, E, 1, 1, 1).contiguous()
    out = blk(grid)                                   # integer coords: window centre == volume diagonal entries
    vol0 = blk.corr_pyramid[0]
    centre = out[0, :, 3 * 7 + 3]                     # level 0, (xoff, yoff) = (0, 0)
    diag = vol0.reshape(E, h * w, h * w).diagonal(dim1=1, dim2=2).reshape(E, h, w)
    assert torch.equal(centre, diag)
    # doubling the volume doubles the output: power-of-two scaling commutes with every rounding except in the
    # subnormal range of half (fixed 2^-24 spacing), so compare exactly away from it and to one spacing inside it
    blk2 = CorrBlock(f1, f2)
    blk2.corr_pyramid = [p * 2 for p in blk.corr_pyramid]
    c = grid + 0.37
    a, b = blk2(c).float(), blk(c).float() * 2
    # (a subnormal PRODUCT can round onto an exact tie of the running sum and flip ONE intermediate rounding, so the
    #  bound is one half-ulp step of the largest intermediate, not of the -- possibly cancelled -- result; measured on
    #  B200: 7 of 33.9M elements differ by more than a result-ulp, 1370 differ at all)
    top = 2 * max(p.abs().max().item() for p in blk.corr_pyramid)
    assert (a - b).abs().max().item() <= top * 2.0 ** -10
    assert (a == b).float().mean() > 0.9999


def _ref_ext():
    from oracle import build_ref
    try:
        return build_ref.load()
    except Exception:
        return None


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16])
def test_reference_kernel_matches_oracle_and_ours(dtype):
    """The reference's OWN corr_index_forward_kernel (correlation_kernels.cu compiled into oracle/_ref/ by
    oracle/build_ref.py) vs the CPU oracle and the B200 kernel: this is what pins the bit-exactness recipe."""
    ref = _ref_ext()
    if ref is None:
        pytest.skip('oracle/_ref/droid_corr_ref.so not present (built only where /root/reference is mounted)')
    from pvo_b200 import droid_backends
    g = torch.Generator().manual_seed(17)
    N, h1, w1, h2, w2 = 4, 20, 28, 20, 28
    vol = torch.randn(N, h1, w1, h2, w2, generator=g).to(dtype).cuda()
    coords = _coords(N, h1, w1, 19, spread=6.0, far=True).cuda()
    it = torch.int16 if dtype == torch.float16 else torch.int32
    r, = ref.corr_index_forward(vol, coords, 3)
    o, = droid_backends.corr_index_forward(vol, coords, 3)
    c = ocorr.corr_index_forward(vol.cpu(), coords.cpu(), 3)
    assert torch.equal(r.cpu().view(it), c.view(it)), 'CPU oracle != reference kernel'
    assert torch.equal(r.view(it), o.view(it)), 'B200 kernel != reference kernel'
    # pyramid levels use coords / 2**i
    for i in (1, 2, 3):
        vi = torch.randn(N, h1, w1, h2 >> i, w2 >> i, generator=g).to(dtype).cuda()
        ci = (coords / 2 ** i).contiguous()
        r, = ref.corr_index_forward(vi, ci, 3)
        o, = droid_backends.corr_index_forward(vi, ci, 3)
        assert torch.equal(r.view(it), o.view(it)), i


def test_reference_altcorr_matches_ours():
    ref = _ref_ext()
    if ref is None:
        pytest.skip('oracle/_ref/droid_corr_ref.so not present')
    from pvo_b200 import droid_backends
    g = torch.Generator().manual_seed(23)
    B, H, W, C = 3, 16, 24, 128
    f1 = (torch.randn(B, H, W, C, generator=g) / 4).cuda()
    f2 = (torch.randn(B, H, W, C, generator=g) / 4).cuda()
    coords = _coords(B, H, W, 29, spread=5.0, far=True).permute(0, 2, 3, 1)[:, None].contiguous().cuda()
    r, = ref.altcorr_forward(f1, f2, coords, 3)
    o, = droid_backends.altcorr_forward(f1, f2, coords, 3)
    c = ocorr.altcorr_forward(f1.cpu(), f2.cpu(), coords.cpu(), 3)
    assert torch.allclose(r, o, atol=2e-4, rtol=1e-4)
    assert torch.allclose(r.cpu(), c, atol=2e-4, rtol=1e-4)


# ---------------------------------------------------------------------------------------------------------------------
# tiled private layout (production half path): csrc/corr_lookup_tiled.cu + the TILED epilogue of csrc/corr_volume_tc.cu
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('hw', [(60, 80), (30, 40), (16, 40), (9, 80), (13, 40)])
def test_tiled_volume_build_equals_planar_build_bit_for_bit(hw):
    """the tiled epilogue writes the same values as the planar one (same TMEM accumulators, same rounding cascade), zero pads
    included: tile(planar) == tiled and untile(tiled) == planar, every level"""
    from pvo_b200.modules.corr import build_pyramid, build_pyramid_tiled, tile, untile
    h, w = hw
    g = torch.Generator().manual_seed(14)
    E = 3
    f1 = torch.randn(E, 128, h, w, generator=g).half().cuda()
    f2 = torch.randn(E, 128, h, w, generator=g).half().cuda()
    planar = build_pyramid(f1, f2, 4, use_tc=True)
    # poison the allocator's next blocks so that unwritten pad elements cannot pass as zeros by luck
    junk = [torch.full((E * h * w * (((h >> l) + 3) // 4) * (((w >> l) + 7) // 8) * 32,), 7.0, dtype=torch.float16, device='cuda') for l in range(4)]
    del junk
    tiled = build_pyramid_tiled(f1, f2)
    for l in range(4):
        h2, w2 = h >> l, w >> l
        assert tiled[l].shape == (E, h * w, (h2 + 3) // 4, (w2 + 7) // 8, 32)
        assert torch.equal(untile(tiled[l], h2, w2).view(planar[l].shape).view(torch.int16), planar[l].view(torch.int16)), l
        assert torch.equal(tile(planar[l]).view(tiled[l].shape).view(torch.int16), tiled[l].view(torch.int16)), l


@pytest.mark.parametrize('hw', [(16, 40), (30, 40), (20, 80)])
def test_tiled_lookup_equals_planar_lookup_bit_for_bit(hw):
    """one code path for interior, border-straddling and fully outside windows (zero-filled tiles == skipped taps), both output
    layouts; coordinates include exact integers, far-out values, NaN and inf"""
    from pvo_b200.modules.corr import CorrBlock
    h, w = hw
    g = torch.Generator().manual_seed(15)
    E = 5
    f1 = torch.randn(1, E, 128, h, w, generator=g).half().cuda()
    f2 = torch.randn(1, E, 128, h, w, generator=g).half().cuda()
    tb = CorrBlock(f1, f2)
    assert tb._tiled is not None
    pb = CorrBlock(f1, f2, tiled=False)
    assert pb._tiled is None
    coords = _coords(E, h, w, 33, spread=7.0, far=True).permute(0, 2, 3, 1)[None].contiguous()
    coords[0, 0, 4, :4, 0] = float('nan')
    coords[0, 1, 4, :4, 1] = float('inf')
    coords[0, 2, 4, :4, 0] = -float('inf')
    coords[0, 3, 4, :4] = 3.0e9
    coords[0, 4, 5] = coords[0, 4, 5] * 0 - 0.5
    coords = coords.cuda()
    a, b = tb(coords), pb(coords)
    assert torch.equal(a.view(torch.int16), b.view(torch.int16))
    pa = torch.full((E * h * w, 208), 7.0, dtype=torch.float16, device='cuda')
    pb_ = torch.full((E * h * w, 208), 7.0, dtype=torch.float16, device='cuda')
    tb(coords, out=pa, layout=1, out_cstride=208)
    pb(coords, out=pb_, layout=1, out_cstride=208)
    assert torch.equal(pa.view(torch.int16), pb_.view(torch.int16))
    # the reference-layout attribute, materialised from the tiled store, equals the planar build
    for x, y in zip(tb.corr_pyramid, pb.corr_pyramid):
        assert torch.equal(x.view(torch.int16), y.view(torch.int16))
    # edge bookkeeping on the tiled store
    keep = torch.tensor([True, False, True, True, False], device='cuda')
    tb2 = CorrBlock(f1, f2)[keep]
    assert tb2.num_edges == 3 and torch.equal(tb2(coords[:, keep].contiguous()).view(torch.int16), a[:, keep].contiguous().view(torch.int16))
    tb3 = CorrBlock(f1[:, :2], f2[:, :2]).cat(CorrBlock(f1[:, 2:], f2[:, 2:]))
    assert torch.equal(tb3(coords).view(torch.int16), a.view(torch.int16))


@pytest.mark.parametrize('hw', [(30, 101), (16, 24), (20, 128), (12, 96), (9, 57)])
def test_tiled_volume_and_lookup_at_widths_padded_to_the_tensor_core_tile(hw):
    """image rows that are not 40 / 80 pixels wide (the VKITTI2 grid is 30 x 101) are zero padded to the next tensor-core tile width
    (40 / 80 / 104 / 128): level 0 within 2 ulp of the CUDA-core build, the pooled levels EXACTLY the avg_pool2d cascade of the
    kernel's own level 0 with the reference's floor widths (columns past w >> l are zeros, not pooled pad values), every pad element
    zero, and the lookup bit-identical to the planar kernel on the same values"""
    from pvo_b200.modules.corr import CorrBlock, build_pyramid, build_pyramid_tiled, pad_width, tile, untile
    h, w = hw
    g = torch.Generator().manual_seed(18)
    E = 3
    f1 = torch.randn(E, 128, h, w, generator=g).half().cuda()
    f2 = torch.randn(E, 128, h, w, generator=g).half().cuda()
    wp = pad_width(w)
    junk = torch.full((64 << 20,), 7.0, dtype=torch.float16, device='cuda')
    del junk
    tiled = build_pyramid_tiled(f1, f2)
    cc = build_pyramid(f1, f2, 4, use_tc=False)
    ulp = torch.finfo(torch.float16).eps
    planar = []
    for l in range(4):
        h2, w2 = h >> l, w >> l
        x = untile(tiled[l], h2, w2, wp >> l)
        planar.append(x.view(E, h, w, h2, w2))
        assert torch.equal(tile(x, wp >> l).view(tiled[l].shape).view(torch.int16), tiled[l].view(torch.int16)), 'pad elements of level %d' % l
    assert (planar[0].float() - cc[0].float()).abs().max().item() <= 2 * ulp * max(1.0, cc[0].float().abs().max().item())
    for l in range(3):
        x = planar[l].cpu()
        pooled = ocorr.avg_pool_level(x.reshape(-1, 1, x.shape[3], x.shape[4])).reshape(planar[l + 1].shape)
        assert torch.equal(pooled, planar[l + 1].cpu()), l
    tb = CorrBlock(f1[None], f2[None])
    assert tb._tiled is not None and tb.wpad == wp
    pb = CorrBlock(f1[None], f2[None], tiled=False)
    pb.corr_pyramid = tb.corr_pyramid
    coords = _coords(E, h, w, 35, spread=6.0, far=True).permute(0, 2, 3, 1)[None].contiguous().cuda()
    assert torch.equal(tb(coords).view(torch.int16), pb(coords).view(torch.int16))


def test_corrblock_slot_pool_append_retire_compact():
    """edges appended into free slots, retired without moving a volume, compacted when the tail runs out: always the same lookups as
    a block built from scratch on the surviving edges in the same order"""
    from pvo_b200.modules.corr import CorrBlock
    g = torch.Generator().manual_seed(19)
    h, w, n = 16, 40, 9
    f1 = torch.randn(1, n, 128, h, w, generator=g).half().cuda()
    f2 = torch.randn(1, n, 128, h, w, generator=g).half().cuda()
    coords = _coords(n, h, w, 36, spread=5.0).permute(0, 2, 3, 1)[None].contiguous().cuda()
    blk = CorrBlock(f1[:, :4], f2[:, :4], capacity=6)
    order = [0, 1, 2, 3]
    p0 = blk._tiled[0].data_ptr()
    blk.append(f1[:, 4:6], f2[:, 4:6]); order += [4, 5]
    assert blk._tiled[0].data_ptr() == p0 and blk.num_edges == 6                 # fitted into the pool
    keep = torch.tensor([True, False, True, True, False, True], device='cuda')
    blk = blk[keep]; order = [o for o, k in zip(order, keep.tolist()) if k]
    assert blk._tiled[0].data_ptr() == p0                                        # nothing moved
    blk.append(f1[:, 6:9], f2[:, 6:9]); order += [6, 7, 8]                       # tail exhausted -> compaction into a new pool
    assert blk._tiled[0].data_ptr() != p0 and blk.num_edges == 7
    ref = CorrBlock(f1[:, order], f2[:, order])
    assert torch.equal(blk(coords[:, order]).view(torch.int16), ref(coords[:, order]).view(torch.int16))
    for a, b in zip(blk.corr_pyramid, ref.corr_pyramid):
        assert torch.equal(a.view(torch.int16), b.view(torch.int16))
