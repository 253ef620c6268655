"""CorrSampler / CorrBlock / CorrLayer / AltCorrBlock with the reference's interface
(VO_Module/droid_slam/modules/corr.py), on the B200 kernels.

Differences that are invisible to callers: the volume + pyramid of CorrBlock.__init__ is one fused launch
(tcgen05 GEMM for half, CUDA cores for float32); __call__ looks up all four levels in one launch and writes the
concatenated (B,E,196,h,w) tensor directly (no per-level zeros + cat).  On the production half path the pyramid is kept in
a TILED private layout (tiles of 4 rows x 8 columns = one 64-byte DRAM access, csrc/corr_lookup_tiled.cu); the reference-layout
`corr_pyramid` attribute is then materialised on demand (and assigning to it switches the block to the planar kernels)."""
import torch

from .. import _lib, droid_backends
from .._lib import check, ptr, stream


class CorrSampler(torch.autograd.Function):
    """Differentiable window lookup of ONE pyramid level (reference interface: corr.py:6-20): forward and backward are the
    pvo_corr_index_forward / _backward kernels."""

    @staticmethod
    def forward(ctx, volume, coords, radius):
        out = droid_backends.corr_index_forward(volume, coords, radius)[0]
        ctx.radius = radius
        ctx.save_for_backward(volume, coords)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        vol, xy = ctx.saved_tensors
        return droid_backends.corr_index_backward(vol, xy, grad_out.contiguous(), ctx.radius)[0], None, None


def to_pixel_major(x):
    """(N,C,H,W) -> (N,H*W,C) contiguous, same dtype, through the transpose kernel."""
    N, C, H, W = x.shape
    out = torch.empty(N, H * W, C, dtype=x.dtype, device=x.device)
    check(_lib.lib().pvo_nchw_to_nhwc(ptr(x), _lib.dtype_code(x), ptr(out), _lib.dtype_code(out), N, C, H * W, C, 0,
                                      stream()), 'nchw_to_nhwc')
    return out


TC_WIDTHS = (40, 80, 104, 128)          # template widths of the tensor-core volume kernel (UMMA N = 2 * width)


def pad_width(w):
    """smallest tensor-core tile width that holds an image row of w pixels, or None"""
    for t in TC_WIDTHS:
        if w <= t:
            return t
    return None


def tc_volume_supported(fmap, num_levels=4):
    N, C, h, w = fmap.shape
    return fmap.dtype == torch.float16 and C == 128 and pad_width(w) is not None and h >= 8 and num_levels == 4


def tiled_shape(N, h, w, l, wpad=None):
    """shape of level l of the tiled pyramid for N plane blocks (edges) of an (h, w) grid built at template width wpad"""
    wpad = wpad or pad_width(w)
    return (N, h * w, ((h >> l) + 3) // 4, ((wpad >> l) + 7) // 8, 32)


def build_pyramid_tiled(fmap1, fmap2, out=None, slots=None):
    """fmap1, fmap2: (N,128,h,w) half -> 4 tiled levels (cap, h*w, ceil(h_l/4), ceil(wpad_l/8), 32), one tcgen05 launch.
    out / slots: write edge e into plane block slots[e] of the given level tensors (a CorrBlock's pool)."""
    _lib.require_cuda(fmap1, fmap2)
    N, C, h, w = fmap1.shape
    wpad = pad_width(w)
    f1 = to_pixel_major(fmap1 / 4.0)
    f2 = fmap2 / 4.0
    if wpad != w:                      # zero features in the pad columns: their correlations are exact zeros
        f2 = torch.nn.functional.pad(f2, (0, wpad - w))
    f2 = to_pixel_major(f2)
    levels = out if out is not None else [torch.empty(tiled_shape(N, h, w, l, wpad), dtype=torch.float16, device=fmap1.device) for l in range(4)]
    check(_lib.lib().pvo_corr_volume_f16_tiled(ptr(f1), ptr(f2), _lib.ptr_array(levels), ptr(slots), N, h, w, wpad, 4, stream()),
          'corr_volume_f16_tiled')
    return levels


def untile(level, h2, w2, w2pad=None):
    """(N, HW, TR, TC, 32) -> (N, HW, h2, w2) planar"""
    N, HW = level.shape[:2]
    out = torch.empty(N, HW, h2, w2, dtype=level.dtype, device=level.device)
    check(_lib.lib().pvo_corr_untile(ptr(level), ptr(out), N * HW, h2, w2, w2pad or w2, stream()), 'corr_untile')
    return out


def tile(level, w2pad=None):
    """(..., h2, w2) planar half -> (prod(...), ceil(h2/4), ceil(w2pad/8), 32) tiled, zero padded"""
    h2, w2 = level.shape[-2:]
    w2pad = w2pad or w2
    n = level.numel() // (h2 * w2)
    out = torch.empty(n, (h2 + 3) // 4, (w2pad + 7) // 8, 32, dtype=level.dtype, device=level.device)
    check(_lib.lib().pvo_corr_tile(ptr(level.contiguous()), ptr(out), n, h2, w2, w2pad, stream()), 'corr_tile')
    return out


def build_pyramid(fmap1, fmap2, num_levels=4, use_tc=None):
    """fmap1, fmap2: (N,C,h,w).  Returns [ (N,h,w,h>>l,w>>l) ] -- corr.py:23-38,63-71 in one call."""
    _lib.require_cuda(fmap1, fmap2)
    N, C, h, w = fmap1.shape
    if use_tc is None:
        use_tc = fmap1.dtype == torch.float16 and C == 128 and w in (40, 80) and h >= 8 and num_levels == 4      # planar: exact widths only
    if use_tc:      # the tensor-core kernel takes the reference's `fmap / 4.0` (a half op, corr.py:66-67) as input
        f1, f2 = to_pixel_major(fmap1 / 4.0), to_pixel_major(fmap2 / 4.0)
    else:           # the CUDA-core kernel applies the same rounding while loading
        f1, f2 = to_pixel_major(fmap1), to_pixel_major(fmap2)
    levels = [torch.empty(N, h, w, h >> l, w >> l, dtype=fmap1.dtype, device=fmap1.device) for l in range(num_levels)]
    arr = _lib.ptr_array(levels)
    if use_tc:
        check(_lib.lib().pvo_corr_volume_f16(ptr(f1), ptr(f2), arr, N, h, w, num_levels, stream()), 'corr_volume_f16')
    else:
        check(_lib.lib().pvo_corr_volume(ptr(f1), ptr(f2), arr, N, h, w, C, num_levels, _lib.dtype_code(fmap1),
                                         stream()), 'corr_volume')
    return levels


class CorrBlock:
    """corr.py:23-71.

    Production half path: the pyramid lives in a POOL of tiled plane blocks (one block = the four levels of one edge) addressed
    through an edge -> slot table.  New edges are built straight into free slots (`append`), retiring edges only drops their table
    entries (`block[mask]`): the per-frame loop adds and removes a few edges of a 48-edge window every frame, and with the
    reference's representation (`torch.cat` / boolean indexing of the whole (E, h, w, h, w) tensors, factor_graph.py:128-131,
    185-187) that copied the entire 1.2 GB volume several times per frame (4.2 ms of 22.5 ms GPU time per frame at the VKITTI
    shape, tools/prof_loop.py).  Slots are handed out from the tail; when it is exhausted the live blocks are compacted into a
    pool twice the live size (amortised)."""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=3, use_tc=None, tiled=None, capacity=None):
        self.num_levels = num_levels
        self.radius = radius
        batch, num, dim, ht, wd = fmap1.shape
        if fmap1.dtype not in (torch.float16, torch.float32):
            raise RuntimeError('CorrBlock: float16 / float32 feature maps only')
        f1 = fmap1.reshape(batch * num, dim, ht, wd).contiguous()
        f2 = fmap2.reshape(batch * num, dim, ht, wd).contiguous()
        self.ht, self.wd = ht, wd
        can_tile = use_tc is not False and radius == 3 and tc_volume_supported(f1, num_levels)
        if tiled is None:
            tiled = can_tile
        if tiled and not can_tile:
            raise RuntimeError('CorrBlock: the tiled layout needs half features, 128 channels, h >= 8, w <= 128, 4 levels, radius 3')
        self._tiled = self._slots = self._planar = None
        self._tail = 0
        if tiled:
            E = batch * num
            self.wpad = pad_width(wd)
            self._alloc_pool(max(E, capacity or 0), fmap1.device)
            self._slots = torch.arange(E, dtype=torch.int32, device=fmap1.device)
            self._tail = E
            build_pyramid_tiled(f1, f2, out=self._tiled, slots=self._slots)
        else:
            self._planar = build_pyramid(f1, f2, num_levels, use_tc)

    # ------------------------------------------------------------------ pool
    def _alloc_pool(self, cap, device):
        self._cap = cap
        self._tiled = [torch.empty(tiled_shape(cap, self.ht, self.wd, l, self.wpad), dtype=torch.float16, device=device) for l in range(4)]

    def _reserve(self, k):
        """make room for k more blocks at the tail: compact the live blocks into a fresh pool when the tail is exhausted"""
        if self._tail + k <= self._cap:
            return
        live = self._slots.long()
        old = self._tiled
        n = int(live.shape[0])
        self._alloc_pool(max(2 * (n + k), 8), old[0].device)
        for dst, src in zip(self._tiled, old):
            if n:
                torch.index_select(src, 0, live, out=dst[:n])
        self._slots = torch.arange(n, dtype=torch.int32, device=old[0].device)
        self._tail = n

    def append(self, fmap1, fmap2):
        """volumes of new edges (fmap1, fmap2: (1,k,128,h,w)) built directly into free slots; returns self"""
        if self._tiled is None:
            return self.cat(CorrBlock(fmap1, fmap2, self.num_levels, self.radius, tiled=False))
        batch, num, dim, ht, wd = fmap1.shape
        k = batch * num
        self._reserve(k)
        new = torch.arange(self._tail, self._tail + k, dtype=torch.int32, device=self._slots.device)
        build_pyramid_tiled(fmap1.reshape(k, dim, ht, wd).contiguous(), fmap2.reshape(k, dim, ht, wd).contiguous(), out=self._tiled, slots=new)
        self._slots = torch.cat([self._slots, new])
        self._tail += k
        self._planar = None
        return self

    def _live(self, l):
        """level l of the live edges, in edge order: (E, HW, TR, TC, 32)"""
        return torch.index_select(self._tiled[l], 0, self._slots.long())

    @property
    def corr_pyramid(self):
        """the reference's attribute (corr.py:29-38): [ (N, h, w, h>>l, w>>l) ]; converted from the tiled store when needed"""
        if self._planar is None:
            self._planar = [untile(self._live(l), self.ht >> l, self.wd >> l, self.wpad >> l).view(-1, self.ht, self.wd, self.ht >> l, self.wd >> l)
                            for l in range(4)]
        return self._planar

    @corr_pyramid.setter
    def corr_pyramid(self, levels):
        self._planar = list(levels)
        self._tiled = self._slots = None               # a caller-supplied pyramid is used as is, through the planar kernels

    def _lookup_tiled(self, coords, batch, num, ht, wd, out, layout, out_cstride):
        if batch * num != self._slots.shape[0]:
            raise RuntimeError('CorrBlock: %d coordinate maps for %d edges' % (batch * num, self._slots.shape[0]))
        if out is None:
            out = torch.empty(batch, num, 196, ht, wd, dtype=torch.float16, device=coords.device)
        check(_lib.lib().pvo_corr_lookup_tiled(_lib.ptr_array(self._tiled), _lib.int_array([ht >> l for l in range(4)]),
                                               _lib.int_array([self.wpad >> l for l in range(4)]), 4, ptr(self._slots), ptr(coords),
                                               ptr(out), batch * num, ht, wd, 3, layout, out_cstride, stream()), 'corr_lookup_tiled')
        return out

    def __call__(self, coords, out=None, layout=0, out_cstride=0):
        """coords (B,E,h,w,2) float32 -> (B,E,L*49,h,w) in the volume's dtype (corr.py:40-50).
        layout=1 writes pixel-major (B*E*h*w, out_cstride) into ``out`` for the fused update operator."""
        batch, num, ht, wd, _ = coords.shape
        coords = coords.contiguous().float()
        if self._tiled is not None and not coords.requires_grad:
            return self._lookup_tiled(coords, batch, num, ht, wd, out, layout, out_cstride)
        if self.radius != 3 or coords.requires_grad or any(p.requires_grad for p in self.corr_pyramid):
            return self._call_per_level(coords)
        pyr = self.corr_pyramid
        L = self.num_levels
        if out is None:
            out = torch.empty(batch, num, L * 49, ht, wd, dtype=pyr[0].dtype, device=coords.device)
        check(_lib.lib().pvo_corr_lookup_pyramid(_lib.ptr_array(pyr), _lib.int_array([p.shape[3] for p in pyr]),
                                                 _lib.int_array([p.shape[4] for p in pyr]), L, ptr(coords), ptr(out),
                                                 batch * num, ht, wd, self.radius, _lib.dtype_code(pyr[0]), layout,
                                                 out_cstride, 0, stream()), 'corr_lookup_pyramid')
        return out

    def _call_per_level(self, coords):
        """The reference's literal per-level path (differentiable through CorrSampler)."""
        out_pyramid = []
        batch, num, ht, wd, _ = coords.shape
        c = coords.permute(0, 1, 4, 2, 3).contiguous().view(batch * num, 2, ht, wd)
        for i in range(self.num_levels):
            corr = CorrSampler.apply(self.corr_pyramid[i], (c / 2 ** i).contiguous(), self.radius)
            out_pyramid.append(corr.view(batch, num, -1, ht, wd))
        return torch.cat(out_pyramid, dim=2)

    @property
    def dtype(self):
        return (self._tiled or self._planar)[0].dtype

    @property
    def num_edges(self):
        return self._slots.shape[0] if self._tiled is not None else self._planar[0].shape[0]

    def cat(self, other):
        """reference interface (corr.py:52-56): append the edges of another block (copies the other block's volumes)"""
        if self._tiled is not None and other._tiled is not None and other.wpad == self.wpad:
            k = other.num_edges
            self._reserve(k)
            for l in range(4):
                self._tiled[l][self._tail:self._tail + k] = other._live(l)
            self._slots = torch.cat([self._slots, torch.arange(self._tail, self._tail + k, dtype=torch.int32, device=self._slots.device)])
            self._tail += k
            self._planar = None
        else:
            self.corr_pyramid = [torch.cat([a, b], 0) for a, b in zip(self.corr_pyramid, other.corr_pyramid)]
        return self

    def __getitem__(self, index):
        """reference interface (corr.py:58-61): keep the edges selected by a mask / index (no volume is moved on the tiled path)"""
        if self._tiled is not None:
            self._slots = self._slots[index].contiguous()
            self._planar = None
        else:
            self._planar = [t[index].contiguous() for t in self._planar]
        return self

    @staticmethod
    def corr(fmap1, fmap2):
        """all-pairs correlation (corr.py:63-71): level 0 only, (B,E,h,w,h,w)."""
        batch, num, dim, ht, wd = fmap1.shape
        lv = build_pyramid(fmap1.reshape(batch * num, dim, ht, wd).contiguous(),
                           fmap2.reshape(batch * num, dim, ht, wd).contiguous(), 1)
        return lv[0].view(batch, num, ht, wd, ht, wd)


class CorrLayer(torch.autograd.Function):
    """On-the-fly correlation of one pyramid level (reference: corr.py:74-88 over altcorr_kernel.cu), differentiable with respect to
    both feature maps (coords get a zero gradient, as in the reference)."""

    @staticmethod
    def forward(ctx, fmap1, fmap2, coords, r):
        out = droid_backends.altcorr_forward(fmap1, fmap2, coords, r)[0]
        ctx.save_for_backward(fmap1, fmap2, coords)
        ctx.r = r
        return out

    @staticmethod
    def backward(ctx, grad_corr):
        g1, g2, gc = droid_backends.altcorr_backward(*ctx.saved_tensors, grad_corr.contiguous(), ctx.r)
        return g1, g2, gc, None


class AltCorrBlock:
    """Correlation features without a materialised volume (global-BA path; reference interface corr.py:91-139):
    ``AltCorrBlock(fmaps)(coords, ii, jj)`` with fmaps (B,N,C,H,W), coords (B,E,H,W,2) or (B,E,H,W,S,2) -> (B,E,L*49,H,W[,S])."""

    def __init__(self, fmaps, num_levels=4, radius=3):
        self.num_levels, self.radius = num_levels, radius
        B, N, C, H, W = fmaps.shape
        level = fmaps.reshape(B * N, C, H, W) / 4.0                     # the same pre-scaling as the volume path
        self.pyramid = []                                              # channels-last float maps, one per level
        for _ in range(num_levels):
            self.pyramid.append(level.permute(0, 2, 3, 1).contiguous().view(B, N, level.shape[2], level.shape[3], C))
            level = torch.nn.functional.avg_pool2d(level, 2, stride=2)

    def _level(self, lvl, src, dst, xy):
        """xy (B,E,S,H,W,2) in level-0 pixels -> (B,E,49,H,W,S) correlation window of pyramid level `lvl`"""
        B, E, S, H, W, _ = xy.shape
        flat = lambda t: t.reshape((B * E,) + tuple(t.shape[2:])).float().contiguous()
        win = CorrLayer.apply(flat(self.pyramid[0][:, src]), flat(self.pyramid[lvl][:, dst]),
                              (xy / float(1 << lvl)).reshape(B * E, S, H, W, 2).contiguous(), self.radius)
        return win.view(B, E, S, -1, H, W).permute(0, 1, 3, 4, 5, 2)

    def corr_fn(self, coords, ii, jj):
        xy = coords.permute(0, 1, 4, 2, 3, 5)                           # (B,E,H,W,S,2) -> (B,E,S,H,W,2)
        return torch.cat([self._level(l, ii, jj, xy) for l in range(self.num_levels)], dim=2)

    def __call__(self, coords, ii, jj):
        single = coords.dim() == 5                                     # one sample per pixel: add / drop the S axis
        out = self.corr_fn(coords.unsqueeze(-2) if single else coords, ii, jj)
        return (out.squeeze(-1) if single else out).contiguous()
