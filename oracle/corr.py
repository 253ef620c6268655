"""Correlation volume / pyramid / lookup on CPU -- oracle, test infrastructure only.

Follows (reference file:line, under VO_Module/droid_slam/modules/corr.py unless noted):
  * :63-71    CorrBlock.corr: (fmap1/4)^T (fmap2/4), all pairs
  * :23-38    pyramid by repeated F.avg_pool2d(2, stride=2) (each level rounded to the storage dtype)
  * :40-50    __call__: coords permuted to (E,2,h,w), level i looks up at coords / 2**i, levels concatenated
  * :6-20     CorrSampler -> droid_backends.corr_index_forward (src/correlation_kernels.cu, see corr_ref.c)
  * :91-139   AltCorrBlock
"""
import ctypes
import os
import subprocess

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libcorr_ref.so')
_lib = None


def build_c(force=False):
    """gcc the C restatement into oracle/_build/ (generic x86-64 so the .so runs on the GPU box too)."""
    src = os.path.join(_HERE, 'corr_ref.c')
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(['gcc', '-O2', '-fPIC', '-shared', '-fopenmp', '-ffp-contract=off',
                               '-std=gnu11', src, '-o', _SO, '-lm'])
    return _SO


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build_c()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def corr_index_forward(volume, coords, radius):
    """volume (N,h1,w1,h2,w2) float32|float16 ; coords (N,2,h1,w1) float32 -> (N,rd,rd,h1,w1)."""
    lib = _load()
    v = np.ascontiguousarray(volume.detach().cpu().numpy())
    c = np.ascontiguousarray(coords.detach().cpu().numpy().astype(np.float32))
    N, h1, w1, h2, w2 = v.shape
    rd = 2 * radius + 1
    out = np.empty((N, rd, rd, h1, w1), dtype=v.dtype)
    fn = {np.dtype(np.float32): lib.corr_index_forward_f32,
          np.dtype(np.float16): lib.corr_index_forward_f16}[v.dtype]
    fn(_ptr(v), _ptr(c), _ptr(out), N, h1, w1, h2, w2, radius)
    return torch.from_numpy(out)


def altcorr_forward(fmap1, fmap2, coords, radius):
    """fmap1 (B,H1,W1,C), fmap2 (B,H2,W2,C), coords (B,S,H1,W1,2) float32 -> (B,S,rd*rd,H1,W1)."""
    lib = _load()
    a = np.ascontiguousarray(fmap1.detach().cpu().float().numpy())
    b = np.ascontiguousarray(fmap2.detach().cpu().float().numpy())
    c = np.ascontiguousarray(coords.detach().cpu().float().numpy())
    B, H1, W1, C = a.shape
    _, H2, W2, _ = b.shape
    S = c.shape[1]
    rd = 2 * radius + 1
    out = np.empty((B, S, rd * rd, H1, W1), dtype=np.float32)
    lib.altcorr_forward_f32(_ptr(a), _ptr(b), _ptr(c), _ptr(out), B, S, H1, W1, H2, W2, C, radius)
    return torch.from_numpy(out)


def corr_volume(fmap1, fmap2):
    """(B,E,C,h,w) x2 -> (B,E,h,w,h,w); arithmetic dtype = input dtype (half inputs -> float32
    accumulate, rounded once to half, which is what a tensor-core / cuBLAS half GEMM produces)."""
    batch, num, dim, ht, wd = fmap1.shape
    dt = fmap1.dtype
    a = (fmap1.reshape(batch * num, dim, ht * wd) / 4.0)
    b = (fmap2.reshape(batch * num, dim, ht * wd) / 4.0)
    corr = torch.matmul(a.float().transpose(1, 2), b.float()).to(dt)
    return corr.view(batch, num, ht, wd, ht, wd)


def avg_pool_level(corr):
    """One F.avg_pool2d(2,2) step on (M,1,h2,w2): float32 accumulate in (dy,dx) raster order, /4, round."""
    dt = corr.dtype
    return F.avg_pool2d(corr.float(), 2, stride=2).to(dt)


def corr_pyramid(fmap1, fmap2, num_levels=4):
    corr = corr_volume(fmap1, fmap2)
    batch, num, h1, w1, h2, w2 = corr.shape
    corr = corr.reshape(batch * num * h1 * w1, 1, h2, w2)
    pyr = []
    for i in range(num_levels):
        pyr.append(corr.view(batch * num, h1, w1, h2 // 2 ** i, w2 // 2 ** i))
        corr = avg_pool_level(corr)
    return pyr


def corr_lookup(pyramid, coords, radius=3):
    """CorrBlock.__call__: coords (B,E,h,w,2) -> (B,E,L*(2r+1)^2,h,w)."""
    batch, num, ht, wd, _ = coords.shape
    c = coords.permute(0, 1, 4, 2, 3).contiguous().view(batch * num, 2, ht, wd)
    outs = []
    for i, vol in enumerate(pyramid):
        o = corr_index_forward(vol, c / 2 ** i, radius)
        outs.append(o.view(batch, num, -1, ht, wd))
    return torch.cat(outs, dim=2)


def alt_corr_lookup(fmaps, coords, ii, jj, num_levels=4, radius=3):
    """AltCorrBlock(fmaps)(coords, ii, jj) with coords (B,E,h,w,2): -> (B,E,L*49,h,w) float32."""
    B, N, C, H, W = fmaps.shape
    f = fmaps.reshape(B * N, C, H, W).float() / 4.0
    pyr = []
    for i in range(num_levels):
        pyr.append(f.permute(0, 2, 3, 1).contiguous().view(B, N, H // 2 ** i, W // 2 ** i, C))
        f = F.avg_pool2d(f, 2, stride=2)
    E = ii.shape[0]
    outs = []
    for i in range(num_levels):
        f1 = pyr[0][:, ii].reshape((B * E,) + pyr[0].shape[2:])
        f2 = pyr[i][:, jj].reshape((B * E,) + pyr[i].shape[2:])
        c = (coords / 2 ** i).reshape(B * E, 1, H, W, 2).contiguous()
        o = altcorr_forward(f1, f2, c, radius)                   # (B*E,1,49,H,W)
        outs.append(o.view(B, E, -1, H, W))
    return torch.cat(outs, dim=2)
