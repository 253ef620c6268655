"""Drop-in for the reference's pybind module ``droid_backends`` (VO_Module/src/droid.cpp:234-247).

Same function names, argument order, in-place behaviour and return arity (callers unpack ``corr, = ...``,
modules/corr.py:12,19,79), implemented as thin ctypes calls into libpvo_b200.so.  Outputs are allocated here with
torch.empty (the kernels write every element; the reference's torch::zeros + read-modify-write is gone).
"""
import torch

from . import _lib
from ._lib import check, ptr, stream, require_cuda

_ba_ws = {}


def corr_index_forward(volume, coords, radius):
    """src/correlation_kernels.cu:126-155.  volume (N,h1,w1,h2,w2) f32|f16; coords (N,2,h1,w1) f32."""
    require_cuda(volume, coords)
    if coords.dtype != torch.float32:
        raise RuntimeError('corr_index_forward: coords must be float32')
    N, h1, w1, h2, w2 = volume.shape
    rd = 2 * radius + 1
    out = torch.empty(N, rd, rd, h1, w1, dtype=volume.dtype, device=volume.device)
    check(_lib.lib().pvo_corr_index_forward(ptr(volume), ptr(coords), ptr(out), N, h1, w1, h2, w2, radius,
                                            _lib.dtype_code(volume), 0, stream()), 'corr_index_forward')
    return [out]


def corr_index_backward(volume, coords, corr_grad, radius):
    """src/correlation_kernels.cu:157-185."""
    require_cuda(volume, coords, corr_grad)
    N, h1, w1, h2, w2 = volume.shape
    grad = torch.zeros_like(volume)
    check(_lib.lib().pvo_corr_index_backward(ptr(coords), ptr(corr_grad), ptr(grad), N, h1, w1, h2, w2, radius,
                                             _lib.dtype_code(volume), stream()), 'corr_index_backward')
    return [grad]


def altcorr_forward(fmap1, fmap2, coords, radius):
    """src/altcorr_kernel.cu:290-319.  fmap1 (B,H1,W1,C), fmap2 (B,H2,W2,C) f32; coords (B,S,H1,W1,2)."""
    require_cuda(fmap1, fmap2, coords)
    if fmap1.dtype != torch.float32 or fmap2.dtype != torch.float32:
        raise RuntimeError('altcorr_forward: float32 feature maps only (the reference caller forces .float(), corr.py:120)')
    B, H1, W1, C = fmap1.shape
    _, H2, W2, _ = fmap2.shape
    S = coords.shape[1]
    rd = 2 * radius + 1
    out = torch.empty(B, S, rd * rd, H1, W1, dtype=torch.float32, device=fmap1.device)
    check(_lib.lib().pvo_altcorr_forward(ptr(fmap1), ptr(fmap2), ptr(coords), ptr(out), B, S, H1, W1, H2, W2, C, radius,
                                         stream()), 'altcorr_forward')
    return [out]


def altcorr_backward(fmap1, fmap2, coords, corr_grad, radius):
    raise NotImplementedError('altcorr_backward (training only, src/altcorr_kernel.cu:152-286) is outside the '
                              'inference hot path -- SURVEY.md section 8(f) item 4')


def frame_distance(poses, disps, intrinsics, ii, jj, beta):
    """src/droid_kernels.cu:1414-1436.  intrinsics (4,)."""
    require_cuda(poses, disps, intrinsics, ii, jj)
    n = ii.shape[0]
    dist = torch.empty(n, dtype=torch.float32, device=poses.device)
    check(_lib.lib().pvo_frame_distance(ptr(poses), ptr(disps), ptr(intrinsics), ptr(ii), ptr(jj), ptr(dist), n,
                                        disps.shape[1], disps.shape[2], float(beta), stream()), 'frame_distance')
    return dist


def ba_workspace(E, nframes, P, K, HW, device):
    need = _lib.lib().pvo_ba_workspace_bytes(E, nframes, P, K, HW)
    if need < 0:
        raise RuntimeError('ba: invalid problem size')
    key = (str(device),)
    ws = _ba_ws.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(int(need * 1.25) + 1024, dtype=torch.uint8, device=device)
        _ba_ws[key] = ws
    return ws


def band_flag(max_span):
    """pvo_ba flags bits 20-27: 1 + block bandwidth of the reduced camera system (0 = unknown)"""
    return 0 if max_span is None else ((min(2 * int(max_span), 254) + 1) << 20)


def ba(poses, disps, intrinsics, targets, weights, eta, ii, jj, t0, t1, iterations, lm, ep, motion_only,
       interleaved=False, reuse_plan=False, workspace=None, max_degree=0, max_span=None):
    """src/droid_kernels.cu:1293-1410.  Mutates poses (buf,7) and disps (buf,ht,wd) in place; returns [dx, dz].

    targets / weights: (E,2,ht,wd) as the reference passes them (factor_graph.py:297-300), or (E,ht,wd,2) with
    ``interleaved=True`` (saves the permute+contiguous copies).  eta: (K,ht,wd).
    max_span: optional max |ii - jj| over the WHOLE graph (all shards): the pose system then has block bandwidth <= 2 max_span
    and the solver skips its own sparsity scan."""
    require_cuda(poses, disps, intrinsics, targets, weights, eta, ii, jj)
    if ii.dtype != torch.long or jj.dtype != torch.long:
        raise RuntimeError('ba: ii/jj must be int64')
    E = ii.shape[0]
    nframes, ht, wd = disps.shape
    K = eta.shape[0]
    P = t1 - t0
    ws = workspace if workspace is not None else ba_workspace(E, nframes, P, K, ht * wd, poses.device)
    dx = torch.empty(P, 6, dtype=torch.float32, device=poses.device)
    dz = torch.empty(K, ht * wd, dtype=torch.float32, device=poses.device)
    check(_lib.lib().pvo_ba(ptr(poses), ptr(disps), ptr(intrinsics), ptr(targets), ptr(weights), ptr(eta), ptr(ii),
                            ptr(jj), E, nframes, ht, wd, int(t0), int(t1), K, int(iterations), float(lm), float(ep),
                            1 if motion_only else 0, 1 if interleaved else 0, ptr(dx), ptr(dz), ptr(ws), ws.numel(),
                            (1 if reuse_plan else 0) | ((int(max_degree) & 0xFFF) << 8) | band_flag(max_span), stream()), 'ba',
          launches=(0 if reuse_plan else 1) + int(iterations) * (3 if motion_only else 4))
    return [dx, dz]


def _out_of_scope(name, where):
    def fn(*a, **k):
        raise NotImplementedError('%s (%s) is not on the VO hot path -- SURVEY.md section 2.2 marks it out of scope' % (name, where))
    return fn


projmap = _out_of_scope('projmap', 'src/droid_kernels.cu:406-495, never called from Python')
depth_filter = _out_of_scope('depth_filter', 'src/droid_kernels.cu:640-754, visualiser only')
iproj = _out_of_scope('iproj', 'src/droid_kernels.cu:758-829, visualiser only')
