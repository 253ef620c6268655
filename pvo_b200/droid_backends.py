"""Drop-in for the reference's pybind module ``droid_backends`` (VO_Module/src/droid.cpp:234-247).

Same function names, argument order, in-place behaviour and return arity (callers unpack ``corr, = ...``,
modules/corr.py:12,19,79), implemented as thin ctypes calls into libpvo_b200.so.  Outputs are allocated here with
torch.empty (the kernels write every element; the reference's torch::zeros + read-modify-write is gone).
"""
import os

import torch

from . import _lib
from ._lib import check, ptr, stream, require_cuda

_ba_ws = {}
_ba_plan = [None]           # who wrote the schedule ("plan") that currently sits in the shared BA workspace


def ba_plan_owner():
    return _ba_plan[0]


def set_ba_plan_owner(token):
    """A caller that wants `reuse_plan=True` on its next call records a token after its call and compares it before the next:
    any other BA call in between (another graph, another window) rewrites the workspace and clears the token."""
    _ba_plan[0] = token


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
    """src/altcorr_kernel.cu:152-286, :321-356 (training path) -> [fmap1_grad, fmap2_grad, coords_grad (zeros, as the reference)]"""
    require_cuda(fmap1, fmap2, coords, corr_grad)
    if fmap1.dtype != torch.float32 or corr_grad.dtype != torch.float32:
        raise RuntimeError('altcorr_backward: float32 only (the reference instantiates its kernel for float)')
    B, H1, W1, C = fmap1.shape
    _, H2, W2, _ = fmap2.shape
    S = coords.shape[1]
    g1 = torch.empty_like(fmap1)
    g2 = torch.zeros_like(fmap2)
    check(_lib.lib().pvo_altcorr_backward(ptr(fmap1.contiguous()), ptr(fmap2.contiguous()), ptr(coords.contiguous()), ptr(corr_grad.contiguous()),
                                          ptr(g1), ptr(g2), B, S, H1, W1, H2, W2, C, radius, stream()), 'altcorr_backward')
    return [g1, g2, torch.zeros_like(coords)]


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


def current_workspace(device):
    """the shared BA workspace of `device` (None before the first call): a captured CUDA graph keeps a reference to the buffer it
    was captured on, so that growth elsewhere cannot free it under the graph"""
    return _ba_ws.get((str(device),))


def band_flag(max_span):
    """pvo_ba flags bits 20-27: 1 + block bandwidth of the reduced camera system (0 = unknown)"""
    return 0 if max_span is None else ((min(2 * int(max_span), 254) + 1) << 20)


def ba(poses, disps, intrinsics, targets, weights, eta, ii, jj, t0, t1, iterations, lm, ep, motion_only,
       interleaved=False, reuse_plan=False, workspace=None, max_degree=0, max_span=None, sys_clean=False):
    """src/droid_kernels.cu:1293-1410.  Mutates poses (buf,7) and disps (buf,ht,wd) in place; returns [dx, dz].

    targets / weights: (E,2,ht,wd) as the reference passes them (factor_graph.py:297-300), or (E,ht,wd,2) with
    ``interleaved=True`` (saves the permute+contiguous copies).  eta: (K,ht,wd).
    max_span: optional max |ii - jj| over the WHOLE graph (all shards): the pose system then has block bandwidth <= 2 max_span
    and the solver skips its own sparsity scan.
    workspace / sys_clean: a caller-owned workspace (zero-initialised once, used by nobody else) whose system replicas are therefore
    known to be zero on entry -- every call leaves them zeroed -- so the schedule kernel need not clear them (CUDA-graph replays)."""
    require_cuda(poses, disps, intrinsics, targets, weights, eta, ii, jj)
    if not reuse_plan:
        _ba_plan[0] = None
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
                            (1 if reuse_plan else 0) | (16 if sys_clean else 0) | ((int(max_degree) & 0xFFF) << 8) | band_flag(max_span), stream()), 'ba',
          launches=(0 if reuse_plan else 1) + int(iterations) * (3 if motion_only else 4))
    if CHECK_BA:
        check_ba_meta(ws, E, nframes, P, K, ht * wd)
    return [dx, dz]


CHECK_BA = bool(int(os.environ.get('PVO_BA_CHECK', '0')))      # tests switch it on: read the device-side error flag after every call


def check_ba_meta(ws, E, nframes, P, K, HW):
    """The schedule kernel records what it could not honour: flag 1 = the number of depth frames unique(arange(t0,t1) U ii) differs
    from eta's row count (the reference raises a shape error at `C + eta`, droid_kernels.cu:1374), flag 2 = a frame's Schur tile
    did not fit in shared memory.  Reading it synchronises, so it is opt-in (PVO_BA_CHECK=1 / droid_backends.CHECK_BA = True)."""
    import ctypes
    meta = (ctypes.c_int * 3)()
    check(_lib.lib().pvo_ba_read_meta(ptr(ws), E, nframes, P, K, HW, meta), 'ba_read_meta', launches=0)
    if meta[1] == 1:
        raise RuntimeError('ba: eta has %d rows but the window and edges span %d depth frames' % (K, meta[0]))
    if meta[1] == 2:
        raise RuntimeError('ba: out-degree %d of a keyframe exceeds the shared-memory Schur tile' % meta[2])


# ------------------------------------------------------------------------------------------------------------------
# The three visualiser-side entry points of the module (src/droid.cpp:234-247).  None is on the per-iteration path
# (visualization.py:98-104 calls iproj / depth_filter when a keyframe is redrawn; projmap has no Python caller), so they are
# tensor expressions over the pose / depth buffers -- a few launches of torch kernels on whatever device the buffers live on --
# not entries of the C ABI.  Same arguments, return arity and quirks as the reference kernels.
def _rel(poses, ix, jx):
    """relSE3 (droid_kernels.cu:96-107): quaternions as stored"""
    ti, qi, tj, qj = poses[ix, :3], poses[ix, 3:], poses[jx, :3], poses[jx, 3:]
    xi, yi, zi, wi = qi.unbind(-1)
    xj, yj, zj, wj = qj.unbind(-1)
    qij = torch.stack([-wj * xi + xj * wi - yj * zi + zj * yi, -wj * yi + yj * wi - zj * xi + xj * zi,
                       -wj * zi + zj * wi - xj * yi + yj * xi, wj * wi + xj * xi + yj * yi + zj * zi], -1)
    return tj - _rot(qij, ti), qij


def _rot(q, X):
    """actSO3 (droid_kernels.cu:58-69), q (...,4) broadcast against X (...,3)"""
    v, w = q[..., :3], q[..., 3:]
    uv = 2.0 * torch.linalg.cross(v.expand_as(X), X, dim=-1)
    return X + w * uv + torch.linalg.cross(v.expand_as(X), uv, dim=-1)


def _rays(intrinsics, ht, wd):
    fx, fy, cx, cy = intrinsics.reshape(-1)[:4].unbind(0)
    v, u = torch.meshgrid(torch.arange(ht, dtype=torch.float32, device=intrinsics.device),
                          torch.arange(wd, dtype=torch.float32, device=intrinsics.device), indexing='ij')
    return torch.stack([(u - cx) / fx, (v - cy) / fy, torch.ones_like(u)], -1), u, v, (fx, fy, cx, cy)


def _iproj(poses, disps, intrinsics):
    n, ht, wd = disps.shape
    X, _, _, _ = _rays(intrinsics, ht, wd)
    d = disps[..., None]
    Y = _rot(poses[:n, None, None, 3:], X[None].expand(n, ht, wd, 3)) + d * poses[:n, None, None, :3]
    return Y / d


def _projmap(poses, disps, intrinsics, ii, jj):
    ht, wd = disps.shape[1:]
    X, u, v, (fx, fy, cx, cy) = _rays(intrinsics, ht, wd)
    tij, qij = _rel(poses, ii, jj)
    E = ii.shape[0]
    Y = _rot(qij[:, None, None], X[None].expand(E, ht, wd, 3)) + disps[ii][..., None] * tij[:, None, None]
    front = Y[..., 2] > 0.01
    coords = torch.zeros(E, ht, wd, 3, dtype=poses.dtype, device=poses.device)        # third channel never written (:1449)
    coords[..., 0] = torch.where(front, fx * (Y[..., 0] / Y[..., 2]) + cx, u.expand(E, ht, wd))
    coords[..., 1] = torch.where(front, fy * (Y[..., 1] / Y[..., 2]) + cy, v.expand(E, ht, wd))
    return coords, (Y[..., 2:3] > 0.25).to(poses.dtype)


def _depth_filter(poses, disps, intrinsics, ix, thresh):
    num, ht, wd = disps.shape
    X, _, _, (fx, fy, cx, cy) = _rays(intrinsics, ht, wd)
    n = ix.shape[0]
    counter = torch.zeros(n, ht, wd, dtype=disps.dtype, device=disps.device)
    t = thresh.reshape(n, 1, 1).double()
    for k in range(6):
        jx = ix - k - 1 if k < 3 else ix + k                       # neighbours -1, -2, -3, +3, +4, +5 (:672)
        ok = (jx >= 0) & (jx < num)
        jc = jx.clamp(0, num - 1)
        tij, qij = _rel(poses, ix, jc)
        di = disps[ix]
        Y = _rot(qij[:, None, None], X[None].expand(n, ht, wd, 3)) + di[..., None] * tij[:, None, None]
        uj, vj = fx * (Y[..., 0] / Y[..., 2]) + cx, fy * (Y[..., 1] / Y[..., 2]) + cy
        inv_dj = 1.0 / (di / Y[..., 2]).double()
        u0f, v0f = torch.floor(uj), torch.floor(vj)
        inside = ok[:, None, None] & (u0f >= 0) & (v0f >= 0) & (u0f < wd - 1) & (v0f < ht - 1)      # NaN compares false
        u0 = torch.nan_to_num(u0f, nan=0.0, posinf=0.0, neginf=0.0).clamp(0, wd - 2).long()
        v0 = torch.nan_to_num(v0f, nan=0.0, posinf=0.0, neginf=0.0).clamp(0, ht - 2).long()
        flat = disps[jc].reshape(n, ht * wd)
        hit = torch.zeros(n, ht, wd, dtype=torch.bool, device=disps.device)
        for dv, du in ((0, 0), (0, 1), (1, 0), (1, 1)):
            dn = torch.gather(flat, 1, ((v0 + dv) * wd + u0 + du).reshape(n, -1)).reshape(n, ht, wd).double()
            hit |= (inv_dj - 1.0 / dn).abs() < t
        counter += (inside & hit).to(disps.dtype)
    return counter


def iproj(poses, disps, intrinsics):
    """src/droid_kernels.cu:758-829, :1494-1516: points (n,h,w,3) = (R_k x + d t_k) / d for the n = disps.shape[0] frames"""
    require_cuda(poses, disps, intrinsics)
    return _iproj(poses, disps, intrinsics)


def projmap(poses, disps, intrinsics, ii, jj):
    """src/droid_kernels.cu:406-495, :1439-1465: -> [coords (E,h,w,3), valid (E,h,w,1)]; pixels behind the camera (Z <= 0.01)
    keep their own coordinates; valid = Z > 0.25"""
    require_cuda(poses, disps, intrinsics, ii, jj)
    return list(_projmap(poses, disps, intrinsics, ii, jj))


def depth_filter(poses, disps, intrinsics, ix, thresh):
    """src/droid_kernels.cu:640-754, :1467-1491: for every frame of `ix` and pixel, in how many of the six neighbour frames
    (ix-1, ix-2, ix-3, ix+3, ix+4, ix+5 -- the reference's `ix + neigh_id`) one of the four depth samples around the
    reprojected pixel agrees with the reprojected depth to `thresh` -> counter (len(ix),h,w) float"""
    require_cuda(poses, disps, intrinsics, ix, thresh)
    return _depth_filter(poses, disps, intrinsics, ix, thresh)
