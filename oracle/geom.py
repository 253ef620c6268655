"""Projective geometry + the *Python* dense bundle adjustment -- oracle, test infrastructure only.

Restates (reference file:line, under VO_Module/droid_slam/):
  * geom/projective_ops.py:6        MIN_DEPTH = 0.2
  * geom/projective_ops.py:13-18    coords_grid
  * geom/projective_ops.py:21-40    iproj
  * geom/projective_ops.py:43-69    proj   (Z < 0.1 -> 1 before the divide)
  * geom/projective_ops.py:72-99    actp   (4x6 Jacobian of the SE3 action)
  * geom/projective_ops.py:102-130  projective_transform
  * geom/ba.py:12-28                safe scatter helpers, pose_retr / disp_retr
  * geom/ba.py:31-106               BA
  * geom/ba.py:109-157              MoBA
  * geom/chol.py:32-73              block_solve / schur_solve (damping on H *before* the Schur complement)
All tensors carry the reference's leading batch dim B.
"""
import torch
from . import se3

MIN_DEPTH = 0.2


def coords_grid(ht, wd, dtype=torch.float32):
    y, x = torch.meshgrid(torch.arange(ht).to(dtype), torch.arange(wd).to(dtype), indexing='ij')
    return torch.stack([x, y], dim=-1)


def _intr(intrinsics):
    return intrinsics[..., None, None, :].unbind(dim=-1)


def iproj(disps, intrinsics):
    ht, wd = disps.shape[2:]
    fx, fy, cx, cy = _intr(intrinsics)
    y, x = torch.meshgrid(torch.arange(ht).to(disps.dtype), torch.arange(wd).to(disps.dtype), indexing='ij')
    i = torch.ones_like(disps)
    X = (x - cx) / fx
    Y = (y - cy) / fy
    return torch.stack([X, Y, i, disps], dim=-1)


def proj(Xs, intrinsics, jacobian=False):
    fx, fy, cx, cy = _intr(intrinsics)
    X, Y, Z, D = Xs.unbind(dim=-1)
    Z = torch.where(Z < 0.5 * MIN_DEPTH, torch.ones_like(Z), Z)
    d = 1.0 / Z
    x = fx * (X * d) + cx
    y = fy * (Y * d) + cy
    coords = torch.stack([x, y], dim=-1)
    if not jacobian:
        return coords, None
    o = torch.zeros_like(d)
    J = torch.stack([fx * d, o, -fx * X * d * d, o,
                     o, fy * d, -fy * Y * d * d, o], dim=-1).reshape(d.shape + (2, 4))
    return coords, J


def actp(Gij, X0, jacobian=False):
    X1 = se3.se3_act4(Gij[:, :, None, None], X0)
    if not jacobian:
        return X1, None
    X, Y, Z, d = X1.unbind(dim=-1)
    o = torch.zeros_like(d)
    Ja = torch.stack([d, o, o, o, Z, -Y,
                      o, d, o, -Z, o, X,
                      o, o, d, Y, -X, o,
                      o, o, o, o, o, o], dim=-1).reshape(d.shape + (4, 6))
    return X1, Ja


def projective_transform(poses, depths, intrinsics, ii, jj, jacobian=False):
    """poses: raw SE3 data (B, F, 7).  Returns coords (B,E,h,w,2), valid (B,E,h,w,1)[, (Ji,Jj,Jz)]."""
    X0 = iproj(depths[:, ii], intrinsics[:, ii])
    Gij = se3.se3_mul(poses[:, jj], se3.se3_inv(poses[:, ii]))
    X1, Ja = actp(Gij, X0, jacobian)
    x1, Jp = proj(X1, intrinsics[:, jj], jacobian)
    valid = ((X1[..., 2] > MIN_DEPTH) & (X0[..., 2] > MIN_DEPTH)).to(depths.dtype).unsqueeze(-1)
    if not jacobian:
        return x1, valid
    Jj = Jp @ Ja
    Ji = -se3.se3_adjT(Gij[:, :, None, None, None], Jj)
    Jz0 = torch.zeros_like(X0)
    Jz0[..., -1] = 1.0
    Jz = se3.se3_act4(Gij[:, :, None, None], Jz0)
    Jz = Jp @ Jz.unsqueeze(-1)
    return x1, valid, (Ji, Jj, Jz)


# ------------------------------------------------------------------ scatter helpers

def scatter_sum(src, index, dim_size):
    """torch_scatter.scatter_sum(src, index, dim=1, dim_size=...)."""
    out = torch.zeros(src.shape[:1] + (dim_size,) + src.shape[2:], dtype=src.dtype, device=src.device)
    if index.numel():
        out.index_add_(1, index, src)
    return out


def scatter_mean(src, index, dim_size=None):
    """torch_scatter.scatter_mean(src, index, dim=1)."""
    if dim_size is None:
        dim_size = int(index.max()) + 1
    s = scatter_sum(src, index, dim_size)
    cnt = torch.zeros(dim_size, dtype=src.dtype, device=src.device).index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
    cnt = cnt.clamp(min=1).reshape((1, dim_size) + (1,) * (src.dim() - 2))
    return s / cnt


def _scatter_mat(A, ii, jj, n, m):
    v = (ii >= 0) & (jj >= 0) & (ii < n) & (jj < m)
    return scatter_sum(A[:, v], ii[v] * m + jj[v], n * m)


def _scatter_vec(b, ii, n):
    v = (ii >= 0) & (ii < n)
    return scatter_sum(b[:, v], ii[v], n)


def _cholesky_solve(H, b):
    """geom/chol.py:5-17: failed factorisation => zero update."""
    try:
        U = torch.linalg.cholesky(H)
        return torch.cholesky_solve(b, U)
    except Exception:
        return torch.zeros_like(b)


def schur_solve(H, E, C, v, w, ep=0.1, lm=0.0001):
    B, P, M, D, HW = E.shape
    H = H.permute(0, 1, 3, 2, 4).reshape(B, P * D, P * D)
    E = E.permute(0, 1, 3, 2, 4).reshape(B, P * D, M * HW)
    Q = (1.0 / C).view(B, M * HW, 1)
    I = torch.eye(P * D, dtype=H.dtype)
    H = H + (ep + lm * H) * I
    v = v.reshape(B, P * D, 1)
    w = w.reshape(B, M * HW, 1)
    Et = E.transpose(1, 2)
    S = H - E @ (Q * Et)
    v = v - E @ (Q * w)
    dx = _cholesky_solve(S, v)
    dz = Q * (w - Et @ dx)
    return dx.reshape(B, P, D), dz.reshape(B, M, HW), S, v


def block_solve(H, b, ep=0.1, lm=0.0001):
    B, N, _, D, _ = H.shape
    I = torch.eye(D, dtype=H.dtype)
    H = H + (ep + lm * H) * I
    H = H.permute(0, 1, 3, 2, 4).reshape(B, N * D, N * D)
    b = b.reshape(B, N * D, 1)
    return _cholesky_solve(H, b).reshape(B, N, D)


def _system(target, weight, poses, disps, intrinsics, ii, jj, fixedp, with_depth):
    B, P, ht, wd = disps.shape
    N = ii.shape[0]
    D = 6
    coords, valid, (Ji, Jj, Jz) = projective_transform(poses, disps, intrinsics, ii, jj, jacobian=True)
    r = (target - coords).view(B, N, -1, 1)
    w = .001 * (valid * weight).view(B, N, -1, 1)
    Ji = Ji.reshape(B, N, -1, D)
    Jj = Jj.reshape(B, N, -1, D)
    wJiT = (w * Ji).transpose(2, 3)
    wJjT = (w * Jj).transpose(2, 3)
    Hii, Hij, Hji, Hjj = wJiT @ Ji, wJiT @ Jj, wJjT @ Ji, wJjT @ Jj
    vi = (wJiT @ r).squeeze(-1)
    vj = (wJjT @ r).squeeze(-1)
    P = P - fixedp
    pi, pj = ii - fixedp, jj - fixedp
    H = _scatter_mat(Hii, pi, pi, P, P) + _scatter_mat(Hij, pi, pj, P, P) + \
        _scatter_mat(Hji, pj, pi, P, P) + _scatter_mat(Hjj, pj, pj, P, P)
    v = _scatter_vec(vi, pi, P) + _scatter_vec(vj, pj, P)
    H = H.view(B, P, P, D, D)
    if not with_depth:
        return H, v, P, None
    Jz = Jz.reshape(B, N, ht * wd, -1)
    Ei = (wJiT.view(B, N, D, ht * wd, -1) * Jz[:, :, None]).sum(dim=-1)
    Ej = (wJjT.view(B, N, D, ht * wd, -1) * Jz[:, :, None]).sum(dim=-1)
    w = w.view(B, N, ht * wd, -1)
    r = r.view(B, N, ht * wd, -1)
    wk = torch.sum(w * r * Jz, dim=-1)
    Ck = torch.sum(w * Jz * Jz, dim=-1)
    kx, kk = torch.unique(ii, return_inverse=True)
    M = kx.shape[0]
    E = _scatter_mat(Ei, pi, kk, P, M) + _scatter_mat(Ej, pj, kk, P, M)
    C = _scatter_vec(Ck, kk, M)
    wv = _scatter_vec(wk, kk, M)
    return H, v, P, (E.view(B, P, M, D, ht * wd), C, wv, kx)


def BA(target, weight, eta, poses, disps, intrinsics, ii, jj, fixedp=1, return_debug=False):
    """One Gauss-Newton step of geom/ba.py:31-106.  poses = raw (B,F,7) data."""
    B, F, ht, wd = disps.shape
    H, v, P, (E, C, w, kx) = _system(target, weight, poses, disps, intrinsics, ii, jj, fixedp, True)
    C = C + eta.view(*C.shape) + 1e-7
    dx, dz, S, vs = schur_solve(H, E, C, v, w)
    upd = scatter_sum(dx, torch.arange(P) + fixedp, F)
    poses = se3.se3_retr(poses, upd)
    disps = disps + scatter_sum(dz.view(B, -1, ht, wd), kx, F)
    disps = torch.where(disps > 10, torch.zeros_like(disps), disps)
    disps = disps.clamp(min=0.0)
    if return_debug:
        return poses, disps, dict(dx=dx, dz=dz, S=S, v=vs)
    return poses, disps


def MoBA(target, weight, eta, poses, disps, intrinsics, ii, jj, fixedp=1):
    """geom/ba.py:109-157."""
    B, F, ht, wd = disps.shape
    H, v, P, _ = _system(target, weight, poses, disps, intrinsics, ii, jj, fixedp, False)
    dx = block_solve(H, v)
    upd = scatter_sum(dx, torch.arange(P) + fixedp, F)
    return se3.se3_retr(poses, upd)
