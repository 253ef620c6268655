"""CPU restatement of ``droid_backends.ba`` (the reference's CUDA dense BA) -- oracle only.

Semantics follow VO_Module/src/droid_kernels.cu (NOT geom/ba.py -- the two differ, see
SURVEY.md section 8(a) addendum):
  * :26        MIN_DEPTH 0.25, weight and 1/Z zeroed when Z_j < MIN_DEPTH (:284-291)
  * :58-107    actSO3 / actSE3 / adjSE3 / relSE3 -- quaternions used as stored, never renormalised
  * :110-176   expSO3 / expSE3.  ``xi[45]`` at :154 is an out-of-bounds read (undefined behaviour);
               this oracle uses ``xi[5]`` as upstream DROID-SLAM and lietorch (se3.h:134-142) do.
  * :177-403   projective_transform_kernel: per-edge Hessian blocks Hs[4], vs[2], Eii, Eij, Cii, bz
  * :833-853, :927-977  accum_kernel / accum_cuda: row sums grouped by index equality
  * :980-1094  EEt6x6, Ev6x1, EvT6x1 (EvT skips rows whose pose index is <= 0, :1084)
  * :1171-1192 solve: L = A - S; diag += ep + lm*diag; LLT in float64; failure => dx = 0
  * :1201-1290 schur_block: which E-rows pair up (same depth frame kk)
  * :1293-1410 ba_cuda driver: iterations, motion_only, in-place retraction
  * :856-925   retrSE3 / pose_retr / disp_retr
Float32 arithmetic for the per-pixel kernels, float64 for the reduced camera solve (as the
reference's Eigen SimplicialLLT).  Summation *order* is not reproduced (1e-4 rel. tolerance).
"""
import numpy as np
import torch

MIN_DEPTH = 0.25


def _act_so3(q, X):
    qv, qw = q[..., :3], q[..., 3:4]
    uv = 2.0 * torch.linalg.cross(qv, X, dim=-1)
    return X + qw * uv + torch.linalg.cross(qv, uv, dim=-1)


def _rel_se3(ti, qi, tj, qj):
    """qij = qj * conj(qi); tij = tj - R(qij) ti   (:96-107)."""
    xi, yi, zi, wi = qi.unbind(-1)
    xj, yj, zj, wj = qj.unbind(-1)
    qij = torch.stack([
        -wj * xi + xj * wi - yj * zi + zj * yi,
        -wj * yi + yj * wi - zj * xi + xj * zi,
        -wj * zi + zj * wi - xj * yi + yj * xi,
        wj * wi + xj * xi + yj * yi + zj * zi], dim=-1)
    tij = tj - _act_so3(qij, ti)
    return tij, qij


def _adj_se3(t, q, X):
    """(:80-94) Y = [R^T X_t ; R^T X_r + R^T (X_t x t)]."""
    qinv = torch.cat([-q[..., :3], q[..., 3:4]], dim=-1)
    top = _act_so3(qinv, X[..., :3])
    u = torch.linalg.cross(X[..., :3], t.expand_as(X[..., :3]), dim=-1)
    bot = _act_so3(qinv, X[..., 3:]) + _act_so3(qinv, u)
    return torch.cat([top, bot], dim=-1)


def exp_se3(xi):
    """(:110-176) with xi[5] in place of the out-of-bounds xi[45]."""
    phi = xi[..., 3:]
    theta_sq = (phi * phi).sum(-1, keepdim=True)
    theta = torch.sqrt(theta_sq)
    theta_p4 = theta_sq * theta_sq
    small = theta_sq < 1e-8
    th = torch.where(small, torch.ones_like(theta), theta)
    imag = torch.where(small, 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_p4,
                       torch.sin(0.5 * th) / th)
    real = torch.where(small, 1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * theta_p4,
                       torch.cos(0.5 * th))
    q = torch.cat([imag * phi, real], dim=-1)
    tau = xi[..., :3]
    t = tau.clone()
    big = theta > 1e-4
    th = torch.where(big, theta, torch.ones_like(theta))
    th2 = torch.where(big, theta_sq, torch.ones_like(theta))
    a = (1 - torch.cos(th)) / th2
    b = (th - torch.sin(th)) / (th * th2)
    c1 = torch.linalg.cross(phi, tau, dim=-1)
    c2 = torch.linalg.cross(phi, c1, dim=-1)
    t = torch.where(big, t + a * c1 + b * c2, t)
    return t, q


def retr_se3(xi, pose):
    """(:856-874) Exp(xi) * T."""
    t, q = pose[..., :3], pose[..., 3:]
    dt, dq = exp_se3(xi)
    dx, dy, dz, dw = dq.unbind(-1)
    qx, qy, qz, qw = q.unbind(-1)
    q1 = torch.stack([
        dw * qx + dx * qw + dy * qz - dz * qy,
        dw * qy + dy * qw + dz * qx - dx * qz,
        dw * qz + dz * qw + dx * qy - dy * qx,
        dw * qw - dx * qx - dy * qy - dz * qz], dim=-1)
    t1 = _act_so3(dq, t) + dt
    return torch.cat([t1, q1], dim=-1)


def edge_blocks(poses, disps, intrinsics, targets, weights, ii, jj):
    """projective_transform_kernel (:177-403) for all edges at once.

    targets/weights: (E,2,h,w).  Returns dict of Hs (4,E,6,6), vs (2,E,6), Eii/Eij (E,6,HW),
    Cii/bz (E,HW), plus the residual/Jacobian intermediates for triage.
    """
    E = ii.shape[0]
    ht, wd = disps.shape[1:]
    fx, fy, cx, cy = [intrinsics[k] for k in range(4)]
    ti, qi = poses[ii, :3], poses[ii, 3:]
    tj, qj = poses[jj, :3], poses[jj, 3:]
    tij, qij = _rel_se3(ti, qi, tj, qj)                      # (E,3), (E,4)
    v, u = torch.meshgrid(torch.arange(ht, dtype=torch.float32),
                          torch.arange(wd, dtype=torch.float32), indexing='ij')
    u = u.reshape(-1)[None]
    v = v.reshape(-1)[None]
    di = disps[ii].reshape(E, -1)
    X0 = torch.stack([((u - cx) / fx).expand_as(di), ((v - cy) / fy).expand_as(di),
                      torch.ones_like(di)], dim=-1)          # (E,HW,3)
    Xj3 = _act_so3(qij[:, None], X0) + di[..., None] * tij[:, None]
    x, y, z = Xj3.unbind(-1)
    h = di
    bad = z < MIN_DEPTH
    d = torch.where(bad, torch.zeros_like(z), 1.0 / torch.where(bad, torch.ones_like(z), z))
    d2 = d * d
    tg = targets.reshape(E, 2, -1)
    wt = weights.reshape(E, 2, -1)
    wu = torch.where(bad, torch.zeros_like(z), .001 * wt[:, 0])
    wv = torch.where(bad, torch.zeros_like(z), .001 * wt[:, 1])
    ru = tg[:, 0] - (fx * d * x + cx)
    rv = tg[:, 1] - (fy * d * y + cy)
    o = torch.zeros_like(x)
    Jju = torch.stack([fx * (h * d), o, fx * (-x * h * d2), fx * (-x * y * d2),
                       fx * (1 + x * x * d2), fx * (-y * d)], dim=-1)
    Jjv = torch.stack([o, fy * (h * d), fy * (-y * h * d2), fy * (-1 - y * y * d2),
                       fy * (x * y * d2), fy * (x * d)], dim=-1)
    Jzu = fx * (tij[:, None, 0] * d - tij[:, None, 2] * (x * d2))
    Jzv = fy * (tij[:, None, 1] * d - tij[:, None, 2] * (y * d2))
    Jiu = -_adj_se3(tij[:, None], qij[:, None], Jju)
    Jiv = -_adj_se3(tij[:, None], qij[:, None], Jjv)

    def outer(w, A, B):
        return torch.einsum('ek,eka,ekb->eab', w, A, B)

    Hs = torch.stack([outer(wu, Jiu, Jiu) + outer(wv, Jiv, Jiv),
                      outer(wu, Jiu, Jju) + outer(wv, Jiv, Jjv),
                      outer(wu, Jju, Jiu) + outer(wv, Jjv, Jiv),
                      outer(wu, Jju, Jju) + outer(wv, Jjv, Jjv)], dim=0)
    vs = torch.stack([torch.einsum('ek,eka->ea', wu * ru, Jiu) + torch.einsum('ek,eka->ea', wv * rv, Jiv),
                      torch.einsum('ek,eka->ea', wu * ru, Jju) + torch.einsum('ek,eka->ea', wv * rv, Jjv)], dim=0)
    Eii = ((wu * Jzu)[..., None] * Jiu + (wv * Jzv)[..., None] * Jiv).transpose(1, 2).contiguous()
    Eij = ((wu * Jzu)[..., None] * Jju + (wv * Jzv)[..., None] * Jjv).transpose(1, 2).contiguous()
    Cii = wu * Jzu * Jzu + wv * Jzv * Jzv
    bz = wu * ru * Jzu + wv * rv * Jzv
    return dict(Hs=Hs, vs=vs, Eii=Eii, Eij=Eij, Cii=Cii, bz=bz, ru=ru, rv=rv,
                Jiu=Jiu, Jiv=Jiv, Jju=Jju, Jjv=Jjv, Jzu=Jzu, Jzv=Jzv, wu=wu, wv=wv)


def _accum(data, ix, jx):
    """accum_cuda (:927-977): out[j] = sum of data rows r with ix[r] == jx[j]."""
    out = torch.zeros((jx.shape[0],) + data.shape[1:], dtype=data.dtype)
    for j in range(jx.shape[0]):
        sel = (ix == jx[j]).nonzero().reshape(-1)
        if sel.numel():
            out[j] = data[sel].sum(0)
    return out


def _solve(A, b, lm, ep):
    """SparseBlock::solve (:1171-1192) on a dense float64 copy."""
    L = A.clone()
    dg = torch.diagonal(L)
    dg += ep + lm * dg.clone()
    try:
        U = torch.linalg.cholesky(L)
        return torch.cholesky_solve(b[:, None], U)[:, 0]
    except Exception:
        return torch.zeros_like(b)


def reduced_system(blk, ii, jj, eta, t0, t1, ht, wd, motion_only=False):
    """Assemble A (pose block), and when not motion_only the Schur terms S, plus Q, w, E rows.
    Returns float64 (A, bA, S, bS) and the float32 pieces needed for back-substitution."""
    P = t1 - t0
    E_ = ii.shape[0]
    A = torch.zeros(P, 6, P, 6, dtype=torch.float64)
    bA = torch.zeros(P, 6, dtype=torch.float64)
    pi, pj = (ii - t0).tolist(), (jj - t0).tolist()
    Hs = blk['Hs'].double()
    vs = blk['vs'].double()
    for e in range(E_):
        for (blkid, a, b) in ((0, pi[e], pi[e]), (1, pi[e], pj[e]), (2, pj[e], pi[e]), (3, pj[e], pj[e])):
            if a >= 0 and b >= 0:
                A[a, :, b, :] += Hs[blkid, e]
        if pi[e] >= 0:
            bA[pi[e]] += vs[0, e]
        if pj[e] >= 0:
            bA[pj[e]] += vs[1, e]
    A = A.reshape(6 * P, 6 * P)
    bA = bA.reshape(-1)
    if motion_only:
        return A, bA, None, None, None
    HW = ht * wd
    ts = torch.arange(t0, t1)
    ii_exp = torch.cat([ts, ii])
    jj_exp = torch.cat([ts, jj])
    kx, kk_exp = torch.unique(ii_exp, sorted=True, return_inverse=True)
    C = _accum(blk['Cii'], ii, kx)
    w = _accum(blk['bz'], ii, kx)
    Q = 1.0 / (C + eta.reshape(-1, HW))
    Ei = _accum(blk['Eii'].reshape(E_, 6 * HW), ii, ts).reshape(P, 6, HW)
    Erows = torch.cat([Ei, blk['Eij']], 0)                       # (P+E, 6, HW)
    # schur_block (:1201-1290): rows paired when they share the depth frame kk
    S = torch.zeros(P, 6, P, 6, dtype=torch.float64)
    bS = torch.zeros(P, 6, dtype=torch.float64)
    rows = [(n, int(jj_exp[n]) - t0, int(kk_exp[n])) for n in range(ii_exp.shape[0])]
    rows_in = [(n, p, k) for (n, p, k) in rows if 0 <= p < P]
    for (n, p, k) in rows_in:
        for (m, q, l) in rows_in:
            if k == l:
                S[p, :, q, :] += torch.einsum('ak,k,bk->ab', Erows[n], Q[k], Erows[m]).double()
    for (n, p, k) in rows_in:
        if True:
            bS[p] += torch.einsum('k,ak->a', Q[k] * w[k], Erows[n]).double()
    S = S.reshape(6 * P, 6 * P)
    bS = bS.reshape(-1)
    return A, bA, S, bS, dict(Q=Q, w=w, Erows=Erows, ii_exp=ii_exp, jj_exp=jj_exp, kx=kx, kk_exp=kk_exp)


def ba(poses, disps, intrinsics, targets, weights, eta, ii, jj, t0, t1, iterations, lm, ep,
       motion_only=False, debug=None):
    """In-place on ``poses`` (buf,7) and ``disps`` (buf,h,w), like droid_backends.ba.  Returns (dx, dz)."""
    ht, wd = disps.shape[1:]
    P = t1 - t0
    dx = dz = None
    for itr in range(iterations):
        blk = edge_blocks(poses, disps, intrinsics, targets, weights, ii, jj)
        A, bA, S, bS, aux = reduced_system(blk, ii, jj, eta, t0, t1, ht, wd, motion_only)
        if motion_only:
            dx = _solve(A, bA, lm, ep).float().reshape(P, 6)
            poses[t0:t1] = retr_se3(dx, poses[t0:t1])
            if debug is not None:
                debug.append(dict(A=A, b=bA, dx=dx))
            continue
        dx = _solve(A - S, bA - bS, lm, ep).float().reshape(P, 6)
        ix = aux['jj_exp'] - t0
        dw = torch.zeros(ix.shape[0], ht * wd)
        for n in range(ix.shape[0]):
            p = int(ix[n])
            if p <= 0 or p >= P:                                  # EvT6x1_kernel :1084
                continue
            dw[n] = torch.einsum('ak,a->k', aux['Erows'][n], dx[p])
        dz = aux['Q'] * (aux['w'] - _accum(dw, aux['ii_exp'], aux['kx']))
        poses[t0:t1] = retr_se3(dx, poses[t0:t1])
        disps[aux['kx']] += dz.reshape(-1, ht, wd)
        if debug is not None:
            debug.append(dict(A=A - S, b=bA - bS, dx=dx, dz=dz, kx=aux['kx']))
    return dx, dz


def depth_video_ba(poses, disps, intrinsics0, target, weight, eta, ii, jj, t0=1, t1=None, itrs=2,
                   lm=1e-4, ep=0.1, motion_only=False):
    """DepthVideo.ba (depth_video.py:197-214): default t1/eta, then clamp_(min=0.001)."""
    if t1 is None:
        t1 = int(max(ii.max(), jj.max())) + 1
    if eta is None:
        k = torch.unique(torch.cat([ii, jj])).shape[0]
        eta = 1e-7 * torch.ones(k, disps.shape[1], disps.shape[2])
    out = ba(poses, disps, intrinsics0, target, weight, eta, ii, jj, t0, t1, itrs, lm, ep, motion_only)
    disps.clamp_(min=0.001)
    return out


def frame_distance(poses, disps, intrinsics, ii, jj, beta):
    """droid_backends.frame_distance (droid_kernels.cu:497-636, wrapper :1413-1436): per pair (i, j) the mean induced flow
    magnitude of frame i's pixels, beta * (flow of the full motion) + (1 - beta) * (flow of the translation alone), over
    the pixels that stay in front of camera j (Z > MIN_DEPTH); 1000 when fewer than 75 % do.  Vectorised form of the
    per-pair restatement in tests/test_geom_gpu.py (_frame_distance_ref), which the GPU tests hold against the reference's
    own kernel (oracle/_ref/droid_ba_ref.so)."""
    ii, jj = ii.reshape(-1).long(), jj.reshape(-1).long()
    ht, wd = disps.shape[1:]
    fx, fy, cx, cy = intrinsics.reshape(-1)[:4].unbind(0)
    v, u = torch.meshgrid(torch.arange(ht, dtype=torch.float32), torch.arange(wd, dtype=torch.float32), indexing='ij')
    X0 = torch.stack([(u - cx) / fx, (v - cy) / fy, torch.ones_like(u)], -1)[None]           # (1,h,w,3)
    tij, qij = _rel_se3(poses[ii, :3], poses[ii, 3:], poses[jj, :3], poses[jj, 3:])
    d = disps[ii][..., None]                                                                   # (n,h,w,1)
    acc = torch.zeros(ii.shape[0], dtype=torch.float64)
    val = torch.zeros(ii.shape[0], dtype=torch.float64)
    tot = 0.0
    for full, wgt in ((True, beta), (False, 1.0 - beta)):
        X = (_act_so3(qij[:, None, None], X0) if full else X0) + d * tij[:, None, None]
        du = fx * X[..., 0] / X[..., 2] + cx - u
        dv = fy * X[..., 1] / X[..., 2] + cy - v
        m = X[..., 2] > MIN_DEPTH
        acc += wgt * torch.where(m, torch.sqrt(du * du + dv * dv), torch.zeros_like(du)).double().sum(dim=(1, 2))
        val += wgt * m.double().sum(dim=(1, 2))
        tot += wgt * ht * wd
    out = torch.where(val / (tot + 1e-8) < 0.75, torch.full_like(acc, 1000.0), acc / val.clamp(min=1e-12))
    return out.float()


# ---- the visualiser-side kernels, pixel by pixel (small inputs only) ----------------------------------------------------
def _act_se3_scalar(t, q, X):
    """actSE3 (droid_kernels.cu:71-78) on numpy float32 scalars: X = (x, y, z, d) -> (R (x,y,z) + d t, d)"""
    import numpy as np
    f = np.float32
    uv = [f(2.0) * (q[1] * X[2] - q[2] * X[1]), f(2.0) * (q[2] * X[0] - q[0] * X[2]), f(2.0) * (q[0] * X[1] - q[1] * X[0])]
    Y = [X[0] + q[3] * uv[0] + (q[1] * uv[2] - q[2] * uv[1]), X[1] + q[3] * uv[1] + (q[2] * uv[0] - q[0] * uv[2]),
         X[2] + q[3] * uv[2] + (q[0] * uv[1] - q[1] * uv[0])]
    return [Y[0] + X[3] * t[0], Y[1] + X[3] * t[1], Y[2] + X[3] * t[2], X[3]]


def _rel_scalar(poses, i, j):
    tij, qij = _rel_se3(poses[i, :3], poses[i, 3:], poses[j, :3], poses[j, 3:])
    return tij.numpy(), qij.numpy()


def iproj_loop(poses, disps, intrinsics):
    """iproj_kernel (droid_kernels.cu:758-829)"""
    import numpy as np
    f = np.float32
    fx, fy, cx, cy = (f(x) for x in intrinsics.tolist())
    n, ht, wd = disps.shape
    out = np.zeros((n, ht, wd, 3), dtype=np.float32)
    P, D = poses.numpy(), disps.numpy()
    for k in range(n):
        for i in range(ht):
            for j in range(wd):
                X = [(f(j) - cx) / fx, (f(i) - cy) / fy, f(1), D[k, i, j]]
                Y = _act_se3_scalar(P[k, :3], P[k, 3:], X)
                out[k, i, j] = [Y[0] / Y[3], Y[1] / Y[3], Y[2] / Y[3]]
    return torch.from_numpy(out)


def projmap_loop(poses, disps, intrinsics, ii, jj):
    """projmap_kernel (droid_kernels.cu:406-495)"""
    import numpy as np
    f = np.float32
    fx, fy, cx, cy = (f(x) for x in intrinsics.tolist())
    ht, wd = disps.shape[1:]
    E = ii.shape[0]
    coords = np.zeros((E, ht, wd, 3), dtype=np.float32)
    valid = np.zeros((E, ht, wd, 1), dtype=np.float32)
    D = disps.numpy()
    for e, (a, b) in enumerate(zip(ii.tolist(), jj.tolist())):
        tij, qij = _rel_scalar(poses, a, b)
        for i in range(ht):
            for j in range(wd):
                Y = _act_se3_scalar(tij, qij, [(f(j) - cx) / fx, (f(i) - cy) / fy, f(1), D[a, i, j]])
                coords[e, i, j, :2] = [f(j), f(i)]
                if Y[2] > 0.01:
                    coords[e, i, j, :2] = [fx * (Y[0] / Y[2]) + cx, fy * (Y[1] / Y[2]) + cy]
                valid[e, i, j, 0] = 1.0 if Y[2] > MIN_DEPTH else 0.0
    return torch.from_numpy(coords), torch.from_numpy(valid)


def depth_filter_loop(poses, disps, intrinsics, ix, thresh):
    """depth_filter_kernel (droid_kernels.cu:640-754)"""
    import math
    import numpy as np
    f = np.float32
    fx, fy, cx, cy = (f(x) for x in intrinsics.tolist())
    num, ht, wd = disps.shape
    D = disps.numpy()
    counter = np.zeros((ix.shape[0], ht, wd), dtype=np.float32)
    for b, a in enumerate(ix.tolist()):
        t = float(thresh[b])
        for neigh in range(6):
            jx = a - neigh - 1 if neigh < 3 else a + neigh
            if jx < 0 or jx >= num:
                continue
            tij, qij = _rel_scalar(poses, a, jx)
            for i in range(ht):
                for j in range(wd):
                    Y = _act_se3_scalar(tij, qij, [(f(j) - cx) / fx, (f(i) - cy) / fy, f(1), D[a, i, j]])
                    with np.errstate(all='ignore'):
                        uj, vj, dj = fx * (Y[0] / Y[2]) + cx, fy * (Y[1] / Y[2]) + cy, Y[3] / Y[2]
                    if not (math.isfinite(uj) and math.isfinite(vj)):
                        continue
                    u0, v0 = int(math.floor(uj)), int(math.floor(vj))
                    if u0 >= 0 and v0 >= 0 and u0 < wd - 1 and v0 < ht - 1:
                        with np.errstate(all='ignore'):
                            inv = 1.0 / np.float64(dj)
                            if any(abs(inv - 1.0 / np.float64(D[jx, v0 + dv, u0 + du])) < t for dv, du in ((0, 0), (0, 1), (1, 0), (1, 1))):
                                counter[b, i, j] += 1.0
    return torch.from_numpy(counter)
