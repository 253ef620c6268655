"""geom.chol with the reference's interface (VO_Module/droid_slam/geom/chol.py:5-73): CholeskySolver (a differentiable SPD solve
that returns a zero update instead of raising when the factorisation fails), block_solve and schur_solve.

These are the TRAINING-path solvers of `geom.ba.BA` when it runs as differentiable tensor algebra; the inference path of this
package never calls them (the fused BA kernels of csrc/ba.cu do the same elimination on the device, pvo_b200/geom/ba.py).  They are
dense linear algebra on a (6P x 6P) system -- cuSOLVER through torch.linalg, as in the reference -- provided so that code written
against the reference's module finds the same names, shapes, damping (H + (ep + lm H) I) and failure behaviour.
"""
import torch


class CholeskySolver(torch.autograd.Function):
    """x = H^-1 b for SPD H.  backward: with g the gradient of x, solve H u = g once; dL/db = u and dL/dH = -x u^T
    (differentiating H x = b; this is the reference's orientation, geom/chol.py:28 -- the transpose of -u x^T, the same on the symmetric
    systems BA builds).  A failed factorisation (not SPD / singular) yields x = 0 and no gradient, so that one bad
    batch element does not stop a training run."""

    @staticmethod
    def forward(ctx, H, b):
        L, info = torch.linalg.cholesky_ex(H)
        ctx.failed = bool((info != 0).any())
        if ctx.failed:
            return torch.zeros_like(b)
        x = torch.cholesky_solve(b, L)
        ctx.save_for_backward(L, x)
        return x

    @staticmethod
    def backward(ctx, grad_x):
        if ctx.failed:
            return None, None
        L, x = ctx.saved_tensors
        u = torch.cholesky_solve(grad_x, L)
        return -x @ u.transpose(-1, -2), u


def _damped(H, ep, lm):
    eye = torch.eye(H.shape[-1], dtype=H.dtype, device=H.device)
    return H + (ep + lm * H) * eye


def block_solve(H, b, ep=0.1, lm=0.0001):
    """normal equations of a pose-only problem: H (B,N,N,D,D) block matrix, b (B,N,D) -> dx (B,N,D)"""
    B, N, _, D, _ = H.shape
    Hd = _damped(H.permute(0, 1, 3, 2, 4).reshape(B, N * D, N * D), ep, lm)
    return CholeskySolver.apply(Hd, b.reshape(B, N * D, 1)).reshape(B, N, D)


def schur_solve(H, E, C, v, w, ep=0.1, lm=0.0001, sless=False):
    """eliminate the (diagonal) depth block C, solve the reduced camera system, back-substitute:
    H (B,P,P,D,D), E (B,P,M,D,HW), C (B,M,HW), v (B,P,D), w (B,M,HW) -> dx (B,P,D) [, dz (B,M,HW)]"""
    B, P, M, D, HW = E.shape
    Hd = _damped(H.permute(0, 1, 3, 2, 4).reshape(B, P * D, P * D), ep, lm)
    Em = E.permute(0, 1, 3, 2, 4).reshape(B, P * D, M * HW)
    Q = (1.0 / C).view(B, M * HW, 1)
    Et = Em.transpose(1, 2)
    wv = w.reshape(B, M * HW, 1)
    S = Hd - Em @ (Q * Et)
    r = v.reshape(B, P * D, 1) - Em @ (Q * wv)
    dx = CholeskySolver.apply(S, r)
    if sless:
        return dx.reshape(B, P, D)
    dz = Q * (wv - Et @ dx)
    return dx.reshape(B, P, D), dz.reshape(B, M, HW)
