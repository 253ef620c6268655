"""geom.ba.BA / MoBA with the reference's signatures and semantics (VO_Module/droid_slam/geom/ba.py:31-157,
geom/chol.py:32-73), executed by the same fused on-device kernels as droid_backends.ba, in "Python-BA" mode
(pvo_ba flags bit 3): valid = Z > 0.2, Z < 0.1 -> 1 before the divide, lietorch quaternion normalisation,
C + eta + 1e-7, LM damping applied to H before the Schur complement, every pose contributes to dz, lietorch
retraction, disps > 10 -> 0 and clamp(min=0).  Functional (inputs are not modified), batch size 1, forward only.
"""
import torch

from .. import _lib, droid_backends
from .._lib import check, ptr, stream
from ..lietorch import SE3


def _run(target, weight, eta, poses, disps, intrinsics, ii, jj, fixedp, motion_only):
    data = poses.data if isinstance(poses, SE3) else poses
    if data.shape[0] != 1:
        raise RuntimeError('BA: batch size 1 only')
    for t in (target, weight, data, disps):
        if t.requires_grad:
            raise NotImplementedError('BA: forward only (training is out of scope)')
    dev = data.device
    P7 = data[0].clone().float().contiguous()
    D = disps[0].clone().float().contiguous()
    F_, ht, wd = D.shape
    ii = ii.to(device=dev, dtype=torch.long).contiguous()
    jj = jj.to(device=dev, dtype=torch.long).contiguous()
    t0, t1 = fixedp, F_
    kx_ba = torch.unique(torch.cat([torch.arange(t0, t1, device=dev), ii]))
    K = kx_ba.shape[0]
    eta_full = torch.ones(K, ht, wd, device=dev)
    if not motion_only:
        kx = torch.unique(ii)
        eta_full[torch.searchsorted(kx_ba, kx)] = eta.reshape(-1, ht, wd).float()
    E = ii.shape[0]
    ws = droid_backends.ba_workspace(E, F_, t1 - t0, K, ht * wd, dev)
    tg = target[0].float().contiguous()
    wg = weight[0].float().contiguous()
    flags = 8                                            # Python-BA semantics
    check(_lib.lib().pvo_ba(ptr(P7), ptr(D), ptr(intrinsics[0, 0].float().contiguous()), ptr(tg), ptr(wg), ptr(eta_full),
                            ptr(ii), ptr(jj), E, F_, ht, wd, t0, t1, K, 1, 1e-4, 0.1, 1 if motion_only else 0, 1, None, None,
                            ptr(ws), ws.numel(), flags, stream()), 'ba', launches=5)
    return P7, D


def BA(target, weight, eta, poses, disps, intrinsics, ii, jj, fixedp=1, rig=1):
    """Full bundle adjustment, one Gauss-Newton step (geom/ba.py:31-106).  Returns (SE3 poses, disps)."""
    if rig != 1:
        raise NotImplementedError('BA: rig != 1')
    P7, D = _run(target, weight, eta, poses, disps, intrinsics, ii, jj, fixedp, False)
    D = torch.where(D > 10, torch.zeros_like(D), D).clamp(min=0.0)
    return SE3(P7[None]), D[None]


def MoBA(target, weight, eta, poses, disps, intrinsics, ii, jj, fixedp=1, rig=1):
    """Motion-only bundle adjustment (geom/ba.py:109-157).  Returns SE3 poses."""
    if rig != 1:
        raise NotImplementedError('MoBA: rig != 1')
    P7, _ = _run(target, weight, eta, poses, disps, intrinsics, ii, jj, fixedp, True)
    return SE3(P7[None])
