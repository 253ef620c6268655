"""geom.projective_ops with the reference's call signatures (VO_Module/droid_slam/geom/projective_ops.py), on the
fused reprojection kernel (csrc/reproject.cu)."""
import torch

from .. import _lib
from .._lib import check, ptr, stream
from ..lietorch import SE3

MIN_DEPTH = 0.2


def coords_grid(ht, wd, **kwargs):
    y, x = torch.meshgrid(torch.arange(ht).to(**kwargs).float(), torch.arange(wd).to(**kwargs).float(), indexing='ij')
    return torch.stack([x, y], dim=-1)


def projective_transform(poses, depths, intrinsics, ii, jj, jacobian=False, return_depth=False):
    """map points from ii->jj (projective_ops.py:102-130).  poses: SE3 (B,F,7); depths (B,F,h,w); intrinsics (B,F,4).
    Returns coords (B,E,h,w,2), valid (B,E,h,w,1).  The Jacobian variant exists only inside the fused BA kernels."""
    if jacobian or return_depth:
        raise NotImplementedError('projective_transform(jacobian=True): the Jacobians are formed inside pvo_ba '
                                  '(csrc/ba.cu) and never materialised -- use geom.ba.BA')
    data = poses.data if isinstance(poses, SE3) else poses
    if data.shape[0] != 1:
        raise RuntimeError('projective_transform: batch size 1 only (PVO inference, depth_video.py:157)')
    P, D, K = data[0].contiguous().float(), depths[0].contiguous().float(), intrinsics[0].contiguous().float()
    _lib.require_cuda(P, D, K)
    ii = ii.to(device=P.device, dtype=torch.long).contiguous()
    jj = jj.to(device=P.device, dtype=torch.long).contiguous()
    E, (h, w) = ii.shape[0], D.shape[1:]
    coords = torch.empty(1, E, h, w, 2, device=P.device)
    valid = torch.empty(1, E, h, w, 1, device=P.device)
    check(_lib.lib().pvo_reproject(ptr(P), ptr(D), ptr(K), ptr(ii), ptr(jj), ptr(coords), ptr(valid), None, None, None,
                                   None, 0, E, h, w, stream()), 'reproject')
    return coords, valid


def induced_flow(poses, disps, intrinsics, ii, jj):
    """optical flow induced by camera motion (projective_ops.py:168-181)"""
    ht, wd = disps.shape[2:]
    coords0 = coords_grid(ht, wd, device=disps.device)
    coords1, valid = projective_transform(poses, disps, intrinsics, ii, jj, False)
    return coords1[..., :2] - coords0, valid
