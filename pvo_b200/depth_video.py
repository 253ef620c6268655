"""DepthVideo: the shared SLAM state buffers + geometric operations on them
(VO_Module/droid_slam/depth_video.py:13-214), backed by the B200 kernels.

Buffers keep the reference's names, shapes and dtypes (poses (buf,7) float, disps (buf,h/8,w/8) float, intrinsics
(buf,4), fmaps/nets/inps (buf,128,h/8,w/8) half, segms (buf,1,h/8,w/8) int32) so the SLAM orchestration code
(motion_filter.py, droid_frontend.py, droid_backend.py) runs on it unchanged.
"""
import torch
from torch.multiprocessing import Value

from . import _lib, droid_backends
from ._lib import check, ptr, stream


# name -> (shape after the buffer dimension in units of (ht, wd, ht//8, wd//8), dtype, fill); depth_video.py:23-58
def _buffer_table(ht, wd):
    h8, w8 = ht // 8, wd // 8
    return {'tstamp': ((), torch.float, 0), 'dirty': ((), torch.bool, 0), 'red': ((), torch.bool, 0), 'poses': ((7,), torch.float, 0),
            'disps': ((h8, w8), torch.float, 1), 'intrinsics': ((4,), torch.float, 0), 'fmaps': ((128, h8, w8), torch.half, 0),
            'nets': ((128, h8, w8), torch.half, 0), 'inps': ((128, h8, w8), torch.half, 0), 'segms': ((1, h8, w8), torch.int, 0),
            'full_flow': ((h8, w8, 2), torch.float, 1)}


# what a video item carries, by position (depth_video.py:66-93): (tstamp, image, pose, disp, intrinsics, fmap, net, inp, segm)
_OPTIONAL = ((2, 'poses'), (3, 'disps'), (4, 'intrinsics'))        # None = leave as is
_TRAILING = ((5, 'fmaps'), (6, 'nets'), (7, 'inps'))              # present only in longer items


class DepthVideo:
    def __init__(self, image_size=(480, 640), buffer=1024, device='cuda:0', segm_filter=False, thresh=0.8,
                 store_images=True):
        self.counter, self.ready = Value('i', 0), Value('i', 0)      # keyframes held / frontend initialised
        self.ht, self.wd = int(image_size[0]), int(image_size[1])
        self.device = device
        for name, (shape, dtype, fill) in _buffer_table(self.ht, self.wd).items():
            setattr(self, name, torch.full((buffer,) + shape, fill, dtype=dtype, device=device))
        self.poses[:, 6] = 1                                         # identity: t = 0, q = (0,0,0,1)
        # full-resolution buffers: only the visualiser / upsampler read them; bench and tests switch them off
        self.images = torch.zeros(buffer, 3, self.ht, self.wd, dtype=torch.uint8, device=device) if store_images else None
        self.disps_up = torch.zeros(buffer, self.ht, self.wd, device=device) if store_images else None
        self.segm_filter, self.thresh = segm_filter, thresh

    def get_lock(self):
        return self.counter.get_lock()

    def _store(self, index, item):
        n = self.counter.value
        if isinstance(index, int):
            if index >= n:
                self.counter.value = index + 1
        elif isinstance(index, torch.Tensor):
            top = int(index.max().item())
            if top > n:
                self.counter.value = top + 1
        self.tstamp[index] = item[0]
        if self.images is not None:
            self.images[index] = item[1]
        for pos, name in _OPTIONAL:
            if item[pos] is not None:
                getattr(self, name)[index] = item[pos]
        for pos, name in _TRAILING:
            if len(item) > pos:
                getattr(self, name)[index] = item[pos]
        if self.segm_filter and len(item) > 8:
            self.segms[index] = item[8]

    def __setitem__(self, index, item):
        with self.get_lock():
            self._store(index, item)

    def append(self, *item):
        with self.get_lock():
            self._store(self.counter.value, item)

    def __getitem__(self, index):
        """(pose, disp, intrinsics, fmap, net, inp) of a keyframe; negative ints count from the newest"""
        with self.get_lock():
            if isinstance(index, int) and index < 0:
                index += self.counter.value
            return tuple(getattr(self, name)[index] for name in ('poses', 'disps', 'intrinsics', 'fmaps', 'nets', 'inps'))

    @staticmethod
    def format_indicies(ii, jj, device):
        as_long = lambda x: torch.as_tensor(x).to(device=device, dtype=torch.long).reshape(-1)
        return as_long(ii), as_long(jj)

    def upsample(self, ix, mask):
        """convex 8x upsampling of the depth of keyframes ix into disps_up (depth_video.py:139-143)"""
        from .droid_net import cvx_upsample
        self.disps_up[ix] = cvx_upsample(self.disps[ix].unsqueeze(-1), mask).squeeze()

    def normalize(self):
        """rescale the map so that the mean inverse depth is 1 (depth_video.py:145-152)"""
        with self.get_lock():
            n = self.counter.value
            scale = self.disps[:n].mean()
            self.disps[:n] /= scale
            self.poses[:n, :3] *= scale
            self.dirty[:n] = True

    # ---------------------------------------------------------------- geometry on the CUDA kernels
    def reproject(self, ii, jj, motion=None):
        """project points from ii -> jj (depth_video.py:154-163).  Returns coords (1,E,h,w,2), valid (1,E,h,w,1).
        motion = (target, delta_dy, raw_mask, out) additionally writes FactorGraph.update's 8 motion channels."""
        ii, jj = DepthVideo.format_indicies(ii, jj, self.device)
        E = ii.shape[0]
        h, w = self.ht // 8, self.wd // 8
        coords = torch.empty(1, E, h, w, 2, device=self.device, dtype=torch.float)
        valid = torch.empty(1, E, h, w, 1, device=self.device, dtype=torch.float)
        tg = dd = rm = mo = None
        mdt = _lib.F32
        if motion is not None:
            tg, dd, rm, mo = motion
            mdt = _lib.dtype_code(mo)
        check(_lib.lib().pvo_reproject(ptr(self.poses), ptr(self.disps), ptr(self.intrinsics), ptr(ii), ptr(jj),
                                       ptr(coords), ptr(valid), ptr(tg), ptr(dd), ptr(rm), ptr(mo), mdt, E, h, w,
                                       stream()), 'reproject')
        return coords, valid

    def distance(self, ii=None, jj=None, beta=0.3, bidirectional=True):
        """frame distance metric (depth_video.py:165-195)"""
        return_matrix = False
        if ii is None:
            return_matrix = True
            N = self.counter.value
            ii, jj = torch.meshgrid(torch.arange(N), torch.arange(N), indexing='ij')
        ii, jj = DepthVideo.format_indicies(ii, jj, self.device)
        intr = self.intrinsics[0].contiguous()
        if bidirectional:
            poses = self.poses[:self.counter.value].clone()
            d1 = droid_backends.frame_distance(poses, self.disps, intr, ii, jj, beta)
            d2 = droid_backends.frame_distance(poses, self.disps, intr, jj, ii, beta)
            d = .5 * (d1 + d2)
        else:
            d = droid_backends.frame_distance(self.poses, self.disps, intr, ii, jj, beta)
        return d.reshape(N, N) if return_matrix else d

    def ba(self, target, weight, eta, ii, jj, t0=1, t1=None, itrs=2, lm=1e-4, ep=0.1, motion_only=False,
           interleaved=False, reuse_plan=False, max_degree=0, max_span=None, workspace=None, sys_clean=False):
        """dense bundle adjustment (depth_video.py:197-214): in place on poses/disps, then disps.clamp_(min=0.001)."""
        with self.get_lock():
            if t1 is None:
                t1 = max(ii.max().item(), jj.max().item()) + 1
            if eta is None:
                k = torch.unique(torch.cat([ii, jj], 0)).shape[0]
                eta = 1e-7 * torch.ones([k, self.ht // 8, self.wd // 8], device=self.device)
            out = droid_backends.ba(self.poses, self.disps, self.intrinsics[0].contiguous(), target, weight, eta, ii, jj,
                                    t0, t1, itrs, lm, ep, motion_only, interleaved=interleaved, reuse_plan=reuse_plan,
                                    max_degree=max_degree, max_span=max_span, workspace=workspace, sys_clean=sys_clean)
            check(_lib.lib().pvo_clamp_min(ptr(self.disps), self.disps.numel(), 0.001, stream()), 'clamp_min')
            return out
