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


class DepthVideo:
    def __init__(self, image_size=(480, 640), buffer=1024, device='cuda:0', segm_filter=False, thresh=0.8,
                 store_images=True):
        self.counter = Value('i', 0)
        self.ready = Value('i', 0)
        self.ht = ht = image_size[0]
        self.wd = wd = image_size[1]
        self.device = device
        z = lambda *s, dtype=torch.float: torch.zeros(*s, device=device, dtype=dtype)
        self.tstamp = z(buffer)
        self.images = z(buffer, 3, ht, wd, dtype=torch.uint8) if store_images else None
        self.dirty = z(buffer, dtype=torch.bool)
        self.red = z(buffer, dtype=torch.bool)
        self.poses = z(buffer, 7)
        self.disps = torch.ones(buffer, ht // 8, wd // 8, device=device, dtype=torch.float)
        self.disps_up = z(buffer, ht, wd) if store_images else None
        self.intrinsics = z(buffer, 4)
        self.fmaps = z(buffer, 128, ht // 8, wd // 8, dtype=torch.half)
        self.nets = z(buffer, 128, ht // 8, wd // 8, dtype=torch.half)
        self.inps = z(buffer, 128, ht // 8, wd // 8, dtype=torch.half)
        self.poses[:] = torch.as_tensor([0, 0, 0, 0, 0, 0, 1], dtype=torch.float, device=device)
        self.segms = z(buffer, 1, ht // 8, wd // 8, dtype=torch.int)
        self.full_flow = torch.ones(buffer, ht // 8, wd // 8, 2, device=device, dtype=torch.float)
        self.segm_filter = segm_filter
        self.thresh = thresh

    def get_lock(self):
        return self.counter.get_lock()

    def __item_setter(self, index, item):
        if isinstance(index, int) and index >= self.counter.value:
            self.counter.value = index + 1
        elif isinstance(index, torch.Tensor) and index.max().item() > self.counter.value:
            self.counter.value = index.max().item() + 1
        self.tstamp[index] = item[0]
        if self.images is not None:
            self.images[index] = item[1]
        for k, buf in ((2, self.poses), (3, self.disps), (4, self.intrinsics)):
            if item[k] is not None:
                buf[index] = item[k]
        for k, buf in ((5, self.fmaps), (6, self.nets), (7, self.inps)):
            if len(item) > k:
                buf[index] = item[k]
        if self.segm_filter and len(item) > 8:
            self.segms[index] = item[8]

    def __setitem__(self, index, item):
        with self.get_lock():
            self.__item_setter(index, item)

    def __getitem__(self, index):
        with self.get_lock():
            if isinstance(index, int) and index < 0:
                index = self.counter.value + index
            return (self.poses[index], self.disps[index], self.intrinsics[index], self.fmaps[index],
                    self.nets[index], self.inps[index])

    def append(self, *item):
        with self.get_lock():
            self.__item_setter(self.counter.value, item)

    @staticmethod
    def format_indicies(ii, jj, device):
        ii = torch.as_tensor(ii).to(device=device, dtype=torch.long).reshape(-1)
        jj = torch.as_tensor(jj).to(device=device, dtype=torch.long).reshape(-1)
        return ii, jj

    def normalize(self):
        with self.get_lock():
            n = self.counter.value
            s = self.disps[:n].mean()
            self.disps[:n] /= s
            self.poses[:n, :3] *= s
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
           interleaved=False, reuse_plan=False, max_degree=0, max_span=None):
        """dense bundle adjustment (depth_video.py:197-214): in place on poses/disps, then disps.clamp_(min=0.001)."""
        with self.get_lock():
            if t1 is None:
                t1 = max(ii.max().item(), jj.max().item()) + 1
            if eta is None:
                k = torch.unique(torch.cat([ii, jj], 0)).shape[0]
                eta = 1e-7 * torch.ones([k, self.ht // 8, self.wd // 8], device=self.device)
            out = droid_backends.ba(self.poses, self.disps, self.intrinsics[0].contiguous(), target, weight, eta, ii, jj,
                                    t0, t1, itrs, lm, ep, motion_only, interleaved=interleaved, reuse_plan=reuse_plan,
                                    max_degree=max_degree, max_span=max_span)
            check(_lib.lib().pvo_clamp_min(ptr(self.disps), self.disps.numel(), 0.001, stream()), 'clamp_min')
            return out
