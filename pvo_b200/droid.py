"""Droid: the visual-odometry system object PVO's scripts drive (VO_Module/droid_slam/droid.py:20-113):
MotionFilter -> DroidFrontend per frame (`track`), two global passes of DroidBackend + PoseTrajectoryFiller at the end
(`terminate`).  `args` is the reference's namespace (evaluation_scripts/test_vo.py:63-83): weights, image_size, buffer, device,
segm_filter, thresh, filter_thresh, warmup, beta, frontend_*, keyframe_thresh, backend_*, disable_vis, use_aff_bri.  The
visualiser process of the reference is not part of this package (`disable_vis` must be True)."""
from collections import OrderedDict

import torch

from . import lietorch
from .depth_video import DepthVideo
from .droid_backend import DroidBackend
from .droid_frontend import DroidFrontend
from .droid_net import DroidNet, upsample_inter
from .lietorch import SE3
from .motion_filter import MotionFilter
from .trajectory_filler import PoseTrajectoryFiller


class Droid:
    def __init__(self, args):
        self.args = args
        self.load_weights(args.weights, getattr(args, 'use_aff_bri', False))
        self.disable_vis = getattr(args, 'disable_vis', True)
        if not self.disable_vis:
            raise NotImplementedError('the Open3D visualiser process is outside this package: run with disable_vis=True')
        self.video = DepthVideo(args.image_size, args.buffer, args.device, args.segm_filter, args.thresh)
        autocast = getattr(args, 'autocast', True)       # False: float32 everywhere (parity runs against the CPU golden)
        if not autocast:
            self.net.update.precision = 'fp32'
        self.filterx = MotionFilter(self.net, self.video, thresh=args.filter_thresh, device=args.device, autocast=autocast)
        self.frontend = DroidFrontend(self.net, self.video, args)
        self.backend = DroidBackend(self.net, self.video, args)
        self.traj_filler = PoseTrajectoryFiller(self.net, self.video, args.device, autocast=autocast)

    def load_weights(self, weights, use_aff_bri=False):
        """weights: checkpoint path or state dict; a DataParallel 'module.' prefix is stripped (droid.py:52-59)"""
        self.net = DroidNet(use_aff_bri)
        sd = torch.load(weights, map_location=self.args.device) if isinstance(weights, str) else weights
        self.net.load_state_dict(OrderedDict((k.replace('module.', ''), v) for k, v in sd.items()))
        self.net.to(self.args.device).eval()

    @torch.no_grad()
    def track(self, tstamp, image, depth=None, intrinsics=None, segments=None):
        self.filterx.track(tstamp, image, depth, intrinsics, segments)
        self.frontend()

    def terminate(self, stream=None, need_inv=True):
        """global BA (7 then 12 steps), then the poses of ALL frames of `stream` as (n,7) [t, q]; need_inv: camera-to-world"""
        del self.frontend
        torch.cuda.empty_cache()
        self.backend(7)
        torch.cuda.empty_cache()
        self.backend(12)
        traj = self.traj_filler(stream)
        return (traj.inv() if need_inv else traj).data.cpu().numpy()

    def get_traj(self):
        return lietorch.cat([SE3(self.video.poses[:self.video.counter.value])], 0).data.cpu().numpy()

    def get_depth(self):
        n = self.video.counter.value
        return upsample_inter(self.video.disps[:n].unsqueeze(0).unsqueeze(4)).squeeze(4).squeeze(0)

    def get_flow(self):
        return upsample_inter(self.video.full_flow[:self.video.counter.value].unsqueeze(0) * 8)
