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


def _plain_state_dict(weights, device):
    """checkpoint path or state dict -> keys without the DataParallel 'module.' prefix (droid.py:52-59)"""
    sd = torch.load(weights, map_location=device) if isinstance(weights, str) else weights
    return OrderedDict((k.replace('module.', ''), v) for k, v in sd.items())


class Droid:
    def __init__(self, args):
        self.args = args
        self.disable_vis = getattr(args, 'disable_vis', True)
        if not self.disable_vis:
            raise NotImplementedError('the Open3D visualiser process is outside this package: run with disable_vis=True')
        self.load_weights(args.weights, getattr(args, 'use_aff_bri', False))
        autocast = getattr(args, 'autocast', True)       # False: float32 everywhere (parity runs against the CPU golden)
        if not autocast:
            self.net.update.precision = 'fp32'
        dev = args.device
        self.video = DepthVideo(args.image_size, args.buffer, dev, args.segm_filter, args.thresh)      # the shared state
        self.filterx = MotionFilter(self.net, self.video, thresh=args.filter_thresh, device=dev, autocast=autocast)
        self.frontend = DroidFrontend(self.net, self.video, args)                                     # sliding window
        self.backend = DroidBackend(self.net, self.video, args)                                       # global passes
        self.traj_filler = PoseTrajectoryFiller(self.net, self.video, dev, autocast=autocast)          # non-keyframe poses

    def load_weights(self, weights, use_aff_bri=False):
        self.net = DroidNet(use_aff_bri)
        self.net.load_state_dict(_plain_state_dict(weights, self.args.device))
        self.net.to(self.args.device).eval()

    @torch.no_grad()
    def track(self, tstamp, image, depth=None, intrinsics=None, segments=None):
        """one incoming frame: keyframe decision + features, then one step of the frontend (droid.py:61-72; the reference
        leaves the per-frame backend call commented out)"""
        self.filterx.track(tstamp, image, depth, intrinsics, segments)
        self.frontend()

    def terminate(self, stream=None, need_inv=True):
        """global BA (7 then 12 steps), then the poses of ALL frames of `stream` as (n,7) [t, q] (droid.py:74-95);
        need_inv: return camera-to-world"""
        del self.frontend
        for steps in (7, 12):
            torch.cuda.empty_cache()
            self.backend(steps)
        traj = self.traj_filler(stream)
        if need_inv:
            traj = traj.inv()
        return traj.data.cpu().numpy()

    # ---- read-outs (droid.py:97-107)
    def _keyframes(self):
        return self.video.counter.value

    def get_traj(self):
        return lietorch.cat([SE3(self.video.poses[:self._keyframes()])], 0).data.cpu().numpy()

    def get_depth(self):
        d = self.video.disps[:self._keyframes()]
        return upsample_inter(d[None, ..., None])[0, ..., 0]

    def get_flow(self):
        return upsample_inter(self.video.full_flow[:self._keyframes()][None] * 8)
