"""MotionFilter: decides which incoming frames become keyframes and extracts their features
(VO_Module/droid_slam/motion_filter.py:13-109), on the B200 kernels.

Per frame: the feature encoder (always); against the last keyframe one all-pairs volume + window lookup at the identity
flow and ONE update-operator call; the frame is kept when the mean predicted flow magnitude exceeds `thresh` (:85).  Kept
frames also run the context encoder and are appended to the DepthVideo -- pose / depth None = "filled in by the frontend"
(:88-89).  The one device->host read per frame is that decision (as in the reference, `.item()` :85).
"""
import torch

from . import lietorch
from .geom import projective_ops as pops
from .modules.corr import CorrBlock

RGB_MEAN = (0.485, 0.456, 0.406)
RGB_STD = (0.229, 0.224, 0.225)


class MotionFilter:
    def __init__(self, net, video, thresh=2.5, device='cuda:0', autocast=True):
        self.autocast = autocast                          # False: float32 encoders / update operator (parity runs)
        self.cnet, self.fnet, self.update = net.cnet, net.fnet, net.update
        self.video, self.thresh, self.device = video, thresh, device
        self.count = 0                                    # frames skipped since the last keyframe (:91)
        self.MEAN = torch.as_tensor(RGB_MEAN, device=device)[:, None, None]
        self.STDV = torch.as_tensor(RGB_STD, device=device)[:, None, None]
        self.net = self.inp = self.fmap = None            # features of the last keyframe

    # ---------------------------------------------------------------------------------------------
    def _normalized(self, image):
        """(3,H,W) BGR 0..255 -> (1,1,3,H,W) RGB, ImageNet-normalised (:57-58)"""
        x = image[None, None, [2, 1, 0]].to(self.device).float() / 255.0
        return x.sub_(self.MEAN).div_(self.STDV)

    def _context(self, x):
        """context encoder -> hidden state (tanh) and input features (relu), (1,128,h,w) each (:34-39; the reference runs
        cnet twice and discards the first result)"""
        with torch.autocast('cuda', enabled=self.autocast):
            net, inp = self.cnet(x).split([128, 128], dim=2)
        return net.tanh().squeeze(0), inp.relu().squeeze(0)

    def _features(self, x):
        with torch.autocast('cuda', enabled=self.autocast):
            return self.fnet(x).squeeze(0)

    def _keep(self, tstamp, image, x, gmap, pose, disp, intrinsics, segments):
        net, inp = self._context(x)
        self.net, self.inp, self.fmap = net, inp, gmap
        self.video.append(tstamp, image, pose, disp, intrinsics / 8.0, gmap[0], net[0], inp[0], segments)

    def flow_magnitude(self, gmap):
        """mean |flow| predicted by one update iteration from the last keyframe to this frame, at zero initial flow (:74-85)"""
        ht, wd = gmap.shape[-2:]
        coords0 = pops.coords_grid(ht, wd, device=self.device)[None, None]
        with torch.autocast('cuda', enabled=self.autocast):
            corr = CorrBlock(self.fmap[None], gmap[None])(coords0)
            _, delta, _, _ = self.update(self.net[None], self.inp[None], corr, segments=None)
        return delta[..., 0:2].float().norm(dim=-1).mean()

    # ---------------------------------------------------------------------------------------------
    @torch.no_grad()
    def track(self, tstamp, image, depth=None, intrinsics=None, segments=None):
        """run on every frame of the video (:44-93)"""
        x = self._normalized(image)
        gmap = self._features(x)
        if self.video.counter.value == 0:                 # the first frame is always a keyframe, at the identity
            identity = lietorch.SE3.Identity(1, device=self.device).data.squeeze()
            self._keep(tstamp, image, x, gmap, identity, 1.0, intrinsics, segments)
            return True
        if self.flow_magnitude(gmap).item() > self.thresh:
            self.count = 0
            self._keep(tstamp, image, x, gmap, None, None, intrinsics, segments)
            return True
        self.count += 1
        return False

    @torch.no_grad()
    def track_vo(self, tstamp, image, depth=None, intrinsics=None, segments=None):
        """every frame is a keyframe (:95-109)"""
        x = self._normalized(image)
        gmap = self._features(x)
        first = self.video.counter.value == 0
        pose = lietorch.SE3.Identity(1, device=self.device).data.squeeze() if first else None
        self._keep(tstamp, image, x, gmap, pose, 1.0 if first else None, intrinsics, segments)
        return True
