"""PoseTrajectoryFiller: poses of the frames that did NOT become keyframes
(VO_Module/droid_slam/trajectory_filler.py:12-104).

Batches of 16 frames: constant-velocity interpolation between the two enclosing keyframes on SE3 (log / exp), features from
the feature encoder, the frames appended temporarily behind the keyframes, two edges each (previous and next keyframe ->
frame), six motion-only iterations (poses of the new frames only; depth and keyframes fixed), poses read back, frames
removed again."""
import torch

from . import lietorch
from .factor_graph import FactorGraph
from .lietorch import SE3
from .motion_filter import RGB_MEAN, RGB_STD

BATCH = 16                  # trajectory_filler.py:93


class PoseTrajectoryFiller:
    def __init__(self, net, video, device='cuda:0', autocast=True):
        self.autocast = autocast
        self.cnet, self.fnet, self.update = net.cnet, net.fnet, net.update
        self.count = 0
        self.video, self.device = video, device
        self.MEAN = torch.as_tensor(RGB_MEAN, device=device)[:, None, None]
        self.STDV = torch.as_tensor(RGB_STD, device=device)[:, None, None]

    def interpolate(self, tstamps):
        """-> (SE3 (M,), t0 (M,), t1 (M,)): constant-velocity poses at `tstamps` between the keyframes before / after each (:42-57)"""
        N = self.video.counter.value
        ts = self.video.tstamp[:N]
        Ps = SE3(self.video.poses[:N])
        tt = torch.as_tensor(tstamps, device=self.device, dtype=ts.dtype)
        t0 = (ts[None, :] <= tt[:, None]).sum(dim=1) - 1            # last keyframe at or before each frame
        t1 = torch.where(t0 < N - 1, t0 + 1, t0)
        dt = ts[t1] - ts[t0] + 1e-3
        v = (Ps[t1] * Ps[t0].inv()).log() / dt.unsqueeze(-1)
        return SE3.exp(v * (tt - ts[t0]).unsqueeze(-1)) * Ps[t0], t0, t1

    def _fill(self, tstamps, images, intrinsics):
        v = self.video
        N, M = v.counter.value, len(tstamps)
        tt = torch.as_tensor(tstamps, device=self.device)
        images = torch.stack(images, 0)
        intrinsics = torch.stack(intrinsics, 0)
        Gs, t0, t1 = self.interpolate(tstamps)
        x = images[None, :, [2, 1, 0]].to(self.device).float() / 255.0
        x = x.sub_(self.MEAN).div_(self.STDV)
        with torch.autocast('cuda', enabled=self.autocast):
            fmap = self.fnet(x).squeeze(0)
        v.counter.value += M
        v[N:N + M] = (tt, images, Gs.data, 1, intrinsics / 8.0, fmap)
        graph = FactorGraph(v, self.update, self.device)
        new = torch.arange(N, N + M, device=self.device)
        graph.add_factors(t0.to(self.device), new)
        graph.add_factors(t1.to(self.device), new)
        for _ in range(6):
            graph.update(N, N + M, motion_only=True)
        Gs = SE3(v.poses[N:N + M].clone())
        v.counter.value -= M
        return [Gs]

    @torch.no_grad()
    def __call__(self, image_stream):
        poses, batch = [], ([], [], [])
        for item in image_stream:
            tstamp, image, intrinsic = item[0], item[1], item[2]
            for lst, x in zip(batch, (tstamp, image, intrinsic)):
                lst.append(x)
            if len(batch[0]) == BATCH:
                poses += self._fill(*batch)
                batch = ([], [], [])
        if batch[0]:
            poses += self._fill(*batch)
        return lietorch.cat(poses, 0)
