"""DroidFrontend: the local sliding-window optimiser (VO_Module/droid_slam/droid_frontend.py:9-112) driving FactorGraph.

State machine, one call per incoming frame:
  * not initialised and exactly `warmup` keyframes collected -> `_initialize` (:72-100): neighbourhood graph (radius 3),
    8 update iterations, proximity edges, 12 more, seed the next keyframe's pose / depth;
  * initialised and a new keyframe arrived -> `_update` (:36-70): retire edges older than `max_age`, add proximity edges for
    the newest keyframes, `iters1` iterations, then either drop the second-newest keyframe (too close to its predecessor)
    or run `iters2` more.
Every iteration is FactorGraph.update = lookup + update operator + post-processing + 2 Gauss-Newton steps on the device.
"""
import torch

from .factor_graph import FactorGraph


class DroidFrontend:
    max_age = 25            # droid_frontend.py:24-26
    iters1 = 4
    iters2 = 2

    def __init__(self, net, video, args):
        self.video = video
        self.update_op = net.update
        self.graph = FactorGraph(video, net.update, args.device, max_factors=48)
        self.t0 = self.t1 = 0           # window of keyframes under optimisation
        self.is_initialized = False
        self.count = 0
        for k in ('warmup', 'beta', 'frontend_nms', 'keyframe_thresh', 'frontend_window', 'frontend_thresh', 'frontend_radius'):
            setattr(self, k, getattr(args, k))

    def _iterate(self, n, t0=None):
        for _ in range(n):
            self.graph.update(t0, None, use_inactive=True)

    def _seed_next(self, disp):
        """pose of the newest keyframe and a constant depth as the initial guess of the next one (:64-65, :89-90)"""
        self.video.poses[self.t1] = self.video.poses[self.t1 - 1]
        self.video.disps[self.t1] = disp

    def _update(self):
        self.count += 1
        self.t1 += 1
        g, v = self.graph, self.video
        if g.corr is not None:
            g.rm_factors(g.age > self.max_age, store=True)
        g.add_proximity_factors(self.t1 - 5, max(self.t1 - self.frontend_window, 0), rad=self.frontend_radius,
                                nms=self.frontend_nms, thresh=self.frontend_thresh, beta=self.beta, remove=True)
        self._iterate(self.iters1)
        d = v.distance([self.t1 - 3], [self.t1 - 2], beta=self.beta, bidirectional=True)
        if d.item() < self.keyframe_thresh:              # the second-newest keyframe adds nothing: drop it
            g.rm_keyframe(self.t1 - 2)
            with v.get_lock():
                v.counter.value -= 1
                self.t1 -= 1
        else:
            self._iterate(self.iters2)
        self._seed_next(v.disps[self.t1 - 1].mean())
        v.dirty[g.ii.min():self.t1] = True

    def _initialize(self):
        g, v = self.graph, self.video
        self.t0, self.t1 = 0, v.counter.value
        g.add_neighborhood_factors(self.t0, self.t1, r=3)
        self._iterate(8, t0=1)
        g.add_proximity_factors(0, 0, rad=2, nms=2, thresh=self.frontend_thresh, remove=False)
        self._iterate(12, t0=1)
        self._seed_next(v.disps[self.t1 - 4:self.t1].mean())
        self.is_initialized = True
        self.last_pose = v.poses[self.t1 - 1].clone()
        self.last_disp = v.disps[self.t1 - 1].clone()
        self.last_time = v.tstamp[self.t1 - 1].clone()
        with v.get_lock():
            v.ready.value = 1
            v.dirty[:self.t1] = True

    @torch.no_grad()
    def __call__(self):
        if not self.is_initialized and self.video.counter.value == self.warmup:
            self._initialize()
        elif self.is_initialized and self.t1 < self.video.counter.value:
            self._update()
