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


_OPTIONS = ('warmup', 'beta', 'frontend_nms', 'keyframe_thresh', 'frontend_window', 'frontend_thresh', 'frontend_radius')


class DroidFrontend:
    max_age = 25            # iterations an edge stays active (droid_frontend.py:24)
    iters1 = 4              # iterations after a new keyframe was linked in (:25)
    iters2 = 2              # ... and more when it is kept (:26)
    max_factors = 48        # active-edge budget of the window graph (:13-14)

    def __init__(self, net, video, args):
        self.video, self.update_op = video, net.update
        self.graph = FactorGraph(video, net.update, args.device, max_factors=self.max_factors)
        self.t0 = self.t1 = 0           # window of keyframes under optimisation: [t0, t1)
        self.is_initialized = False
        self.count = 0
        for name in _OPTIONS:
            setattr(self, name, getattr(args, name))

    # ---------------------------------------------------------------------------------------------
    def _iterate(self, n, t0=None):
        for _ in range(n):
            self.graph.update(t0, None, use_inactive=True)

    def _seed_next(self, disp):
        """the newest keyframe's pose and a constant depth are the initial guess of the keyframe that comes next (:64-65, :89-90)"""
        k = self.t1
        self.video.poses[k] = self.video.poses[k - 1]
        self.video.disps[k] = disp

    def _redundant(self, k):
        """is keyframe k too close to keyframe k-1 (mean induced flow, both directions) to be worth keeping? (:53-56)"""
        d = self.video.distance([k - 1], [k], beta=self.beta, bidirectional=True)
        return d.item() < self.keyframe_thresh

    def _link_newest(self):
        """proximity edges between the newest keyframes and the window (:45-46)"""
        lo = max(self.t1 - self.frontend_window, 0)
        self.graph.add_proximity_factors(self.t1 - 5, lo, rad=self.frontend_radius, nms=self.frontend_nms, thresh=self.frontend_thresh,
                                         beta=self.beta, remove=True)

    # ---------------------------------------------------------------------------------------------
    def _update(self):
        g, v = self.graph, self.video
        self.count += 1
        self.t1 += 1
        if g.corr is not None:
            g.rm_factors(g.age > self.max_age, store=True)
        self._link_newest()
        self._iterate(self.iters1)
        if self._redundant(self.t1 - 2):                  # the second-newest keyframe adds nothing: drop it, shift the newest down
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
        last = self.t1 - 1
        self.last_pose, self.last_disp, self.last_time = v.poses[last].clone(), v.disps[last].clone(), v.tstamp[last].clone()
        self.is_initialized = True
        with v.get_lock():
            v.ready.value = 1
            v.dirty[:self.t1] = True

    @torch.no_grad()
    def __call__(self):
        held = self.video.counter.value
        if self.is_initialized:
            if self.t1 < held:
                self._update()
        elif held == self.warmup:
            self._initialize()
