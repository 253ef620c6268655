"""DroidBackend: global bundle adjustment over all keyframes (VO_Module/droid_slam/droid_backend.py:9-41).

One call = one global pass: the map is rescaled to unit mean inverse depth, a throw-away FactorGraph over the WHOLE video is
built with the memory-lean correlation (`corr_impl='alt'`: no stored volumes) and proximity edges between all keyframes,
`steps` iterations of FactorGraph.update_lowmem run on it, and the edges are dropped again."""
import torch

from .factor_graph import FactorGraph

EDGE_BUDGET = 100000                      # droid_backend.py:31: effectively unlimited
_OPTIONS = ('beta', 'backend_thresh', 'backend_radius', 'backend_nms')


class DroidBackend:
    def __init__(self, net, video, args):
        self.video, self.update_op, self.device = video, net.update, args.device
        self.t0 = self.t1 = 0             # kept for interface parity; the pass always covers [0, counter)
        for name in _OPTIONS:
            setattr(self, name, getattr(args, name))

    def global_graph(self):
        """all keyframes, proximity edges chosen with the backend's radius / NMS / distance threshold"""
        g = FactorGraph(self.video, self.update_op, self.device, corr_impl='alt', max_factors=EDGE_BUDGET)
        g.add_proximity_factors(rad=self.backend_radius, nms=self.backend_nms, thresh=self.backend_thresh, beta=self.beta)
        return g

    @torch.no_grad()
    def __call__(self, steps=12):
        n = self.video.counter.value
        self.video.normalize()
        g = self.global_graph()
        g.update_lowmem(steps=steps)
        g.clear_edges()
        self.video.dirty[:n] = True
