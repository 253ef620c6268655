"""FactorGraph: edge bookkeeping + the hot loop body ``update()`` -- one "update-op iteration"
(VO_Module/droid_slam/factor_graph.py:12-307), on the fused B200 path.

update() does, with no host synchronisation and ~30 kernel launches (CUDA-graph capturable):
    reproject + 8 motion channels          1 launch   (reference: ~30 torch kernels, factor_graph.py:231-237)
    4-level correlation lookup             1 launch   (reference: 4 x (zeros + kernel) + cat, :240)
    update operator                        ~14 launches (reference: ~20 cuDNN + ~12 elementwise, :242)
    target / mask / segment filter / weights  1-3 launches (reference: host numpy round trip, :250-283)
    dense BA, 2 Gauss-Newton iterations    9 launches (reference: >= 12 host round trips, :302)
State that the reference keeps as (1,E,C,h,w) tensors (net, inp) lives pixel-major inside the update engine between
iterations; the reference-layout views are materialised only when the edge set changes (add/rm factors).
"""
import itertools

import numpy as np
import torch

from . import _lib, droid_backends
from . import nn as pnn
from ._lib import check, ptr, stream
from .droid_net import DynamicUpdateModule
from .modules.corr import CorrBlock, AltCorrBlock


_SERIAL = itertools.count()         # one number per prepared edge set, process wide


def coords_grid(ht, wd, device):
    """geom/projective_ops.py:13-18"""
    y, x = torch.meshgrid(torch.arange(ht, device=device).float(), torch.arange(wd, device=device).float(), indexing='ij')
    return torch.stack([x, y], dim=-1)


class FactorGraph:
    def __init__(self, video, update_op, device='cuda:0', corr_impl='volume', max_factors=-1, precision='auto',
                 use_cuda_graph=False):
        self.video = video
        self.update_op = update_op
        self.device = device
        self.max_factors = max_factors
        self.corr_impl = corr_impl
        self.precision = precision
        self.use_cuda_graph = use_cuda_graph   # replay the whole iteration as one CUDA graph (no per-launch CPU cost)
        self.use_launch_plan = True            # eager path: record the launches of an edge set once, re-issue them afterwards
        self.ht = ht = video.ht // 8
        self.wd = wd = video.wd // 8
        self.dy_thresh = 0.5
        self.mask_num = 2
        self.coords0 = coords_grid(ht, wd, device)
        lz = lambda: torch.as_tensor([], dtype=torch.long, device=device)
        fz = lambda c=2: torch.zeros([1, 0, ht, wd, c], device=device, dtype=torch.float)
        self.ii, self.jj, self.age = lz(), lz(), lz()
        self.corr = None
        self._net = self._inp = None          # reference-layout caches (1,E,128,h,w)
        self._state_in_engine = False         # True: engine.X[:, :256] holds the authoritative net|inp
        self.segm = None
        self.damping = 1e-6 * torch.ones_like(self.video.disps)
        self.target_cam, self.weight, self.delta_dy, self.full_flow = fz(), fz(), fz(), fz()
        self.raw_mask = fz(self.mask_num)
        self.ii_inac, self.jj_inac, self.ii_bad, self.jj_bad = lz(), lz(), lz(), lz()
        self.target_cam_inac, self.weight_inac, self.delta_dy_inac = fz(), fz(), fz()
        self.raw_mask_inac = fz(self.mask_num)
        self._cache = None                    # per-edge-set constants (unique(ii), K, default t0, workspaces)
        self.timers = None                    # dict -> per-stage CUDA events (bench.py)
        self.launch_count = None              # kernels launched by the last update()

    # ------------------------------------------------------------------ reference-layout state (lazy)
    def _engine(self):
        if not isinstance(self.update_op, DynamicUpdateModule):
            raise RuntimeError('FactorGraph: update_op must be a pvo_b200.droid_net.DynamicUpdateModule')
        prec = self.precision if self.precision != 'auto' else self.update_op.precision
        mode = prec if prec in ('fp32', 'fp16') else 'fp16'       # the reference's update() is autocast (:227)
        return self.update_op.engine(mode, self.device, owner=self)

    def _pull_state(self):
        """engine (pixel-major) -> reference layout; called before the edge set is edited."""
        if self._state_in_engine:
            eng = self._engine()
            E = self.ii.shape[0]
            self._net = eng.store_nchw(eng.X, 448, 0, 128, E).view(1, E, 128, self.ht, self.wd)
            self._state_in_engine = False

    @property
    def net(self):
        self._pull_state()
        return self._net

    @net.setter
    def net(self, v):
        self._net = v
        self._state_in_engine = False
        self._cache = None

    @property
    def inp(self):
        return self._inp

    @inp.setter
    def inp(self, v):
        self._inp = v
        self._state_in_engine = False
        self._cache = None

    # ------------------------------------------------------------------ edge bookkeeping (factor_graph.py:60-225)
    def __filter_repeated_edges(self, ii, jj):
        """drop the candidate edges that are already active or inactive (factor_graph.py:60-70).  The reference walks a
        Python set with one `.item()` per element; here one `isin` over packed (i, j) keys -- no host synchronisation."""
        big = self.video.poses.shape[0] + 1
        have = torch.cat([self.ii * big + self.jj, self.ii_inac * big + self.jj_inac])
        keep = ~torch.isin(ii * big + jj, have)
        return ii[keep], jj[keep]

    def clear_edges(self):
        self.rm_factors(self.ii >= 0)
        self.net = None
        self.inp = None

    def add_factors(self, ii, jj, remove=False):
        """add edges (factor_graph.py:106-161): volume pyramid for the new edges, hidden state from the video."""
        ii = torch.as_tensor(ii, dtype=torch.long, device=self.device)
        jj = torch.as_tensor(jj, dtype=torch.long, device=self.device)
        ii, jj = self.__filter_repeated_edges(ii, jj)
        if ii.shape[0] == 0:
            return
        if self.max_factors > 0 and self.ii.shape[0] + ii.shape[0] > self.max_factors and self.corr is not None and remove:
            ix = torch.arange(len(self.age))[torch.argsort(self.age).cpu()]
            self.rm_factors(ix >= self.max_factors - ii.shape[0], store=True)
        self._pull_state()
        net = self.video.nets[ii].to(self.device).unsqueeze(0)
        if self.corr_impl == 'volume':
            f1, f2 = self.video.fmaps[ii].to(self.device).unsqueeze(0), self.video.fmaps[jj].to(self.device).unsqueeze(0)
            pooled = hasattr(CorrBlock, 'append')     # (a reference-style CorrBlock substituted for ours has only cat / __getitem__)
            if self.corr is None:     # volumes live in a slot pool: later edges are built into free slots, retired ones free theirs
                self.corr = CorrBlock(f1, f2, capacity=2 * self.max_factors if self.max_factors > 0 else None) if pooled else CorrBlock(f1, f2)
            elif pooled:
                self.corr.append(f1, f2)
            else:
                self.corr = self.corr.cat(CorrBlock(f1, f2))
            inp = self.video.inps[ii].to(self.device).unsqueeze(0)
            self._inp = inp if self._inp is None else torch.cat([self._inp, inp], 1)
        target, _ = self.video.reproject(ii, jj)
        self.ii = torch.cat([self.ii, ii], 0)
        self.jj = torch.cat([self.jj, jj], 0)
        self.age = torch.cat([self.age, torch.zeros_like(ii)], 0)
        # torch.cat promotes like the reference's cat of a float state with half features (factor_graph.py:147): the
        # hidden state of the existing edges is never narrowed
        self._net = net if self._net is None else torch.cat([self._net, net.to(torch.promote_types(self._net.dtype, net.dtype))], 1)
        self.target_cam = torch.cat([self.target_cam, target], 1)
        self.weight = torch.cat([self.weight, torch.zeros_like(target)], 1)
        self.raw_mask = torch.cat([self.raw_mask, torch.zeros_like(target)[..., 0:self.mask_num]], 1)
        self.delta_dy = torch.cat([self.delta_dy, torch.zeros_like(target)], 1)
        segm = self.video.segms[ii].to(self.device).unsqueeze(0)
        self.segm = segm if self.segm is None else torch.cat([self.segm, segm], 1)
        self._cache = None

    def rm_factors(self, mask, store=False):
        """drop edges (factor_graph.py:163-199)"""
        mask = mask.to(self.device)
        self._pull_state()
        if store:
            self.ii_inac = torch.cat([self.ii_inac, self.ii[mask]], 0)
            self.jj_inac = torch.cat([self.jj_inac, self.jj[mask]], 0)
            self.target_cam_inac = torch.cat([self.target_cam_inac, self.target_cam[:, mask]], 1)
            self.weight_inac = torch.cat([self.weight_inac, self.weight[:, mask]], 1)
            self.raw_mask_inac = torch.cat([self.raw_mask_inac, self.raw_mask[:, mask]], 1)
            self.delta_dy_inac = torch.cat([self.delta_dy_inac, self.delta_dy[:, mask]], 1)
        keep = ~mask
        self.ii, self.jj, self.age = self.ii[keep], self.jj[keep], self.age[keep]
        if self.corr_impl == 'volume' and self.corr is not None:
            self.corr = self.corr[keep]
        if self._net is not None:
            self._net = self._net[:, keep]
        if self._inp is not None:
            self._inp = self._inp[:, keep]
        if self.segm is not None:
            self.segm = self.segm[:, keep]
        self.target_cam = self.target_cam[:, keep]
        self.weight = self.weight[:, keep]
        self.raw_mask = self.raw_mask[:, keep]
        self.delta_dy = self.delta_dy[:, keep]
        self._cache = None

    def rm_keyframe(self, ix):
        """drop a keyframe and its edges (factor_graph.py:201-225)"""
        with self.video.get_lock():
            for buf in (self.video.poses, self.video.disps, self.video.intrinsics, self.video.nets, self.video.inps,
                        self.video.fmaps):
                buf[ix] = buf[ix + 1]
            if self.video.segm_filter:
                self.video.segms[ix] = self.video.segms[ix + 1]
        m = (self.ii == ix) | (self.jj == ix)
        self.ii[self.ii >= ix] -= 1
        self.jj[self.jj >= ix] -= 1
        self.ii_inac[self.ii_inac >= ix] -= 1
        self.jj_inac[self.jj_inac >= ix] -= 1
        self.rm_factors(m, store=False)

    def add_neighborhood_factors(self, t0, t1, r=3):
        """edges between all frames within distance r (factor_graph.py:362-370)"""
        ii, jj = torch.meshgrid(torch.arange(t0, t1), torch.arange(t0, t1), indexing='ij')
        ii, jj = ii.reshape(-1).to(self.device), jj.reshape(-1).to(self.device)
        keep = ((ii - jj).abs() > 0) & ((ii - jj).abs() <= r)
        self.add_factors(ii[keep], jj[keep])

    def add_proximity_factors(self, t0=0, t1=0, rad=2, nms=2, beta=0.25, thresh=16.0, remove=False):
        """add edges based on frame distance (factor_graph.py:372-429): `frame_distance` on the device, one copy of the
        distance matrix to the host, greedy selection + NMS there (graph_select.py)"""
        from .graph_select import select_proximity_edges
        t = self.video.counter.value
        ii, jj = torch.meshgrid(torch.arange(t0, t), torch.arange(t1, t), indexing='ij')
        d = self.video.distance(ii.reshape(-1), jj.reshape(-1), beta=beta).float().cpu().numpy()
        existing = torch.stack([torch.cat([self.ii, self.ii_bad, self.ii_inac], 0),
                                torch.cat([self.jj, self.jj_bad, self.jj_inac], 0)], 1).cpu().tolist()
        es = select_proximity_edges(d, existing, t0, t1, t, rad=rad, nms=nms, thresh=thresh)
        if not es:
            return
        e = torch.as_tensor(es, dtype=torch.long, device=self.device)
        self.add_factors(e[:, 0], e[:, 1], remove)

    # ------------------------------------------------------------------ the hot loop body
    def _prepare(self):
        """Per-edge-set constants.  The only host synchronisation of the hot path, paid when edges change."""
        eng = self._engine()
        E = self.ii.shape[0]
        kx = torch.unique(self.ii)
        K = int(kx.shape[0])
        nframes = self.video.disps.shape[0]
        eng.reserve(E, K, self.ht, self.wd, nframes)
        if not self._state_in_engine:
            eng.load_nchw(self._net.reshape(E, 128, self.ht, self.wd).contiguous(), eng.X, 448, 0)
            eng.load_nchw(self._inp.reshape(E, 128, self.ht, self.wd).contiguous(), eng.X, 448, 128)
            self._state_in_engine = True
        eng.group(self.ii)
        L = _lib.lib()
        c = dict(E=E, K=K, kx=kx, t0_default=max(1, int(self.ii.min().item()) + 1),
                 t1_default=int(max(self.ii.max().item(), self.jj.max().item())) + 1,
                 serial=next(_SERIAL), ba_kx={}, graphs={}, plans={},
                 max_degree=int(torch.bincount(self.ii).max().item()),
                 max_span=int((self.ii - self.jj).abs().max().item()))
        if getattr(self, '_span_reduce', None) is not None:    # sharded graph: the bandwidth bound must hold for the summed system.
            c['max_span'] = self._span_reduce(c['max_span'])   # One collective per edge-set change, issued by every rank (dist.attach)
        c['coords1'] = torch.empty(1, E, self.ht, self.wd, 2, device=self.device)
        c['seg_ws'] = torch.zeros(int(L.pvo_update_post_workspace_ints(E)), dtype=torch.int32, device=self.device)
        c['segm'] = self.segm[0, :, 0].contiguous() if self.segm is not None else None
        for name in ('target_cam', 'weight', 'raw_mask', 'delta_dy'):
            setattr(self, name, getattr(self, name).contiguous())
        self.full_flow = torch.empty_like(self.target_cam)
        self._cache = c
        return c

    @torch.no_grad()
    def update(self, t0=None, t1=None, itrs=2, use_inactive=False, EP=1e-7, motion_only=False):
        """run update operator on factor graph (factor_graph.py:227-307)"""
        if not (self.use_cuda_graph and self.timers is None):
            return self._update(t0, t1, itrs, use_inactive, EP, motion_only)
        # CUDA-graph path: 1st call with a given signature runs eagerly (lazy initialisation, attribute setup; with
        # use_inactive it also builds the inactive slab, the only step with a host synchronisation), 2nd call captures, later
        # calls replay.  Every launch of the iteration is capture-safe.  The BA of such a graph runs on a private workspace
        # held by this cache entry (see _iteration), so other BA calls on the device between replays (another FactorGraph, the
        # backend, a different window) can neither leave a foreign schedule behind nor free the buffer the graph writes to.
        c = self._cache or self._prepare()
        key = (t0, t1, itrs, EP, motion_only, use_inactive)
        entry = c['graphs'].get(key)
        if entry is None:
            c['graphs'][key] = 'warm'
            return self._update(t0, t1, itrs, use_inactive, EP, motion_only)
        if entry == 'warm':
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            # with a sharded graph the NCCL all-reduce of the BA system is captured too; thread_local keeps the
            # process group's watchdog thread from tripping the capture
            mode = 'thread_local' if getattr(self, 'world', 1) > 1 else 'global'
            with torch.cuda.graph(g, capture_error_mode=mode):
                self._update(t0, t1, itrs, use_inactive, EP, motion_only, capturing=True)
            c['graphs'][key] = g
            entry = g
        entry.replay()

    def _inactive_slab(self, c, t0, t1):
        """`use_inactive=True` (every frontend call, droid_frontend.py:49,62,81,87): the BA also sees the stored edges whose
        frames are still near the window (factor_graph.py:284-291).  The reference concatenates them with the active edges on
        every iteration (4 torch.cat + torch.unique); here the selection is made ONCE per (edge set, window): one buffer holds
        [selected inactive | active] targets and weights, and `self.target_cam` / `self.weight` BECOME views of its tail, so the
        update operator's post-processing writes the active part in place and the BA reads the whole buffer -- no copies, no
        host synchronisation per iteration.  The only syncs (selected-edge count, window end, degree / span bounds) are here."""
        m = (self.ii_inac >= t0 - 3) & (self.jj_inac >= t0 - 3)
        sel = torch.nonzero(m).flatten()
        n, E = int(sel.shape[0]), c['E']
        ii = torch.cat([self.ii_inac[sel], self.ii], 0)
        jj = torch.cat([self.jj_inac[sel], self.jj], 0)
        tw = torch.empty(2, 1, n + E, self.ht, self.wd, 2, device=self.device)
        if n:
            tw[0, :, :n] = self.target_cam_inac[:, sel]
            tw[1, :, :n] = self.weight_inac[:, sel]
        tw[0, :, n:] = self.target_cam
        tw[1, :, n:] = self.weight
        self.target_cam, self.weight = tw[0, :, n:], tw[1, :, n:]
        t1_ = int(max(ii.max().item(), jj.max().item())) + 1 if t1 is None else t1        # depth_video.py:203-204
        kx = torch.unique(torch.cat([torch.arange(t0, t1_, device=self.device), ii]))
        # the reference indexes the damping by unique(ii) and its kernel by unique(arange(t0,t1) U ii) (droid_kernels.cu:1311-1319);
        # when the two differ its `C + eta` raises a shape error -- so does this
        if int(torch.unique(ii).shape[0]) != int(kx.shape[0]):
            raise RuntimeError('FactorGraph.update: a frame of the BA window [%d, %d) has no out-edge (damping rows would misalign)'
                               % (t0, t1_))
        max_span = int((ii - jj).abs().max().item())
        if getattr(self, '_span_reduce', None) is not None:
            max_span = self._span_reduce(max_span)
        ia = dict(key=(t0, t1), t1=t1_, n=n, ii=ii, jj=jj, tw=tw, kx=kx, serial=next(_SERIAL),
                  max_degree=int(torch.bincount(ii).max().item()), max_span=max_span)
        c['inac'] = ia
        c['graphs'] = {k: g for k, g in c['graphs'].items() if not k[5]}      # graphs captured on the previous slab are stale
        c['plans'] = {k: g for k, g in c['plans'].items() if not k[5]}
        return ia

    def _update(self, t0=None, t1=None, itrs=2, use_inactive=False, EP=1e-7, motion_only=False, capturing=False):
        c = self._cache or self._prepare()
        eng = self._engine()
        L = _lib.lib()
        E, K, ht, wd = c['E'], c['K'], self.ht, self.wd
        v = self.video
        coords1 = c['coords1']
        T = lambda name: _lib.timed(self.timers, name)
        n0 = _lib.LAUNCHES[0]
        if t0 is None:
            t0 = c['t0_default']
        ia = None
        if use_inactive:
            ia = c.get('inac')
            if ia is None or ia['key'] != (t0, t1):
                if capturing:
                    raise RuntimeError('FactorGraph: the inactive slab must exist before capture')
                ia = self._inactive_slab(c, t0, t1)
        # Launch plan (_lib.LaunchPlan): the first iteration on an edge set records its launches, the following ones re-issue them
        # without any Python argument marshalling.  Not with per-stage timers, the mixed-precision lookup (torch ops in between),
        # a sharded graph (NCCL call in between) or while a CUDA graph is being captured.
        key = (t0, t1, itrs, EP, motion_only, use_inactive)
        plannable = (self.timers is None and not capturing and self.use_launch_plan and self.corr.dtype == eng.tdtype
                     and getattr(self, 'world', 1) == 1)
        plan = c['plans'].get(key) if plannable else None
        if plan is not None:
            plan.replay()
            droid_backends.set_ba_plan_owner(None)            # the replay rewrote the shared schedule
            self._state_in_engine = True
            self.launch_count = plan.launches
            return
        with _lib.recording(plannable) as rec:
            self._iteration(c, eng, ia, t0, t1, itrs, EP, motion_only, use_inactive, capturing or plannable, capturing)
        if rec is not None and rec.calls is not None:
            c['plans'][key] = rec
        self.launch_count = _lib.LAUNCHES[0] - n0

    def _iteration(self, c, eng, ia, t0, t1, itrs, EP, motion_only, use_inactive, fresh_schedule, capturing=False):
        L = _lib.lib()
        E, ht, wd = c['E'], self.ht, self.wd
        v = self.video
        coords1 = c['coords1']
        T = lambda name: _lib.timed(self.timers, name)
        # (1) reprojection + motion features, written straight into the flow encoder's input buffer
        with T('reproject'):
            check(L.pvo_reproject(ptr(v.poses), ptr(v.disps), ptr(v.intrinsics), ptr(self.ii), ptr(self.jj), ptr(coords1),
                                  None, ptr(self.target_cam), ptr(self.delta_dy), ptr(self.raw_mask), ptr(eng.motn),
                                  eng.dcode, E, ht, wd, stream()), 'reproject')
        # (2) correlation features, pixel-major into the corr encoder's input buffer
        with T('corr_lookup'):
            if self.corr.dtype == eng.tdtype:
                self.corr(coords1, out=eng.corr, layout=1, out_cstride=pnn.CORR_STRIDE)
            else:
                self._lookup_mixed(coords1, eng)
        # (3) update operator (hidden state advances in place inside the engine)
        eng.run(with_agg=True)
        self._state_in_engine = True
        # (4) target, masks, segment filter, weights (in place on the graph's state tensors)
        segm = c['segm'] if v.segm_filter else None
        with T('update_post'):
            check(L.pvo_update_post(ptr(coords1), ptr(eng.head_out), ptr(segm), float(v.thresh), ptr(self.target_cam),
                                    ptr(self.weight), ptr(self.raw_mask), ptr(self.delta_dy), ptr(self.full_flow),
                                    ptr(c['seg_ws']), c['seg_ws'].numel(), E, ht, wd, stream()), 'update_post',
                  launches=3 if segm is not None else 1)
        kx, eta, dampall = c['kx'], eng.eta, self.damping
        _lib.torch_op(lambda: dampall.index_copy_(0, kx, eta))
        if use_inactive:
            ii, jj, target_cam, weight, t1, kx_ba = ia['ii'], ia['jj'], ia['tw'][0], ia['tw'][1], ia['t1'], ia['kx']
            max_degree, max_span, serial = ia['max_degree'], ia['max_span'], ia['serial']
        else:
            ii, jj, target_cam, weight = self.ii, self.jj, self.target_cam, self.weight
            if t1 is None:
                t1 = c['t1_default']
            # the BA's depth frames are unique(arange(t0,t1) U ii) (droid_kernels.cu:1311-1319); normally == unique(ii)
            kx_ba = c['ba_kx'].get((t0, t1))
            if kx_ba is None:
                kx_ba = torch.unique(torch.cat([torch.arange(t0, t1, device=self.device), ii]))
                c['ba_kx'][(t0, t1)] = kx_ba
            max_degree, max_span, serial = c['max_degree'], c['max_span'], c['serial']
        damping = c.get(('damp', kx_ba.shape[0]))
        if damping is None:
            damping = c[('damp', kx_ba.shape[0])] = torch.empty(kx_ba.shape[0], ht, wd, device=self.device)

        def damp():
            torch.index_select(dampall, 0, kx_ba, out=damping)
            damping.mul_(.2).add_(EP)
        _lib.torch_op(damp)
        # the BA schedule in the (shared) workspace can be reused iff the last BA call anywhere was this graph's, on this edge set and
        # window; a recorded iteration always rebuilds it.  A graph that replays as a CUDA graph owns a PRIVATE workspace per
        # (edges, window): nobody else can leave a foreign schedule there or dirty its system replicas, so the captured iteration
        # carries neither the schedule kernel nor the clearing of the replicas (43 us per step at P = 63), and growth of the shared
        # workspace elsewhere cannot free the buffer under the graph.
        token = (id(self), serial, t0, t1, ii.shape[0], motion_only)
        ws, clean = None, False
        if self.use_cuda_graph:
            wkey = ('ba_ws', ii.shape[0], t0, t1, kx_ba.shape[0])
            ws = c.get(wkey)
            if ws is None:
                if capturing:
                    raise RuntimeError('FactorGraph: the private BA workspace must exist before capture')
                need = _lib.lib().pvo_ba_workspace_bytes(ii.shape[0], v.disps.shape[0], t1 - t0, kx_ba.shape[0], ht * wd)
                ws = c[wkey] = torch.zeros(int(need) + 1024, dtype=torch.uint8, device=self.device)
            reuse = c.get(wkey + ('token',)) == token
            clean = reuse
        else:
            reuse = (not fresh_schedule) and droid_backends.ba_plan_owner() == token
        # (5) dense bundle adjustment, in place on video.poses / video.disps
        with T('ba_total'):
            v.ba(target_cam[0], weight[0], damping, ii, jj, t0, t1, itrs=itrs, lm=1e-4, ep=0.1, motion_only=motion_only,
                 interleaved=True, reuse_plan=reuse, max_degree=max_degree, max_span=max_span, workspace=ws, sys_clean=clean)
        if ws is not None:
            c[wkey + ('token',)] = token
        else:
            droid_backends.set_ba_plan_owner(token)
        age = self.age
        _lib.torch_op(lambda: age.add_(1))

    @torch.no_grad()
    def update_lowmem(self, t0=None, t1=None, itrs=2, use_inactive=False, EP=1e-7, steps=8):
        """run the update operator on the factor graph, reduced-memory form (factor_graph.py:309-360; the global-BA path of
        droid_backend.py): no materialised volumes -- AltCorrBlock computes the correlation features of each chunk of
        edges (grouped by source frame, 8 frames per chunk) on the fly -- and one dense BA over the whole window [1, t)
        with lm=1e-5, ep=1e-2 per step.  Same state updates as the reference, chunk by chunk."""
        v = self.video
        t = v.counter.value
        corr_op = AltCorrBlock(v.fmaps[None, :t])
        ht, wd = self.ht, self.wd
        for _ in range(steps):
            coords1, _ = v.reproject(self.ii, self.jj)
            motn = torch.cat([self.target_cam - self.coords0, self.target_cam - self.coords0 + self.delta_dy,
                              self.target_cam - coords1, self.raw_mask * torch.ones_like(coords1)[..., 0:self.mask_num]], dim=-1)
            motn = motn.permute(0, 1, 4, 2, 3).clamp(-64.0, 64.0)
            net_all = self.net                              # reference-layout hidden state (pulled out of the engine)
            s = 8
            for i in range(0, int(self.jj.max()) + 1, s):
                m = (self.ii >= i) & (self.ii < i + s)
                if not bool(m.any()):
                    continue
                iis, jjs = self.ii[m], self.jj[m]
                corr1 = corr_op(coords1[:, m], iis, jjs)
                net, delta, weight, damping, _, delta_m = self.update_op(net_all[:, m], v.inps[None, iis], corr1, motn[:, m], iis, jjs)
                net_all[:, m] = net.to(net_all.dtype)
                self.target_cam[:, m] = coords1[:, m] + delta[..., 0:2].float()
                self.damping[torch.unique(iis)] = damping[0].float()
                self.raw_mask[:, m] = self.raw_mask[:, m] + delta_m.float()
                bin_mask = (torch.sigmoid(self.raw_mask[:, m]) >= self.dy_thresh).float()
                self.delta_dy[:, m] = delta[..., 2:4].float() * (1 - bin_mask)
                self.weight[:, m] = torch.sigmoid(weight.float() + (1 - bin_mask) * 10)
            self.net = net_all
            # rows = the BA's depth frames unique(arange(1,t) U ii) (droid_kernels.cu:1311-1319; the reference indexes unique(ii) and
            # raises at `C + eta` when the two differ -- pvo_ba flags it, droid_backends.check_ba_meta)
            damping = self.damping[torch.unique(torch.cat([torch.arange(1, t, device=self.device), self.ii]))].contiguous() + EP
            target_cam = self.target_cam.view(-1, ht, wd, 2).permute(0, 3, 1, 2).contiguous()
            weight = self.weight.view(-1, ht, wd, 2).permute(0, 3, 1, 2).contiguous()
            v.ba(target_cam, weight, damping, self.ii, self.jj, 1, t, itrs=itrs, lm=1e-5, ep=1e-2, motion_only=False)
            if hasattr(v, 'dirty'):
                v.dirty[:t] = True

    def _lookup_mixed(self, coords1, eng):
        """volume dtype != engine dtype (e.g. half volume feeding the float32 parity engine)"""
        E = coords1.shape[1]
        tmp = torch.empty(E * self.ht * self.wd, pnn.CORR_STRIDE, dtype=self.corr.dtype, device=self.device)
        self.corr(coords1, out=tmp, layout=1, out_cstride=pnn.CORR_STRIDE)
        eng.corr.copy_(tmp)
