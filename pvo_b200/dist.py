"""Multi-GPU frame graphs: edges sharded by SOURCE keyframe, one all-reduce of the reduced camera system per
Gauss-Newton iteration.

Why this sharding: every per-iteration quantity is either per-edge (correlation lookup, update-operator convolutions,
Jacobians) or a reduction over edges that share the same source frame ii -- GraphAgg's scatter_mean
(droid_net.py:83-88), the depth blocks C / w / Ei (droid_kernels.cu:1374-1378) and the Schur products, which pair only
rows of the same depth frame (droid_kernels.cu:1236-1251).  With all out-edges of a keyframe on one GPU everything is
local except the pose system: each rank forms its partial  A_g - S_g  (6P x 6P) and  b_g  (6P) in float64, ONE
all-reduce (sum) over NCCL / NVLink combines them (7 KB at P=7, 1.1 MB at P=63 -- latency bound, issued in-stream
on the compute stream), every rank solves the identical system redundantly and applies dx to its pose replica and dz
to the depth maps of the frames it owns.  Nothing else crosses GPUs per iteration.

The reference has no multi-GPU inference path; this is new capability, validated against the single-GPU result
(tests/test_dist_cpu.py for the host logic over gloo, tests/test_dist_gpu.py on 2 GPUs).
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib, droid_backends
from ._lib import check, ptr, stream


def owner_of(frame, frames_per_rank):
    return frame // frames_per_rank


def shard_edges(ii, jj, rank, world, frames_per_rank):
    """Edges whose SOURCE frame is owned by `rank` (contiguous blocks of frames_per_rank keyframes per rank)."""
    own = (torch.div(ii, frames_per_rank, rounding_mode='floor') == rank) if world > 1 else torch.ones_like(ii, dtype=torch.bool)
    return ii[own], jj[own], own


def allreduce_system(buf, group=None):
    """Sum the (H | b) float64 block across ranks; NCCL on the GPU box, gloo in the CPU tests."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf


class PeerAllReduce:
    """One-shot sum of a small fp64 vector across the ranks of one NVLink domain, in one kernel launch
    (csrc/allreduce_p2p.cu): every rank pushes its partial into a staging slot in each peer's memory and sums what
    arrived in its own.  The staging buffers are a torch symmetric-memory allocation (CUDA IPC / fabric handles exchanged
    by `rendezvous`); only raw peer pointers cross into the C ABI."""
    N_MAX = 160_000          # doubles; the reduced camera system of up to P = 64 free poses (36 P^2 + 7 P)

    def __init__(self, device, group=None):
        import torch.distributed._symmetric_memory as symm
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        stage_bytes = 2 * self.world * self.N_MAX * 8
        self.buf = symm.empty(stage_bytes + self.world * 256 * 4, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.stage_ptrs = (ctypes.c_void_p * self.world)(*ptrs)
        self.flag_ptrs = (ctypes.c_void_p * self.world)(*[p + stage_bytes for p in ptrs])
        self.epochs = torch.zeros(256, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group)                                   # every rank's flags are zero before anyone pushes

    def __call__(self, t):
        assert t.dtype == torch.float64 and t.is_contiguous() and t.numel() <= self.N_MAX
        check(_lib.lib().pvo_allreduce_oneshot(ptr(t), self.stage_ptrs, self.flag_ptrs, ptr(self.epochs), self.rank, self.world,
                                               t.numel(), self.N_MAX, stream()), 'allreduce_oneshot')
        return t


def make_reducer(device, group=None):
    """PeerAllReduce when the ranks share an NVLink / P2P domain, else None (allreduce_system then uses NCCL in-stream)."""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1 and dist.get_backend(group) == 'nccl'):
        return None
    ok = torch.ones(1, device=device)
    try:
        red = PeerAllReduce(device, group)
    except Exception as ex:                                   # no symmetric memory on this system: stay on NCCL, say so
        red, ok = None, torch.zeros(1, device=device)
        print('[pvo_b200.dist] peer-memory all-reduce unavailable (%r); using ncclAllReduce' % (ex,), flush=True)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)    # all ranks must take the same path
    return red if ok.item() > 0 else None


def system_view(ws, E, nframes, P, K, HW, flags=0, motion_only=False):
    """float64 view of what the ranks exchange inside a BA workspace (torch.uint8 tensor): the band-packed reduced camera system when
    `flags` carry the graph's bandwidth and the window is large (pvo_ba_exchange_range), else the dense {H | b | dA} block."""
    off, nb = ctypes.c_long(0), ctypes.c_long(0)
    check(_lib.lib().pvo_ba_exchange_range(E, nframes, P, K, HW, int(flags), 1 if motion_only else 0, ctypes.byref(off), ctypes.byref(nb)),
          'ba_exchange_range', launches=0)
    return ws[off.value:off.value + nb.value].view(torch.float64)


def ba_sharded(poses, disps, intrinsics, targets, weights, eta, ii, jj, t0, t1, iterations, lm, ep, motion_only,
               interleaved=False, group=None, workspace=None, max_degree=0, reducer=None, max_span=None, reuse_plan=False,
               sys_clean=False):
    """droid_backends.ba for a rank's shard of the edges: assemble -> all-reduce -> solve + update, per iteration.
    poses must be identical on all ranks on entry and are identical on exit."""
    E = ii.shape[0]
    nframes, ht, wd = disps.shape
    K, P = eta.shape[0], t1 - t0
    ws = workspace if workspace is not None else droid_backends.ba_workspace(E, nframes, P, K, ht * wd, poses.device)
    hints = ((int(max_degree) & 0xFFF) << 8) | droid_backends.band_flag(max_span)
    sysbuf = system_view(ws, E, nframes, P, K, ht * wd, hints, motion_only)
    dx = torch.empty(P, 6, dtype=torch.float32, device=poses.device)
    dz = torch.empty(K, ht * wd, dtype=torch.float32, device=poses.device)
    L = _lib.lib()

    def call(phase, reuse):
        flags = (1 if reuse else 0) | (phase << 1) | (16 if sys_clean else 0) | hints
        check(L.pvo_ba(ptr(poses), ptr(disps), ptr(intrinsics), ptr(targets), ptr(weights), ptr(eta), ptr(ii), ptr(jj), E,
                       nframes, ht, wd, int(t0), int(t1), K, 1, float(lm), float(ep), 1 if motion_only else 0,
                       1 if interleaved else 0, ptr(dx), ptr(dz), ptr(ws), ws.numel(), flags, stream()), 'ba',
              launches=(0 if reuse else 1) + 2)
    for it in range(int(iterations)):
        call(1, it > 0 or reuse_plan)
        if reducer is not None:
            reducer(sysbuf)
        else:
            allreduce_system(sysbuf, group)
        call(2, True)
    return [dx, dz]


def attach(fg, rank, world, frames_per_rank, group=None):
    """Make a FactorGraph's BA go through ba_sharded.  The caller adds only this rank's edges (shard_edges)."""
    video = fg.video
    reducer = make_reducer(video.poses.device, group) if video.poses.is_cuda else None
    fg.allreduce_kind = 'one-shot push over NVLink peer memory (own kernel)' if reducer is not None else 'ncclAllReduce'

    def ba(target, weight, eta, ii, jj, t0=1, t1=None, itrs=2, lm=1e-4, ep=0.1, motion_only=False, interleaved=False,
           reuse_plan=False, max_degree=0, max_span=None, workspace=None, sys_clean=False):
        # max_span arrives already maximised over ranks (fg._span_reduce below, called by FactorGraph when the edge set changes)
        out = ba_sharded(video.poses, video.disps, video.intrinsics[0].contiguous(), target, weight, eta, ii, jj, t0, t1,
                         itrs, lm, ep, motion_only, interleaved=interleaved, group=group, max_degree=max_degree, reducer=reducer,
                         max_span=max_span, workspace=workspace, reuse_plan=reuse_plan, sys_clean=sys_clean)
        check(_lib.lib().pvo_clamp_min(ptr(video.disps), video.disps.numel(), 0.001, stream()), 'clamp_min')
        return out
    def span_reduce(local_span):
        """The bandwidth bound must hold for the all-reduced system: max over ranks.  Called by FactorGraph exactly once per
        edge-set change (and once per inactive-slab build) on EVERY rank -- edge sets change in lockstep -- so the collective can
        never be issued by a subset of the ranks (a per-value cache here could hit on one rank and miss on another)."""
        if not (dist.is_initialized() and dist.get_world_size(group) > 1):
            return int(local_span)
        t = torch.tensor([int(local_span)], device=video.poses.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        return int(t.item())
    video.ba = ba
    fg._span_reduce = span_reduce
    fg.rank, fg.world = rank, world
    return fg
