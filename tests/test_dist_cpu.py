"""Multi-GPU host logic on CPU (gloo, world_size 2): sharding edges by source keyframe makes the reduced camera
system additive across ranks, so one all-reduce of (H | b) reproduces the single-device system exactly
(pvo_b200/dist.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ba_cuda, geom
from pvo_b200 import dist as pdist
from tests.golden import cases


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case():
    c = cases.graph_case(F=8, radius=3, h=10, w=12, seed=5, fdtype=torch.float32)
    g = torch.Generator().manual_seed(55)
    ii, jj = c['ii'], c['jj']
    tgt, _ = geom.projective_transform(c['poses'][None], c['disps'][None], c['intrinsics'][None], ii, jj)
    tgt = (tgt + 0.3 * torch.randn(tgt.shape, generator=g))[0].permute(0, 3, 1, 2).contiguous()
    wgt = torch.rand(tgt.shape, generator=g)
    eta = 0.01 * torch.rand(8, 10, 12, generator=g) + 1e-6
    return c, tgt, wgt, eta


def _system(c, tgt, wgt, eta, sel, t0, t1):
    ii, jj = c['ii'][sel], c['jj'][sel]
    blk = ba_cuda.edge_blocks(c['poses'], c['disps'], c['intrinsics'][0], tgt[sel], wgt[sel], ii, jj)
    kx = torch.unique(torch.cat([torch.arange(t0, t1), ii]))
    A, bA, S, bS, _ = ba_cuda.reduced_system(blk, ii, jj, eta[kx], t0, t1, 10, 12)
    return torch.cat([(A - S).reshape(-1), bA - bS])


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    c, tgt, wgt, eta = _case()
    _, _, own = pdist.shard_edges(c['ii'], c['jj'], rank, world, 4)
    buf = _system(c, tgt, wgt, eta, own, 1, 8)
    pdist.allreduce_system(buf)
    if rank == 0:
        torch.save(buf, out)
    dist.destroy_process_group()


def test_shard_edges_partitions_by_source_frame():
    c, *_ = _case()
    seen = torch.zeros(c['E'], dtype=torch.int32)
    for r in range(2):
        ii, jj, own = pdist.shard_edges(c['ii'], c['jj'], r, 2, 4)
        assert ((ii // 4) == r).all()
        seen += own.int()
    assert (seen == 1).all()


def test_allreduced_system_equals_single_device_system(tmp_path):
    out = str(tmp_path / 'sys.pt')
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    c, tgt, wgt, eta = _case()
    full = _system(c, tgt, wgt, eta, torch.ones(c['E'], dtype=torch.bool), 1, 8)
    assert torch.allclose(got, full, rtol=1e-9, atol=1e-9 * full.abs().max())


def _span_worker(rank, world, port, out):
    """dist.attach on a stand-in graph: the bandwidth bound of the pose system must come out as the maximum over ranks, through a
    collective that every rank issues on every call (a rank-local cache in front of it could hit on one rank and miss on another)"""
    import types
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    video = types.SimpleNamespace(poses=torch.zeros(4, 7), disps=torch.ones(4, 2, 2), intrinsics=torch.ones(4, 4))
    fg = types.SimpleNamespace(video=video)
    pdist.attach(fg, rank, world, 2)
    assert fg.world == world and fg.rank == rank and fg.allreduce_kind == 'ncclAllReduce'      # no peer memory on CPU
    seq = [(3, 5), (3, 3), (7, 2), (3, 5)]            # (rank 0's local span, rank 1's): repeats included -- nothing may be cached
    got = [fg._span_reduce(a if rank == 0 else b) for a, b in seq]
    if rank == 0:
        torch.save(torch.tensor(got), out)
    dist.destroy_process_group()


def test_span_bound_is_the_maximum_over_ranks_on_every_call(tmp_path):
    out = str(tmp_path / 'span.pt')
    mp.spawn(_span_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert torch.load(out).tolist() == [5, 3, 7, 5]
