"""2-GPU parity of the sharded frame graph (pvo_b200/dist.py): edges split by source keyframe, one NCCL all-reduce of
the reduced camera system per Gauss-Newton iteration -- must reproduce the single-GPU update."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import update_op as oupd
from tests.golden import cases

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _graph(c, ii, jj, dev, rank=None, world=1, fpr=4):
    from pvo_b200.depth_video import DepthVideo
    from pvo_b200.droid_net import DynamicUpdateModule
    from pvo_b200.factor_graph import FactorGraph
    from pvo_b200 import dist as pdist
    F_, h, w = c['F'], c['h'], c['w']
    v = DepthVideo(image_size=(8 * h, 8 * w), buffer=F_ + 2, device=dev, store_images=False)
    v.poses[:F_] = c['poses'].to(dev)
    v.disps[:F_] = c['disps'].to(dev)
    v.intrinsics[:] = c['intrinsics'][0].to(dev)
    v.fmaps[:F_] = c['fmaps'].to(dev)
    v.nets[:F_] = c['nets'].to(dev)
    v.inps[:F_] = c['inps'].to(dev)
    v.counter.value = F_
    m = DynamicUpdateModule(precision='fp32').to(dev)
    m.load_state_dict(oupd.init_params(0))
    fg = FactorGraph(v, m, device=dev, precision='fp32')
    if world > 1:
        pdist.attach(fg, rank, world, fpr)
    fg.add_factors(ii, jj)
    return v, fg


def _worker(rank, world, port, out, F=8, fpr=4, h=16, w=24):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda:%d' % rank))
    from pvo_b200 import dist as pdist
    dev = 'cuda:%d' % rank
    c = cases.graph_case(F=F, radius=3, h=h, w=w, seed=6)
    ii, jj, own = pdist.shard_edges(c['ii'], c['jj'], rank, world, fpr)
    v, fg = _graph(c, ii, jj, dev, rank, world, fpr)
    for _ in range(2):
        fg.update(t0=1, t1=F, itrs=2)
    torch.cuda.synchronize()
    torch.save({'poses': v.poses[:F].cpu(), 'disps': v.disps[:F].cpu(), 'flow': fg.full_flow.cpu(), 'own': own}, out % rank)
    dist.destroy_process_group()


# F = 32: 31 free poses do not fit the dense shared-memory solve, so the ranks exchange the BAND-PACKED system (pvo_ba_exchange_range)
@pytest.mark.parametrize('F,fpr,h,w', [(8, 4, 16, 24), (32, 16, 8, 12)])
def test_sharded_update_matches_single_gpu(tmp_path, F, fpr, h, w):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    out = str(tmp_path / 'r%d.pt')
    mp.spawn(_worker, args=(2, _free_port(), out, F, fpr, h, w), nprocs=2, join=True)
    c = cases.graph_case(F=F, radius=3, h=h, w=w, seed=6)
    v, fg = _graph(c, c['ii'], c['jj'], 'cuda:0')
    for _ in range(2):
        fg.update(t0=1, t1=F, itrs=2)
    ref_p, ref_d, ref_f = v.poses[:F].cpu(), v.disps[:F].cpu(), fg.full_flow.cpu()
    r = [torch.load(out % k) for k in range(2)]
    for k in range(2):
        assert (r[k]['poses'] - ref_p).abs().max() < 2e-5, k                       # pose replicas identical to single GPU
        frames = slice(fpr * k, fpr * k + fpr)
        assert (r[k]['disps'][frames] - ref_d[frames]).abs().max() < 2e-5, k      # each rank owns its frames' depth
        assert (r[k]['flow'] - ref_f[:, r[k]['own']]).abs().max() < 2e-4, k
    assert torch.equal(r[0]['poses'], r[1]['poses'])                               # redundant solves are bit-identical


def _reduce_worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda:%d' % rank))
    from pvo_b200 import dist as pdist
    dev = 'cuda:%d' % rank
    red = pdist.make_reducer(torch.device(dev))
    res = {'kind': 'peer' if red is not None else 'nccl', 'ok': True, 'err': 0.0}
    if red is not None:
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        for it, n in enumerate([1, 7, 1806, 2048, 2049, 50000, 143262, 1806, 1806, 1806]):     # repeated sizes exercise both parities
            x = torch.randn(n, dtype=torch.float64, device=dev, generator=g)
            want = x.clone()
            dist.all_reduce(want)
            got = red(x.clone())
            res['err'] = max(res['err'], (got - want).abs().max().item() / max(want.abs().max().item(), 1e-30))
        # inside a CUDA graph: the epoch advances on the device, so replays stay in step across ranks
        x = torch.randn(1806, dtype=torch.float64, device=dev, generator=g)
        y = x.clone()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, capture_error_mode='thread_local'):
            red(y)
        for _ in range(5):
            y.copy_(x)
            gr.replay()
        want = x.clone()
        dist.all_reduce(want)
        res['err'] = max(res['err'], (y - want).abs().max().item() / want.abs().max().item())
    torch.cuda.synchronize()
    torch.save(res, out % rank)
    dist.destroy_process_group()


def test_peer_memory_allreduce_matches_nccl(tmp_path):
    """the one-shot NVLink push all-reduce (csrc/allreduce_p2p.cu) == ncclAllReduce on fp64 vectors, eagerly and replayed
    from a CUDA graph"""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    out = str(tmp_path / 'a%d.pt')
    mp.spawn(_reduce_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = [torch.load(out % k) for k in range(2)]
    if r[0]['kind'] != 'peer':
        pytest.skip('symmetric memory unavailable on this box; the NCCL path is covered by test_sharded_update_matches_single_gpu')
    assert all(x['err'] < 1e-14 for x in r), r
