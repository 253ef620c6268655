"""The orchestration layer (pvo_b200.motion_filter / droid_frontend / droid_backend / trajectory_filler) against the reference's
own, on CPU.  tests/golden/slam_loop.npz was produced by the reference's MotionFilter, DroidFrontend, DroidBackend and
PoseTrajectoryFiller driving the reference's FactorGraph / DepthVideo / CorrBlock (oracle shims for the native extensions).
Here the SAME lower layer (reference FactorGraph etc., imported from /root/reference -- build container only, the test is
skipped elsewhere) is driven by THIS package's orchestration classes: every decision (frames kept, keyframes removed, edges
after every frame) must be identical and the state equal to float rounding.  The full stack on the B200 kernels is compared
with the same golden file in tests/test_zz_slam_loop_gpu.py."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import ref_import

G = os.path.join(os.path.dirname(__file__), 'golden')
pytestmark = pytest.mark.skipif(not ref_import.available(), reason='/root/reference is only mounted in the build container')


def test_orchestration_reproduces_the_reference_loop(monkeypatch):
    from tests.golden.make_golden import droid_net_state
    from tests.golden.make_golden_slam_loop import loop_args, video_frames
    import tests.golden.make_golden_slam_loop as gen
    import pvo_b200.droid_backend as pb
    import pvo_b200.droid_frontend as pf
    import pvo_b200.motion_filter as pm
    import pvo_b200.trajectory_filler as pt
    gold = np.load(os.path.join(G, 'slam_loop.npz'))
    torch.manual_seed(0)
    ref = ref_import.load()
    for mod in (pf, pb, pt):
        monkeypatch.setattr(mod, 'FactorGraph', ref.factor_graph.FactorGraph)
    monkeypatch.setattr(pm, 'CorrBlock', ref.corr.CorrBlock)
    monkeypatch.setattr(gen, 'FILTER_THRESH', float(gold['filter_thresh']))
    monkeypatch.setattr(gen, 'KEYFRAME_THRESH', float(gold['keyframe_thresh']))
    m = ref.droid_net.DroidNet()
    m.load_state_dict(droid_net_state(), strict=True)
    m.eval()

    def update(net, inp, corr, *a, **k):
        r = m.update(net.float(), inp.float(), corr.float(), *a, **k)
        return ((r[0].half(),) + tuple(r[1:])) if update.half_state else r
    update.half_state = False
    net = types.SimpleNamespace(cnet=m.cnet, fnet=m.fnet, update=update)
    args = loop_args()
    frames = video_frames()
    with torch.no_grad():
        video = ref.depth_video.DepthVideo(args.image_size, args.buffer, args.device, args.segm_filter, args.thresh)
        filt = pm.MotionFilter(net, video, thresh=args.filter_thresh, device='cpu', autocast=False)
        front = pf.DroidFrontend(net, video, args)
        kept = []
        for n, (t, img, intr, seg) in enumerate(frames):
            kept.append(filt.track(t, img, None, intr, seg))
            front()
            assert video.counter.value == int(gold['count%d' % n]), n
            assert front.t1 == int(gold['t1_%d' % n]), n
            es = np.array(sorted(zip(front.graph.ii.tolist(), front.graph.jj.tolist())), dtype=np.int64).reshape(-1, 2)
            assert np.array_equal(es, gold['edges%d' % n]), n
            ia = np.array(sorted(zip(front.graph.ii_inac.tolist(), front.graph.jj_inac.tolist())), dtype=np.int64).reshape(-1, 2)
            assert np.array_equal(ia, gold['inactive%d' % n]), n
            c = video.counter.value
            assert np.abs(video.poses[:c + 1].numpy() - gold['poses%d' % n]).max() < 1e-5, n
            assert np.abs(video.disps[:c + 1].numpy() - gold['disps%d' % n]).max() < 1e-5, n
        n = video.counter.value
        assert kept == [True] * 8 + [False] * 6 and n == 7                     # 8 kept, one removed by the frontend
        assert np.array_equal(video.tstamp[:n].numpy(), gold['tstamp'])
        assert np.abs(video.poses[:n + 1].numpy() - gold['front_poses']).max() < 1e-5
        assert np.abs(video.disps[:n + 1].numpy() - gold['front_disps']).max() < 1e-5
        update.half_state = True
        pb.DroidBackend(net, video, args)(2)
        update.half_state = False
        assert np.abs(video.poses[:n].numpy() - gold['back_poses']).max() < 1e-5
        assert np.abs(video.disps[:n].numpy() - gold['back_disps']).max() < 1e-5
        traj = pt.PoseTrajectoryFiller(net, video, device='cpu', autocast=False)(frames)
        assert traj.data.shape == (len(frames), 7)
        assert np.abs(traj.data.numpy() - gold['trajectory']).max() < 1e-4
        assert video.counter.value == n                                           # the temporary frames are gone again


def test_depth_video_bookkeeping_equals_the_reference_class():
    """DepthVideo's host side (buffers, append / item assignment with ints, tensors and slices, negative indexing, normalize,
    counter rules) against the reference's DepthVideo on CPU tensors, operation by operation"""
    from pvo_b200.depth_video import DepthVideo
    ref = ref_import.load()
    g = torch.Generator().manual_seed(1)
    H, W, B = 64, 96, 12
    mine = DepthVideo(image_size=[H, W], buffer=B, device='cpu', segm_filter=True, thresh=0.7)
    theirs = ref.depth_video.DepthVideo(image_size=[H, W], buffer=B, device='cpu', segm_filter=True, thresh=0.7)
    names = ('tstamp', 'images', 'dirty', 'red', 'poses', 'disps', 'disps_up', 'intrinsics', 'fmaps', 'nets', 'inps', 'segms', 'full_flow')

    def same():
        assert mine.counter.value == theirs.counter.value
        for n in names:
            a, b = getattr(mine, n), getattr(theirs, n)
            assert a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b), n
    same()

    def item(k, full=True, batch=None):
        lead = () if batch is None else (batch,)
        r = lambda *s, **kw: torch.randn(*(lead + s), generator=g, **kw)
        it = [float(k) if batch is None else torch.arange(batch).float() + k, torch.randint(0, 255, lead + (3, H, W), generator=g, dtype=torch.uint8),
              r(7), r(H // 8, W // 8).abs() + 0.1, r(4).abs() + 1]
        if full:
            it += [r(128, H // 8, W // 8).half(), r(128, H // 8, W // 8).half(), r(128, H // 8, W // 8).half(),
                   torch.randint(0, 9, lead + (1, H // 8, W // 8), generator=g, dtype=torch.int32)]
        return tuple(it)
    for k in range(4):
        it = item(k)
        mine.append(*it), theirs.append(*it)
        same()
    it = item(9)
    it = it[:2] + (None, None) + it[4:]                    # pose / depth left alone, as the motion filter appends
    mine.append(*it), theirs.append(*it)
    same()
    it = item(20, full=False)
    mine[7] = it
    theirs[7] = it                                         # int index beyond the counter moves it
    same()
    idx = torch.tensor([2, 9])
    it = item(30, batch=2)
    mine[idx] = it
    theirs[idx] = it                                       # tensor index
    same()
    it = item(40, full=False, batch=3)
    mine[3:6] = it
    theirs[3:6] = it                                       # slice: counter untouched (trajectory filler)
    same()
    for i in (0, -1, 4):
        for a, b in zip(mine[i], theirs[i]):
            assert torch.equal(a, b)
    mine.normalize(), theirs.normalize()
    same()
    assert DepthVideo.format_indicies([1, 2], torch.tensor([[3], [4]]), 'cpu')[1].tolist() == [3, 4]


def test_factor_graph_bookkeeping_equals_the_reference_class(monkeypatch):
    """FactorGraph's edge bookkeeping -- add_factors (duplicate filter, eviction of the oldest edges at max_factors),
    rm_factors with and without storing the removed edges as inactive, rm_keyframe (index shift of active and inactive edges,
    video buffers moved down), clear_edges -- is tensor code: run it on CPU next to the reference's FactorGraph through the
    same operation sequence.  Only the two device stages add_factors touches are replaced in this test (the volume by a
    stand-in that counts edges, reprojection by the reference's)."""
    import pvo_b200.factor_graph as pfg
    from pvo_b200.depth_video import DepthVideo
    ref = ref_import.load()
    g = torch.Generator().manual_seed(2)
    H, W, B, F_ = 128, 160, 16, 9
    mine_v = DepthVideo(image_size=[H, W], buffer=B, device='cpu', segm_filter=True)
    ref_v = ref.depth_video.DepthVideo(image_size=[H, W], buffer=B, device='cpu', segm_filter=True)
    from oracle import se3 as ose3
    for k in range(F_):
        item = (float(k), torch.zeros(3, H, W, dtype=torch.uint8), ose3.se3_exp(0.05 * torch.randn(6, generator=g)),
                0.5 + torch.rand(H // 8, W // 8, generator=g), torch.tensor([W / 16.0, W / 16.0, W / 16.0, H / 16.0]),
                torch.randn(128, H // 8, W // 8, generator=g).half(), torch.randn(128, H // 8, W // 8, generator=g).half(),
                torch.randn(128, H // 8, W // 8, generator=g).half(), torch.randint(0, 5, (1, H // 8, W // 8), generator=g, dtype=torch.int32))
        mine_v.append(*item), ref_v.append(*item)

    class CountingVolume:                      # stands in for CorrBlock: FactorGraph only cats and masks it
        def __init__(self, f1, f2=None):
            self.n = f1 if isinstance(f1, int) else f1.shape[1]

        def cat(self, other):
            return CountingVolume(self.n + other.n)

        def __getitem__(self, keep):
            return CountingVolume(int(keep.sum()))
    monkeypatch.setattr(pfg, 'CorrBlock', CountingVolume)
    mine_v.reproject = lambda ii, jj: ref_v.reproject(ii, jj)
    mine = pfg.FactorGraph(mine_v, None, device='cpu', max_factors=14)
    theirs = ref.factor_graph.FactorGraph(ref_v, None, device='cpu', max_factors=14)

    def same(tag):
        for n in ('ii', 'jj', 'age', 'ii_inac', 'jj_inac'):
            assert torch.equal(getattr(mine, n), getattr(theirs, n)), (tag, n)
        assert mine.corr.n == mine.ii.shape[0] == theirs.corr.corr_pyramid[0].shape[0], tag      # the volume follows the edges
        for n in ('target_cam', 'weight', 'raw_mask', 'delta_dy', 'target_cam_inac', 'weight_inac', 'raw_mask_inac', 'delta_dy_inac'):
            a, b = getattr(mine, n), getattr(theirs, n)
            assert a.shape == b.shape and torch.equal(a, b), (tag, n)
        for a, b in ((mine.net, theirs.net), (mine.inp, theirs.inp), (mine.segm, theirs.segm)):
            assert (a is None) == (b is None), tag
            if a is not None:
                assert a.shape == b.shape and torch.equal(a.float(), b.float()), tag
        for n in ('poses', 'disps', 'intrinsics', 'fmaps', 'nets', 'inps', 'segms'):
            assert torch.equal(getattr(mine_v, n), getattr(ref_v, n)), (tag, n)

    def both(fn):
        fn(mine), fn(theirs)
    both(lambda f: f.add_neighborhood_factors(0, 5, r=2))
    same('neighbourhood')
    both(lambda f: f.add_factors(torch.tensor([0, 1, 5, 5]), torch.tensor([1, 0, 3, 4])))          # two repeats, two new
    same('repeats filtered')
    for f in (mine, theirs):                                                                        # what update() leaves behind
        f.age += torch.arange(f.age.shape[0]) % 7
        f.weight = torch.rand(f.weight.shape, generator=torch.Generator().manual_seed(3))
        f.target_cam = f.target_cam + 0.5
    same('aged')
    both(lambda f: f.rm_factors(f.age > 4, store=True))
    same('stored as inactive')
    both(lambda f: f.add_factors(torch.tensor([6, 7, 6, 5, 7]), torch.tensor([7, 6, 5, 6, 5]), remove=True))   # over budget -> evict oldest
    same('evicted at max_factors')
    assert mine.ii.shape[0] <= 14 and mine.ii_inac.shape[0] > 0
    both(lambda f: f.rm_keyframe(5))
    same('keyframe removed')
    both(lambda f: f.add_factors(torch.tensor([0, 2]), torch.tensor([1, 0])))                       # inactive edges are repeats too
    same('inactive are repeats')
    both(lambda f: f.rm_factors(f.ii == 0, store=False))
    same('dropped')
    both(lambda f: f.clear_edges())
    assert mine.ii.shape[0] == 0 and theirs.ii.shape[0] == 0


def test_gpu_test_restates_the_generator_frames_exactly():
    """tests/test_zz_slam_loop_gpu.py cannot import the generator on the GPU box (it needs the reference tree), so it restates
    the synthetic video: the two must be identical"""
    import importlib
    from tests.golden.make_golden_slam_loop import video_frames as generator_frames
    restated = importlib.import_module('tests.test_zz_slam_loop_gpu').video_frames()
    for a, b in zip(restated, generator_frames()):
        assert a[0] == b[0] and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and torch.equal(a[3], b[3])
    assert len(restated) == 14
