"""The whole per-frame loop on the B200 kernels -- pvo_b200.droid.Droid: MotionFilter -> DroidFrontend per frame, then
DroidBackend and PoseTrajectoryFiller -- against tests/golden/slam_loop.npz, the reference's own loop run on CPU in float32
(tests/golden/make_golden_slam_loop.py; thresholds sit in gaps of the decision variables at least 7e-4 (flow magnitude) and
18 % (keyframe distance) wide).  Float32 path (`autocast=False`), so every discrete decision must match; the continuous state
is compared loosely (about 40 update iterations feed back into each other).

The file name sorts last on purpose: should a path trap on the device, no other test file runs after it in the same process.
First hardware run (round 2): every discrete decision (keyframes kept / removed, edge lists after each frame, t1), the poses after
every frame (2e-3) and after the frontend matched; the largest relative depth difference over the 8 x 16 x 20 final depths was 3.9e-2
on a handful of weakly constrained pixels (confidence ~ 0: their depth is set by the damping term alone, so a 1e-6 difference in the
network output is amplified), hence depth is gated by its distribution: median, 99th percentile and a looser maximum."""
import os
import types

import numpy as np
import pytest
import torch

# the timeout ends the process instead of leaving the box hung should an untried path ever spin on the device
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(420, method='thread')]
G = os.path.join(os.path.dirname(__file__), 'golden')

H, W = 128, 160
GAINS = (1.0, 0.95, 0.35, 0.9, 0.3, 1.0, 0.55, 0.25, 0.85, 1.0, 0.4, 0.9, 0.3, 0.8)
STEPS = (0, 1, 1, 3, 6, 6, 7, 10, 14, 15, 19, 24, 25, 30)


def video_frames():
    """tests/golden/make_golden_slam_loop.py video_frames(), restated (the generator imports the reference tree)"""
    g = torch.Generator().manual_seed(5)
    big = torch.rand(1, 3, 52, 64, generator=g)
    tex = torch.nn.functional.interpolate(big, size=(2 * H, 2 * W), mode='bicubic', align_corners=True).clamp(0, 1)[0]
    intr = torch.tensor([W / 2.0, W / 2.0, W / 2.0, H / 2.0])
    seg = torch.zeros(1, H // 8, W // 8, dtype=torch.int32)
    seg[:, :, :5] = 10000
    seg[:, 4:, 5:] = 30000
    seg[:, 2:6, 8:12] = 120001
    out = []
    for k, step in enumerate(STEPS):
        y0, x0 = 10 + step, 20 + 2 * step
        gain, offset = GAINS[k], 0.5 * (1 - GAINS[k])
        img = ((tex[:, y0:y0 + H, x0:x0 + W] * gain + offset) * 255).round().to(torch.uint8)
        out.append((float(k), img, intr.clone(), seg.clone()))
    return out


def make_droid(gold):
    from pvo_b200.droid import Droid
    from tests.golden.make_golden import droid_net_state
    args = types.SimpleNamespace(weights=droid_net_state(), device='cuda:0', image_size=[H, W], buffer=32, segm_filter=True, thresh=0.8,
                                 filter_thresh=float(gold['filter_thresh']), warmup=5, beta=0.6, frontend_nms=1,
                                 keyframe_thresh=float(gold['keyframe_thresh']), frontend_window=25, frontend_thresh=12.0, frontend_radius=2,
                                 backend_thresh=15.0, backend_radius=2, backend_nms=3, disable_vis=True, autocast=False)
    return Droid(args)


def test_frontend_decisions_and_state():
    gold = np.load(os.path.join(G, 'slam_loop.npz'))
    droid = make_droid(gold)
    for n, (t, img, intr, seg) in enumerate(video_frames()):
        droid.track(t, img.cuda(), None, intr.cuda(), seg.cuda())
        assert droid.video.counter.value == int(gold['count%d' % n]), n
        assert droid.frontend.t1 == int(gold['t1_%d' % n]), n
        g = droid.frontend.graph
        es = np.array(sorted(zip(g.ii.tolist(), g.jj.tolist())), dtype=np.int64).reshape(-1, 2)
        assert np.array_equal(es, gold['edges%d' % n]), n
        c = droid.video.counter.value                # per-frame state: says at which frame a drift starts
        perr = np.abs(droid.video.poses[:c + 1].cpu().numpy() - gold['poses%d' % n]).max()
        assert perr < 2e-3, (n, perr)
    n = droid.video.counter.value
    assert np.array_equal(droid.video.tstamp[:n].cpu().numpy(), gold['tstamp'])
    assert np.abs(droid.video.poses[:n + 1].cpu().numpy() - gold['front_poses']).max() < 2e-3
    d, dg = droid.video.disps[:n + 1].cpu().numpy(), gold['front_disps']
    rel = np.abs(d - dg) / np.maximum(np.abs(dg), 1e-3)
    print('final depth relative difference: median %.2e  p99 %.2e  max %.2e' % (np.median(rel), np.quantile(rel, 0.99), rel.max()))
    assert np.median(rel) < 1e-3 and np.quantile(rel, 0.99) < 2e-2 and rel.max() < 1e-1


def test_backend_and_trajectory():
    gold = np.load(os.path.join(G, 'slam_loop.npz'))
    droid = make_droid(gold)
    frames = video_frames()
    for t, img, intr, seg in frames:
        droid.track(t, img.cuda(), None, intr.cuda(), seg.cuda())
    n = droid.video.counter.value
    droid.backend(2)
    assert np.abs(droid.video.poses[:n].cpu().numpy() - gold['back_poses']).max() < 5e-3
    traj = droid.traj_filler([(t, img.cuda(), intr.cuda(), seg) for t, img, intr, seg in frames])
    assert traj.data.shape == (len(frames), 7) and droid.video.counter.value == n
    assert np.abs(traj.data.cpu().numpy() - gold['trajectory']).max() < 1e-2
