"""Generate tests/golden/slam_loop.npz: the REFERENCE's per-frame orchestration -- MotionFilter.track, DroidFrontend
(initialisation, sliding-window updates, keyframe removal), DroidBackend (global pass, update_lowmem) and
PoseTrajectoryFiller -- run unmodified on CPU over a short synthetic video (build container only; needs /root/reference).

The reference classes are imported from /root/reference/VO_Module/droid_slam with the oracle shims of oracle/ref_import.py
standing in for the native extensions (lietorch, torch_scatter, droid_backends); the network is the reference's DroidNet
with the seeded weights shared with the other golden files; `net.update` is wrapped only to cast the half-stored state to
float (autocast does nothing on CPU).  Stored: per frame the keyframe count, the frontend's window end t1 and the edge list
after the call; the flow magnitudes the motion filter compared with its threshold; poses / depths of the keyframes after the
frontend, after the backend, and the filled-in trajectory of all frames."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_import                       # noqa: E402
from tests.golden.make_golden import droid_net_state   # noqa: E402

H, W, NFRAMES = 128, 160, 14


def loop_args():
    """the reference's defaults (evaluation_scripts/test_vo.py:63-83) with a shorter warm-up and thresholds placed inside the
    range this synthetic video produces, so that frames get skipped, kept, and one keyframe gets removed"""
    return types.SimpleNamespace(device='cpu', image_size=[H, W], buffer=32, segm_filter=True, thresh=0.8, filter_thresh=FILTER_THRESH,
                                 warmup=5, beta=0.6, frontend_nms=1, keyframe_thresh=KEYFRAME_THRESH, frontend_window=25, frontend_thresh=12.0,
                                 frontend_radius=2, backend_thresh=15.0, backend_radius=2, backend_nms=3, disable_vis=True)


FILTER_THRESH = 0.0        # set in main() after a dry run reports the magnitudes (kept in the npz)
KEYFRAME_THRESH = 0.0


GAINS = (1.0, 0.95, 0.35, 0.9, 0.3, 1.0, 0.55, 0.25, 0.85, 1.0, 0.4, 0.9, 0.3, 0.8)


def video_frames():
    """a smooth random texture sliding / zooming slowly + synthetic panoptic ids at 1/8 resolution; (tstamp, image uint8
    (3,H,W) BGR, intrinsics (4,), segments (1,H/8,W/8) int32)"""
    g = torch.Generator().manual_seed(5)
    big = torch.rand(1, 3, 52, 64, generator=g)
    tex = torch.nn.functional.interpolate(big, size=(2 * H, 2 * W), mode='bicubic', align_corners=True).clamp(0, 1)[0]
    intr = torch.tensor([W / 2.0, W / 2.0, W / 2.0, H / 2.0])
    seg = torch.zeros(1, H // 8, W // 8, dtype=torch.int32)
    seg[:, :, :5] = 10000
    seg[:, 4:, 5:] = 30000
    seg[:, 2:6, 8:12] = 120001
    out = []
    for k in range(NFRAMES):
        step = (0, 1, 1, 3, 6, 6, 7, 10, 14, 15, 19, 24, 25, 30)[k]          # uneven motion: some frames barely move
        y0, x0 = 10 + step, 20 + 2 * step
        # random weights do not measure motion; what moves the motion filter's flow magnitude is the image statistics the
        # (un-normalised) context encoder sees, so exposure changes from frame to frame
        gain, offset = GAINS[k], 0.5 * (1 - GAINS[k])
        img = ((tex[:, y0:y0 + H, x0:x0 + W] * gain + offset) * 255).round().to(torch.uint8)
        out.append((float(k), img, intr.clone(), seg.clone()))
    return out


def main():
    global FILTER_THRESH, KEYFRAME_THRESH
    torch.set_grad_enabled(False)
    torch.manual_seed(0)
    ref = ref_import.load()
    import droid_backend
    import droid_frontend
    import motion_filter
    import trajectory_filler
    m = ref.droid_net.DroidNet()
    m.load_state_dict(droid_net_state(), strict=True)
    m.eval()

    def update(net, inp, corr, *a, **k):
        r = m.update(net.float(), inp.float(), corr.float(), *a, **k)
        # update_lowmem writes the new hidden state back into its half buffer (factor_graph.py:339; under CUDA autocast the
        # operator returns half there): same cast here, as in make_golden_lowmem.py
        return ((r[0].half(),) + tuple(r[1:])) if update.half_state else r
    update.half_state = False
    net = types.SimpleNamespace(cnet=m.cnet, fnet=m.fnet, update=update)
    frames = video_frames()

    def run(record):
        args = loop_args()
        video = ref.depth_video.DepthVideo(args.image_size, args.buffer, args.device, args.segm_filter, args.thresh)
        filt = motion_filter.MotionFilter(net, video, thresh=args.filter_thresh, device='cpu')
        front = droid_frontend.DroidFrontend(net, video, args)
        mags, dists = [], []
        orig_update, orig_distance = filt.update, video.distance

        def spy_update(*a, **k):
            r = orig_update(*a, **k)
            mags.append(float(r[1][..., 0:2].norm(dim=-1).mean()))
            return r

        def spy_distance(ii=None, jj=None, beta=0.3, bidirectional=True):
            d = orig_distance(ii, jj, beta=beta, bidirectional=bidirectional)
            if ii is not None and len(ii) == 1:
                dists.append(float(d))
            return d
        filt.update, video.distance = spy_update, spy_distance
        for n, (t, img, intr, seg) in enumerate(frames):
            filt.track(t, img, None, intr, seg)
            front()
            es = sorted(zip(front.graph.ii.tolist(), front.graph.jj.tolist()))
            record('count%d' % n, np.array(video.counter.value))
            record('t1_%d' % n, np.array(front.t1))
            record('edges%d' % n, np.array(es, dtype=np.int64).reshape(-1, 2))
            record('inactive%d' % n, np.array(sorted(zip(front.graph.ii_inac.tolist(), front.graph.jj_inac.tolist())), dtype=np.int64).reshape(-1, 2))
            c = video.counter.value                      # state after this frame (triage aid for the GPU comparison)
            record('poses%d' % n, video.poses[:c + 1].numpy().copy())
            record('disps%d' % n, video.disps[:c + 1].numpy().copy())
        return args, video, front, mags, dists

    # The weights are random, so the predicted flow magnitude barely depends on the motion: place the two thresholds inside
    # gaps of the observed values, wide enough (margins below) that float32 rounding differences between implementations
    # cannot flip a decision, and such that frames get skipped, kept, and keyframes get removed.
    _, _, _, mags, _ = run(lambda k, v: None)
    print('dry run flow magnitudes', np.round(mags, 4))
    best = None
    srt = np.sort(mags)
    for a, b in zip(srt[:-1], srt[1:]):
        FILTER_THRESH = float(a + b) / 2
        _, video, _, m1, d1 = run(lambda k, v: None)
        margin = float(np.min(np.abs(np.array(m1) - FILTER_THRESH)))
        kept = video.counter.value
        print('filter %.5f -> %d keyframes, margin %.5f, %d distances' % (FILTER_THRESH, kept, margin, len(d1)))
        if 7 <= kept <= 11 and len(d1) >= 3 and (best is None or margin > best[0]):
            best = (margin, FILTER_THRESH)
    assert best is not None and best[0] > 2e-4, best
    FILTER_THRESH = best[1]
    _, _, _, _, d1 = run(lambda k, v: None)
    print('keyframe distances (nothing removed)', np.round(d1, 4))
    best = None
    srt = np.sort(d1)
    for a, b in zip(srt[:-1], srt[1:]):
        KEYFRAME_THRESH = float(a + b) / 2
        _, video, _, m2, d2 = run(lambda k, v: None)
        removed = sum(1 for d in d2 if d < KEYFRAME_THRESH)
        margin = min(float(np.min(np.abs(np.array(d2) - KEYFRAME_THRESH) / KEYFRAME_THRESH)), float(np.min(np.abs(np.array(m2) - FILTER_THRESH))) / 2e-4 * 0.05)
        print('keyframe thresh %.5f -> removed %d of %d, relative margin %.3f' % (KEYFRAME_THRESH, removed, len(d2), margin))
        if 1 <= removed < len(d2) and (best is None or margin > best[0]):
            best = (margin, KEYFRAME_THRESH)
    assert best is not None and best[0] > 0.03, best
    KEYFRAME_THRESH = best[1]
    out = {}
    args, video, front, mags, dists = run(out.__setitem__)
    print('filter', FILTER_THRESH, 'keyframe thresh', KEYFRAME_THRESH, 'magnitudes', np.round(mags, 4), 'distances', np.round(dists, 4))
    n = video.counter.value
    out.update(filter_thresh=np.array(FILTER_THRESH), keyframe_thresh=np.array(KEYFRAME_THRESH), flow_magnitudes=np.array(mags),
               keyframe_distances=np.array(dists), tstamp=video.tstamp[:n].numpy(), front_poses=video.poses[:n + 1].numpy().copy(),
               front_disps=video.disps[:n + 1].numpy().copy())
    backend = droid_backend.DroidBackend(net, video, args)
    update.half_state = True
    backend(2)
    update.half_state = False
    out.update(back_poses=video.poses[:n].numpy().copy(), back_disps=video.disps[:n].numpy().copy())
    # PoseTrajectoryFiller builds its FactorGraph with the default device argument ("cuda:0", trajectory_filler.py:67): point that
    # default at the CPU for this run
    init = ref.factor_graph.FactorGraph.__init__
    init.__defaults__ = tuple('cpu' if d == 'cuda:0' else d for d in init.__defaults__)
    filler = trajectory_filler.PoseTrajectoryFiller(net, video, device='cpu')
    traj = filler(frames)
    out['trajectory'] = traj.data.numpy()
    np.savez_compressed(os.path.join(HERE, 'slam_loop.npz'), **out)
    print('keyframes', n, 'tstamps', video.tstamp[:n].tolist(), 'trajectory', traj.data.shape)


if __name__ == '__main__':
    main()
