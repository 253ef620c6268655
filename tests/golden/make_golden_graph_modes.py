"""Generate tests/golden/graph_modes.npz: the REFERENCE's FactorGraph.update in the two modes the per-frame loop uses besides
the plain one -- `use_inactive=True` with stored (inactive) edges, after an eviction and a keyframe removal, and
`motion_only=True` on a window of appended frames (the trajectory filler's call) -- run unmodified on CPU (build container
only; needs /root/reference).  Same set-up as graph_step.npz in make_golden.py: a CPU DepthVideo, the reference update
operator with the seeded weights (inputs cast to float: autocast does nothing on CPU), oracle shims for the native extensions."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_import, update_op as oupd   # noqa: E402
from tests.golden import cases                      # noqa: E402

CASE = dict(F=7, radius=2, h=16, w=24, seed=6, with_segm=True)
EXTRA = 3                                            # frames appended behind the keyframes for the motion-only window


def fill_video(video, c, extra_poses=None):
    F_ = c['F']
    for name in ('poses', 'disps', 'intrinsics', 'fmaps', 'nets', 'inps', 'segms'):
        getattr(video, name)[:F_] = c[name]
    video.counter.value = F_


def script(fg, video, c, record):
    """the operation sequence; `record(tag)` is called after every update"""
    F_ = c['F']
    fg.add_factors(c['ii'], c['jj'])
    fg.update(None, None, itrs=2, use_inactive=True)             # nothing inactive yet
    record('a')
    old = (fg.ii <= 1) & (fg.jj <= 2)                             # retire the edges among the oldest frames, keep their estimates
    fg.rm_factors(old, store=True)
    fg.update(None, None, itrs=2, use_inactive=True)             # the stored edges still constrain the BA
    record('b')
    fg.rm_keyframe(F_ - 2)                                        # second-newest keyframe goes; indices above shift down
    video.counter.value = F_ - 1
    fg.update(None, None, itrs=2, use_inactive=True)
    record('c')
    fg.update(1, None, itrs=2, use_inactive=True)                # explicit t0, as the initialisation does
    record('d')


def motion_only_script(FG, video, update_op, c, record):
    """the trajectory filler's use: EXTRA frames appended behind the keyframes, each tied to two keyframes, poses-only BA"""
    n = video.counter.value
    g = torch.Generator().manual_seed(9)
    for k in range(EXTRA):
        video.poses[n + k] = video.poses[n - 1 - k]
        video.disps[n + k] = 1.0
        video.intrinsics[n + k] = video.intrinsics[0]
        video.fmaps[n + k] = torch.randn(128, c['h'], c['w'], generator=g).half()
    video.counter.value = n + EXTRA
    fg = FG(video, update_op, 'cpu')
    new = torch.arange(n, n + EXTRA)
    fg.add_factors(torch.tensor([n - 1, n - 2, n - 3]), new)
    fg.add_factors(torch.tensor([n - 2, n - 3, n - 4]), new)
    for it in range(2):
        fg.update(n, n + EXTRA, motion_only=True)
        record('m%d' % it, fg)
    video.counter.value = n


def main():
    torch.set_grad_enabled(False)
    ref = ref_import.load()
    m = ref.droid_net.DynamicUpdateModule()
    m.load_state_dict(oupd.init_params(0), strict=True)

    def update_fn(net, inp, corr, motn, ii, jj, flag):
        return m(net.float(), inp.float(), corr.float(), motn, ii, jj, flag)
    c = cases.graph_case(**CASE)
    video = ref.depth_video.DepthVideo(image_size=[8 * c['h'], 8 * c['w']], buffer=c['F'] + EXTRA + 2, device='cpu', segm_filter=True, thresh=0.45)
    fill_video(video, c)
    fg = ref.factor_graph.FactorGraph(video, update_fn, device='cpu', corr_impl='volume', max_factors=-1)
    out = {}

    def record(tag, graph=None):
        g_ = graph or fg
        n = video.counter.value
        out.update({'poses_' + tag: video.poses[:n].clone().numpy(), 'disps_' + tag: video.disps[:n].clone().numpy(),
                    'target_' + tag: g_.target_cam.clone().numpy(), 'weight_' + tag: g_.weight.clone().numpy(),
                    'ii_' + tag: g_.ii.clone().numpy(), 'jj_' + tag: g_.jj.clone().numpy(),
                    'ii_inac_' + tag: g_.ii_inac.clone().numpy(), 'jj_inac_' + tag: g_.jj_inac.clone().numpy()})
    script(fg, video, c, record)
    motion_only_script(ref.factor_graph.FactorGraph, video, update_fn, c, record)
    np.savez_compressed(os.path.join(HERE, 'graph_modes.npz'), **out)
    print({k: v.shape for k, v in out.items() if k.startswith(('poses', 'ii_'))})


if __name__ == '__main__':
    main()
