"""Generate tests/golden/lowmem.npz: the REFERENCE's FactorGraph.update_lowmem (factor_graph.py:309-360) run unmodified on a CPU
DepthVideo (build container only; needs /root/reference).  Its AltCorrBlock / droid_backends.ba calls land in the oracle's
restatements (oracle/ref_import.py); what the fixture pins is the control flow of the global-BA form -- chunking by source
frame, per-chunk state updates, one BA per step with lm=1e-5 / ep=1e-2 -- i.e. oracle.step.graph_update_lowmem."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_import, update_op as oupd   # noqa: E402
from tests.golden import cases                      # noqa: E402


def main():
    torch.set_grad_enabled(False)
    ref = ref_import.load()
    m = ref.droid_net.DynamicUpdateModule()
    m.load_state_dict(oupd.init_params(0), strict=True)
    c = cases.graph_case(F=12, radius=2, h=16, w=24, seed=12)
    F_, h, w = c['F'], c['h'], c['w']
    video = ref.depth_video.DepthVideo(image_size=[8 * h, 8 * w], buffer=16, device='cpu')
    video.poses[:F_] = c['poses']
    video.disps[:F_] = c['disps']
    video.intrinsics[:F_] = c['intrinsics']
    video.fmaps[:F_] = c['fmaps']
    video.nets[:F_] = c['nets']
    video.inps[:F_] = c['inps']
    video.counter.value = F_

    def update_fn(net, inp, corr, motn, ii, jj, flag):      # autocast is a no-op on CPU: cast by hand
        out = m(net.float(), inp.float(), corr.float(), motn, ii, jj, flag)
        return (out[0].half(),) + tuple(out[1:])                # under autocast the hidden state comes back in half

    fg = ref.factor_graph.FactorGraph(video, update_fn, device='cpu', corr_impl='volume', max_factors=-1)
    fg.add_factors(c['ii'], c['jj'])
    fg.update_lowmem(itrs=2, steps=2)
    out = dict(poses=video.poses[:F_].clone(), disps=video.disps[:F_].clone(), target=fg.target_cam.clone(), weight=fg.weight.clone(),
               delta_dy=fg.delta_dy.clone(), raw_mask=fg.raw_mask.clone(), net=fg.net.float()[:, :, ::8].clone(),
               damping=fg.damping[:F_].clone())
    np.savez_compressed(os.path.join(HERE, 'lowmem.npz'), **{k: v.numpy() for k, v in out.items()})
    print({k: tuple(v.shape) for k, v in out.items()})


if __name__ == '__main__':
    main()
