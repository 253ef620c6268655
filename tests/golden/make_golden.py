"""Generate tests/golden/*.npz by running the REFERENCE's own Python on CPU (build container only).

    python tests/golden/make_golden.py

Needs /root/reference (read-only mount).  The reference modules are imported unmodified through
oracle/ref_import.py (shims for lietorch / torch_scatter / droid_backends, whose native code cannot be
built here -- SURVEY.md section 0).  What each fixture pins:

  update_op.npz   reference DynamicUpdateModule(+ConvGRU+GraphAgg).forward, float32       -> oracle.update_op
  geom.npz        reference geom.projective_ops.projective_transform (with Jacobians),
                  geom.ba.BA (2 steps) and MoBA on the shimmed SE3                       -> oracle.geom
  graph_step.npz  reference FactorGraph.update x2 on a CPU DepthVideo with the panoptic
                  segment filter ON (control flow, motion features, weights, damping);
                  its droid_backends.ba / corr_index_forward calls land in the oracle's
                  restatements (reference CUDA kernels cannot run here)                  -> oracle.step
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_import, se3, update_op as oupd   # noqa: E402
from tests.golden import cases                            # noqa: E402


def save(name, **arrs):
    out = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrs.items()}
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, {k: v.shape for k, v in out.items()})


def main():
    torch.set_grad_enabled(False)
    ref = ref_import.load()

    # ---- update operator ------------------------------------------------------------------
    sd = oupd.init_params(0)
    m = ref.droid_net.DynamicUpdateModule()
    m.load_state_dict(sd, strict=True)
    net, inp, corr, flow = cases.update_inputs(E=3, h=12, w=16, seed=0)
    ii = torch.tensor([0, 0, 2])
    jj = torch.tensor([1, 2, 0])
    r_net, r_delta, r_weight, r_eta, r_up, r_dm = m(net, inp, corr, flow, ii, jj)
    save('update_op.npz', net=r_net, delta=r_delta, weight=r_weight, eta=r_eta,
         upmask=r_up['disp'][:, :, ::16], delta_m=r_dm, ii=ii, jj=jj)

    # ---- geometry + python BA --------------------------------------------------------------
    c = cases.graph_case(F=5, radius=2, h=12, w=16, seed=3, fdtype=torch.float32)
    ii, jj = c['ii'], c['jj']
    g = torch.Generator().manual_seed(30)
    Gs = se3.SE3(c['poses'][None])
    disps, intr = c['disps'][None], c['intrinsics'][None]
    x1, valid, (Ji, Jj, Jz) = ref.pops.projective_transform(Gs, disps, intr, ii, jj, jacobian=True)
    target = x1 + 0.5 * torch.randn(x1.shape, generator=g)
    weight = torch.rand(x1.shape, generator=g)
    eta = 0.01 * torch.rand(1, torch.unique(ii).shape[0], 12, 16, generator=g)
    G1, d1 = ref.ba.BA(target, weight, eta, Gs, disps, intr, ii, jj, fixedp=2)
    G2, d2 = ref.ba.BA(target, weight, eta, G1, d1, intr, ii, jj, fixedp=2)
    Gm = ref.ba.MoBA(target, weight, eta, Gs, disps, intr, ii, jj, fixedp=1)
    save('geom.npz', coords=x1, valid=valid, Ji=Ji, Jj=Jj, Jz=Jz, target=target, weight=weight, eta=eta,
         ba1_poses=G1.data, ba1_disps=d1, ba2_poses=G2.data, ba2_disps=d2, moba_poses=Gm.data)

    # ---- reference FactorGraph.update on a CPU DepthVideo -----------------------------------
    c = cases.graph_case(F=5, radius=2, h=16, w=24, seed=4, with_segm=True)
    F_, h, w = c['F'], c['h'], c['w']
    video = ref.depth_video.DepthVideo(image_size=[8 * h, 8 * w], buffer=8, device='cpu',
                                       segm_filter=True, thresh=0.45)
    video.poses[:F_] = c['poses']
    video.disps[:F_] = c['disps']
    video.intrinsics[:F_] = c['intrinsics']
    video.fmaps[:F_] = c['fmaps']
    video.nets[:F_] = c['nets']
    video.inps[:F_] = c['inps']
    video.segms[:F_] = c['segms']
    video.counter.value = F_

    def update_fn(net, inp, corr, motn, ii, jj, flag):      # autocast is a no-op on CPU: cast by hand
        return m(net.float(), inp.float(), corr.float(), motn, ii, jj, flag)

    fg = ref.factor_graph.FactorGraph(video, update_fn, device='cpu', corr_impl='volume', max_factors=-1)
    fg.add_factors(c['ii'], c['jj'])
    out = {}
    for it in range(2):
        fg.update(t0=None, t1=None, itrs=2, use_inactive=False)
        out.update({'poses%d' % it: video.poses[:F_].clone(), 'disps%d' % it: video.disps[:F_].clone(),
                    'target%d' % it: fg.target_cam.clone(), 'weight%d' % it: fg.weight.clone(),
                    'delta_dy%d' % it: fg.delta_dy.clone(), 'raw_mask%d' % it: fg.raw_mask.clone(),
                    'full_flow%d' % it: fg.full_flow.clone(), 'net%d' % it: fg.net.float()[:, :, ::8].clone(),
                    'damping%d' % it: fg.damping[:F_].clone()})
    save('graph_step.npz', **out)
    make_droid_net(ref)


def droid_net_state():
    """seeded DroidNet state dict (fnet.*, cnet.*, update.*) shared by the golden generator and the GPU test"""
    from oracle import encoder as oenc
    sd = {}
    sd.update({'fnet.' + k: v for k, v in oenc.init_params(128, seed=11).items()})
    sd.update({'cnet.' + k: v for k, v in oenc.init_params(256, seed=12).items()})
    sd.update({'update.' + k: v for k, v in oupd.init_params(0).items()})
    return sd


def droid_net_inputs():
    from collections import OrderedDict
    g = torch.Generator().manual_seed(21)
    F_, H, W = 3, 128, 192
    images = torch.randint(0, 256, (1, F_, 3, H, W), generator=g).float()
    images = torch.nn.functional.avg_pool2d(images.view(F_, 3, H, W), 5, stride=1, padding=2).view(1, F_, 3, H, W)   # some structure
    poses = cases.trajectory(F_, g)[None]
    disps = cases.smooth_disps(F_, H // 8, W // 8, g)[None]
    intr = cases.intrinsics_for(H // 8, W // 8).repeat(1, F_, 1)
    graph = OrderedDict([(0, [1, 2]), (1, [0, 2]), (2, [0, 1])])
    return images, poses, disps, intr, graph


def make_droid_net(ref):
    m = ref.droid_net.DroidNet()
    m.load_state_dict(droid_net_state(), strict=True)
    m.eval()
    images, poses, disps, intr, graph = droid_net_inputs()
    fm, net, inp = m.extract_features(images.clone())
    Gs_list, disp_list, res_list, flow_list, mask_list = m(se3.SE3(poses), images.clone(), disps, intr, graph, num_steps=2,
                                                           fixedp=2, ret_flow=True, downsample=True)
    save('droid_net.npz', fmaps=fm[:, :, ::8], net=net[:, :, ::8], inp=inp[:, :, ::8], poses=Gs_list[-1].data,
         disps_up=disp_list[-1][:, :, ::4, ::4], residual=res_list[-1], flow=flow_list[-1], mask=mask_list[-1][:, :, ::4, ::4])


if __name__ == '__main__':
    main()
