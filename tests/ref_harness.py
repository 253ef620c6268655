"""Helpers shared by the GPU tests and bench.py's parity-gate leg: build the REFERENCE's own FactorGraph / DepthVideo /
DynamicUpdateModule (staged bytecode, oracle/ref_stage.py) on the GPU over pvo_b200.droid_backends + pvo_b200.lietorch,
and this package's graph on the same seeded case.  Test infrastructure -- nothing in pvo_b200/ imports this."""
import torch


def reference_graph(c, sd, fp32, segm_filter=False, thresh=0.8, buffer=None):
    """-> (video, fg, update_module): the reference classes, unmodified, on cuda:0.  fp32=True switches the reference's
    `autocast(enabled=True)` regions to float32 (what its CPU run -- the golden vectors -- computes)."""
    from oracle import ref_stage
    ref = ref_stage.load()
    ref_stage.force_fp32(fp32)
    F_, h, w = c['F'], c['h'], c['w']
    video = ref.depth_video.DepthVideo(image_size=[8 * h, 8 * w], buffer=buffer or F_ + 3, device='cuda:0',
                                       segm_filter=segm_filter, thresh=thresh)
    video.poses[:F_] = c['poses'].cuda()
    video.disps[:F_] = c['disps'].cuda()
    video.intrinsics[:F_] = c['intrinsics'].cuda()
    video.fmaps[:F_] = c['fmaps'].cuda()
    video.nets[:F_] = c['nets'].cuda()
    video.inps[:F_] = c['inps'].cuda()
    if 'segms' in c:
        video.segms[:F_] = c['segms'].cuda()
    video.counter.value = F_
    m = ref.droid_net.DynamicUpdateModule().cuda()
    m.load_state_dict(sd, strict=True)
    m.eval()
    if fp32:
        def update_fn(net, inp, corr, motn, ii, jj, flag=False):      # as the golden generator: cast by hand
            return m(net.float(), inp.float(), corr.float(), motn, ii, jj, flag)
    else:
        update_fn = m
    fg = ref.factor_graph.FactorGraph(video, update_fn, device='cuda:0', corr_impl='volume', max_factors=-1)
    fg.add_factors(c['ii'].cuda(), c['jj'].cuda())
    return video, fg, m


def our_graph(c, sd, precision, segm_filter=False, thresh=0.8, buffer=None, use_cuda_graph=False):
    from pvo_b200.depth_video import DepthVideo
    from pvo_b200.droid_net import DynamicUpdateModule
    from pvo_b200.factor_graph import FactorGraph
    F_, h, w = c['F'], c['h'], c['w']
    v = DepthVideo(image_size=(8 * h, 8 * w), buffer=buffer or F_ + 3, device='cuda', segm_filter=segm_filter, thresh=thresh,
                   store_images=False)
    v.poses[:F_] = c['poses'].cuda()
    v.disps[:F_] = c['disps'].cuda()
    v.intrinsics[:] = c['intrinsics'][0].cuda()
    v.fmaps[:F_] = c['fmaps'].cuda()
    v.nets[:F_] = c['nets'].cuda()
    v.inps[:F_] = c['inps'].cuda()
    if 'segms' in c:
        v.segms[:F_] = c['segms'].cuda()
    v.counter.value = F_
    m = DynamicUpdateModule(precision=precision).cuda()
    m.load_state_dict(sd)
    fg = FactorGraph(v, m, device='cuda', precision=precision, use_cuda_graph=use_cuda_graph)
    fg.add_factors(c['ii'], c['jj'])
    return v, fg


def pose_err(a, b):
    """max relative translation error and max rotation angle between (n,7) pose sets (SURVEY.md 8(d) observables)"""
    dt = (a[:, :3] - b[:, :3]).norm(dim=-1) / b[:, :3].norm(dim=-1).clamp(min=1e-2)
    qa, qb = a[:, 3:] / a[:, 3:].norm(dim=-1, keepdim=True), b[:, 3:] / b[:, 3:].norm(dim=-1, keepdim=True)
    av, aw, bv, bw = -qa[:, :3], qa[:, 3:], qb[:, :3], qb[:, 3:]
    rv = aw * bv + bw * av + torch.cross(av, bv, dim=-1)
    rw = aw[:, 0] * bw[:, 0] - (av * bv).sum(-1)
    ang = 2 * torch.atan2(rv.double().norm(dim=-1), rw.double().abs())
    return dt.max().item(), ang.max().item()


def deviation(poses, disps, flow, ref_poses, ref_disps, ref_flow):
    """the four observables of the parity contract, as a dict of floats"""
    dt, ang = pose_err(poses.float().cpu(), ref_poses.float().cpu())
    drel = ((disps - ref_disps).abs() / ref_disps.abs().clamp(min=1e-3))
    return dict(pose_rel=dt, rot_rad=ang, depth_rel_mean=drel.mean().item(), depth_rel_max=drel.max().item(),
                flow_epe=(flow - ref_flow).norm(dim=-1).mean().item())


def fp16_gate(steps=2, F=8, radius=3, h=60, w=80, seed=1):
    """at config 2, how far is the benchmarked half / tensor-core path from the float32 path, next to how far
    the REFERENCE's own inference mode (update operator under torch.autocast on cuDNN, half correlation volume) is from it on the
    same GPU.  float32 yardstick = this package's float32 path, itself gated against the CPU oracle at this size
    (tests/test_config2_gpu.py).  Returns {'ours_fp16': {...}, 'reference_autocast': {...}}."""
    from oracle import update_op as oupd
    from tests.golden import cases
    c = cases.graph_case(F=F, radius=radius, h=h, w=w, seed=seed)
    sd = oupd.init_params(0)
    out = {}
    with torch.no_grad():
        v32, g32 = our_graph(c, sd, 'fp32')
        for _ in range(steps):
            g32.update(itrs=2)
        p32, d32, f32 = v32.poses[:F].clone(), v32.disps[:F].clone(), g32.full_flow.clone()
        del v32, g32
        torch.cuda.empty_cache()
        v16, g16 = our_graph(c, sd, 'fp16')
        for _ in range(steps):
            g16.update(itrs=2)
        out['ours_fp16'] = deviation(v16.poses[:F], v16.disps[:F], g16.full_flow, p32, d32, f32)
        del v16, g16
        torch.cuda.empty_cache()
        vr, gr, _ = reference_graph(c, sd, fp32=False)
        for _ in range(steps):
            gr.update(itrs=2)
        out['reference_autocast'] = deviation(vr.poses[:F], vr.disps[:F], gr.full_flow, p32, d32, f32)
        del vr, gr
        torch.cuda.empty_cache()
    return out
