"""One frame-graph "update-op iteration" on CPU -- oracle, test infrastructure only.

Restates FactorGraph.update (VO_Module/droid_slam/factor_graph.py:227-307, the hot loop body):
  * :231-237  reproject (depth_video.py:154-163 -> geom/projective_ops.py:102-130) and the 8 motion channels
              [target-coords0 | target-coords0+delta_dy | target-coords1 | raw_mask], clamped to +-64
  * :240      4-level correlation lookup at coords1
  * :242-244  update operator (DynamicUpdateModule)
  * :246-247  default t0 = max(1, ii.min()+1)
  * :250-253  target_cam, raw weight logits, raw_mask += delta_m, bin_mask = sigmoid(raw_mask) >= 0.5
  * :256-276  panoptic segment filter: per (edge,label) segment, dynamic fraction > thresh => whole segment dynamic
  * :278-280  delta_dy = delta[...,2:4]*(1-bin_mask); weight = sigmoid(logit + 10*(1-bin_mask))
  * :283      damping[unique(ii)] = eta
  * :295-302  BA call with eta = .2*damping[unique(ii)] + EP, targets/weights permuted to (E,2,h,w)
  * :303-304  full_flow = coords1 + delta_dy - coords0
The update operator runs in float32 here (the reference autocasts it to fp16 on the GPU; parity of
the CUDA path is defined against float32, SURVEY.md section 7 "Tolerance vs precision").
"""
import numpy as np
import torch

from . import corr as ocorr
from . import geom
from . import ba_cuda
from . import update_op as oupd

DY_THRESH = 0.5


def segment_filter(bin_mask, segm, thresh):
    """factor_graph.py:256-276 on (1,E,h,w,2) bool bin_mask and (1,E,1,h,w) int32 segm.

    Integer restatement of the numpy code: key = (edge+1)*1e6 + label; a segment (edge,label!=0) whose
    dynamic-pixel count / pixel count > thresh is forced dynamic (bin_mask -> False on both channels).
    The reference compares ``dy_n / ori_dic[label] > thresh`` in Python float64.
    """
    E = bin_mask.shape[1]
    seg = segm[0, :, 0].numpy().astype(np.int64)                       # (E,h,w)
    dyn = ((bin_mask[0, ..., 0] == 0) | (bin_mask[0, ..., 1] == 0)).numpy()
    keep = np.ones(seg.shape, dtype=bool)
    for e in range(E):
        labels, inv, cnt = np.unique(seg[e], return_inverse=True, return_counts=True)
        dcnt = np.bincount(inv.reshape(-1), weights=dyn[e].reshape(-1).astype(np.float64),
                           minlength=labels.shape[0])
        for k, lab in enumerate(labels):
            # reference skips keys with label % 1e6 == 0, and dynamic counts of 0 never appear in dy_ky
            if lab == 0 or dcnt[k] == 0:
                continue
            if (dcnt[k] / cnt[k]) > thresh:
                keep[e][seg[e] == lab] = False
    keep = torch.from_numpy(keep)[None]
    out = bin_mask.clone()
    out[..., 0] = out[..., 0] & keep
    out[..., 1] = out[..., 1] & keep
    return out


def reproject(poses, disps, intrinsics, ii, jj):
    """DepthVideo.reproject: (F,7),(F,h,w),(F,4) -> coords (1,E,h,w,2), valid (1,E,h,w,1)."""
    return geom.projective_transform(poses[None], disps[None], intrinsics[None], ii, jj)


def graph_update(sd, st, t0=None, t1=None, itrs=2, EP=1e-7, motion_only=False,
                 segm_filter=False, thresh=0.8, lm=1e-4, ep=0.1, debug=None):
    """Mutates the state dict ``st`` exactly as FactorGraph.update mutates self / self.video.

    st keys: poses (buf,7), disps (buf,h,w), intrinsics (buf,4), ii, jj (E,), pyramid (list of L vols),
    net, inp (1,E,128,h,w), target_cam, weight, delta_dy (1,E,h,w,2), raw_mask (1,E,h,w,2),
    damping (buf,h,w), segm (1,E,1,h,w) int32 [if segm_filter], age (E,)
    """
    ii, jj = st['ii'], st['jj']
    ht, wd = st['disps'].shape[1:]
    coords0 = geom.coords_grid(ht, wd)
    coords1, _ = reproject(st['poses'], st['disps'], st['intrinsics'], ii, jj)
    motn = torch.cat([st['target_cam'] - coords0, st['target_cam'] - coords0 + st['delta_dy'],
                      st['target_cam'] - coords1, st['raw_mask']], dim=-1)
    motn = motn.permute(0, 1, 4, 2, 3).clamp(-64.0, 64.0)
    corr = ocorr.corr_lookup(st['pyramid'], coords1, 3).float()
    net, delta, weight, damping, upmask, delta_m = oupd.update_op(
        sd, st['net'].float(), st['inp'].float(), corr, motn, ii)
    st['net'] = net
    if t0 is None:
        t0 = max(1, int(ii.min()) + 1)
    st['target_cam'] = coords1 + delta[..., 0:2]
    st['raw_mask'] = st['raw_mask'] + delta_m
    bin_mask = torch.sigmoid(st['raw_mask']) >= DY_THRESH
    if segm_filter:
        bin_mask = segment_filter(bin_mask, st['segm'], thresh)
    bin_mask = bin_mask.float()
    st['delta_dy'] = delta[..., 2:4] * (1 - bin_mask)
    st['weight'] = torch.sigmoid(weight + (1 - bin_mask) * 10)
    kx = torch.unique(ii)
    st['damping'][kx] = damping[0]
    eta = .2 * st['damping'][kx].contiguous() + EP
    target = st['target_cam'].view(-1, ht, wd, 2).permute(0, 3, 1, 2).contiguous()
    wgt = st['weight'].view(-1, ht, wd, 2).permute(0, 3, 1, 2).contiguous()
    if debug is not None:
        debug.update(coords1=coords1, motn=motn, corr=corr, delta=delta, weight_logits=weight,
                     eta=damping, delta_m=delta_m, target=target.clone(), wgt=wgt.clone(),
                     ba_eta=eta.clone(), t0=t0, upmask=upmask)
        debug['ba'] = []
    ba_cuda.depth_video_ba(st['poses'], st['disps'], st['intrinsics'][0], target, wgt, eta, ii, jj,
                           t0, t1, itrs, lm, ep, motion_only)
    st['full_flow'] = coords1 + st['delta_dy'] - coords0
    if 'age' in st:
        st['age'] = st['age'] + 1
    return st


def graph_update_lowmem(sd, st, fmaps, inps, t, itrs=2, EP=1e-7, steps=2, chunk=8):
    """FactorGraph.update_lowmem (factor_graph.py:309-360): the global-BA form of the update -- correlation features
    computed on the fly from the feature pyramid (AltCorrBlock), the update operator run on chunks of edges grouped by
    source frame (8 frames per chunk), ONE dense BA over the whole window [1, t) with lm=1e-5, ep=1e-2 per step.
    Mutates ``st`` (same keys as graph_update; no correlation pyramid needed).  fmaps / inps: (buf,128,h,w)."""
    ii, jj = st['ii'], st['jj']
    ht, wd = st['disps'].shape[1:]
    coords0 = geom.coords_grid(ht, wd)
    for _ in range(steps):
        coords1, _ = reproject(st['poses'], st['disps'], st['intrinsics'], ii, jj)
        motn = torch.cat([st['target_cam'] - coords0, st['target_cam'] - coords0 + st['delta_dy'],
                          st['target_cam'] - coords1, st['raw_mask']], dim=-1)
        motn = motn.permute(0, 1, 4, 2, 3).clamp(-64.0, 64.0)
        for i in range(0, int(jj.max()) + 1, chunk):
            v = (ii >= i) & (ii < i + chunk)
            if not bool(v.any()):
                continue
            iis, jjs = ii[v], jj[v]
            corr1 = ocorr.alt_corr_lookup(fmaps[None, :t].float(), coords1[:, v], iis, jjs)
            net, delta, weight, damping, _, delta_m = oupd.update_op(sd, st['net'][:, v].float(), inps[None, iis].float(), corr1,
                                                                     motn[:, v], iis)
            st['net'][:, v] = net.to(st['net'].dtype)
            st['target_cam'][:, v] = coords1[:, v] + delta[..., 0:2]
            st['damping'][torch.unique(iis)] = damping[0]
            st['raw_mask'][:, v] = st['raw_mask'][:, v] + delta_m
            bin_mask = (torch.sigmoid(st['raw_mask'][:, v]) >= DY_THRESH).float()
            st['delta_dy'][:, v] = delta[..., 2:4] * (1 - bin_mask)
            st['weight'][:, v] = torch.sigmoid(weight + (1 - bin_mask) * 10)
        eta = st['damping'][torch.unique(ii)].contiguous() + EP
        target = st['target_cam'].view(-1, ht, wd, 2).permute(0, 3, 1, 2).contiguous()
        wgt = st['weight'].view(-1, ht, wd, 2).permute(0, 3, 1, 2).contiguous()
        ba_cuda.depth_video_ba(st['poses'], st['disps'], st['intrinsics'][0], target, wgt, eta, ii, jj, 1, t, itrs, 1e-5, 1e-2, False)
    return st


def make_state(case, fdtype=None):
    """FactorGraph.__init__ + add_factors (factor_graph.py:12-58, 107-161) for a tests/golden/cases.py
    graph case: volume pyramid for every edge, net/inp gathered by ii, target = reprojection,
    zero weight / raw_mask / delta_dy, damping = 1e-6."""
    ii, jj = case['ii'], case['jj']
    fm = case['fmaps'] if fdtype is None else case['fmaps'].to(fdtype)
    st = dict(ii=ii.clone(), jj=jj.clone(), poses=case['poses'].clone(), disps=case['disps'].clone(),
              intrinsics=case['intrinsics'].clone())
    st['pyramid'] = ocorr.corr_pyramid(fm[ii][None], fm[jj][None])
    st['net'] = case['nets'][ii][None].clone()
    st['inp'] = case['inps'][ii][None].clone()
    target, _ = reproject(st['poses'], st['disps'], st['intrinsics'], ii, jj)
    st['target_cam'] = target
    st['weight'] = torch.zeros_like(target)
    st['raw_mask'] = torch.zeros_like(target)
    st['delta_dy'] = torch.zeros_like(target)
    st['damping'] = 1e-6 * torch.ones_like(st['disps'])
    st['age'] = torch.zeros_like(ii)
    if 'segms' in case:
        st['segm'] = case['segms'][ii][None].clone()
    return st
