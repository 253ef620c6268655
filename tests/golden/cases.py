"""Seeded synthetic inputs shared by the golden-vector generator, the oracle tests and the GPU parity tests.

Everything is produced by CPU ``torch.Generator``s, which are deterministic across machines, so the
committed golden OUTPUTS (tests/golden/*.npz, made by make_golden.py from the reference's own code)
can be compared against re-generated inputs anywhere.  Shapes follow SURVEY.md section 8(d).
"""
import math

import torch


def _gen(seed):
    return torch.Generator().manual_seed(seed)


def neighborhood_edges(F, radius):
    """FactorGraph.add_neighborhood_factors (factor_graph.py:362-370): all i != j with |i-j| <= radius."""
    ii, jj = [], []
    for i in range(F):
        for j in range(F):
            if i != j and abs(i - j) <= radius:
                ii.append(i)
                jj.append(j)
    return torch.tensor(ii, dtype=torch.long), torch.tensor(jj, dtype=torch.long)


def smooth_disps(F, h, w, g, lo=0.2, hi=2.0):
    coarse = torch.rand(F, 1, max(2, h // 8 + 1), max(2, w // 8 + 1), generator=g) * (hi - lo) + lo
    return torch.nn.functional.interpolate(coarse, size=(h, w), mode='bilinear', align_corners=True)[:, 0].contiguous()


def trajectory(F, g, noise=0.002):
    """frame k = Exp(k * [0.05,0,0.02, 0,0.01,0]) + small noise; quaternions renormalised."""
    from oracle import se3
    base = torch.tensor([0.05, 0.0, 0.02, 0.0, 0.01, 0.0])
    xi = torch.arange(F, dtype=torch.float32)[:, None] * base[None] + noise * torch.randn(F, 6, generator=g)
    xi[0] = 0
    return se3.se3_exp(xi)


def intrinsics_for(h, w):
    """TartanAir-like calibration at 1/8 resolution scaled to the grid: [fx,fy,cx,cy]."""
    return torch.tensor([w / 2.0, w / 2.0, w / 2.0, h / 2.0])


def update_inputs(E, h, w, seed=0, dtype=torch.float32):
    g = _gen(seed)
    net = torch.tanh(torch.randn(1, E, 128, h, w, generator=g))
    inp = torch.relu(torch.randn(1, E, 128, h, w, generator=g))
    corr = torch.randn(1, E, 196, h, w, generator=g)
    flow = 4.0 * torch.randn(1, E, 8, h, w, generator=g)
    return net.to(dtype), inp.to(dtype), corr.to(dtype), flow.to(dtype)


def graph_case(F=8, radius=3, h=60, w=80, seed=1, with_segm=False, fdtype=torch.float16):
    """A full frame-graph state (SURVEY.md section 8(d) config 2 when called with the defaults)."""
    g = _gen(seed)
    ii, jj = neighborhood_edges(F, radius)
    E = ii.shape[0]
    poses = trajectory(F, g)
    disps = smooth_disps(F, h, w, g)
    intr = intrinsics_for(h, w).repeat(F, 1)
    fmaps = torch.randn(F, 128, h, w, generator=g).to(fdtype)
    nets = torch.tanh(torch.randn(F, 128, h, w, generator=g)).to(fdtype)
    inps = torch.relu(torch.randn(F, 128, h, w, generator=g)).to(fdtype)
    case = dict(F=F, E=E, h=h, w=w, ii=ii, jj=jj, poses=poses, disps=disps, intrinsics=intr,
                fmaps=fmaps, nets=nets, inps=inps)
    if with_segm:
        segm = torch.zeros(F, 1, h, w, dtype=torch.int32)
        for f in range(F):
            # 4 stuff bands + a few rectangular "things" with panoptic ids trainId*10000 + k
            for b in range(4):
                segm[f, 0, b * h // 4:(b + 1) * h // 4] = (b + 3) * 10000 if b else 0   # label 0 is skipped by the filter
            n_things = 3 + int(torch.randint(0, 3, (1,), generator=g))
            for k in range(n_things):
                y0 = int(torch.randint(0, max(1, h - 4), (1,), generator=g))
                x0 = int(torch.randint(0, max(1, w - 4), (1,), generator=g))
                hh = int(torch.randint(2, max(3, h // 3), (1,), generator=g))
                ww = int(torch.randint(2, max(3, w // 3), (1,), generator=g))
                segm[f, 0, y0:y0 + hh, x0:x0 + ww] = (k % 3) * 10000 + k + 1
        case['segms'] = segm
    return case
