"""Generate tests/golden/proximity.npz: the REFERENCE's FactorGraph.add_proximity_factors (factor_graph.py:372-429) run
unmodified on prescribed frame-distance matrices (build container only; needs /root/reference).

The method is called on a bare FactorGraph object whose `video.distance` returns the prescribed distances and whose
`add_factors` records the selected edges -- the selection logic itself (masking, suppression around existing edges,
greedy pick with NMS) is the reference's code, line for line."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_import   # noqa: E402


def run_case(ref, seed, t, t0, t1, rad, nms, thresh, n_exist):
    g = torch.Generator().manual_seed(seed)
    n = (t - t0) * (t - t1)
    d = torch.rand(n, generator=g) * 40.0
    d[torch.rand(n, generator=g) < 0.05] = 150.0                     # > 100 -> masked
    ex = torch.randint(0, t, (n_exist, 2), generator=g)
    fg = ref.factor_graph.FactorGraph.__new__(ref.factor_graph.FactorGraph)
    fg.device = 'cpu'
    fg.ii, fg.jj = ex[:n_exist // 2, 0].clone(), ex[:n_exist // 2, 1].clone()
    fg.ii_bad, fg.jj_bad = torch.zeros(0, dtype=torch.long), torch.zeros(0, dtype=torch.long)
    fg.ii_inac, fg.jj_inac = ex[n_exist // 2:, 0].clone(), ex[n_exist // 2:, 1].clone()
    fg.video = types.SimpleNamespace(counter=types.SimpleNamespace(value=t), distance=lambda ii, jj, beta: d.clone())
    got = {}
    fg.add_factors = lambda ii, jj, remove=False: got.update(ii=ii.clone(), jj=jj.clone())
    fg.add_proximity_factors(t0=t0, t1=t1, rad=rad, nms=nms, beta=0.25, thresh=thresh)
    return d, ex, torch.stack([got['ii'], got['jj']], 1)


def main():
    ref = ref_import.load()
    out = {}
    cfgs = [(0, 12, 0, 0, 2, 2, 16.0, 10), (1, 20, 0, 0, 2, 2, 16.0, 24), (2, 16, 3, 0, 2, 1, 12.0, 8), (3, 25, 5, 2, 3, 2, 20.0, 30),
            (4, 9, 0, 0, 2, 2, 16.0, 0), (5, 30, 0, 0, 2, 2, 8.0, 60)]
    for k, (seed, t, t0, t1, rad, nms, thresh, ne) in enumerate(cfgs):
        d, ex, es = run_case(ref, seed, t, t0, t1, rad, nms, thresh, ne)
        out['d%d' % k], out['ex%d' % k], out['es%d' % k] = d.numpy(), ex.numpy(), es.numpy()
        out['cfg%d' % k] = np.array([t, t0, t1, rad, nms, thresh, ne], dtype=np.float64)
        print(k, 'edges', es.shape[0])
    np.savez_compressed(os.path.join(HERE, 'proximity.npz'), **out)


if __name__ == '__main__':
    main()
