"""Edge selection for FactorGraph.add_proximity_factors (droid_slam/factor_graph.py:372-429): host-side graph logic.

The reference walks the distance matrix with Python loops and one `.item()` device sync per candidate.  Here the
distances come to the host once (one D2H copy of a (t-t0) x (t-t1) matrix) and the same greedy selection with the same
non-maximum suppression runs on numpy arrays -- identical edge list, no per-element synchronisation."""
import numpy as np


def _suppress(d, i, j, t0, t1, t, nms):
    """factor_graph.py:391-399 / 417-425: blank the (di, dj) diamond around edge (i, j)"""
    w = t - t1
    lim = max(min(abs(i - j) - 2, nms), 0)
    for di in range(-nms, nms + 1):
        for dj in range(-nms, nms + 1):
            if abs(di) + abs(dj) <= lim:
                i1, j1 = i + di, j + dj
                if t0 <= i1 < t and t1 <= j1 < t:
                    d[(i1 - t0) * w + (j1 - t1)] = np.inf


def select_proximity_edges(d, existing, t0, t1, t, rad=2, nms=2, thresh=16.0):
    """d: flat float array over the (i in [t0,t), j in [t1,t)) grid of frame distances (row-major, as meshgrid gives);
    existing: iterable of (i, j) edges already in the graph (active, bad and inactive).  Returns the list of (i, j)."""
    d = np.array(d, dtype=np.float32).copy()
    w = t - t1
    ii, jj = np.meshgrid(np.arange(t0, t), np.arange(t1, t), indexing='ij')
    ii, jj = ii.reshape(-1), jj.reshape(-1)
    d[ii - rad < jj] = np.inf
    d[d > 100] = np.inf
    for i, j in existing:
        i, j = int(i), int(j)
        if abs(i - j) <= 2:
            continue
        _suppress(d, i, j, t0, t1, t, nms)
    es = []
    for i in range(t0, t):
        for j in range(i + 1, min(i + rad + 1, t)):
            es.append((i, j))
            es.append((j, i))
    for k in np.argsort(d, kind='stable'):
        if d[k] > thresh:
            continue
        i, j = int(ii[k]), int(jj[k])
        es.append((i, j))
        es.append((j, i))
        _suppress(d, i, j, t0, t1, t, nms)
    return es
