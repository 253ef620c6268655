"""Import the UNMODIFIED reference Python (VO_Module/droid_slam) from /root/reference on CPU.

Only usable in the build container (``/root/reference`` does not exist on the GPU box); used by
``tests/golden/make_golden.py`` to generate the committed golden vectors that pin the oracle, and by
the ``-m "not gpu"`` tests that re-check them when the reference is present.

The reference's native dependencies cannot be built here, so three modules are
shimmed in ``sys.modules`` before import:
  * ``lietorch``       -> oracle.se3 (SE3 only; PVO instantiates no other group, projective_ops.py:81)
  * ``torch_scatter``  -> oracle.geom.scatter_sum / scatter_mean (call sites droid_net.py:87, geom/ba.py:14-28)
  * ``droid_backends`` -> oracle CPU restatements (corr_index_forward, altcorr_forward, ba, frame_distance)
plus ``numpy.range = numpy.arange`` for the dead-code bug at droid_net.py:295 and a stub ``matplotlib``.
"""
import os
import sys
import types

import numpy as np
import torch

REF_ROOT = '/root/reference/VO_Module'


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'droid_slam'))


def _make_lietorch():
    from . import se3
    m = types.ModuleType('lietorch')
    m.SE3 = se3.SE3

    class _Unused:
        pass
    m.SO3 = type('SO3', (_Unused,), {})
    m.Sim3 = type('Sim3', (_Unused,), {})
    m.RxSO3 = type('RxSO3', (_Unused,), {})
    m.cat = se3.cat
    m.stack = se3.stack
    return m


def _make_scatter():
    from . import geom
    m = types.ModuleType('torch_scatter')

    def scatter_sum(src, index, dim=1, dim_size=None):
        assert dim == 1
        if dim_size is None:
            dim_size = int(index.max()) + 1
        return geom.scatter_sum(src, index, dim_size)

    def scatter_mean(src, index, dim=1, dim_size=None):
        assert dim == 1
        return geom.scatter_mean(src, index, dim_size)
    m.scatter_sum = scatter_sum
    m.scatter_mean = scatter_mean
    return m


def _make_backends():
    from . import corr as ocorr
    from . import ba_cuda
    m = types.ModuleType('droid_backends')
    m.corr_index_forward = lambda volume, coords, radius: [ocorr.corr_index_forward(volume, coords, radius)]
    m.altcorr_forward = lambda f1, f2, coords, radius: [ocorr.altcorr_forward(f1, f2, coords, radius)]

    def ba(poses, disps, intrinsics, targets, weights, eta, ii, jj, t0, t1, itrs, lm, ep, motion_only):
        return list(ba_cuda.ba(poses, disps, intrinsics, targets, weights, eta, ii, jj,
                               t0, t1, itrs, lm, ep, motion_only))
    m.ba = ba
    m.frame_distance = lambda poses, disps, intrinsics, ii, jj, beta: ba_cuda.frame_distance(poses, disps, intrinsics, ii, jj, beta)
    return m


def load():
    """Returns a namespace with the reference modules: droid_net, gru, corr, ba, pops, chol."""
    if not available():
        raise RuntimeError('/root/reference is not mounted here')
    if not hasattr(np, 'range'):
        np.range = np.arange                                  # droid_net.py:295
    sys.modules.setdefault('lietorch', _make_lietorch())
    sys.modules.setdefault('torch_scatter', _make_scatter())
    sys.modules.setdefault('droid_backends', _make_backends())
    if 'matplotlib' not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mp = types.ModuleType('matplotlib')
            mp.pyplot = types.ModuleType('matplotlib.pyplot')
            sys.modules['matplotlib'] = mp
            sys.modules['matplotlib.pyplot'] = mp.pyplot
    path = os.path.join(REF_ROOT, 'droid_slam')
    if path not in sys.path:
        sys.path.insert(0, path)
    import droid_net            # noqa: E402
    import modules.gru as gru   # noqa: E402
    import modules.corr as corr  # noqa: E402
    import geom.ba as ba        # noqa: E402
    import geom.projective_ops as pops  # noqa: E402
    import geom.chol as chol    # noqa: E402
    ns = types.SimpleNamespace(droid_net=droid_net, gru=gru, corr=corr, ba=ba, pops=pops, chol=chol)
    try:
        import factor_graph     # noqa: E402
        import depth_video      # noqa: E402
        ns.factor_graph = factor_graph
        ns.depth_video = depth_video
    except Exception as e:      # pragma: no cover
        ns.factor_graph_error = repr(e)
    return ns
