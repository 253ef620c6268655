"""The reference's Python VO layer, staged for the GPU box as BYTECODE (test infrastructure only).

`/root/reference` does not exist on the GPU box and reference sources are never copied into this repository.  Like
`build_ref.py` compiles the reference's CUDA files from where they lie into `oracle/_ref/*.so`, `stage()` compiles the
reference's own `droid_slam/*.py` files -- unmodified, from where they lie -- into sourceless `.pyc` files under
`oracle/_ref/py/droid_slam/` (git-ignored, shipped with the snapshot like the `.so` files).  `load()` imports them on the
GPU box with the native dependencies bound to THIS package:

    lietorch       -> pvo_b200.lietorch
    droid_backends -> pvo_b200.droid_backends
    torch_scatter  -> three lines of torch.index_add_ (the reference's only uses: droid_net.py:87, geom/ba.py:14-28)

which is exactly the substitution a PVO maintainer makes.  Used by
  * tests/test_dropin_gpu.py: the reference's unmodified FactorGraph / DepthVideo / CorrBlock / DynamicUpdateModule run on the
    B200 kernels and reproduce tests/golden/graph_step.npz (the drop-in proof);
  * the fp16 parity gate (tests/test_config2_gpu.py, bench.py's gate leg): the reference's update operator under
    torch.autocast (cuDNN) on the same GPU is the yardstick for the half/tensor-core path's deviation from float32.
"""
import importlib
import os
import py_compile
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_PY = '/root/reference/VO_Module/droid_slam'
OUT = os.path.join(HERE, '_ref', 'py', 'droid_slam')
FILES = ['droid_net.py', 'factor_graph.py', 'depth_video.py', 'motion_filter.py', 'droid_frontend.py',
         'modules/__init__.py', 'modules/corr.py', 'modules/gru.py', 'modules/extractor.py', 'modules/clipping.py',
         'geom/__init__.py', 'geom/projective_ops.py', 'geom/ba.py', 'geom/chol.py', 'geom/graph_utils.py']


def stage(force=False):
    """build container only: reference .py -> oracle/_ref/py/droid_slam/**/*.pyc (sourceless modules)"""
    if not os.path.isdir(REF_PY):
        return None
    for rel in FILES:
        src = os.path.join(REF_PY, rel)
        dst = os.path.join(OUT, rel + 'c')
        if not force and os.path.exists(dst) and os.path.getmtime(dst) >= os.path.getmtime(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        # dfile: tracebacks cite the reference path; unchecked-hash pyc so the (absent) source is never consulted
        py_compile.compile(src, cfile=dst, dfile='VO_Module/droid_slam/' + rel, doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    return OUT


def available():
    return os.path.exists(os.path.join(OUT, 'factor_graph.pyc'))


def _scatter_module():
    import torch
    m = types.ModuleType('torch_scatter')

    def scatter_sum(src, index, dim=1, dim_size=None):
        assert dim == 1
        n = int(index.max()) + 1 if dim_size is None else dim_size
        out = torch.zeros(src.shape[:1] + (n,) + src.shape[2:], dtype=src.dtype, device=src.device)
        return out.index_add_(1, index, src)

    def scatter_mean(src, index, dim=1, dim_size=None):
        s = scatter_sum(src, index, dim, dim_size)
        cnt = torch.zeros(s.shape[1], dtype=src.dtype, device=src.device).index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
        return s / cnt.clamp(min=1).view((1, -1) + (1,) * (src.dim() - 2))
    m.scatter_sum, m.scatter_mean = scatter_sum, scatter_mean
    return m


_NS = [None]
PRECISION = {'fp32': False}     # True: every `torch.cuda.amp.autocast(enabled=True)` region of the reference runs in float32


def force_fp32(on):
    """The reference decorates FactorGraph.update / add_factors with `torch.cuda.amp.autocast(enabled=True)` (factor_graph.py:106,
    227).  On CPU (where the golden vectors were generated) those regions run in float32 because CUDA autocast is inert there; this
    switch gives the same float32 run on the GPU without touching the reference: the decorator objects are created while the staged
    modules import, from the switchable subclass installed by load()."""
    PRECISION['fp32'] = bool(on)


def _switchable_autocast():
    import torch
    import warnings
    base = torch.cuda.amp.autocast

    class autocast(torch.amp.autocast_mode.autocast):
        def __init__(self, enabled=True, dtype=torch.float16, cache_enabled=True):
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                super().__init__('cuda', enabled=enabled, dtype=dtype, cache_enabled=cache_enabled)
            self._want = enabled

        def __enter__(self):
            self._enabled = self._want and not PRECISION['fp32'] and torch.cuda.is_available()
            return super().__enter__()
    return base, autocast


def load():
    """import the staged reference modules with their native dependencies bound to pvo_b200; returns a namespace"""
    if _NS[0] is not None:
        return _NS[0]
    if not available():
        raise RuntimeError('oracle/_ref/py is not staged (python -m oracle.ref_stage in the build container)')
    import numpy as np
    if not hasattr(np, 'range'):
        np.range = np.arange                                   # droid_net.py:295 (dead value)
    import pvo_b200.droid_backends as backends
    import pvo_b200.lietorch as lie
    if not hasattr(lie, 'SO3'):
        lie.SO3 = type('SO3', (), {})
        lie.Sim3 = type('Sim3', (), {})
    sys.modules['lietorch'] = lie
    sys.modules['droid_backends'] = backends
    sys.modules['torch_scatter'] = _scatter_module()
    if 'matplotlib' not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mp = types.ModuleType('matplotlib')
            mp.pyplot = types.ModuleType('matplotlib.pyplot')
            sys.modules['matplotlib'], sys.modules['matplotlib.pyplot'] = mp, mp.pyplot
    dr = types.ModuleType('data_readers')
    dru = types.ModuleType('data_readers.rgbd_utils')
    dru.compute_distance_matrix_flow = dru.compute_distance_matrix_flow2 = None      # training-only (geom/graph_utils.py:7)
    dr.rgbd_utils = dru
    sys.modules.setdefault('data_readers', dr)
    sys.modules.setdefault('data_readers.rgbd_utils', dru)
    if OUT not in sys.path:
        sys.path.insert(0, OUT)
    ns = types.SimpleNamespace()
    import torch
    base, sw = _switchable_autocast()
    torch.cuda.amp.autocast = sw
    try:
        _import_all(ns)
    finally:
        torch.cuda.amp.autocast = base
    _NS[0] = ns
    return ns


def _import_all(ns):
    for name, mod in (('droid_net', 'droid_net'), ('factor_graph', 'factor_graph'), ('depth_video', 'depth_video'),
                      ('corr', 'modules.corr'), ('gru', 'modules.gru'), ('pops', 'geom.projective_ops'), ('ba', 'geom.ba')):
        setattr(ns, name, importlib.import_module(mod))


if __name__ == '__main__':
    print(stage(force='--force' in sys.argv))
