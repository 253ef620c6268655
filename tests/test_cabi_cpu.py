"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, and exports every symbol that
include/pvo_b200.h declares (no compute calls -- there is no GPU here), the Python binding table matches the header,
and the product package never touches the oracle or a CPU fallback."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'pvo_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(pvo_[a-z0-9_]+)\s*\(', src)))


def test_library_loads_and_exports_every_declared_symbol():
    from pvo_b200 import build, _lib
    path = build.build()
    assert os.path.exists(path)
    h = ctypes.CDLL(path)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(h, s), 'libpvo_b200.so does not export %s' % s
    assert _lib.lib().pvo_version() >= 100


def test_binding_table_matches_header():
    from pvo_b200 import _lib
    assert sorted(_lib.symbols()) == header_symbols()


def test_product_does_not_import_oracle_or_fall_back():
    """A product path that routes through the oracle (or any CPU fallback) voids every parity claim."""
    pkg = os.path.join(ROOT, 'pvo_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', txt, flags=re.M), f
                assert 'from oracle' not in txt and 'import oracle' not in txt, f
                assert '/root/reference' not in txt, f


def test_ops_fail_loudly_without_cuda():
    import torch
    from pvo_b200 import droid_backends
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    vol = torch.zeros(1, 4, 4, 4, 4)
    coords = torch.zeros(1, 2, 4, 4)
    with pytest.raises(RuntimeError):
        droid_backends.corr_index_forward(vol, coords, 3)


def test_out_of_scope_entry_points_say_so():
    import torch
    from pvo_b200 import droid_backends
    for name in ('projmap', 'depth_filter', 'iproj'):           # provided (tensor expressions), CUDA tensors only like the rest
        assert callable(getattr(droid_backends, name))
    with pytest.raises(RuntimeError):
        droid_backends.iproj(torch.zeros(2, 7), torch.ones(2, 4, 4), torch.ones(4))
    with pytest.raises(RuntimeError):                           # provided since round 2; CUDA tensors only
        droid_backends.altcorr_backward(torch.zeros(1, 4, 4, 8), torch.zeros(1, 4, 4, 8), torch.zeros(1, 1, 4, 4, 2), torch.zeros(1, 1, 49, 4, 4), 3)


def test_ba_workspace_size_is_monotone():
    from pvo_b200 import _lib
    L = _lib.lib()
    a = L.pvo_ba_workspace_bytes(36, 1024, 7, 8, 4800)
    b = L.pvo_ba_workspace_bytes(72, 1024, 7, 8, 4800)
    assert 0 < a < b
    assert L.pvo_ba_workspace_bytes(36, 1024, 0, 8, 4800) == -1


def test_ba_exchange_range_is_the_band_for_large_windows():
    """What the ranks all-reduce (host arithmetic only, no GPU): the dense {H | b | dA} block for windows that are solved densely in
    shared memory, for an unknown bandwidth and for the MoBA variant; the band-packed block [6P][12r+6] | b | dA otherwise."""
    import ctypes
    from pvo_b200 import _lib, droid_backends
    L = _lib.lib()

    def rng(P, flags, motion_only=0):
        off, nb = ctypes.c_long(0), ctypes.c_long(0)
        assert L.pvo_ba_exchange_range(42, 64, P, 8, 4800, flags, motion_only, ctypes.byref(off), ctypes.byref(nb)) == 0
        return off.value, nb.value

    dense = lambda P: 8 * (36 * P * P + 12 * P)
    band3 = droid_backends.band_flag(3)                        # radius 3 -> block bandwidth 6
    assert rng(7, band3)[1] == dense(7)                        # 7 poses: dense shared-memory solve
    assert rng(63, 0)[1] == dense(63)                          # bandwidth unknown
    assert rng(63, band3)[1] == 8 * (378 * 42 + 2 * 378)       # 133 KB instead of 1.1 MB
    assert rng(63, band3)[0] != rng(63, 0)[0]                  # its own block of the workspace
    assert rng(63, band3 | (1 << 3), motion_only=1)[1] == dense(63)      # geom/ba.py MoBA damps every block: stays dense
    assert rng(63, droid_backends.band_flag(8))[1] == 8 * (378 * 102 + 2 * 378)   # config 5: windowed factorisation, packed all the same
    off, nb = rng(63, band3)
    assert off + nb <= L.pvo_ba_workspace_bytes(42, 64, 63, 8, 4800)
