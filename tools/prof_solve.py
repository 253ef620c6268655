"""Where the cycles of the large-window BA solve go: builds csrc/ba.cu ALONE with -DPVO_SOLVE_PROF into tools/_prof/libpvo_ba_prof.so
(cycle probes on thread 0, see PROF() in ba.cu; never part of the product library) and runs one rank's BA of the 8-GPU window.
  python tools/prof_solve.py build      (here, no GPU)        python tools/prof_solve.py [F ...]      (on the B200)"""
import ctypes, os, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
SO = os.path.join(ROOT, 'tools', '_prof', 'libpvo_ba_prof.so')
if sys.argv[1:2] == ['build']:
    from pvo_b200 import build as b
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    cmd = [b._nvcc()] + b.NVCC_FLAGS + ['-DPVO_SOLVE_PROF', '-I', b.CSRC, '-shared', os.path.join(b.CSRC, 'ba.cu'), '-o', SO]
    subprocess.run(cmd, check=True)
    print('built', SO)
    sys.exit(0)
import torch
from tests.golden import cases
from pvo_b200 import droid_backends, _lib
L = ctypes.CDLL(SO)
real = _lib.lib()
RADIUS = int(os.environ.get('DIAG_BA_RADIUS', '3'))
NAMES = {0: 'panel (thread 0 row)', 1: 'wait panel barrier', 2: 'next diag: 21-entry update', 3: 'next diag: factorisation', 4: 'wait trailing barrier',
         9: 'damp + load', 10: 'two eliminations', 11: 'merge', 12: 'separator', 13: 'back substitution'}
for F in ([int(a) for a in sys.argv[1:]] or [64]):
    c = cases.graph_case(F=F, radius=RADIUS, h=60, w=80, seed=1)
    ii, jj = c['ii'].cuda(), c['jj'].cuda()
    own = ii < 8
    ii, jj = ii[own].contiguous(), jj[own].contiguous()
    E = ii.shape[0]
    g = torch.Generator(device='cuda').manual_seed(0)
    poses, disps = c['poses'].cuda(), c['disps'].cuda()
    intr = c['intrinsics'][0].cuda().contiguous()
    tgt = torch.rand(E, 2, 60, 80, device='cuda', generator=g) * torch.tensor([80.0, 60.0], device='cuda').view(1, 2, 1, 1)
    wgt = torch.rand(E, 2, 60, 80, device='cuda', generator=g)
    kx = torch.unique(torch.cat([torch.arange(1, F, device='cuda'), ii]))
    K, P, HW = kx.shape[0], F - 1, 4800
    eta = torch.full((K, 60, 80), 0.02, device='cuda')
    nb = real.pvo_ba_workspace_bytes(E, F, P, K, HW)
    ws = torch.zeros(nb, dtype=torch.uint8, device='cuda')
    flags = droid_backends.band_flag(RADIUS)
    fn = L.pvo_ba
    fn.restype = ctypes.c_int
    fn.argtypes = _lib._SIGS['pvo_ba'][1]
    calls = 10
    buf = (ctypes.c_longlong * 32)()
    for rep in range(2):
        L.pvo_solve_prof_read(buf, 1)
        for _ in range(calls):
            p, d = poses.clone(), disps.clone()
            rc = fn(p.data_ptr(), d.data_ptr(), intr.data_ptr(), tgt.data_ptr(), wgt.data_ptr(), eta.data_ptr(), ii.data_ptr(), jj.data_ptr(), E, F, 60, 80, 1, F, K,
                    2, 1e-4, 0.1, 0, 0, None, None, ws.data_ptr(), nb, flags, torch.cuda.current_stream().cuda_stream)
            assert rc == 0, rc
        torch.cuda.synchronize()
        L.pvo_solve_prof_read(buf, 0)
    print('radius %d F=%d P=%d: cycles of thread 0 per solve (x %d solves); slots 0-4 are summed over ALL ldl_forward calls' % (RADIUS, F, P, 2 * calls))
    for i in range(32):
        if buf[i]:
            print('  slot %2d  %-32s %9.0f cycles  = %6.1f us at 1.965 GHz' % (i, NAMES.get(i, ''), buf[i] / (2 * calls), buf[i] / (2 * calls) / 1965.0))
