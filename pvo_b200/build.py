"""Build libpvo_b200.so (all CUDA kernels + the C ABI) for sm_100a, in-tree.

    python -m pvo_b200.build [--force]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, '_obj')
LIB = os.path.join(HERE, 'libpvo_b200.so')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-Xptxas', '-v']


def _nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError('nvcc not found')


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _newest_header():
    return max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h')))


def _compile(src, force):
    obj = os.path.join(OBJ, src[:-3] + '.o')
    sp = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(sp), _newest_header()):
        return obj, None
    cmd = [_nvcc()] + NVCC_FLAGS + ['-I', CSRC, '-c', sp, '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
    with open(obj + '.ptxas.txt', 'w') as f:
        f.write(r.stderr)
    return obj, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, force), srcs))
    objs = [o for o, _ in results]
    rebuilt = any(log is not None for _, log in results)
    if rebuilt or force or not os.path.exists(LIB):
        cmd = [_nvcc(), '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    if verbose:
        for (o, log) in results:
            if log:
                print(log)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
