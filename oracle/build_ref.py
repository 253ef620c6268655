"""Compile the REFERENCE's own correlation kernels into oracle/_ref/ (build container only).

The reference's pybind module cannot be built as shipped (vendored Eigen lacks Eigen/Core), but
its two Eigen-free files compile for sm_100a once `X.type()` is replaced by `X.scalar_type()` at
correlation_kernels.cu:145,176 and altcorr_kernel.cu:309 (torch >= 2 removed the implicit conversion).  This recipe
  1. reads the sources WHERE THEY LIE under /root/reference (never copied into the repo),
  2. applies that 3-line patch to a temporary copy under /tmp,
  3. builds them with a binding file of our own (oracle/ref_binding.cpp) into oracle/_ref/droid_corr_ref.so.
The .so is git-ignored but travels to the GPU box, where tests/test_corr_gpu.py uses it to pin the bit-exactness of
both the CPU oracle and the B200 lookup kernel against the reference's actual CUDA code, and bench.py can time it.
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = '/root/reference/VO_Module/src'
OUT_DIR = os.path.join(HERE, '_ref')
OUT = os.path.join(OUT_DIR, 'droid_corr_ref.so')
OUT_BA = os.path.join(OUT_DIR, 'droid_ba_ref.so')


def build_if_possible(force=False):
    if not os.path.isdir(REF_SRC):
        return None
    if os.path.exists(OUT) and not force:
        return OUT
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix='pvo_ref_')
    try:
        srcs = []
        for name in ('correlation_kernels.cu', 'altcorr_kernel.cu'):
            txt = open(os.path.join(REF_SRC, name)).read()
            txt = re.sub(r'AT_DISPATCH_FLOATING_TYPES_AND_HALF\((\w+)\.type\(\)', r'AT_DISPATCH_FLOATING_TYPES_AND_HALF(\1.scalar_type()', txt)
            p = os.path.join(tmp, name)
            open(p, 'w').write(txt)
            srcs.append(p)
        shutil.copy(os.path.join(HERE, 'ref_binding.cpp'), os.path.join(tmp, 'ref_binding.cpp'))
        srcs.append(os.path.join(tmp, 'ref_binding.cpp'))
        inc = ce.include_paths('cuda') if hasattr(ce, 'include_paths') else ce.include_paths(True)
        import sysconfig
        inc = list(inc) + [sysconfig.get_paths()['include']]
        objs = []
        for s in srcs:
            o = s + '.o'
            if s.endswith('.cu'):
                cmd = ['nvcc', '-O3', '-gencode', 'arch=compute_100a,code=sm_100a', '-Xcompiler', '-fPIC', '-std=c++17',
                       '-DTORCH_EXTENSION_NAME=droid_corr_ref', '-D_GLIBCXX_USE_CXX11_ABI=%d' % int(torch._C._GLIBCXX_USE_CXX11_ABI),
                       '--expt-relaxed-constexpr', '-c', s, '-o', o] + ['-I' + i for i in inc]
            else:
                cmd = ['g++', '-O2', '-fPIC', '-std=c++17', '-DTORCH_EXTENSION_NAME=droid_corr_ref',
                       '-D_GLIBCXX_USE_CXX11_ABI=%d' % int(torch._C._GLIBCXX_USE_CXX11_ABI), '-c', s, '-o', o] + ['-I' + i for i in inc]
            subprocess.check_call(cmd)
            objs.append(o)
        libdir = os.path.join(os.path.dirname(torch.__file__), 'lib')
        subprocess.check_call(['g++', '-shared', '-o', OUT] + objs + ['-L' + libdir, '-lc10', '-ltorch_cpu', '-ltorch', '-ltorch_python',
                                                                    '-lc10_cuda', '-ltorch_cuda', '-L/usr/local/cuda/lib64', '-lcudart',
                                                                    '-Wl,-rpath,' + libdir])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return OUT


def patch_droid_kernels(txt):
    """droid_kernels.cu cannot compile as shipped: the vendored Eigen tree has no Eigen/Core.  Three
    documented patches make a PATCHED baseline (never the pristine reference): (1) drop the Eigen includes / typedefs, (2)
    splice our dense-ATen `SparseBlock` (oracle/ref_sparseblock_dense.inc: same interface, float64 Cholesky instead of
    Eigen::SimplicialLLT) in place of the reference's class at :1096-1198, (3) `xi[45]` -> `xi[5]` in expSE3 (:154, an
    out-of-bounds read; upstream DROID-SLAM has xi[5]).  All CUDA kernels (K5-K12) compile untouched."""
    txt = re.sub(r'#include <Eigen/[A-Za-z]+>\n', '', txt)
    txt = re.sub(r'typedef Eigen::SparseMatrix<double> SpMat;\n', '', txt)
    txt = re.sub(r'typedef Eigen::Triplet<double> T;\n', '', txt)
    a = txt.index('class SparseBlock {')
    b = txt.index('SparseBlock schur_block(')
    end = txt.rindex('};', a, b) + 2
    assert 'Eigen' not in txt[:a] + txt[end:]              # nothing else in the file touches Eigen
    txt = txt[:a] + open(os.path.join(HERE, 'ref_sparseblock_dense.inc')).read() + txt[end:]
    txt = txt.replace('xi[45]', 'xi[5]')
    return re.sub(r'\.type\(\)', '.scalar_type()', txt)


def build_ba_if_possible(force=False):
    """oracle/_ref/droid_ba_ref.so: the reference's BA / frame-distance CUDA kernels (patched as documented above)."""
    if not os.path.isdir(REF_SRC):
        return None
    if os.path.exists(OUT_BA) and not force:
        return OUT_BA
    import sysconfig
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix='pvo_ref_ba_')
    try:
        cu = os.path.join(tmp, 'droid_kernels.cu')
        open(cu, 'w').write(patch_droid_kernels(open(os.path.join(REF_SRC, 'droid_kernels.cu')).read()))
        shutil.copy(os.path.join(HERE, 'ref_binding_ba.cpp'), os.path.join(tmp, 'ref_binding_ba.cpp'))
        inc = list(ce.include_paths('cuda') if hasattr(ce, 'include_paths') else ce.include_paths(True)) + [sysconfig.get_paths()['include']]
        abi = '-D_GLIBCXX_USE_CXX11_ABI=%d' % int(torch._C._GLIBCXX_USE_CXX11_ABI)
        subprocess.check_call(['nvcc', '-O3', '-gencode', 'arch=compute_100a,code=sm_100a', '-Xcompiler', '-fPIC', '-std=c++17',
                               '-DTORCH_EXTENSION_NAME=droid_ba_ref', abi, '--expt-relaxed-constexpr', '-c', cu, '-o', cu + '.o'] + ['-I' + i for i in inc])
        bp = os.path.join(tmp, 'ref_binding_ba.cpp')
        subprocess.check_call(['g++', '-O2', '-fPIC', '-std=c++17', '-DTORCH_EXTENSION_NAME=droid_ba_ref', abi, '-c', bp, '-o', bp + '.o'] + ['-I' + i for i in inc])
        libdir = os.path.join(os.path.dirname(torch.__file__), 'lib')
        subprocess.check_call(['g++', '-shared', '-o', OUT_BA, cu + '.o', bp + '.o', '-L' + libdir, '-lc10', '-ltorch_cpu', '-ltorch', '-ltorch_python',
                               '-lc10_cuda', '-ltorch_cuda', '-L/usr/local/cuda/lib64', '-lcudart', '-Wl,-rpath,' + libdir])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return OUT_BA


def load_ba():
    """Import oracle/_ref/droid_ba_ref.so.  Returns None when absent."""
    if not os.path.exists(OUT_BA):
        return None
    import importlib.util
    import torch  # noqa: F401
    spec = importlib.util.spec_from_file_location('droid_ba_ref', OUT_BA)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load():
    """Import oracle/_ref/droid_corr_ref.so (needs torch + a GPU to be useful).  Returns None when absent."""
    if not os.path.exists(OUT):
        return None
    import importlib.util
    import torch  # noqa: F401
    spec = importlib.util.spec_from_file_location('droid_corr_ref', OUT)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == '__main__':
    print(build_if_possible(force='--force' in sys.argv))
    print(build_ba_if_possible(force='--force' in sys.argv))
