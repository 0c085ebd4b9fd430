"""Build the CUDA library in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m dynamicbayesiannetworks_b200.build [--force] [--verbose]

Produces dynamicbayesiannetworks_b200/libdbn_b200.so, which ctypes loads (see _lib.py) and which travels to the
GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, 'csrc')
LIB = os.path.join(PKG, 'libdbn_b200.so')
SOURCES = [os.path.join(CSRC, 'dbn_kernels.cu')]
DEPS = SOURCES + [os.path.join(CSRC, f) for f in ('dbn_chain.cuh', 'dbn_chain_fast.cuh', 'dbn_fp.cuh', 'dbn_score.cuh', 'dbn_rng.cuh')] + [
    os.path.join(ROOT, 'include', 'dbn_b200.h')]


def nvcc_path():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return 'nvcc'


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False, out=None, defines=()):
    """`out` / `defines`: build a variant (e.g. a different launch bound) beside the default library."""
    if out is None and not force and not is_stale():
        return LIB
    out = out or LIB
    cmd = [nvcc_path(), '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
           '-shared', '-Xcompiler', '-fPIC', '-Xcompiler', '-O2', '-I', os.path.join(ROOT, 'include'),
           '-o', out] + ['-D' + d for d in defines] + SOURCES
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n%s\n%s' % (' '.join(cmd), res.stdout))
    if verbose:
        print(res.stdout)
    return out


if __name__ == '__main__':
    argv = sys.argv[1:]
    out = argv[argv.index('--out') + 1] if '--out' in argv else None
    defines = [a[2:] for a in argv if a.startswith('-D')]
    print(build(force='--force' in argv, verbose='--verbose' in argv, out=out, defines=defines))
