"""Build the CUDA library in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m dynamicbayesiannetworks_b200.build [--force] [--verbose] [--slim] [-DNAME[=VALUE] ...] [--out path.so]

Produces dynamicbayesiannetworks_b200/libdbn_b200.so, which ctypes loads (see _lib.py) and which travels to the
GPU box with the repo snapshot.  The kernel families are separate objects (csrc/dbn_inst.cu compiled once per
family / trace flag), built in parallel and cached under build/obj by source time stamp and flags.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, 'csrc')
LIB = os.path.join(PKG, 'libdbn_b200.so')
OBJDIR = os.path.join(ROOT, 'build', 'obj')
_H = lambda *names: [os.path.join(CSRC, f) for f in names] + [os.path.join(ROOT, 'include', 'dbn_b200.h')]
COMMON = _H('dbn_score.cuh', 'dbn_rng.cuh', 'dbn_chain.cuh', 'dbn_launch.h')
# object name -> (source, defines, headers it depends on beside COMMON)
OBJECTS = {
    'abi': ('dbn_kernels.cu', [], ['dbn_chain_fast.cuh', 'dbn_fp.cuh', 'dbn_fp_big.cuh', 'dbn_coupled.cuh', 'dbn_warp.cuh']),
    'fast_plain': ('dbn_inst.cu', ['DBN_INST_FAMILY=1', 'DBN_INST_TRACE=0'], ['dbn_chain_fast.cuh']),
    'fast_trace': ('dbn_inst.cu', ['DBN_INST_FAMILY=1', 'DBN_INST_TRACE=1'], ['dbn_chain_fast.cuh']),
    'chain_plain': ('dbn_inst.cu', ['DBN_INST_FAMILY=2', 'DBN_INST_TRACE=0'], ['dbn_coupled.cuh']),
    'chain_trace': ('dbn_inst.cu', ['DBN_INST_FAMILY=2', 'DBN_INST_TRACE=1'], ['dbn_coupled.cuh']),
    'fp_plain_small': ('dbn_inst.cu', ['DBN_INST_FAMILY=3', 'DBN_INST_TRACE=0', 'DBN_INST_BIG=0'], ['dbn_fp.cuh', 'dbn_fp_big.cuh']),
    'fp_plain_big': ('dbn_inst.cu', ['DBN_INST_FAMILY=3', 'DBN_INST_TRACE=0', 'DBN_INST_BIG=1'], ['dbn_fp.cuh', 'dbn_fp_big.cuh']),
    'fp_trace_small': ('dbn_inst.cu', ['DBN_INST_FAMILY=3', 'DBN_INST_TRACE=1', 'DBN_INST_BIG=0'], ['dbn_fp.cuh', 'dbn_fp_big.cuh']),
    'fp_trace_big': ('dbn_inst.cu', ['DBN_INST_FAMILY=3', 'DBN_INST_TRACE=1', 'DBN_INST_BIG=1'], ['dbn_fp.cuh', 'dbn_fp_big.cuh']),
    'stubs': ('dbn_stub_launchers.cu', [], []),   # --slim only
}
SLIM = ('abi', 'fast_plain', 'fast_trace', 'stubs')   # kernel A/B builds of the h / nh family (build_variants/*.so)


def nvcc_path():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return 'nvcc'


def _deps(name):
    src, _, extra = OBJECTS[name]
    return [os.path.join(CSRC, src)] + COMMON + [p for p in (os.path.join(CSRC, f) for f in extra) if os.path.exists(p)]


def _obj_path(name, defines):
    tag = hashlib.sha1(' '.join(sorted(defines)).encode()).hexdigest()[:10]
    return os.path.join(OBJDIR, '%s_%s.o' % (name, tag))


def _stale(obj, deps):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(d) > t for d in deps)


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for name in OBJECTS if name != 'stubs' for d in _deps(name))


def _compile(name, defines, force, verbose):
    src, own, _ = OBJECTS[name]
    alld = list(own) + list(defines)
    obj = _obj_path(name, alld)
    if not force and not _stale(obj, _deps(name)):
        return obj, ''
    cmd = [nvcc_path(), '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-c',
           '-Xcompiler', '-fPIC', '-Xcompiler', '-O2', '-I', os.path.join(ROOT, 'include'), '-I', CSRC,
           '-o', obj] + ['-D' + d for d in alld] + [os.path.join(CSRC, src)]
    if verbose:
        cmd[1:1] = ['-Xptxas', '-v']
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n%s\n%s' % (' '.join(cmd), res.stdout))
    return obj, res.stdout


def build(force=False, verbose=False, out=None, defines=(), jobs=None, slim=False):
    """`out` / `defines`: build a variant (e.g. a different launch bound) beside the default library."""
    if out is None and not force and not defines and not is_stale():
        return LIB
    out = out or LIB
    os.makedirs(OBJDIR, exist_ok=True)
    jobs = jobs or min(len(OBJECTS), os.cpu_count() or 1)
    names = [n for n in OBJECTS if (n in SLIM if slim else n != 'stubs')]
    with ThreadPoolExecutor(jobs) as ex:
        results = list(ex.map(lambda n: _compile(n, defines, force, verbose), names))
    if verbose:
        for name, (_, log) in zip(names, results):
            if log:
                print('---- %s\n%s' % (name, log))
    cmd = [nvcc_path(), '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', out] + [o for o, _ in results]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError('link failed:\n%s\n%s' % (' '.join(cmd), res.stdout))
    return out


if __name__ == '__main__':
    argv = sys.argv[1:]
    out = argv[argv.index('--out') + 1] if '--out' in argv else None
    defines = [a[2:] for a in argv if a.startswith('-D')]
    print(build(force='--force' in argv, verbose='--verbose' in argv, out=out, defines=defines, slim='--slim' in argv))
