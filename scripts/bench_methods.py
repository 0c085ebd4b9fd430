#!/usr/bin/env python
"""Device throughput of every method on the BASELINE configs that are not the bench.py line (parity-test cases):
   c2 yeast nh-dbn 64 chains; c3 mTOR seq-dbn / glob-dbn 256 chains; h-dbn and fixed nh on yeast; fp-glob on yeast /
   mTOR / a 40-gene synthetic series.  Data come from the golden fixtures (the series the reference ships) so that the
   script runs on the GPU box.  One JSON line per case: unit-iterations/s with the table resident (CUDA events).

       python scripts/bench_methods.py [label-substring] [> profiles/r1_v9/methods.jsonl]
"""
import glob
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynamicbayesiannetworks_b200.engine import DbnEngine          # noqa: E402
from dynamicbayesiannetworks_b200.synth import make_synthetic     # noqa: E402


def series(name):
    g = dict(np.load(glob.glob(os.path.join(ROOT, 'tests', 'golden', 'chain_%s_varying_nh_dbn.npz' % name))[0]))
    return [g['data_seg%d' % s] for s in range(int(g['n_data_segments']))]


def run(label, data, method, chains, window, reps, handles=1, **kw):
    """handles > 1: the units are split into contiguous ranges, one library handle and stream each (their launches overlap)."""
    from dynamicbayesiannetworks_b200.parallel import shard_units
    n = data[0].shape[1]
    engs = []
    for h in range(handles):
        rg = None if handles == 1 else shard_units(n * chains, h, handles)
        eng = DbnEngine(n_chains=chains, seed=11, unit_range=rg)
        s = torch.cuda.Stream()
        eng.set_stream(s.cuda_stream)
        eng.build_stats(data)
        eng.chain_init(method, **kw)
        if method == 'fixed_nh_dbn':
            N = eng.N
            eng.set_changepoints([N // 3, 2 * N // 3, N + 2])
        engs.append((eng, s))
    steps = [(e.run_fp if method.startswith('fp_') else e.run) for e, _ in engs]
    for _ in range(3):
        for st in steps:
            st(window)
    torch.cuda.synchronize()
    for e, _ in engs:
        e.reset_counters()
    e0 = torch.cuda.Event(enable_timing=True)
    ends = [torch.cuda.Event(enable_timing=True) for _ in engs]
    e0.record()
    for _, s in engs:
        s.wait_event(e0)
    for _ in range(reps):
        for st in steps:
            st(window)
    for (_, s), ev in zip(engs, ends):
        ev.record(s)
    torch.cuda.synchronize()
    ms = max(e0.elapsed_time(ev) for ev in ends)
    c = {}
    for e, _ in engs:
        for k, v in e.counters().items():
            c[k] = c.get(k, 0) + v
    eng = engs[0][0]
    units = eng.n * chains
    out = dict(case=label, method=method, n_genes=eng.n, N=eng.N, chains=chains, units=units, handles=handles, iterations=window * reps,
               ms_per_iteration=ms / (window * reps), unit_iterations_per_sec=units * window * reps / ms * 1e3,
               ml_evals_per_sec=c['ml_evals'] / ms * 1e3, algorithmic_gbs=c['algorithmic_bytes'] / ms / 1e6,
               algorithmic_gflops=c['algorithmic_flops'] / ms / 1e6)
    for e, _ in engs:
        e.close()
    print(json.dumps(out), flush=True)


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else ''   # substring filter on the case label
    global run
    run_all = run

    def run(label, *a, **k):
        if only in label:
            run_all(label, *a, **k)
    yeast, mtor = series('yeast'), series('mtor')
    syn40, _, _ = make_synthetic(40, 400, 3, 4, 5)
    run('c2 yeast nh-dbn, 64 chains', yeast, 'varying_nh_dbn', 64, 100, 20)
    run('yeast h-dbn, 64 chains', yeast, 'h_dbn', 64, 100, 20)
    run('yeast fixed nh-dbn, 64 chains', yeast, 'fixed_nh_dbn', 64, 100, 20)
    run('c3 mTOR seq-dbn, 256 chains', mtor, 'seq_coup_nh_dbn', 256, 50, 10)
    run('c3 mTOR glob-dbn, 256 chains', mtor, 'glob_coup_nh_dbn', 256, 50, 10)
    run('mTOR var-glob-dbn, 256 chains', mtor, 'var_glob_coup_nh_dbn', 256, 50, 10)
    run('mTOR nh-dbn, 4096 chains (device filled)', mtor, 'varying_nh_dbn', 4096, 50, 10)
    run('mTOR seq-dbn, 4096 chains (device filled)', mtor, 'seq_coup_nh_dbn', 4096, 50, 10)
    run('mTOR glob-dbn, 4096 chains (device filled)', mtor, 'glob_coup_nh_dbn', 4096, 50, 10)
    run('mTOR var-glob-dbn, 4096 chains (device filled)', mtor, 'var_glob_coup_nh_dbn', 4096, 50, 10)
    run('mTOR nh-dbn, 4096 chains, two handles', mtor, 'varying_nh_dbn', 4096, 50, 10, handles=2)
    run('mTOR seq-dbn, 4096 chains, two handles', mtor, 'seq_coup_nh_dbn', 4096, 50, 10, handles=2)
    run('mTOR glob-dbn, 4096 chains, two handles', mtor, 'glob_coup_nh_dbn', 4096, 50, 10, handles=2)
    run('mTOR var-glob-dbn, 4096 chains, two handles', mtor, 'var_glob_coup_nh_dbn', 4096, 50, 10, handles=2)
    run('yeast fp-glob-dbn, 64 chains', yeast, 'fp_glob_coup_nh_dbn', 64, 20, 5)
    run('mTOR fp-glob-dbn, 256 chains', mtor, 'fp_glob_coup_nh_dbn', 256, 10, 3)
    run('synthetic 40 genes T=400 fp-glob-dbn, 64 chains', [syn40], 'fp_glob_coup_nh_dbn', 64, 5, 2)
    run('mTOR fp-nh-dbn, 256 chains', mtor, 'fp_varying_nh_dbn', 256, 10, 3)
    run('mTOR fp-h-dbn, 256 chains', mtor, 'fp_h_dbn', 256, 10, 3)
    run('mTOR fp-var-glob-dbn, 256 chains', mtor, 'fp_var_glob_coup_nh_dbn', 256, 10, 3)
    run('mTOR fp-seq-dbn, 256 chains', mtor, 'fp_seq_coup_nh_dbn', 256, 10, 3)


if __name__ == '__main__':
    main()
