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


def run(label, data, method, chains, window, reps, **kw):
    eng = DbnEngine(n_chains=chains, seed=11)
    s = torch.cuda.Stream()
    torch.cuda.set_stream(s)
    eng.set_stream(s.cuda_stream)
    eng.build_stats(data)
    eng.chain_init(method, **kw)
    if method == 'fixed_nh_dbn':
        N = eng.N
        eng.set_changepoints([N // 3, 2 * N // 3, N + 2])
    step = eng.run_fp if method.startswith('fp_') else eng.run
    for _ in range(3):
        step(window)
    torch.cuda.synchronize()
    eng.reset_counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step(window)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    c = eng.counters()
    units = eng.n * chains
    out = dict(case=label, method=method, n_genes=eng.n, N=eng.N, chains=chains, units=units, iterations=window * reps,
               ms_per_iteration=ms / (window * reps), unit_iterations_per_sec=units * window * reps / ms * 1e3,
               ml_evals_per_sec=c['ml_evals'] / ms * 1e3, algorithmic_gbs=c['algorithmic_bytes'] / ms / 1e6,
               algorithmic_gflops=c['algorithmic_flops'] / ms / 1e6)
    eng.close()
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
    run('yeast fp-glob-dbn, 64 chains', yeast, 'fp_glob_coup_nh_dbn', 64, 20, 5)
    run('mTOR fp-glob-dbn, 256 chains', mtor, 'fp_glob_coup_nh_dbn', 256, 10, 3)
    run('synthetic 40 genes T=400 fp-glob-dbn, 64 chains', [syn40], 'fp_glob_coup_nh_dbn', 64, 5, 2)
    run('mTOR fp-nh-dbn, 256 chains', mtor, 'fp_varying_nh_dbn', 256, 10, 3)
    run('mTOR fp-h-dbn, 256 chains', mtor, 'fp_h_dbn', 256, 10, 3)
    run('mTOR fp-var-glob-dbn, 256 chains', mtor, 'fp_var_glob_coup_nh_dbn', 256, 10, 3)
    run('mTOR fp-seq-dbn, 256 chains', mtor, 'fp_seq_coup_nh_dbn', 256, 10, 3)


if __name__ == '__main__':
    main()
