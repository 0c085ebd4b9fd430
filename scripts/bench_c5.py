#!/usr/bin/env python
"""BASELINE configs[4] at full size: synthetic 500 genes, T = 10000, full-parents globally coupled chains
(fp_glob_coup_nh_dbn, design width k = 500), SURVEY.md 8(d) recipe (seed 20260120, 5 changepoints, chains started at the
true changepoints).  First 1 chain per node (500 units), then as many chains as make nodes x chains >= 4096 if the first
run says that fits the time budget.  One JSON line per case: unit-iterations/s, executed FP64 rate against the
measured DFMA peak, prefix-table build time and bandwidth.

    python scripts/bench_c5.py [budget_seconds]  [> profiles/<round>/c5.jsonl]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynamicbayesiannetworks_b200.engine import DbnEngine          # noqa: E402
from dynamicbayesiannetworks_b200.synth import make_synthetic     # noqa: E402

N_GENES, T = 500, 10000


def run(data, cps, chains, iters, peak):
    eng = DbnEngine(n_chains=chains, seed=11)
    s = torch.cuda.Stream()
    torch.cuda.set_stream(s)
    eng.set_stream(s.cuda_stream)
    t0 = time.perf_counter()
    eng.build_stats([data])
    eng.sync()
    build_s = time.perf_counter() - t0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.rebuild_stats(); e1.record(); torch.cuda.synchronize()
    prefix_ms = e0.elapsed_time(e1)
    info = eng.stats_info()
    eng.chain_init('fp_glob_coup_nh_dbn')
    eng.set_changepoints([c + 1 for c in cps] + [eng.N + 2])
    eng.run_fp(1)
    torch.cuda.synchronize()
    eng.reset_counters()
    e0.record()
    eng.run_fp(iters)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    c = eng.counters()
    st = eng.state()
    units = eng.n * chains
    out = dict(case='c5 synthetic 500 genes, T=10000, fp-glob-dbn, %d chains' % chains, n_genes=eng.n, N=eng.N, chains=chains,
               units=units, iterations=iters, ms_per_iteration=ms / iters, unit_iterations_per_sec=units * iters / ms * 1e3,
               ml_evals_per_sec=c['ml_evals'] / ms * 1e3, segment_evals=c['segment_evals'],
               algorithmic_gflops=c['algorithmic_flops'] / ms / 1e6, fp64_peak_gflops_measured=peak,
               fp64_frac=c['algorithmic_flops'] / ms / 1e6 / peak, executed_level3_gflops=c['executed_level3_flops'] / ms / 1e6,
               executed_fp64_frac=c['executed_level3_flops'] / ms / 1e6 / peak, algorithmic_gbs=c['algorithmic_bytes'] / ms / 1e6,
               prefix_table_gb=info['algorithmic_bytes'] / 1e9, prefix_build_ms=prefix_ms,
               prefix_gbs=info['algorithmic_bytes'] / prefix_ms / 1e6, build_stats_host_call_s=build_s,
               mean_segments=float(st['S'].mean()), finite_state=bool(np.isfinite(st['lambda2']).all() and (st['sigma2'] > 0).all()))
    eng.close()
    print(json.dumps(out), flush=True)
    return ms / iters


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 240.0
    data, cps, _ = make_synthetic(N_GENES, T, n_changepoints=5, max_fan_in=4, seed=20260120)
    probe = DbnEngine(n_chains=1)
    peak = probe.fp64_peak_gflops()
    probe.close()
    per_it = run(data, cps, 1, 2, peak) / 1e3
    chains = -(-4096 // N_GENES)
    iters = 2
    if per_it * chains * (iters + 1) < budget:
        run(data, cps, chains, iters, peak)
    else:
        print(json.dumps(dict(case='c5 at %d chains skipped' % chains, reason='%.1f s per iteration at 1 chain per node: %d chains x %d '
                              'iterations would exceed the %.0f s budget' % (per_it, chains, iters + 1, budget))), flush=True)


if __name__ == '__main__':
    main()
