#!/usr/bin/env python
"""Throughput of kernel 2, the batched scorer (dbn_score_batch): B random (node, parent set, changepoint list) items on
the c4 series (100 genes, T = 2000), device time of the scorer kernel alone (dbn_score_batch_time), as marginal-likelihood
evaluations/s, segment evaluations/s, algorithmic GB/s and FP64 GFLOP/s (SURVEY 8d counts of the executed evaluations).

    python scripts/bench_scorer.py [B=262144] [> profiles/r2_scorer/scorer.jsonl]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynamicbayesiannetworks_b200.engine import DbnEngine          # noqa: E402
from dynamicbayesiannetworks_b200.synth import make_synthetic     # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
    n, T = 100, 2000
    data, _, _ = make_synthetic(n, T, 3, 4, 20260120)
    rng = np.random.default_rng(7)
    eng = DbnEngine(n_chains=1)
    eng.build_stats([data])
    peak = eng.fp64_peak_gflops()
    N = eng.N
    for label, fan_in, n_cp, seq in (('fan-in 4, 8 segments, shared prior mean (glob form)', 4, 7, False),
                                     ('fan-in 4, 8 segments, sequentially coupled', 4, 7, True),
                                     ('fan-in 3, 3 segments, shared prior mean', 3, 2, False)):
        node = rng.integers(0, n, B).astype(np.int32)
        c = np.argsort(rng.random((B, n - 1)), axis=1)[:, :fan_in]          # fan_in distinct genes out of the n - 1 others
        c = np.sort(c + (c >= node[:, None]), axis=1)
        cps = np.sort(rng.integers(2, N - n_cp, (B, n_cp)), axis=1) + np.arange(n_cp)   # strictly increasing changepoints
        pi = c.tolist()
        tau = np.concatenate([cps, np.full((B, 1), N + 2)], axis=1).tolist()
        lam = rng.gamma(2.0, 0.5, B) + 0.1
        dlt = rng.gamma(2.0, 0.5, B) + 0.1
        mu = rng.normal(0, 0.3, (B, 5))
        eng.reset_counters()
        out = eng.score_batch(node, pi, tau, lam, delta2=dlt, mu=mu, seq=seq)
        ms = eng.score_batch_ms()
        eng.reset_counters()
        out = eng.score_batch(node, pi, tau, lam, delta2=dlt, mu=mu, seq=seq)   # second call: scratch allocated, caches warm
        ms = eng.score_batch_ms()
        c = eng.counters()
        print(json.dumps(dict(case='scorer, c4 series, ' + label, B=B, kernel_ms=ms, ml_evals_per_sec=B / ms * 1e3,
                              segment_evals_per_sec=c['segment_evals'] / ms * 1e3, algorithmic_gbs=c['algorithmic_bytes'] / ms / 1e6,
                              algorithmic_gflops=c['algorithmic_flops'] / ms / 1e6, fp64_peak_gflops_measured=peak,
                              fp64_frac=c['algorithmic_flops'] / ms / 1e6 / peak, finite=bool(np.isfinite(out).all()))), flush=True)
    eng.close()


if __name__ == '__main__':
    main()
