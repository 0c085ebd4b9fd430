#!/usr/bin/env python
"""One launch of the full-parents kernel at BASELINE configs[4] size (500 genes, T = 10000, 1 chain per node) for ncu:
    ncu --set full --import-source on -k regex:fp_glob --launch-skip 1 -c 1 python scripts/c5_one.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynamicbayesiannetworks_b200.engine import DbnEngine          # noqa: E402
from dynamicbayesiannetworks_b200.synth import make_synthetic     # noqa: E402

n, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (500, 10000)
data, cps, _ = make_synthetic(n, T, n_changepoints=5, max_fan_in=4, seed=20260120)
eng = DbnEngine(n_chains=1, seed=11)
eng.build_stats([data])
eng.chain_init('fp_glob_coup_nh_dbn')
eng.set_changepoints([c + 1 for c in cps] + [eng.N + 2])
eng.run_fp(1)
eng.run_fp(1)
eng.sync()
print(eng.counters())
