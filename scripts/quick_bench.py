"""Quick device timing of the c4 workload (development aid; bench.py is the contract)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, '.')
from dynamicbayesiannetworks_b200.engine import DbnEngine
from dynamicbayesiannetworks_b200.synth import make_synthetic

n, T, C = 100, 2000, int(sys.argv[1]) if len(sys.argv) > 1 else 1024
method = sys.argv[2] if len(sys.argv) > 2 else 'varying_nh_dbn'
fan = 4
data, _, _ = make_synthetic(n, T, 3, 4, 20260119)
e = DbnEngine(n_chains=C)
s = torch.cuda.Stream()
torch.cuda.set_stream(s)
e.set_stream(s.cuda_stream)
t0 = time.time(); e.build_stats([data]); print('build_stats wall %.3f s' % (time.time() - t0), e.stats_info())
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    ev0.record(); e.rebuild_stats(); ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1); info = e.stats_info()
    print('prefix rebuild %.3f ms  -> %.1f GB/s' % (ms, info['algorithmic_bytes'] / ms / 1e6))
print('fp64 peak GFLOP/s', e.fp64_peak_gflops())
e.chain_init(method, fan_in=fan, burn_in=0)
W = 10
for rep in range(30):
    e.reset_counters()
    ev0.record(); e.run(W); ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    c = e.counters()
    if rep % 3 == 0 or rep > 26:
        st = e.state()
        print('win %2d: %.3f ms/window  %.1f us/sweep  %.3e unit-it/s  evals/it %.2f  segs/it %.1f  alg GB/s %.1f  GF/s %.1f  S mean %.2f  npi mean %.2f' % (
            rep, ms, ms * 1e3 / W, n * C * W / ms * 1e3, c['ml_evals'] / c['unit_iterations'], c['segment_evals'] / c['unit_iterations'],
            c['algorithmic_bytes'] / ms / 1e6, c['algorithmic_flops'] / ms / 1e6, st['S'].mean(), st['npi'].mean()))
