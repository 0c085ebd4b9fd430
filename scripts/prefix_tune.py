import sys, os
sys.path.insert(0, '.')
import torch
from dynamicbayesiannetworks_b200.engine import DbnEngine
from dynamicbayesiannetworks_b200.synth import make_synthetic
data, _, _ = make_synthetic(100, 2000, 3, 4, 20260119)
e = DbnEngine(n_chains=8)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); e.set_stream(s.cuda_stream)
e.build_stats([data])
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for _ in range(20):
    ev0.record(); e.rebuild_stats(); ev1.record(); torch.cuda.synchronize()
    best = min(best, ev0.elapsed_time(ev1))
info = e.stats_info()
print('Lc=%s prefix rebuild best %.4f ms -> %.1f GB/s algorithmic' % (os.environ.get('DBN_PREFIX_LC'), best, info['algorithmic_bytes'] / best / 1e6))
