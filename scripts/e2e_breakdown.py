"""Where the end-to-end step of bench.py spends its time (host wall clock, synchronised after each call)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, '.')
from dynamicbayesiannetworks_b200.engine import DbnEngine
from dynamicbayesiannetworks_b200.synth import make_synthetic
data, _, _ = make_synthetic(100, 2000, 3, 4, 20260119)
data = torch.from_numpy(data).pin_memory().numpy()
e = DbnEngine(n_chains=1024)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); e.set_stream(s.cuda_stream)
e.build_stats([data]); e.chain_init('varying_nh_dbn', fan_in=4, burn_in=0)
for _ in range(30): e.run(10)
torch.cuda.synchronize()
host = torch.empty(e.counts_device_ptr()[1], dtype=torch.int64, pin_memory=True).numpy()
acc = dict(build=0.0, run_refresh=0.0, run_plain=0.0, counts=0.0)
K = 30
for _ in range(K):
    t0 = time.perf_counter(); e.build_stats([data]); torch.cuda.synchronize(); t1 = time.perf_counter()
    e.run(10); torch.cuda.synchronize(); t2 = time.perf_counter()
    c = e.counts(out=host); t3 = time.perf_counter()
    e.run(10); torch.cuda.synchronize(); t4 = time.perf_counter()
    acc['build'] += t1 - t0; acc['run_refresh'] += t2 - t1; acc['counts'] += t3 - t2; acc['run_plain'] += t4 - t3
print({k: '%.3f ms' % (v / K * 1e3) for k, v in acc.items()})
