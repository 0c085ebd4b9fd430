#!/usr/bin/env python
"""Units-per-warp A/B at the per-GPU load of the c4 strong-scaling run (100 genes, T = 2000, nh-dbn): DBN_LANES=32|16|8|.. or auto.
    DBN_LANES=8 python scripts/lanes_c4.py [chains=128]"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynamicbayesiannetworks_b200.engine import DbnEngine          # noqa: E402
from dynamicbayesiannetworks_b200.synth import make_synthetic     # noqa: E402
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 128
data, cps, _ = make_synthetic(100, 2000, 3, 4, 20260120)
eng = DbnEngine(n_chains=chains, seed=11)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); eng.set_stream(s.cuda_stream)
eng.build_stats([data]); eng.chain_init('varying_nh_dbn', fan_in=4)
for _ in range(30): eng.run(10)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): eng.run(10)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(json.dumps(dict(lanes=os.environ.get('DBN_LANES', 'auto'), units=100 * chains, ms_per_step=ms / 50, unit_iterations_per_sec=100 * chains * 500 / ms * 1e3)))
