"""Development aid: device time of the prefix build (kernel 1) on BASELINE configs[3] (CUDA events, best and median of 30),
for the default three-launch build and, with DBN_PREFIX_ONEPASS=1, the one-launch build (decoupled look-back)."""
import sys, os
sys.path.insert(0, '.')
import numpy as np
import torch
from dynamicbayesiannetworks_b200.engine import DbnEngine
from dynamicbayesiannetworks_b200.synth import make_synthetic
n, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (100, 2000)
data, _, _ = make_synthetic(n, T, 3, 4, 20260119)
e = DbnEngine(n_chains=8)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); e.set_stream(s.cuda_stream)
e.build_stats([data])
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(30):
    ev0.record(); e.rebuild_stats(); ev1.record(); torch.cuda.synchronize()
    ts.append(ev0.elapsed_time(ev1))
info = e.stats_info()
zz, zy, yy = e.stats_row(info['N'])
X = data[:-1]; Z = np.concatenate([np.ones((X.shape[0], 1)), X], axis=1)
ok = np.allclose(zz, Z.T @ Z, rtol=1e-11, atol=1e-9) and np.allclose(yy, (data[1:] ** 2).sum(axis=0), rtol=1e-11)
print('n=%d T=%d onepass=%s Lc=%s: prefix rebuild best %.4f ms median %.4f ms -> %.1f / %.1f GB/s algorithmic (%.1f MB); last row ok: %s' % (
    n, T, os.environ.get('DBN_PREFIX_ONEPASS', '0'), os.environ.get('DBN_PREFIX_LC'), min(ts), float(np.median(ts)),
    info['algorithmic_bytes'] / min(ts) / 1e6, info['algorithmic_bytes'] / float(np.median(ts)) / 1e6, info['algorithmic_bytes'] / 1e6, ok))
