#!/usr/bin/env python
"""Does the tail of a window's launch (102 400 units = 2.7 waves of resident blocks at 2 x 128 threads per SM; 1.8 at the earlier 2 x 192) cost throughput?  The same chains as one
handle on one stream, or as K handles with contiguous unit ranges on K streams (their launches overlap: the second wave of one
range runs beside the first wave of another).      python scripts/two_stream_ab.py [K=2]"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynamicbayesiannetworks_b200.engine import DbnEngine          # noqa: E402
from dynamicbayesiannetworks_b200.parallel import shard_units      # noqa: E402
from dynamicbayesiannetworks_b200.synth import make_synthetic     # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 2
METHOD = sys.argv[2] if len(sys.argv) > 2 else 'varying_nh_dbn'
if len(sys.argv) > 3 and sys.argv[3] == 'mtor':      # the mTOR series of the golden fixtures, 4096 chains (bench_methods' device-filled case)
    import glob
    import numpy as np
    g = dict(np.load(glob.glob(os.path.join(ROOT, 'tests', 'golden', 'chain_mtor_varying_nh_dbn.npz'))[0]))
    DATA = [g['data_seg%d' % s_] for s_ in range(int(g['n_data_segments']))]
    n, chains, window, reps, FAN = DATA[0].shape[1], 4096, 50, 8, 3
else:
    n, T, chains, window, reps, FAN = 100, 2000, 1024, 10, 60, 4
    data, cps, _ = make_synthetic(n, T, 3, 4, 20260120)
    DATA = [data]


def run(parts):
    engs, streams = [], []
    for r in range(parts):
        first, count = shard_units(n * chains, r, parts)
        e = DbnEngine(n_chains=chains, seed=11, unit_range=None if parts == 1 else (first, count))
        s = torch.cuda.Stream()
        e.set_stream(s.cuda_stream)
        e.build_stats(DATA)
        e.chain_init(METHOD, fan_in=FAN)
        engs.append(e); streams.append(s)
    for _ in range(3 if window > 10 else 25):
        for e in engs:
            e.run(window)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    ends = [torch.cuda.Event(enable_timing=True) for _ in engs]
    cur = torch.cuda.current_stream()
    e0.record(cur)
    for s in streams:
        s.wait_event(e0)
    for _ in range(reps):
        for e in engs:
            e.run(window)
    for s, ev in zip(streams, ends):
        ev.record(s)
    torch.cuda.synchronize()
    ms = max(e0.elapsed_time(ev) for ev in ends)
    for e in engs:
        e.close()
    return ms / reps


for parts in (1, K, 1, K):
    ms = run(parts)
    print(json.dumps(dict(method=METHOD, units=n * chains, handles=parts, ms_per_step=ms, unit_iterations_per_sec=n * chains * window / ms * 1e3)), flush=True)
