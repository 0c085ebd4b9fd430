"""How far apart are the segment counts S of the units that share a warp / a block?  (bench workload, configs[3])
A warp's segment loops run max(S) times over its lanes, and the block barriers of the h / nh kernel tie a block's warps together,
so mean(max S) / mean(S) bounds what sorting the units of a node by S could save."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from dynamicbayesiannetworks_b200.engine import DbnEngine
from dynamicbayesiannetworks_b200.synth import make_synthetic
data, _, _ = make_synthetic(100, 2000, 3, 4, 20260119)
e = DbnEngine(n_chains=1024, seed=0xDB20B200)
e.build_stats([data]); e.chain_init('varying_nh_dbn', fan_in=4, burn_in=0)
for it in (300, 700):
    e.run(it)
    st = e.state()
    S = st['S'].astype(np.int64); k = st['npi'].astype(np.int64)
    print('after', it, 'more iterations: S mean %.2f sd %.2f min %d max %d | npi mean %.2f' % (S.mean(), S.std(), S.min(), S.max(), k.mean()))
    for g in (32, 128):
        print('  groups of %d: mean of max S %.2f (unsorted)' % (g, S.reshape(-1, g).max(1).mean()), end='')
        Ss = np.sort(S.reshape(100, 1024), axis=1).reshape(-1, g)
        print(' | sorted within node %.2f' % Ss.max(1).mean())
    print('  per node mean S: min %.2f max %.2f' % (S.reshape(100, 1024).mean(1).min(), S.reshape(100, 1024).mean(1).max()))
