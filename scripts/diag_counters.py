"""Development aid: counters of the plain build of the h / nh kernel on the yeast series (see
tests/test_gpu_pinning.py::test_work_counters_match_host_recount)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from helpers import CHAIN_FILES, load, data_list
from dynamicbayesiannetworks_b200.engine import DbnEngine
data = data_list(load([p for p in CHAIN_FILES if 'yeast_varying' in p][0]))
for method in ('varying_nh_dbn', 'h_dbn'):
    for C, n_iter, seed, burn in ((64, 20, 5, 0), (64, 200, 5, 0), (256, 200, 5, 0), (256, 1500, 5, 0), (256, 1500, 1234, 500)):
        with DbnEngine(n_chains=C, seed=seed) as e:
            e.build_stats(data)
            e.chain_init(method, burn_in=burn)
            e.run(n_iter)
            e.sync()
            c = e.counters()
            st = e.state()
            print(method, C, n_iter, seed, burn, 'OK' if c['unit_iterations'] == 5 * C * n_iter else 'BAD', ['%x' % v for v in c.values()],
                  'verify', e.verify_cache(), 'npi', np.bincount(st['npi'], minlength=5).tolist(), 'finite', bool(np.isfinite(st['sigma2']).all()), flush=True)
