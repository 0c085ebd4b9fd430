import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
from helpers import CHAIN_FILES, load, data_list
from dynamicbayesiannetworks_b200.engine import DbnEngine
np.set_printoptions(linewidth=250, precision=6)
g = load([p for p in CHAIN_FILES if 'yeast_varying' in p][0])
NIT = int(sys.argv[1]) if len(sys.argv) > 1 else 50
C = 256
states = []
with DbnEngine(n_chains=C, seed=1234) as e:
    e.build_stats(data_list(g))
    e.chain_init('varying_nh_dbn', burn_in=500)
    for it in range(NIT):
        e.run(1)
        states.append(e.state())
with DbnEngine(n_chains=C, seed=1234) as e:
    e.build_stats(data_list(g))
    e.chain_init('varying_nh_dbn', burn_in=500)
    t = e.run(NIT, trace=True)
U = t['S'].shape[0]
mv = {0: 'B', 1: 'D', 2: 'R'}
pm = {0: 'add', 1: 'del', 2: 'exc'}
first = {}
for it in range(NIT):
    st = states[it]
    for u in range(U):
        if u in first: continue
        S = int(t['S_after'][u, it]); npi = int(t['npi'][u, it])
        same = (st['S'][u] == S and list(st['tau'][u][:S]) == list(t['tau_after'][u, it][:S]) and st['npi'][u] == npi
                and list(st['pi'][u][:npi]) == list(t['pi_after'][u, it][:npi])
                and abs(st['sigma2'][u] - t['sigma2'][u, it]) <= 1e-9 * abs(t['sigma2'][u, it]) and abs(st['lambda2'][u] - t['lambda2'][u, it]) <= 1e-9 * abs(t['lambda2'][u, it]))
        if not same:
            first[u] = it
print('units diverging:', len(first))
for u, it in sorted(first.items(), key=lambda kv: kv[1])[:12]:
    st = states[it]
    print('unit', u, 'warp', u // 32, 'lane', u % 32, 'first divergence at iteration', it)
    print('   plain: S %d tau %s npi %d pi %s sig %.10g lam %.10g' % (st['S'][u], list(st['tau'][u][:st['S'][u]]), st['npi'][u], list(st['pi'][u][:st['npi'][u]]), st['sigma2'][u], st['lambda2'][u]))
    for j in range(max(0, it - 2), it + 1):
        print('   trace it %d: S %d->%d tau_after %s | t:%s v%d a%d | p:%s v%d a%d npi %d pi %s | sig %.10g lam %.10g logml %s' % (
            j, t['S'][u, j], t['S_after'][u, j], list(map(int, t['tau_after'][u, j][:t['S_after'][u, j]])),
            mv[int(t['tau_move'][u, j])], t['tau_valid'][u, j], t['tau_accept'][u, j],
            pm[int(t['pi_move'][u, j])], t['pi_valid'][u, j], t['pi_accept'][u, j], t['npi'][u, j], list(map(int, t['pi_after'][u, j][:t['npi'][u, j]])),
            t['sigma2'][u, j], t['lambda2'][u, j], t['logml'][u, j]))
# what the warp of the earliest divergence did in the iteration before
if first:
    u0, it0 = sorted(first.items(), key=lambda kv: kv[1])[0]
    w = u0 // 32
    for j in range(max(0, it0 - 1), it0 + 1):
        print('=== warp', w, 'iteration', j)
        for u in range(32 * w, 32 * w + 32):
            print('lane %2d S %d->%d tau_after %s | t:%s v%d a%d | p:%s v%d a%d npi %d pi %s' % (
                u % 32, t['S'][u, j], t['S_after'][u, j], list(map(int, t['tau_after'][u, j][:t['S_after'][u, j]])),
                mv[int(t['tau_move'][u, j])], t['tau_valid'][u, j], t['tau_accept'][u, j],
                pm[int(t['pi_move'][u, j])], t['pi_valid'][u, j], t['pi_accept'][u, j], t['npi'][u, j], list(map(int, t['pi_after'][u, j][:t['npi'][u, j]]))))
