#!/usr/bin/env python
"""Free-running ORACLE chains on the yeast series -> tests/golden/free_running_yeast.npz.

Per method and node: the edge-frequency row (proposed_adj_matrix row, network.py:355-368 / scores.py:155-194) of each
of N_CHAINS independent oracle chains (oracle/dbn_oracle.py:run_chain with numpy draws).  The GPU test
tests/test_gpu_pinning.py::test_free_running_edge_posteriors_within_mc_error compares the device's Philox-driven chains
with the mean of these rows, with the Monte-Carlo error taken from their spread.  The oracle itself is
pinned to the reference by tests/test_oracle_golden.py; the chains are generated here (minutes of CPU) so that the GPU
box does not spend its time on them.

    python tests/golden/make_free_running.py        # ~10 min on 8 cores
"""
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

METHODS = ['h_dbn', 'varying_nh_dbn', 'seq_coup_nh_dbn', 'glob_coup_nh_dbn', 'var_glob_coup_nh_dbn']
N_CHAINS, N_ITER, BURN = 64, 800, 200


def work(task):
    from helpers import CHAIN_FILES, load, data_list
    from oracle import dbn_oracle as O
    method, node, cix = task
    data = data_list(load([p for p in CHAIN_FILES if 'yeast_varying' in p][0]))
    X, Y = O.lagged_matrices(data)
    N, n = X.shape
    rng = np.random.default_rng([METHODS.index(method), node, cix, 20261020])
    edge = np.zeros((n, n)); nonempty = np.zeros(n); cph = np.zeros((n, N + 3)); cpk = np.zeros(n)
    t = O.run_chain(method, X, Y, node, N_ITER, O.RngDraws(rng))
    O.accumulate_counts(t, node, n, N, BURN, edge, nonempty, cph, cpk)
    return method, node, cix, edge[node] / max(nonempty[node], 1), cph[node] / max(cpk[node], 1)


def main():
    n = 5
    tasks = [(m, node, c) for m in METHODS for node in range(n) for c in range(N_CHAINS)]
    out = {m: np.zeros((N_CHAINS, n, n)) for m in METHODS}
    cps = {}
    with mp.get_context('spawn').Pool(os.cpu_count()) as pool:
        for m, node, c, row, cp in pool.imap_unordered(work, tasks, chunksize=4):
            out[m][c, node] = row
            cps.setdefault(m, np.zeros((N_CHAINS, n, cp.size)))[c, node] = cp
    save = {'edge_freq_' + m: out[m] for m in METHODS}
    save.update({'cp_freq_' + m: cps[m] for m in METHODS})
    save.update(n_chains=N_CHAINS, n_iter=N_ITER, burn_in=BURN)
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'free_running_yeast.npz'), **save)
    for m in METHODS:
        print(m, np.round(out[m].mean(axis=0), 3).tolist())


if __name__ == '__main__':
    main()
