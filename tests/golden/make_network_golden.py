#!/usr/bin/env python
"""Network-level fixtures from the UNMODIFIED reference -> tests/golden/network_<data>_<method>.npz.

Run in the build container only (needs /root/reference):  python tests/golden/make_network_golden.py

What `Network.score_edges` (network.py:283-396) derives on the host from a finished chain: `padded_betas`
(`transform_beta_coef`, bayesianLinearRegression.py:40-67), the thinned changepoint chain `cps_over_response`, the
per-time-point beta matrices `betas_over_time` (scores.py:125-137), `scores_over_time` (scores.py:106-123, full-parents
methods) and the rows of `proposed_adj_matrix`.  The chain itself (pi / tau / beta histories) is stored as the INPUT, so
the CPU test `tests/test_host_logic.py::test_network_diagnostics_vs_reference` feeds the reference's own chain to this
repo's `scores.py` / `Network._diagnostics` and compares the outputs exactly.  Also the lag-2 design dict of
`set_network_configuration` (network.py:92-122).
"""
import contextlib
import io
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refboot  # noqa: E402

dyban = _refboot.import_reference()
from dyban import Network  # noqa: E402
import dyban.scores as ref_scores  # noqa: E402

NAN = float('nan')
SM = 10


def _pad_list_of_vecs(lst, width):
    out = np.full((SM, width), NAN)
    for h, v in enumerate(lst):
        v = np.asarray(v, dtype=float).ravel()
        out[h, :v.size] = v
    return out


def run(name, data, method, n_iter, burn_in, seed):
    n = data[0].shape[1]
    fp = method.startswith('fp_')
    np.random.seed(seed)
    random.seed(seed)
    net = Network([d.copy() for d in data], n_iter, burn_in, 1)
    g = dict(method=np.array(method), n_iter=n_iter, burn_in=burn_in, n_data_segments=len(data))
    for s, d in enumerate(data):
        g['data_seg%d' % s] = d
    kb = n if fp else 5
    pi = np.full((n, n_iter + 1, n), -1, dtype=np.int32); npi = np.zeros((n, n_iter + 1), dtype=np.int32)
    tau = np.full((n, n_iter, SM), -1, dtype=np.int32)
    betas = np.full((n, n_iter + 1, SM, kb), NAN); nseg = np.zeros((n, n_iter + 1), dtype=np.int32)
    padded = np.full((n, n_iter + 1, SM, n), NAN)
    bot, sot, cps_out, adj = [], [], [], []
    with contextlib.redirect_stdout(io.StringIO()):
        for i in range(n):
            net.set_network_configuration(i)
            net.fit(method)
            res = net.chain_results
            feats = [int(s_[1:]) for s_ in list(net.network_configuration['features'])]
            if fp:
                # network.py:303-347 with the plotting call of score_beta_matrix left out (the stub matplotlib has no hist)
                burned = res['betas_vector'][2:][burn_in:]
                thinned = [burned[x] for x in range(len(burned)) if x % 10 == 0]
                if method != 'fp_h_dbn':
                    bc = res['tau_vector'][burn_in:]
                    tcps = [bc[x] for x in range(len(bc)) if x % 10 == 0]
                    N = net.network_configuration['response']['y'].shape[0]
                    b_ = ref_scores.get_betas_over_time(N, tcps, thinned, n)
                    net.betas_over_time.append(b_)
                    net.scores_over_time.append(ref_scores.get_scores_over_time(b_, feats, n))
                else:
                    tcps = []
                if method == 'fp_glob_coup_nh_dbn':
                    bm_ = res['mu_vector'][burn_in:]
                    thinned = [[bm_[x]] for x in range(len(bm_)) if x % 10 == 0]
                bm = ref_scores.beta_post_matrix(thinned)
                row = [0.0] * n
                for j, f in enumerate(feats):
                    row[f] = ref_scores.credible_score(bm[:, j + 1])
                net.proposed_adj_matrix.append(row)
                net.cps_over_response.append(tcps)
            else:
                net.score_edges(i, method)
            for it, p in enumerate(res.get('pi_vector', [])):      # 'fp_varying_nh_dbn' returns no 'pi_vector'
                p = np.asarray(p, dtype=np.int32).ravel()
                pi[i, it, :p.size] = p; npi[i, it] = p.size
            for it, t in enumerate(res.get('tau_vector', [])):
                tau[i, it, :len(t)] = t
            for it, b in enumerate(res.get('betas_vector', [])):   # the h-dbn regressor returns no 'betas_vector'
                b = b if isinstance(b, list) else [b]
                betas[i, it] = _pad_list_of_vecs(b, kb); nseg[i, it] = len(b)
            if 'padded_betas' in res:
                for it, b in enumerate(res['padded_betas']):
                    b = b if isinstance(b, list) else [b]
                    padded[i, it] = _pad_list_of_vecs(b, n)
            adj.append(net.proposed_adj_matrix[-1])
            c = net.cps_over_response[-1]
            ca = np.full((len(c), SM), -1, dtype=np.int32)
            for x, t in enumerate(c):
                ca[x, :len(t)] = t
            cps_out.append(ca)
            if net.betas_over_time:
                bot.append(np.stack(net.betas_over_time[-1]) if len(net.betas_over_time) == i + 1 else None)
            if net.scores_over_time and len(net.scores_over_time) == i + 1:
                sot.append(net.scores_over_time[-1])
    g.update(pi=pi, npi=npi, tau=tau, betas=betas, nseg=nseg, padded_betas=padded, proposed_adj_matrix=np.array(adj, dtype=float))
    g['cps_over_response'] = np.stack(cps_out) if cps_out and all(c.shape == cps_out[0].shape for c in cps_out) else np.zeros((n, 0, SM), dtype=np.int32)
    if bot and all(b is not None for b in bot):
        g['betas_over_time'] = np.stack(bot)          # [n, N, n_kept, n]
    if sot:
        g['scores_over_time'] = np.stack(sot)         # [n, N, n - 1]
    path = os.path.join(HERE, 'network_%s_%s.npz' % (name, method))
    np.savez_compressed(path, **g)
    print(path, {k: getattr(v, 'shape', None) for k, v in g.items() if k in ('betas_over_time', 'scores_over_time', 'cps_over_response')})


def lag2(name, data):
    """network.py:92-122 with lag = 2: the lagged design dict of every node."""
    n = data[0].shape[1]
    net = Network([d.copy() for d in data], 10, 0, 2)
    g = dict(n_data_segments=len(data))
    for s, d in enumerate(data):
        g['data_seg%d' % s] = d
    for i in range(n):
        net.set_network_configuration(i)
        cfg = net.network_configuration
        labels = list(cfg['features'].keys())
        g['labels_%d' % i] = np.array([int(l[1:]) for l in labels], dtype=np.int32)
        g['features_%d' % i] = np.stack([cfg['features'][l] for l in labels], axis=1)
        g['response_%d' % i] = cfg['response']['y']
    path = os.path.join(HERE, 'network_%s_lag2_config.npz' % name)
    np.savez_compressed(path, **g)
    print(path)


def main():
    yeast = _refboot.read_yeast()
    for method in ('varying_nh_dbn', 'h_dbn', 'glob_coup_nh_dbn', 'fp_varying_nh_dbn', 'fp_glob_coup_nh_dbn'):
        run('yeast', yeast, method, 150, 30, 11)
    lag2('yeast', yeast)


if __name__ == '__main__' and 'evaluation-only' not in sys.argv:
    main()


def evaluation(name):
    """examples/utils.py:24-131 on the reference's own adjacency matrices (lag 1 and lag 2 fixtures) against the yeast truth
    of examples/yeast_tests.py:58-64: transformResults' outputs and the two AUCs the example drivers print."""
    import importlib.util
    import types
    for m in ('pandas',):
        pass
    spec = importlib.util.spec_from_file_location('ref_examples_utils', '/root/reference/examples/utils.py')
    mod = importlib.util.module_from_spec(spec)
    cwd = os.getcwd()
    import tempfile
    os.chdir(tempfile.mkdtemp(prefix='dyban_ref_'))     # the module opens ./output.log at import time
    try:
        spec.loader.exec_module(mod)
    finally:
        os.chdir(cwd)
    from sklearn.metrics import roc_curve, auc, precision_recall_curve
    true_inc = [[0, 0, 1, 0, 1], [1, 0, 0, 1, 0], [0, 1, 0, 0, 0], [0, 1, 1, 0, 0], [0, 0, 1, 0, 0]]
    g = dict(true_inc=np.array(true_inc))
    for tag, path in (('lag1', os.path.join(HERE, 'chain_yeast_varying_nh_dbn.npz')), ('lag2', os.path.join(HERE, 'lag2chain_yeast_varying_nh_dbn.npz'))):
        adj = np.load(path)['proposed_adj_matrix']
        flat_true, flat_scores = mod.transformResults([list(r) for r in true_inc], [list(map(float, r)) for r in adj])
        fpr, tpr, _ = roc_curve(flat_true, flat_scores)
        pr, rc, _ = precision_recall_curve(flat_true, flat_scores)
        g['adj_' + tag] = adj
        g['flat_true_' + tag] = np.array(flat_true)
        g['flat_scores_' + tag] = np.array(flat_scores, dtype=float)
        g['roc_auc_' + tag] = auc(fpr, tpr)
        g['pr_auc_' + tag] = auc(rc, pr)
    path = os.path.join(HERE, 'evaluation_%s.npz' % name)
    np.savez_compressed(path, **g)
    print(path, float(g['roc_auc_lag1']), float(g['pr_auc_lag1']), float(g['roc_auc_lag2']), float(g['pr_auc_lag2']))


if __name__ == '__main__' and 'evaluation-only' in sys.argv:
    evaluation('yeast')
