"""The oracle (oracle/dbn_oracle.py) against outputs of the unmodified reference (tests/golden/*.npz)."""
import math

import numpy as np
import pytest

from helpers import LAG_CHAIN_FILES, labels_to_columns, FP_PLAIN_FILES, FP_VV_FILES, FP_SEQ_FILES, KAT_FILES, CHAIN_FILES, FP_CHAIN_FILES, LOGML_RTOL, load, data_list, ids, assert_close
from oracle import dbn_oracle as O


@pytest.mark.parametrize('path', KAT_FILES, ids=ids(KAT_FILES))
def test_function_kats(path):
    g = load(path)
    X, Y = O.lagged_matrices(data_list(g))
    N = int(g['N'])
    assert X.shape[0] == N
    n_cases = g['node'].shape[0]
    for c in range(n_cases):
        node, npi, S = int(g['node'][c]), int(g['npi'][c]), int(g['S'][c])
        pi = [int(v) for v in g['pi'][c, :npi]]
        tau = [int(v) for v in g['tau'][c, :S]]
        lam, dlt, sig = float(g['lambda2'][c]), float(g['delta2'][c]), float(g['sigma2'][c])
        a, b, T = float(g['alpha'][c]), float(g['beta'][c]), float(g['T'][c])
        k = npi + 1
        mu = g['mu'][c, :k]
        y = Y[:, node]
        stats = O.segment_stats(X, y, pi, tau)
        Xs, ys = O.segment_design(X, y, pi, tau)
        zero = [np.zeros(k)] * S
        # log-ML, Gram form and literal form, vs the reference (1e-9 relative, north_star)
        for name, mus, seq in (('logml_plain', zero, False), ('logml_mu', [mu] * S, False)):
            ref = float(g[name][c])
            if not math.isfinite(ref):
                continue  # reference overflow (T >= 344 never reached in these fixtures)
            assert_close(O.log_ml(stats, mus, a, b, lam, T), ref, LOGML_RTOL, what='%s[%d]' % (name, c))
            assert_close(O.log_ml_literal(Xs, ys, mus, a, b, lam, T), ref, LOGML_RTOL, what='%s-literal[%d]' % (name, c))
        bt = O.beta_tilde(stats, lam, dlt)
        for h in range(S):
            assert_close(bt[h], g['beta_tilde'][c, h, :k], 1e-8, 1e-11, what='beta_tilde[%d,%d]' % (c, h))
        assert_close(O.log_ml(stats, bt, a, b, lam, T, True, dlt), float(g['logml_seq'][c]), LOGML_RTOL,
                     what='logml_seq[%d]' % c)
        assert_close(O.log_ml_literal(Xs, ys, bt, a, b, lam, T, True, dlt), float(g['logml_seq'][c]), LOGML_RTOL,
                     what='logml_seq-literal[%d]' % c)
        raw = float(g['rawml_h'][c])
        if S == 1 and math.isfinite(raw) and raw > 0:
            assert_close(O.log_ml_h(stats, a, b, lam, N), math.log(raw), LOGML_RTOL, what='rawml_h[%d]' % c)
        # priors
        assert_close(O.log_cp_prior(tau), math.log(float(g['cp_prior'][c])), 1e-12, what='cp_prior[%d]' % c)
        F = X.shape[1] - 1
        assert_close(O.pi_prior(pi, F, int(g['fan_in'][c])), float(g['pi_prior'][c]), 1e-14, what='pi_prior[%d]' % c)
        # muSampler: MVN arguments and the density of the input mu
        mean, Sigma, Q = O.mu_posterior(stats, mu, sig, lam)
        assert_close(mean, g['mus_mean'][c, :k], 1e-8, 1e-11, what='mus_mean[%d]' % c)
        assert_close(Sigma, g['mus_cov'][c, :k, :k], 1e-8, 1e-12, what='mus_cov[%d]' % c)
        ref_ld = float(g['mus_logdens_before'][c])
        if math.isfinite(ref_ld):
            assert_close(O.mvn_logpdf_prec(mu, mean, Q), ref_ld, 1e-8, 1e-9, what='mus_logdens_before[%d]' % c)


@pytest.mark.parametrize('path', CHAIN_FILES, ids=ids(CHAIN_FILES))
def test_chain_replay(path):
    """Replay the reference's recorded draw stream: accept flags, pi and tau must match bit-exactly, every Gibbs
    parameter and log-ML to 1e-9 relative."""
    g = load(path)
    method = str(g['method'])
    X, Y = O.lagged_matrices(data_list(g))
    n = X.shape[1]
    n_iter, burn_in, N = int(g['n_iter']), int(g['burn_in']), int(g['N'])
    edge = np.zeros((n, n)); nonempty = np.zeros(n)
    cp_hist = np.zeros((n, N + 3)); cp_kept = np.zeros(n)
    for node in range(n):
        tr = O.run_chain(method, X, Y, node, n_iter, O.ReplayDraws(g, node), init_tau=g.get('change_points'))
        for key in ('pi_valid', 'pi_accept', 'npi', 'pi_after', 'S', 'tau_after'):
            assert np.array_equal(tr[key], g[key][node]), (key, node)
        if method != 'h_dbn':
            for key in ('tau_valid', 'tau_accept'):
                assert np.array_equal(tr[key], g[key][node]), (key, node)
        for key in ('sig_shape', 'sig_scale', 'lam_shape', 'lam_scale', 'del_shape', 'del_scale', 'beta_mean'):
            assert_close(tr[key], g[key][node], 1e-9, 1e-12, what='%s node %d' % (key, node))
        assert_close(tr['beta_cov'], g['beta_cov'][node], 1e-9, 1e-13, what='beta_cov node %d' % node)
        assert_close(tr['logml'], g['logml'][node], LOGML_RTOL, what='logml node %d' % node)
        if method == 'var_glob_coup_nh_dbn':
            for key in ('sig_shape_v', 'sig_scale_v'):
                assert_close(tr[key], g[key][node], 1e-9, 1e-12, what='%s node %d' % (key, node))
        if method in ('glob_coup_nh_dbn', 'var_glob_coup_nh_dbn'):
            assert_close(tr['mus_mean'], g['mus_mean'][node], 1e-8, 1e-11, what='mus_mean node %d' % node)
            assert_close(tr['mus_cov'], g['mus_cov'][node], 1e-8, 1e-12, what='mus_cov node %d' % node)
            ref = g['mus_logdens'][node].copy()
            mine = tr['mus_logdens']
            ref[np.isnan(mine)] = np.nan       # the oracle only fills the densities the acceptance ratio uses
            assert_close(mine, ref, 1e-8, 1e-9, what='mus_logdens node %d' % node)
            assert_close(tr['mu_after'], g['mu_after'][node], 0, 0, what='mu_after node %d' % node)
        O.accumulate_counts(tr, node, n, N, burn_in, edge, nonempty, cp_hist, cp_kept)
    adj = O.edge_scores(edge, nonempty)
    ref_adj = g['proposed_adj_matrix']
    ok = np.isfinite(ref_adj).all(axis=1)
    assert_close(adj[ok], ref_adj[ok], 1e-14, what='proposed_adj_matrix')
    if 'cp_hist' in g:
        assert np.array_equal(cp_hist, g['cp_hist'])
        assert np.array_equal(cp_kept, g['cp_kept'])


@pytest.mark.parametrize('path', FP_CHAIN_FILES, ids=ids(FP_CHAIN_FILES))
def test_fp_glob_chain_replay(path):
    """Full-parents globally coupled chains (fpGlobCoupBpwLinReg.py:16-119) replayed from the reference's recorded
    draws: changepoint lists and accept flags bit-exact, Gibbs parameters / log-MLs / mu-sampler means and densities
    to 1e-9..1e-8, and the credible-score row of network.py:303-347 from the thinned mu chain."""
    g = load(path)
    X, Y = O.lagged_matrices(data_list(g))
    n = X.shape[1]
    n_iter, burn_in = int(g['n_iter']), int(g['burn_in'])
    adj = np.zeros((n, n))
    for node in range(n):
        tr = O.run_chain('fp_glob_coup_nh_dbn', X, Y, node, n_iter, O.ReplayDraws(g, node))
        for key in ('S', 'tau_after', 'tau_valid', 'tau_accept'):
            assert np.array_equal(tr[key], g[key][node]), (key, node)
        for key in ('sig_shape', 'sig_scale', 'lam_shape', 'lam_scale'):
            assert_close(tr[key], g[key][node], 1e-9, 1e-12, what='%s node %d' % (key, node))
        # n-wide posterior means on the GP-smoothed (near-collinear) mTOR series: the reference's own inv() carries
        # ~1e-9 relative noise in single entries, so entries are compared at 1e-7 of the vector's scale
        assert_close(tr['beta_mean'], g['beta_mean'][node], 1e-7, 1e-9, what='beta_mean node %d' % node)
        assert_close(np.diagonal(tr['beta_cov'], axis1=-2, axis2=-1), g['beta_cov_diag'][node], 1e-7, 1e-13, what='beta_cov')
        assert_close(tr['logml'], g['logml'][node], LOGML_RTOL, what='logml node %d' % node)
        assert_close(tr['mus_mean'], g['mus_mean'][node], 1e-7, 1e-9, what='mus_mean node %d' % node)
        assert_close(np.diagonal(tr['mus_cov'], axis1=-2, axis2=-1), g['mus_cov_diag'][node], 1e-7, 1e-12, what='mus_cov')
        ref = g['mus_logdens'][node].copy()
        mine = tr['mus_logdens']
        ref[np.isnan(mine)] = np.nan
        assert_close(mine, ref, 1e-8, 1e-8, what='mus_logdens node %d' % node)
        assert_close(tr['mu_after'], g['mu_after'][node], 0, 0, what='mu_after node %d' % node)
        # thinned mu chain: index 0 is the initial zero vector (fpGlobCoupBpwLinReg.py:66)
        mu_vec = [np.zeros(n)] + [tr['mu_after'][it, :n] for it in range(n_iter)]
        kept = [mu_vec[i] for i in range(len(mu_vec)) if O.mu_kept(i, burn_in)]
        adj[node] = O.credible_scores(np.array(kept), [j for j in range(n) if j != node], n)
    assert_close(adj, g['proposed_adj_matrix'], 1e-14, what='proposed_adj_matrix (credible scores)')


@pytest.mark.parametrize('path', FP_PLAIN_FILES + FP_VV_FILES + FP_SEQ_FILES, ids=ids(FP_PLAIN_FILES + FP_VV_FILES + FP_SEQ_FILES))
def test_fp_plain_chain_replay(path):
    """fp_varying_nh_dbn (fullParentsBpwLinReg.py:41-140) and fp_h_dbn (fpBayesianLinearRegression.py:40-125) replayed from
    the reference's recorded draws: changepoint lists / accept flags bit-exact, Gibbs parameters and log-MLs, and the
    credible-score row of network.py:316-347 from the thinned beta chain (every segment's vector is a sample)."""
    g = load(path)
    method = str(g['method'])
    X, Y = O.lagged_matrices(data_list(g))
    n = X.shape[1]
    n_iter, burn_in, N = int(g['n_iter']), int(g['burn_in']), int(g['N'])
    adj = np.zeros((n, n))
    cp_hist = np.zeros((n, N + 3)); cp_kept = np.zeros(n)
    for node in range(n):
        tr = O.run_chain(method, X, Y, node, n_iter, O.ReplayDraws(g, node))
        assert np.array_equal(tr['S'], g['S'][node]) and np.array_equal(tr['tau_after'], g['tau_after'][node]), node
        if method != 'fp_h_dbn':
            for key in ('tau_valid', 'tau_accept'):
                assert np.array_equal(tr[key], g[key][node]), (key, node)
        for key in ('sig_shape', 'sig_scale', 'lam_shape', 'lam_scale') + (('del_shape', 'del_scale') if 'fp_seq' in method else ()):
            assert_close(tr[key], g[key][node], 1e-8 if 'fp_seq' in method else 1e-9, 1e-12, what='%s node %d' % (key, node))
        assert_close(tr['beta_mean'], g['beta_mean'][node], 1e-7, 1e-9, what='beta_mean node %d' % node)
        assert_close(np.diagonal(tr['beta_cov'], axis1=-2, axis2=-1), g['beta_cov_diag'][node], 1e-7, 1e-13, what='beta_cov')
        assert_close(tr['logml'], g['logml'][node], LOGML_RTOL, what='logml node %d' % node)
        if method == 'fp_var_glob_coup_nh_dbn':      # fpvvGlobCoup.py: per-segment variances, mu redrawn every iteration
            for key in ('sig_shape_v', 'sig_scale_v'):
                assert_close(tr[key], g[key][node], 1e-9, 1e-12, what='%s node %d' % (key, node))
            assert_close(tr['mus_mean'][:, 0], g['mus_mean'][node][:, 0], 1e-7, 1e-9, what='mus_mean node %d' % node)
            assert_close(np.diagonal(tr['mus_cov'][:, 0], axis1=-2, axis2=-1), g['mus_cov_diag'][node][:, 0], 1e-7, 1e-12, what='mus_cov')
            assert_close(tr['mu_after'], g['mu_after'][node], 0, 0, what='mu_after node %d' % node)
        samples = [tr['beta'][it, h, :n] for it in range(n_iter) if O.beta_kept(it, burn_in) for h in range(tr['S'][it])]
        adj[node] = O.credible_scores(np.array(samples), [j for j in range(n) if j != node], n)
        for it in range(n_iter):
            if O.tau_kept(it, burn_in) and method != 'fp_h_dbn':
                cp_kept[node] += 1
                for c in [c for c in tr['tau_after'][it] if c >= 0][:-1]:
                    cp_hist[node, c] += 1
    assert_close(adj, g['proposed_adj_matrix'], 1e-14, what='proposed_adj_matrix (credible scores of the beta chain)')
    if 'cp_hist' in g:
        assert np.array_equal(cp_hist, g['cp_hist']) and np.array_equal(cp_kept, g['cp_kept'])


@pytest.mark.parametrize('path', LAG_CHAIN_FILES, ids=ids(LAG_CHAIN_FILES))
def test_lagged_design_chain_replay(path):
    """lag = 2 (network.py:92-122: zero-padded lagged copies, relabelled): the oracle on the lagged design reproduces the
    reference's recorded chains -- parents, changepoints, accept flags, log-MLs, Gibbs parameters, adjacency rows."""
    g = load(path)
    method, lag = str(g['method']), int(g['lag'])
    data = data_list(g)
    n = data[0].shape[1]
    X, Y = O.lagged_matrices(data, lag)
    assert X.shape[1] == n * lag
    N = X.shape[0]
    n_iter = g['sigma2'].shape[1]
    cols = labels_to_columns(g['pi_after'], n)
    assert (cols >= n).any()                     # the fixture does use lag-2 parents
    edge = np.zeros((n, n)); ne = np.zeros(n); cph = np.zeros((n, N + 3)); cpk = np.zeros(n)
    for node in range(n):
        tr = O.run_chain(method, X, Y, node, n_iter, O.ReplayDraws(g, node))
        assert np.array_equal(tr['pi_after'], cols[node]), node
        lab = np.vectorize(lambda c: -1 if c < 0 else O.column_to_label(c, node, n))(tr['pi_after'])
        assert np.array_equal(lab, g['pi_after'][node])
        assert np.array_equal(tr['tau_after'], g['tau_after'][node])
        for key in ('pi_valid', 'pi_accept', 'npi', 'S'):
            assert np.array_equal(tr[key], g[key][node]), key
        assert_close(tr['logml'], g['logml'][node], LOGML_RTOL, what='logml')
        assert_close(tr['sig_scale'], g['sig_scale'][node], 1e-9, what='sig_scale')
        assert_close(tr['lam_scale'], g['lam_scale'][node], 1e-9, what='lam_scale')
        O.accumulate_counts(tr, node, n, N, int(g['burn_in']), edge, ne, cph, cpk)
    ref = g['proposed_adj_matrix']
    ok = np.isfinite(ref).all(axis=1)
    assert ref.shape[1] == lag * (n - 1) + 1 and (ref[ok][:, n:] == 0).all()      # scores.py:156: len(features) + 1 entries
    assert_close(O.edge_scores(edge, ne)[ok], ref[ok][:, :n], 1e-14, what='proposed_adj_matrix')
