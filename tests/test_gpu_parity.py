"""GPU parity tests proper: the CUDA path (through the C ABI) against the oracle and the reference goldens."""
import math

import numpy as np
import pytest

from helpers import KAT_FILES, CHAIN_FILES, FP_CHAIN_FILES, FP_PLAIN_FILES, FP_VV_FILES, FP_SEQ_FILES, LOGML_RTOL, load, data_list, ids, assert_close, assert_close_mat, golden_replay
from oracle import dbn_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def eng_mod():
    from dynamicbayesiannetworks_b200.engine import DbnEngine
    return DbnEngine


# ---------------------------------------------------------------------------------------------------------
# kernel 1: prefix statistics == cumulative sums of outer products of the lag-1 rows
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('path', KAT_FILES, ids=ids(KAT_FILES))
def test_prefix_rows(eng_mod, path):
    g = load(path)
    data = data_list(g)
    X, Y = O.lagged_matrices(data)
    N, n = X.shape
    Z = np.concatenate([np.ones((N, 1)), X], axis=1)
    with eng_mod(n_chains=1) as e:
        e.build_stats(data)
        assert e.stats_info()['N'] == N
        for t in sorted({0, 1, 2, N // 3, N // 2, N - 1, N}):
            zz, zy, yy = e.stats_row(t)
            assert_close(zz, Z[:t].T @ Z[:t], 1e-12, 1e-12, what='ZZ[%d]' % t)
            assert_close(zy, (Z[:t].T @ Y[:t]).T, 1e-12, 1e-12, what='ZY[%d]' % t)
            assert_close(yy, (Y[:t] ** 2).sum(axis=0), 1e-12, 1e-12, what='YY[%d]' % t)


# ---------------------------------------------------------------------------------------------------------
# kernel 2: batched scorer vs the reference's own log-ML values (1e-9 relative, north_star)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('path', KAT_FILES, ids=ids(KAT_FILES))
def test_score_batch_vs_reference(eng_mod, path):
    g = load(path)
    B = g['node'].shape[0]
    pi = [list(g['pi'][c, :g['npi'][c]]) for c in range(B)]
    tau = [list(g['tau'][c, :g['S'][c]]) for c in range(B)]
    mu = [g['mu'][c, :g['npi'][c] + 1] for c in range(B)]
    with eng_mod(n_chains=1) as e:
        e.build_stats(data_list(g))
        for ab in (0.005, 0.01):
            for T in sorted(set(g['T'])):
                sel = np.nonzero((g['alpha'] == ab) & (g['T'] == T))[0]
                if sel.size == 0:
                    continue
                sub = lambda lst: [lst[c] for c in sel]
                args = (list(g['node'][sel]), sub(pi), sub(tau), g['lambda2'][sel])
                plain = e.score_batch(*args, alpha=ab, beta=ab, T=T)
                with_mu = e.score_batch(*args, mu=sub(mu), alpha=ab, beta=ab, T=T)
                seq = e.score_batch(*args, delta2=g['delta2'][sel], alpha=ab, beta=ab, T=T, seq=True)
                assert_close(plain, g['logml_plain'][sel], LOGML_RTOL, what='plain')
                assert_close(with_mu, g['logml_mu'][sel], LOGML_RTOL, what='mu')
                assert_close(seq, g['logml_seq'][sel], LOGML_RTOL, what='seq')
        # h-dbn: log of the un-logged calculateMarginalLikelihood (marginalLikelihood.py:152-166), N rows, T = N
        N = int(g['N'])
        raw = g['rawml_h']
        sel = np.nonzero(np.isfinite(raw) & (raw > 0))[0]
        if sel.size:
            for ab in (0.005, 0.01):
                s2 = sel[g['alpha'][sel] == ab]
                if s2.size:
                    out = e.score_batch(list(g['node'][s2]), [pi[c] for c in s2], [[N + 2]] * s2.size, g['lambda2'][s2],
                                        alpha=ab, beta=ab, T=N)
                    assert_close(out, np.log(raw[s2]), LOGML_RTOL, what='h')


def test_score_batch_large_T_vs_oracle(eng_mod):
    """BASELINE config 4 shape (n=100, T=2000): beyond the reference's overflow limit, the oracle is the check."""
    from dynamicbayesiannetworks_b200.synth import make_synthetic
    data, _, _ = make_synthetic(100, 2000, n_changepoints=3, max_fan_in=4, seed=20260119)
    X, Y = O.lagged_matrices([data])
    N, n = X.shape
    rng = np.random.default_rng(5)
    B = 300
    node = rng.integers(0, n, B)
    pi, tau, mu, lam, dlt = [], [], [], [], []
    for b in range(B):
        feats = [j for j in range(n) if j != node[b]]
        pi.append([int(v) for v in rng.choice(feats, size=rng.integers(0, 5), replace=False)])
        S = int(rng.integers(1, 10))
        tau.append(sorted(int(v) for v in rng.choice(np.arange(2, N + 1), size=S - 1, replace=False)) + [N + 2])
        mu.append(rng.normal(0, 0.3, len(pi[-1]) + 1))
        lam.append(float(np.exp(rng.uniform(-3, 3)))); dlt.append(float(np.exp(rng.uniform(-3, 3))))
    with eng_mod(n_chains=1) as e:
        e.build_stats([data])
        out_mu = e.score_batch(list(node), pi, tau, lam, mu=mu, alpha=0.005, beta=0.005, T=N)
        out_seq = e.score_batch(list(node), pi, tau, lam, delta2=dlt, alpha=0.005, beta=0.005, T=N - 1, seq=True)
    for b in range(B):
        stats = O.segment_stats(X, Y[:, node[b]], pi[b], tau[b])
        ref = O.log_ml(stats, [mu[b]] * len(tau[b]), 0.005, 0.005, lam[b], N)
        assert_close(out_mu[b], ref, LOGML_RTOL, what='large-T mu[%d]' % b)
        ref = O.log_ml(stats, O.beta_tilde(stats, lam[b], dlt[b]), 0.005, 0.005, lam[b], N - 1, True, dlt[b])
        assert_close(out_seq[b], ref, LOGML_RTOL, what='large-T seq[%d]' % b)


# ---------------------------------------------------------------------------------------------------------
# kernel 3: replay of the reference's recorded draw stream
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('path', CHAIN_FILES, ids=ids(CHAIN_FILES))
def test_chain_replay_vs_reference(eng_mod, path):
    """Given the identical injected draw stream, accept/reject flags, parent sets and changepoints must match the
    reference bit-exactly; Gibbs parameters and log-MLs to 1e-9 relative."""
    g = load(path)
    method = str(g['method'])
    n, n_iter = g['sigma2'].shape
    N, burn_in = int(g['N']), int(g['burn_in'])
    with eng_mod(n_chains=1) as e:
        e.build_stats(data_list(g))
        e.chain_init(method, burn_in=burn_in)
        if 'change_points' in g:            # 'fixed_nh_dbn' (network.py:178-189)
            e.set_changepoints(g['change_points'])
        tr = e.run(n_iter, replay=golden_replay(g), trace=True)
        counts = e.counts()
    for key in ('pi_valid', 'pi_accept', 'npi', 'S'):
        assert np.array_equal(tr[key], g[key]), key
    assert np.array_equal(tr['pi_after'], g['pi_after']), 'pi_after'
    if method not in ('h_dbn', 'fixed_nh_dbn'):
        for key in ('tau_valid', 'tau_accept'):
            assert np.array_equal(tr[key], g[key]), key
    assert np.array_equal(tr['tau_after'], g['tau_after']), 'tau_after'
    assert_close(tr['sig_shape'], g['sig_shape'], 1e-12, what='sig_shape')
    assert_close(1.0 / tr['sig_rate'], g['sig_scale'], 1e-9, what='sig_scale')
    assert_close(tr['lam_shape'], g['lam_shape'], 1e-12, what='lam_shape')
    assert_close(1.0 / tr['lam_rate'], g['lam_scale'], 1e-9, what='lam_scale')
    if method == 'seq_coup_nh_dbn':
        assert_close(tr['del_shape'], g['del_shape'], 1e-12, what='del_shape')
        assert_close(1.0 / tr['del_rate'], g['del_scale'], 1e-9, what='del_scale')
    assert_close(tr['logml'], g['logml'], LOGML_RTOL, what='logml')
    assert_close(tr['beta_mean'], g['beta_mean'], 1e-8, 1e-11, what='beta_mean')
    assert_close_mat(tr['beta_cov'], g['beta_cov'], 1e-9, what='beta_cov')
    if method == 'var_glob_coup_nh_dbn':
        assert_close(tr['sig_shape_v'], g['sig_shape_v'], 1e-12, what='sig_shape_v')
        assert_close(1.0 / tr['sig_rate_v'], g['sig_scale_v'], 1e-9, what='sig_scale_v')
    if method in ('glob_coup_nh_dbn', 'var_glob_coup_nh_dbn'):
        assert_close(tr['mus_mean'], g['mus_mean'], 1e-8, 1e-11, what='mus_mean')
        assert_close_mat(tr['mus_cov'], g['mus_cov'], 1e-9, what='mus_cov')
        ref = np.stack([g['mus_logdens'][:, :, 0, 0], g['mus_logdens'][:, :, 1, 1], g['mus_logdens'][:, :, 2, 1],
                        g['mus_logdens'][:, :, 3, 0]], axis=-1)
        assert_close(tr['mus_logdens'], ref, 1e-8, 1e-9, what='mus_logdens')
        assert_close(tr['mu_after'], g['mu_after'], 0, 0, what='mu_after')
    # thinned-chain counts -> proposed_adj_matrix / changepoint histogram (network.py:355-368,388-389)
    adj = counts['edge'] / np.maximum(counts['nonempty'], 1)[:, None]
    ref_adj = g['proposed_adj_matrix']
    ok = np.isfinite(ref_adj).all(axis=1)
    assert_close(adj[ok], ref_adj[ok], 1e-14, what='proposed_adj_matrix')
    if 'cp_hist' in g:
        assert np.array_equal(counts['cp_hist'], g['cp_hist'].astype(np.int64))
        assert np.array_equal(counts['cp_kept'], g['cp_kept'].astype(np.int64))


def test_chain_replay_split_launches(eng_mod):
    """State carried across dbn_run calls: two launches of n/2 iterations == one launch of n."""
    g = load([p for p in CHAIN_FILES if 'yeast_varying' in p][0])
    n, n_iter = g['sigma2'].shape
    r = golden_replay(g)
    half = n_iter // 2
    with eng_mod(n_chains=1) as e:
        e.build_stats(data_list(g))
        e.chain_init('varying_nh_dbn', burn_in=int(g['burn_in']))
        a = e.run(half, replay={k: v[:, :half] for k, v in r.items()}, trace=True)
        b = e.run(n_iter - half, replay={k: v[:, half:] for k, v in r.items()}, trace=True)
    assert np.array_equal(np.concatenate([a['tau_after'], b['tau_after']], axis=1), g['tau_after'])
    assert np.array_equal(np.concatenate([a['pi_after'], b['pi_after']], axis=1), g['pi_after'])


# ---------------------------------------------------------------------------------------------------------
# free-running Philox chains: statistical agreement with the oracle's free-running chains
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('method', ['h_dbn', 'varying_nh_dbn', 'seq_coup_nh_dbn', 'glob_coup_nh_dbn', 'var_glob_coup_nh_dbn'])
def test_free_running_state_sanity(eng_mod, method):
    g = load([p for p in CHAIN_FILES if 'yeast_varying' in p][0])
    data = data_list(g)
    X, Y = O.lagged_matrices(data)
    n = X.shape[1]
    N = X.shape[0]
    n_iter, burn = 1500, 500
    C = 256
    with eng_mod(n_chains=C, seed=1234) as e:
        e.build_stats(data)
        e.chain_init(method, burn_in=burn)
        e.run(n_iter)
        c = e.counts()
        st = e.state()
        ctr = e.counters()
        assert e.verify_cache() == (0, 0)   # h / nh: the per-unit segment-statistics cache still equals the table's values
    assert ctr['unit_iterations'] == n * C * n_iter
    assert np.isfinite(st['sigma2']).all() and (st['sigma2'] > 0).all()
    assert np.isfinite(st['lambda2']).all() and (st['lambda2'] > 0).all()
    gpu = c['edge'] / np.maximum(c['nonempty'], 1)[:, None]
    assert np.isfinite(gpu).all() and (gpu >= 0).all() and (gpu <= 1).all() and (np.diag(gpu) == 0).all()
    # the comparison with the oracle's free-running chains (Monte-Carlo error bound) is
    # tests/test_gpu_pinning.py::test_free_running_edge_posteriors_within_mc_error


@pytest.mark.parametrize('method', ['h_dbn', 'varying_nh_dbn'])
def test_plain_and_trace_builds_follow_the_same_trajectory(eng_mod, method):
    """The TRACE instantiation of the h / nh kernel (the one the replay fixtures pin) and the plain one (the one that is
    timed) draw the same Philox stream, so free-running chains must agree state by state: discrete state exactly,
    sigma^2 / lambda^2 to rounding.  Many chains per warp, launches of mixed lengths, cache checked after each."""
    g = load([p for p in CHAIN_FILES if 'yeast_varying' in p][0])
    data = data_list(g)
    C = 256
    windows = [1, 7, 40, 152, 200]
    with eng_mod(n_chains=C, seed=77) as e:
        e.build_stats(data)
        e.chain_init(method, burn_in=0)
        tr = e.run(sum(windows), trace=True)
    with eng_mod(n_chains=C, seed=77) as e:
        e.build_stats(data)
        e.chain_init(method, burn_in=0)
        done = 0
        for w in windows:
            e.run(w)
            done += w
            assert e.verify_cache() == (0, 0), done
            st = e.state()
            it = done - 1
            S = tr['S_after'][:, it] if method != 'h_dbn' else st['S']
            assert np.array_equal(st['S'], S), done
            assert np.array_equal(st['npi'], tr['npi'][:, it]), done
            for u in range(st['S'].size):
                assert np.array_equal(st['pi'][u, :st['npi'][u]], tr['pi_after'][u, it, :st['npi'][u]]), (done, u)
                if method != 'h_dbn':
                    assert np.array_equal(st['tau'][u, :S[u]], tr['tau_after'][u, it, :S[u]]), (done, u)
            assert np.allclose(st['sigma2'], tr['sigma2'][:, it], rtol=1e-9, atol=0), done
            assert np.allclose(st['lambda2'], tr['lambda2'][:, it], rtol=1e-9, atol=0), done


def test_network_dropin(eng_mod):
    """dyban.Network surface on the device path: result attributes, history lengths, pickling."""
    import pickle
    from dynamicbayesiannetworks_b200 import Network
    g = load([p for p in CHAIN_FILES if 'yeast_varying' in p][0])
    data = data_list(g)
    for method in ('h_dbn', 'varying_nh_dbn', 'seq_coup_nh_dbn', 'glob_coup_nh_dbn', 'var_glob_coup_nh_dbn'):
        net = Network(data, 300, 100, 1)
        net.infer_network(method)
        adj = np.array(net.proposed_adj_matrix)
        assert adj.shape == (5, 5) and np.isfinite(adj).all() and (adj >= 0).all() and (adj <= 1).all()
        assert np.allclose(np.diag(adj), 0)
        r = net.chain_results
        assert len(r['pi_vector']) == 301 and len(r['sigma_sqr_vector']) == 301 and len(r['lambda_sqr_vector']) == 301
        if method != 'h_dbn':
            assert len(r['tau_vector']) == 300 and all(t[-1] == 35 for t in r['tau_vector'])
            assert len(net.cps_over_response) == 5 and len(net.cps_over_response[0]) == 20
        if method == 'seq_coup_nh_dbn':
            assert len(r['delta_sqr_vector']) == 301
        if method == 'var_glob_coup_nh_dbn':   # vvglobCoup.py:70-72: one variance per segment of the iteration's tau
            assert all(len(r['sigma_sqr_vector'][it + 1]) == (1 if it == 0 else len(r['tau_vector'][it - 1])) for it in range(300))
            assert len(r['mu_vector']) == 301
        pickle.loads(pickle.dumps(net))
    fixed = Network(data, 300, 100, 1, [12, 23, 35])          # network.py:178-189: predefined change points
    fixed.infer_network('fixed_nh_dbn')
    assert fixed.chain_results['tau_vector'] == []      # only the 'varying_nh' regressor records tau (bayesianPwLinearRegression.py:117-121)
    assert all(c == [35] for c in fixed.cps_over_response[0])   # network.py:376-383: the artificial [N + 2] list
    assert np.isfinite(np.array(fixed.proposed_adj_matrix)).all()
    assert fixed.network_args['type'] == 'Varying Parents-Fixed changepoints'
    with pytest.raises(RuntimeError):                          # the list must end with the pseudo changepoint N + 2
        Network(data, 10, 0, 1, [12, 23]).infer_network('fixed_nh_dbn')
    big = Network(data, 400, 100, 1, n_chains=128)
    big.infer_network('nh-dbn')
    assert big.chain_results is None and np.array(big.proposed_adj_matrix).shape == (5, 5)
    with pytest.raises(ValueError):
        Network(data, 10, 0, 1).infer_network('no_such_dbn')


# ---------------------------------------------------------------------------------------------------------
# kernel 4: full-parents globally coupled chains (design width k = n)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('path', FP_CHAIN_FILES, ids=ids(FP_CHAIN_FILES))
def test_fp_glob_replay_vs_reference(eng_mod, path):
    """fpGlobCoupBpwLinReg.py:68-110 replayed from the reference's recorded draws: changepoint lists and accept flags
    bit-exact; log-MLs 1e-9; Gibbs parameters, muSampler means / densities as in the CPU oracle test; credible scores."""
    from dynamicbayesiannetworks_b200 import parallel
    g = load(path)
    n, n_iter = g['sigma2'].shape
    burn_in = int(g['burn_in'])
    replay = dict(sigma2=g['sigma2'], lambda2=g['lambda2'], u_tau=g['u_tau'], mu_tau=g['mus_val'][:, :, 3, :],
                  beta=g['beta'], tau_move=g['tau_move'], tau_pick=g['tau_pick'])
    with eng_mod(n_chains=1) as e:
        e.build_stats(data_list(g))
        e.chain_init('fp_glob_coup_nh_dbn', burn_in=burn_in)
        tr = e.run_fp(n_iter, replay=replay, trace=True)
        counts = e.counts()
    for key in ('S', 'tau_valid', 'tau_accept', 'tau_after'):
        assert np.array_equal(tr[key], g[key]), key
    assert_close(tr['sig_shape'], g['sig_shape'], 1e-12, what='sig_shape')
    assert_close(1.0 / tr['sig_rate'], g['sig_scale'], 1e-9, what='sig_scale')
    assert_close(tr['lam_shape'], g['lam_shape'], 1e-12, what='lam_shape')
    assert_close(1.0 / tr['lam_rate'], g['lam_scale'], 1e-9, what='lam_scale')
    assert_close(tr['logml'], g['logml'][:, :, 2:4], LOGML_RTOL, what='logml')
    assert_close(tr['beta_mean'], g['beta_mean'][..., :n], 1e-7, 1e-9, what='beta_mean')
    assert_close(tr['beta_cov_diag'], g['beta_cov_diag'][..., :n], 1e-7, 1e-13, what='beta_cov_diag')
    ref_means = np.stack([g['mus_mean'][:, :, 2, :n], g['mus_mean'][:, :, 3, :n]], axis=2)
    assert_close(tr['mus_mean'], ref_means, 1e-7, 1e-9, what='mus_mean')
    ref_ld = np.stack([g['mus_logdens'][:, :, 2, 1], g['mus_logdens'][:, :, 3, 0]], axis=-1)
    assert_close(tr['mus_logdens'], ref_ld, 1e-8, 1e-8, what='mus_logdens')
    assert_close(tr['mu_after'], g['mu_after'][..., :n], 0, 0, what='mu_after')
    adj = parallel.credible_scores_from_counts(counts['edge'], counts['neg'], counts['nonempty'])
    assert_close(adj, g['proposed_adj_matrix'], 1e-14, what='proposed_adj_matrix (credible scores)')
    assert np.array_equal(counts['cp_hist'], g['cp_hist'].astype(np.int64))
    assert np.array_equal(counts['cp_kept'], g['cp_kept'].astype(np.int64))


@pytest.mark.parametrize('path', FP_PLAIN_FILES + FP_VV_FILES + FP_SEQ_FILES, ids=ids(FP_PLAIN_FILES + FP_VV_FILES + FP_SEQ_FILES))
def test_fp_plain_replay_vs_reference(eng_mod, path):
    """fp_varying_nh_dbn (fullParentsBpwLinReg.py:41-140) / fp_h_dbn (fpBayesianLinearRegression.py:40-125) replayed from
    the reference's recorded draws: changepoint lists and accept flags bit-exact, log-MLs 1e-9, Gibbs parameters, and the
    credible scores of the thinned beta chain (network.py:316-347) from the device's sign counts."""
    from dynamicbayesiannetworks_b200 import parallel
    g = load(path)
    method = str(g['method'])
    n, n_iter = g['sigma2'].shape
    burn_in = int(g['burn_in'])
    replay = dict(sigma2=g['sigma2'], lambda2=g['lambda2'], u_tau=g['u_tau'], beta=g['beta'], tau_move=g['tau_move'],
                  tau_pick=g['tau_pick'])
    vv = method == 'fp_var_glob_coup_nh_dbn'      # fpvvGlobCoup.py: per-segment variances, mu redrawn every iteration
    if vv:
        replay.update(sigma2v=g['sigma2v'], mu_tau=g['mus_val'][:, :, 0, :])
    seq = method == 'fp_seq_coup_nh_dbn'          # fpSeqCoupBpwlinReg.py: beta-tilde prior means, delta^2
    if seq:
        replay.update(delta2=g['delta2'])
    with eng_mod(n_chains=1) as e:
        e.build_stats(data_list(g))
        e.chain_init(method, burn_in=burn_in)
        tr = e.run_fp(n_iter, replay=replay, trace=True)
        counts = e.counts()
    assert np.array_equal(tr['S'], g['S']) and np.array_equal(tr['tau_after'], g['tau_after'])
    if method != 'fp_h_dbn':
        for key in ('tau_valid', 'tau_accept'):
            assert np.array_equal(tr[key], g[key]), key
    assert_close(tr['sig_shape'], g['sig_shape'], 1e-12, what='sig_shape')
    assert_close(1.0 / tr['sig_rate'], g['sig_scale'], 1e-9, what='sig_scale')
    assert_close(tr['lam_shape'], g['lam_shape'], 1e-12, what='lam_shape')
    assert_close(1.0 / tr['lam_rate'], g['lam_scale'], 1e-9, what='lam_scale')
    assert_close(tr['logml'], g['logml'][:, :, 2:4], LOGML_RTOL, what='logml')
    if seq:
        # the beta-tilde recursion chains n-wide solves on the near-collinear mTOR series: single entries that are ~1e-9 of
        # their vector carry the reference's own inv() noise, so entries are compared at 1e-7 of the vector's scale
        ref = g['beta_mean'][..., :n]
        fin = np.isfinite(ref)
        assert np.array_equal(np.isfinite(tr['beta_mean']), fin)
        scale = np.max(np.where(fin, np.abs(ref), 0.0), axis=-1, keepdims=True)
        excess = np.where(fin, np.abs(np.where(fin, tr['beta_mean'], 0.0) - np.where(fin, ref, 0.0)) - (1e-7 * scale + 1e-9), -1.0)
        assert excess.max() <= 0, ('beta_mean: |a - b| <= 1e-7 max|b| + 1e-9 violated by', float(excess.max()))
    else:
        assert_close(tr['beta_mean'], g['beta_mean'][..., :n], 1e-7, 1e-9, what='beta_mean')
    assert_close(tr['beta_cov_diag'], g['beta_cov_diag'][..., :n], 1e-7, 1e-13, what='beta_cov_diag')
    if seq:      # delta^2's shape / rate travel in the first per-segment variance slots
        assert_close(tr['sig_shape_v'][:, :, 0], g['del_shape'], 1e-12, what='del_shape')
        assert_close(1.0 / tr['sig_rate_v'][:, :, 0], g['del_scale'], 1e-8, what='del_scale')
    if vv:
        assert_close(tr['sig_shape_v'], g['sig_shape_v'], 1e-12, what='sig_shape_v')
        assert_close(1.0 / tr['sig_rate_v'], g['sig_scale_v'], 1e-9, what='sig_scale_v')
        assert_close(tr['mus_mean'][:, :, 0], g['mus_mean'][:, :, 0, :n], 1e-7, 1e-9, what='vvMuSampler mean')
        assert_close(tr['mu_after'], g['mu_after'][..., :n], 0, 0, what='mu_after')
    adj = parallel.credible_scores_from_counts(counts['edge'], counts['neg'], counts['nonempty'])
    assert_close(adj, g['proposed_adj_matrix'], 1e-14, what='proposed_adj_matrix (credible scores of the beta chain)')
    if 'cp_hist' in g:
        assert np.array_equal(counts['cp_hist'], g['cp_hist'].astype(np.int64))
        assert np.array_equal(counts['cp_kept'], g['cp_kept'].astype(np.int64))


@pytest.mark.parametrize('method', ['fp_varying_nh_dbn', 'fp_h_dbn', 'fp_var_glob_coup_nh_dbn', 'fp_seq_coup_nh_dbn'])
def test_fp_plain_free_running_and_dropin(eng_mod, method):
    """Free-running Philox chains of the plain full-parents methods: the device's sign counts equal the credible scores
    recomputed on the host from the traced beta chain (same run), the posterior-mean betas agree with the oracle's
    free-running chain within Monte-Carlo error, and the Network drop-in returns the reference's result keys."""
    import pickle
    from dynamicbayesiannetworks_b200 import Network
    g = load([p for p in CHAIN_FILES if 'yeast_varying' in p][0])
    data = data_list(g)
    X, Y = O.lagged_matrices(data)
    n = X.shape[1]
    net = Network(data, 400, 100, 1, keep_chain=True)
    net.infer_network(method)
    adj = np.array(net.proposed_adj_matrix)
    assert adj.shape == (n, n) and (adj >= 0.5 - 1e-12)[~np.eye(n, dtype=bool)].all() and (adj <= 1).all()
    assert np.allclose(np.diag(adj), 0)
    for node in range(n):
        r = net.chain_results_per_node[node]
        assert len(r['betas_vector']) == 401 and len(r['sigma_sqr_vector']) == 401
        assert (len(r['tau_vector']) == 400) == (method != 'fp_h_dbn')
        burned = r['betas_vector'][2:][100:]                      # network.py:316-320
        samples = np.array([v for x in range(len(burned)) if x % 10 == 0 for v in burned[x]])
        row = O.credible_scores(samples, [j for j in range(n) if j != node], n)
        assert_close(adj[node], row, 1e-14, what='credible scores node %d' % node)
    pickle.loads(pickle.dumps(net))
    # posterior mean of beta (segments pooled) vs the oracle's free-running chains
    node = 0
    mine = np.array([v for it in range(100, 400) for v in net.chain_results_per_node[node]['betas_vector'][it + 1]]).mean(axis=0)
    rng = np.random.default_rng(3)
    ref = []
    for _ in range(4):
        tr = O.run_chain(method, X, Y, node, 300, O.RngDraws(rng))
        ref += [tr['beta'][it, h, :n] for it in range(100, 300) for h in range(tr['S'][it])]
    ref = np.array(ref).mean(axis=0)
    assert np.abs(mine - ref).max() < 0.25, (mine, ref)


@pytest.mark.parametrize('n,T,n_iter', [(72, 160, 6), (300, 420, 3)])
def test_fp_glob_large_n_workspace_path(eng_mod, n, T, n_iter):
    """n = 72 / 300 do not fit the shared-memory matrices: the global-workspace path (blocked Cholesky / triangular
    inverse / rank-k update; n = 300 takes two column passes and several tiles) must give the same log-MLs as the oracle
    on the same injected stream (free draws from the oracle's own chain)."""
    from dynamicbayesiannetworks_b200.synth import make_synthetic
    data, _, _ = make_synthetic(n, T, n_changepoints=2, max_fan_in=3, seed=5)
    X, Y = O.lagged_matrices([data])
    # only the first 4 nodes are checked; the other units replay node 0's stream (their results are not compared).
    # the oracle's RngDraws do not record their picks: wrap them in a recording draw source
    class Rec(O.RngDraws):
        def __init__(self, rng):
            super().__init__(rng); self.log = {}
        def move(self, it, which):
            v = super().move(it, which); self.log[(it, 'move')] = v; return v
        def picker(self, it, which):
            inner = super().picker(it, which); lst = self.log.setdefault((it, 'pick'), [])
            def f(m):
                v = inner(m); lst.append(v); return v
            return f
        def uniform(self, it, which):
            v = super().uniform(it, which); self.log[(it, 'u')] = v; return v
        def mu_draw(self, it, slot, mean, Sigma):
            v = super().mu_draw(it, slot, mean, Sigma); self.log[(it, 'mu', slot)] = v; return v
    recs, trs = [], []
    for node in range(4):
        r = Rec(np.random.default_rng(100 + node))
        trs.append(O.run_chain('fp_glob_coup_nh_dbn', X, Y, node, n_iter, r)); recs.append(r)
    rp = dict(sigma2=np.ones((n, n_iter)), lambda2=np.ones((n, n_iter)), u_tau=np.full((n, n_iter), 0.5),
              mu_tau=np.zeros((n, n_iter, n)), beta=np.zeros((n, n_iter, 10, n)),
              tau_move=np.zeros((n, n_iter), dtype=np.int32), tau_pick=np.zeros((n, n_iter, 2), dtype=np.int32))
    for u in range(n):
        t, r = (trs[u], recs[u]) if u < 4 else (trs[0], recs[0])
        rp['sigma2'][u], rp['lambda2'][u] = t['sigma2'], t['lambda2']
        rp['beta'][u] = np.nan_to_num(t['beta'][:, :, :n])
        for it in range(n_iter):
            rp['tau_move'][u, it] = r.log[(it, 'move')]
            picks = r.log.get((it, 'pick'), [])
            rp['tau_pick'][u, it, :len(picks)] = picks
            rp['u_tau'][u, it] = r.log.get((it, 'u'), 0.5)
            if (it, 'mu', 3) in r.log:
                rp['mu_tau'][u, it] = r.log[(it, 'mu', 3)]
    with eng_mod(n_chains=1) as e:
        e.build_stats([data])
        e.chain_init('fp_glob_coup_nh_dbn', burn_in=0)
        tr = e.run_fp(n_iter, replay=rp, trace=True)
    for node in range(4):
        assert np.array_equal(tr['tau_after'][node], trs[node]['tau_after']), node
        assert_close(tr['logml'][node], trs[node]['logml'][:, 2:4], LOGML_RTOL, what='logml node %d' % node)
        assert_close(1.0 / tr['sig_rate'][node], trs[node]['sig_scale'], 1e-9, what='sig_scale')
        assert_close(tr['mu_after'][node], trs[node]['mu_after'][:, :n], 0, 0, what='mu_after')


def test_fp_glob_network_dropin(eng_mod):
    import pickle
    from dynamicbayesiannetworks_b200 import Network
    g = load([p for p in CHAIN_FILES if 'yeast_varying' in p][0])
    net = Network(data_list(g), 200, 50, 1)
    net.infer_network('fp_glob_coup_nh_dbn')
    adj = np.array(net.proposed_adj_matrix)
    assert adj.shape == (5, 5) and (adj >= 0.5 - 1e-12)[~np.eye(5, dtype=bool)].all() and (adj <= 1).all()
    assert np.allclose(np.diag(adj), 0)
    r = net.chain_results
    assert len(r['mu_vector']) == 201 and len(r['tau_vector']) == 200 and r['pi_vector'] == [[0, 1, 2, 3]]
    assert net.network_args['type'] == 'Full Parents'
    pickle.loads(pickle.dumps(net))
    many = Network(data_list(g), 300, 100, 1, n_chains=32)
    many.infer_network('fp_glob_coup_nh_dbn')
    assert np.array(many.proposed_adj_matrix).shape == (5, 5)


# ---------------------------------------------------------------------------------------------------------
# BASELINE configs[3] at full size (100 genes, T = 2000, 1024 chains, fan-in 4): size-independent properties
# ---------------------------------------------------------------------------------------------------------
def test_c4_full_size_properties(eng_mod):
    from dynamicbayesiannetworks_b200.synth import make_synthetic
    n, T, C, fan = 100, 2000, 1024, 4
    data, _, _ = make_synthetic(n, T, 3, 4, 20260119)
    X, Y = O.lagged_matrices([data])
    N = X.shape[0]
    n_iter, burn, thin = 60, 20, 10
    with eng_mod(n_chains=C, seed=5) as e:
        e.build_stats([data])
        e.chain_init('varying_nh_dbn', fan_in=fan, burn_in=burn, thin=thin)
        for w in (7, 13, 40):                     # launches of different lengths carry the state and the cache on
            e.run(w)
            assert e.verify_cache() == (0, 0)
        st = e.state()
        c = e.counts()
        ctr = e.counters()
        # the device's log-ML of states the chains actually reached == the oracle's (the reference overflows at this T)
        rng = np.random.default_rng(0)
        pick = rng.choice(n * C, size=48, replace=False)
        node = [int(u // C) for u in pick]
        pi = [[int(v) for v in st['pi'][u, :st['npi'][u]]] for u in pick]
        tau = [[int(v) for v in st['tau'][u, :st['S'][u]]] for u in pick]
        lam = st['lambda2'][pick]
        got = e.score_batch(node, pi, tau, lam, alpha=0.005, beta=0.005, T=N)
    U = n * C
    assert ctr['unit_iterations'] == U * n_iter
    assert np.isfinite(st['sigma2']).all() and (st['sigma2'] > 0).all()
    assert np.isfinite(st['lambda2']).all() and (st['lambda2'] > 0).all()
    # structural invariants of every state: sorted distinct changepoints in [2, N] ending with N + 2, <= 9 segments,
    # distinct parents != node, fan-in respected
    S = st['S']
    assert S.min() >= 1 and S.max() <= 9
    for u in rng.choice(U, size=4000, replace=False):
        t = st['tau'][u, :S[u]]
        assert t[-1] == N + 2 and (np.diff(t) > 0).all() and (S[u] == 1 or (t[0] >= 2 and t[-2] <= N)), (u, t)
        p = st['pi'][u, :st['npi'][u]]
        assert st['npi'][u] <= fan and len(set(p.tolist())) == p.size and (u // C) not in p.tolist() and (p >= 0).all() and (p < n).all()
    # count bookkeeping (network.py:358-359, :388-389): kept tau samples per node = C * #{it in [burn, n_iter): (it - burn) % thin == 0}
    kept_tau = sum(1 for it in range(burn, n_iter) if (it - burn) % thin == 0)
    assert (c['cp_kept'] == C * kept_tau).all()
    kept_pi = sum(1 for idx in range(burn, n_iter + 1) if (idx - burn) % thin != 0)
    assert (c['nonempty'] <= C * kept_pi).all() and (c['edge'].sum(axis=1) <= fan * c['nonempty']).all()
    assert (np.diag(c['edge']) == 0).all()
    assert (c['cp_hist'][:, :2] == 0).all() and (c['cp_hist'][:, N + 1:] == 0).all()
    assert (c['cp_hist'].sum(axis=1) <= 8 * c['cp_kept']).all()
    for b in range(len(pick)):
        stt = O.segment_stats(X, Y[:, node[b]], pi[b], tau[b])
        ref = O.log_ml(stt, [np.zeros(len(pi[b]) + 1)] * len(tau[b]), 0.005, 0.005, lam[b], N)
        assert abs(got[b] - ref) <= LOGML_RTOL * abs(ref), (b, got[b], ref)


# ---------------------------------------------------------------------------------------------------------
# multi-GPU: chains sharded over two ranks + NCCL all-reduce of the counts == the unsharded run (needs 2 GPUs)
# ---------------------------------------------------------------------------------------------------------
# ---------------------------------------------------------------------------------------------------------
# BASELINE configs[4] at full size (500 genes, T = 10000, full-parents globally coupled): the blocked k = 500 routines
# ---------------------------------------------------------------------------------------------------------
def test_c5_full_size_logml_and_state(eng_mod):
    """One traced iteration of every node at n = 500, T = 10000 (30 GB prefix table, design width 500): the device's
    log marginal likelihoods of the current and of the accepted proposed state equal the oracle's (1e-9), the Gibbs
    rates too, and the state invariants hold for all 500 units."""
    from dynamicbayesiannetworks_b200.synth import make_synthetic
    n, T = 500, 10000
    data, cps, _ = make_synthetic(n, T, n_changepoints=5, max_fan_in=4, seed=20260120)
    X, Y = O.lagged_matrices([data])
    N = X.shape[0]
    tau0 = [c + 1 for c in cps] + [N + 2]
    with eng_mod(n_chains=1, seed=3) as e:
        e.build_stats([data])
        zz, zy, yy = e.stats_row(N)
        e.chain_init('fp_glob_coup_nh_dbn')
        e.set_changepoints(tau0)
        tr = e.run_fp(1, trace=True)
        st = e.state()
    # the last prefix row is the full-series Gram matrix
    Z = np.concatenate([np.ones((N, 1)), X], axis=1)
    assert_close(zz, Z.T @ Z, 1e-11, 1e-9, what='ZZ[N]')
    assert_close(yy, (Y * Y).sum(axis=0), 1e-11, what='YY[N]')
    S_after = tr['S_after'][:, 0]
    assert ((S_after >= 1) & (S_after <= 9)).all() and (tr['S'][:, 0] == len(tau0)).all()
    for u in range(n):
        t = tr['tau_after'][u, 0, :S_after[u]]
        assert t[-1] == N + 2 and (np.diff(t) > 0).all() and t[0] >= 2
    assert np.isfinite(tr['sigma2']).all() and (tr['sigma2'] > 0).all() and (tr['lambda2'] > 0).all()
    assert np.array_equal(st['S'], S_after)
    valid = tr['tau_valid'][:, 0] == 1
    assert valid.sum() > n // 2 and np.isfinite(tr['logml'][valid, 0]).all()
    c = O.method_constants('fp_glob_coup_nh_dbn', N)
    acc_nodes = [u for u in range(n) if tr['tau_accept'][u, 0] == 1][:2]
    for node in sorted(set([0, 250, 499] + acc_nodes)):
        y = Y[:, node]
        pi = [j for j in range(n) if j != node]
        stats = O.segment_stats(X, y, pi, tau0)
        zeros = [np.zeros(n)] * len(stats)
        shape, rate = O.sigma2_params(stats, zeros, 1.0, c['a_sig'], c['b_sig'], c['T_sigma'])
        assert_close(tr['sig_rate'][node, 0], rate, 1e-9, what='sigma^2 rate node %d' % node)
        if valid[node]:
            lam = float(tr['lambda2'][node, 0])
            ref = O.log_ml(stats, zeros, c['a_sig'], c['b_sig'], lam, c['T_ml'])
            assert_close(tr['logml'][node, 0, 0], ref, LOGML_RTOL, what='log-ML(tau) node %d' % node)
            if tr['tau_accept'][node, 0] == 1:
                tau1 = [int(v) for v in tr['tau_after'][node, 0, :S_after[node]]]
                mu1 = tr['mu_after'][node, 0, :n]
                stats1 = O.segment_stats(X, y, pi, tau1)
                ref = O.log_ml(stats1, [mu1] * len(stats1), c['a_sig'], c['b_sig'], lam, c['T_ml'])
                assert_close(tr['logml'][node, 0, 1], ref, LOGML_RTOL, what='log-ML(tau*) node %d' % node)


def test_network_sharded_over_two_gpus_matches_unsharded():
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    from helpers import ROOT
    import os
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29571', os.path.join(ROOT, 'scripts', 'dist_network_check.py')]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:]
    assert res.stdout.count('sharded == unsharded: True') == 22, res.stdout[-3000:]   # 11 cases x 2 ranks
