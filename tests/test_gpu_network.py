"""GPU tests of the drop-in surface added in round 2: `Network.fit` / `Network.score_edges` (network.py:143,283), the
flat (node, chain) unit range behind them (`dbn_set_unit_range`), `padded_betas` / `betas_over_time` /
`scores_over_time` of a device-run chain, bounded trace windows."""
import numpy as np
import pytest

from helpers import CHAIN_FILES, LAG_CHAIN_FILES, LOGML_RTOL, load, data_list, ids, assert_close, assert_close_mat, golden_replay, labels_to_columns
from oracle import dbn_oracle as O

pytestmark = pytest.mark.gpu


def _series(name):
    return data_list(load([p for p in CHAIN_FILES if ('%s_varying' % name) in p][0]))


@pytest.mark.parametrize('method', ['varying_nh_dbn', 'seq_coup_nh_dbn', 'fp_glob_coup_nh_dbn', 'fp_varying_nh_dbn'])
def test_unit_ranges_reproduce_the_full_grid(method):
    """Three engines on the slices [0, 7), [7, 8), [8, U) of the unit grid: their states are the matching slices of the
    full grid's state and their counts add up to the full grid's counts, bit for bit."""
    from dynamicbayesiannetworks_b200.engine import DbnEngine
    data = _series('yeast')
    C, n_iter = 3, 60
    fp = method.startswith('fp_')

    def run(rng):
        with DbnEngine(n_chains=C, seed=9, unit_range=rng) as e:
            e.build_stats(data)
            e.chain_init(method, burn_in=10)
            (e.run_fp if fp else e.run)(n_iter)
            return e.state(), e.counts()
    full_st, full_c = run(None)
    U = 5 * C
    acc = None
    for first, count in ((0, 7), (7, 1), (8, U - 8)):
        st, c = run((first, count))
        for key in ('S', 'sigma2', 'lambda2') + (() if fp else ('npi',)):
            assert np.array_equal(st[key], full_st[key][first:first + count]), (method, key, first)
        for u in range(count):      # entries past the list lengths are not part of the state
            assert np.array_equal(st['tau'][u, :st['S'][u]], full_st['tau'][first + u, :st['S'][u]]), (method, 'tau', first + u)
            if not fp:
                assert np.array_equal(st['pi'][u, :st['npi'][u]], full_st['pi'][first + u, :st['npi'][u]]), (method, 'pi', first + u)
        acc = c if acc is None else {k: acc[k] + c[k] for k in c}
    for k in full_c:
        assert np.array_equal(acc[k], full_c[k]), (method, k)


def test_unit_range_table_holds_only_its_nodes():
    """A handle restricted to a unit range builds the ZZ block and the ZY | YY blocks of the range's nodes only (a rank
    keeps its own columns): rows equal the matching slices of the full table's rows, the scorer works on those nodes and refuses
    the others, and moving the range afterwards rebuilds the table for the new nodes from the series on the device."""
    from dynamicbayesiannetworks_b200.engine import DbnEngine, DbnError
    data = _series('yeast')
    n, C = data[0].shape[1], 3
    with DbnEngine(n_chains=C) as full:
        full.build_stats(data)
        N = full.N
        info_full = full.stats_info()
        rows_full = {t: full.stats_row(t) for t in (0, 1, N // 2, N)}
        ref = full.score_batch([2], [[0, 4]], [[N // 2, N + 2]], [1.3])
    with DbnEngine(n_chains=C, unit_range=(4, 7)) as e:      # units 4..10 = nodes 1..3
        e.build_stats(data)
        info = e.stats_info()
        assert info['row_doubles'] == (n + 1) * (n + 2) // 2 + 3 * (n + 2)
        assert info['row_doubles'] < info_full['row_doubles'] and info['algorithmic_bytes'] < info_full['algorithmic_bytes']
        for t, (zz, zy, yy) in rows_full.items():
            zz_, zy_, yy_ = e.stats_row(t)
            assert np.array_equal(zz_, zz) and np.array_equal(zy_, zy[1:4]) and np.array_equal(yy_, yy[1:4]), t
        assert np.array_equal(e.score_batch([2], [[0, 4]], [[N // 2, N + 2]], [1.3]), ref)
        with pytest.raises(DbnError, match='holds nodes 1..3 only'):
            e.score_batch([0], [[1, 4]], [[N // 2, N + 2]], [1.3])
        # the range moves after the build: nodes 3..4 now
        e.unit_range = (3 * C + 1, 4)
        e.chain_init('varying_nh_dbn', burn_in=0)
        assert e.stats_info()['row_doubles'] == (n + 1) * (n + 2) // 2 + 2 * (n + 2)
        for t, (zz, zy, yy) in rows_full.items():
            zz_, zy_, yy_ = e.stats_row(t)
            assert np.array_equal(zz_, zz) and np.array_equal(zy_, zy[3:5]) and np.array_equal(yy_, yy[3:5]), t
        e.run(40)
        assert e.verify_cache() == (0, 0)
        st = e.state()
    with DbnEngine(n_chains=C) as full:
        full.build_stats(data)
        full.chain_init('varying_nh_dbn', burn_in=0)
        full.run(40)
        fs = full.state()
    for key in ('S', 'npi', 'sigma2', 'lambda2'):
        assert np.array_equal(st[key], fs[key][3 * C + 1:3 * C + 5]), key


@pytest.mark.parametrize('method', ['h_dbn', 'varying_nh_dbn', 'glob_coup_nh_dbn', 'fp_varying_nh_dbn', 'fp_glob_coup_nh_dbn'])
def test_fit_score_edges_sequence_equals_infer_network(method):
    """The reference's own calling sequence (network.py:413-416) node by node gives what infer_network gives for all
    nodes at once: adjacency rows, changepoint chains, histories (same Philox streams, keyed by (chain, node))."""
    from dynamicbayesiannetworks_b200 import Network
    data = _series('yeast')
    n = data[0].shape[1]
    a = Network(data, 220, 60, 1)
    a.infer_network(method)
    b = Network(data, 220, 60, 1)
    for i in range(n):
        b.set_network_configuration(i)
        b.fit(method)
        b.score_edges(i, method)
        for key in a.chain_results_per_node[i]:
            x, y = a.chain_results_per_node[i][key], b.chain_results[key]
            assert len(x) == len(y), (method, i, key)
    assert np.array_equal(np.array(a.proposed_adj_matrix), np.array(b.proposed_adj_matrix))
    assert a.cps_over_response == b.cps_over_response
    assert a.network_args['scoring_method'] == b.network_args['scoring_method']
    if method != 'fp_h_dbn':
        assert len(a.betas_over_time) == n and len(b.betas_over_time) == n
        for i in range(n):
            assert np.array_equal(np.stack(a.betas_over_time[i]), np.stack(b.betas_over_time[i]))
    with pytest.raises(RuntimeError):
        Network(data, 10, 0, 1).fit(method)                       # no configuration set


def test_chain_results_keys_and_diagnostics_shapes():
    from dynamicbayesiannetworks_b200 import Network
    data = _series('yeast')
    n, N = 5, 33
    net = Network(data, 300, 100, 1)
    net.infer_network('varying_nh_dbn')
    r = net.chain_results
    assert len(r['padded_betas']) == 301 and r['padded_betas'][0][0].shape == (n,)
    for it in range(1, 301):       # one padded vector per segment of the draw; zeros exactly where the gene is no parent
        pi = set(int(p) for p in r['pi_vector'][it - 1])
        assert len(r['padded_betas'][it]) == len(r['betas_vector'][it])
        for v in r['padded_betas'][it]:
            assert v.shape == (n,) and all(v[d] == 0 for d in range(1, n) if d not in pi)
    assert len(net.betas_over_time) == n and len(net.betas_over_time[0]) == N and net.betas_over_time[0][0].shape == (20, n)
    assert net.scores_over_time == []                   # varying-parents methods have none (network.py:370-396)
    fp = Network(data, 300, 100, 1)
    fp.infer_network('fp_varying_nh_dbn')
    assert len(fp.scores_over_time) == n and fp.scores_over_time[0].shape == (N, n - 1)
    assert (np.abs(fp.scores_over_time[0]) <= 0.5).all()
    h = Network(data, 300, 100, 1)
    h.infer_network('h_dbn')
    assert all(c == [N + 2] for c in h.cps_over_response[0]) and len(h.cps_over_response) == n     # network.py:376-383


def test_kept_history_is_traced_in_bounded_windows():
    """keep_chain runs are traced window by window (one launch for the whole chain allocated
    units x iterations x 3 KB on the device): a small window gives the same histories as one launch."""
    from dynamicbayesiannetworks_b200 import Network, network as NW
    data = _series('yeast')
    a = Network(data, 130, 20, 1, window=130)
    a.infer_network('seq_coup_nh_dbn')
    b = Network(data, 130, 20, 1, window=17)
    b.infer_network('seq_coup_nh_dbn')
    assert np.array_equal(np.array(a.proposed_adj_matrix), np.array(b.proposed_adj_matrix))
    assert a.chain_results['tau_vector'] == b.chain_results['tau_vector']
    assert a.chain_results['sigma_sqr_vector'] == b.chain_results['sigma_sqr_vector']
    saved = NW.TRACE_WINDOW_BYTES
    try:
        NW.TRACE_WINDOW_BYTES = 64 << 10           # forces several windows through the default planner
        c = Network(data, 130, 20, 1)
        keep, window = c._plan(2, 5, 5, 130, 1)
        assert keep and 1 <= window < 130
        c.infer_network('seq_coup_nh_dbn')
        assert c.chain_results['tau_vector'] == a.chain_results['tau_vector']
    finally:
        NW.TRACE_WINDOW_BYTES = saved
    with pytest.raises(MemoryError):                # a history that cannot fit is refused before anything is allocated
        Network(data, 10 ** 9, 0, 1, keep_chain=True)._plan(1, 5, 5, 10 ** 9, 1)


# ---------------------------------------------------------------------------------------------------------
# lag = 2 (network.py:92-122): zero-padded lagged copies as extra feature columns
# ---------------------------------------------------------------------------------------------------------
def test_lag2_prefix_rows():
    from dynamicbayesiannetworks_b200.engine import DbnEngine
    data = _series('yeast')                   # two series segments: the lag-2 block restarts with a zero row in each
    X, Y = O.lagged_matrices(data, 2)
    N, nv = X.shape
    Z = np.concatenate([np.ones((N, 1)), X], axis=1)
    with DbnEngine(n_chains=1) as e:
        e.build_stats(data, lag=2)
        assert e.stats_info()['N'] == N
        for t in (0, 1, 2, 17, 18, 19, N):
            zz, zy, yy = e.stats_row(t)
            assert_close(zz, Z[:t].T @ Z[:t], 1e-12, 1e-12, what='ZZ[%d]' % t)
            assert_close(zy, (Z[:t].T @ Y[:t]).T, 1e-12, 1e-12, what='ZY[%d]' % t)
            assert_close(yy, (Y[:t] ** 2).sum(axis=0), 1e-12, 1e-12, what='YY[%d]' % t)


@pytest.mark.parametrize('path', LAG_CHAIN_FILES, ids=ids(LAG_CHAIN_FILES))
def test_lag2_chain_replay_vs_reference(path):
    """The reference's recorded lag-2 chains replayed on the device: parents (as feature columns), changepoints and accept
    flags bit-exact, log-MLs and Gibbs parameters 1e-9, adjacency rows 1e-14."""
    from dynamicbayesiannetworks_b200.engine import DbnEngine
    g = load(path)
    method, lag = str(g['method']), int(g['lag'])
    data = data_list(g)
    n, n_iter = g['sigma2'].shape
    cols = labels_to_columns(g['pi_after'], n)
    with DbnEngine(n_chains=1) as e:
        e.build_stats(data, lag=lag)
        e.chain_init(method, burn_in=int(g['burn_in']))
        tr = e.run(n_iter, replay=golden_replay(g), trace=True)
        counts = e.counts()
        assert e.verify_cache() == (0, 0)
    for key in ('pi_valid', 'pi_accept', 'npi', 'S'):
        assert np.array_equal(tr[key], g[key]), key
    assert np.array_equal(tr['pi_after'], cols), 'pi_after'
    assert np.array_equal(tr['tau_after'], g['tau_after']), 'tau_after'
    if method != 'h_dbn':
        for key in ('tau_valid', 'tau_accept'):
            assert np.array_equal(tr[key], g[key]), key
    assert_close(tr['logml'], g['logml'], LOGML_RTOL, what='logml')
    assert_close(1.0 / tr['sig_rate'], g['sig_scale'], 1e-9, what='sig_scale')
    assert_close(1.0 / tr['lam_rate'], g['lam_scale'], 1e-9, what='lam_scale')
    assert_close(tr['beta_mean'], g['beta_mean'], 1e-8, 1e-11, what='beta_mean')
    assert_close_mat(tr['beta_cov'], g['beta_cov'], 1e-9, what='beta_cov')
    adj = counts['edge'] / np.maximum(counts['nonempty'], 1)[:, None]
    ref = g['proposed_adj_matrix']
    ok = np.isfinite(ref).all(axis=1)
    assert_close(adj[ok], ref[ok][:, :n], 1e-14, what='proposed_adj_matrix')
    if 'cp_hist' in g:
        assert np.array_equal(counts['cp_hist'], g['cp_hist'].astype(np.int64))


def test_lag2_network_dropin():
    from dynamicbayesiannetworks_b200 import Network
    data = _series('yeast')
    n = 5
    net = Network(data, 300, 100, 2)
    net.infer_network('varying_nh_dbn')
    adj = np.array(net.proposed_adj_matrix)
    assert adj.shape == (n, 2 * (n - 1) + 1) and (adj[:, n:] == 0).all() and (adj >= 0).all() and (adj <= 1).all()
    r = net.chain_results
    labels = np.concatenate([np.asarray(p).ravel() for p in r['pi_vector'][1:]])
    assert labels.max() >= n and labels.max() <= n + (n - 1) - 1            # lag-2 labels n .. 2n - 2 (network.py:92-104)
    assert r['padded_betas'][5][0].shape == (2 * (n - 1) + 1,)
    assert len(net.betas_over_time) == n and net.betas_over_time[0][0].shape[1] == 2 * (n - 1) + 1
    with pytest.raises(NotImplementedError):
        Network(data, 10, 0, 2).infer_network('fp_glob_coup_nh_dbn')
    with pytest.raises(NotImplementedError):
        Network(data, 10, 0, 3).infer_network('varying_nh_dbn')


def test_network_two_handles_per_gpu_same_result(monkeypatch):
    """Runs larger than one wave of resident blocks are split over two handles (two streams) per GPU: the adjacency matrix and
    the changepoint histogram must be exactly those of the single-handle run."""
    from dynamicbayesiannetworks_b200 import Network
    from dynamicbayesiannetworks_b200 import network as NW
    from dynamicbayesiannetworks_b200.synth import make_synthetic
    data, _, _ = make_synthetic(60, 300, 2, 3, 7)
    chains = NW.ONE_WAVE_UNITS // 60 + 8          # a little more than one wave of units
    res = {}
    for label, limit in (('two', NW.ONE_WAVE_UNITS), ('one', 1 << 40)):
        monkeypatch.setattr(NW, 'ONE_WAVE_UNITS', limit)
        net = Network([data], 40, 10, 1, n_chains=chains, keep_chain=False)
        net.infer_network('varying_nh_dbn')
        res[label] = (np.array(net.proposed_adj_matrix), net.counters['unit_iterations'], net.counters['ml_evals'])
    assert np.array_equal(res['two'][0], res['one'][0])
    assert res['two'][1:] == res['one'][1:]
