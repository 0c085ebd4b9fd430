"""Round-2 pins: the instantiations that are TIMED (fan-in 4 / KM = 5, non-trace builds of every kernel family) tied to the
oracle at trajectory level, the work counters against host-recomputed values, and the free-running draws by moments.

The reference cannot generate these fixtures (fan-in 3 is hard-coded, `globCoupBayesianPwLinReg.py:22`, and its log-ML
overflows for T >= 344, `marginalLikelihood.py:116`), so the draw stream is recorded from the oracle's own chain
(`oracle.run_chain(..., fan_in=4)`, itself pinned to the reference by tests/test_oracle_golden.py) and injected into
the device chain: accept flags, parent sets and changepoint lists bit-exact, log-MLs 1e-9 (moves.py:528-639, :129-236).
"""
import numpy as np
import pytest

from helpers import CHAIN_FILES, LOGML_RTOL, load, data_list, assert_close, assert_close_mat, recording_draws, oracle_replay
from oracle import dbn_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def eng_mod():
    from dynamicbayesiannetworks_b200.engine import DbnEngine
    return DbnEngine


@pytest.fixture(scope='module')
def c4_series():
    from dynamicbayesiannetworks_b200.synth import make_synthetic
    data, cps, _ = make_synthetic(100, 2000, 3, 4, 20260119)
    X, Y = O.lagged_matrices([data])
    return data, X, Y, cps


def _series(name):
    return data_list(load([p for p in CHAIN_FILES if ('%s_varying' % name) in p][0]))


def compare_with_oracle(tr, u, ref, method, what):
    """Device trace of unit `u` against the oracle trace `ref` of the same injected stream."""
    for key in ('pi_valid', 'pi_accept', 'npi', 'S'):
        assert np.array_equal(tr[key][u], ref[key]), (what, key, np.nonzero(tr[key][u] != ref[key])[0][:5])
    assert np.array_equal(tr['pi_after'][u], ref['pi_after']), (what, 'pi_after')
    if method != 'h_dbn':
        for key in ('tau_valid', 'tau_accept'):
            assert np.array_equal(tr[key][u], ref[key]), (what, key)
    assert np.array_equal(tr['tau_after'][u], ref['tau_after']), (what, 'tau_after')
    assert_close(tr['sig_shape'][u], ref['sig_shape'], 1e-12, what=what + ' sig_shape')
    assert_close(1.0 / tr['sig_rate'][u], ref['sig_scale'], 1e-9, what=what + ' sig_scale')
    assert_close(tr['lam_shape'][u], ref['lam_shape'], 1e-12, what=what + ' lam_shape')
    assert_close(1.0 / tr['lam_rate'][u], ref['lam_scale'], 1e-9, what=what + ' lam_scale')
    if method == 'seq_coup_nh_dbn':
        assert_close(tr['del_shape'][u], ref['del_shape'], 1e-12, what=what + ' del_shape')
        assert_close(1.0 / tr['del_rate'][u], ref['del_scale'], 1e-9, what=what + ' del_scale')
    assert_close(tr['logml'][u], ref['logml'], LOGML_RTOL, what=what + ' logml')
    assert_close(tr['beta_mean'][u], ref['beta_mean'], 1e-8, 1e-11, what=what + ' beta_mean')
    assert_close_mat(tr['beta_cov'][u], ref['beta_cov'], 1e-9, what=what + ' beta_cov')
    if method == 'var_glob_coup_nh_dbn':
        assert_close(tr['sig_shape_v'][u], ref['sig_shape_v'], 1e-12, what=what + ' sig_shape_v')
        assert_close(1.0 / tr['sig_rate_v'][u], ref['sig_scale_v'], 1e-9, what=what + ' sig_scale_v')
    if method in ('glob_coup_nh_dbn', 'var_glob_coup_nh_dbn'):
        assert_close(tr['mus_mean'][u], ref['mus_mean'], 1e-8, 1e-11, what=what + ' mus_mean')
        assert_close_mat(tr['mus_cov'][u], ref['mus_cov'], 1e-9, what=what + ' mus_cov')
        r = ref['mus_logdens']
        assert_close(tr['mus_logdens'][u], np.stack([r[:, 0, 0], r[:, 1, 1], r[:, 2, 1], r[:, 3, 0]], axis=-1), 1e-8, 1e-9,
                     what=what + ' mus_logdens')
        assert_close(tr['mu_after'][u], ref['mu_after'], 0, 0, what=what + ' mu_after')


# ---------------------------------------------------------------------------------------------------------
# (a) fan-in 4 (KM = 5) replay of an oracle-recorded stream on the BASELINE configs[3] series: the instantiation
#     bench.py times (chain_fast_kernel<nh, 5, .>) and the KM = 5 builds of the other varying-parents kernels
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('method,start', [('varying_nh_dbn', 'empty'), ('varying_nh_dbn', 'true_cps'), ('h_dbn', 'empty'),
                                          ('seq_coup_nh_dbn', 'true_cps'), ('glob_coup_nh_dbn', 'empty'),
                                          ('var_glob_coup_nh_dbn', 'true_cps')])
def test_fan_in_4_replay_vs_oracle(eng_mod, c4_series, method, start):
    data, X, Y, cps = c4_series
    N, n = X.shape
    n_iter = 260 if method in ('varying_nh_dbn', 'h_dbn') else 160
    init_tau = None if start == 'empty' or method == 'h_dbn' else [c + 1 for c in cps] + [N + 2]
    nodes = (0, 17, 42, 58, 77, 99)
    chains = {}
    for node in nodes:
        rec = recording_draws(np.random.default_rng(4000 + node))
        chains[node] = (O.run_chain(method, X, Y, node, n_iter, rec, fan_in=4, init_tau=init_tau), rec)
    # the fixture is only worth something if the chains actually use the fifth column and move their changepoints
    assert sum(int((t['npi'] == 4).sum()) for t, _ in chains.values()) > n_iter // 2
    assert sum(int(t['pi_accept'].sum()) for t, _ in chains.values()) >= 20
    if method != 'h_dbn':
        assert sum(int(t['tau_accept'].sum()) for t, _ in chains.values()) >= 20
    with eng_mod(n_chains=1) as e:
        e.build_stats([data])
        e.chain_init(method, fan_in=4, burn_in=0)
        if init_tau is not None:
            e.set_changepoints(init_tau)
        tr = e.run(n_iter, replay=oracle_replay(n, n_iter, chains), trace=True)
        assert e.verify_cache() == (0, 0)
    for node in nodes:
        compare_with_oracle(tr, node, chains[node][0], method, '%s node %d' % (method, node))


# ---------------------------------------------------------------------------------------------------------
# (b) the non-trace build of every kernel family follows the trajectory of its trace build (same Philox stream =>
#     same discrete states; the trace builds are the ones the replay fixtures pin)
# ---------------------------------------------------------------------------------------------------------
def _check_state_against_trace(st, tr, it, with_pi, with_tau, what):
    S = tr['S_after'][:, it] if with_tau else st['S']
    assert np.array_equal(st['S'], S), (what, 'S')
    if with_pi:
        assert np.array_equal(st['npi'], tr['npi'][:, it]), (what, 'npi')
    for u in range(st['S'].size):
        if with_pi:
            assert np.array_equal(st['pi'][u, :st['npi'][u]], tr['pi_after'][u, it, :st['npi'][u]]), (what, 'pi', u)
        if with_tau:
            assert np.array_equal(st['tau'][u, :S[u]], tr['tau_after'][u, it, :S[u]]), (what, 'tau', u)
    assert np.allclose(st['sigma2'], tr['sigma2'][:, it], rtol=1e-9, atol=0), (what, 'sigma2')
    assert np.allclose(st['lambda2'], tr['lambda2'][:, it], rtol=1e-9, atol=0), (what, 'lambda2')


@pytest.mark.parametrize('method', ['varying_nh_dbn', 'h_dbn', 'seq_coup_nh_dbn', 'glob_coup_nh_dbn', 'var_glob_coup_nh_dbn'])
def test_plain_and_trace_builds_same_trajectory_fan_in_4(eng_mod, c4_series, method):
    """KM = 5 on the BASELINE configs[3] series: 3 chains x 100 nodes (ten warps, ragged last block)."""
    data = c4_series[0]
    C, windows = 3, [1, 9, 50, 60]
    with eng_mod(n_chains=C, seed=2026) as e:
        e.build_stats([data])
        e.chain_init(method, fan_in=4, burn_in=0)
        tr = e.run(sum(windows), trace=True)
    assert (tr['npi'] == 4).any()
    with eng_mod(n_chains=C, seed=2026) as e:
        e.build_stats([data])
        e.chain_init(method, fan_in=4, burn_in=0)
        done = 0
        for w in windows:
            e.run(w)
            done += w
            assert e.verify_cache() == (0, 0), done
            _check_state_against_trace(e.state(), tr, done - 1, True, method != 'h_dbn', '%s after %d' % (method, done))


@pytest.mark.parametrize('method', ['seq_coup_nh_dbn', 'glob_coup_nh_dbn', 'var_glob_coup_nh_dbn'])
def test_plain_and_trace_builds_same_trajectory_coupled(eng_mod, method):
    """KM = 4 (the reference's fan-in 3) on the mTOR series, 64 chains."""
    data = _series('mtor')
    C, windows = 64, [3, 40, 77]
    with eng_mod(n_chains=C, seed=91) as e:
        e.build_stats(data)
        e.chain_init(method, burn_in=0)
        tr = e.run(sum(windows), trace=True)
    with eng_mod(n_chains=C, seed=91) as e:
        e.build_stats(data)
        e.chain_init(method, burn_in=0)
        done = 0
        for w in windows:
            e.run(w)
            done += w
            _check_state_against_trace(e.state(), tr, done - 1, True, True, '%s after %d' % (method, done))


@pytest.mark.parametrize('method,series', [('fp_glob_coup_nh_dbn', 'mtor'), ('fp_varying_nh_dbn', 'mtor'), ('fp_h_dbn', 'mtor'),
                                           ('fp_var_glob_coup_nh_dbn', 'mtor'), ('fp_seq_coup_nh_dbn', 'yeast'),
                                           ('fp_glob_coup_nh_dbn', 'synth72'), ('fp_varying_nh_dbn', 'synth72')])
def test_plain_and_trace_builds_same_trajectory_full_parents(eng_mod, method, series):
    """fp_glob_kernel<false, ., .> vs <true, ., .>: shared-memory instantiation (mTOR / yeast) and the blocked
    global-workspace one (n = 72)."""
    if series == 'synth72':
        from dynamicbayesiannetworks_b200.synth import make_synthetic
        data, C, windows = [make_synthetic(72, 160, 2, 3, 5)[0]], 2, [2, 5]
    else:
        data, C, windows = _series(series), 16, [2, 11, 30]
    with eng_mod(n_chains=C, seed=17) as e:
        e.build_stats(data)
        e.chain_init(method, burn_in=0)
        tr = e.run_fp(sum(windows), trace=True)
    with eng_mod(n_chains=C, seed=17) as e:
        e.build_stats(data)
        e.chain_init(method, burn_in=0)
        done = 0
        for w in windows:
            e.run_fp(w)
            done += w
            _check_state_against_trace(e.state(), tr, done - 1, False, method != 'fp_h_dbn', '%s after %d' % (method, done))


# ---------------------------------------------------------------------------------------------------------
# (d) work counters: evaluations / segments / algorithmic bytes / flops recomputed on the host from the trace
# ---------------------------------------------------------------------------------------------------------
def _seg_bytes(k):
    return 16 * (k * (k + 1) // 2 + k + 1)


def _seg_flops(k):
    return (k * k * k) // 3 + 2 * k * k + 5 * k


@pytest.mark.parametrize('method', ['h_dbn', 'varying_nh_dbn'])
def test_work_counters_match_host_recount(eng_mod, method):
    """SURVEY 8d accounting of chain_fast_kernel: per iteration one Gibbs sweep (S segments at k), the parent-set sweep
    (current state: 1 evaluation of S segments at k whenever either move is valid; the proposal: 1 evaluation of S
    segments at k*), the changepoint move (1 evaluation of S* segments).  The trace and the plain build must report the
    same totals on the same stream, and both must equal the recount from the traced flags."""
    data = _series('yeast')
    C, n_iter = 64, 120
    with eng_mod(n_chains=C, seed=5) as e:
        e.build_stats(data)
        e.chain_init(method, burn_in=0)
        tr = e.run(n_iter, trace=True)
        ctr_trace = e.counters()
    with eng_mod(n_chains=C, seed=5) as e:
        e.build_stats(data)
        e.chain_init(method, burn_in=0)
        e.run(n_iter)
        ctr_plain = e.counters()
    U = tr['S'].shape[0]
    evals = segs = nbytes = flops = 0
    npi_before = np.concatenate([np.zeros((U, 1), dtype=np.int64), tr['npi'][:, :-1]], axis=1)
    for u in range(U):
        for it in range(n_iter):
            S, k = int(tr['S'][u, it]), int(npi_before[u, it]) + 1
            pv = tr['pi_valid'][u, it] == 1
            tv = method != 'h_dbn' and tr['tau_valid'][u, it] == 1
            segs += S; nbytes += S * _seg_bytes(k); flops += S * _seg_flops(k)            # Gibbs sweep
            if pv or tv:
                evals += 1; segs += S; nbytes += S * _seg_bytes(k); flops += S * _seg_flops(k) + 12
            if pv:
                move = int(tr['pi_move'][u, it])
                ks = k + (1 if move == 0 else (-1 if move == 1 else 0))
                evals += 1; segs += S; nbytes += S * _seg_bytes(ks); flops += S * _seg_flops(ks) + 12
            if tv:
                move = int(tr['tau_move'][u, it])
                Ss = S + (1 if move == 0 else (-1 if move == 1 else 0))
                k2 = int(tr['npi'][u, it]) + 1
                evals += 1; segs += Ss; nbytes += Ss * _seg_bytes(k2); flops += Ss * _seg_flops(k2) + 12
    want = dict(unit_iterations=U * n_iter, ml_evals=evals, segment_evals=segs, algorithmic_bytes=nbytes, algorithmic_flops=flops,
                pi_accepts=int((tr['pi_accept'] == 1).sum()),
                tau_accepts=0 if method == 'h_dbn' else int((tr['tau_accept'] == 1).sum()))
    for key, v in want.items():
        assert ctr_trace[key] == v, ('trace build', key, ctr_trace[key], v)
        assert ctr_plain[key] == v, ('plain build', key, ctr_plain[key], v)
    assert ctr_plain['segment_factorisations'] == ctr_trace['segment_factorisations'] > 0


# ---------------------------------------------------------------------------------------------------------
# (c) free-running chains: edge posteriors within Monte-Carlo error of the oracle's free-running chains, the error
#     estimated from the between-chain spread on both sides
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('method', ['h_dbn', 'varying_nh_dbn', 'seq_coup_nh_dbn', 'glob_coup_nh_dbn', 'var_glob_coup_nh_dbn'])
def test_free_running_edge_posteriors_within_mc_error(eng_mod, method):
    """Oracle side: tests/golden/free_running_yeast.npz (64 free-running oracle chains per node and method, generated by
    tests/golden/make_free_running.py).  Device side: 8 batches of 64 Philox chains with different seeds.  Every
    off-diagonal edge frequency and every node's mean number of changepoints must agree within 4.5 combined standard
    errors (between-chain / between-batch spread), and the z-scores must not be systematically large."""
    import os
    from helpers import GOLDEN
    f = load(os.path.join(GOLDEN, 'free_running_yeast.npz'))
    data = _series('yeast')
    n = data[0].shape[1]
    n_iter, burn = int(f['n_iter']), int(f['burn_in'])
    batches, cp_batches = [], []
    for b in range(8):
        with eng_mod(n_chains=64, seed=7000 + b) as e:
            e.build_stats(data)
            e.chain_init(method, burn_in=burn)
            e.run(n_iter)
            c = e.counts()
        batches.append(c['edge'] / np.maximum(c['nonempty'], 1)[:, None])
        cp_batches.append(c['cp_hist'].sum(axis=1) / np.maximum(c['cp_kept'], 1))
    gpu = np.mean(batches, axis=0)
    gpu_se = np.std(batches, axis=0, ddof=1) / np.sqrt(len(batches))
    per_chain = f['edge_freq_' + method]
    ref = per_chain.mean(axis=0)
    ref_se = per_chain.std(axis=0, ddof=1) / np.sqrt(per_chain.shape[0])
    off = ~np.eye(n, dtype=bool)
    se = np.sqrt(gpu_se ** 2 + ref_se ** 2)[off]
    z = np.abs(gpu - ref)[off] / np.maximum(se, 2e-3)
    assert z.max() < 4.5, (method, float(z.max()), gpu, ref)
    assert np.sqrt(np.mean(z ** 2)) < 2.0, (method, z)
    if method != 'h_dbn':
        # mean number of changepoints per kept sample and node (tau_vector thinning, network.py:388-389)
        g_cp = np.mean(cp_batches, axis=0)
        g_se = np.std(cp_batches, axis=0, ddof=1) / np.sqrt(len(cp_batches))
        r_all = f['cp_freq_' + method].sum(axis=2)
        r_cp, r_se = r_all.mean(axis=0), r_all.std(axis=0, ddof=1) / np.sqrt(r_all.shape[0])
        zc = np.abs(g_cp - r_cp) / np.maximum(np.sqrt(g_se ** 2 + r_se ** 2), 2e-3)
        assert zc.max() < 4.5, (method, zc, g_cp, r_cp)


# ---------------------------------------------------------------------------------------------------------
# free-running Gibbs draws by moments: from a FIXED state, one iteration of many chains; the sample moments of the
# device's own draws must match the Gibbs parameters the oracle computes for that state (samplers.py:103-127,
# :189-218, :265-294)
# ---------------------------------------------------------------------------------------------------------
def test_free_running_gibbs_draw_moments(eng_mod):
    data = _series('yeast')
    X, Y = O.lagged_matrices(data)
    N, n = X.shape
    C = 4096
    tau0 = [12, 23, N + 2]
    with eng_mod(n_chains=C, seed=31337) as e:
        e.build_stats(data)
        e.chain_init('fixed_nh_dbn', burn_in=0)       # no changepoint move; parents start empty: k = 1, S = 3
        e.set_changepoints(tau0)
        tr = e.run(1, trace=True)
    cst = O.method_constants('fixed_nh_dbn', N)
    for node in range(n):
        u = slice(node * C, (node + 1) * C)
        stats = O.segment_stats(X, Y[:, node], [], tau0)
        zeros = [np.zeros(1)] * len(stats)
        shape, rate = O.sigma2_params(stats, zeros, 1.0, cst['a_sig'], cst['b_sig'], cst['T_sigma'])
        # sigma^2 ~ InvGamma(shape, rate): 1 / sigma^2 ~ Gamma(shape, 1 / rate), mean shape / rate, var shape / rate^2
        prec = 1.0 / tr['sigma2'][u, 0]
        assert abs(prec.mean() - shape / rate) < 5 * np.sqrt(shape) / rate / np.sqrt(C), node
        assert abs(prec.var() / (shape / rate ** 2) - 1) < 0.15, node
        # beta_h | sigma^2 ~ N(mean_h, sigma^2 cov_h): standardised residuals are N(0, 1)
        for h, (mean, cov) in enumerate(O.beta_params(stats, zeros, 1.0, 1.0)):
            zs = (tr['beta'][u, 0, h, 0] - mean[0]) / np.sqrt(tr['sigma2'][u, 0] * cov[0, 0])
            assert abs(zs.mean()) < 5 / np.sqrt(C), (node, h, zs.mean())
            assert abs(zs.var() - 1) < 5 * np.sqrt(2.0 / C), (node, h, zs.var())
            assert abs(np.mean(zs ** 3)) < 5 * np.sqrt(15.0 / C) and abs(np.mean(zs ** 4) - 3) < 5 * np.sqrt(96.0 / C), (node, h)
        # lambda^2 | beta, sigma^2 ~ InvGamma(a + S k / 2, b + sum |beta_h|^2 / (2 sigma^2)): rate * (1 / lambda^2) ~ Gamma(shape, 1)
        lshape = tr['lam_shape'][u, 0]
        g = tr['lam_rate'][u, 0] / tr['lambda2'][u, 0]
        assert np.allclose(lshape, lshape[0])
        assert abs(g.mean() - lshape[0]) < 5 * np.sqrt(lshape[0] / C), node
        assert abs(g.var() / lshape[0] - 1) < 0.15, node


@pytest.mark.parametrize('method,series,chains', [('varying_nh_dbn', 'mtor', 64), ('h_dbn', 'yeast', 64), ('seq_coup_nh_dbn', 'mtor', 64),
                                                   ('glob_coup_nh_dbn', 'mtor', 64)])
def test_units_per_warp_do_not_change_the_chains(eng_mod, monkeypatch, method, series, chains):
    """Small unit counts are spread over more warps (DBN_LANES units per warp, chosen from the unit count): the Philox counters
    are keyed by the (chain, node) of a unit, not by its thread, so every mapping must give the same states and counts."""
    data = _series(series)
    got = {}
    for lanes in ('32', '5', '1', None):   # dense, a ragged value, one unit per warp, automatic
        if lanes is None:
            monkeypatch.delenv('DBN_LANES', raising=False)
        else:
            monkeypatch.setenv('DBN_LANES', lanes)
        with eng_mod(n_chains=chains, seed=77) as e:
            e.build_stats(data)
            e.chain_init(method, burn_in=0)
            e.run(30)
            e.run(45)
            st = e.state()
            cnt = e.counts()
            got[lanes] = (st['S'].copy(), st['tau'].copy(), st['npi'].copy(), st['pi'].copy(), st['lambda2'].copy(),
                          np.concatenate([np.asarray(v).ravel() for _, v in sorted(cnt.items())]))
    ref = got['32']
    for lanes, g in got.items():
        for a, b, name in zip(g, ref, ('S', 'tau', 'npi', 'pi', 'lambda2', 'counts')):
            assert np.array_equal(a, b), (method, lanes, name)


def test_score_batch_device_time(eng_mod):
    """dbn_score_batch keeps its device scratch between calls and reports the scorer kernel's device time."""
    data = _series('yeast')
    with eng_mod(n_chains=1) as e:
        e.build_stats(data)
        with pytest.raises(Exception):
            e.score_batch_ms()   # no call yet
        for B in (3, 700, 5):      # the scratch grows and is reused
            node = [i % e.n for i in range(B)]
            pi = [[(nd + 1) % e.n] for nd in node]
            tau = [[e.N + 2]] * B
            out = e.score_batch(node, pi, tau, 1.0)
            assert np.isfinite(out).all() and e.score_batch_ms() > 0.0


@pytest.mark.parametrize('shape', [(100, 2000), (20, 3001), (7, 130)], ids=lambda s: 'n%d_T%d' % s)
def test_one_pass_prefix_build_equals_three_pass_build(eng_mod, shape, monkeypatch):
    """The one-launch prefix build (DBN_PREFIX_ONEPASS: chunk offsets by decoupled look-back inside the write pass; measured slower
    than the default and kept as an alternative) writes the same bits as the three-launch build (totals, scan, write): both add chunk aggregates to the left in chunk order.  63 / 94 / 5 chunks: two,
    three and one look-back windows; with a unit range (fewer column tiles) as well; rebuilt several times (generation stamps)."""
    from dynamicbayesiannetworks_b200.synth import make_synthetic
    n, T = shape
    data, _, _ = make_synthetic(n, T, 3, min(4, n - 1), 7)
    X, Y = O.lagged_matrices([data])
    N = X.shape[0]
    Z = np.concatenate([np.ones((N, 1)), X], axis=1)
    ts = sorted({0, 1, 31, 32, 33, 64, N // 2, N - 33, N - 1, N} & set(range(N + 1)))
    for rng in (None, (n + 1, 2 * n - 1)):      # unit range: 3 chains per node, units n + 1 .. 3 n - 1 = nodes (n + 1) // 3 .. n - 1
        monkeypatch.setenv('DBN_PREFIX_ONEPASS', '1')      # read when the table is laid out
        with eng_mod(n_chains=3, unit_range=rng) as e:
            e.build_stats([data])
            for _ in range(3):
                e.rebuild_stats()
            one = [e.stats_row(t) for t in ts]
        monkeypatch.delenv('DBN_PREFIX_ONEPASS')
        with eng_mod(n_chains=3, unit_range=rng) as e:
            e.build_stats([data])
            three = [e.stats_row(t) for t in ts]
        for t, a, b in zip(ts, one, three):
            for x, y in zip(a, b):
                assert np.array_equal(x, y), (shape, rng, t)
        zz, zy, yy = one[-1]
        assert_close(zz, Z.T @ Z, 1e-12, 1e-12, what='ZZ[N]')
        if rng is None:
            assert_close(zy, (Z.T @ Y).T, 1e-12, 1e-12, what='ZY[N]')
            assert_close(yy, (Y ** 2).sum(axis=0), 1e-12, 1e-12, what='YY[N]')
