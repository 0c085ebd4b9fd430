#!/usr/bin/env python
"""Mint the golden vectors in tests/golden/*.npz by running the UNMODIFIED reference here.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
The reference's own tests pin nothing for this path, so every fixture is an output
of the reference itself, generated with both of its RNGs seeded and every draw recorded:

  kat_logml_<data>.npz   calculateMarginalLikelihoodWithChangepoints / calculateMarginalLikelihood
                         (marginalLikelihood.py:91-166) on enumerated (node, pi, tau, lambda2, delta2, mu),
                         design matrices built by the reference's own set_network_configuration /
                         constructNdArray / constructResponseNdArray; plus betaTildeSampler
                         (samplers.py:4-54), muSampler (samplers.py:307-351) and both priors (priors.py).
  chain_<data>_<method>.npz  a full Network.infer_network run with every random draw, every Gibbs
                         parameter, every log-ML and the resulting state per iteration (replay stream).
"""
import io
import os
import random
import sys
import contextlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _refboot  # noqa: E402

dyban = _refboot.import_reference()
from dyban import Network  # noqa: E402
import dyban.moves as ref_moves  # noqa: E402
import dyban.samplers as ref_samplers  # noqa: E402
import dyban.priors as ref_priors  # noqa: E402
import dyban.utils as ref_utils  # noqa: E402
import dyban.marginalLikelihood as ref_ml  # noqa: E402
from dynamicbayesiannetworks_b200.synth import make_synthetic  # noqa: E402

KM = 5      # padded design width (intercept + fan-in <= 4)
SM = 10     # padded number of segments
NAN = float('nan')

# ----------------------------------------------------------------------------------------------
# draw recorder
# ----------------------------------------------------------------------------------------------
EVENTS = []
_orig = {}


def _install_hooks():
    _orig['gamma'] = np.random.gamma
    _orig['mvn'] = np.random.multivariate_normal
    _orig['choice'] = np.random.choice
    _orig['uniform'] = np.random.uniform
    _orig['randint'] = ref_moves.randint
    _orig['logml_cp'] = ref_moves.calculateMarginalLikelihoodWithChangepoints
    _orig['logml_h'] = ref_moves.calculateMarginalLikelihood
    _orig['musampler'] = ref_moves.muSampler
    _orig['betatilde'] = ref_samplers.betaTildeSampler
    _orig['logml_vv'] = ref_moves.vvLogMargLikelihood
    _orig['vvmusampler'] = ref_moves.vvMuSampler

    def gamma(shape, scale=1.0, size=None):
        v = _orig['gamma'](shape, scale=scale, size=size)
        EVENTS.append(('gamma', float(shape), float(np.asarray(scale).item()), float(np.asarray(v).item())))
        return v

    def mvn(mean, cov, *a, **k):
        v = _orig['mvn'](mean, cov, *a, **k)
        EVENTS.append(('mvn', np.array(mean, dtype=float).ravel().copy(), np.array(cov, dtype=float).copy(),
                       np.array(v, dtype=float).ravel().copy()))
        return v

    def choice(a, size=None, *args, **k):
        v = _orig['choice'](a, size, *args, **k)
        arr = np.asarray(a).ravel()
        val = np.asarray(v).ravel()[0]
        idx = int(np.nonzero(arr == val)[0][0])
        EVENTS.append(('choice', idx, int(val), int(arr.size)))
        return v

    def uniform(low=0.0, high=1.0, size=None):
        v = _orig['uniform'](low, high, size)
        EVENTS.append(('uniform', float(v)))
        return v

    def randint(a, b):
        v = _orig['randint'](a, b)
        EVENTS.append(('randint', int(v)))
        return v

    def logml_cp(*a, **k):
        v = _orig['logml_cp'](*a, **k)
        EVENTS.append(('logml', float(np.asarray(v).item())))
        return v

    def logml_h(*a, **k):
        v = _orig['logml_h'](*a, **k)
        EVENTS.append(('rawml', float(np.asarray(v).item())))
        return v

    def musampler(*a, **k):
        s, d, db = _orig['musampler'](*a, **k)
        EVENTS.append(('musampler', float(d), float(db)))
        return s, d, db

    def logml_vv(*a, **k):
        v = _orig['logml_vv'](*a, **k)
        EVENTS.append(('logml', float(np.asarray(v).item())))
        return v

    def vvmusampler(*a, **k):
        s, d, db = _orig['vvmusampler'](*a, **k)
        EVENTS.append(('musampler', float(d), float(db)))
        return s, d, db

    np.random.gamma = gamma
    np.random.multivariate_normal = mvn
    np.random.choice = choice
    np.random.uniform = uniform
    ref_moves.randint = randint
    ref_moves.calculateMarginalLikelihoodWithChangepoints = logml_cp
    ref_moves.calculateMarginalLikelihood = logml_h
    ref_moves.muSampler = musampler
    ref_moves.vvLogMargLikelihood = logml_vv
    ref_moves.vvMuSampler = vvmusampler


def _remove_hooks():
    np.random.gamma = _orig['gamma']
    np.random.multivariate_normal = _orig['mvn']
    np.random.choice = _orig['choice']
    np.random.uniform = _orig['uniform']
    ref_moves.randint = _orig['randint']
    ref_moves.calculateMarginalLikelihoodWithChangepoints = _orig['logml_cp']
    ref_moves.calculateMarginalLikelihood = _orig['logml_h']
    ref_moves.muSampler = _orig['musampler']
    ref_moves.vvLogMargLikelihood = _orig['logml_vv']
    ref_moves.vvMuSampler = _orig['vvmusampler']


# ----------------------------------------------------------------------------------------------
# chain goldens
# ----------------------------------------------------------------------------------------------
def _pad_vec(v, n):
    out = np.full(n, NAN)
    v = np.asarray(v, dtype=float).ravel()
    out[:v.size] = v
    return out


def _pad_mat(m, n):
    out = np.full((n, n), NAN)
    m = np.asarray(m, dtype=float)
    out[:m.shape[0], :m.shape[1]] = m
    return out


def _parse_move(ev, pos, end):
    """ev[pos] is the 'randint' of a move; returns (record dict, next pos)."""
    assert ev[pos][0] == 'randint', ev[pos]
    rec = dict(move=ev[pos][1], picks=[], ncand=[], mvn=[], logml=[], rawml=[], mus=[], u=NAN, valid=0)
    pos += 1
    while pos < end and ev[pos][0] not in ('randint', 'gamma'):
        e = ev[pos]
        if e[0] == 'choice':
            rec['picks'].append(e[1]); rec['ncand'].append(e[3])
        elif e[0] == 'mvn':
            rec['mvn'].append(e)
        elif e[0] == 'logml':
            rec['logml'].append(e[1])
        elif e[0] == 'rawml':
            rec['rawml'].append(e[1])
        elif e[0] == 'musampler':
            rec['mus'].append((e[1], e[2]))
        elif e[0] == 'uniform':
            rec['u'] = e[1]; rec['valid'] = 1
        else:
            raise AssertionError(e)
        pos += 1
    return rec, pos


def run_chain_golden(name, data_list, method, n_iter, burn_in, seed, change_points=None, lag=1):
    """Run Network.infer_network(method) with recording and write chain_<name>_<method>.npz.
    `change_points`: the predefined list of 'fixed_nh_dbn' (network.py:178-189), ending with the pseudo changepoint."""
    n = data_list[0].shape[1]
    is_fixed = method == 'fixed_nh_dbn'              # segments but no changepoint move
    has_tau = method not in ('h_dbn', 'fixed_nh_dbn', 'fp_h_dbn')
    multi_seg = method != 'h_dbn'
    is_seq = method in ('seq_coup_nh_dbn', 'fp_seq_coup_nh_dbn')
    is_fp = method.startswith('fp_')              # full parents: no parent-set move, design width = n
    is_fp_glob = method == 'fp_glob_coup_nh_dbn'
    is_fp_vv = method == 'fp_var_glob_coup_nh_dbn'  # fpvvGlobCoup.py: mu redrawn every iteration (always accepted, :99-103)
    is_vv = method == 'var_glob_coup_nh_dbn' or is_fp_vv   # segment-specific sigma^2_h, mu kept in the changepoint move
    is_glob = method == 'glob_coup_nh_dbn' or is_fp_glob or is_vv
    KM = globals()['KM'] if not is_fp else max(globals()['KM'], n)
    EVENTS.clear()
    np.random.seed(seed)
    random.seed(seed)
    _install_hooks()
    try:
        net = Network([d.copy() for d in data_list], n_iter, burn_in, lag, list(change_points) if is_fixed else [])
        per_node = []
        with contextlib.redirect_stdout(io.StringIO()):
            for i in range(n):
                start = len(EVENTS)
                net.set_network_configuration(i)
                net.fit(method)
                res = net.chain_results
                per_node.append((start, len(EVENTS), res))
                if is_fp:
                    # network.py:303-347 without the plotting of scores.py:223-231 (credible_interval draws a
                    # histogram): thin the mu chain, stack it, credible_score per feature (scores.py:196-205,:245-246)
                    import dyban.scores as ref_scores
                    feats = [int(s_[1:]) for s_ in list(net.network_configuration['features'])]
                    if is_fp_glob:
                        burned = res['mu_vector'][burn_in:]
                        thinned = [[burned[x]] for x in range(len(burned)) if x % 10 == 0]
                    else:   # network.py:316-320: the beta chain shifted by 2, every segment's vector is a sample
                        burned = res['betas_vector'][2:][burn_in:]
                        thinned = [burned[x] for x in range(len(burned)) if x % 10 == 0]
                    bm = ref_scores.beta_post_matrix(thinned)
                    row = [0.0] * n
                    for j, f in enumerate(feats):
                        row[f] = ref_scores.credible_score(bm[:, j + 1])
                    net.proposed_adj_matrix.append(row)
                    burned_cps = res['tau_vector'][burn_in:] if has_tau else []
                    net.cps_over_response.append([burned_cps[x] for x in range(len(burned_cps)) if x % 10 == 0])
                    continue
                try:
                    net.score_edges(i, method)
                except AttributeError:
                    # scores.py:184 calls .size on the initial python-list pi when a chain never accepted
                    net.proposed_adj_matrix.append([NAN] * n)
                    net.cps_over_response.append([])
    finally:
        _remove_hooks()

    N = net.network_configuration['response']['y'].shape[0]
    g = {}
    shp = (n, n_iter)
    for key in ('sigma2', 'lambda2', 'delta2', 'u_pi', 'u_tau', 'sig_shape', 'sig_scale', 'lam_shape', 'lam_scale',
                'del_shape', 'del_scale'):
        g[key] = np.full(shp, NAN)
    if is_vv:
        for key in ('sigma2v', 'sig_shape_v', 'sig_scale_v'):
            g[key] = np.full(shp + (SM,), NAN)
    g['beta'] = np.full(shp + (SM, KM), NAN)
    g['beta_mean'] = np.full(shp + (SM, KM), NAN)
    g['beta_cov'] = np.full(shp + (SM, KM, KM), NAN)
    g['logml'] = np.full(shp + (4,), NAN)          # [pi*, pi, tau, tau*] in reference call order
    for key in ('pi_move', 'tau_move', 'pi_valid', 'tau_valid', 'pi_accept', 'tau_accept', 'npi', 'S'):
        g[key] = np.full(shp, -1, dtype=np.int32)
    g['pi_pick'] = np.full(shp + (2,), -1, dtype=np.int32)
    g['tau_pick'] = np.full(shp + (2,), -1, dtype=np.int32)
    g['pi_ncand'] = np.full(shp + (2,), -1, dtype=np.int32)
    g['tau_ncand'] = np.full(shp + (2,), -1, dtype=np.int32)
    g['pi_after'] = np.full(shp + (KM - 1,), -1, dtype=np.int32)   # reference (unsorted) order
    g['tau_after'] = np.full(shp + (SM,), -1, dtype=np.int32)
    if is_glob:
        g['mus_mean'] = np.full(shp + (4, KM), NAN)    # MVN calls in order [pi#1, pi#2, tau#1, tau#2]
        g['mus_cov'] = np.full(shp + (4, KM, KM), NAN)
        g['mus_val'] = np.full(shp + (4, KM), NAN)
        g['mus_logdens'] = np.full(shp + (4, 2), NAN)  # log(density), log(densityBefore) per muSampler call
        g['mu_after_pi'] = np.full(shp + (KM,), NAN)
        g['mu_after'] = np.full(shp + (KM,), NAN)

    for i, (start, end, res) in enumerate(per_node):
        ev = EVENTS
        pos = start
        pi_vec = res.get('pi_vector', [])
        tau_vec = res.get('tau_vector', [])
        cur_tau = [int(c) for c in change_points] if is_fixed else [N + 2]
        cur_pi = [j for j in range(n) if j != i] if is_fp else []
        cur_mu = np.zeros(len(cur_pi) + 1)
        for it in range(n_iter):
            S = len(cur_tau)
            g['S'][i, it] = S
            for h in range(S if is_vv else 1):      # segmentSigmaSampler (samplers.py:90-101): one draw per segment
                e = ev[pos]; assert e[0] == 'gamma', (i, it, e)
                if h == 0:
                    g['sig_shape'][i, it], g['sig_scale'][i, it], g['sigma2'][i, it] = e[1], e[2], 1.0 / e[3]
                if is_vv:
                    g['sig_shape_v'][i, it, h], g['sig_scale_v'][i, it, h], g['sigma2v'][i, it, h] = e[1], e[2], 1.0 / e[3]
                pos += 1
            for h in range(S if multi_seg else 1):
                e = ev[pos]; assert e[0] == 'mvn', (i, it, h, e)
                g['beta_mean'][i, it, h] = _pad_vec(e[1], KM)
                g['beta_cov'][i, it, h] = _pad_mat(e[2], KM)
                g['beta'][i, it, h] = _pad_vec(e[3], KM)
                pos += 1
            e = ev[pos]; assert e[0] == 'gamma', (i, it, e)
            g['lam_shape'][i, it], g['lam_scale'][i, it], g['lambda2'][i, it] = e[1], e[2], 1.0 / e[3]
            pos += 1
            if is_seq:
                e = ev[pos]; assert e[0] == 'gamma', (i, it, e)
                g['del_shape'][i, it], g['del_scale'][i, it], g['delta2'][i, it] = e[1], e[2], 1.0 / e[3]
                pos += 1
            if is_fp_vv:   # the vvMuSampler call of fpvvGlobCoup.py:100 (bound at import, so only its MVN draw is recorded)
                e = ev[pos]; assert e[0] == 'mvn', (i, it, e)
                g['mus_mean'][i, it, 0] = _pad_vec(e[1], KM)
                g['mus_cov'][i, it, 0] = _pad_mat(e[2], KM)
                g['mus_val'][i, it, 0] = _pad_vec(e[3], KM)
                cur_mu = e[3].copy()
                pos += 1
            # ---- pi move
            if is_fp:
                rec = dict(move=-1, valid=-1, u=NAN, picks=[], ncand=[], mvn=[], mus=[], logml=[], rawml=[])
            else:
                rec, pos = _parse_move(ev, pos, end)
            g['pi_move'][i, it] = rec['move']
            g['pi_valid'][i, it] = rec['valid']
            g['u_pi'][i, it] = rec['u']
            for j, (p, c) in enumerate(zip(rec['picks'], rec['ncand'])):
                g['pi_pick'][i, it, j] = p
                g['pi_ncand'][i, it, j] = c
            new_pi = cur_pi if is_fp else [int(v) for v in np.asarray(pi_vec[it + 1]).ravel()]
            g['pi_accept'][i, it] = -1 if is_fp else int(new_pi != cur_pi)
            if method == 'h_dbn':
                if rec['valid']:
                    g['logml'][i, it, 0] = np.log(rec['rawml'][0])
                    g['logml'][i, it, 1] = np.log(rec['rawml'][1])
            elif rec['valid'] > 0:
                g['logml'][i, it, 0] = rec['logml'][0]
                g['logml'][i, it, 1] = rec['logml'][1]
            if is_glob and rec['valid'] > 0:
                for j, m in enumerate(rec['mvn']):
                    g['mus_mean'][i, it, j] = _pad_vec(m[1], KM)
                    g['mus_cov'][i, it, j] = _pad_mat(m[2], KM)
                    g['mus_val'][i, it, j] = _pad_vec(m[3], KM)
                for j, (d, db) in enumerate(rec['mus']):
                    g['mus_logdens'][i, it, j] = (np.log(d), np.log(db))
                if g['pi_accept'][i, it]:
                    cur_mu = rec['mvn'][0][3].copy()
            cur_pi = new_pi
            g['npi'][i, it] = len(cur_pi)
            g['pi_after'][i, it, :len(cur_pi)] = cur_pi
            if is_glob:
                g['mu_after_pi'][i, it] = _pad_vec(cur_mu, KM)
            # ---- tau move
            if has_tau:
                rec, pos = _parse_move(ev, pos, end)
                g['tau_move'][i, it] = rec['move']
                g['tau_valid'][i, it] = rec['valid']
                g['u_tau'][i, it] = rec['u']
                for j, (p, c) in enumerate(zip(rec['picks'], rec['ncand'])):
                    g['tau_pick'][i, it, j] = p
                    g['tau_ncand'][i, it, j] = c
                new_tau = [int(v) for v in tau_vec[it]]
                g['tau_accept'][i, it] = int(new_tau != cur_tau)
                if rec['valid']:
                    g['logml'][i, it, 2] = rec['logml'][0]
                    g['logml'][i, it, 3] = rec['logml'][1]
                if is_glob and not is_vv and rec['valid']:
                    for j, m in enumerate(rec['mvn']):
                        g['mus_mean'][i, it, 2 + j] = _pad_vec(m[1], KM)
                        g['mus_cov'][i, it, 2 + j] = _pad_mat(m[2], KM)
                        g['mus_val'][i, it, 2 + j] = _pad_vec(m[3], KM)
                    for j, (d, db) in enumerate(rec['mus']):
                        g['mus_logdens'][i, it, 2 + j] = (np.log(d), np.log(db))
                    if g['tau_accept'][i, it]:
                        cur_mu = rec['mvn'][1][3].copy()
                cur_tau = new_tau
            g['tau_after'][i, it, :len(cur_tau)] = cur_tau
            if is_glob:
                g['mu_after'][i, it] = _pad_vec(cur_mu, KM)
            # cross-check the recorded Gibbs outputs against the reference's own history
            if is_vv:
                assert np.allclose(res['sigma_sqr_vector'][it + 1], g['sigma2v'][i, it, :S], rtol=1e-12, atol=0)
            else:
                assert abs(res['sigma_sqr_vector'][it + 1] - g['sigma2'][i, it]) <= 1e-12 * abs(g['sigma2'][i, it])
            assert abs(float(np.asarray(res['lambda_sqr_vector'][it + 1]).item()) - g['lambda2'][i, it]) \
                <= 1e-12 * abs(g['lambda2'][i, it])
        assert pos == end, (i, pos, end, ev[pos] if pos < end else None)

    out = dict(g)
    if is_fp:
        # n-wide covariances would make the fixture several MB: keep their diagonals (the means, log-densities and
        # log-MLs recorded beside them pin the same matrices through Q^-1 and ln det Q)
        out['beta_cov_diag'] = np.diagonal(out.pop('beta_cov'), axis1=-2, axis2=-1).copy()
        if 'mus_cov' in out:
            out['mus_cov_diag'] = np.diagonal(out.pop('mus_cov'), axis1=-2, axis2=-1).copy()
    out['method'] = np.array(method)
    if is_fixed:
        out['change_points'] = np.array([int(c) for c in change_points], dtype=np.int32)
    out['n_iter'] = np.array(n_iter)
    out['burn_in'] = np.array(burn_in)
    out['seed'] = np.array(seed)
    out['N'] = np.array(N)
    out['n_data_segments'] = np.array(len(data_list))
    for s, d in enumerate(data_list):
        out['data_seg%d' % s] = np.asarray(d, dtype=np.float64)
    adj = np.array([[float(v) for v in row] for row in net.proposed_adj_matrix], dtype=np.float64)
    out['proposed_adj_matrix'] = adj
    if has_tau:
        # thinned tau chain kept by network.py:388-389 -> changepoint histogram definition
        hist = np.zeros((n, N + 3))
        kept = np.zeros(n)
        for i, cps in enumerate(net.cps_over_response):
            for tau in cps:
                kept[i] += 1
                for c in tau[:-1]:
                    hist[i, c] += 1
        out['cp_hist'] = hist
        out['cp_kept'] = kept
    out['lag'] = np.array(lag)
    # lag > 1 (network.py:92-122): parents are LABELS -- gene ids for the lag-1 block, n + position for the lag-2 block
    path = os.path.join(HERE, ('chain_%s_%s.npz' if lag == 1 else 'lag%dchain_%%s_%%s.npz' % lag) % (name, method))
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


# ----------------------------------------------------------------------------------------------
# function-level known-answer vectors
# ----------------------------------------------------------------------------------------------
def _random_tau(rng, N, S, lo=2):
    """S-1 distinct real changepoints in [2, N] + the pseudo changepoint N+2 (network.py:162)."""
    real = sorted(int(v) for v in rng.choice(np.arange(lo, N + 1), size=S - 1, replace=False))
    return real + [N + 2]


def run_kat(name, data_list, n_cases, seed, max_parents=4):
    rng = np.random.default_rng(seed)
    n = data_list[0].shape[1]
    net = Network([d.copy() for d in data_list], 10, 0, 1)
    cfgs = []
    for i in range(n):
        net.set_network_configuration(i)
        cfgs.append(net.network_configuration)
    N = cfgs[0]['response']['y'].shape[0]

    keys_f = ('lambda2', 'delta2', 'sigma2', 'alpha', 'beta', 'T', 'logml_plain', 'logml_mu', 'logml_seq', 'rawml_h',
              'cp_prior', 'pi_prior', 'mus_logdens', 'mus_logdens_before')
    K = {k: np.full(n_cases, NAN) for k in keys_f}
    K['node'] = np.zeros(n_cases, dtype=np.int32)
    K['npi'] = np.zeros(n_cases, dtype=np.int32)
    K['pi'] = np.full((n_cases, KM - 1), -1, dtype=np.int32)
    K['S'] = np.zeros(n_cases, dtype=np.int32)
    K['tau'] = np.full((n_cases, SM), -1, dtype=np.int32)
    K['mu'] = np.full((n_cases, KM), NAN)
    K['beta_tilde'] = np.full((n_cases, SM, KM), NAN)
    K['mus_mean'] = np.full((n_cases, KM), NAN)
    K['mus_cov'] = np.full((n_cases, KM, KM), NAN)
    K['fan_in'] = np.zeros(n_cases, dtype=np.int32)

    for c in range(n_cases):
        node = int(rng.integers(0, n))
        feats = [j for j in range(n) if j != node]
        npi = int(rng.integers(0, min(max_parents, len(feats)) + 1))
        pi = [int(v) for v in rng.choice(feats, size=npi, replace=False)]
        S = int(rng.integers(1, min(9, N - 1) + 1))
        if c % 7 == 0:
            S = 1
        tau = _random_tau(rng, N, S)
        lam = float(np.exp(rng.uniform(np.log(0.02), np.log(50.0))))
        dlt = float(np.exp(rng.uniform(np.log(0.02), np.log(50.0))))
        sig = float(np.exp(rng.uniform(np.log(0.05), np.log(5.0))))
        alpha = beta = 0.005 if c % 2 == 0 else 0.01
        T = N if c % 3 else N - 1
        k = npi + 1
        mu = rng.normal(0, 0.5, size=(k, 1))
        data = cfgs[node]
        partial = ref_utils.selectData(data, pi)
        X = ref_utils.constructNdArray(partial, N, tau)
        y = ref_utils.constructResponseNdArray(data['response']['y'], tau)
        zero = ref_utils.constructMuMatrix(pi)

        K['node'][c] = node; K['npi'][c] = npi; K['pi'][c, :npi] = pi
        K['S'][c] = S; K['tau'][c, :S] = tau
        K['lambda2'][c] = lam; K['delta2'][c] = dlt; K['sigma2'][c] = sig
        K['alpha'][c] = alpha; K['beta'][c] = beta; K['T'][c] = T
        K['mu'][c, :k] = mu.ravel()
        K['logml_plain'][c] = np.asarray(ref_ml.calculateMarginalLikelihoodWithChangepoints(
            X, y, zero, alpha, beta, lam, T, tau)).item()
        K['logml_mu'][c] = np.asarray(ref_ml.calculateMarginalLikelihoodWithChangepoints(
            X, y, mu, alpha, beta, lam, T, tau)).item()
        bt = ref_samplers.betaTildeSampler(y, X, zero, tau, lam, dlt)
        for h in range(S):
            K['beta_tilde'][c, h, :k] = np.asarray(bt[h]).ravel()
        K['logml_seq'][c] = np.asarray(ref_ml.calculateMarginalLikelihoodWithChangepoints(
            X, y, bt, alpha, beta, lam, T, tau, 'seq-coup', dlt)).item()
        if S == 1:
            Xh = ref_utils.constructDesignMatrix(partial, N)
            K['rawml_h'][c] = np.asarray(ref_ml.calculateMarginalLikelihood(
                Xh, data['response']['y'], zero, alpha, beta, lam, N)).item()
        K['cp_prior'][c] = ref_priors.calculateChangePointsSetPrior(tau)
        fan = 3 if c % 2 else 4
        K['fan_in'][c] = fan
        K['pi_prior'][c] = ref_priors.calculateFeatureSetPriorProb(pi, len(feats), fan)
        # muSampler: capture the MVN arguments (mean/cov) and both densities
        cap = {}
        orig = np.random.multivariate_normal

        def mvn(mean, cov, *a, **kw):
            cap['mean'] = np.array(mean, dtype=float).ravel(); cap['cov'] = np.array(cov, dtype=float)
            return orig(mean, cov, *a, **kw)
        np.random.multivariate_normal = mvn
        try:
            np.random.seed(seed + c)
            _, dens, dens_before = ref_samplers.muSampler(mu, tau, X, y, sig, lam)
        finally:
            np.random.multivariate_normal = orig
        K['mus_mean'][c, :k] = cap['mean']
        K['mus_cov'][c, :k, :k] = cap['cov']
        with np.errstate(divide='ignore'):
            K['mus_logdens'][c] = np.log(dens)
            K['mus_logdens_before'][c] = np.log(dens_before)

    out = dict(K)
    out['N'] = np.array(N)
    out['n_data_segments'] = np.array(len(data_list))
    for s, d in enumerate(data_list):
        out['data_seg%d' % s] = np.asarray(d, dtype=np.float64)
    path = os.path.join(HERE, 'kat_logml_%s.npz' % name)
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


def main():
    yeast = _refboot.read_yeast()
    mtor = _refboot.read_mtor()
    syn, _, _ = make_synthetic(8, 200, n_changepoints=3, max_fan_in=3, seed=7)
    syn_long, _, _ = make_synthetic(6, 341, n_changepoints=2, max_fan_in=3, seed=11)

    if 'lag2-only' in sys.argv:     # lag-2 designs (zero-padded lagged copies, network.py:92-122)
        run_chain_golden('yeast', yeast, 'varying_nh_dbn', n_iter=240, burn_in=40, seed=5, lag=2)
        run_chain_golden('yeast', yeast, 'h_dbn', n_iter=240, burn_in=40, seed=5, lag=2)
        run_chain_golden('yeast', yeast, 'seq_coup_nh_dbn', n_iter=160, burn_in=40, seed=5, lag=2)
        run_chain_golden('yeast', yeast, 'glob_coup_nh_dbn', n_iter=160, burn_in=40, seed=5, lag=2)
        run_chain_golden('synth200', [syn], 'varying_nh_dbn', n_iter=120, burn_in=20, seed=6, lag=2)
        return
    if 'fixed-only' in sys.argv:    # add the fixed-changepoint fixtures without re-minting the others
        run_chain_golden('yeast', yeast, 'fixed_nh_dbn', n_iter=240, burn_in=40, seed=1, change_points=[12, 23, 35])
        run_chain_golden('synth200', [syn], 'fixed_nh_dbn', n_iter=120, burn_in=20, seed=3, change_points=[51, 101, 151, 201])
        return
    if 'vv-only' in sys.argv:       # add the segment-specific-variance fixtures without re-minting the others
        run_chain_golden('yeast', yeast, 'var_glob_coup_nh_dbn', n_iter=240, burn_in=40, seed=1)
        run_chain_golden('mtor', mtor, 'var_glob_coup_nh_dbn', n_iter=80, burn_in=20, seed=2)
        run_chain_golden('synth200', [syn], 'var_glob_coup_nh_dbn', n_iter=120, burn_in=20, seed=3)
        return
    if 'fpseq-only' in sys.argv:
        run_chain_golden('yeast', yeast, 'fp_seq_coup_nh_dbn', n_iter=240, burn_in=40, seed=1)
        run_chain_golden('mtor', mtor, 'fp_seq_coup_nh_dbn', n_iter=60, burn_in=20, seed=2)
        run_chain_golden('synth200', [syn], 'fp_seq_coup_nh_dbn', n_iter=100, burn_in=20, seed=3)
        return
    if 'fpvv-only' in sys.argv:
        run_chain_golden('yeast', yeast, 'fp_var_glob_coup_nh_dbn', n_iter=240, burn_in=40, seed=1)
        run_chain_golden('mtor', mtor, 'fp_var_glob_coup_nh_dbn', n_iter=60, burn_in=20, seed=2)
        run_chain_golden('synth200', [syn], 'fp_var_glob_coup_nh_dbn', n_iter=100, burn_in=20, seed=3)
        return
    if 'fp2-only' in sys.argv:      # full-parents nh / h fixtures
        run_chain_golden('yeast', yeast, 'fp_varying_nh_dbn', n_iter=240, burn_in=40, seed=1)
        run_chain_golden('mtor', mtor, 'fp_varying_nh_dbn', n_iter=60, burn_in=20, seed=2)
        run_chain_golden('synth200', [syn], 'fp_varying_nh_dbn', n_iter=100, burn_in=20, seed=3)
        run_chain_golden('yeast', yeast, 'fp_h_dbn', n_iter=240, burn_in=40, seed=1)
        run_chain_golden('mtor', mtor, 'fp_h_dbn', n_iter=60, burn_in=20, seed=2)
        return
    if 'fp-only' in sys.argv:       # add the full-parents fixtures without re-minting the others
        run_chain_golden('yeast', yeast, 'fp_glob_coup_nh_dbn', n_iter=240, burn_in=40, seed=1)
        run_chain_golden('mtor', mtor, 'fp_glob_coup_nh_dbn', n_iter=60, burn_in=20, seed=2)
        run_chain_golden('synth200', [syn], 'fp_glob_coup_nh_dbn', n_iter=100, burn_in=20, seed=3)
        return
    run_kat('yeast', yeast, 160, seed=101)
    run_kat('mtor', mtor, 160, seed=102)
    run_kat('synth200', [syn], 120, seed=103)
    run_kat('synth341', [syn_long], 60, seed=104)

    methods = ['h_dbn', 'varying_nh_dbn', 'seq_coup_nh_dbn', 'glob_coup_nh_dbn']
    for m in methods:
        run_chain_golden('yeast', yeast, m, n_iter=240, burn_in=40, seed=1)
    for m in methods:
        run_chain_golden('mtor', mtor, m, n_iter=80, burn_in=20, seed=2)
    for m in methods[1:]:
        run_chain_golden('synth200', [syn], m, n_iter=120, burn_in=20, seed=3)
    run_chain_golden('yeast', yeast, 'fp_glob_coup_nh_dbn', n_iter=240, burn_in=40, seed=1)
    run_chain_golden('mtor', mtor, 'fp_glob_coup_nh_dbn', n_iter=60, burn_in=20, seed=2)
    run_chain_golden('synth200', [syn], 'fp_glob_coup_nh_dbn', n_iter=100, burn_in=20, seed=3)
    run_chain_golden('yeast', yeast, 'fixed_nh_dbn', n_iter=240, burn_in=40, seed=1, change_points=[12, 23, 35])
    run_chain_golden('synth200', [syn], 'fixed_nh_dbn', n_iter=120, burn_in=20, seed=3, change_points=[51, 101, 151, 201])
    run_chain_golden('yeast', yeast, 'var_glob_coup_nh_dbn', n_iter=240, burn_in=40, seed=1)
    run_chain_golden('mtor', mtor, 'var_glob_coup_nh_dbn', n_iter=80, burn_in=20, seed=2)
    run_chain_golden('synth200', [syn], 'var_glob_coup_nh_dbn', n_iter=120, burn_in=20, seed=3)
    run_chain_golden('yeast', yeast, 'fp_varying_nh_dbn', n_iter=240, burn_in=40, seed=1)
    run_chain_golden('mtor', mtor, 'fp_varying_nh_dbn', n_iter=60, burn_in=20, seed=2)
    run_chain_golden('synth200', [syn], 'fp_varying_nh_dbn', n_iter=100, burn_in=20, seed=3)
    run_chain_golden('yeast', yeast, 'fp_h_dbn', n_iter=240, burn_in=40, seed=1)
    run_chain_golden('mtor', mtor, 'fp_h_dbn', n_iter=60, burn_in=20, seed=2)
    run_chain_golden('yeast', yeast, 'fp_var_glob_coup_nh_dbn', n_iter=240, burn_in=40, seed=1)
    run_chain_golden('mtor', mtor, 'fp_var_glob_coup_nh_dbn', n_iter=60, burn_in=20, seed=2)
    run_chain_golden('synth200', [syn], 'fp_var_glob_coup_nh_dbn', n_iter=100, burn_in=20, seed=3)
    run_chain_golden('yeast', yeast, 'fp_seq_coup_nh_dbn', n_iter=240, burn_in=40, seed=1)
    run_chain_golden('mtor', mtor, 'fp_seq_coup_nh_dbn', n_iter=60, burn_in=20, seed=2)
    run_chain_golden('synth200', [syn], 'fp_seq_coup_nh_dbn', n_iter=100, burn_in=20, seed=3)


if __name__ == '__main__':
    main()
