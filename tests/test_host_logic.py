"""CPU tests of the host-side logic: method table vs the oracle's, chain sharding, count reduction (gloo, 2 ranks)."""
import glob
import os
import sys

import numpy as np
import pytest

from helpers import ROOT, GOLDEN, CHAIN_FILES, load, data_list
from oracle import dbn_oracle as O


def test_method_table_matches_oracle():
    from dynamicbayesiannetworks_b200.methods import run_params, METHOD_IDS
    for name in ('h_dbn', 'varying_nh_dbn', 'seq_coup_nh_dbn', 'glob_coup_nh_dbn', 'h-dbn', 'nh-dbn', 'seq-dbn', 'glob-dbn', 'fixed_nh_dbn'):
        for N in (33, 119, 1999):
            p = run_params(name, N)
            c = O.method_constants(name, N)
            if name == 'fixed_nh_dbn':      # same device kernel as nh, without the changepoint move
                assert METHOD_IDS[name] == O.NH_DBN and p['with_tau'] == 0
            else:
                assert METHOD_IDS[name] == O.METHOD_IDS[name]
            assert (p['T_ml'], p['T_sigma'], p['num_samples']) == (c['T_ml'], c['T_sigma'], c['num_samples'])
            assert (p['a_sigma'], p['b_sigma'], p['a_lambda'], p['b_lambda'], p['a_delta'], p['b_delta']) == (
                c['a_sig'], c['b_sig'], c['a_lam'], c['b_lam'], c['a_del'], c['b_del'])
            assert p['fan_in'] == 3 and p['thin'] == 10 and p['cp_p'] == 0.125
    p = run_params('fp_seq_coup_nh_dbn', 33)      # the last of the reference's eleven method strings (network.py:221-233)
    assert (p['num_samples'], p['T_ml'], p['a_sigma'], p['with_tau']) == (32, 32.0, 0.005, 1)
    with pytest.raises(ValueError):
        run_params('nonsense', 33)


def test_shard_chains_partitions_everything():
    from dynamicbayesiannetworks_b200.parallel import shard_chains
    for total in (1, 7, 64, 256, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                off, cnt = shard_chains(total, r, world)
                seen += list(range(off, off + cnt))
            assert seen == list(range(total))


def test_network_api_surface():
    """Constructor signature, attributes and method strings of dyban.Network (network.py:28-52, :398)."""
    import inspect
    from dynamicbayesiannetworks_b200 import Network
    sig = inspect.signature(Network.__init__)
    assert list(sig.parameters)[:6] == ['self', 'data', 'chain_length', 'burn_in', 'lag', 'change_points']
    g = load(CHAIN_FILES[0])
    net = Network(data_list(g), 100, 10, 1)
    for attr in ('proposed_adj_matrix', 'chain_results', 'cps_over_response', 'betas_over_time', 'scores_over_time',
                 'network_configurations', 'network_args', 'edge_scores', 'true_adj_matrix'):
        assert hasattr(net, attr)
    assert net.network_args['thinning'] == 'modulo 10' and net.network_args['length'] == '100'
    net.set_network_configuration(2)
    X, Y = O.lagged_matrices(data_list(g))
    cfg = net.network_configuration
    assert sorted(cfg['features']) == ['X0', 'X1', 'X3', 'X4'] if X.shape[1] == 5 else True
    assert np.array_equal(cfg['response']['y'], Y[:, 2])
    assert np.array_equal(cfg['features']['X0'], X[:, 0])
    import pickle
    pickle.loads(pickle.dumps(net))


def _gloo_worker(rank, world, port, tmp, mode='chains', total_chains=4):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from dynamicbayesiannetworks_b200 import parallel
    from dynamicbayesiannetworks_b200.engine import split_counts
    from helpers import CHAIN_FILES, load, data_list
    from oracle import dbn_oracle as O
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    g = load([p for p in CHAIN_FILES if 'yeast_varying' in p][0])
    X, Y = O.lagged_matrices(data_list(g))
    N, n = X.shape
    n_iter, burn = 60, 10
    if mode == 'chains':     # every node on every rank, the chains split (the weak-scaling bench)
        off, cnt = parallel.shard_chains(total_chains, rank, world)
        mine = [(node, c) for node in range(n) for c in range(off, off + cnt)]
    else:                    # the flat unit list u = node * C + chain split over the ranks
        first, cnt = parallel.shard_units(n * total_chains, rank, world)
        mine = [(u // total_chains, u % total_chains) for u in range(first, first + cnt)]
    edge = np.zeros((n, n)); ne = np.zeros(n); cph = np.zeros((n, N + 3)); cpk = np.zeros(n)
    for node, c in mine:     # streams are keyed by the grid's (node, chain)
        tr = O.run_chain('varying_nh_dbn', X, Y, node, n_iter, O.RngDraws(np.random.default_rng(1000 * node + c)))
        O.accumulate_counts(tr, node, n, N, burn, edge, ne, cph, cpk)
    flat = np.concatenate([edge.ravel(), ne, cph.ravel(), cpk]).astype(np.int64)
    local = torch.from_numpy(flat.copy())
    out = parallel.allreduce_counts(local)
    assert np.array_equal(local.numpy(), flat)       # local accumulators untouched
    np.save(os.path.join(tmp, 'sum_%d.npy' % rank), out.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize('mode,total_chains', [('chains', 4), ('units', 1), ('units', 3)])
def test_count_allreduce_two_ranks_gloo(tmp_path, mode, total_chains):
    """world_size-2 gloo run of the N>1 host path: shard chains (or the flat unit list, one chain included), accumulate
    per-rank counts, all-reduce; the result must equal the single-process accumulation over all units."""
    import torch.multiprocessing as mp
    from dynamicbayesiannetworks_b200.engine import split_counts
    from dynamicbayesiannetworks_b200 import parallel
    port = 29500 + (os.getpid() % 2000) + (0 if mode == 'chains' else 7 * total_chains)
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path), mode, total_chains), nprocs=2, join=True)
    s0 = np.load(tmp_path / 'sum_0.npy'); s1 = np.load(tmp_path / 'sum_1.npy')
    assert np.array_equal(s0, s1)
    g = load([p for p in CHAIN_FILES if 'yeast_varying' in p][0])
    X, Y = O.lagged_matrices(data_list(g))
    N, n = X.shape
    edge = np.zeros((n, n)); ne = np.zeros(n); cph = np.zeros((n, N + 3)); cpk = np.zeros(n)
    for node in range(n):
        for c in range(total_chains):
            tr = O.run_chain('varying_nh_dbn', X, Y, node, 60, O.RngDraws(np.random.default_rng(1000 * node + c)))
            O.accumulate_counts(tr, node, n, N, 10, edge, ne, cph, cpk)
    got = split_counts(s0, n, N)
    assert np.array_equal(got['edge'], edge.astype(np.int64))
    assert np.array_equal(got['nonempty'], ne.astype(np.int64))
    assert np.array_equal(got['cp_hist'], cph.astype(np.int64))
    assert np.array_equal(got['cp_kept'], cpk.astype(np.int64))
    adj = parallel.edge_scores_from_counts(got['edge'], got['nonempty'])
    assert np.allclose(adj, np.nan_to_num(O.edge_scores(edge, ne)))


def test_overlapped_allreduce_group_is_limited_to_a_few_ctas():
    """The all-reduce that runs beside the chain kernel must not take its SMs (profiles/r1_v13/nccl_diag.txt)."""
    dist = pytest.importorskip('torch.distributed')
    if not hasattr(dist, 'ProcessGroupNCCL'):
        pytest.skip('torch built without NCCL')
    from dynamicbayesiannetworks_b200 import parallel
    opts = parallel.nccl_small_footprint_options()
    assert opts.config.max_ctas == 2 and opts.config.min_ctas == 1
    assert parallel.nccl_small_footprint_options(1).config.max_ctas == 1


# ---------------------------------------------------------------------------------------------------------
# Network.score_edges' host-side derivations against the reference's own outputs (tests/golden/network_*.npz,
# generated by tests/golden/make_network_golden.py from the unmodified reference): padded_betas, cps_over_response,
# betas_over_time, scores_over_time (network.py:316-338, :370-396; scores.py:106-137;
# bayesianLinearRegression.py:40-67)
# ---------------------------------------------------------------------------------------------------------
NETWORK_FILES = sorted(p for p in glob.glob(os.path.join(GOLDEN, 'network_*.npz')) if 'lag2' not in p)


@pytest.mark.parametrize('path', NETWORK_FILES, ids=[os.path.basename(p)[8:-4] for p in NETWORK_FILES])
def test_network_diagnostics_vs_reference(path):
    from dynamicbayesiannetworks_b200 import Network, scores as S
    from dynamicbayesiannetworks_b200.methods import method_id, FP_METHODS
    g = dict(np.load(path, allow_pickle=False))
    method = str(g['method'])
    mid = method_id(method)
    data = [g['data_seg%d' % s] for s in range(int(g['n_data_segments']))]
    n, n_iter, burn = data[0].shape[1], int(g['n_iter']), int(g['burn_in'])
    N = sum(d.shape[0] - 1 for d in data)
    fp = mid in FP_METHODS
    net = Network(data, n_iter, burn, 1, diagnostics=True)
    for node in range(n):
        pi_vec = [g['pi'][node, it, :g['npi'][node, it]] for it in range(n_iter + 1)]
        has_tau = (g['tau'][node, 0] >= 0).any()
        tau_vec = [[int(c) for c in g['tau'][node, it] if c >= 0] for it in range(n_iter)] if has_tau else []
        width = n if fp else None
        betas = []
        for it in range(n_iter + 1):
            S_ = int(g['nseg'][node, it])
            betas.append([g['betas'][node, it, h][np.isfinite(g['betas'][node, it, h])] for h in range(S_)])
        res = {'pi_vector': pi_vec, 'tau_vector': tau_vec, 'betas_vector': betas}
        if not fp:
            if not g['nseg'][node].any():
                # the h-dbn and glob regressors return no 'betas_vector': their padded_betas ARE the only record of the
                # draws, so only the downstream functions are checked for them
                padded = [[v for v in g['padded_betas'][node, it] if np.isfinite(v).all()] for it in range(n_iter + 1)]
            else:
                # transform_beta_coef on the draw of iteration it - 1, conditioned on the parent set before it
                padded = [[np.zeros(n)]] + [S.transform_beta_coef(betas[it], pi_vec[it - 1], n - 1) for it in range(1, n_iter + 1)]
                for it in range(n_iter + 1):
                    for h, v in enumerate(padded[it]):
                        assert np.array_equal(v, g['padded_betas'][node, it, h]), (node, it, h)
            res['padded_betas'] = padded
        net._diagnostics(mid, method, node, N, res)
        ref_cps = [[int(c) for c in row if c >= 0] for row in g['cps_over_response'][node]]
        assert net.cps_over_response[node] == ref_cps, node
        if 'betas_over_time' in g:
            got = np.stack(net.betas_over_time[node])
            assert got.shape == g['betas_over_time'][node].shape
            assert np.array_equal(got, g['betas_over_time'][node]), node
        if 'scores_over_time' in g:
            assert np.array_equal(net.scores_over_time[node], g['scores_over_time'][node]), node


def test_lag2_network_configuration_vs_reference():
    """set_network_configuration with lag = 2 (network.py:92-122): labels, zero-padded lagged columns, response."""
    from dynamicbayesiannetworks_b200 import Network
    g = dict(np.load(os.path.join(GOLDEN, 'network_yeast_lag2_config.npz'), allow_pickle=False))
    data = [g['data_seg%d' % s] for s in range(int(g['n_data_segments']))]
    n = data[0].shape[1]
    net = Network(data, 10, 0, 2)
    for i in range(n):
        net.set_network_configuration(i)
        cfg = net.network_configuration
        labels = [int(l[1:]) for l in cfg['features']]
        assert labels == g['labels_%d' % i].tolist(), i
        assert np.array_equal(np.stack([cfg['features']['X%d' % l] for l in labels], axis=1), g['features_%d' % i]), i
        assert np.array_equal(cfg['response']['y'], g['response_%d' % i]), i
        assert net.network_args['network_configs'][i]['response'] == 'X%d' % i


def test_shard_units_covers_the_grid():
    from dynamicbayesiannetworks_b200 import parallel
    for total, world in ((4500, 8), (11 * 256, 4), (5, 8), (100, 1), (7, 7)):
        got = [parallel.shard_units(total, r, world) for r in range(world)]
        assert sum(c for _, c in got) == total
        pos = 0
        for first, count in got:
            assert first == pos
            pos += count
        assert max(c for _, c in got) - min(c for _, c in got) <= 1


def test_evaluation_helpers_vs_reference(tmp_path):
    """examples/utils.py:24-131, :226-261: transformResults (diagonal removal, lag-2 fold), the ROC / PR AUCs the example
    drivers print, save_chain / load_chain -- against outputs of the reference's own helpers (tests/golden/evaluation_yeast.npz,
    generated by tests/golden/make_network_golden.py evaluation-only)."""
    from dynamicbayesiannetworks_b200 import evaluation as E, Network
    g = dict(np.load(os.path.join(GOLDEN, 'evaluation_yeast.npz'), allow_pickle=False))
    true_inc = [list(map(int, r)) for r in g['true_inc']]
    for tag in ('lag1', 'lag2'):
        adj = [list(map(float, r)) for r in g['adj_' + tag]]
        ft, fs = E.transformResults(true_inc, adj)
        assert ft == g['flat_true_' + tag].tolist()
        assert np.array_equal(np.array(fs), g['flat_scores_' + tag])
        res = E.adjMatrixRoc(fs, ft, False)
        assert abs(res['roc_auc'] - float(g['roc_auc_' + tag])) < 1e-12
        assert abs(res['pr_auc'] - float(g['pr_auc_' + tag])) < 1e-12
    # the AUCs agree with scikit-learn on random scores with ties as well
    sk = pytest.importorskip('sklearn.metrics')
    rng = np.random.default_rng(0)
    for _ in range(20):
        y = rng.integers(0, 2, 40)
        if y.min() == y.max():
            continue
        s = np.round(rng.random(40), 1)
        fpr, tpr, _ = sk.roc_curve(y, s)
        pr, rc, _ = sk.precision_recall_curve(y, s)
        assert abs(E.roc_auc(y, s) - sk.auc(fpr, tpr)) < 1e-12
        assert abs(E.pr_auc(y, s) - sk.auc(rc, pr)) < 1e-12
    net = Network([g['adj_lag1']], 10, 0, 1)
    net.proposed_adj_matrix = [list(map(float, r)) for r in g['adj_lag1']]
    E.save_chain('net.pckl', net, folder=str(tmp_path))
    back = E.load_chain('net.pckl', folder=str(tmp_path))
    assert back.proposed_adj_matrix == net.proposed_adj_matrix
    assert E.load_chain('missing.pckl', folder=str(tmp_path)) is None
