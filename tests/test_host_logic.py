"""CPU tests of the host-side logic: method table vs the oracle's, chain sharding, count reduction (gloo, 2 ranks)."""
import os
import sys

import numpy as np
import pytest

from helpers import ROOT, CHAIN_FILES, load, data_list
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


def _gloo_worker(rank, world, port, tmp):
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
    total_chains, n_iter, burn = 4, 60, 10
    off, cnt = parallel.shard_chains(total_chains, rank, world)
    edge = np.zeros((n, n)); ne = np.zeros(n); cph = np.zeros((n, N + 3)); cpk = np.zeros(n)
    for node in range(n):
        for c in range(off, off + cnt):      # chain streams are keyed by the GLOBAL chain id
            tr = O.run_chain('varying_nh_dbn', X, Y, node, n_iter, O.RngDraws(np.random.default_rng(1000 * node + c)))
            O.accumulate_counts(tr, node, n, N, burn, edge, ne, cph, cpk)
    flat = np.concatenate([edge.ravel(), ne, cph.ravel(), cpk]).astype(np.int64)
    local = torch.from_numpy(flat.copy())
    out = parallel.allreduce_counts(local)
    assert np.array_equal(local.numpy(), flat)       # local accumulators untouched
    np.save(os.path.join(tmp, 'sum_%d.npy' % rank), out.numpy())
    dist.destroy_process_group()


def test_count_allreduce_two_ranks_gloo(tmp_path):
    """world_size-2 gloo run of the N>1 host path: shard chains, accumulate per-rank counts, all-reduce; the result
    must equal the single-process accumulation over all chains."""
    import torch.multiprocessing as mp
    from dynamicbayesiannetworks_b200.engine import split_counts
    from dynamicbayesiannetworks_b200 import parallel
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    s0 = np.load(tmp_path / 'sum_0.npy'); s1 = np.load(tmp_path / 'sum_1.npy')
    assert np.array_equal(s0, s1)
    g = load([p for p in CHAIN_FILES if 'yeast_varying' in p][0])
    X, Y = O.lagged_matrices(data_list(g))
    N, n = X.shape
    edge = np.zeros((n, n)); ne = np.zeros(n); cph = np.zeros((n, N + 3)); cpk = np.zeros(n)
    for node in range(n):
        for c in range(4):
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
