"""Drop-in for `dyban.Network` (network.py:15-416) whose per-node regressor loops run on the B200.

    from dynamicbayesiannetworks_b200 import Network          # or:  import dynamicbayesiannetworks_b200 as dyban
    net = Network([on, off], chain_length=10000, burn_in=5000, lag=1)
    net.infer_network('varying_nh_dbn')                       # 'h_dbn' | 'varying_nh_dbn' | 'seq_coup_nh_dbn' | 'glob_coup_nh_dbn'
    net.proposed_adj_matrix                                   # list of n lists of n floats, row = child

Same constructor signature, method strings and result attributes as the reference.  Extensions are keyword-only and
default to the reference's literals: `n_chains` (the reference runs one chain), `fan_in`, `seed`, `device`,
`keep_chain` (per-iteration histories of chain 0 in `chain_results`, as the reference's lists), `window`
(iterations per kernel launch).  The object holds only plain Python / NumPy data after `infer_network` returns, so
it stays picklable (examples/utils.py:226-240).  There is no CPU fallback.
"""
import numpy as np

from .engine import DbnEngine
from .methods import method_id, H_DBN, SEQ_DBN, GLOB_DBN, FP_GLOB_DBN, VV_GLOB_DBN, FP_H_DBN, FP_VV_DBN, FP_SEQ_DBN, FP_METHODS, METHOD_NAMES, FIXED_CP
from . import parallel

_MODEL_ARGS = {  # network.py:155-156, :191-192, :211-212, :235-236
    'varying_nh_dbn': ('Bayesian Non-Homogeneous', 'Varying Parents'),
    'h_dbn': ('Bayesian Homogeneous', 'Varying Parents'),
    'seq_coup_nh_dbn': ('Sequentilly Coupled Non-Homogeneous', 'Varying Parents'),
    'glob_coup_nh_dbn': ('Globally Coupled Non-Homogeneous', 'Varying Parents'),
    'fp_glob_coup_nh_dbn': ('Globally Coupled Non-Homogeneous', 'Full Parents'),   # network.py:246-247
    'fixed_nh_dbn': ('Bayesian Non-Homogeneous', 'Varying Parents-Fixed changepoints'),   # network.py:179-180
    'fp_varying_nh_dbn': ('Bayesian Non-Homogeneous', 'Full Parents'),   # network.py:166-167
    'fp_h_dbn': ('Bayesian Homogeneous', 'Full Parents'),                # network.py:201-202
    'fp_var_glob_coup_nh_dbn': ('Varying Globally Coupled Non-Homogeneous', 'Full Parents'),   # network.py:271-272
    'fp_seq_coup_nh_dbn': ('Sequentilly Coupled Non-Homogeneous', 'Full Parents'),             # network.py:222-223
    'var_glob_coup_nh_dbn': ('Varying Globally Coupled Non-Homogeneous', 'Varying Parents'),   # network.py:259-260
}


class Network():
    def __init__(self, data, chain_length, burn_in, lag, change_points=[], *, n_chains=1, fan_in=3, seed=0xDB20B200,
                 device=0, keep_chain=None, window=None, thin=10):
        self.data = data
        self.lag = lag
        self.change_points = change_points
        self.network_configuration = None
        self.chain_length = chain_length
        self.burn_in = burn_in
        self.true_adj_matrix = None
        self.proposed_adj_matrix = []
        self.edge_scores = None
        self.chain_results = None
        self.scores_over_time = []
        self.betas_over_time = []
        self.cps_over_response = []
        self.network_configurations = []
        self.network_args = {
            'model': None, 'type': None, 'length': str(self.chain_length), 'burn-in': str(self.burn_in),
            'thinning': 'modulo 10', 'scoring_method': None, 'network_configs': []
        }
        # extensions
        self.n_chains = int(n_chains)
        self.fan_in = int(fan_in)
        self.seed = int(seed)
        self.device = int(device)
        self.keep_chain = (self.n_chains == 1 and chain_length <= 200000) if keep_chain is None else bool(keep_chain)
        self.window = window
        self.thin = int(thin)
        self.changepoint_histogram = None
        self.chain_results_per_node = []
        self.counters = None

    # ---- reference-compatible helpers -----------------------------------------------------------------
    def set_network_configuration(self, configuration):
        """network.py:54-141 — kept for API compatibility: builds the {'features', 'response'} dict of one node."""
        if self.lag != 1:
            raise NotImplementedError('lag > 1 (network.py:92-122) is a later row of the scope table')
        dims = self.data[0].shape[1]
        feats = [x for x in range(dims) if x != configuration]
        data_dict = {'features': {}, 'response': {}}
        for el in feats:
            data_dict['features']['X' + str(el)] = np.concatenate([seg[:seg.shape[0] - 1, el] for seg in self.data])
        data_dict['response']['y'] = np.concatenate([seg[1:, configuration] for seg in self.data])
        self.network_configuration = data_dict
        self.network_configurations.append(data_dict)
        self.network_args['network_configs'].append(
            {'features': list(data_dict['features'].keys()), 'response': 'X' + str(configuration)})

    # ---- the hot path ----------------------------------------------------------------------------------
    def infer_network(self, method):
        """network.py:398-416: all nodes (and all chains) advance together on the device instead of one regressor per
        node in a Python loop; edge scores are the thinned-chain frequencies of scores.py:155-194."""
        mid = method_id(method)
        name = method if method in FIXED_CP else METHOD_NAMES[mid]
        if self.lag != 1:
            raise NotImplementedError('lag > 1 (network.py:92-122) is a later row of the scope table')
        dims = self.data[0].shape[1]
        self.network_args['model'], self.network_args['type'] = _MODEL_ARGS[name]
        self.network_args['scoring_method'] = 'edge-scores'
        for i in range(dims):
            self.set_network_configuration(i)
        n_iter = int(self.chain_length)
        # under torch.distributed (one process per GPU) the chains are sharded over the ranks and the count matrices
        # all-reduced at the end: every rank returns the scores of ALL chains
        rank, world = parallel.dist_rank_world()
        chain_offset, n_local = parallel.shard_chains(self.n_chains, rank, world)
        if n_local < 1:
            raise ValueError('n_chains = %d is fewer than the %d ranks' % (self.n_chains, world))
        if world > 1 and self.keep_chain:
            raise ValueError('keep_chain (per-iteration histories) is a single-process feature')
        with DbnEngine(n_chains=n_local, device=self.device, seed=self.seed, chain_offset=chain_offset) as eng:
            eng.build_stats(self.data, lag=self.lag)
            eng.chain_init(name, fan_in=self.fan_in, burn_in=int(self.burn_in), thin=self.thin)
            if name in FIXED_CP:   # network.py:178-189: the predefined change points, pseudo changepoint included
                eng.set_changepoints(self.change_points)
            N = eng.N
            traces = []
            window = self.window or (n_iter if self.keep_chain else 100)
            done = 0
            run = eng.run_fp if mid in FP_METHODS else eng.run
            while done < n_iter:
                w = min(window, n_iter - done)
                if self.keep_chain:
                    traces.append(run(w, trace=True))
                else:
                    run(w)
                done += w
            counts = eng.counts() if world == 1 else parallel.global_counts(eng)
            self.counters = eng.counters()
        self._finish(name, mid, dims, N, counts, traces)
        return self

    def _finish(self, name, mid, dims, N, counts, traces):
        if mid in FP_METHODS:
            return self._finish_fp(mid, dims, N, counts, traces)
        adj = parallel.edge_scores_from_counts(counts['edge'], counts['nonempty'])
        self.proposed_adj_matrix = [[float(v) for v in row] for row in adj]
        self.edge_scores = self.proposed_adj_matrix[-1]
        kept = np.maximum(counts['cp_kept'], 1)[:, None]
        self.changepoint_histogram = (counts['cp_hist'] / kept)[:, :N + 1]
        if traces:
            cat = {k: np.concatenate([t[k] for t in traces], axis=1) for k in traces[0]}
            C = self.n_chains
            self.chain_results_per_node = [self._chain_results(cat, node * C, mid, dims, N) for node in range(dims)]
            self.chain_results = self.chain_results_per_node[-1]
            if mid != H_DBN:
                b = int(self.burn_in)
                for r in self.chain_results_per_node:   # network.py:388-389
                    burned = r['tau_vector'][b:]
                    self.cps_over_response.append([burned[x] for x in range(len(burned)) if x % 10 == 0])

    def _finish_fp(self, mid, dims, N, counts, traces):
        """network.py:303-347: edge scores are credible scores of the thinned global-mean chain ('fp_glob_coup_nh_dbn') or of
        the thinned beta chain, every segment's vector a sample ('fp_varying_nh_dbn', 'fp_h_dbn')."""
        adj = parallel.credible_scores_from_counts(counts['edge'], counts['neg'], counts['nonempty'])
        self.network_args['scoring_method'] = 'fraq-score'
        self.proposed_adj_matrix = [[float(v) for v in row] for row in adj]
        self.edge_scores = self.proposed_adj_matrix[-1]
        kept = np.maximum(counts['cp_kept'], 1)[:, None]
        self.changepoint_histogram = (counts['cp_hist'] / kept)[:, :N + 1]
        if traces:
            cat = {k: np.concatenate([t[k] for t in traces], axis=1) for k in traces[0]}
            C, b = self.n_chains, int(self.burn_in)
            n_iter = cat['sigma2'].shape[1]
            for node in range(dims):
                u = node * C
                pi = [j for j in range(dims) if j != node]
                res = {   # fpGlobCoupBpwLinReg.py:112-119
                    'lambda_sqr_vector': [1] + [float(v) for v in cat['lambda2'][u]],
                    'sigma_sqr_vector': [1] + [float(v) for v in cat['sigma2'][u]],
                    'pi_vector': [pi],
                    'tau_vector': [] if mid == FP_H_DBN else
                                  [[int(c) for c in cat['tau_after'][u, it, :cat['S_after'][u, it]]] for it in range(n_iter)],
                    'betas_vector': [[np.zeros(dims)]] + [[np.array(cat['beta'][u, it, h]) for h in range(int(cat['S'][u, it]))]
                                                          for it in range(n_iter)],
                }
                if mid == FP_SEQ_DBN:  # fpSeqCoupBpwlinReg.py:67,88-90,127: delta^2 history (first per-segment variance slot)
                    res['delta_sqr_vector'] = [1] + [float(v) for v in cat['sigma2v'][u, :, 0]]
                if mid == FP_VV_DBN:   # fpvvGlobCoup.py:68,76: one variance per segment and iteration
                    res['sigma_sqr_vector'] = [[1]] + [[float(v) for v in cat['sigma2v'][u, it, :int(cat['S'][u, it])]] for it in range(n_iter)]
                if mid in (FP_GLOB_DBN, FP_VV_DBN):
                    res['mu_vector'] = [np.zeros((dims, 1))] + [np.array(cat['mu_after'][u, it]).reshape(-1, 1) for it in range(n_iter)]
                self.chain_results_per_node.append(res)
                burned = res['tau_vector'][b:]
                self.cps_over_response.append([burned[x] for x in range(len(burned)) if x % 10 == 0])
            self.chain_results = self.chain_results_per_node[-1]

    @staticmethod
    def _chain_results(t, u, mid, dims, N):
        """Per-iteration histories of one unit as the reference's result dict (bayesianPwLinearRegression.py:134-141):
        lists of length num_iter + 1 starting from the initial state."""
        n_iter = t['sigma2'].shape[1]
        pi_vec = [[]] + [np.array(t['pi_after'][u, it, :t['npi'][u, it]]) for it in range(n_iter)]
        res = {
            'lambda_sqr_vector': [1] + [float(v) for v in t['lambda2'][u]],
            'sigma_sqr_vector': [1] + [float(v) for v in t['sigma2'][u]],
            'pi_vector': pi_vec,
            'tau_vector': [] if mid == H_DBN else [[int(c) for c in t['tau_after'][u, it, :t['S_after'][u, it]]] for it in range(n_iter)],
        }
        betas = [[np.zeros(1)]]
        for it in range(n_iter):
            S = 1 if mid == H_DBN else int(t['S'][u, it])
            k = 1 if it == 0 else int(t['npi'][u, it - 1]) + 1
            betas.append([np.array(t['beta'][u, it, h, :k]) for h in range(S)])
        res['betas_vector'] = betas
        if mid == SEQ_DBN:
            res['delta_sqr_vector'] = [1] + [float(v) for v in t['delta2'][u]]
        if mid == VV_GLOB_DBN:   # vvglobCoup.py:59,70-72,125: one variance per segment and iteration
            res['sigma_sqr_vector'] = [[1]] + [[float(v) for v in t['sigma2v'][u, it, :int(t['S'][u, it])]] for it in range(n_iter)]
        if mid in (GLOB_DBN, VV_GLOB_DBN):
            res['mu_vector'] = [np.zeros((1, 1))] + [np.array(t['mu_after'][u, it, :t['npi'][u, it] + 1]).reshape(-1, 1) for it in range(n_iter)]
        return res
