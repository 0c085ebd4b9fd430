"""Drop-in for `dyban.Network` (network.py:15-416) whose per-node regressor loops run on the B200.

    from dynamicbayesiannetworks_b200 import Network          # or:  import dynamicbayesiannetworks_b200 as dyban
    net = Network([on, off], chain_length=10000, burn_in=5000, lag=1)
    net.infer_network('varying_nh_dbn')                       # every method string of network.py:154-281
    net.proposed_adj_matrix                                   # list of n lists of n floats, row = child

Same constructor signature, method strings, public methods (`set_network_configuration`, `fit`, `score_edges`,
`infer_network`) and result attributes as the reference.  Extensions are keyword-only and default to the reference's
literals: `n_chains` (the reference runs one chain), `fan_in`, `seed`, `device`, `keep_chain` (per-iteration histories
of chain 0 in `chain_results`, as the reference's lists), `window` (iterations per kernel launch), `thin`,
`diagnostics` (per-time-point beta matrices).  The object holds only plain Python / NumPy data after a call returns, so
it stays picklable (examples/utils.py:226-240).  There is no CPU fallback.

`infer_network` advances ALL nodes (and chains) together; the reference's own calling sequence
`set_network_configuration(i); fit(method); score_edges(i, method)` (network.py:413-416) works too and runs one node's
slice of the unit grid per `fit`.  Under `torch.distributed` (one process per GPU) the flat (node, chain) unit list is
split over the ranks and the count matrices are all-reduced, whatever the number of chains.
"""
import warnings

import numpy as np

from . import _lib as L
from .engine import DbnEngine

# resident threads of the h / nh kernel and of the coupled kernel (2 blocks of 128 per SM each) on one B200: a run with more
# units than one wave is split over two handles (two streams), whose launches fill each other's partial last wave
ONE_WAVE_UNITS = 148 * 2 * 128
ONE_WAVE_UNITS_COUPLED = 148 * 2 * 128
from .methods import method_id, H_DBN, SEQ_DBN, GLOB_DBN, FP_GLOB_DBN, VV_GLOB_DBN, FP_H_DBN, FP_VV_DBN, FP_SEQ_DBN, FP_METHODS, METHOD_NAMES, FIXED_CP
from . import parallel
from . import scores as S

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
TRACE_WINDOW_BYTES = 256 << 20      # device + host bytes of one traced launch (keep_chain runs are traced in windows)
KEEP_CHAIN_BYTES = 4 << 30          # default keep_chain only if the whole history fits this many host bytes
DIAGNOSTIC_BYTES = 1 << 30          # default betas_over_time only below this many bytes


class Network():
    def __init__(self, data, chain_length, burn_in, lag, change_points=[], *, n_chains=1, fan_in=3, seed=0xDB20B200,
                 device=None, keep_chain=None, window=None, thin=10, diagnostics=None):
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
        self.thin = int(thin)
        self.network_args = {
            'model': None, 'type': None, 'length': str(self.chain_length), 'burn-in': str(self.burn_in),
            'thinning': 'modulo %d' % self.thin, 'scoring_method': None, 'network_configs': []
        }
        # extensions
        self.n_chains = int(n_chains)
        self.fan_in = int(fan_in)
        self.seed = int(seed)
        self.device = None if device is None else int(device)
        self.keep_chain = None if keep_chain is None else bool(keep_chain)
        self.window = window
        self.diagnostics = diagnostics
        self.changepoint_histogram = None
        self.chain_results_per_node = []
        self.counters = None
        self._current = None            # node of the last set_network_configuration
        self._node_counts = None        # counts of the last fit()

    # ---- reference-compatible helpers -----------------------------------------------------------------
    def set_network_configuration(self, configuration):
        """network.py:54-141: the {'features': {'X<label>': column}, 'response': {'y': column}} dict of one node.  Lag-1
        columns are labelled by gene id; every further lag adds a block of zero-padded copies labelled by a running
        counter (network.py:92-122), so with lag 2 and n genes the second block's labels are n .. 2n - 2."""
        dims = self.data[0].shape[1]
        feats = [x for x in range(dims) if x != configuration]
        data_dict = {'features': {}, 'response': {}}
        count_label = 1
        for lag in range(int(self.lag)):
            for el in feats:
                label = count_label if lag + 1 > 1 else el
                cols = []
                for seg in self.data:
                    col = seg[:seg.shape[0] - (lag + 1), el]
                    cols.append(np.concatenate([np.zeros(lag), col]) if lag else col)
                data_dict['features']['X' + str(label)] = np.concatenate(cols)
                count_label += 1
        data_dict['response']['y'] = np.concatenate([seg[1:, configuration] for seg in self.data])
        self.network_configuration = data_dict
        self.network_configurations.append(data_dict)
        self.network_args['network_configs'].append(
            {'features': list(data_dict['features'].keys()), 'response': 'X' + str(configuration)})
        self._current = int(configuration)

    # ---- plumbing ---------------------------------------------------------------------------------------
    def _resolve(self, method):
        mid = method_id(method)
        name = method if method in FIXED_CP else METHOD_NAMES[mid]
        if int(self.lag) not in (1, 2) or (int(self.lag) != 1 and mid in FP_METHODS):
            raise NotImplementedError('the device path takes lag 1 or 2 (varying-parents methods) and lag 1 (full-parents methods); '
                                      'set_network_configuration builds the lagged design dict for any lag')
        self.network_args['model'], self.network_args['type'] = _MODEL_ARGS[name]
        return mid, name

    def _trace_stride(self, mid, n):
        return L.fp_td_stride(n) if mid in FP_METHODS else L.TD_STRIDE

    def _plan(self, mid, n, n_units, n_iter, world):
        """(keep_chain, window): histories are kept by default only for a single chain whose whole history fits
        KEEP_CHAIN_BYTES on the host; traced runs go in windows of at most TRACE_WINDOW_BYTES per launch."""
        stride = self._trace_stride(mid, n) * 8 + L.TI_STRIDE * 4
        total = n_units * n_iter * stride
        keep = self.keep_chain
        if keep is None:
            keep = self.n_chains == 1 and world == 1 and total <= KEEP_CHAIN_BYTES
        if keep and world > 1:
            raise ValueError('keep_chain (per-iteration histories) is a single-process feature')
        if keep and total > 16 * KEEP_CHAIN_BYTES:
            raise MemoryError('keep_chain would hold %.1f GB of per-iteration history (%d units x %d iterations x %d bytes); '
                              'pass keep_chain=False (edge scores need no history) or a shorter chain' % (total / 1e9, n_units, n_iter, stride))
        if self.window:
            window = int(self.window)
        elif keep:
            window = int(max(1, min(n_iter, TRACE_WINDOW_BYTES // max(1, n_units * stride))))
        else:
            window = 100
        return keep, window

    def _run(self, mid, name, nodes):
        """Advance the units of `nodes` (None: all nodes) for chain_length iterations; returns (N, counts, traces, units
        per node in the trace, first traced node)."""
        n = self.data[0].shape[1]
        n_iter = int(self.chain_length)
        rank, world = parallel.dist_rank_world()
        C = self.n_chains
        if nodes is None:
            first, count = parallel.shard_units(n * C, rank, world)
        else:
            first, count = nodes * C, C
            world = 1
        if count < 1:
            raise ValueError('%d units cannot be split over %d ranks' % (n * C, world))
        keep, window = self._plan(mid, n, count, n_iter, world)
        device = parallel.default_device() if self.device is None else self.device
        # Large varying-parents runs without histories use TWO handles per GPU, each a contiguous half of the unit range on its own stream:
        # 102 400 units are 2.7 waves of resident blocks (2 x 128 threads per SM), and the launches of one half fill the partial last wave of the other
        # (+10 %, scripts/two_stream_ab.py).  The chains are the same chains (Philox is keyed by (chain, node)).
        # (measured: c4 nh 6.4e8 -> 7.0e8; mTOR 4096 chains seq 1.95e8 -> 2.35e8, glob 2.2e8 -> 2.67e8, profiles/r2_small/two_handles.txt)
        fast = name in ('h_dbn', 'varying_nh_dbn', 'fixed_nh_dbn')
        two = (not keep) and mid not in FP_METHODS and count > (ONE_WAVE_UNITS if fast else ONE_WAVE_UNITS_COUPLED)
        ranges = [(first, count)] if not two else [(first + f_, c_) for f_, c_ in (parallel.shard_units(count, i, 2) for i in range(2))]
        engines = []
        try:
            for rg in ranges:
                eng = DbnEngine(n_chains=C, device=device, seed=self.seed, unit_range=rg)
                engines.append(eng)
                eng.build_stats(self.data, lag=self.lag)
                eng.chain_init(name, fan_in=self.fan_in, burn_in=int(self.burn_in), thin=self.thin)
                if name in FIXED_CP:   # network.py:178-189: the predefined change points, pseudo changepoint included
                    eng.set_changepoints(self.change_points)
            N = engines[0].N
            traces = []
            done = 0
            while done < n_iter:
                w = min(window, n_iter - done)
                for eng in engines:      # asynchronous launches, one stream per handle
                    run = eng.run_fp if mid in FP_METHODS else eng.run
                    if keep:
                        traces.append(run(w, trace=True))
                    else:
                        run(w)
                done += w
            counts, ctrs = None, {}
            for eng in engines:
                c_ = eng.counts() if world == 1 else parallel.global_counts(eng)
                counts = c_ if counts is None else {k: counts[k] + c_[k] for k in counts}
                for k, v in eng.counters().items():
                    ctrs[k] = ctrs.get(k, 0) + v
            self.counters = ctrs
        finally:
            for eng in engines:
                eng.close()
        cat = {k: np.concatenate([t[k] for t in traces], axis=1) for k in traces[0]} if traces else None
        return N, counts, cat, first

    # ---- the reference's per-node calls (network.py:143-396) ---------------------------------------------
    def fit(self, method):
        """network.py:143-281: run the regressor of `method` on the CURRENT configuration (the node of the last
        set_network_configuration) and set `chain_results`."""
        if self._current is None:
            raise RuntimeError('fit: call set_network_configuration(i) first')
        mid, name = self._resolve(method)
        n = self.data[0].shape[1]
        node = self._current
        N, counts, cat, first = self._run(mid, name, node)
        self._node_counts = (node, N, counts)
        if cat is not None:
            u = 0      # chain 0 of the node's slice
            self.chain_results = (self._chain_results_fp(cat, u, mid, n, node) if mid in FP_METHODS
                                  else self._chain_results(cat, u, mid, n, name, node, int(self.lag)))
        else:
            self.chain_results = None
        return self

    def score_edges(self, currResponse, method):
        """network.py:283-396: edge scores of the current configuration -> one more row of `proposed_adj_matrix`; the
        thinned changepoint chain and the per-time-point beta diagnostics when histories were kept."""
        mid, name = self._resolve(method)
        if self._node_counts is None or self._node_counts[0] != currResponse:
            raise RuntimeError('score_edges(%d): call set_network_configuration(%d) and fit(method) first' % (currResponse, currResponse))
        node, N, counts = self._node_counts
        self._score_node(mid, name, node, N, counts, self.chain_results)

    def _score_node(self, mid, name, node, N, counts, res):
        n = self.data[0].shape[1]
        if mid in FP_METHODS:
            adj = parallel.credible_scores_from_counts(counts['edge'], counts['neg'], counts['nonempty'])
            self.network_args['scoring_method'] = 'fraq-score'
        else:
            adj = parallel.edge_scores_from_counts(counts['edge'], counts['nonempty'])
            self.network_args['scoring_method'] = 'edge-scores'
        row = [float(v) for v in adj[node]]
        if int(self.lag) > 1:      # calculateFeatureScores returns len(features) + 1 entries, the genes first (scores.py:156,192)
            row = row + [0.0] * (int(self.lag) * (n - 1) + 1 - n)
        self.proposed_adj_matrix.append(row)
        self.edge_scores = row
        if res is not None:
            self._diagnostics(mid, name, node, N, res)

    def _diagnostics(self, mid, name, node, N, res):
        """cps_over_response, betas_over_time, scores_over_time of one node from its kept histories
        (network.py:316-338, :370-396)."""
        n = self.data[0].shape[1]
        feats = [j for j in range(n) if j != node]
        b, th = int(self.burn_in), self.thin
        thin_eq = lambda chain: [chain[x] for x in range(len(chain)) if x % th == 0]
        key = 'betas_vector' if mid in FP_METHODS else 'padded_betas'
        betas = thin_eq(res[key][2:][b:])
        as_list = lambda v: v if isinstance(v, list) else [v]
        betas = [as_list(v) for v in betas]
        if len(res['tau_vector']) == 0:
            # homogeneous model (and 'fixed_nh_dbn', whose regressor never records tau): one artificial changepoint
            cps = [[N + 2] for _ in range(len(betas))] if mid != FP_H_DBN else []
        else:
            cps = thin_eq(res['tau_vector'][b:])
        self.cps_over_response.append(cps)
        want = self.diagnostics
        if want is None:
            want = N * max(len(betas), 1) * n * 8 * n <= DIAGNOSTIC_BYTES
        if not want or mid == FP_H_DBN:      # network.py:329-333: no betas over time for 'fp_h_dbn'
            return
        # network.py:386-394: beta_dim = len(currFeatures) + 1 (lag copies included) for the varying-parents methods
        bot = S.get_betas_over_time(N, cps, betas, n if mid in FP_METHODS else int(self.lag) * (n - 1) + 1)
        self.betas_over_time.append(bot)
        if mid in FP_METHODS:
            self.scores_over_time.append(S.get_scores_over_time(bot, feats, n))

    # ---- the hot path ----------------------------------------------------------------------------------
    def infer_network(self, method):
        """network.py:398-416: all nodes (and all chains) advance together on the device instead of one regressor per
        node in a Python loop; edge scores are the thinned-chain frequencies of scores.py:155-194."""
        mid, name = self._resolve(method)
        dims = self.data[0].shape[1]
        for i in range(dims):
            self.set_network_configuration(i)
        N, counts, cat, first = self._run(mid, name, None)
        kept = np.maximum(counts['cp_kept'], 1)[:, None]
        self.changepoint_histogram = (counts['cp_hist'] / kept)[:, :N + 1]
        C = self.n_chains
        for node in range(dims):
            res = None
            if cat is not None:
                u = node * C - first
                res = (self._chain_results_fp(cat, u, mid, dims, node) if mid in FP_METHODS
                       else self._chain_results(cat, u, mid, dims, name, node, int(self.lag)))
                self.chain_results_per_node.append(res)
            self._score_node(mid, name, node, N, counts, res)
        if cat is not None:
            self.chain_results = self.chain_results_per_node[-1]
        elif self.keep_chain is None and self.n_chains == 1:
            warnings.warn('per-iteration histories were not kept (they would not fit %d GB); chain_results is None. '
                          'Pass keep_chain=True to force them.' % (KEEP_CHAIN_BYTES >> 30))
        return self

    # ---- histories in the reference's result-dict form ---------------------------------------------------
    def _chain_results_fp(self, cat, u, mid, dims, node):
        """fpGlobCoupBpwLinReg.py:112-119 and siblings: lists of length num_iter + 1 starting from the initial state."""
        n_iter = cat['sigma2'].shape[1]
        pi = [j for j in range(dims) if j != node]
        res = {
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
        return res

    @staticmethod
    def _chain_results(t, u, mid, dims, name, node=0, lag=1):
        """Per-iteration histories of one unit as the reference's result dict (bayesianPwLinearRegression.py:134-141):
        lists of length num_iter + 1 starting from the initial state.  The device works on feature COLUMNS (column
        l * n + g = lag-(l + 1) copy of gene g); the reference labels the lag-1 block by gene id and every further block
        by a running counter over the node's features (network.py:92-104), which is what `pi_vector` carries."""
        n_iter = t['sigma2'].shape[1]
        F = lag * (dims - 1)

        def label(col):
            col = int(col)
            if col < dims:
                return col
            blk, g = divmod(col, dims)
            return dims + (blk - 1) * (dims - 1) + (g if g < node else g - 1)
        pi_vec = [[]] + [np.array([label(c) for c in t['pi_after'][u, it, :t['npi'][u, it]]], dtype=np.int64) for it in range(n_iter)]
        fixed = name in FIXED_CP
        res = {
            'lambda_sqr_vector': [1] + [float(v) for v in t['lambda2'][u]],
            'sigma_sqr_vector': [1] + [float(v) for v in t['sigma2'][u]],
            'pi_vector': pi_vec,
            # only the 'varying_nh' regressor records its changepoints (bayesianPwLinearRegression.py:117-121)
            'tau_vector': [] if (mid == H_DBN or fixed) else
                          [[int(c) for c in t['tau_after'][u, it, :t['S_after'][u, it]]] for it in range(n_iter)],
        }
        homog_like = mid in (H_DBN, GLOB_DBN, VV_GLOB_DBN)   # their initial padded entry is a bare vector (e.g. globCoupBayesianPwLinReg.py:58)
        betas = [[np.zeros(1)]]
        padded = [np.zeros(F + 1) if homog_like else [np.zeros(F + 1)]]
        for it in range(n_iter):
            Sn = 1 if mid == H_DBN else int(t['S'][u, it])
            k = 1 if it == 0 else int(t['npi'][u, it - 1]) + 1
            sample = [np.array(t['beta'][u, it, h, :k]) for h in range(Sn)]
            betas.append(sample)
            padded.append(S.transform_beta_coef(sample, pi_vec[it], F))   # the parent set the draw was conditioned on
        res['betas_vector'] = betas
        res['padded_betas'] = padded
        if mid == SEQ_DBN:
            res['delta_sqr_vector'] = [1] + [float(v) for v in t['delta2'][u]]
        if mid == VV_GLOB_DBN:   # vvglobCoup.py:59,70-72,125: one variance per segment and iteration
            res['sigma_sqr_vector'] = [[1]] + [[float(v) for v in t['sigma2v'][u, it, :int(t['S'][u, it])]] for it in range(n_iter)]
        if mid in (GLOB_DBN, VV_GLOB_DBN):
            res['mu_vector'] = [np.zeros((1, 1))] + [np.array(t['mu_after'][u, it, :t['npi'][u, it] + 1]).reshape(-1, 1) for it in range(n_iter)]
        return res
