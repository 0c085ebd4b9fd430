"""Host-side handle on the CUDA library: one engine = one device + one stream (include/dbn_b200.h)."""
import ctypes as C

import numpy as np

from . import _lib as L
from .methods import run_params


class DbnError(RuntimeError):
    pass


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class DbnEngine:
    """Device-resident RJ-MCMC engine for all (node, chain) units of one network.

    Replaces the reference's per-node regressor objects (`Cls(data_dict, _type, num_samples, num_iter,
    change_points).fit()`, bayesianPwLinearRegression.py:35-141 and siblings): `build_stats` takes the same list of
    time-series segments `Network` receives, `chain_init(method)` + `run(n_iter)` advance every node and chain.
    """

    def __init__(self, n_chains=1, device=0, seed=0xDB20B200, chain_offset=0, unit_range=None):
        """`unit_range` = (first, count): this engine runs only that contiguous slice of the (node, chain) grid
        u = node * n_chains + chain (dbn_set_unit_range); state / trace / replay arrays are then indexed by the local
        unit.  None: the whole grid."""
        self._lib = L.load()
        self._h = C.c_void_p()
        cfg = L.DbnConfig(int(device), int(n_chains), int(chain_offset), int(seed))
        rc = self._lib.dbn_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            msg = self._lib.dbn_last_error(None)
            self._h = None
            raise DbnError('dbn_create failed (%d): %s' % (rc, (msg or b'').decode()))
        self.n_chains = int(n_chains)
        self.device = int(device)
        self.unit_range = None if unit_range is None else (int(unit_range[0]), int(unit_range[1]))
        self.n = self.N = None
        if self.unit_range is not None:     # before build_stats: the table then holds only the ZY | YY blocks of the range's nodes
            self._check(self._lib.dbn_set_unit_range(self._h, self.unit_range[0], self.unit_range[1]), 'dbn_set_unit_range')
        self.params = None

    # ---- plumbing
    def _check(self, rc, what):
        if rc != 0:
            raise DbnError('%s failed (%d): %s' % (what, rc, (self._lib.dbn_last_error(self._h) or b'').decode()))

    def close(self):
        if getattr(self, '_h', None):
            self._lib.dbn_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_stream(self, cuda_stream_ptr):
        self._check(self._lib.dbn_set_stream(self._h, C.c_void_p(cuda_stream_ptr or 0)), 'dbn_set_stream')

    def sync(self):
        self._check(self._lib.dbn_sync(self._h), 'dbn_sync')

    # ---- kernel 1
    def build_stats(self, data, lag=1):
        """`data`: list of [len_s, n] arrays (what Network(data=...) receives, network.py:28)."""
        segs = [np.ascontiguousarray(np.asarray(s, dtype=np.float64)) for s in data]
        n = segs[0].shape[1]
        if any(s.ndim != 2 or s.shape[1] != n for s in segs):
            raise ValueError('all data segments must be [len, n] with the same n')
        # a single series is passed as it is (no host copy: a pinned caller buffer stays pinned for the H2D copy)
        flat = segs[0] if len(segs) == 1 else np.ascontiguousarray(np.concatenate(segs, axis=0))
        seg_len = np.array([s.shape[0] for s in segs], dtype=np.int32)
        self._check(self._lib.dbn_build_stats(self._h, _dptr(flat), _iptr(seg_len), len(segs), n, int(lag)), 'dbn_build_stats')
        self.n = n
        self.lag = int(lag)
        self.nv = n * int(lag)      # candidate feature columns: column l * n + g is the lag-(l + 1) copy of gene g
        self.N = int(sum(s.shape[0] - 1 for s in segs))
        self.h2d_bytes = flat.nbytes
        return self

    def rebuild_stats(self):
        self._check(self._lib.dbn_rebuild_stats(self._h), 'dbn_rebuild_stats')

    def stats_info(self):
        N, R, B = C.c_int32(), C.c_int64(), C.c_int64()
        self._check(self._lib.dbn_stats_info(self._h, C.byref(N), C.byref(R), C.byref(B)), 'dbn_stats_info')
        return dict(N=N.value, row_doubles=R.value, algorithmic_bytes=B.value)

    def stats_row(self, t):
        """Table row t split into ZZ[nv+1, nv+1] (symmetric), ZY[nodes, nv+1] (node, column), YY[nodes]; nv = n * lag.
        `nodes` = every node, or the nodes of the engine's unit range (the table holds only their blocks)."""
        info = self.stats_info()
        row = np.empty(info['row_doubles'])
        self._check(self._lib.dbn_read_stats_row(self._h, int(t), _dptr(row)), 'dbn_read_stats_row')
        n, nv = self.n, self.nv
        zz = np.zeros((nv + 1, nv + 1))
        zz[np.tril_indices(nv + 1)] = row[:(nv + 1) * (nv + 2) // 2]
        zz = zz + np.tril(zz, -1).T
        blocks = row[(nv + 1) * (nv + 2) // 2:].reshape(-1, nv + 2)
        return zz, blocks[:, :nv + 1].copy(), blocks[:, nv + 1].copy()

    # ---- kernel 2
    def score_batch(self, node, pi, tau, lambda2, delta2=None, mu=None, alpha=0.005, beta=0.005, T=None, seq=False):
        """Log marginal likelihoods of a batch of (node, parent set, changepoint list) states.

        pi / tau are lists of lists (gene ids; 1-based changepoints ending with N+2); mu is None (zeros) or a list of
        per-item vectors in design-column order.  Replaces marginalLikelihood.py:91-166.
        """
        B = len(node)
        node_a = np.asarray(node, dtype=np.int32)
        pi_a = np.full((B, L.PMAX), -1, dtype=np.int32)
        pl = np.zeros(B, dtype=np.int32)
        tau_a = np.full((B, L.SMAX), -1, dtype=np.int32)
        tl = np.zeros(B, dtype=np.int32)
        mu_a = np.zeros((B, L.KMAX))
        for b in range(B):
            if len(pi[b]) > L.PMAX or len(tau[b]) > L.SMAX:
                raise ValueError('item %d exceeds the padded widths' % b)
            pl[b] = len(pi[b]); pi_a[b, :pl[b]] = pi[b]
            tl[b] = len(tau[b]); tau_a[b, :tl[b]] = tau[b]
            if mu is not None and mu[b] is not None:
                m = np.asarray(mu[b], dtype=np.float64).ravel()
                mu_a[b, :m.size] = m
        lam = np.ascontiguousarray(np.broadcast_to(np.asarray(lambda2, dtype=np.float64), (B,)))
        dlt = np.ascontiguousarray(np.broadcast_to(np.asarray(1.0 if delta2 is None else delta2, dtype=np.float64), (B,)))
        out = np.empty(B)
        T = float(self.N if T is None else T)
        self._check(self._lib.dbn_score_batch(self._h, int(bool(seq)), float(alpha), float(beta), T, B, _iptr(node_a), _iptr(pi_a),
                                              _iptr(pl), _iptr(tau_a), _iptr(tl), _dptr(lam), _dptr(dlt), _dptr(mu_a), _dptr(out)),
                    'dbn_score_batch')
        return out

    # ---- kernel 3
    def chain_init(self, method, **kw):
        p = run_params(method, self.N, **kw)
        self.params = p
        self.method = method
        rp = L.DbnRunParams(**p)
        if self.unit_range is not None:
            self._check(self._lib.dbn_set_unit_range(self._h, self.unit_range[0], self.unit_range[1]), 'dbn_set_unit_range')
        self._check(self._lib.dbn_chain_init(self._h, C.byref(rp)), 'dbn_chain_init')
        self.units = self.n * self.n_chains if self.unit_range is None else self.unit_range[1]
        return self

    def set_changepoints(self, tau):
        """dbn_set_changepoints: start every unit from the changepoint list `tau` (ends with the pseudo changepoint N + 2)."""
        t = np.ascontiguousarray(np.asarray(tau, dtype=np.int32).ravel())
        self._check(self._lib.dbn_set_changepoints(self._h, _iptr(t), int(t.size)), 'dbn_set_changepoints')
        return self

    def run(self, n_iter, replay=None, trace=False):
        """Advance every unit by n_iter iterations.  `replay`: dict of recorded draws (see pack_replay);
        returns the unpacked trace dict when trace (or replay) is requested, else None (asynchronous)."""
        U = self.units
        if replay is None and not trace:
            self._check(self._lib.dbn_run(self._h, int(n_iter), None, None, None, None), 'dbn_run')
            return None
        td = np.empty((U, n_iter, L.TD_STRIDE))
        ti = np.empty((U, n_iter, L.TI_STRIDE), dtype=np.int32)
        if replay is not None:
            rd, ri = pack_replay(replay, U, n_iter)
            self._check(self._lib.dbn_run(self._h, int(n_iter), _dptr(rd), _iptr(ri), _dptr(td), _iptr(ti)), 'dbn_run')
        else:
            self._check(self._lib.dbn_run(self._h, int(n_iter), None, None, _dptr(td), _iptr(ti)), 'dbn_run')
        return unpack_trace(td, ti)

    def run_fp(self, n_iter, replay=None, trace=False):
        """dbn_run_fp: advance every unit of a full-parents chain ('fp_glob_coup_nh_dbn', 'fp_varying_nh_dbn', 'fp_h_dbn').
        `replay`: dict with sigma2, lambda2, u_tau [U, n_iter]; mu_tau [U, n_iter, k]; beta [U, n_iter, SMAX, k];
        tau_move [U, n_iter]; tau_pick [U, n_iter, 2].  Returns the unpacked trace dict when requested."""
        U, k = self.units, self.n
        if replay is None and not trace:
            self._check(self._lib.dbn_run_fp(self._h, int(n_iter), None, None, None, None), 'dbn_run_fp')
            return None
        tds = L.fp_td_stride(k)
        td = np.empty((U, n_iter, tds))
        ti = np.empty((U, n_iter, L.FP_TI_STRIDE), dtype=np.int32)
        if replay is not None:
            rd = np.full((U, n_iter, L.fp_rd_stride(k)), np.nan)
            ri = np.full((U, n_iter, L.FP_RI_STRIDE), -1, dtype=np.int32)
            rd[:, :, 0], rd[:, :, 1], rd[:, :, 2] = replay['sigma2'], replay['lambda2'], replay['u_tau']
            if 'mu_tau' in replay:     # globally coupled only
                rd[:, :, 3:3 + k] = np.asarray(replay['mu_tau'])[:, :, :k]
            rd[:, :, 3 + k:3 + k + L.SMAX * k] = np.asarray(replay['beta'])[:, :, :, :k].reshape(U, n_iter, L.SMAX * k)
            if 'sigma2v' in replay:    # 'fp_var_glob_coup_nh_dbn': one variance per segment
                rd[:, :, 3 + k + L.SMAX * k:] = replay['sigma2v']
            if 'delta2' in replay:     # 'fp_seq_coup_nh_dbn': delta^2 travels in the first per-segment variance slot
                rd[:, :, 3 + k + L.SMAX * k] = replay['delta2']
            ri[:, :, 0] = replay['tau_move']
            ri[:, :, 1:3] = replay['tau_pick']
            rd, ri = np.ascontiguousarray(rd), np.ascontiguousarray(ri)
            self._check(self._lib.dbn_run_fp(self._h, int(n_iter), _dptr(rd), _iptr(ri), _dptr(td), _iptr(ti)), 'dbn_run_fp')
        else:
            self._check(self._lib.dbn_run_fp(self._h, int(n_iter), None, None, _dptr(td), _iptr(ti)), 'dbn_run_fp')
        o = 10
        return dict(
            sig_shape=td[:, :, 0], sig_rate=td[:, :, 1], lam_shape=td[:, :, 2], lam_rate=td[:, :, 3],
            sigma2=td[:, :, 4], lambda2=td[:, :, 5], logml=td[:, :, 6:8], mus_logdens=td[:, :, 8:10],
            mu_after=td[:, :, o:o + k], mus_mean=td[:, :, o + k:o + 3 * k].reshape(U, n_iter, 2, k),
            beta_mean=td[:, :, o + 3 * k:o + 3 * k + L.SMAX * k].reshape(U, n_iter, L.SMAX, k),
            beta_cov_diag=td[:, :, o + 3 * k + L.SMAX * k:o + 3 * k + 2 * L.SMAX * k].reshape(U, n_iter, L.SMAX, k),
            beta=td[:, :, o + 3 * k + 2 * L.SMAX * k:o + 3 * k + 3 * L.SMAX * k].reshape(U, n_iter, L.SMAX, k),
            sig_shape_v=td[:, :, o + 3 * k + 3 * L.SMAX * k:o + 3 * k + 3 * L.SMAX * k + L.SMAX],
            sig_rate_v=td[:, :, o + 3 * k + 3 * L.SMAX * k + L.SMAX:o + 3 * k + 3 * L.SMAX * k + 2 * L.SMAX],
            sigma2v=td[:, :, o + 3 * k + 3 * L.SMAX * k + 2 * L.SMAX:],
            tau_move=ti[:, :, 0], tau_valid=ti[:, :, 1], tau_accept=ti[:, :, 2], S=ti[:, :, 3], S_after=ti[:, :, 4],
            tau_after=ti[:, :, 5:5 + L.SMAX])

    def counts(self, out=None):
        """Count matrices of the thinned chain (views of one flat int64 block).  `out`: a caller-owned flat int64 array of
        dbn_counts_size elements to read into -- a PINNED one is written by the copy engine directly (no bounce buffer, no
        fresh pages per call); the returned matrices are views of it."""
        ne = C.c_int64()
        self._check(self._lib.dbn_counts_size(self._h, C.byref(ne)), 'dbn_counts_size')
        if out is None:
            flat = np.empty(ne.value, dtype=np.int64)
        else:
            flat = out
            if flat.dtype != np.int64 or flat.ndim != 1 or flat.size != ne.value or not flat.flags['C_CONTIGUOUS']:
                raise ValueError('counts(out=): need a contiguous flat int64 array of %d elements' % ne.value)
        self._check(self._lib.dbn_read_counts(self._h, flat.ctypes.data_as(C.POINTER(C.c_int64))), 'dbn_read_counts')
        return split_counts(flat, self.n, self.N)

    def counts_device_ptr(self):
        p = C.c_void_p()
        ne = C.c_int64()
        self._check(self._lib.dbn_device_counts(self._h, C.byref(p)), 'dbn_device_counts')
        self._check(self._lib.dbn_counts_size(self._h, C.byref(ne)), 'dbn_counts_size')
        return p.value, ne.value

    def state(self):
        U = self.units
        pi = np.empty((U, L.PMAX), dtype=np.int32); pl = np.empty(U, dtype=np.int32)
        tau = np.empty((U, L.SMAX), dtype=np.int32); tl = np.empty(U, dtype=np.int32)
        sc = np.empty((U, 3)); mu = np.empty((U, L.KMAX))
        self._check(self._lib.dbn_read_state(self._h, _iptr(pi), _iptr(pl), _iptr(tau), _iptr(tl), _dptr(sc), _dptr(mu)), 'dbn_read_state')
        return dict(pi=pi, npi=pl, tau=tau, S=tl, sigma2=sc[:, 0], lambda2=sc[:, 1], delta2=sc[:, 2], mu=mu)

    def counters(self):
        out = np.zeros(8, dtype=np.uint64)
        self._check(self._lib.dbn_read_counters(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64))), 'dbn_read_counters')
        keys = ('unit_iterations', 'ml_evals', 'segment_evals', 'algorithmic_bytes', 'algorithmic_flops', 'pi_accepts',
                'segment_factorisations', 'tau_accepts')
        d = {k: int(out[i]) for i, k in enumerate(keys)}
        if self.params is not None and str(getattr(self, 'method', '')).startswith('fp_'):
            # full-parents kernels: slot 5 = factorisations that also formed the inverse factor and I - A^-1, slot 6 = all k x k
            # Cholesky factorisations (segments + the muSampler precision): executed level-3 flops = k^3 / 3 per factor + 2 k^3 / 3 per inverse form
            k = self.n
            d['inverse_form_evals'] = d.pop('pi_accepts')
            d['executed_level3_flops'] = (d['segment_factorisations'] + 2 * d['inverse_form_evals']) * (k ** 3) // 3
        return d

    def verify_cache(self):
        """dbn_verify_cache: (units, entries) of the h / nh kernel's segment-statistics cache that differ from what the
        prefix table and the current (pi, tau) imply; (0, 0) when consistent."""
        out = np.zeros(2, dtype=np.uint64)
        self._check(self._lib.dbn_verify_cache(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64))), 'dbn_verify_cache')
        return int(out[0]), int(out[1])

    def reset_counters(self):
        self._check(self._lib.dbn_reset_counters(self._h), 'dbn_reset_counters')

    def score_batch_ms(self):
        """dbn_score_batch_time: device time of the scorer kernel of the last score_batch call."""
        g = C.c_double()
        self._check(self._lib.dbn_score_batch_time(self._h, C.byref(g)), 'dbn_score_batch_time')
        return g.value

    def fp64_peak_gflops(self):
        g = C.c_double()
        self._check(self._lib.dbn_measure_fp64_peak(self._h, C.byref(g)), 'dbn_measure_fp64_peak')
        return g.value


def split_counts(flat, n, N):
    """Contiguous count block -> named matrices (layout documented at dbn_read_counts in include/dbn_b200.h)."""
    o = 0
    edge = flat[o:o + n * n].reshape(n, n); o += n * n
    nonempty = flat[o:o + n]; o += n
    cp_hist = flat[o:o + n * (N + 3)].reshape(n, N + 3); o += n * (N + 3)
    cp_kept = flat[o:o + n]; o += n
    out = dict(edge=edge, nonempty=nonempty, cp_hist=cp_hist, cp_kept=cp_kept)
    if flat.size >= o + n * n:   # full-parents method: edge = #(mu_j >= 0), nonempty = #kept, neg = #(mu_j <= 0)
        out['neg'] = flat[o:o + n * n].reshape(n, n)
    return out


def pack_replay(r, U, n_iter):
    """Named arrays [U, n_iter, ...] -> the ABI's replay records.  NaN / -1 entries are 'not drawn'."""
    rd = np.full((U, n_iter, L.RD_STRIDE), np.nan)
    ri = np.full((U, n_iter, L.RI_STRIDE), -1, dtype=np.int32)
    rd[:, :, L.RD_SIGMA2] = r['sigma2']
    rd[:, :, L.RD_LAMBDA2] = r['lambda2']
    if 'delta2' in r:
        rd[:, :, L.RD_DELTA2] = r['delta2']
    rd[:, :, L.RD_U_PI] = r['u_pi']
    if 'u_tau' in r:
        rd[:, :, L.RD_U_TAU] = r['u_tau']
    if 'mu_pi' in r:
        rd[:, :, L.RD_MU_PI:L.RD_MU_PI + L.KMAX] = r['mu_pi']
        rd[:, :, L.RD_MU_TAU:L.RD_MU_TAU + L.KMAX] = r['mu_tau']
    rd[:, :, L.RD_BETA:L.RD_BETA + L.SMAX * L.KMAX] = np.asarray(r['beta']).reshape(U, n_iter, L.SMAX * L.KMAX)
    if 'sigma2v' in r:      # segment-specific variances of 'var_glob_coup_nh_dbn'
        rd[:, :, L.RD_SIGMA2V:L.RD_SIGMA2V + L.SMAX] = r['sigma2v']
    ri[:, :, L.RI_PI_MOVE] = r['pi_move']
    ri[:, :, L.RI_PI_PICK:L.RI_PI_PICK + 2] = r['pi_pick']
    if 'tau_move' in r:
        ri[:, :, L.RI_TAU_MOVE] = r['tau_move']
        ri[:, :, L.RI_TAU_PICK:L.RI_TAU_PICK + 2] = r['tau_pick']
    return np.ascontiguousarray(rd), np.ascontiguousarray(ri)


def _unpack_tri(a):
    """[..., KTRI] packed lower triangles -> [..., KMAX, KMAX] symmetric."""
    out = np.full(a.shape[:-1] + (L.KMAX, L.KMAX), np.nan)
    il, jl = np.tril_indices(L.KMAX)
    out[..., il, jl] = a
    out[..., jl, il] = a
    return out


def unpack_trace(td, ti):
    U, n_iter = td.shape[:2]
    t = dict(
        sig_shape=td[:, :, L.TD_SIG_SHAPE], sig_rate=td[:, :, L.TD_SIG_RATE],
        lam_shape=td[:, :, L.TD_LAM_SHAPE], lam_rate=td[:, :, L.TD_LAM_RATE],
        del_shape=td[:, :, L.TD_DEL_SHAPE], del_rate=td[:, :, L.TD_DEL_RATE],
        logml=td[:, :, L.TD_LOGML:L.TD_LOGML + 4],
        sigma2=td[:, :, L.TD_SIGMA2], lambda2=td[:, :, L.TD_LAMBDA2], delta2=td[:, :, L.TD_DELTA2],
        mu_after=td[:, :, L.TD_MU:L.TD_MU + L.KMAX],
        mus_logdens=td[:, :, L.TD_MUS_LOGDENS:L.TD_MUS_LOGDENS + 4],
        mus_mean=td[:, :, L.TD_MUS_MEAN:L.TD_MUS_MEAN + 4 * L.KMAX].reshape(U, n_iter, 4, L.KMAX),
        mus_cov=_unpack_tri(td[:, :, L.TD_MUS_COV:L.TD_MUS_COV + 4 * L.KTRI].reshape(U, n_iter, 4, L.KTRI)),
        beta_mean=td[:, :, L.TD_BETA_MEAN:L.TD_BETA_MEAN + L.SMAX * L.KMAX].reshape(U, n_iter, L.SMAX, L.KMAX),
        beta_cov=_unpack_tri(td[:, :, L.TD_BETA_COV:L.TD_BETA_COV + L.SMAX * L.KTRI].reshape(U, n_iter, L.SMAX, L.KTRI)),
        beta=td[:, :, L.TD_BETA:L.TD_BETA + L.SMAX * L.KMAX].reshape(U, n_iter, L.SMAX, L.KMAX),
        sig_shape_v=td[:, :, L.TD_SIGV_SHAPE:L.TD_SIGV_SHAPE + L.SMAX], sig_rate_v=td[:, :, L.TD_SIGV_RATE:L.TD_SIGV_RATE + L.SMAX],
        sigma2v=td[:, :, L.TD_SIGMA2V:L.TD_SIGMA2V + L.SMAX],
        pi_valid=ti[:, :, L.TI_PI_VALID], pi_accept=ti[:, :, L.TI_PI_ACCEPT],
        tau_valid=ti[:, :, L.TI_TAU_VALID], tau_accept=ti[:, :, L.TI_TAU_ACCEPT],
        npi=ti[:, :, L.TI_NPI], pi_after=ti[:, :, L.TI_PI:L.TI_PI + L.PMAX],
        S_after=ti[:, :, L.TI_S], tau_after=ti[:, :, L.TI_TAU:L.TI_TAU + L.SMAX],
        pi_move=ti[:, :, L.TI_PI_MOVE], tau_move=ti[:, :, L.TI_TAU_MOVE], S=ti[:, :, L.TI_S_BEFORE],
    )
    return t
