"""Shared helpers for the parity tests (loading golden fixtures, tolerances)."""
import glob
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')

# north_star: "log marginal likelihoods match the reference implementation within 1e-9 relative in FP64"
LOGML_RTOL = 1e-9

KAT_FILES = sorted(glob.glob(os.path.join(GOLDEN, 'kat_logml_*.npz')))
_ALL_CHAINS = sorted(glob.glob(os.path.join(GOLDEN, 'chain_*.npz')))
_FP_ALL = [p for p in _ALL_CHAINS if '_fp_' in os.path.basename(p)]              # full-parents methods (n-wide records)
FP_CHAIN_FILES = [p for p in _FP_ALL if 'fp_glob' in os.path.basename(p)]        # globally coupled: scored from the mu chain
FP_VV_FILES = [p for p in _FP_ALL if 'fp_var_glob' in os.path.basename(p)]       # segment-specific variances, mu redrawn every iteration
FP_SEQ_FILES = [p for p in _FP_ALL if 'fp_seq' in os.path.basename(p)]           # sequentially coupled at width n (oracle only so far)
FP_PLAIN_FILES = [p for p in _FP_ALL if p not in FP_CHAIN_FILES + FP_VV_FILES + FP_SEQ_FILES]   # fp_varying_nh / fp_h: scored from the beta chain
CHAIN_FILES = [p for p in _ALL_CHAINS if p not in _FP_ALL]
# lag-2 designs (network.py:92-122): parents are LABELS (gene id for the lag-1 block, n + position for the lag-2 block)
LAG_CHAIN_FILES = sorted(glob.glob(os.path.join(GOLDEN, 'lag*chain_*.npz')))


def load(path):
    return dict(np.load(path, allow_pickle=False))


def data_list(g):
    return [g['data_seg%d' % s] for s in range(int(g['n_data_segments']))]


def ids(paths):
    return [os.path.basename(p)[:-4] for p in paths]


def assert_close(a, b, rtol, atol=0.0, what=''):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    fa, fb = np.isfinite(a), np.isfinite(b)
    assert (fa == fb).all(), '%s: finite masks differ at %s' % (what, np.argwhere(fa != fb)[:5])
    if not fa.any():
        return
    err = np.abs(a[fa] - b[fa])
    tol = atol + rtol * np.abs(b[fb])
    bad = err > tol
    assert not bad.any(), '%s: max rel err %.3e (rtol %.1e) at %d/%d entries' % (
        what, float((err / np.maximum(np.abs(b[fb]), 1e-300)).max()), rtol, int(bad.sum()), bad.size)


def labels_to_columns(pi_after, n):
    """[n, n_iter, k] parent LABELS of a lag > 1 golden -> feature columns (column l * n + g = lag-(l + 1) copy of gene g)."""
    out = np.array(pi_after, dtype=np.int32)
    for node in range(out.shape[0]):
        lab = out[node]
        m = lab >= n
        blk, j = np.divmod(lab[m] - n, n - 1)
        lab[m] = (blk + 1) * n + np.where(j < node, j, j + 1)
    return out


def golden_replay(g):
    """Replay dict (engine.pack_replay layout, unit = node, one chain) from a chain golden file."""
    n, n_iter = g['sigma2'].shape
    K, S = 5, 10
    r = dict(sigma2=g['sigma2'], lambda2=g['lambda2'], delta2=g['delta2'], u_pi=g['u_pi'], u_tau=g['u_tau'],
             beta=g['beta'], pi_move=g['pi_move'], pi_pick=g['pi_pick'], tau_move=g['tau_move'], tau_pick=g['tau_pick'])
    if 'mus_val' in g:
        r['mu_pi'] = g['mus_val'][:, :, 0, :]
        r['mu_tau'] = g['mus_val'][:, :, 3, :]
    if 'sigma2v' in g:      # var_glob_coup_nh_dbn: one variance per segment
        r['sigma2v'] = g['sigma2v']
    return r


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def assert_close_mat(a, b, rtol, what=''):
    """Matrix-wise tolerance: |a - b| <= rtol * max|b| over the trailing two axes (covariances: entries that are
    tiny relative to the matrix scale carry cancellation noise in the reference's own np.linalg.inv)."""
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    fa, fb = np.isfinite(a), np.isfinite(b)
    assert (fa == fb).all(), '%s: finite masks differ at %s' % (what, np.argwhere(fa != fb)[:5])
    scale = np.nanmax(np.where(fb, np.abs(b), 0.0), axis=(-1, -2), keepdims=True)
    err = np.where(fb, np.abs(np.where(fa, a, 0.0) - np.where(fb, b, 0.0)), 0.0)
    bad = err > rtol * scale
    assert not bad.any(), '%s: max matrix-relative err %.3e (rtol %.1e) at %d entries' % (
        what, float((err / np.maximum(scale, 1e-300)).max()), rtol, int(bad.sum()))


# ---------------------------------------------------------------------------------------------------------
# oracle chains with every draw recorded -> replay records for the device (the fixtures the reference cannot give:
# fan-in 4, T beyond its overflow limit)
# ---------------------------------------------------------------------------------------------------------
def recording_draws(rng):
    """oracle.RngDraws that logs every draw so that the same stream can be injected into the device chain."""
    from oracle import dbn_oracle as O

    class Rec(O.RngDraws):
        def __init__(self, rng_):
            super().__init__(rng_)
            self.log = {}

        def inv_gamma(self, it, which, shape, rate):
            v = super().inv_gamma(it, which, shape, rate); self.log[(it, which)] = v; return v

        def inv_gamma_seg(self, it, h, shape, rate):
            v = super().inv_gamma_seg(it, h, shape, rate); self.log[(it, 'sigv', h)] = v; return v

        def move(self, it, which):
            v = super().move(it, which); self.log[(it, which, 'move')] = v; return v

        def picker(self, it, which):
            inner = super().picker(it, which); lst = self.log.setdefault((it, which, 'pick'), [])

            def f(m):
                v = inner(m); lst.append(v); return v
            return f

        def uniform(self, it, which):
            v = super().uniform(it, which); self.log[(it, which, 'u')] = v; return v

        def mu_draw(self, it, slot, mean, Sigma):
            v = super().mu_draw(it, slot, mean, Sigma); self.log[(it, 'mu', slot)] = v; return v
    return Rec(rng)


def oracle_replay(n_units, n_iter, chains):
    """Replay dict (engine.pack_replay layout) for `n_units` units from recorded oracle chains `chains` = {unit: (trace,
    Rec)}.  Units without a recorded chain get a stream of moves that are always valid to execute: delete / death with
    pick 0 (invalid once the parent set is empty / only the pseudo changepoint is left), unit Gibbs draws."""
    K, S = 5, 10
    r = dict(sigma2=np.ones((n_units, n_iter)), lambda2=np.ones((n_units, n_iter)), delta2=np.ones((n_units, n_iter)),
             u_pi=np.full((n_units, n_iter), 0.5), u_tau=np.full((n_units, n_iter), 0.5),
             beta=np.zeros((n_units, n_iter, S, K)), pi_move=np.ones((n_units, n_iter), dtype=np.int32),
             pi_pick=np.zeros((n_units, n_iter, 2), dtype=np.int32), tau_move=np.ones((n_units, n_iter), dtype=np.int32),
             tau_pick=np.zeros((n_units, n_iter, 2), dtype=np.int32), mu_pi=np.zeros((n_units, n_iter, K)),
             mu_tau=np.zeros((n_units, n_iter, K)), sigma2v=np.ones((n_units, n_iter, S)))
    for u, (t, rec) in chains.items():
        r['sigma2'][u], r['lambda2'][u] = t['sigma2'], t['lambda2']
        r['delta2'][u] = np.nan_to_num(t['delta2'], nan=1.0)
        r['beta'][u] = np.nan_to_num(t['beta'][:, :, :K])
        r['sigma2v'][u] = np.nan_to_num(t['sigma2v'], nan=1.0)
        for it in range(n_iter):
            for which, mv, pk, uu in (('pi', 'pi_move', 'pi_pick', 'u_pi'), ('tau', 'tau_move', 'tau_pick', 'u_tau')):
                if (it, which, 'move') in rec.log:
                    r[mv][u, it] = rec.log[(it, which, 'move')]
                    picks = rec.log.get((it, which, 'pick'), [])
                    r[pk][u, it, :len(picks)] = picks
                    r[uu][u, it] = rec.log.get((it, which, 'u'), 0.5)
            for slot, key in ((0, 'mu_pi'), (3, 'mu_tau')):
                if (it, 'mu', slot) in rec.log:
                    v = rec.log[(it, 'mu', slot)]
                    r[key][u, it, :len(v)] = v
    return r
