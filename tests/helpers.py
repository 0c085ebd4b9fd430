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
