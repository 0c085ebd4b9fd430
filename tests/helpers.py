"""Shared helpers for the parity tests (loading golden fixtures, tolerances)."""
import glob
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')

# north_star: "log marginal likelihoods match the reference implementation within 1e-9 relative in FP64"
LOGML_RTOL = 1e-9

KAT_FILES = sorted(glob.glob(os.path.join(GOLDEN, 'kat_logml_*.npz')))
CHAIN_FILES = sorted(glob.glob(os.path.join(GOLDEN, 'chain_*.npz')))


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
