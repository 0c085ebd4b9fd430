"""Host-side scoring helpers of the drop-in: the reference's `dyban.scores` functions that `Network.score_edges` calls on
a finished chain (scores.py:92-253) and `transform_beta_coef` (bayesianLinearRegression.py:40-67), restated on NumPy
arrays.  They run on the per-iteration histories a `keep_chain` run returns; multi-chain runs get the same numbers from
the device's count matrices (`parallel.edge_scores_from_counts` / `credible_scores_from_counts`).  Function names and
argument meaning follow the reference so that code written against `dyban.scores` keeps working; plotting / ROC helpers
are in `evaluation.py`.
"""
import numpy as np


def transform_beta_coef(beta, pi, dims):
    """bayesianLinearRegression.py:40-67: pad every segment's coefficient vector to `dims + 1` entries, a zero where the
    feature is not a parent.  Quirks kept: entry d >= 1 is matched against the parent LABEL d, so
    gene 0 is never matched, and the coefficients are consumed in order of the matches -- with gene 0 in the parent set
    the first matched entry receives gene 0's coefficient."""
    if not isinstance(beta, list):
        beta = [beta]
    srtd = sorted(int(p) for p in pi)
    out = []
    for b in beta:
        b = np.asarray(b, dtype=float).ravel()
        v = np.zeros(dims + 1)
        v[0] = b[0]
        idx = 0
        for d in range(1, dims + 1):
            if d in srtd:
                idx += 1
                v[d] = b[idx]
        out.append(v)
    return out


def fraction_score(posterior_sample):
    """scores.py:92-104: 0.5 - fraction of samples <= 0."""
    s = np.asarray(posterior_sample, dtype=float)
    return 0.5 - np.count_nonzero(s <= 0) / s.size


def credible_score(posterior_sample):
    """scores.py:196-205: max(fraction >= 0, fraction <= 0)."""
    s = np.asarray(posterior_sample, dtype=float)
    return max(np.count_nonzero(s >= 0) / s.size, np.count_nonzero(s <= 0) / s.size)


def get_betas_over_time(time_pts, thinned_changepoints, thinned_chain, dims):
    """scores.py:125-137: for every time point t (0-based), the [n_kept, dims] matrix whose row i is the coefficient
    vector of kept sample i in the segment that contains t (the first changepoint c with t + 1 < c)."""
    n_kept = min(len(thinned_changepoints), len(thinned_chain))
    per_sample = np.zeros((n_kept, time_pts, dims))
    t1 = np.arange(1, time_pts + 1)
    for i in range(n_kept):
        cps = np.asarray(thinned_changepoints[i], dtype=np.int64)
        seg = np.searchsorted(cps, t1, side='right')           # first j with t + 1 < cps[j]
        vecs = np.stack([np.asarray(v, dtype=float).reshape(dims) for v in thinned_chain[i]])
        per_sample[i] = vecs[np.minimum(seg, len(vecs) - 1)]
    return [per_sample[:, t, :].copy() for t in range(time_pts)]


def get_scores_over_time(betas_over_time, currFeatures, dims):
    """scores.py:106-123: [time_pts, dims - 1] matrix of fraction scores, column j = feature j of the design matrix."""
    out = np.zeros((len(betas_over_time), dims - 1))
    for t, m in enumerate(betas_over_time):
        for j in range(len(currFeatures)):
            out[t, j] = fraction_score(m[:, j + 1])
    return out


def beta_post_matrix(thinned_chain):
    """scores.py:139-153: every segment's vector of every kept iteration as one row."""
    rows = [np.asarray(v, dtype=float).ravel() for row in thinned_chain for v in row]
    return np.stack(rows) if rows else np.zeros((0, 0))


def score_beta_matrix(betas_matrix, currFeatures, currResponse):
    """scores.py:223-253 without the histogram plot: credible score of every feature's column."""
    dims = betas_matrix.shape[1]
    edge_scores = [0 for _ in range(dims)]
    for j, feat in enumerate(currFeatures):
        edge_scores[feat] = credible_score(betas_matrix[:, j + 1])
    return edge_scores


def calculateFeatureScores(selectedFeaturesVector, totalDims, currentFeatures, currentResponse):
    """scores.py:155-194: frequency of every feature in the kept parent sets, over the kept NON-EMPTY sets; with lag > 1
    a gene counts when either its lag-1 or its highest-lag copy is in the set (:172-178)."""
    adj = [0 for _ in range(len(currentFeatures) + 1)]
    denom = len([x for x in selectedFeaturesVector if np.size(x) != 0])
    sub = currentFeatures[:totalDims - 1]
    for idx, feat in enumerate(sub):
        if len(currentFeatures) > totalDims:
            lagged = currentFeatures[idx + (len(currentFeatures) - totalDims) + 1]
            freq = sum(1 for pi in selectedFeaturesVector if (feat in pi) or (lagged in pi))
        else:
            freq = sum(1 for pi in selectedFeaturesVector if feat in pi)
        adj[feat] = freq / denom
    return adj
