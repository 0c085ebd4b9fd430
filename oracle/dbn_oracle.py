"""CPU oracle for the RJ-MCMC proposal-scoring hot path of dyban (charx7/DynamicBayesianNetworks).

TEST INFRASTRUCTURE ONLY.  Nothing under `dynamicbayesiannetworks_b200/` imports this module; only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may use it,
and only as the checker / the timed CPU baseline.  The product path is the CUDA library.

What it is: a NumPy FP64 restatement of the reference's algorithm for the path, function by function, each
citing the reference file:line it follows (paths relative to /root/reference/src/dyban/).  Two forms of the
scoring arithmetic are provided:

  * `log_ml_literal`  — the reference's own T_h x T_h formulation (C_h = I + w X_h X_h^T, det and inverse
                        per segment) with only the overflow-prone scalars moved to the log domain
                        (lgamma for log(gamma()), -(T/2) ln(pi) for log(pi**(-T/2)), slogdet for
                        log(det**0.5)).  O(T_h^3); the literal CPU cost of the reference.
  * `log_ml`          — the same quantity on k x k sufficient statistics G_h = X_h^T X_h, g_h = X_h^T y_h,
                        s_h = y_h^T y_h (matrix-determinant lemma + Woodbury).  The
                        statistics are accumulated directly from the data rows of each segment (NOT from
                        prefix sums), so it is an independent check of the GPU's prefix-difference layout.

Parity pinning: PINNED.  `tests/test_oracle_golden.py` checks every function here against outputs of the
unmodified reference run in the build container (`tests/golden/make_golden.py`, fixtures `tests/golden/*.npz`):
log-ML / beta-tilde / mu-sampler / prior known-answer vectors on yeast (N=33), mTOR (N=119) and synthetic
N=199 / N=340 series, and full replayed chains (every Gibbs parameter, every log-ML, every accept flag, pi and
tau after every iteration) for h-dbn, nh-dbn, seq-dbn and glob-dbn.  The reference's own unit tests pin
nothing for this path, and its log-ML overflows for T >= 344 (marginalLikelihood.py:116),
so beyond that size this oracle stands on the algebraic identity plus the `log_ml` == `log_ml_literal` check.
"""
import math

import numpy as np

# ------------------------------------------------------------------------------------------------
# method table: the per-method constants the reference hard-codes
# ------------------------------------------------------------------------------------------------
H_DBN, NH_DBN, SEQ_DBN, GLOB_DBN, FP_GLOB_DBN, FIXED_NH_DBN, VV_GLOB_DBN, FP_NH_DBN, FP_H_DBN, FP_VV_DBN, FP_SEQ_DBN = range(11)
METHOD_IDS = {
    'fp_seq_coup_nh_dbn': FP_SEQ_DBN,       # network.py:221-233, fpSeqCoupBpwlinReg.py (the seq loop without the parent-set move)
    'fp_var_glob_coup_nh_dbn': FP_VV_DBN,   # network.py:270-281, fpvvGlobCoup.py
    'fp_varying_nh_dbn': FP_NH_DBN,    # network.py:165-177, fullParentsBpwLinReg.py
    'fp_h_dbn': FP_H_DBN,              # network.py:200-209, fpBayesianLinearRegression.py
    'var_glob_coup_nh_dbn': VV_GLOB_DBN, 'var-glob-dbn': VV_GLOB_DBN,   # network.py:258-269, vvglobCoup.py
    'fp_glob_coup_nh_dbn': FP_GLOB_DBN, 'fp-glob-dbn': FP_GLOB_DBN,
    'h_dbn': H_DBN, 'h-dbn': H_DBN,
    'varying_nh_dbn': NH_DBN, 'nh-dbn': NH_DBN,
    'fixed_nh_dbn': FIXED_NH_DBN,      # network.py:178-189: nh regressor, predefined changepoints, no changepoint move
    'seq_coup_nh_dbn': SEQ_DBN, 'seq-dbn': SEQ_DBN,
    'glob_coup_nh_dbn': GLOB_DBN, 'glob-dbn': GLOB_DBN,
}


def method_constants(method, N):
    """Constructor arguments and hyper-parameters per method.

    network.py:157-163 (nh: num_samples=N), :193-197 (h: N+1), :213-219 (seq: N-1), :237-243 (glob: N);
    sigma^2 ~ IG(0.005,0.005) for nh/seq (bayesianPwLinearRegression.py:69-70,
    seqCoupledBayesianPwLinReg.py:44-45) and IG(0.01,0.01) for h/glob (bayesianLinearRegression.py:110-111,
    globCoupBayesianPwLinReg.py:43-44); lambda^2, delta^2 ~ IG(2,0.2) everywhere.
    Returns dict(T_ml, T_sigma, num_samples, a_sig, b_sig, a_lam, b_lam, a_del, b_del).
    """
    m = METHOD_IDS[method] if isinstance(method, str) else method
    c = dict(a_lam=2.0, b_lam=0.2, a_del=2.0, b_del=0.2)
    if m == H_DBN:
        # fit(): T = self.num_samples = N+1 (bayesianLinearRegression.py:96); ML gets num_samples-1 = N (:134-135)
        c.update(T_ml=N, T_sigma=N + 1, num_samples=N, a_sig=0.01, b_sig=0.01)
    elif m in (NH_DBN, FP_NH_DBN):
        # fp: network.py:168-174 (num_samples = N), fullParentsBpwLinReg.py:73-74 (IG(0.005, 0.005))
        c.update(T_ml=N, T_sigma=N, num_samples=N, a_sig=0.005, b_sig=0.005)
    elif m == FP_H_DBN:
        # network.py:203-207 (num_samples + 1), fpBayesianLinearRegression.py:70,84-85,90-91: shape a + (N+1)/2, no moves
        c.update(T_ml=N, T_sigma=N + 1, num_samples=N, a_sig=0.01, b_sig=0.01)
    elif m in (SEQ_DBN, FIXED_NH_DBN, FP_SEQ_DBN):   # fp-seq: network.py:224-230, fpSeqCoupBpwlinReg.py:50-51
        # seq: network.py:213-219; fixed nh: network.py:182-188 also passes num_samples - 1
        c.update(T_ml=N - 1, T_sigma=N - 1, num_samples=N - 1, a_sig=0.005, b_sig=0.005)
    elif m in (GLOB_DBN, FP_GLOB_DBN, VV_GLOB_DBN, FP_VV_DBN):   # fp-vv: network.py:273-279, fpvvGlobCoup.py:49-50
        # network.py:237-243, :249-255, :261-267 (num_samples = N); fpGlobCoupBpwLinReg.py:49-50, vvglobCoup.py:42-43
        # (IG(0.01, 0.01)); the segment-specific-variance model uses T_h per segment instead of T (log_ml_vv)
        c.update(T_ml=N, T_sigma=N, num_samples=N, a_sig=0.01, b_sig=0.01)
    else:
        raise ValueError(method)
    return c


# ------------------------------------------------------------------------------------------------
# a1 — data slicing (network.py:54-141, utils.py:128-269)
# ------------------------------------------------------------------------------------------------
def lagged_matrices(data_list, lag=1):
    """Feature / response matrices shared by all nodes.

    network.py:105-118: features of a segment are rows [:len-1]; :126-130: responses are rows [1:len];
    segments are concatenated.  Returns (X[N, n * lag], Y[N, n]); node i regresses Y[:, i] on the columns of X that
    are not copies of gene i.  lag > 1 (network.py:92-122): block l of X holds the series shifted by l more rows,
    zero-padded at the start of every segment; column l * n + g is the lag-(l + 1) copy of gene g.
    """
    n = np.asarray(data_list[0]).shape[1]
    blocks = []
    for l in range(int(lag)):
        rows = []
        for s in data_list:
            s = np.asarray(s, dtype=np.float64)
            rows.append(np.concatenate([np.zeros((l, n)), s[:s.shape[0] - (l + 1), :]], axis=0))
        blocks.append(np.concatenate(rows, axis=0))
    X = np.concatenate(blocks, axis=1)
    Y = np.concatenate([np.asarray(s, dtype=np.float64)[1:, :] for s in data_list], axis=0)
    return np.ascontiguousarray(X), np.ascontiguousarray(Y)


def label_to_column(label, node, n):
    """Reference parent label -> column of X for node `node` (network.py:92-104): the lag-1 block is labelled by gene id,
    every further block by a running counter over the node's features, so label n + j (j-th feature of the node) is the
    lag-2 copy of gene j (+ 1 if j >= node)."""
    label = int(label)
    if label < n:
        return label
    blk, j = divmod(label - n, n - 1)
    return (blk + 1) * n + (j if j < node else j + 1)


def column_to_label(col, node, n):
    col = int(col)
    if col < n:
        return col
    blk, g = divmod(col, n)
    return n + (blk - 1) * (n - 1) + (g if g < node else g - 1)


def segment_bounds(tau, N):
    """0-based [a, b) row ranges of the segments defined by the changepoint list `tau`.

    utils.py:162-224 / :226-269: segment h covers rows [c_{h-1}-1, c_h-1) with c_{-1}=1; the last entry of tau is
    the pseudo changepoint (N+2), reduced by one for X (utils.py:179) and clipped by slicing for y (:262), i.e. the
    last segment ends at N.
    """
    bounds = []
    lo = 1
    for h, c in enumerate(tau):
        a = lo - 1
        b = N if h == len(tau) - 1 else c - 1
        bounds.append((a, b))
        lo = c
    return bounds


def design_columns(pi):
    """Design-matrix column order = intercept, then parents sorted by label (utils.py:202-204)."""
    return sorted(int(p) for p in pi)


def segment_design(X, y, pi, tau):
    """Per-segment design matrices [1, X_pi] and responses, as the reference's constructNdArray builds them."""
    cols = design_columns(pi)
    N = X.shape[0]
    Xs, ys = [], []
    for a, b in segment_bounds(tau, N):
        D = np.ones((b - a, 1 + len(cols)))
        if cols:
            D[:, 1:] = X[a:b, cols]
        Xs.append(D)
        ys.append(y[a:b])
    return Xs, ys


def segment_stats(X, y, pi, tau):
    """(G_h, g_h, s_h, T_h) per segment, accumulated directly from the rows."""
    Xs, ys = segment_design(X, y, pi, tau)
    return [(D.T @ D, D.T @ yy, float(yy @ yy), D.shape[0]) for D, yy in zip(Xs, ys)]


# ------------------------------------------------------------------------------------------------
# a2 / a3 — marginal likelihood
# ------------------------------------------------------------------------------------------------
def _omega(h, lam, dlt, seq):
    # marginalLikelihood.py:104-107: delta^2 replaces lambda^2 for idx > 0 when method == 'seq-coup'
    return dlt if (seq and h > 0) else lam


def log_ml_literal(Xs, ys, mus, alpha, beta, lam, T, seq=False, dlt=None):
    """marginalLikelihood.py:91-150 in its own T_h x T_h form; `mus` is a list (one prior mean per segment)."""
    acc_logdet = 0.0
    acc_q = 0.0
    for h, (D, yy) in enumerate(zip(Xs, ys)):
        w = _omega(h, lam, dlt, seq)
        C = np.eye(D.shape[0]) + w * (D @ D.T)                      # :104-107
        sign, ld = np.linalg.slogdet(C)                             # :110-112 (log(det ** 0.5))
        acc_logdet += 0.5 * ld
        r = yy.reshape(-1, 1) - D @ np.asarray(mus[h], dtype=float).reshape(-1, 1)   # :133
        acc_q += float((r.T @ np.linalg.solve(C, r)).item())               # :137-144
    el1 = math.lgamma(T / 2 + alpha) - math.lgamma(alpha)           # :116
    el2 = -(T / 2) * math.log(math.pi) + alpha * math.log(2 * beta) - acc_logdet   # :118
    el3 = -(T / 2 + alpha) * math.log(2 * beta + acc_q)             # :146
    return el1 + el2 + el3


def seg_quadratic(G, g, s, mu, w):
    """Returns (ln det(I + w X X^T), r^T C^-1 r) for one segment from k x k statistics.

    ln det C = ln det(I_k + w G) (matrix determinant lemma), r^T C^-1 r = rho - w v^T A^-1 v with
    v = g - G mu, rho = s - 2 mu^T g + mu^T G mu, A = I_k + w G (Woodbury) — SURVEY.md Appendix A.
    """
    k = G.shape[0]
    A = np.eye(k) + w * G
    L = np.linalg.cholesky(A)
    logdet = 2.0 * float(np.sum(np.log(np.diag(L))))
    mu = np.asarray(mu, dtype=float).ravel()
    v = g - G @ mu
    rho = s - 2.0 * float(mu @ g) + float(mu @ G @ mu)
    z = np.linalg.solve(L, v)
    return logdet, rho - w * float(z @ z)


def log_ml(stats, mus, alpha, beta, lam, T, seq=False, dlt=None):
    """marginalLikelihood.py:91-150 on sufficient statistics (the form the GPU kernels evaluate)."""
    acc_logdet = 0.0
    acc_q = 0.0
    for h, (G, g, s, _) in enumerate(stats):
        ld, q = seg_quadratic(G, g, s, mus[h], _omega(h, lam, dlt, seq))
        acc_logdet += 0.5 * ld
        acc_q += q
    return (math.lgamma(T / 2 + alpha) - math.lgamma(alpha) - (T / 2) * math.log(math.pi)
            + alpha * math.log(2 * beta) - acc_logdet - (T / 2 + alpha) * math.log(2 * beta + acc_q))


def log_ml_h(stats, alpha, beta, lam, T):
    """log of calculateMarginalLikelihood (marginalLikelihood.py:152-166): one segment, mu = 0."""
    k = stats[0][0].shape[0]
    return log_ml(stats[:1], [np.zeros(k)], alpha, beta, lam, T)


def log_ml_vv(stats, mu, alpha, beta, lam):
    """vvLogMargLikelihood (marginalLikelihood.py:5-43): segments with their own variance, i.e. the sum over segments
    of single-segment log marginal likelihoods, each with its own length T_h in the Gamma / pi / exponent terms
    (:29-30, :39), all with lambda^2 and the same prior mean mu (:32)."""
    acc = 0.0
    for G, g, s, T_h in stats:
        ld, q = seg_quadratic(G, g, s, mu, lam)
        acc += (math.lgamma(T_h / 2 + alpha) - math.lgamma(alpha) - (T_h / 2) * math.log(math.pi)
                + alpha * math.log(2 * beta) - 0.5 * ld - (T_h / 2 + alpha) * math.log(2 * beta + q))
    return acc


# ------------------------------------------------------------------------------------------------
# a7 — priors
# ------------------------------------------------------------------------------------------------
CP_PRIOR_P = 0.125


def log_cp_prior(tau):
    """log of calculateChangePointsSetPrior (priors.py:4-46); -inf if more than 9 entries (:23-24)."""
    cc = [1] + list(tau)
    if len(cc) > 10:
        return -math.inf
    p = CP_PRIOR_P
    out = ((cc[-1] - 1) - cc[-2]) * math.log(1 - p)                 # :28-31
    for idx in range(len(tau) - 1):                                 # :35-42
        out += math.log(p) + (cc[idx + 1] - cc[idx] - 1) * math.log(1 - p)
    return out


def pi_prior(pi, F, fan_in):
    """calculateFeatureSetPriorProb (priors.py:48-68): uniform over sets of size <= fan_in, else 0."""
    if len(pi) > fan_in:
        return 0.0
    return 1.0 / sum(math.comb(F, i) for i in range(fan_in + 1))


# ------------------------------------------------------------------------------------------------
# a8 — sequentially coupled prior means
# ------------------------------------------------------------------------------------------------
def beta_tilde(stats, lam, dlt):
    """betaTildeSampler (samplers.py:4-54) on sufficient statistics.

    bt[0] = mu (zeros, :47-48); bt[1] = (I/lam + G_1)^-1 g_1 from segment 1's OWN data (:36-39);
    bt[h>=2] = (I/dlt + G_{h-1})^-1 (bt[h-2]/dlt + g_{h-1}) (:40-46, note the idx-2).
    """
    k = stats[0][0].shape[0]
    bt = []
    for h in range(len(stats)):
        if h == 0:
            bt.append(np.zeros(k))
        elif h == 1:
            G, g = stats[1][0], stats[1][1]
            bt.append(np.linalg.solve(np.eye(k) / lam + G, g))
        else:
            G, g = stats[h - 1][0], stats[h - 1][1]
            bt.append(np.linalg.solve(np.eye(k) / dlt + G, bt[h - 2] / dlt + g))
    return bt


# ------------------------------------------------------------------------------------------------
# a9 — global-coupling mean sampler
# ------------------------------------------------------------------------------------------------
def mu_posterior(stats, mu_in, sigma2, lam):
    """Mean and covariance of the MVN that muSampler draws from (samplers.py:307-351).

    X_h^T (s2 C_h)^-1 X_h = A_h^-1 G_h / s2 and X_h^T (s2 C_h)^-1 y_h = A_h^-1 g_h / s2 with A_h = I + lam G_h;
    Sigma = (I + sum_h ...)^-1 (:324); mean = Sigma (sum_h ... + mu_in) (:336-337: the INPUT mu is added).
    Returns (mean, Sigma, precision).
    """
    k = stats[0][0].shape[0]
    Q = np.eye(k)
    # vvMuSampler (samplers.py:353-394): segment-specific sigma^2_h (:362) and NO mu_in term in the mean (:379)
    per_seg = isinstance(sigma2, (list, tuple, np.ndarray))
    m = np.zeros(k) if mu_in is None else np.asarray(mu_in, dtype=float).ravel().copy()
    for h, (G, g, _, _) in enumerate(stats):
        A = np.eye(k) + lam * G
        s2 = sigma2[h] if per_seg else sigma2
        Q = Q + np.linalg.solve(A, G) / s2
        m = m + np.linalg.solve(A, g) / s2
    Sigma = np.linalg.inv(Q)
    return Sigma @ m, Sigma, Q


def mvn_logpdf_prec(x, mean, Q):
    """log N(x; mean, Q^-1) (what scipy.stats.multivariate_normal.pdf returns, logged; moves.py:106-116)."""
    x = np.asarray(x, dtype=float).ravel()
    d = x - np.asarray(mean, dtype=float).ravel()
    L = np.linalg.cholesky(Q)
    logdetQ = 2.0 * float(np.sum(np.log(np.diag(L))))
    return -0.5 * (x.size * math.log(2 * math.pi) - logdetQ + float(d @ Q @ d))


def std_normal_logpdf(x):
    """log N(x; 0, I) — the mu prior in moves.py:102-109, :501-508."""
    x = np.asarray(x, dtype=float).ravel()
    return -0.5 * (x.size * math.log(2 * math.pi) + float(x @ x))


# ------------------------------------------------------------------------------------------------
# a10-a12 — Gibbs full conditionals (parameters only; the draw itself comes from the draw source)
# ------------------------------------------------------------------------------------------------
def sigma2_params(stats, mus, lam, a_sig, b_sig, T, seq=False, dlt=None):
    """(shape, rate) of the Gamma whose reciprocal is sigma^2.

    samplers.py:103-127 (nh/glob), :56-88 (seq: beta-tilde means, delta^2 for h>0), :129-143 (h):
    shape = a + T/2, rate = b + 0.5 sum_h r_h^T C_h^-1 r_h.
    """
    q = 0.0
    for h, (G, g, s, _) in enumerate(stats):
        q += seg_quadratic(G, g, s, mus[h], _omega(h, lam, dlt, seq))[1]
    return a_sig + T / 2, b_sig + 0.5 * q


def beta_params(stats, mus, lam, sigma2, seq=False, dlt=None):
    """Per-segment (mean, cov) passed to np.random.multivariate_normal (samplers.py:207-218).

    mean = (I/w + G)^-1 (mu/w + g), cov = sigma2 (I/w + G)^-1, with w = lambda^2, or delta^2 for h > 0 in the
    sequentially coupled model (samplers.py:158-163) whose prior means are the beta-tildes (:156).
    """
    out = []
    for h, (G, g, _, _) in enumerate(stats):
        w = _omega(h, lam, dlt, seq)
        k = G.shape[0]
        P = np.eye(k) / w + G
        Pinv = np.linalg.inv(P)
        s2 = sigma2[h] if isinstance(sigma2, (list, tuple, np.ndarray)) else sigma2   # vvBetaSamplerWithChangepoints (samplers.py:169-187)
        out.append((Pinv @ (np.asarray(mus[h], dtype=float).ravel() / w + g), s2 * Pinv))
    return out


def lambda2_params(method, betas, mu, sigma2, a_lam, b_lam):
    """(shape, rate) for lambda^2.

    h:       samplers.py:296-305  shape a + k/2,        rate b + |beta - mu|^2 / (2 sigma2)
    nh/glob: samplers.py:265-294  shape a + S k / 2,    rate b + 0.5 sum_h |beta_h - mu|^2 / sigma2
    seq:     samplers.py:249-263  shape a + (k + 1)/2,  rate b + beta_0^T beta_0 / (2 sigma2)
    """
    k = len(betas[0])
    mu = np.asarray(mu, dtype=float).ravel()
    if method == H_DBN:
        d = betas[0] - mu
        return a_lam + k / 2, b_lam + 0.5 * float(d @ d) / sigma2
    if method == SEQ_DBN:
        return a_lam + (k + 1) / 2, b_lam + 0.5 * float(betas[0] @ betas[0]) / sigma2
    acc = 0.0
    for h, b in enumerate(betas):
        d = b - mu
        # segment-specific variances when sigma2 is a list (samplers.py:275-282)
        acc += float(d @ d) / (sigma2[h] if isinstance(sigma2, (list, tuple, np.ndarray)) else sigma2)
    return a_lam + len(betas) * k / 2, b_lam + 0.5 * acc


def delta2_params(betas, bt, sigma2, a_del, b_del):
    """(shape, rate) for delta^2 (samplers.py:220-247): beta_h is centred on bt[h-1] (:228) for h >= 1."""
    k = len(betas[0])
    acc = 0.0
    for h in range(1, len(betas)):
        d = betas[h] - bt[h - 1]
        acc += float(d @ d)
    return a_del + (len(betas) - 1) * k / 2, b_del + 0.5 * acc / sigma2


# ------------------------------------------------------------------------------------------------
# a4 / a5 — proposals (pure index manipulation; `pick(n)` returns an index in [0, n))
# ------------------------------------------------------------------------------------------------
ADD, DELETE, EXCHANGE = 0, 1, 2
BIRTH, DEATH, REALLOC = 0, 1, 2


def propose_pi(pi, move, features, fan_in, pick):
    """addMove / deleteMove / exchangeMove (utils.py:271-323). Returns (pi_star | None, hr).

    `pi` keeps the reference's order: add/exchange append (utils.py:305,323), delete returns sorted (:290).
    `features` = the candidate parent labels (all genes but the node), ascending (np.setdiff1d sorts).
    hr per moves.py:610-618.
    """
    pi = [int(p) for p in pi]
    F = len(features)
    if move == ADD:
        if len(pi) > fan_in - 1:                                    # utils.py:294-295
            return None, 0.0
        cand = [f for f in features if f not in pi]                 # :301
        if not cand:
            return None, 0.0
        star = pi + [cand[pick(len(cand))]]                         # :303-305
        return star, (F - len(pi)) / len(star)
    if move == DELETE:
        if len(pi) < 1:                                             # :285-286
            return None, 0.0
        el = pi[pick(len(pi))]                                      # :289
        star = sorted(p for p in pi if p != el)                     # :290
        return star, len(pi) / (F - len(star))
    if len(pi) < 1:                                                 # :308-309
        return None, 0.0
    el = pi[pick(len(pi))]                                          # :314
    cand = [f for f in features if f not in pi]                     # :317
    rest = sorted(p for p in pi if p != el)                         # :318
    if not cand:
        return None, 0.0
    return rest + [cand[pick(len(cand))]], 1.0                      # :320-323


def propose_tau(tau, move, num_samples, pick, glob=False):
    """cpBirthMove / cpDeathMove / cpRellocationMove (changepointMoves.py) + validity and hr.

    nh/seq: moves.py:154-175 (birth invalid when len == 10, hr = (n - |tau|)/(|tau*| - 1);
            death hr = (|tau| - 1)/(n - |tau*|)).
    glob:   moves.py:40-64 (birth invalid when len > 9, hr = (n - 1 - |tau|)/|tau*|; death hr = |tau|/(n - 1 - |tau*|)).
    Returns (tau_star | None, hr).
    """
    tau = [int(c) for c in tau]
    n = num_samples
    if move == BIRTH:
        cand = [c for c in range(2, n + 1) if c not in tau]        # changepointMoves.py:112-117
        if not cand:
            return None, 0.0
        star = sorted(tau + [cand[pick(len(cand))]])                # :118-122
        if len(star) > 9:
            return None, 0.0
        hr = (n - 1 - len(tau)) / len(star) if glob else (n - len(tau)) / (len(star) - 1)
        return star, hr
    if move == DEATH:
        if len(tau) < 2:                                            # :81-82
            return None, 0.0
        star = list(tau)
        star.pop(pick(len(tau) - 1))                                # :86-89
        hr = len(tau) / (n - 1 - len(star)) if glob else (len(tau) - 1) / (n - len(star))
        return star, hr
    if len(tau) < 2:                                                # :21-22
        return None, 0.0
    idx = pick(len(tau) - 1)                                        # :28-30
    last = tau[-1] - 1                                              # :26
    left = 1 if idx == 0 else tau[idx - 1]                          # :33-37
    right = last if idx + 1 == len(tau) - 1 else tau[idx + 1]       # :39-43
    cand = [c for c in range(left + 1, right) if c != tau[idx]]     # :47-49
    if not cand:                                                    # :50-51
        return None, 0.0
    new = cand[pick(len(cand))]                                     # :55
    star = sorted(tau[:idx] + tau[idx + 1:-1] + [new]) + [tau[-1]]  # :56-58
    return star, 1.0


# ------------------------------------------------------------------------------------------------
# draw sources
# ------------------------------------------------------------------------------------------------
class ReplayDraws:
    """Feeds a recorded reference run (tests/golden/chain_*.npz) back, one node at a time."""

    def __init__(self, golden, node):
        self.g, self.i = golden, node

    def inv_gamma(self, it, which, shape, rate):
        return float(self.g[{'sigma': 'sigma2', 'lambda': 'lambda2', 'delta': 'delta2'}[which]][self.i, it])

    def inv_gamma_seg(self, it, h, shape, rate):
        return float(self.g['sigma2v'][self.i, it, h])

    def beta(self, it, h, mean, cov):
        k = len(mean)
        return np.array(self.g['beta'][self.i, it, h, :k])

    def move(self, it, which):
        return int(self.g[which + '_move'][self.i, it])

    def picker(self, it, which):
        picks = list(self.g[which + '_pick'][self.i, it])
        return lambda n, _p=picks: int(_p.pop(0))

    def mu_draw(self, it, slot, mean, Sigma):
        k = len(mean)
        return np.array(self.g['mus_val'][self.i, it, slot, :k])

    def uniform(self, it, which):
        return float(self.g['u_' + which][self.i, it])


class RngDraws:
    """Free-running draws from a numpy Generator (statistical parity only)."""

    def __init__(self, rng):
        self.rng = rng

    def inv_gamma(self, it, which, shape, rate):
        return 1.0 / self.rng.gamma(shape, 1.0 / rate)

    def inv_gamma_seg(self, it, h, shape, rate):
        return 1.0 / self.rng.gamma(shape, 1.0 / rate)

    def beta(self, it, h, mean, cov):
        L = np.linalg.cholesky(cov)
        return mean + L @ self.rng.standard_normal(len(mean))

    def move(self, it, which):
        return int(self.rng.integers(0, 3))

    def picker(self, it, which):
        return lambda n: int(self.rng.integers(0, n))

    def mu_draw(self, it, slot, mean, Sigma):
        L = np.linalg.cholesky(Sigma)
        return mean + L @ self.rng.standard_normal(len(mean))

    def uniform(self, it, which):
        return float(self.rng.uniform())


# ------------------------------------------------------------------------------------------------
# a13 — the per-node chain (regressor loops)
# ------------------------------------------------------------------------------------------------
def _accept_plain(log_ratio, u):
    """moves.py:225-232, :620-637: u < min(1, exp(delta))."""
    return u < min(1.0, math.exp(min(log_ratio, 700.0)))


def _accept_glob(log_ratio, u):
    """moves.py:112-120, :512-520: u < exp(min(1, delta))."""
    return u < math.exp(min(1.0, log_ratio))


def run_chain(method, X, Y, node, n_iter, draws, fan_in=3, with_tau=True, init_tau=None):
    # X may hold several lag blocks (lagged_matrices(..., lag)): the node's own copies in every block are excluded
    """One (node, chain) unit for `n_iter` iterations; returns a trace dict shaped like the golden files.

    Order of updates per iteration follows bayesianLinearRegression.py:114-147 (h),
    bayesianPwLinearRegression.py:89-132 (nh), seqCoupledBayesianPwLinReg.py:68-121 (seq),
    globCoupBayesianPwLinReg.py:64-116 (glob): sigma^2, beta, lambda^2 (, delta^2), pi move, tau move.
    Full-parents globally coupled (fpGlobCoupBpwLinReg.py:16-119): pi is fixed to all other genes (:27-31), there is
    no parent-set move, the changepoint move is the globally coupled one (:88-98); trace arrays are n wide.
    """
    m = METHOD_IDS[method] if isinstance(method, str) else method
    N, n = X.shape
    y = Y[:, node]
    n_genes = Y.shape[1]
    features = [j for j in range(n) if j % n_genes != node]
    F = len(features)
    c = method_constants(m, N)
    fp = m in (FP_GLOB_DBN, FP_NH_DBN, FP_H_DBN, FP_VV_DBN, FP_SEQ_DBN)    # full parents: pi = every other gene, no parent-set move
    vv = m in (VV_GLOB_DBN, FP_VV_DBN)   # vvglobCoup.py:66-123: segment-specific sigma^2_h, mu kept in the changepoint move
    seq, glob = m in (SEQ_DBN, FP_SEQ_DBN), m in (GLOB_DBN, FP_GLOB_DBN, VV_GLOB_DBN, FP_VV_DBN)
    has_tau = with_tau and m not in (H_DBN, FIXED_NH_DBN, FP_H_DBN)   # bayesianPwLinearRegression.py:117: only 'varying_nh' moves tau
    KM, SM = (max(5, n) if fp else 5), 10

    pi, tau = (list(features) if fp else []), ([int(c) for c in init_tau] if init_tau is not None else [N + 2])
    lam = sig = dlt = 1.0
    mu = np.zeros(len(pi) + 1)                                      # glob: current global mean (k entries)
    tr = {k_: np.full(n_iter, np.nan) for k_ in (
        'sigma2', 'lambda2', 'delta2', 'sig_shape', 'sig_scale', 'lam_shape', 'lam_scale', 'del_shape', 'del_scale')}
    tr['beta_mean'] = np.full((n_iter, SM, KM), np.nan)
    tr['beta_cov'] = np.full((n_iter, SM, KM, KM), np.nan)
    tr['logml'] = np.full((n_iter, 4), np.nan)
    for k_ in ('pi_valid', 'tau_valid', 'pi_accept', 'tau_accept', 'npi', 'S'):
        tr[k_] = np.full(n_iter, -1, dtype=np.int32)
    tr['pi_after'] = np.full((n_iter, KM - 1), -1, dtype=np.int32)
    tr['tau_after'] = np.full((n_iter, SM), -1, dtype=np.int32)
    tr['mus_mean'] = np.full((n_iter, 4, KM), np.nan)
    tr['mus_cov'] = np.full((n_iter, 4, KM, KM), np.nan)
    tr['mus_logdens'] = np.full((n_iter, 4, 2), np.nan)
    tr['mu_after'] = np.full((n_iter, KM), np.nan)
    for k_ in ('sigma2v', 'sig_shape_v', 'sig_scale_v'):
        tr[k_] = np.full((n_iter, SM), np.nan)
    tr['n_evals'] = 0

    def prior_means(stats, lam_, dlt_, mu_):
        if seq:
            return beta_tilde(stats, lam_, dlt_)
        k = stats[0][0].shape[0]
        return [np.asarray(mu_, dtype=float).ravel() if glob else np.zeros(k)] * len(stats)

    def score(stats, lam_, dlt_, mu_):
        tr['n_evals'] += 1
        if vv:
            return log_ml_vv(stats, mu_, c['a_sig'], c['b_sig'], lam_)
        return log_ml(stats, prior_means(stats, lam_, dlt_, mu_), c['a_sig'], c['b_sig'], lam_, c['T_ml'], seq, dlt_)

    for it in range(n_iter):
        stats = segment_stats(X, y, pi, tau)
        S, k = len(stats), len(pi) + 1
        tr['S'][it] = S
        # ---- sigma^2
        mus = prior_means(stats, lam, dlt, mu)
        if vv:
            # segmentSigmaSampler (samplers.py:90-101): sigmaSqrSampler per segment with numSamples = T = T_h
            sig = []
            for h, (G, g_, s_, T_h) in enumerate(stats):
                shape, rate = c['a_sig'] + T_h / 2, c['b_sig'] + 0.5 * seg_quadratic(G, g_, s_, mu, lam)[1]
                sig.append(draws.inv_gamma_seg(it, h, shape, rate))
                tr['sig_shape_v'][it, h], tr['sig_scale_v'][it, h], tr['sigma2v'][it, h] = shape, 1.0 / rate, sig[-1]
            tr['sig_shape'][it], tr['sig_scale'][it], tr['sigma2'][it] = (
                tr['sig_shape_v'][it, 0], tr['sig_scale_v'][it, 0], sig[0])
        else:
            shape, rate = sigma2_params(stats, mus, lam, c['a_sig'], c['b_sig'], c['T_sigma'], seq, dlt)
            tr['sig_shape'][it], tr['sig_scale'][it] = shape, 1.0 / rate
            sig = draws.inv_gamma(it, 'sigma', shape, rate)
            tr['sigma2'][it] = sig
        # ---- beta
        betas = []
        for h, (mean, cov) in enumerate(beta_params(stats, mus, lam, sig, seq, dlt)):
            tr['beta_mean'][it, h, :k] = mean
            tr['beta_cov'][it, h, :k, :k] = cov
            betas.append(draws.beta(it, h, mean, cov))
        # ---- lambda^2 (and delta^2)
        tr.setdefault('beta', np.full((n_iter, SM, KM), np.nan))
        for h, b_ in enumerate(betas):
            tr['beta'][it, h, :k] = b_
        # (fp-h: S = 1, so the nh / glob shape a + S k / 2 is samplers.py:296-305's a + k / 2)
        shape, rate = lambda2_params(SEQ_DBN if seq else (GLOB_DBN if fp else m), betas, mu if glob else np.zeros(k), sig, c['a_lam'], c['b_lam'])
        tr['lam_shape'][it], tr['lam_scale'][it] = shape, 1.0 / rate
        lam_new = draws.inv_gamma(it, 'lambda', shape, rate)
        tr['lambda2'][it] = lam_new
        dlt_new = dlt
        if seq:
            bt = beta_tilde(stats, lam_new, dlt)                    # samplers.py:223 (lambda[it+1], delta[it])
            shape, rate = delta2_params(betas, bt, sig, c['a_del'], c['b_del'])
            tr['del_shape'][it], tr['del_scale'][it] = shape, 1.0 / rate
            dlt_new = draws.inv_gamma(it, 'delta', shape, rate)
            tr['delta2'][it] = dlt_new
        lam, dlt = lam_new, dlt_new

        mu_next = None
        if m == FP_VV_DBN:
            # fpvvGlobCoup.py:99-103: mu ~ vvMuSampler at (tau, sigma^2_h, lambda^2 new), always accepted, but the
            # changepoint move below still conditions on the OLD mu (muVector[it], :106-108)
            mean_s, Sig_s, _ = mu_posterior(stats, None, sig, lam)
            mu_next = draws.mu_draw(it, 0, mean_s, Sig_s)
            tr['mus_mean'][it, 0, :k] = mean_s
            tr['mus_cov'][it, 0, :k, :k] = Sig_s
        # ---- pi move
        star = None
        if not fp:
            move = draws.move(it, 'pi')
            star, hr = propose_pi(pi, move, features, fan_in, draws.picker(it, 'pi'))
            tr['pi_valid'][it] = int(star is not None)
            tr['pi_accept'][it] = 0
        if star is not None:
            stats_star = segment_stats(X, y, star, tau)
            lp = math.log(pi_prior(star, F, fan_in)) - math.log(pi_prior(pi, F, fan_in))
            if glob:
                # moves.py:476-491: mu* ~ q(.|pi*, mu_in = 0); reverse density with mu_in = current mu
                # (vv: moves.py:361-375, vvMuSampler has no mu_in term in its mean, samplers.py:379)
                mean_s, Sig_s, Q_s = mu_posterior(stats_star, None if vv else np.zeros(len(star) + 1), sig, lam)
                mu_star = draws.mu_draw(it, 0, mean_s, Sig_s)
                ml_star = score(stats_star, lam, dlt, mu_star)
                ml_cur = score(stats, lam, dlt, mu)
                mean_c, Sig_c, Q_c = mu_posterior(stats, None if vv else mu, sig, lam)
                ld_star = mvn_logpdf_prec(mu_star, mean_s, Q_s)
                ld_cur = mvn_logpdf_prec(mu, mean_c, Q_c)
                tr['mus_mean'][it, 0, :len(mean_s)] = mean_s
                tr['mus_cov'][it, 0, :len(mean_s), :len(mean_s)] = Sig_s
                tr['mus_mean'][it, 1, :k] = mean_c
                tr['mus_cov'][it, 1, :k, :k] = Sig_c
                tr['mus_logdens'][it, 0, 0] = ld_star
                tr['mus_logdens'][it, 1, 1] = ld_cur
                ratio = (ml_star - ml_cur + lp + ld_cur - ld_star
                         + std_normal_logpdf(mu_star) - std_normal_logpdf(mu) + math.log(hr))
                acc = _accept_glob(ratio, draws.uniform(it, 'pi'))
            else:
                if m == H_DBN:
                    ml_star = log_ml_h(stats_star, c['a_sig'], c['b_sig'], lam, c['T_ml']); tr['n_evals'] += 1
                    ml_cur = log_ml_h(stats, c['a_sig'], c['b_sig'], lam, c['T_ml']); tr['n_evals'] += 1
                else:
                    ml_star = score(stats_star, lam, dlt, None)
                    ml_cur = score(stats, lam, dlt, None)
                acc = _accept_plain(ml_star - ml_cur + lp + math.log(hr), draws.uniform(it, 'pi'))
            tr['logml'][it, 0], tr['logml'][it, 1] = ml_star, ml_cur
            tr['pi_accept'][it] = int(acc)
            if acc:
                pi, stats = star, stats_star
                if glob:
                    mu = mu_star
        tr['npi'][it] = len(pi)
        tr['pi_after'][it, :len(pi)] = pi

        # ---- tau move
        if has_tau:
            move = draws.move(it, 'tau')
            star, hr = propose_tau(tau, move, c['num_samples'], draws.picker(it, 'tau'), glob)
            tr['tau_valid'][it] = int(star is not None)
            tr['tau_accept'][it] = 0
            if star is not None:
                stats_star = segment_stats(X, y, pi, star)
                lp = log_cp_prior(star) - log_cp_prior(tau)
                k = len(pi) + 1
                if vv:
                    # vvGlobCoupTauMove (moves.py:238-307): mu is kept, no proposal densities; glob-style hr and accept
                    ml_cur = score(stats, lam, dlt, mu)
                    ml_star = score(stats_star, lam, dlt, mu)
                    acc = _accept_glob(ml_star - ml_cur + lp + math.log(hr), draws.uniform(it, 'tau'))
                elif glob:
                    # moves.py:66-93: current density first (mu_in = mu), then mu* ~ q(.|tau*, mu_in = 0)
                    ml_cur = score(stats, lam, dlt, mu)
                    mean_c, Sig_c, Q_c = mu_posterior(stats, mu, sig, lam)
                    ld_cur = mvn_logpdf_prec(mu, mean_c, Q_c)
                    mean_s, Sig_s, Q_s = mu_posterior(stats_star, np.zeros(k), sig, lam)
                    mu_star = draws.mu_draw(it, 3, mean_s, Sig_s)
                    ld_star = mvn_logpdf_prec(mu_star, mean_s, Q_s)
                    ml_star = score(stats_star, lam, dlt, mu_star)
                    tr['mus_mean'][it, 2, :k] = mean_c
                    tr['mus_cov'][it, 2, :k, :k] = Sig_c
                    tr['mus_mean'][it, 3, :k] = mean_s
                    tr['mus_cov'][it, 3, :k, :k] = Sig_s
                    tr['mus_logdens'][it, 2, 1] = ld_cur
                    tr['mus_logdens'][it, 3, 0] = ld_star
                    ratio = (ml_star - ml_cur + lp + std_normal_logpdf(mu_star) - std_normal_logpdf(mu)
                             + ld_cur - ld_star + math.log(hr))
                    acc = _accept_glob(ratio, draws.uniform(it, 'tau'))
                else:
                    ml_cur = score(stats, lam, dlt, None)
                    ml_star = score(stats_star, lam, dlt, None)
                    acc = _accept_plain(ml_star - ml_cur + lp + math.log(hr), draws.uniform(it, 'tau'))
                tr['logml'][it, 2], tr['logml'][it, 3] = ml_cur, ml_star
                tr['tau_accept'][it] = int(acc)
                if acc:
                    tau = star
                    if glob and not vv:
                        mu = mu_star
        if mu_next is not None:
            mu = mu_next
        tr['tau_after'][it, :len(tau)] = tau
        if glob:
            tr['mu_after'][it, :len(mu)] = mu
    return tr


# ------------------------------------------------------------------------------------------------
# a14 — edge-frequency and changepoint-histogram reductions
# ------------------------------------------------------------------------------------------------
THIN = 10


def edge_kept(idx, burn_in):
    """pi_vector[idx] is kept iff idx >= burn_in and (idx - burn_in) % 10 != 0 (network.py:358-359)."""
    return idx >= burn_in and (idx - burn_in) % THIN != 0


def tau_kept(j, burn_in):
    """tau_vector[j] is kept iff j >= burn_in and (j - burn_in) % 10 == 0 (network.py:388-389)."""
    return j >= burn_in and (j - burn_in) % THIN == 0


def accumulate_counts(trace, node, n, N, burn_in, edge, nonempty, cp_hist, cp_kept):
    """Add one unit's samples to the count matrices that get all-reduced (scores.py:155-194,
    examples/plot_diagnostics.py:226-240).  pi_vector[0] is the initial empty list, pi_vector[it+1] = pi_after[it];
    tau_vector[it] = tau_after[it]."""
    n_iter = trace['npi'].shape[0]
    for it in range(n_iter):
        if edge_kept(it + 1, burn_in):
            k = int(trace['npi'][it])
            if k > 0:
                nonempty[node] += 1
                # lag > 1: a gene counts once when any of its copies is a parent (scores.py:172-178 for lag 2)
                for gene in sorted({int(p) % n for p in trace['pi_after'][it, :k]}):
                    edge[node, gene] += 1
        if tau_kept(it, burn_in) and trace['tau_valid'][it] >= 0:
            cp_kept[node] += 1
            for c_ in trace['tau_after'][it]:
                if 0 <= c_ <= N:
                    cp_hist[node, int(c_)] += 1


def credible_scores(mu_samples, features, n):
    """Edge-score row of a full-parents globally coupled node (network.py:303-347, scores.py:196-205, :223-253):
    for every parent, max(fraction of thinned mu samples >= 0, fraction <= 0); `mu_samples` is [n_kept, k] with
    column 0 the intercept and column 1 + j the j-th feature (sorted by label); the node's own column stays 0."""
    mu_samples = np.asarray(mu_samples, dtype=float)
    row = np.zeros(n)
    for j, f in enumerate(features):
        col = mu_samples[:, j + 1]
        row[f] = max(np.mean(col >= 0), np.mean(col <= 0))
    return row


def beta_kept(it, burn_in):
    """Full-parents methods scored from the beta chain (network.py:316-320): betas_vector[2:][burn_in:][x], x % 10 == 0,
    i.e. betas_vector[it + 1] (the draw of iteration it) is kept iff it - 1 >= burn_in and (it - 1 - burn_in) % 10 == 0;
    every segment's vector of a kept iteration is one posterior sample (beta_post_matrix, scores.py:139-153)."""
    return it - 1 >= burn_in and (it - 1 - burn_in) % THIN == 0


def mu_kept(idx, burn_in):
    """mu_vector[idx] (idx = 0 is the initial zero vector, idx = it + 1 the state after iteration it) is kept iff
    idx >= burn_in and (idx - burn_in) % 10 == 0 (network.py:309-311)."""
    return idx >= burn_in and (idx - burn_in) % THIN == 0


def edge_scores(edge, nonempty):
    """proposed_adj_matrix (row = child): count / number of kept non-empty parent sets (scores.py:184-192)."""
    with np.errstate(invalid='ignore', divide='ignore'):
        return edge / nonempty[:, None]
