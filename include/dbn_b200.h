/*
 * dbn_b200.h — C ABI of the B200-native RJ-MCMC proposal-scoring path for dyban-style dynamic Bayesian networks.
 *
 * The reference (charx7/DynamicBayesianNetworks, package `dyban`) is pure Python/NumPy and has no FFI of its own
 * (SURVEY.md section 8b).  These entry points are what a `ctypes` binding inside the reference would call in place
 * of its per-node regressor loops; each declaration cites the reference code it replaces (paths relative to
 * /root/reference/src/dyban/).  Plain pointers and sizes only: no C++ or torch types cross this boundary.
 *
 * Ownership: the caller owns every host buffer (C-contiguous, FP64 / int32 / int64).  The library owns all device
 * memory behind the opaque handle; no pointer returned by the library outlives dbn_destroy().
 * Errors: every function returns 0 on success or a negative dbn_status; dbn_last_error() gives the message.
 * No exception crosses the ABI.  Invalid MCMC proposals are data (flags in the trace), never errors.
 * Threading: one handle = one CUDA device + one stream.  Calls on one handle must be serialised by the caller;
 * distinct handles may be driven from distinct host threads (ctypes releases the GIL), which covers the
 * chain-level ThreadPoolExecutor pattern of examples/simulated_data.py:184-202.
 */
#ifndef DBN_B200_H
#define DBN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DBN_ABI_VERSION 1

/* padded widths used by every array layout in this header */
#define DBN_KMAX 5   /* design columns: intercept + up to 4 parents (reference fan-in is 3)      */
#define DBN_PMAX 4   /* parents per node                                                          */
#define DBN_SMAX 10  /* segments: the reference allows at most 9 changepoint entries (priors.py:23) */

typedef enum {
  DBN_OK = 0,
  DBN_ERR_INVALID = -1,  /* bad argument                            */
  DBN_ERR_CUDA = -2,     /* CUDA runtime error (message has detail) */
  DBN_ERR_STATE = -3,    /* call order (e.g. run before build_stats) */
  DBN_ERR_NOMEM = -4
} dbn_status;

/* model switch of network.py:154-281 (first-pass methods) */
typedef enum {
  DBN_H_DBN = 0,    /* 'h_dbn'            bayesianLinearRegression.py:70-155      */
  DBN_NH_DBN = 1,   /* 'varying_nh_dbn'   bayesianPwLinearRegression.py:41-141    */
  DBN_SEQ_DBN = 2,  /* 'seq_coup_nh_dbn'  seqCoupledBayesianPwLinReg.py:16-130    */
  DBN_GLOB_DBN = 3, /* 'glob_coup_nh_dbn' globCoupBayesianPwLinReg.py:16-124      */
  DBN_FP_GLOB_DBN = 4, /* 'fp_glob_coup_nh_dbn' fpGlobCoupBpwLinReg.py:16-119: every other gene is a parent (design
                         width k = n), no parent-set move, globally coupled changepoint move; run with dbn_run_fp */
  DBN_VV_GLOB_DBN = 5, /* 'var_glob_coup_nh_dbn' vvglobCoup.py:16-123: globally coupled with a segment-specific
                         variance sigma^2_h (segmentSigmaSampler samplers.py:90-101), vvLogMargLikelihood
                         (marginalLikelihood.py:5-43), vvMuSampler (samplers.py:353-394), vvGlobCoupPiMove /
                         vvGlobCoupTauMove (moves.py:238-415)                                                  */
  DBN_FP_NH_DBN = 6,   /* 'fp_varying_nh_dbn' fullParentsBpwLinReg.py:41-140: full parents, prior mean 0, the nh changepoint
                         move (moves.py:129-236); run with dbn_run_fp                                          */
  DBN_FP_H_DBN = 7,    /* 'fp_h_dbn' fpBayesianLinearRegression.py:40-125: full parents, one segment, Gibbs draws only;
                         run with dbn_run_fp                                                                   */
  DBN_FP_VV_DBN = 8,   /* 'fp_var_glob_coup_nh_dbn' fpvvGlobCoup.py:16-130: full parents, segment-specific variances, mu
                         redrawn every iteration (vvMuSampler), vvGlobCoupTauMove on the previous mu; run with dbn_run_fp */
  DBN_FP_SEQ_DBN = 9   /* 'fp_seq_coup_nh_dbn' fpSeqCoupBpwlinReg.py:16-130: full parents, sequentially coupled (beta-tilde
                         prior means, delta^2), nh changepoint move; run with dbn_run_fp.  Its delta^2 is injected /
                         traced through the first entries of the per-segment variance slots (shape, rate, draw)  */
} dbn_method;

typedef struct dbn_handle dbn_handle;

typedef struct {
  int32_t device;       /* CUDA device ordinal                                                            */
  int32_t n_chains;     /* chains per node held by this handle (extension: the reference runs one)        */
  int64_t chain_offset; /* global index of local chain 0: Philox counters use the global chain id, so a   */
                        /* sharded run draws the same streams as an unsharded one                          */
  uint64_t seed;        /* Philox4x32-10 key                                                               */
} dbn_config;

/*
 * Per-run model constants.  The reference hard-codes all of them inside each regressor's fit()
 * (SURVEY.md section 5 "Config / flags", Appendix C); the host wrapper fills in those literals by method.
 */
typedef struct {
  int32_t method;       /* dbn_method                                                                      */
  int32_t fan_in;       /* maximum parents; 3 in every reference regressor (e.g. bayesianPwLinearRegression.py:47) */
  int32_t with_tau;     /* 1: changepoint moves enabled ('varying_nh' / 'seq_coup_nh' / 'glob_coup_nh')    */
  int32_t num_samples;  /* regressor's self.num_samples: birth candidates {2..num_samples} and the cp hr   */
                        /* (N for nh/glob, N-1 for seq: network.py:160,216,240)                            */
  double T_ml;          /* T in the Gamma/pi/exponent terms of the marginal likelihood (Appendix C-1)      */
  double T_sigma;       /* T in the sigma^2 Gibbs shape (N+1 for h-dbn: bayesianLinearRegression.py:96)    */
  double a_sigma, b_sigma;   /* sigma^2 ~ IG(a,b): 0.005 (nh, seq) / 0.01 (h, glob)                        */
  double a_lambda, b_lambda; /* lambda^2 ~ IG(2, 0.2)                                                      */
  double a_delta, b_delta;   /* delta^2 ~ IG(2, 0.2)  (seq only)                                           */
  double cp_p;          /* changepoint prior hyper-parameter p = 0.125 (priors.py:20)                      */
  int32_t burn_in;      /* Network.burn_in (network.py:34)                                                 */
  int32_t thin;         /* 10 (network.py:359,389)                                                         */
} dbn_run_params;

/* ---- lifetime ---------------------------------------------------------------------------------------- */
int dbn_abi_version(void);
int dbn_create(const dbn_config* cfg, dbn_handle** out);
void dbn_destroy(dbn_handle* h);
/* message of the last failure on this handle (or of the last failed dbn_create when h == NULL) */
const char* dbn_last_error(const dbn_handle* h);
/* launch on an externally owned cudaStream_t (e.g. torch's current stream); NULL restores the handle's own.
 * Ordering contract: everything the handle queued on its previous stream is ordered before the work later calls
 * queue on the new one (an event recorded on the old stream is waited for on the new stream). */
int dbn_set_stream(dbn_handle* h, void* cuda_stream);
int dbn_sync(dbn_handle* h);

/* ---- kernel 1: time-prefix sufficient statistics ----------------------------------------------------- */
/*
 * Replaces the per-call re-slicing of design matrices: Network.set_network_configuration (network.py:54-141),
 * constructNdArray / constructResponseNdArray / constructDesignMatrix / selectData (utils.py:42-64,128-269).
 * `data` is the row-major [sum(seg_len), n] matrix of the concatenated time-series segments that
 * Network(data=[seg0, seg1, ...]) receives; lag must be 1.  Builds, on the device, for t = 0..N
 * (N = sum(seg_len - 1) regression rows, z_r = [1, x_r], y_r = x_{r+1}):
 *    ZZ[t] = sum_{r<t} z_r z_r^T (packed lower triangle), ZY[t][i] = sum_{r<t} z_r y_{r,i}, YY[t][i] = sum y_{r,i}^2
 * so any segment's Gram statistics are one difference of two rows.
 * The call returns once `data` has been handed to the copy engine: a pageable buffer is free again on return; a
 * PINNED (page-locked) buffer is read asynchronously and must stay unchanged until the next synchronising call on the
 * handle (dbn_sync, dbn_read_*), in exchange for no host synchronisation on this path.
 */
int dbn_build_stats(dbn_handle* h, const double* data, const int32_t* seg_len, int32_t n_seg, int32_t n, int32_t lag);
/* re-run the prefix kernels on the data already resident on the device (timing / roofline of kernel 1) */
int dbn_rebuild_stats(dbn_handle* h);
/* geometry of the table: regression rows N, doubles per table row, algorithmic bytes of one build */
int dbn_stats_info(const dbn_handle* h, int32_t* N, int64_t* row_doubles, int64_t* algorithmic_bytes);
/* copy table row t (0..N) to the host: [ZZ packed | per node of the table (all, or those of the unit range): ZY[0..n], YY]
 * (row_doubles values) */
int dbn_read_stats_row(dbn_handle* h, int32_t t, double* out_row);

/* ---- kernel 2: batched marginal-likelihood scorer ---------------------------------------------------- */
/*
 * Replaces calculateMarginalLikelihoodWithChangepoints (marginalLikelihood.py:91-150) and, with one segment, the
 * log of calculateMarginalLikelihood (:152-166).  Item b scores node[b] with parents pi[b, :pi_len[b]] (gene
 * ids), changepoints tau[b, :tau_len[b]] (1-based, last = N+2), lambda2[b], and either
 *   mode 0: the same prior mean mu[b, :k] for every segment (zeros for nh/h, the global mean for glob), or
 *   mode 1: the sequentially coupled prior means of betaTildeSampler (samplers.py:4-54) computed on the
 *           device from lambda2[b], delta2[b] ('seq-coup': delta2 replaces lambda2 for segments h > 0).
 * mu is [B, DBN_KMAX] in design-column order (intercept, then parents sorted by id); pi is [B, DBN_PMAX],
 * tau is [B, DBN_SMAX].  out_logml[b] receives the log marginal likelihood.
 */
int dbn_score_batch(dbn_handle* h, int32_t mode, double alpha, double beta, double T, int32_t B,
                    const int32_t* node, const int32_t* pi, const int32_t* pi_len, const int32_t* tau,
                    const int32_t* tau_len, const double* lambda2, const double* delta2, const double* mu,
                    double* out_logml);

/* ---- kernel 3: chain step (Gibbs draws, proposals, scoring, MH accept, count accumulation) ------------ */
/*
 * dbn_chain_init: initial state of every (node, chain) unit as in the regressors' fit() preambles —
 * pi = {}, tau = [N+2], sigma^2 = lambda^2 = delta^2 = 1, mu = 0 (bayesianPwLinearRegression.py:41-86) —
 * and zeroed count matrices.  Unit index u = node * n_chains + chain.
 */
int dbn_chain_init(dbn_handle* h, const dbn_run_params* p);
/*
 * Restrict the handle to the contiguous range [first_unit, first_unit + n_units) of the (node, chain) grid
 * u = node * n_chains + chain (n_units < 0: the whole grid again).  This is SURVEY.md 8e's partitioning: the reference
 * loops over the nodes (network.py:413-416) and every (node, chain) unit is independent, so a rank takes a slice of the
 * flat unit list, whatever the number of chains; `Network.fit(i)` of one node is the slice [i * n_chains, (i + 1) *
 * n_chains).  Philox counters use the grid's (chain, node), so the union of the slices is bit-identical to one handle
 * running the whole grid.  State, trace and replay arrays are indexed by the LOCAL unit (0 .. n_units - 1); the count
 * matrices keep their full [n, ...] shape (rows of nodes outside the slice stay zero) so that they can be summed over
 * ranks.  Takes effect with the next dbn_chain_init (an existing chain is dropped).
 * The prefix table follows the range: a handle with a unit range builds and keeps the ZZ block and the ZY | YY blocks of the
 * nodes its range touches only (row = [ZZ | blocks of nodes first_node .. last_node]; dbn_stats_info reports the shorter row,
 * dbn_score_batch refuses other nodes).  Set the range BEFORE dbn_build_stats to build the smaller table straight away; set
 * after it, the table is rebuilt for the new nodes from the series already on the device.
 */
int dbn_set_unit_range(dbn_handle* h, int64_t first_unit, int64_t n_units);

/* replay record per (unit, iteration): the reference's recorded draws */
#define DBN_RD_SIGMA2 0
#define DBN_RD_LAMBDA2 1
#define DBN_RD_DELTA2 2
#define DBN_RD_U_PI 3
#define DBN_RD_U_TAU 4
#define DBN_RD_MU_PI 5                     /* [DBN_KMAX] mu* drawn in the parent-set move (glob)  */
#define DBN_RD_MU_TAU (5 + DBN_KMAX)       /* [DBN_KMAX] mu* drawn in the changepoint move (glob) */
#define DBN_RD_BETA (5 + 2 * DBN_KMAX)     /* [DBN_SMAX][DBN_KMAX] beta_h                         */
#define DBN_RD_SIGMA2V (5 + 2 * DBN_KMAX + DBN_SMAX * DBN_KMAX) /* [DBN_SMAX] sigma^2_h (DBN_VV_GLOB_DBN) */
#define DBN_RD_STRIDE (5 + 2 * DBN_KMAX + DBN_SMAX * DBN_KMAX + DBN_SMAX)
#define DBN_RI_PI_MOVE 0
#define DBN_RI_PI_PICK 1                   /* [2] index into the candidate list of each np.random.choice */
#define DBN_RI_TAU_MOVE 3
#define DBN_RI_TAU_PICK 4                  /* [2] */
#define DBN_RI_STRIDE 8

/* trace record per (unit, iteration) */
#define DBN_KTRI (DBN_KMAX * (DBN_KMAX + 1) / 2)
#define DBN_TD_SIG_SHAPE 0
#define DBN_TD_SIG_RATE 1
#define DBN_TD_LAM_SHAPE 2
#define DBN_TD_LAM_RATE 3
#define DBN_TD_DEL_SHAPE 4
#define DBN_TD_DEL_RATE 5
#define DBN_TD_LOGML 6                     /* [4] pi*, pi, tau, tau* (NaN when the move was invalid)      */
#define DBN_TD_SIGMA2 10
#define DBN_TD_LAMBDA2 11
#define DBN_TD_DELTA2 12
#define DBN_TD_MU 13                       /* [DBN_KMAX] global mean after the iteration (glob)           */
#define DBN_TD_MUS_LOGDENS 18              /* [4] log q(mu*|pi*), log q(mu|pi), log q(mu|tau), log q(mu*|tau*) */
#define DBN_TD_MUS_MEAN 22                 /* [4][DBN_KMAX] MVN means of the four muSampler calls         */
#define DBN_TD_MUS_COV (22 + 4 * DBN_KMAX) /* [4][DBN_KTRI] their covariances, packed lower triangle     */
#define DBN_TD_BETA_MEAN (DBN_TD_MUS_COV + 4 * DBN_KTRI)            /* [DBN_SMAX][DBN_KMAX]              */
#define DBN_TD_BETA_COV (DBN_TD_BETA_MEAN + DBN_SMAX * DBN_KMAX)    /* [DBN_SMAX][DBN_KTRI]              */
#define DBN_TD_BETA (DBN_TD_BETA_COV + DBN_SMAX * DBN_KTRI)         /* [DBN_SMAX][DBN_KMAX] drawn beta_h */
/* DBN_VV_GLOB_DBN: per-segment sigma^2_h Gibbs shape / rate / draw; the scalar slots above hold segment 0's */
#define DBN_TD_SIGV_SHAPE (DBN_TD_BETA + DBN_SMAX * DBN_KMAX)       /* [DBN_SMAX] */
#define DBN_TD_SIGV_RATE (DBN_TD_SIGV_SHAPE + DBN_SMAX)             /* [DBN_SMAX] */
#define DBN_TD_SIGMA2V (DBN_TD_SIGV_RATE + DBN_SMAX)                /* [DBN_SMAX] */
#define DBN_TD_STRIDE (DBN_TD_SIGMA2V + DBN_SMAX)
#define DBN_TI_PI_VALID 0
#define DBN_TI_PI_ACCEPT 1
#define DBN_TI_TAU_VALID 2
#define DBN_TI_TAU_ACCEPT 3
#define DBN_TI_NPI 4
#define DBN_TI_PI 5                        /* [DBN_PMAX] parents after the iteration, reference order     */
#define DBN_TI_S 9                         /* len(tau) after the iteration                                 */
#define DBN_TI_TAU 10                      /* [DBN_SMAX]                                                   */
#define DBN_TI_PI_MOVE 20
#define DBN_TI_TAU_MOVE 21
#define DBN_TI_S_BEFORE 22                 /* len(tau) at the start of the iteration                       */
#define DBN_TI_STRIDE 24

/*
 * dbn_run: advance every unit by n_iter iterations of the regressor loop (a13): sigma^2, beta, lambda^2
 * (, delta^2) Gibbs draws (samplers.py), parent-set move (moves.py:528-706 / :417-526), changepoint move
 * (moves.py:129-236 / :25-126), and accumulation of the edge / changepoint counts of the thinned chain
 * (network.py:355-368,388-389).
 *   replay_d/replay_i : NULL for free-running Philox draws, else host arrays [units, n_iter, DBN_RD_STRIDE] /
 *                       [units, n_iter, DBN_RI_STRIDE] with the recorded draw stream to consume instead.
 *   trace_d/trace_i   : NULL, or host arrays [units, n_iter, DBN_TD_STRIDE] / [.., DBN_TI_STRIDE] to fill.
 * Without replay/trace the call is asynchronous on the handle's stream.
 */
/*
 * Start every unit from the changepoint list `tau` (1-based cut positions, increasing, the last one the pseudo
 * changepoint N + 2) instead of [N + 2].  Call between dbn_chain_init and the first dbn_run.  With
 * dbn_run_params.with_tau == 0 this is 'fixed_nh_dbn' (network.py:178-189: BayesianPieceWiseLinearRegression with
 * predefined change points, bayesianPwLinearRegression.py:117 skips the changepoint move).
 */
int dbn_set_changepoints(dbn_handle* h, const int32_t* tau, int32_t len);

int dbn_run(dbn_handle* h, int32_t n_iter, const double* replay_d, const int32_t* replay_i, double* trace_d,
            int32_t* trace_i);

/*
 * dbn_run_fp: dbn_run for DBN_FP_GLOB_DBN (fpGlobCoupBpwLinReg.py:68-110): sigma^2, beta_h, lambda^2 Gibbs draws with
 * the global mean mu as prior mean, then the globally coupled changepoint move (moves.py:25-126) with
 * mu* ~ muSampler (samplers.py:307-351).  Records are n = k wide, so their strides depend on n:
 *   replay_d [units, n_iter, 3 + k + DBN_SMAX*k + DBN_SMAX]: sigma2, lambda2, u_tau, mu*[k], beta[DBN_SMAX][k],
 *            sigma2_h[DBN_SMAX] (DBN_FP_VV_DBN only; its mu*[k] is the per-iteration vvMuSampler draw)
 *   replay_i [units, n_iter, 4]: tau_move, tau_pick[2], (pad)
 *   trace_d  [units, n_iter, 10 + 3k + 3*DBN_SMAX*k + 3*DBN_SMAX] (the last 3*DBN_SMAX: per-segment sigma^2_h shape, rate,
 *            draw of DBN_FP_VV_DBN): sig_shape, sig_rate, lam_shape, lam_rate, sigma2, lambda2,
 *            logml(tau), logml(tau*), log q(mu | tau), log q(mu* | tau*), mu after [k], muSampler mean (tau) [k],
 *            muSampler mean (tau*) [k], beta posterior means [DBN_SMAX][k], diag of the beta covariances [DBN_SMAX][k],
 *            drawn beta_h [DBN_SMAX][k]
 *   trace_i  [units, n_iter, 16]: tau_move, tau_valid, tau_accept, S before, S after, tau[DBN_SMAX], (pad)
 * DBN_FP_NH_DBN / DBN_FP_H_DBN use the same records (the mu fields stay NaN / zero); their sign counts are taken over the
 * kept draws of the beta chain instead: every beta_h of iteration it with it - 1 >= burn_in and (it - 1 - burn_in) % thin
 * == 0 is one sample (network.py:316-320, scores.py:139-153), nonempty[i] = #samples.
 * Counts for this method reuse the block below with: edge[i][j] = #kept mu samples of node i with mu_j >= 0,
 * nonempty[i] = #kept mu samples (network.py:309-311), and one more matrix appended after cp_kept:
 * neg[n][n] = #kept samples with mu_j <= 0; the edge score is max(edge, neg) / nonempty (credible_score,
 * scores.py:196-205 via network.py:339-347).
 */
int dbn_run_fp(dbn_handle* h, int32_t n_iter, const double* replay_d, const int32_t* replay_i, double* trace_d,
               int32_t* trace_i);

/*
 * Count matrices of the thinned chain, summed over this handle's chains (what gets all-reduced over ranks):
 *   edge[n][n]      edge[i][j] = #kept samples of node i whose parent set contains j   (scores.py:173-182)
 *   nonempty[n]     #kept samples with a non-empty parent set (the denominator, scores.py:184)
 *   cp_hist[n][N+3] cp_hist[i][c] = #kept tau samples of node i containing changepoint c
 *                   (examples/plot_diagnostics.py:226-240)
 *   cp_kept[n]      #kept tau samples
 * dbn_counts_size gives the total int64 count; the block is contiguous in that order on the device.
 */
int dbn_counts_size(const dbn_handle* h, int64_t* n_elems);
int dbn_read_counts(dbn_handle* h, int64_t* out /* n_elems */);
int dbn_device_counts(dbn_handle* h, void** device_ptr);

/* current state of every unit: pi [U, DBN_PMAX], pi_len [U], tau [U, DBN_SMAX], tau_len [U],
 * scalars [U, 3] = (sigma^2, lambda^2, delta^2), mu [U, DBN_KMAX]; any pointer may be NULL */
int dbn_read_state(dbn_handle* h, int32_t* pi, int32_t* pi_len, int32_t* tau, int32_t* tau_len, double* scalars,
                   double* mu);

/* work counters accumulated by dbn_run since the last dbn_chain_init / dbn_reset_counters:
 * [0] unit-iterations, [1] marginal-likelihood evaluations executed, [2] segment evaluations (all passes),
 * [3] algorithmic bytes of those segment evaluations (2*8*(k(k+1)/2+k+1) per segment, actual k),
 * [4] algorithmic FP64 flops of those segment evaluations, [5] accepted parent-set moves, [6] segment
 * factorisations actually executed (the h / nh kernel scores pi and pi* in one sweep and re-evaluates only the
 * segments a changepoint move changes, so this is smaller than [2]), [7] accepted changepoint moves
 * ([5]-[7] are filled by the h / nh kernel only) */
int dbn_read_counters(dbn_handle* h, uint64_t out[8]);
int dbn_reset_counters(dbn_handle* h);

/* Test hook for the h / nh kernel's per-unit cache of segment statistics: recomputes every cached row from the prefix
 * table and the current (pi, tau) on the device and compares bit for bit.  out[0] = units with a differing entry,
 * out[1] = differing entries; both 0 for methods without a cache. */
int dbn_verify_cache(dbn_handle* h, uint64_t out[2]);

/* dependent-free DFMA microbenchmark: measured FP64 FMA throughput of this device in GFLOP/s */
int dbn_measure_fp64_peak(dbn_handle* h, double* gflops);

/* device time (CUDA events on the handle's stream, milliseconds) of the scorer kernel of the last dbn_score_batch call:
 * the throughput figure of kernel 2 (marginalLikelihood.py:91-166 evaluated for a batch) without the host copies */
int dbn_score_batch_time(dbn_handle* h, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* DBN_B200_H */
