// Kernel 3 — the chain step: one thread per (node, chain) unit runs the regressor loop of the reference
// (a13: bayesianLinearRegression.py:114-147, bayesianPwLinearRegression.py:89-132,
// seqCoupledBayesianPwLinReg.py:68-121, globCoupBayesianPwLinReg.py:64-116) for `n_iter` iterations with its
// state in registers / shared memory, scoring every proposal from the prefix table.
#pragma once
#include "dbn_score.cuh"
#include "dbn_rng.cuh"
#include "../../include/dbn_b200.h"

namespace dbn {

constexpr int KMAX = DBN_KMAX;
constexpr int PMAX = DBN_PMAX;
constexpr int SMAX = DBN_SMAX;
constexpr int CHAIN_BLOCK = 128;

struct ChainArgs {
  Table tb;
  int method, fan_in, with_tau, num_samples, burn_in, thin;
  int C;        // local chains per node
  int U;        // units of this handle (a contiguous range of the n * C grid)
  int unit0;    // grid index node * C + chain of local unit 0
  int n_genes;  // nodes = genes; tb.n = n_genes * lag candidate feature columns (column l * n_genes + g = lag-(l+1) copy of gene g)
  int lag;
  int lanes;    // lanes of a warp that hold a unit (32 = every lane; fewer spread a small unit count over more warps and SMs)
  int it0;      // global index of the first iteration of this launch
  int n_iter;
  unsigned int chain_offset;
  unsigned long long seed;
  double T_sigma_half, a_sig, b_sig, a_lam, b_lam, a_del, b_del, log_p, log_1mp;
  double c0, T_half_plus_a, two_b;  // marginal-likelihood constants
  // state, struct of arrays over units
  int* st_npi;
  int* st_pi;   // [PMAX][U]
  int* st_S;
  int* st_tau;  // [SMAX][U]
  double* st_sig;
  double* st_lam;
  double* st_del;
  double* st_mu;  // [KMAX][U]
  // h / nh fast path (dbn_chain_fast.cuh): per-unit cache of the current state's boundary rows
  double* cache;  // [BSLOTS][NP][U] pairs of doubles: prefix-table rows at the segment boundaries (dbn_chain_fast.cuh)
  int* st_cg;     // [PMAX][U] gene id held by column slot s+1, -1 = empty
  int refresh;    // 1: rebuild the cache from the table at kernel start
  // counts (contiguous block): edge[n][n], nonempty[n], cp_hist[n][N+3], cp_kept[n]
  unsigned long long* edge;
  unsigned long long* nonempty;
  unsigned long long* cp_hist;
  unsigned long long* cp_kept;
  unsigned long long* counters;  // [8] (+ 16 probe counters of the diagnostic builds)
  // replay / trace (device copies, may be null)
  const double* rd;
  const int* ri;
  double* td;
  int* ti;
};

// ---- small sorted-set helpers on register arrays -------------------------------------------------------------
__device__ __forceinline__ void sort4(int (&a)[PMAX], int n) {
  // insertion sort of the first n entries, n <= 4, fully predicated
#pragma unroll
  for (int i = 1; i < PMAX; ++i) {
#pragma unroll
    for (int j = i; j > 0; --j) {
      if (j < n && a[j - 1] > a[j]) {
        int t = a[j];
        a[j] = a[j - 1];
        a[j - 1] = t;
      }
    }
  }
}

// p-th (0-based) column in {0..n_cols-1} \ ({node, node + n_genes, ..} u pi), ascending
// (np.setdiff1d(possibleFeaturesSet, pi)[p]; with lag > 1 the node's own copy in every lag block is no feature,
// network.py:85,92-104).  `sorted_pi` ascending; the own copies are ascending too, merged on the fly.
__device__ __forceinline__ int nth_candidate(int p, int node, const int (&sorted_pi)[PMAX], int npi, int n_genes = 0, int lag = 1) {
  int c = p;
  int own = node;        // next own copy not yet passed
  int own_left = lag;
#pragma unroll
  for (int j = 0; j < PMAX; ++j) {
    if (j < npi) {
      const int e = sorted_pi[j];
      for (; own_left > 0 && own < e; --own_left, own += n_genes)
        if (own <= c) ++c;
      if (e <= c) ++c;
    }
  }
  for (; own_left > 0; --own_left, own += n_genes)
    if (own <= c) ++c;
  return c;
}

// kept parent set -> edge counts: a gene counts once when any of its lag copies is a parent (scores.py:172-178)
__device__ __forceinline__ void count_edges(unsigned long long* edge_row, const int (&pi)[PMAX], int npi, int n_genes) {
#pragma unroll
  for (int j = 0; j < PMAX; ++j) {
    if (j < npi) {
      const int gene = pi[j] % n_genes;
      bool dup = false;
#pragma unroll
      for (int q = 0; q < PMAX; ++q) dup = dup || (q < j && pi[q] % n_genes == gene);
      if (!dup) atomicAdd(&edge_row[gene], 1ull);
    }
  }
}

// design columns of a parent set: f[0] = 0, f[1..] = sorted ids + 1
__device__ __forceinline__ void design_cols(const int (&pi)[PMAX], int npi, int (&f)[KMAX]) {
  int s[PMAX];
#pragma unroll
  for (int j = 0; j < PMAX; ++j) s[j] = pi[j];
  sort4(s, npi);
  f[0] = 0;
#pragma unroll
  for (int j = 0; j < PMAX; ++j) f[j + 1] = (j < npi) ? s[j] + 1 : 0;
}

// log of calculateChangePointsSetPrior (priors.py:4-46), tau(j) accessor, S = len(tau) <= 9
template <class TauFn>
__device__ __forceinline__ double log_cp_prior(TauFn tau, int S, double log_p, double log_1mp) {
  int prev = 1;
  double acc = 0.0;
  for (int j = 0; j < S - 1; ++j) {
    const int c = tau(j);
    acc += log_p + (double)(c - prev - 1) * log_1mp;
    prev = c;
  }
  return acc + (double)((tau(S - 1) - 1) - prev) * log_1mp;
}

// packed lower-triangular inverse covariance -> covariance (trace only)
template <int KM>
__device__ void invert_spd(const double (&Q)[KM * (KM + 1) / 2], double scale, double* out_packed) {
  double A[KM * (KM + 1) / 2], dinv[KM];
#pragma unroll
  for (int e = 0; e < KM * (KM + 1) / 2; ++e) A[e] = Q[e];
  ldl_factor<KM>(A, dinv);
#pragma unroll
  for (int col = 0; col < KM; ++col) {
    double x[KM];
#pragma unroll
    for (int i = 0; i < KM; ++i) x[i] = (i == col) ? 1.0 : 0.0;
    solve_ldl<KM>(A, dinv, x);
#pragma unroll
    for (int i = col; i < KM; ++i) out_packed[DBN_TRI(i, col)] = scale * x[i];
  }
}

// Gaussian defined by precision Q and mean numerator m (mean = Q^-1 m): factor once, then sample / evaluate
template <int KM>
struct GaussQ {
  double L[KM * (KM + 1) / 2], dinv[KM], mean[KM], logdetQ;
  __device__ __forceinline__ void setup(const double (&Q)[KM * (KM + 1) / 2], const double (&m)[KM]) {
#pragma unroll
    for (int e = 0; e < KM * (KM + 1) / 2; ++e) L[e] = Q[e];
    logdetQ = log(ldl_factor<KM>(L, dinv));
#pragma unroll
    for (int i = 0; i < KM; ++i) mean[i] = m[i];
    solve_ldl<KM>(L, dinv, mean);
  }
  // log N(x; mean, Q^-1) with k real dimensions
  __device__ __forceinline__ double logpdf(const double (&x)[KM], const double (&Q)[KM * (KM + 1) / 2], int k) const {
    double d[KM], Qd[KM];
#pragma unroll
    for (int i = 0; i < KM; ++i) d[i] = x[i] - mean[i];
    symv<KM>(Q, d, Qd);
    return -0.5 * ((double)k * 1.8378770664093454836 - logdetQ + dot<KM>(d, Qd));
  }
  // x = mean + L^-T D^-1/2 z
  __device__ __forceinline__ void sample(const double (&z)[KM], double (&x)[KM]) const {
#pragma unroll
    for (int i = 0; i < KM; ++i) x[i] = z[i] * sqrt(dinv[i]);
    solve_upper<KM>(L, x);
#pragma unroll
    for (int i = 0; i < KM; ++i) x[i] += mean[i];
  }
};

template <int KM>
__device__ __forceinline__ double sqnorm_k(const double (&x)[KM]) {
  return dot<KM>(x, x);
}

// ---- boundary-row cache of the coupled chain kernel ----------------------------------------------------------------------
// Row j of a unit holds the prefix-table row at the end boundary of the current state's segment j, restricted to the
// state's design columns in design order ([ZZ packed | ZY | YY | pad], zeros in the padding), as 16-byte pairs: pair p of row
// j of unit u at doubles ((j * NP + p) * U + u) * 2, so a warp moves a pair of consecutive units as one coalesced 512-byte
// request instead of 64 scattered sectors.  Rows are exact copies of table entries and a segment is the same difference of
// two rows, so the statistics are bit-identical to the gathered ones.  A pass streams the rows through two shared-memory
// staging buffers with cp.async, one row ahead: the L2 latency of row h + 1 is hidden behind the factorisation of segment h
// (reading the rows with plain loads was measured SLOWER than the round-1 gathers on the small mTOR table, whose gathers hit
// L1: profiles/r2_coupled).  Row SEG_ROW_NEW holds the one boundary a changepoint proposal adds; the row of boundary j of
// the scored changepoint list is  j == r_new ? SEG_ROW_NEW : j + (j > r_from ? r_shift : 0).
constexpr int SEG_ROW_NEW = 9, SEG_ROWS_CACHED = 10;
template <int KM>
struct SegDims {
  static constexpr int KT = KM * (KM + 1) / 2, NE = KT + KM + 1, NP = (NE + 1) / 2;
};
inline size_t chain_smem_bytes(int km) { return (size_t)2 * ((km * (km + 1) / 2 + km + 2) / 2) * 2 * CHAIN_BLOCK * sizeof(double); }
inline size_t chain_cache_doubles_per_unit(int km) { return (size_t)SEG_ROWS_CACHED * ((km * (km + 1) / 2 + km + 2) / 2) * 2; }

__device__ __forceinline__ void seg_cp_async16(double* smem, const double* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void seg_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void seg_cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int KM, class TauFn>
struct SegFromCache {
  static constexpr int KT = SegDims<KM>::KT, NP = SegDims<KM>::NP;
  const double* cu;   // cache + 2 u
  long long U;
  TauFn tau;
  int S, N;
  int r_new, r_from, r_shift;
  double* st;         // this thread's slice of the staging area: pair p of buffer b at st + (b * NP + p) * 2 * CHAIN_BLOCK
  __device__ __forceinline__ int row(int j) const { return j == r_new ? SEG_ROW_NEW : j + (j > r_from ? r_shift : 0); }
  __device__ __forceinline__ void issue(int j, int b) const {
    const double* p = cu + (long long)row(j) * NP * (2 * U);
    double* d = st + b * NP * (2 * CHAIN_BLOCK);
#pragma unroll
    for (int q = 0; q < NP; ++q) seg_cp_async16(d + q * (2 * CHAIN_BLOCK), p + (long long)q * (2 * U));
  }
  __device__ __forceinline__ void begin() const {
    if (S > 0) issue(0, 0);
    seg_cp_commit();
  }
  __device__ __forceinline__ int operator()(int h, double (&G)[KT], double (&g)[KM], double& s) const {
    const int lo = h ? DBN_SEG_END(tau(h - 1), h - 1, S, N) : 0;
    const int hi = DBN_SEG_END(tau(h), h, S, N);
    seg_cp_wait_all();
    const double* bh = st + (h & 1) * NP * (2 * CHAIN_BLOCK);
    const double* bl = st + ((h & 1) ^ 1) * NP * (2 * CHAIN_BLOCK);
    const bool hl = h > 0;   // segment 0 starts at table row 0, which is all zeros
#define DBN_SEGE(b, e) (b)[((e) >> 1) * (2 * CHAIN_BLOCK) + ((e) & 1)]
#pragma unroll
    for (int e = 0; e < KT; ++e) G[e] = DBN_SEGE(bh, e) - (hl ? DBN_SEGE(bl, e) : 0.0);
#pragma unroll
    for (int i = 0; i < KM; ++i) g[i] = DBN_SEGE(bh, KT + i) - (hl ? DBN_SEGE(bl, KT + i) : 0.0);
    s = DBN_SEGE(bh, KT + KM) - (hl ? DBN_SEGE(bl, KT + KM) : 0.0);
#undef DBN_SEGE
    // the operands are in registers: the next row may overwrite the buffer of row h - 1
    if (h + 1 < S) issue(h + 1, (h + 1) & 1);
    seg_cp_commit();
    return hi - lo;
  }
};
// one table row -> cache row j of this unit (cu = cache + 2 u)
template <int KM>
__device__ __forceinline__ void seg_cache_fill_row(double* cu, long long U, int j, const Table& t, const Cols<KM>& c, int table_row) {
  constexpr int KT = SegDims<KM>::KT, NE = SegDims<KM>::NE, NP = SegDims<KM>::NP;
  const double* __restrict__ p = t.base + (long long)table_row * t.row_stride;
  double v[2 * NP];
#pragma unroll
  for (int i = 0; i < KM; ++i) {
    const bool on = i < c.k;
    v[KT + i] = on ? __ldg(p + c.zy[i]) : 0.0;
#pragma unroll
    for (int q = 0; q <= i; ++q) v[DBN_TRI(i, q)] = on ? __ldg(p + c.zz[DBN_TRI(i, q)]) : 0.0;
  }
  v[KT + KM] = __ldg(p + c.yy);
  if (NE < 2 * NP) v[2 * NP - 1] = 0.0;
  double2* dst = reinterpret_cast<double2*>(cu) + (long long)j * NP * U;
#pragma unroll
  for (int q = 0; q < NP; ++q) dst[(long long)q * U] = make_double2(v[2 * q], v[2 * q + 1]);
}

#ifndef DBN_CHAIN_MINB
#define DBN_CHAIN_MINB 1
#endif
template <int METHOD, int KM, bool TRACE>
__global__ void __launch_bounds__(CHAIN_BLOCK, DBN_CHAIN_MINB) chain_kernel(const ChainArgs a) {
  constexpr bool SEQ = METHOD == DBN_SEQ_DBN;
  constexpr bool VV = METHOD == DBN_VV_GLOB_DBN;   // globally coupled, segment-specific variances (vvglobCoup.py:66-123)
  constexpr bool GLOB = METHOD == DBN_GLOB_DBN || VV;
  constexpr bool HOMOG = METHOD == DBN_H_DBN;
  constexpr int KT = KM * (KM + 1) / 2;
  constexpr int NB = (TRACE || SEQ) ? SMAX : 1;

  __shared__ int s_tau[SMAX * CHAIN_BLOCK];
  __shared__ int s_tas[SMAX * CHAIN_BLOCK];
  __shared__ double s_sigv[VV ? SMAX * CHAIN_BLOCK : 1];   // VV: sigma^2_h of the current iteration
  const int tid = threadIdx.x;
  // unit of this thread: dense, or `lanes` units per warp when the unit count does not fill the device (diverging lanes of
  // a warp serialise; a unit alone in its warp runs at single-thread latency, profiles/r2_small)
  const int u_raw = a.lanes >= 32 ? blockIdx.x * CHAIN_BLOCK + tid
                                  : ((tid & 31) < a.lanes ? (blockIdx.x * (CHAIN_BLOCK / 32) + (tid >> 5)) * a.lanes + (tid & 31) : a.U);
  const bool active = u_raw < a.U;   // lanes past the end run zero iterations so warp-wide reductions stay full
  const int u = active ? u_raw : 0;
  const int n_it = active ? a.n_iter : 0;
  const int node = (a.unit0 + u) / a.C;
  const int chain = (a.unit0 + u) - node * a.C;
  const int n = a.tb.n, N = a.tb.N, F = n - a.lag;   // n feature columns (genes x lag), the node's own copies are no candidates
  auto TAU = [&](int j) -> int& { return s_tau[j * CHAIN_BLOCK + tid]; };
  auto TAS = [&](int j) -> int& { return s_tas[j * CHAIN_BLOCK + tid]; };
  auto tau_cur = [&](int j) { return s_tau[j * CHAIN_BLOCK + tid]; };
  auto tau_star = [&](int j) { return s_tas[j * CHAIN_BLOCK + tid]; };
  auto SIGV = [&](int h) -> double& { return s_sigv[VV ? h * CHAIN_BLOCK + tid : 0]; };
  const double vv_c = VV ? a.a_sig * log(a.two_b) - lgamma(a.a_sig) : 0.0;

  // ---- load state
  int npi = a.st_npi[u];
  int pi[PMAX];
#pragma unroll
  for (int j = 0; j < PMAX; ++j) pi[j] = a.st_pi[j * a.U + u];
  int S = a.st_S[u];
  for (int j = 0; j < SMAX; ++j) TAU(j) = a.st_tau[j * a.U + u];
  double sig2 = a.st_sig[u], lam = a.st_lam[u], dlt = a.st_del[u];
  double mu[KM];
#pragma unroll
  for (int i = 0; i < KM; ++i) mu[i] = GLOB ? a.st_mu[i * a.U + u] : 0.0;

  Rng rng;
  rng.init(a.seed, a.chain_offset + (unsigned)chain, (unsigned)node);
  Work wk = {0, 0, 0, 0};
  const bool replay = TRACE && a.rd != nullptr;
  // per-unit boundary-row cache of the current state (SegFromCache, dbn_score.cuh): every pass over the current parent set
  // reads coalesced cached rows; only the proposed parent set is gathered from the table.  Round 1 gathered 2 x 21 scattered
  // doubles per segment in every pass and was bound by the L1 request path (profiles/r1_v0).
  extern __shared__ __align__(16) double s_stage[];   // [2][NP][CHAIN_BLOCK][2] staging of cached rows
  double* const cu = a.cache + 2ll * u;
  const long long U = a.U;
  double* const st = s_stage + 2 * tid;

  for (int li = 0; li < n_it; ++li) {
    const int it = a.it0 + li;
    rng.start_iteration((unsigned)it);
    const double* rd = replay ? a.rd + ((long long)u * a.n_iter + li) * DBN_RD_STRIDE : nullptr;
    const int* ri = replay ? a.ri + ((long long)u * a.n_iter + li) * DBN_RI_STRIDE : nullptr;
    double* td = (TRACE && a.td) ? a.td + ((long long)u * a.n_iter + li) * DBN_TD_STRIDE : nullptr;
    int* ti = (TRACE && a.ti) ? a.ti + ((long long)u * a.n_iter + li) * DBN_TI_STRIDE : nullptr;
    if (TRACE && ti) ti[DBN_TI_S_BEFORE] = S;

    int f[KMAX];
    design_cols(pi, npi, f);
    int k = npi + 1;
    Cols<KM> cols;
    {
      int fk[KM];
#pragma unroll
      for (int i = 0; i < KM; ++i) fk[i] = f[i];
      make_cols<KM>(cols, a.tb, node, fk, k);
    }
    // rows 0..S-1 of the cache <- table rows at the current segment ends, current design columns (launch start, accepted moves)
    auto fill_rows = [&]() {
      for (int j = 0; j < S; ++j) seg_cache_fill_row<KM>(cu, U, j, a.tb, cols, DBN_SEG_END(tau_cur(j), j, S, N));
      __threadfence_block();   // the stores are ordered before the asynchronous copies that read the rows back
    };
    if (li == 0) fill_rows();
    auto seg_cur = [&]() { return SegFromCache<KM, decltype(tau_cur)>{cu, U, tau_cur, S, N, -1, 99, 0, st}; };

    // =========================== Gibbs pass: sigma^2, beta, lambda^2 (, delta^2) ===========================
    // One sweep over the segments with the OLD lambda^2/delta^2.  beta_h = mean_h + sqrt(sigma^2) e_h with
    // e_h = sqrt(w) L^-T D^-1/2 z_h independent of sigma^2, so sum_h |beta_h - c_h|^2 is assembled from
    // (|d|^2, d.e, |e|^2) after sigma^2 has been drawn: no second sweep, no per-segment storage (nh/glob/h).
    double q_sum = 0.0, acc_dd = 0.0, acc_de = 0.0, acc_ee = 0.0;
    double acc_ss = 0.0;  // VV: sum_h |beta_h - mu|^2 / sigma^2_h (samplers.py:275-282)
    double bm[NB][KM], be[NB][KM];  // seq / trace: per-segment mean and e (or injected beta in replay)
    {
      double bt_cur[KM], bt_prev[KM];
#pragma unroll
      for (int i = 0; i < KM; ++i) bt_cur[i] = bt_prev[i] = 0.0;
      int lo = 0;
      seg_cur().begin();
      for (int h = 0; h < S; ++h) {
        const int hi = DBN_SEG_END(tau_cur(h), h, S, N);
        double G[KT], g[KM], s, A[KT], dinv[KM], wv[KM];
        seg_cur()(h, G, g, s);
        if (SEQ && h == 1) {
          form_a<KM>(G, lam, A);
          ldl_factor<KM>(A, dinv);
#pragma unroll
          for (int i = 0; i < KM; ++i) bt_cur[i] = lam * g[i];
          solve_ldl<KM>(A, dinv, bt_cur);
          wk.segment(k, false);
        }
        const double w = (SEQ && h > 0) ? dlt : lam;
        form_a<KM>(G, w, A);
        ldl_factor<KM>(A, dinv);
        double pm[KM];  // prior mean of this segment
#pragma unroll
        for (int i = 0; i < KM; ++i) pm[i] = SEQ ? bt_cur[i] : mu[i];
        double q_h;
        if constexpr (SEQ || GLOB)
          q_h = segment_q<KM, true>(G, g, s, pm, w, A, dinv, wv);
        else
          q_h = segment_q<KM, false>(G, g, s, pm, w, A, dinv, wv);
        q_sum += q_h;
        wk.segment(k, SEQ || GLOB);
        double sig_h = 1.0;
        if constexpr (VV) {
          // segmentSigmaSampler (samplers.py:90-101): sigmaSqrSampler on segment h alone, numSamples = T = T_h
          const double shape = a.a_sig + 0.5 * (double)(hi - lo);
          const double rate = a.b_sig + 0.5 * q_h;
          sig_h = replay ? rd[DBN_RD_SIGMA2V + h] : rate / rng.gamma(shape);
          SIGV(h) = sig_h;
          if (TRACE && td) {
            td[DBN_TD_SIGV_SHAPE + h] = shape;
            td[DBN_TD_SIGV_RATE + h] = rate;
            td[DBN_TD_SIGMA2V + h] = sig_h;
            if (h == 0) {
              td[DBN_TD_SIG_SHAPE] = shape;
              td[DBN_TD_SIG_RATE] = rate;
              td[DBN_TD_SIGMA2] = sig_h;
            }
          }
          if (h == 0) sig2 = sig_h;
        }
        // posterior mean of beta_h: (I/w + G)^-1 (pm/w + g) = A^-1 (pm + w g)   (samplers.py:210-212)
        double mean[KM];
#pragma unroll
        for (int i = 0; i < KM; ++i) mean[i] = pm[i] + w * g[i];
        solve_ldl<KM>(A, dinv, mean);
        double e[KM];
        if (!replay) {
          // e = sqrt(w) L^-T D^-1/2 z  (cov = sigma^2 w A^-1, samplers.py:214-216)
          double z[KM];
#pragma unroll
          for (int i = 0; i < KM; i += 2) {
            double z0, z1;
            rng.normal2(z0, z1);
            z[i] = (i < k) ? z0 : 0.0;
            if (i + 1 < KM) z[i + 1] = (i + 1 < k) ? z1 : 0.0;
          }
          const double sw = sqrt(w);
#pragma unroll
          for (int i = 0; i < KM; ++i) e[i] = sw * z[i] * sqrt(dinv[i]);
          solve_upper<KM>(A, e);
        } else {
#pragma unroll
          for (int i = 0; i < KM; ++i) e[i] = (i < k) ? rd[DBN_RD_BETA + h * KMAX + i] : 0.0;  // injected beta_h
        }
        if (TRACE && td) {
#pragma unroll
          for (int i = 0; i < KM; ++i) td[DBN_TD_BETA_MEAN + h * KMAX + i] = (i < k) ? mean[i] : NAN;
          // cov / sigma^2 = w A^-1 ; scaled by sigma^2 below
          double Aorig[KT];
          form_a<KM>(G, w, Aorig);
          double covp[KT];
          invert_spd<KM>(Aorig, w, covp);
#pragma unroll
          for (int i = 0; i < KM; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) td[DBN_TD_BETA_COV + h * DBN_KTRI + DBN_TRI(i, j)] = (i < k) ? covp[DBN_TRI(i, j)] : NAN;
        }
        if constexpr (NB > 1) {
#pragma unroll
          for (int i = 0; i < KM; ++i) {
            bm[h][i] = mean[i];
            be[h][i] = e[i];
          }
        }
        if (!SEQ) {
          // centre c_h = mu (zeros for nh/h): samplers.py:279, :299
          if (!replay) {
            double d[KM];
#pragma unroll
            for (int i = 0; i < KM; ++i) d[i] = mean[i] - pm[i];
            if constexpr (VV) {
              acc_ss += (dot<KM>(d, d) + 2.0 * sqrt(sig_h) * dot<KM>(d, e)) / sig_h + dot<KM>(e, e);
            } else {
              acc_dd += dot<KM>(d, d);
              acc_de += dot<KM>(d, e);
              acc_ee += dot<KM>(e, e);
            }
          } else {
            double d[KM];
#pragma unroll
            for (int i = 0; i < KM; ++i) d[i] = (i < k) ? e[i] - pm[i] : 0.0;
            if constexpr (VV)
              acc_ss += dot<KM>(d, d) / sig_h;
            else
              acc_dd += dot<KM>(d, d);
          }
        }
        if (SEQ && h >= 1) {
          double nxt[KM];
#pragma unroll
          for (int i = 0; i < KM; ++i) nxt[i] = bt_prev[i] + dlt * g[i];
          solve_ldl<KM>(A, dinv, nxt);
#pragma unroll
          for (int i = 0; i < KM; ++i) {
            bt_prev[i] = bt_cur[i];
            bt_cur[i] = nxt[i];
          }
        }
        lo = hi;
      }
    }
    // ---- sigma^2 ~ IG: shape a + T/2, rate b + q/2 (samplers.py:118-125, :81-86, :135-141)
    if constexpr (!VV) {
      const double shape = a.a_sig + a.T_sigma_half;
      const double rate = a.b_sig + 0.5 * q_sum;
      sig2 = replay ? rd[DBN_RD_SIGMA2] : rate / rng.gamma(shape);
      if (TRACE && td) {
        td[DBN_TD_SIG_SHAPE] = shape;
        td[DBN_TD_SIG_RATE] = rate;
        td[DBN_TD_SIGMA2] = sig2;
      }
    }
    const double sg = sqrt(sig2);
    auto sig_of = [&](int h) -> double { return VV ? SIGV(h) : sig2; };   // variance of segment h
    if (TRACE && td) {
      for (int h = 0; h < S; ++h) {
        const double s2h = sig_of(h), sgh = sqrt(s2h);
        for (int i = 0; i < KM; ++i) {
          const double b = replay ? be[h < NB ? h : 0][i] : bm[h < NB ? h : 0][i] + sgh * be[h < NB ? h : 0][i];
          td[DBN_TD_BETA + h * KMAX + i] = (i < k) ? b : NAN;
        }
        for (int e = 0; e < DBN_KTRI; ++e) td[DBN_TD_BETA_COV + h * DBN_KTRI + e] *= s2h;
      }
    }
    // ---- lambda^2
    {
      double shape, rate;
      if constexpr (SEQ) {
        // samplers.py:249-263: shape a + (k+1)/2, rate b + beta_0.beta_0 / (2 sigma^2)
        double b0[KM];
#pragma unroll
        for (int i = 0; i < KM; ++i) b0[i] = replay ? be[0][i] : bm[0][i] + sg * be[0][i];
        shape = a.a_lam + 0.5 * (double)(k + 1);
        rate = a.b_lam + 0.5 * dot<KM>(b0, b0) / sig2;
      } else if constexpr (VV) {
        // samplers.py:265-294 with the list of sigma^2_h: shape a + S k / 2, rate b + sum_h |beta_h - mu|^2 / (2 sigma^2_h)
        shape = a.a_lam + 0.5 * (double)(S * k);
        rate = a.b_lam + 0.5 * acc_ss;
      } else {
        const double ss = replay ? acc_dd : (acc_dd + 2.0 * sg * acc_de + sig2 * acc_ee);
        // nh/glob: samplers.py:285-289 shape a + S k/2 ; h: :301 shape a + k/2 ; rate b + ss / (2 sigma^2)
        shape = a.a_lam + 0.5 * (double)(HOMOG ? k : S * k);
        rate = a.b_lam + 0.5 * ss / sig2;
      }
      const double lam_new = replay ? rd[DBN_RD_LAMBDA2] : rate / rng.gamma(shape);
      if (TRACE && td) {
        td[DBN_TD_LAM_SHAPE] = shape;
        td[DBN_TD_LAM_RATE] = rate;
        td[DBN_TD_LAMBDA2] = lam_new;
      }
      if constexpr (SEQ) {
        // delta^2 (samplers.py:220-247): beta-tildes recomputed with (lambda^2 new, delta^2 old), beta_h centred
        // on bt[h-1] for h >= 1
        double acc = 0.0;
        double bt_cur[KM], bt_prev[KM];
#pragma unroll
        for (int i = 0; i < KM; ++i) bt_cur[i] = bt_prev[i] = 0.0;
        int lo = 0;
        seg_cur().begin();
        for (int h = 0; h < S; ++h) {
          const int hi = DBN_SEG_END(tau_cur(h), h, S, N);
          double G[KT], g[KM], s, A[KT], dinv[KM];
          seg_cur()(h, G, g, s);   // the rows are streamed in order: segment 0 is staged although only its end row is used
          if (h >= 1) {
            if (h == 1) {
              form_a<KM>(G, lam_new, A);
              ldl_factor<KM>(A, dinv);
#pragma unroll
              for (int i = 0; i < KM; ++i) bt_cur[i] = lam_new * g[i];
              solve_ldl<KM>(A, dinv, bt_cur);
              wk.segment(k, false);
            }
            double d[KM];
#pragma unroll
            for (int i = 0; i < KM; ++i) {
              const double b = replay ? be[h < NB ? h : 0][i] : bm[h < NB ? h : 0][i] + sg * be[h < NB ? h : 0][i];
              d[i] = b - bt_prev[i];
            }
            acc += dot<KM>(d, d);
            if (h + 1 < S) {
              form_a<KM>(G, dlt, A);
              ldl_factor<KM>(A, dinv);
              double nxt[KM];
#pragma unroll
              for (int i = 0; i < KM; ++i) nxt[i] = bt_prev[i] + dlt * g[i];
              solve_ldl<KM>(A, dinv, nxt);
#pragma unroll
              for (int i = 0; i < KM; ++i) {
                bt_prev[i] = bt_cur[i];
                bt_cur[i] = nxt[i];
              }
              wk.segment(k, false);
            }
          }
          lo = hi;
        }
        const double dshape = a.a_del + 0.5 * (double)((S - 1) * k);
        const double drate = a.b_del + 0.5 * acc / sig2;
        const double dlt_new = replay ? rd[DBN_RD_DELTA2] : drate / rng.gamma(dshape);
        if (TRACE && td) {
          td[DBN_TD_DEL_SHAPE] = dshape;
          td[DBN_TD_DEL_RATE] = drate;
          td[DBN_TD_DELTA2] = dlt_new;
        }
        dlt = dlt_new;
      }
      lam = lam_new;
    }

    // unified scorer for the current method; `seg` = segment source (table gather or cached rows)
    auto score = [&](auto seg, int S_, int k_, const double(&mu_)[KM]) -> double {
      if constexpr (SEQ)
        return score_seq_seg<KM>(seg, S_, k_, lam, dlt, a.c0, a.T_half_plus_a, a.two_b, wk);
      else if constexpr (VV)
        return score_vv_seg<KM>(seg, S_, k_, lam, mu_, a.a_sig, vv_c, a.two_b, wk);
      else if constexpr (GLOB)
        return score_plain_seg<KM, true>(seg, S_, k_, lam, mu_, a.c0, a.T_half_plus_a, a.two_b, wk);
      else
        return score_plain_seg<KM, false>(seg, S_, k_, lam, mu_, a.c0, a.T_half_plus_a, a.two_b, wk);
    };

    // ======================================= parent-set move (a4 / a6) =======================================
    double ml_now = NAN;   // log-ML of the state after the pi move (reused by the tau move)
    bool have_ml_now = false;
    double Qc[KT], m0c[KM];  // glob: muSampler precision / mean numerator of the current state
    bool have_qc = false;
    {
      const int move = replay ? ri[DBN_RI_PI_MOVE] : rng.below(3);
      int star[PMAX];
      int nstar = npi;
      bool valid = true;
      double log_hr = 0.0;
      int spi[PMAX];
#pragma unroll
      for (int j = 0; j < PMAX; ++j) spi[j] = pi[j];
      sort4(spi, npi);
#pragma unroll
      for (int j = 0; j < PMAX; ++j) star[j] = pi[j];
      if (move == 0) {  // addMove (utils.py:293-305)
        const int ncand = F - npi;
        if (npi > a.fan_in - 1 || ncand <= 0) {
          valid = false;
        } else {
          const int p = replay ? ri[DBN_RI_PI_PICK] : rng.below(ncand);
          const int c = nth_candidate(p, node, spi, npi, a.n_genes, a.lag);
#pragma unroll
          for (int j = 0; j < PMAX; ++j)
            if (j == npi) star[j] = c;
          nstar = npi + 1;
          log_hr = log((double)(F - npi) / (double)nstar);  // moves.py:612
        }
      } else {
        if (npi < 1) {
          valid = false;  // utils.py:285-286, :308-309
        } else {
          const int idx = replay ? ri[DBN_RI_PI_PICK] : rng.below(npi);
          int el = pi[0];
#pragma unroll
          for (int j = 1; j < PMAX; ++j)
            if (j == idx) el = pi[j];
          // sorted(pi \ el)  (np.setdiff1d, utils.py:290, :318)
          int w_ = 0;
#pragma unroll
          for (int j = 0; j < PMAX; ++j) {
            if (j < npi && spi[j] != el) {
#pragma unroll
              for (int q = 0; q < PMAX; ++q)
                if (q == w_) star[q] = spi[j];
              ++w_;
            }
          }
          if (move == 1) {  // deleteMove
            nstar = npi - 1;
            log_hr = log((double)npi / (double)(F - nstar));  // moves.py:615
          } else {  // exchangeMove (utils.py:307-323)
            const int ncand = F - npi;
            if (ncand <= 0) {
              valid = false;
            } else {
              const int p = replay ? ri[DBN_RI_PI_PICK + 1] : rng.below(ncand);
              const int c = nth_candidate(p, node, spi, npi, a.n_genes, a.lag);
#pragma unroll
              for (int j = 0; j < PMAX; ++j)
                if (j == npi - 1) star[j] = c;
              nstar = npi;
            }
          }
        }
      }
      if (TRACE && ti) {
        ti[DBN_TI_PI_MOVE] = move;
        ti[DBN_TI_PI_VALID] = valid;
        ti[DBN_TI_PI_ACCEPT] = 0;
      }
      if (TRACE && td) {
        td[DBN_TD_LOGML + 0] = NAN;
        td[DBN_TD_LOGML + 1] = NAN;
        td[DBN_TD_LOGML + 2] = NAN;
        td[DBN_TD_LOGML + 3] = NAN;
        for (int e = DBN_TD_MUS_LOGDENS; e < DBN_TD_BETA_MEAN; ++e) td[e] = NAN;
      }
      if (valid) {
        int fs[KMAX];
        design_cols(star, nstar, fs);
        const int ks = nstar + 1;
        Cols<KM> cstar;
        {
          int fk[KM];
#pragma unroll
          for (int i = 0; i < KM; ++i) fk[i] = fs[i];
          make_cols<KM>(cstar, a.tb, node, fk, ks);
        }
        double ml_star, ml_cur, log_ratio;
        double mu_star[KM];
#pragma unroll
        for (int i = 0; i < KM; ++i) mu_star[i] = 0.0;
        double Qs[KT], m0s[KM];
        if constexpr (GLOB) {
          // moves.py:476-491: mu* ~ muSampler(0 | pi*), reverse density muSampler(mu | pi) at the current mu
          // (VV: moves.py:361-375; vvMuSampler has segment-specific variances and no mu_in term, samplers.py:353-394)
          // one sweep per parent set (not VV): the muSampler sums also give the log-ML at any mean (ml_from_sums)
          constexpr bool ONE = !VV;
          const MuSums sums_s = mu_posterior_seg<KM, ONE>(seg_from_table<KM>(a.tb, cstar, tau_cur, S), S, ks, lam, sig_of, Qs, m0s, wk);
          GaussQ<KM> gs;
          gs.setup(Qs, m0s);
          if (replay) {
#pragma unroll
            for (int i = 0; i < KM; ++i) mu_star[i] = (i < ks) ? rd[DBN_RD_MU_PI + i] : 0.0;
          } else {
            double z[KM];
#pragma unroll
            for (int i = 0; i < KM; i += 2) {
              double z0, z1;
              rng.normal2(z0, z1);
              z[i] = (i < ks) ? z0 : 0.0;
              if (i + 1 < KM) z[i + 1] = (i + 1 < ks) ? z1 : 0.0;
            }
            gs.sample(z, mu_star);
          }
          const double ld_star = gs.logpdf(mu_star, Qs, ks);
          if constexpr (ONE) {
            ml_star = ml_from_sums<KM>(sums_s, Qs, m0s, sig2, lam, mu_star, a.c0, a.T_half_plus_a, a.two_b, wk);
            const MuSums sums_c = mu_posterior_seg<KM, true>(seg_cur(), S, k, lam, sig_of, Qc, m0c, wk);
            ml_cur = ml_from_sums<KM>(sums_c, Qc, m0c, sig2, lam, mu, a.c0, a.T_half_plus_a, a.two_b, wk);
          } else {
            ml_star = score(seg_from_table<KM>(a.tb, cstar, tau_cur, S), S, ks, mu_star);
            // var-glob: the current state's vvLogMargLikelihood (at the known mean) and its muSampler sums in one sweep
            double acc_vv = 0.0;
            mu_posterior_seg<KM, false>(seg_cur(), S, k, lam, sig_of, Qc, m0c, wk,
                                        [&](int len, const double(&G)[KT], const double(&g)[KM], double s, const double(&A)[KT],
                                            const double(&dinv)[KM], double prod) {
                                          double wv[KM];
                                          const double q = segment_q<KM, true>(G, g, s, mu, lam, A, dinv, wv);
                                          const double Th_half = 0.5 * (double)len;
                                          acc_vv += lgamma(Th_half + a.a_sig) + vv_c - Th_half * 1.1447298858494002 - 0.5 * log(prod) -
                                                    (Th_half + a.a_sig) * log(a.two_b + q);
                                          wk.flops += 12;
                                        });
            wk.eval();
            ml_cur = acc_vv;
          }
          have_qc = true;
          double mc[KM];
#pragma unroll
          for (int i = 0; i < KM; ++i) mc[i] = m0c[i] + (VV ? 0.0 : mu[i]);  // samplers.py:336: the input mu is the prior mean term
          GaussQ<KM> gc;
          gc.setup(Qc, mc);
          const double ld_cur = gc.logpdf(mu, Qc, k);
          // moves.py:501-516
          const double lp_star = -0.5 * ((double)ks * 1.8378770664093454836 + sqnorm_k<KM>(mu_star));
          const double lp_cur = -0.5 * ((double)k * 1.8378770664093454836 + sqnorm_k<KM>(mu));
          log_ratio = ml_star - ml_cur + ld_cur - ld_star + lp_star - lp_cur + log_hr;
          if (TRACE && td) {
            td[DBN_TD_MUS_LOGDENS + 0] = ld_star;
            td[DBN_TD_MUS_LOGDENS + 1] = ld_cur;
            for (int i = 0; i < KM; ++i) {
              td[DBN_TD_MUS_MEAN + 0 * KMAX + i] = (i < ks) ? gs.mean[i] : NAN;
              td[DBN_TD_MUS_MEAN + 1 * KMAX + i] = (i < k) ? gc.mean[i] : NAN;
            }
            double cv[KT];
            invert_spd<KM>(Qs, 1.0, cv);
            for (int i = 0; i < KM; ++i)
              for (int j = 0; j <= i; ++j) td[DBN_TD_MUS_COV + 0 * DBN_KTRI + DBN_TRI(i, j)] = (i < ks) ? cv[DBN_TRI(i, j)] : NAN;
            invert_spd<KM>(Qc, 1.0, cv);
            for (int i = 0; i < KM; ++i)
              for (int j = 0; j <= i; ++j) td[DBN_TD_MUS_COV + 1 * DBN_KTRI + DBN_TRI(i, j)] = (i < k) ? cv[DBN_TRI(i, j)] : NAN;
          }
        } else {
          ml_star = score(seg_from_table<KM>(a.tb, cstar, tau_cur, S), S, ks, mu_star);   // moves.py:574-576 / :569-571 / :672
          ml_cur = score(seg_cur(), S, k, mu);          // moves.py:599-600 / :594-595 / :681
          log_ratio = ml_star - ml_cur + log_hr;         // uniform parent-set prior cancels (priors.py:48-68)
        }
        const double uacc = replay ? rd[DBN_RD_U_PI] : rng.uniform();
        // nh/seq/h: u < min(1, exp(d)) (moves.py:620-633, :699-702); glob: u < exp(min(1, d)) (moves.py:512-520)
        const bool acc = GLOB ? (uacc < exp(fmin(1.0, log_ratio))) : (uacc < fmin(1.0, exp(log_ratio)));
        if (TRACE && td) {
          td[DBN_TD_LOGML + 0] = ml_star;
          td[DBN_TD_LOGML + 1] = ml_cur;
        }
        if (TRACE && ti) ti[DBN_TI_PI_ACCEPT] = acc;
        ml_now = acc ? ml_star : ml_cur;
        have_ml_now = true;
        if (acc) {
          npi = nstar;
          k = ks;
#pragma unroll
          for (int j = 0; j < PMAX; ++j) pi[j] = star[j];
          cols = cstar;
          fill_rows();   // the cached rows follow the accepted parent set
          if (GLOB) {
#pragma unroll
            for (int i = 0; i < KM; ++i) {
              mu[i] = mu_star[i];
              m0c[i] = m0s[i];
            }
#pragma unroll
            for (int e = 0; e < KT; ++e) Qc[e] = Qs[e];
          }
        }
      }
      // ---- counts: pi_vector[it+1] is kept iff idx >= burn_in and (idx - burn_in) % thin != 0 (network.py:358-359)
      const int idx = it + 1;
      if (idx >= a.burn_in && ((idx - a.burn_in) % a.thin) != 0 && npi > 0) {
        atomicAdd(&a.nonempty[node], 1ull);
        count_edges(a.edge + (long long)node * a.n_genes, pi, npi, a.n_genes);
      }
      if (TRACE && ti) {
        ti[DBN_TI_NPI] = npi;
        for (int j = 0; j < PMAX; ++j) ti[DBN_TI_PI + j] = (j < npi) ? pi[j] : -1;
      }
    }

    // ======================================= changepoint move (a5 / a6) ======================================
    if (!HOMOG && a.with_tau) {
      const int move = replay ? ri[DBN_RI_TAU_MOVE] : rng.below(3);
      const int ns = a.num_samples;
      bool valid = true;
      int Ss = S;
      double log_hr = 0.0;
      int r_new = -1, r_from = 99, r_shift = 0, new_cp = 0;   // cached row of boundary j of tau* (SegFromCache) and the added changepoint
      if (move == 0) {  // cpBirthMove (changepointMoves.py:93-124)
        int inside = 0;
        for (int j = 0; j < S - 1; ++j) inside += (TAU(j) >= 2 && TAU(j) <= ns) ? 1 : 0;
        const int ncand = ns - 1 - inside;
        if (ncand <= 0) {
          valid = false;
        } else {
          const int p = replay ? ri[DBN_RI_TAU_PICK] : rng.below(ncand);
          int c = 2 + p;
          for (int j = 0; j < S - 1; ++j)
            if (TAU(j) <= c) ++c;
          int w_ = 0;
          bool placed = false;
          for (int j = 0; j < S; ++j) {
            if (!placed && c < TAU(j)) {
              r_new = r_from = w_;
              r_shift = -1;
              new_cp = c;
              TAS(w_++) = c;
              placed = true;
            }
            TAS(w_++) = TAU(j);
          }
          Ss = S + 1;
          if (Ss > 9) valid = false;  // moves.py:156 (== 10) / :42 (> 9)
          // moves.py:159 (nh/seq) / :46 (glob)
          log_hr = GLOB ? log((double)(ns - 1 - S) / (double)Ss) : log((double)(ns - S) / (double)(Ss - 1));
        }
      } else if (S < 2) {
        valid = false;  // changepointMoves.py:81-82, :21-22
      } else if (move == 1) {  // cpDeathMove (:63-91)
        const int idx = replay ? ri[DBN_RI_TAU_PICK] : rng.below(S - 1);
        int w_ = 0;
        for (int j = 0; j < S; ++j)
          if (j != idx) TAS(w_++) = TAU(j);
        Ss = S - 1;
        r_from = idx - 1;
        r_shift = 1;
        // moves.py:163 (nh/seq) / :55 (glob)
        log_hr = GLOB ? log((double)S / (double)(ns - 1 - Ss)) : log((double)(S - 1) / (double)(ns - Ss));
      } else {  // cpRellocationMove (:3-61)
        const int idx = replay ? ri[DBN_RI_TAU_PICK] : rng.below(S - 1);
        const int last = TAU(S - 1) - 1;
        const int left = (idx == 0) ? 1 : TAU(idx - 1);
        const int right = (idx + 1 == S - 1) ? last : TAU(idx + 1);
        const int ncand = right - left - 2;
        if (ncand <= 0) {
          valid = false;
        } else {
          const int p = replay ? ri[DBN_RI_TAU_PICK + 1] : rng.below(ncand);
          int c = left + 1 + p;
          if (c >= TAU(idx)) ++c;
          r_new = idx;
          new_cp = c;
          for (int j = 0; j < S; ++j) TAS(j) = (j == idx) ? c : TAU(j);
        }
      }
      if (TRACE && ti) {
        ti[DBN_TI_TAU_MOVE] = move;
        ti[DBN_TI_TAU_VALID] = valid;
        ti[DBN_TI_TAU_ACCEPT] = 0;
      }
      if (valid) {
        if (r_new >= 0) {   // the one boundary tau* adds
          seg_cache_fill_row<KM>(cu, U, SEG_ROW_NEW, a.tb, cols, new_cp - 1);
          __threadfence_block();
        }
        const SegFromCache<KM, decltype(tau_star)> seg_star{cu, U, tau_star, Ss, N, r_new, r_from, r_shift, st};
        double ml_cur = NAN;
        if constexpr (GLOB && !VV) {   // one sweep: muSampler sums of the current state and its log-ML (moves.py:67-73)
          if (!have_qc) {
            const MuSums sums_c = mu_posterior_seg<KM, true>(seg_cur(), S, k, lam, sig_of, Qc, m0c, wk);
            ml_cur = ml_from_sums<KM>(sums_c, Qc, m0c, sig2, lam, mu, a.c0, a.T_half_plus_a, a.two_b, wk);
            have_qc = true;
          } else {
            ml_cur = ml_now;
          }
        } else {
          ml_cur = have_ml_now ? ml_now : score(seg_cur(), S, k, mu);  // moves.py:184-191 / :67-69
        }
        double ml_star, log_ratio;
        const double lprior = log_cp_prior(tau_star, Ss, a.log_p, a.log_1mp) - log_cp_prior(tau_cur, S, a.log_p, a.log_1mp);
        double mu_star[KM];
#pragma unroll
        for (int i = 0; i < KM; ++i) mu_star[i] = 0.0;
        if constexpr (VV) {
          // vvGlobCoupTauMove (moves.py:238-307): mu is kept, so only the two marginal likelihoods enter
#pragma unroll
          for (int i = 0; i < KM; ++i) mu_star[i] = mu[i];
          ml_star = score(seg_star, Ss, k, mu);        // :283-284
          log_ratio = ml_star - ml_cur + lprior + log_hr;  // :291-295
        } else if constexpr (GLOB) {
          // moves.py:72-93: density of the current mu under muSampler(mu | tau), then mu* ~ muSampler(0 | tau*)
          if (!have_qc) mu_posterior_seg<KM>(seg_cur(), S, k, lam, sig_of, Qc, m0c, wk);
          double mc[KM];
#pragma unroll
          for (int i = 0; i < KM; ++i) mc[i] = m0c[i] + mu[i];
          GaussQ<KM> gc;
          gc.setup(Qc, mc);
          const double ld_cur = gc.logpdf(mu, Qc, k);
          double Qs[KT], m0s[KM];
          const MuSums sums_s = mu_posterior_seg<KM, true>(seg_star, Ss, k, lam, sig_of, Qs, m0s, wk);
          GaussQ<KM> gs;
          gs.setup(Qs, m0s);
          if (replay) {
#pragma unroll
            for (int i = 0; i < KM; ++i) mu_star[i] = (i < k) ? rd[DBN_RD_MU_TAU + i] : 0.0;
          } else {
            double z[KM];
#pragma unroll
            for (int i = 0; i < KM; i += 2) {
              double z0, z1;
              rng.normal2(z0, z1);
              z[i] = (i < k) ? z0 : 0.0;
              if (i + 1 < KM) z[i + 1] = (i + 1 < k) ? z1 : 0.0;
            }
            gs.sample(z, mu_star);
          }
          const double ld_star = gs.logpdf(mu_star, Qs, k);
          ml_star = ml_from_sums<KM>(sums_s, Qs, m0s, sig2, lam, mu_star, a.c0, a.T_half_plus_a, a.two_b, wk);
          const double lp_star = -0.5 * ((double)k * 1.8378770664093454836 + sqnorm_k<KM>(mu_star));
          const double lp_cur = -0.5 * ((double)k * 1.8378770664093454836 + sqnorm_k<KM>(mu));
          log_ratio = ml_star - ml_cur + lprior + lp_star - lp_cur + ld_cur - ld_star + log_hr;  // moves.py:112-117
          if (TRACE && td) {
            td[DBN_TD_MUS_LOGDENS + 2] = ld_cur;
            td[DBN_TD_MUS_LOGDENS + 3] = ld_star;
            for (int i = 0; i < KM; ++i) {
              td[DBN_TD_MUS_MEAN + 2 * KMAX + i] = (i < k) ? gc.mean[i] : NAN;
              td[DBN_TD_MUS_MEAN + 3 * KMAX + i] = (i < k) ? gs.mean[i] : NAN;
            }
            double cv[KT];
            invert_spd<KM>(Qc, 1.0, cv);
            for (int i = 0; i < KM; ++i)
              for (int j = 0; j <= i; ++j) td[DBN_TD_MUS_COV + 2 * DBN_KTRI + DBN_TRI(i, j)] = (i < k) ? cv[DBN_TRI(i, j)] : NAN;
            invert_spd<KM>(Qs, 1.0, cv);
            for (int i = 0; i < KM; ++i)
              for (int j = 0; j <= i; ++j) td[DBN_TD_MUS_COV + 3 * DBN_KTRI + DBN_TRI(i, j)] = (i < k) ? cv[DBN_TRI(i, j)] : NAN;
          }
        } else {
          ml_star = score(seg_star, Ss, k, mu_star);  // moves.py:209-216
          log_ratio = ml_star - ml_cur + lprior + log_hr;  // moves.py:225
        }
        const double uacc = replay ? rd[DBN_RD_U_TAU] : rng.uniform();
        const bool acc = GLOB ? (uacc < exp(fmin(1.0, log_ratio))) : (uacc < fmin(1.0, exp(log_ratio)));
        if (TRACE && td) {
          td[DBN_TD_LOGML + 2] = ml_cur;
          td[DBN_TD_LOGML + 3] = ml_star;
        }
        if (TRACE && ti) ti[DBN_TI_TAU_ACCEPT] = acc;
        if (acc) {
          S = Ss;
          for (int j = 0; j < Ss; ++j) TAU(j) = TAS(j);
          fill_rows();   // the cached rows follow the accepted changepoints
          if (GLOB) {
#pragma unroll
            for (int i = 0; i < KM; ++i) mu[i] = mu_star[i];
          }
        }
      }
      // ---- counts: tau_vector[it] is kept iff it >= burn_in and (it - burn_in) % thin == 0 (network.py:388-389)
      if (it >= a.burn_in && ((it - a.burn_in) % a.thin) == 0) {
        atomicAdd(&a.cp_kept[node], 1ull);
        for (int j = 0; j < S - 1; ++j) atomicAdd(&a.cp_hist[(long long)node * (N + 3) + TAU(j)], 1ull);
      }
    }
    if (TRACE && ti) {
      ti[DBN_TI_S] = S;
      for (int j = 0; j < SMAX; ++j) ti[DBN_TI_TAU + j] = (j < S) ? TAU(j) : -1;
    }
    if (TRACE && td) {
      if (!SEQ) td[DBN_TD_DEL_SHAPE] = td[DBN_TD_DEL_RATE] = td[DBN_TD_DELTA2] = NAN;
      for (int i = 0; i < KMAX; ++i) td[DBN_TD_MU + i] = (GLOB && i < k && i < KM) ? mu[i < KM ? i : 0] : NAN;
    }
  }

  // ---- store state
  if (active) {
  a.st_npi[u] = npi;
#pragma unroll
  for (int j = 0; j < PMAX; ++j) a.st_pi[j * a.U + u] = pi[j];
  a.st_S[u] = S;
  for (int j = 0; j < SMAX; ++j) a.st_tau[j * a.U + u] = TAU(j);
  a.st_sig[u] = sig2;
  a.st_lam[u] = lam;
  a.st_del[u] = dlt;
  if (GLOB) {
#pragma unroll
    for (int i = 0; i < KM; ++i) a.st_mu[i * a.U + u] = mu[i];
  }
  }
  // ---- work counters: summed per block through shared memory (no warp shuffles: see profiles/r1_notes.md), one
  // global atomic per counter and block
  const unsigned long long v[5] = {(unsigned long long)n_it, wk.evals, wk.segs, wk.bytes, wk.flops};
  unsigned long long* const s_cnt = reinterpret_cast<unsigned long long*>(s_tau);  // 5 x CHAIN_BLOCK x 8 B <= s_tau
  static_assert(5 * CHAIN_BLOCK * sizeof(unsigned long long) <= SMAX * CHAIN_BLOCK * sizeof(int), "counter scratch fits s_tau");
  __syncthreads();
#pragma unroll
  for (int c = 0; c < 5; ++c) s_cnt[c * CHAIN_BLOCK + tid] = v[c];
  __syncthreads();
  if (tid < 5) {
    unsigned long long t = 0;
    for (int i = 0; i < CHAIN_BLOCK; ++i) t += s_cnt[tid * CHAIN_BLOCK + i];
    if (t) atomicAdd(&a.counters[tid], t);
  }
}

}  // namespace dbn
