// Kernel 4 — full-parents globally coupled chain step (fp_glob_coup_nh_dbn): the regressor loop of
// fpGlobCoupBpwLinReg.py:68-110 for design width k = n (every other gene is a parent, :27-31), any n.
//
// One thread block per (node, chain) unit; the k x k matrices live in shared memory when they fit (k <= ~56) and in
// a per-block global workspace otherwise.  Per segment the block forms A_h = I + lam G_h from two prefix-table
// rows, factorises it (Cholesky, A = L L^T) and inverts the factor (W = L^-1), after which everything the iteration
// needs is a matrix-vector product:
//   quadratic form   r^T C^-1 r = rho - lam |W v|^2,  v = g - G mu           (samplers.py:103-127)
//   posterior mean   A^-1 (mu + lam g) = W^T W (mu + lam g),  cov = s2 lam W^T W  (samplers.py:189-218)
//   draw             beta = mean + sqrt(s2 lam) W^T z
// The changepoint move (moves.py:25-126) needs, for tau and tau*, the marginal likelihood at a mean vector and the
// Gaussian muSampler draws from (samplers.py:307-351).  Both follow from five sums over the segments,
//   M = sum_h (I - A_h^-1),  a = sum_h A_h^-1 g_h,  c = sum_h g_h^T A_h^-1 g_h,  s = sum_h y_h^T y_h,
//   ld = sum_h ln det A_h,
// because G_h A_h^-1 = (I - A_h^-1) / lam:
//   precision  Q = I + M / (lam s2),   mean numerator  a / s2 (+ mu_in)         (samplers.py:319-337)
//   sum_h r_h^T C_h^-1 r_h at mean mu  =  s - lam c - 2 mu.a + mu^T M mu / lam   (marginalLikelihood.py:91-150)
// so one sweep per changepoint list scores ANY mu (the proposed mu* is only known after the sweep).
#pragma once
#include "dbn_chain.cuh"
#include "dbn_fp_big.cuh"

namespace dbn {

constexpr int FP_BLOCK_MAX = 512;
// threads per block: one warp while a matrix row fits it (k <= 32: the block barriers are then warp-local and the idle
// warps of a wider block no longer wait in them — 54 % of all stall samples at k = 11, profiles/r1_v9), else 64 / 128
// the large-k path (matrices in the global workspace, tiled routines of dbn_fp_big.cuh) always runs FP_BIG_BLOCK threads, one block per SM
constexpr int FP_BIG_BLOCK = FPB_THREADS;
inline int fp_block_threads(int k, bool mats_in_smem) { return !mats_in_smem ? FP_BIG_BLOCK : (k <= 32 ? 32 : (k <= 64 ? 64 : 128)); }
#define FP_BLOCK ((int)blockDim.x)
constexpr int FP_NMAT = 5;  // A/L, W, M (tau), M (tau*), Q
constexpr int FP_NVEC = 12;

struct FpArgs {
  Table tb;
  int num_samples, burn_in, thin, with_tau;
  int C, U, it0, n_iter;
  int unit0;       // grid index node * C + chain of local unit 0
  unsigned int chain_offset;
  unsigned long long seed;
  double T_sigma_half, a_sig, b_sig, a_lam, b_lam, log_p, log_1mp, c0, T_half_plus_a, two_b;
  int* st_S;
  int* st_tau;     // [SMAX][U]
  double* st_sig;
  double* st_lam;
  double* st_mu;   // [k][U]
  double* st_del;  // delta^2 (sequentially coupled)
  double a_del, b_del;
  unsigned long long* pos;      // [n][n]  #kept mu samples with mu_j >= 0   (credible_score, scores.py:196-205)
  unsigned long long* kept;     // [n]     #kept mu samples
  unsigned long long* cp_hist;  // [n][N+3]
  unsigned long long* cp_kept;  // [n]
  unsigned long long* neg;      // [n][n]  #kept mu samples with mu_j <= 0
  unsigned long long* counters;
  double* ws;      // per-block workspace of FP_NMAT * ld * k doubles (null: matrices in shared memory)
  double* bsc;     // plain full-parents methods: per-block [SMAX][2][k] posterior mean and direction of every beta_h
  const double* rd;
  const int* ri;
  double* td;
  int* ti;
};

__host__ __device__ inline int fp_ld(int k) { return k | 1; }
// replay / trace record strides (doubles / ints per (unit, iteration)); see include/dbn_b200.h
__host__ __device__ inline int fp_rd_stride(int k) { return 3 + k + DBN_SMAX * k + DBN_SMAX; }   // ... , sigma^2_h[SMAX] (vv)
__host__ __device__ inline int fp_td_stride(int k) { return 10 + 3 * k + 3 * DBN_SMAX * k + 3 * DBN_SMAX; }   // ... , sigma^2_h shape / rate / draw (vv)

// ---- block-level building blocks (all threads of the block call them) -----------------------------------------------
// sum of x over the block, the same value (same order of additions) in every thread.  Through shared memory and block
// barriers only: warp shuffles are emitted without a warp barrier where ptxas considers the warp convergent, which is
// not safe after divergent loops (profiles/r1_notes.md).
__device__ __forceinline__ double fp_block_sum(double x, double* red) {
  __syncthreads();  // red is free (every thread has read the previous sum)
  red[threadIdx.x] = x;
  __syncthreads();
  double t = 0.0;
  if (FP_BLOCK <= 128) {
    for (int w = 0; w < FP_BLOCK; ++w) t += red[w];
    return t;
  }
  // wide blocks (large k): 32 partial sums of FP_BLOCK / 32 consecutive entries, then every thread adds the 32
  double* red2 = red + FP_BLOCK_MAX;
  if (threadIdx.x < 32) {
    const int g = FP_BLOCK >> 5;
    double p = 0.0;
    for (int w = 0; w < g; ++w) p += red[threadIdx.x * g + w];
    red2[threadIdx.x] = p;
  }
  __syncthreads();
#pragma unroll
  for (int w = 0; w < 32; ++w) t += red2[w];
  return t;
}

__device__ __forceinline__ double fp_dot(const double* a, const double* b, int k, double* red) {
  double x = 0.0;
  for (int i = threadIdx.x; i < k; i += FP_BLOCK) x += a[i] * b[i];
  return fp_block_sum(x, red);
}

// design column a of node `node` -> table column (0 = intercept, gene j -> j + 1); columns are sorted by gene id
__device__ __forceinline__ int fp_col(int a, int node) { return a == 0 ? 0 : (a - 1 < node ? a : a + 1); }

// A = I + lam (P[hi] - P[lo]) restricted to the node's design columns (full symmetric), g, s
__device__ inline void fp_load_segment(const Table& tb, int node, int lo, int hi, double lam, double* A, int ld, int k, double* g,
                                double* s_out) {
  const double* __restrict__ ph = tb.base + (long long)hi * tb.row_stride;
  const double* __restrict__ pl = tb.base + (long long)lo * tb.row_stride;
  const int nb = tb.zy_base + node * (tb.n + 2);
  if (FP_BLOCK > 128) {
    // large k: a warp per matrix row, lanes along the row (table entries of a row are contiguous but for the node's own column);
    // sixteen table loads in flight per lane before the first store (one pair per step left the loop waiting for DRAM:
    // 15 % of all stall samples of the c5 launch, profiles/r2_c5/c5_fp_glob_stall_lines.txt)
    for (int a = threadIdx.x >> 5; a < k; a += FP_BLOCK >> 5) {
      const int fa = fp_col(a, node);
      const long long ro = (long long)fa * (fa + 1) / 2;
      for (int b0 = threadIdx.x & 31; b0 <= a; b0 += 256) {
        double vh[8], vl[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int b = b0 + 32 * u;
          const long long off = ro + fp_col(b <= a ? b : a, node);
          vh[u] = __ldg(ph + off);
          vl[u] = __ldg(pl + off);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int b = b0 + 32 * u;
          if (b <= a) {
            const double v = lam * (vh[u] - vl[u]) + (a == b ? 1.0 : 0.0);
            A[(long long)a * ld + b] = v;
            A[(long long)b * ld + a] = v;   // (dropping the mirrored, scattered store was measured: no gain, profiles/r2_c5 notes)
          }
        }
      }
    }
  } else {
  for (int e = threadIdx.x; e < k * (k + 1) / 2; e += FP_BLOCK) {
    int a = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);  // row of packed entry e; corrected for rounding below
    while (a * (a + 1) / 2 > e) --a;
    while ((a + 1) * (a + 2) / 2 <= e) ++a;
    const int b = e - a * (a + 1) / 2;
    const int fa = fp_col(a, node), fb = fp_col(b, node);
    const int off = fa * (fa + 1) / 2 + fb;
    const double v = lam * (__ldg(ph + off) - __ldg(pl + off)) + (a == b ? 1.0 : 0.0);
    A[a * ld + b] = v;
    A[b * ld + a] = v;
  }
  }
  for (int a = threadIdx.x; a < k; a += FP_BLOCK) {
    const int fa = fp_col(a, node);
    g[a] = __ldg(ph + nb + fa) - __ldg(pl + nb + fa);
  }
  if (threadIdx.x == 0) *s_out = __ldg(ph + nb + tb.n + 1) - __ldg(pl + nb + tb.n + 1);
  __syncthreads();
}

// y = M x for a full symmetric (or any square) k x k matrix; thread per output row, x broadcast from shared memory
__device__ inline void fp_symv(const double* M, int ld, int k, const double* x, double* y) {
  for (int i = threadIdx.x; i < k; i += FP_BLOCK) {
    double acc = 0.0;
    int j = 0;
    if (FP_BLOCK > 128) {   // large k (matrix in global memory): eight loads in flight
      double s4[4] = {0.0, 0.0, 0.0, 0.0};
      for (; j + 8 <= k; j += 8) {
        double v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = M[(long long)(j + q) * ld + i];
#pragma unroll
        for (int q = 0; q < 8; ++q) s4[q & 3] = fma(v[q], x[j + q], s4[q & 3]);
      }
      acc = (s4[0] + s4[1]) + (s4[2] + s4[3]);
    }
    for (; j < k; ++j) acc += M[j * ld + i] * x[j];  // column i of a symmetric matrix: coalesced over i
    y[i] = acc;
  }
  __syncthreads();
}

// in-place lower Cholesky of the full symmetric matrix A (upper part left untouched); returns ln det A
__device__ inline double fp_cholesky(double* A, int ld, int k, double* idiag, double* stage) {  // idiag[j] = 1 / L[j][j]
  if (stage) return fpb_factor(A, nullptr, ld, k, idiag, stage);
  double mant = 1.0;  // product of the pivots l_j = sqrt(d_j) as mantissa x 2^expo: one log per factorisation
  int expo = 0;
  for (int j = 0; j < k; ++j) {
    __syncthreads();  // the previous column's trailing update is complete
    const double l = sqrt(A[j * ld + j]);
    {
      int ex;
      mant = frexp(mant * l, &ex);
      expo += ex;
    }
    __syncthreads();  // every thread has read the pivot before it is overwritten
    const double il = 1.0 / l;
    if (threadIdx.x == 0) {
      A[j * ld + j] = l;
      idiag[j] = il;
    }
    for (int i = j + 1 + threadIdx.x; i < k; i += FP_BLOCK) A[i * ld + j] *= il;
    __syncthreads();
    // trailing update, thread per column c: A[i][c] -= A[i][j] A[c][j] for i >= c
    for (int c = j + 1 + threadIdx.x; c < k; c += FP_BLOCK) {
      const double lcj = A[c * ld + j];
      for (int i = c; i < k; ++i) A[i * ld + c] -= A[i * ld + j] * lcj;
    }
  }
  __syncthreads();
  return 2.0 * (log(mant) + (double)expo * 0.69314718055994530942);
}

// W = L^-1 (lower triangular, zeros above the diagonal): forward substitution on all columns at once
__device__ inline void fp_trinv(const double* L, double* W, int ld, int k, const double* idiag, double* stage) {
  for (int c = threadIdx.x; c < k; c += FP_BLOCK) {
    for (int i = 0; i < c; ++i) W[i * ld + c] = 0.0;
    for (int i = c; i < k; ++i) {
      double acc = (i == c) ? 1.0 : 0.0;
      for (int j = c; j < i; ++j) acc -= L[i * ld + j] * W[j * ld + c];
      W[i * ld + c] = acc * idiag[i];
    }
  }
  __syncthreads();
}

// Cholesky factor and inverse factor W = L^-1 in one call (large k: fused per panel, dbn_fp_big.cuh); returns ln det A
__device__ inline double fp_factor_inv(double* A, double* W, int ld, int k, double* idiag, double* stage) {
  if (stage) return fpb_factor(A, W, ld, k, idiag, stage);
  const double ldet = fp_cholesky(A, ld, k, idiag, nullptr);
  fp_trinv(A, W, ld, k, idiag, nullptr);
  return ldet;
}

// t = W x (lower triangular W), thread per row
__device__ inline void fp_trmv(const double* W, int ld, int k, const double* x, double* t, double* stage) {
  if (stage) return fpb_trmv(W, ld, k, x, t, stage);
  for (int i = threadIdx.x; i < k; i += FP_BLOCK) {
    double acc = 0.0;
    for (int j = 0; j <= i; ++j) acc += W[i * ld + j] * x[j];
    t[i] = acc;
  }
  __syncthreads();
}

// y = W^T t (lower triangular W), thread per output entry: coalesced over a
__device__ inline void fp_trmv_t(const double* W, int ld, int k, const double* t, double* y, double* stage, double* red) {
  if (stage) return fpb_trmv_t(W, ld, k, t, y);
  for (int a = threadIdx.x; a < k; a += FP_BLOCK) {
    double acc = 0.0;
    for (int i = a; i < k; ++i) acc += W[i * ld + a] * t[i];
    y[a] = acc;
  }
  __syncthreads();
}

// M += I - W^T W (full symmetric)
__device__ inline void fp_accumulate_i_minus_ainv(const double* W, double* M, int ld, int k, double* stage, double scale) {
  if (stage) return fpb_syrk(W, M, ld, k, stage, scale);
  for (int a = 0; a < k; ++a) {
    for (int b = threadIdx.x; b <= a; b += FP_BLOCK) {
      double acc = 0.0;
      for (int i = a; i < k; ++i) acc += W[i * ld + a] * W[i * ld + b];
      const double v = scale * (((a == b) ? 1.0 : 0.0) - acc);
      M[a * ld + b] += v;
      if (a != b) M[b * ld + a] += v;
    }
  }
  __syncthreads();
}


// z[0..k) ~ N(0,1), Philox slots base + i/2
__device__ inline void fp_normals(double* z, int k, uint32_t chain, uint32_t node, uint32_t it, uint32_t base, uint2 key) {
  for (int p = threadIdx.x; 2 * p < k; p += FP_BLOCK) {
    const double2 zz = normal_pair(chain, node, it, base + (uint32_t)p, key);
    z[2 * p] = zz.x;
    if (2 * p + 1 < k) z[2 * p + 1] = zz.y;
  }
  __syncthreads();
}

// the five sums of one changepoint list at lam (see the header comment); M must be zeroed by the caller
struct FpSummary {
  double c, s, ld;
};

template <class TauFn>
__device__ FpSummary fp_sweep(const Table& tb, int node, TauFn tau, int S, double lam, double* A, double* W, double* M, int ld, int k,
                              double* g, double* t, double* x, double* a_sum, double* red, double* s_tmp, double* stage,
                              const double* wgt = nullptr) {   // wgt[h] = sigma^2_h: segment h enters M and a with 1 / sigma^2_h (vvMuSampler)
  FpSummary r = {0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < k; i += FP_BLOCK) a_sum[i] = 0.0;
  for (int i = threadIdx.x; i < k * ld; i += FP_BLOCK) M[i] = 0.0;
  __syncthreads();
  int lo = 0;
  for (int h = 0; h < S; ++h) {
    const int hi = DBN_SEG_END(tau(h), h, S, tb.N);
    fp_load_segment(tb, node, lo, hi, lam, A, ld, k, g, s_tmp);
    r.s += *s_tmp;
    r.ld += fp_factor_inv(A, W, ld, k, t, stage);
    fp_trmv(W, ld, k, g, t, stage);          // t = L^-1 g
    r.c += fp_dot(t, t, k, red);             // g^T A^-1 g
    fp_trmv_t(W, ld, k, t, x, stage, red);   // x = A^-1 g
    const double wh = wgt ? 1.0 / wgt[h] : 1.0;   // wgt holds the variances
    for (int i = threadIdx.x; i < k; i += FP_BLOCK) a_sum[i] += wh * x[i];
    fp_accumulate_i_minus_ainv(W, M, ld, k, stage, wh);
    lo = hi;
  }
  return r;
}

// does the changepoint list `tau` (S segments) contain the segment (lo, hi]?
template <class TauFn>
__device__ __forceinline__ bool fp_has_segment(TauFn tau, int S, int N, int lo, int hi) {
  int l = 0;
  for (int h = 0; h < S; ++h) {
    const int e = DBN_SEG_END(tau(h), h, S, N);
    if (l == lo && e == hi) return true;
    l = e;
  }
  return false;
}

// Delta form of fp_sweep for the changepoint move: a birth / death / reallocation changes at most two segments, and the
// five sums are sums over segments, so the segments `tau` shares with `other` are evaluated once for both lists.
// Segments of `tau` that also occur in `other` add to set 0 (M0, a0, r0), the others to set 1; with skip_common the shared
// segments are not evaluated at all (second pass: only the proposal's new segments).  Returns the number of segments
// evaluated.  The caller zeroes what it wants zeroed.
template <class TauFn, class OtherFn>
__device__ int fp_sweep_split(const Table& tb, int node, TauFn tau, int S, OtherFn other, int So, bool skip_common, double lam, double* A,
                              double* W, int ld, int k, double* M0, double* a0, FpSummary& r0, double* M1, double* a1, FpSummary& r1,
                              double* g, double* t, double* x, double* red, double* s_tmp, double* stage) {
  int lo = 0, n_eval = 0;
  for (int h = 0; h < S; ++h) {
    const int hi = DBN_SEG_END(tau(h), h, S, tb.N);
    const bool common = fp_has_segment(other, So, tb.N, lo, hi);
    if (!(common && skip_common)) {
      FpSummary& r = common ? r0 : r1;
      double* a_sum = common ? a0 : a1;
      fp_load_segment(tb, node, lo, hi, lam, A, ld, k, g, s_tmp);
      r.s += *s_tmp;
      r.ld += fp_factor_inv(A, W, ld, k, t, stage);
      fp_trmv(W, ld, k, g, t, stage);          // t = L^-1 g
      r.c += fp_dot(t, t, k, red);             // g^T A^-1 g
      fp_trmv_t(W, ld, k, t, x, stage, red);   // x = A^-1 g
      for (int i = threadIdx.x; i < k; i += FP_BLOCK) a_sum[i] += x[i];
      fp_accumulate_i_minus_ainv(W, common ? M0 : M1, ld, k, stage, 1.0);
      ++n_eval;
    }
    lo = hi;
  }
  return n_eval;
}

// t = L^-1 g by forward substitution on the Cholesky factor (idiag[j] = 1 / L[j][j]); column-oriented, k block steps
__device__ inline void fp_trsv(const double* L, int ld, int k, const double* idiag, const double* g, double* t, bool upper) {
  for (int i = threadIdx.x; i < k; i += FP_BLOCK) t[i] = g[i];
  __syncthreads();
  for (int j = 0; j < k; ++j) {
    if (threadIdx.x == 0) t[j] *= idiag[j];
    __syncthreads();
    const double tj = t[j];
    for (int i = j + 1 + threadIdx.x; i < k; i += FP_BLOCK) t[i] -= (upper ? L[(long long)j * ld + i] : L[i * ld + j]) * tj;
    __syncthreads();
  }
}

// plain (mu = 0) marginal likelihood of one changepoint list (marginalLikelihood.py:91-150): only
// s = sum y_h^T y_h, c = sum g_h^T A_h^-1 g_h and ld = sum ln det A_h are needed, so no inverse is formed
template <class TauFn>
__device__ FpSummary fp_sweep_plain(const Table& tb, int node, TauFn tau, int S, double lam, double* A, int ld, int k, double* g,
                                    double* t, double* idiag, double* red, double* s_tmp, double* stage) {
  FpSummary r = {0.0, 0.0, 0.0};
  int lo = 0;
  for (int h = 0; h < S; ++h) {
    const int hi = DBN_SEG_END(tau(h), h, S, tb.N);
    fp_load_segment(tb, node, lo, hi, lam, A, ld, k, g, s_tmp);
    r.s += *s_tmp;
    r.ld += fp_cholesky(A, ld, k, idiag, stage);
    fp_trsv(A, ld, k, idiag, g, t, stage != nullptr);
    r.c += fp_dot(t, t, k, red);
    lo = hi;
  }
  return r;
}

// x = A^-1 rhs through the inverse factor (t is scratch)
__device__ __forceinline__ void fp_apply_ainv(const double* W, int ld, int k, const double* rhs, double* t, double* x, double* stage,
                                              double* red) {
  fp_trmv(W, ld, k, rhs, t, stage);
  fp_trmv_t(W, ld, k, t, x, stage, red);
}

// sequentially coupled log marginal likelihood (marginalLikelihood.py:91-150 with method == 'seq-coup', prior means from
// betaTildeSampler, samplers.py:4-54): bt[0] = 0, bt[1] = A_lam(G_1)^-1 lam g_1, bt[h+1] = A_dlt(G_h)^-1 (bt[h-1] + dlt g_h).
// bt_cur / bt_prev / nxt are k-vectors of scratch.
template <class TauFn>
__device__ inline double fp_score_seq(const Table& tb, int node, TauFn tau, int S, double lam, double dlt, double* bt_cur, double* bt_prev,
                               double* nxt, double* A, double* W, int ld, int k, double* g, double* t, double* x, double* tmp,
                               double* rhs, double* red, double* s_tmp, double* stage, double c0, double T_half_plus_a, double two_b) {
  for (int i = threadIdx.x; i < k; i += FP_BLOCK) bt_cur[i] = bt_prev[i] = 0.0;
  __syncthreads();
  double acc_ld = 0.0, acc_q = 0.0;
  int lo = 0;
  for (int h = 0; h < S; ++h) {
    const int hi = DBN_SEG_END(tau(h), h, S, tb.N);
    if (h == 1) {
      fp_load_segment(tb, node, lo, hi, lam, A, ld, k, g, s_tmp);
      fp_factor_inv(A, W, ld, k, t, stage);
      for (int i = threadIdx.x; i < k; i += FP_BLOCK) rhs[i] = lam * g[i];
      __syncthreads();
      fp_apply_ainv(W, ld, k, rhs, t, bt_cur, stage, red);
    }
    const double w = (h == 0) ? lam : dlt;
    fp_load_segment(tb, node, lo, hi, w, A, ld, k, g, s_tmp);
    const double s = *s_tmp;
    fp_symv(A, ld, k, bt_cur, tmp);
    for (int i = threadIdx.x; i < k; i += FP_BLOCK) {
      const double gm = (tmp[i] - bt_cur[i]) / w;
      tmp[i] = gm;
      x[i] = g[i] - gm;
    }
    __syncthreads();
    const double rho = s - 2.0 * fp_dot(bt_cur, g, k, red) + fp_dot(bt_cur, tmp, k, red);
    acc_ld += fp_factor_inv(A, W, ld, k, t, stage);
    fp_trmv(W, ld, k, x, t, stage);
    acc_q += rho - w * fp_dot(t, t, k, red);
    if (h >= 1) {
      for (int i = threadIdx.x; i < k; i += FP_BLOCK) rhs[i] = bt_prev[i] + dlt * g[i];
      __syncthreads();
      fp_apply_ainv(W, ld, k, rhs, t, nxt, stage, red);
      for (int i = threadIdx.x; i < k; i += FP_BLOCK) {
        bt_prev[i] = bt_cur[i];
        bt_cur[i] = nxt[i];
      }
      __syncthreads();
    }
    lo = hi;
  }
  return c0 - 0.5 * acc_ld - T_half_plus_a * log(two_b + acc_q);
}

// vvLogMargLikelihood (marginalLikelihood.py:5-43) at prior mean mu: per segment its own length T_h in the Gamma / pi /
// exponent terms; needs the factor and one forward substitution per segment, no inverse.  c_seg = a ln(2b) - lgamma(a).
template <class TauFn>
__device__ inline double fp_score_vv(const Table& tb, int node, TauFn tau, int S, double lam, const double* mu, double* A, int ld, int k,
                              double* g, double* t, double* x, double* tmp, double* idiag, double* red, double* s_tmp, double* stage,
                              double a_sig, double c_seg, double two_b) {
  double acc = 0.0;
  int lo = 0;
  for (int h = 0; h < S; ++h) {
    const int hi = DBN_SEG_END(tau(h), h, S, tb.N);
    fp_load_segment(tb, node, lo, hi, lam, A, ld, k, g, s_tmp);
    const double s = *s_tmp;
    fp_symv(A, ld, k, mu, tmp);
    for (int i = threadIdx.x; i < k; i += FP_BLOCK) {
      const double gm = (tmp[i] - mu[i]) / lam;   // (G mu)_i
      tmp[i] = gm;
      x[i] = g[i] - gm;
    }
    __syncthreads();
    const double rho = s - 2.0 * fp_dot(mu, g, k, red) + fp_dot(mu, tmp, k, red);
    const double ldh = fp_cholesky(A, ld, k, idiag, stage);
    fp_trsv(A, ld, k, idiag, x, t, stage != nullptr);
    const double q = rho - lam * fp_dot(t, t, k, red);
    const double Th = 0.5 * (double)(hi - lo);
    acc += lgamma(Th + a_sig) + c_seg - Th * 1.1447298858494002 - 0.5 * ldh - (Th + a_sig) * log(two_b + q);
    lo = hi;
  }
  return acc;
}

// sum_h r_h^T C_h^-1 r_h at mean mu from the sums
__device__ inline double fp_quad_at(const FpSummary& sm, const double* M, int ld, int k, const double* a_sum, const double* mu, double lam,
                             double* tmp, double* red) {
  fp_symv(M, ld, k, mu, tmp);
  const double mMm = fp_dot(mu, tmp, k, red);
  const double ma = fp_dot(mu, a_sum, k, red);
  return sm.s - lam * sm.c - 2.0 * ma + mMm / lam;
}

// GLOBM: the globally coupled model (fpGlobCoupBpwLinReg.py).  !GLOBM: the plain full-parents models, prior mean 0 —
// fp_varying_nh_dbn (fullParentsBpwLinReg.py:41-140: nh changepoint move of moves.py:129-236) and, with with_tau == 0,
// fp_h_dbn (fpBayesianLinearRegression.py:40-125: Gibbs draws only); their edge scores come from the thinned beta chain.
// FPM_VV: fp_var_glob_coup_nh_dbn (fpvvGlobCoup.py:72-121): segment-specific variances (segmentSigmaSampler), mu redrawn every
// iteration from vvMuSampler (always accepted), the changepoint move of vvGlobCoupTauMove on the OLD mu scored by
// vvLogMargLikelihood (per-segment T_h); edge scores from the beta chain like the plain models.
// FPM_SEQ: fp_seq_coup_nh_dbn (fpSeqCoupBpwlinReg.py:68-106): the sequentially coupled loop at width n — prior mean of segment h is
// betaTildeSampler's beta-tilde_h (samplers.py:4-54), delta^2 replaces lambda^2 for h > 0, delta^2 Gibbs draw, nh changepoint move
// scored by the sequentially coupled marginal likelihood.  The `mu` / `mu_star` vectors hold beta-tilde_h / beta-tilde_{h-1}.
constexpr int FPM_GLOB = 0, FPM_PLAIN = 1, FPM_VV = 2, FPM_SEQ = 3;
// BIG: the matrices live in the global workspace and the blocked routines run (256 threads, 2 blocks per SM); !BIG: matrices in
// shared memory, at most 128 threads and 96 registers so that 19+ one-warp blocks of a small network are resident per SM.
template <bool TRACE, int FPM = FPM_GLOB, bool BIG = false>
__global__ void __launch_bounds__(BIG ? FP_BIG_BLOCK : 128, BIG ? 1 : 5) fp_glob_kernel(const FpArgs a) {
  constexpr bool GLOBM = FPM == FPM_GLOB, VVM = FPM == FPM_VV, SEQM = FPM == FPM_SEQ;
  constexpr bool GHR = GLOBM || VVM;   // Hastings ratios and acceptance rule of the globally coupled changepoint moves
  extern __shared__ double sm[];
  const int n = a.tb.n, N = a.tb.N, k = n, ld = BIG ? fp_ld_big(k) : fp_ld(k);
  const int tid = threadIdx.x;
  // shared layout: vectors | reduction scratch | ints | (matrices, when they fit)
  double* vec = sm;
  double* g = vec + 0 * k, *mu = vec + 1 * k, *mu_star = vec + 2 * k, *t = vec + 3 * k, *x = vec + 4 * k, *z = vec + 5 * k;
  double* mean = vec + 6 * k, *a_cur = vec + 7 * k, *a_star = vec + 8 * k, *tmp = vec + 9 * k, *rhs = vec + 10 * k, *e = vec + 11 * k;
  double* red = vec + FP_NVEC * k;       // [FP_BLOCK_MAX + 32]
  double* sc = red + FP_BLOCK_MAX + 32;                  // scalars published by thread 0: [0] sig2 [1] lam [2] u [3] s_tmp ...
  double* sigv = sc + 8;                 // [SMAX] sigma^2_h of the iteration (vv)
  int* s_tau = reinterpret_cast<int*>(sc + 8 + SMAX);
  int* s_tas = s_tau + SMAX;
  int* s_int = s_tas + SMAX;             // [8]
  double* mats = BIG ? a.ws + (long long)blockIdx.x * FP_NMAT * ld * k : reinterpret_cast<double*>(s_int + 8);
  // large k: staging area of the tiled routines (16-byte aligned: cp.async / 128-bit shared loads)
  double* stage = BIG ? reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(s_int + 8) + 15) & ~(uintptr_t)15) : nullptr;
  double* A = mats, *W = mats + (long long)ld * k, *Mc = mats + 2ll * ld * k, *Ms = mats + 3ll * ld * k, *Q = mats + 4ll * ld * k;
  double* bsc = a.bsc ? a.bsc + (long long)blockIdx.x * (2 * SMAX) * k : nullptr;
  const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
  const bool replay = TRACE && a.rd != nullptr;
  const int rds = fp_rd_stride(k), tds = fp_td_stride(k);
  auto tau_cur = [&](int j) { return s_tau[j]; };
  auto tau_star = [&](int j) { return s_tas[j]; };

  for (int u = blockIdx.x; u < a.U; u += gridDim.x) {
    const int node = (a.unit0 + u) / a.C, chain = (a.unit0 + u) - node * a.C;
    const uint32_t gchain = a.chain_offset + (unsigned)chain, unode = (unsigned)node;
    __syncthreads();
    int S = a.st_S[u];
    if (tid < SMAX) s_tau[tid] = a.st_tau[tid * a.U + u];
    double sig2 = a.st_sig[u], lam = a.st_lam[u];
    double dlt = SEQM ? a.st_del[u] : 1.0;
    for (int i = tid; i < k; i += FP_BLOCK) mu[i] = a.st_mu[(long long)i * a.U + u];
    __syncthreads();
    unsigned long long n_evals = 0, n_segs = 0;
    unsigned long long n_inv = 0, n_qfac = 0;   // factorisations that also form the inverse and I - A^-1; factorisations of the muSampler precision

    for (int li = 0; li < a.n_iter; ++li) {
      const int it = a.it0 + li;
      const uint32_t uit = (uint32_t)it;
      const double* rd = replay ? a.rd + ((long long)u * a.n_iter + li) * rds : nullptr;
      const int* ri = replay ? a.ri + ((long long)u * a.n_iter + li) * 4 : nullptr;
      double* td = (TRACE && a.td) ? a.td + ((long long)u * a.n_iter + li) * tds : nullptr;
      int* ti = (TRACE && a.ti) ? a.ti + ((long long)u * a.n_iter + li) * 16 : nullptr;
      if (TRACE && ti && tid == 0) ti[3] = S;
      // the initial zero vector mu_vector[0] is part of the thinned chain when burn_in == 0 (network.py:309-311)
      if (GLOBM && it == 0 && a.burn_in == 0) {
        if (tid == 0) atomicAdd(&a.kept[node], 1ull);
        for (int j = 1 + tid; j < k; j += FP_BLOCK) {
          const int gene = fp_col(j, node) - 1;
          atomicAdd(&a.pos[(long long)node * n + gene], 1ull);
          atomicAdd(&a.neg[(long long)node * n + gene], 1ull);
        }
      }

      // ======================= Gibbs pass: sigma^2, beta_h, lambda^2 (fpGlobCoupBpwLinReg.py:70-86) ================
      double q_sum = 0.0, acc_dd = 0.0, acc_de = 0.0, acc_ee = 0.0;
      double acc_ss = 0.0;   // vv: sum_h |beta_h - mu|^2 / sigma^2_h (samplers.py:275-282)
      auto w_of = [&](int h) -> double { return (SEQM && h > 0) ? dlt : lam; };   // weight of segment h (samplers.py:158-163)
      if constexpr (SEQM) {
        for (int i = tid; i < k; i += FP_BLOCK) mu[i] = mu_star[i] = 0.0;   // beta-tilde_0 = 0
        __syncthreads();
      }
      {
        int lo = 0;
        for (int h = 0; h < S; ++h) {
          const int hi = DBN_SEG_END(s_tau[h], h, S, N);
          const double w = w_of(h);
          if (SEQM && h == 1) {   // beta-tilde_1 from this segment's own data with lambda^2 (samplers.py:36-39)
            fp_load_segment(a.tb, node, lo, hi, lam, A, ld, k, g, sc + 3);
            fp_factor_inv(A, W, ld, k, t, stage);
            for (int i = tid; i < k; i += FP_BLOCK) rhs[i] = lam * g[i];
            __syncthreads();
            fp_apply_ainv(W, ld, k, rhs, t, mu, stage, red);
            ++n_segs;
          }
          fp_load_segment(a.tb, node, lo, hi, w, A, ld, k, g, sc + 3);
          const double s = sc[3];
          // G mu = (A mu - mu) / lam ; v = g - G mu ; rho = s - 2 mu.g + mu.G mu
          fp_symv(A, ld, k, mu, tmp);
          for (int i = tid; i < k; i += FP_BLOCK) {
            const double gm = (tmp[i] - mu[i]) / w;
            tmp[i] = gm;
            x[i] = g[i] - gm;              // v
            rhs[i] = mu[i] + w * g[i];     // numerator of the posterior mean (samplers.py:210-212)
          }
          __syncthreads();
          const double rho = s - 2.0 * fp_dot(mu, g, k, red) + fp_dot(mu, tmp, k, red);
          // large k (solve form): the Gibbs pass needs A^-1 only applied to vectors, so the factor and the inverted diagonal
          // blocks suffice (blocked two-vector solves, dbn_fp_big.cuh) and the k^3 / 3 of the full inverse factor is saved.
          // Trace builds also form W (the traced covariance diagonal reads it); seq needs it for the beta-tilde recursion.
          const bool solve_form = BIG && !SEQM;
          if (solve_form) {
            fpb_factor(A, W, ld, k, t, stage, TRACE);
            fpb_solve_fwd2(A, W, ld, k, x, rhs, stage);   // x <- L^-1 v, rhs <- L^-1 (mu + w g)
            for (int i = tid; i < k; i += FP_BLOCK) t[i] = x[i];
            __syncthreads();
          } else {
            fp_factor_inv(A, W, ld, k, t, stage);
            fp_trmv(W, ld, k, x, t, stage);
          }
          const double q_h = rho - w * fp_dot(t, t, k, red);
          q_sum += q_h;
          double sig_h = 1.0;
          if constexpr (VVM) {
            // segmentSigmaSampler (samplers.py:90-101): sigmaSqrSampler on segment h alone, numSamples = T = T_h
            if (tid == 0) {
              const double shape = a.a_sig + 0.5 * (double)(hi - lo), rate = a.b_sig + 0.5 * q_h;
              const double s2h = replay ? rd[3 + k + DBN_SMAX * k + h] : rate / gamma_draw(gchain, unode, uit, 20000u + 256u * (unsigned)h, key, shape);
              sigv[h] = s2h;
              if (TRACE && td) {
                const int o = 10 + 3 * k + 3 * DBN_SMAX * k;
                td[o + h] = shape; td[o + DBN_SMAX + h] = rate; td[o + 2 * DBN_SMAX + h] = s2h;
                if (h == 0) { td[0] = shape; td[1] = rate; td[4] = s2h; }
              }
            }
            __syncthreads();
            sig_h = sigv[h];
          }
          if (solve_form) {
            if (!replay) {
              fp_normals(z, k, gchain, unode, uit, (uint32_t)(h * 512), key);
            } else {
              for (int i = tid; i < k; i += FP_BLOCK) z[i] = 0.0;
              __syncthreads();
            }
            fpb_solve_bwd2(A, W, ld, k, rhs, z, stage);   // rhs <- L^-T L^-1 (mu + w g) = posterior mean, z <- L^-T z
            for (int i = tid; i < k; i += FP_BLOCK) {
              mean[i] = rhs[i];
              e[i] = z[i];
            }
            __syncthreads();
          } else {
            fp_trmv(W, ld, k, rhs, t, stage);
            fp_trmv_t(W, ld, k, t, mean, stage, red);
          }
          ++n_segs;
          if (!replay) {
            if (!solve_form) {
              fp_normals(z, k, gchain, unode, uit, (uint32_t)(h * 512), key);
              fp_trmv_t(W, ld, k, z, e, stage, red);  // e = L^-T z ; beta = mean + sqrt(sig2 lam) e   (samplers.py:214-216)
            }
            for (int i = tid; i < k; i += FP_BLOCK) tmp[i] = mean[i] - mu[i];
            __syncthreads();
            if constexpr (VVM) {
              const double dd = fp_dot(tmp, tmp, k, red), de = fp_dot(tmp, e, k, red), ee = fp_dot(e, e, k, red);
              acc_ss += (dd + 2.0 * sqrt(sig_h * lam) * de) / sig_h + lam * ee;
            } else if (!SEQM || h == 0) {   // seq: only beta_0 enters the lambda^2 rate (samplers.py:249-263)
              acc_dd += fp_dot(tmp, tmp, k, red);
              acc_de += fp_dot(tmp, e, k, red);
              acc_ee += fp_dot(e, e, k, red);
            }
            if constexpr (!GLOBM) {   // kept for the sign counts below (read back by the thread that writes it)
              for (int i = tid; i < k; i += FP_BLOCK) {
                bsc[(2 * h) * k + i] = mean[i];
                bsc[(2 * h + 1) * k + i] = e[i];
              }
            }
          } else {
            for (int i = tid; i < k; i += FP_BLOCK) tmp[i] = rd[3 + k + h * k + i] - mu[i];  // injected beta_h
            __syncthreads();
            if constexpr (VVM)
              acc_ss += fp_dot(tmp, tmp, k, red) / sig_h;
            else if (!SEQM || h == 0)
              acc_dd += fp_dot(tmp, tmp, k, red);
          }
          if (TRACE && td) {
            for (int i = tid; i < k; i += FP_BLOCK) {
              td[10 + 3 * k + 2 * DBN_SMAX * k + h * k + i] = replay ? rd[3 + k + h * k + i] : e[i];  // completed below
              td[10 + 3 * k + h * k + i] = mean[i];
              double dg = 0.0;  // diag of A^-1 = column norms of W
              for (int r = i; r < k; ++r) {
                const double wv = W[(long long)r * ld + i];
                dg += wv * wv;
              }
              td[10 + 3 * k + DBN_SMAX * k + h * k + i] = w * dg;  // x sigma^2 below
            }
          }
          __syncthreads();
          if (SEQM && h >= 1) {   // beta-tilde_{h+1} = A_dlt(G_h)^-1 (beta-tilde_{h-1} + dlt g_h)  (samplers.py:40-46)
            for (int i = tid; i < k; i += FP_BLOCK) rhs[i] = mu_star[i] + dlt * g[i];
            __syncthreads();
            fp_apply_ainv(W, ld, k, rhs, t, z, stage, red);
            for (int i = tid; i < k; i += FP_BLOCK) {
              mu_star[i] = mu[i];
              mu[i] = z[i];
            }
            __syncthreads();
          }
          lo = hi;
        }
      }
      if (tid == 0) {
        const double shape = a.a_sig + a.T_sigma_half, rate = a.b_sig + 0.5 * q_sum;
        const double s2 = VVM ? sigv[0] : (replay ? rd[0] : rate / gamma_draw(gchain, unode, uit, 20000u, key, shape));
        const double sg = sqrt(s2 * lam);
        const double ss = replay ? acc_dd : acc_dd + 2.0 * sg * acc_de + s2 * lam * acc_ee;
        // samplers.py:285-289: shape a + S k / 2, rate b + sum |beta_h - mu|^2 / (2 sigma^2)  (vv: / sigma^2_h per segment)
        // (seq: shape a + (k + 1) / 2, rate b + beta_0.beta_0 / (2 sigma^2), samplers.py:249-263; acc_* hold segment 0 only)
        const double lshape = a.a_lam + 0.5 * (double)(SEQM ? k + 1 : S * k), lrate = a.b_lam + 0.5 * (VVM ? acc_ss : ss / s2);
        const double l2 = replay ? rd[1] : lrate / gamma_draw(gchain, unode, uit, 24000u, key, lshape);
        sc[0] = s2;
        sc[1] = l2;
        if (TRACE && td) {
          if (!VVM) { td[0] = shape; td[1] = rate; td[4] = s2; }
          td[2] = lshape; td[3] = lrate; td[5] = l2;
          td[6] = td[7] = td[8] = td[9] = NAN;
        }
      }
      __syncthreads();
      sig2 = sc[0];
      if (TRACE && td) {
        for (int i = tid; i < S * k; i += FP_BLOCK) {
          const double s2h = VVM ? sigv[i / k] : sig2, sg = sqrt(s2h * w_of(i / k));
          td[10 + 3 * k + DBN_SMAX * k + i] *= s2h;
          if (!replay) td[10 + 3 * k + 2 * DBN_SMAX * k + i] = td[10 + 3 * k + i] + sg * td[10 + 3 * k + 2 * DBN_SMAX * k + i];
        }
      }
      if constexpr (!GLOBM) {
        // ---- counts: betas_vector[2:][burn_in:][x % 10 == 0] (network.py:316-320) keeps the draw of iteration it iff
        // it - 1 >= burn_in and (it - 1 - burn_in) % thin == 0; every segment's beta_h is one posterior sample
        // (beta_post_matrix, scores.py:139-153); pos / neg feed credible_score (scores.py:196-205)
        if (it - 1 >= a.burn_in && ((it - 1 - a.burn_in) % a.thin) == 0) {
          if (tid == 0) atomicAdd(&a.kept[node], (unsigned long long)S);
          for (int j = 1 + tid; j < k; j += FP_BLOCK) {
            unsigned long long npos = 0, nneg = 0;
            for (int h = 0; h < S; ++h) {
              const double sgb = sqrt((VVM ? sigv[h] : sig2) * w_of(h));
              const double b = replay ? rd[3 + k + h * k + j] : bsc[(2 * h) * k + j] + sgb * bsc[(2 * h + 1) * k + j];
              npos += (b >= 0.0) ? 1 : 0;
              nneg += (b <= 0.0) ? 1 : 0;
            }
            const int gene = fp_col(j, node) - 1;
            if (npos) atomicAdd(&a.pos[(long long)node * n + gene], npos);
            if (nneg) atomicAdd(&a.neg[(long long)node * n + gene], nneg);
          }
        }
      }
      if constexpr (SEQM) {
        // ---- delta^2 (samplers.py:220-247): beta-tildes recomputed with (lambda^2 new, delta^2 old), beta_h centred on
        // beta-tilde_{h-1} for h >= 1; shape a + (S - 1) k / 2, rate b + sum |beta_h - bt_{h-1}|^2 / (2 sigma^2)
        const double lam_new = sc[1];
        for (int i = tid; i < k; i += FP_BLOCK) mu[i] = mu_star[i] = 0.0;
        __syncthreads();
        double accd = 0.0;
        const double sgd = sqrt(sig2 * dlt);
        int lo = 0;
        for (int h = 0; h < S; ++h) {
          const int hi = DBN_SEG_END(s_tau[h], h, S, N);
          if (h >= 1) {
            if (h == 1) {
              fp_load_segment(a.tb, node, lo, hi, lam_new, A, ld, k, g, sc + 3);
              fp_factor_inv(A, W, ld, k, t, stage);
              for (int i = tid; i < k; i += FP_BLOCK) rhs[i] = lam_new * g[i];
              __syncthreads();
              fp_apply_ainv(W, ld, k, rhs, t, mu, stage, red);
              ++n_segs;
            }
            for (int i = tid; i < k; i += FP_BLOCK) {
              const double b = replay ? rd[3 + k + h * k + i] : bsc[(2 * h) * k + i] + sgd * bsc[(2 * h + 1) * k + i];
              tmp[i] = b - mu_star[i];
            }
            __syncthreads();
            accd += fp_dot(tmp, tmp, k, red);
            if (h + 1 < S) {
              fp_load_segment(a.tb, node, lo, hi, dlt, A, ld, k, g, sc + 3);
              fp_factor_inv(A, W, ld, k, t, stage);
              for (int i = tid; i < k; i += FP_BLOCK) rhs[i] = mu_star[i] + dlt * g[i];
              __syncthreads();
              fp_apply_ainv(W, ld, k, rhs, t, z, stage, red);
              for (int i = tid; i < k; i += FP_BLOCK) {
                mu_star[i] = mu[i];
                mu[i] = z[i];
              }
              __syncthreads();
              ++n_segs;
            }
          }
          lo = hi;
        }
        if (tid == 0) {
          const double dshape = a.a_del + 0.5 * (double)((S - 1) * k), drate = a.b_del + 0.5 * accd / sig2;
          const double d2 = replay ? rd[3 + k + DBN_SMAX * k] : drate / gamma_draw(gchain, unode, uit, 28000u, key, dshape);
          sc[4] = d2;
          if (TRACE && td) {
            const int o = 10 + 3 * k + 3 * DBN_SMAX * k;   // entry 0 of the per-segment variance slots: delta^2's shape, rate, draw
            td[o] = dshape; td[o + DBN_SMAX] = drate; td[o + 2 * DBN_SMAX] = d2;
          }
        }
        __syncthreads();
      }
      lam = sc[1];
      if constexpr (SEQM) dlt = sc[4];

      if constexpr (VVM) {
        // ---- fpvvGlobCoup.py:99-103: mu_next ~ vvMuSampler(tau, sigma^2_h, lambda^2 new), always accepted; it takes effect
        // after the changepoint move, which still conditions on the old mu (:106-108).  Precision Q = I + sum_h A_h^-1 G_h /
        // sigma^2_h = I + M_w / lam, mean = Q^-1 sum_h A_h^-1 g_h / sigma^2_h (samplers.py:353-379: no mu_in term)
        fp_sweep(a.tb, node, tau_cur, S, lam, A, W, Mc, ld, k, g, t, x, a_cur, red, sc + 3, stage, sigv);
        n_segs += S;
        n_inv += S;
        ++n_qfac;
        const double il = 1.0 / lam;
        for (int i = tid; i < k * ld; i += FP_BLOCK) Q[i] = Mc[i] * il;
        __syncthreads();
        for (int i = tid; i < k; i += FP_BLOCK) {
          Q[i * ld + i] += 1.0;
          rhs[i] = a_cur[i];
        }
        __syncthreads();
        fp_factor_inv(Q, W, ld, k, t, stage);
        fp_trmv(W, ld, k, rhs, t, stage);
        fp_trmv_t(W, ld, k, t, mean, stage, red);
        if (replay) {
          for (int i = tid; i < k; i += FP_BLOCK) mu_star[i] = rd[3 + i];
        } else {
          fp_normals(z, k, gchain, unode, uit, 8192u, key);
          fp_trmv_t(W, ld, k, z, e, stage, red);
          for (int i = tid; i < k; i += FP_BLOCK) mu_star[i] = mean[i] + e[i];
        }
        if (TRACE && td)
          for (int i = tid; i < k; i += FP_BLOCK) td[10 + k + i] = mean[i];
        __syncthreads();
      }

      // ======================= changepoint move (moves.py:25-126) =================================================
      if (a.with_tau) {
        if (tid == 0) {
          uint4 w = make_uint4(0, 0, 0, 0);
          if (!replay) w = philox_block(gchain, unode, uit, 16384u, key);
          const int move = replay ? ri[0] : (int)__umulhi(w.x, 3u);
          const int ns = a.num_samples;
          int valid = 1, Ss = S, hr_num = 1, hr_den = 1;
          if (move == 0) {  // cpBirthMove (changepointMoves.py:93-124)
            int inside = 0;
            for (int j = 0; j < S - 1; ++j) inside += (s_tau[j] >= 2 && s_tau[j] <= ns) ? 1 : 0;
            const int ncand = ns - 1 - inside;
            if (ncand <= 0) {
              valid = 0;
            } else {
              const int p = replay ? ri[1] : (int)__umulhi(w.y, (uint32_t)ncand);
              int c = 2 + p;
              for (int j = 0; j < S - 1; ++j)
                if (s_tau[j] <= c) ++c;
              int w_ = 0;
              bool placed = false;
              for (int j = 0; j < S; ++j) {
                if (!placed && c < s_tau[j]) {
                  s_tas[w_++] = c;
                  placed = true;
                }
                s_tas[w_++] = s_tau[j];
              }
              Ss = S + 1;
              if (Ss > 9) valid = 0;  // moves.py:42 (glob) / :156 (nh: == 10)
              hr_num = GHR ? ns - 1 - S : ns - S;    // moves.py:46 / :159
              hr_den = GHR ? Ss : Ss - 1;
            }
          } else if (S < 2) {
            valid = 0;
          } else if (move == 1) {  // cpDeathMove (:63-91)
            const int idx = replay ? ri[1] : (int)__umulhi(w.y, (uint32_t)(S - 1));
            int w_ = 0;
            for (int j = 0; j < S; ++j)
              if (j != idx) s_tas[w_++] = s_tau[j];
            Ss = S - 1;
            hr_num = GHR ? S : S - 1;  // moves.py:55 / :163
            hr_den = GHR ? ns - 1 - Ss : ns - Ss;
          } else {  // cpRellocationMove (:3-61)
            const int idx = replay ? ri[1] : (int)__umulhi(w.y, (uint32_t)(S - 1));
            const int last = s_tau[S - 1] - 1;
            const int left = (idx == 0) ? 1 : s_tau[idx - 1];
            const int right = (idx + 1 == S - 1) ? last : s_tau[idx + 1];
            const int ncand = right - left - 2;
            if (ncand <= 0) {
              valid = 0;
            } else {
              const int p = replay ? ri[2] : (int)__umulhi(w.z, (uint32_t)ncand);
              int c = left + 1 + p;
              if (c >= s_tau[idx]) ++c;
              for (int j = 0; j < S; ++j) s_tas[j] = (j == idx) ? c : s_tau[j];
            }
          }
          s_int[0] = move; s_int[1] = valid; s_int[2] = Ss; s_int[3] = hr_num; s_int[4] = hr_den;
          sc[2] = replay ? rd[2] : (double)((((uint64_t)w.w << 32) | philox_block(gchain, unode, uit, 16385u, key).x) >> 11) * 0x1.0p-53;
          if (TRACE && ti) { ti[0] = move; ti[1] = valid; ti[2] = 0; }
        }
        __syncthreads();
        const int valid = s_int[1], Ss = s_int[2];
        if (VVM && valid) {
          // vvGlobCoupTauMove (moves.py:238-307): mu is kept, two vvLogMargLikelihood evaluations, u < exp(min(1, delta))
          const double c_seg = a.a_sig * log(a.two_b) - lgamma(a.a_sig);
          const double ml_cur = fp_score_vv(a.tb, node, tau_cur, S, lam, mu, A, ld, k, g, t, x, tmp, rhs, red, sc + 3, stage, a.a_sig,
                                            c_seg, a.two_b);
          const double ml_star = fp_score_vv(a.tb, node, tau_star, Ss, lam, mu, A, ld, k, g, t, x, tmp, rhs, red, sc + 3, stage,
                                             a.a_sig, c_seg, a.two_b);
          n_segs += S + Ss;
          n_evals += 2;
          const double lprior = log_cp_prior(tau_star, Ss, a.log_p, a.log_1mp) - log_cp_prior(tau_cur, S, a.log_p, a.log_1mp);
          const double log_ratio = ml_star - ml_cur + lprior + log((double)s_int[3] / (double)s_int[4]);
          const bool acc = sc[2] < exp(fmin(1.0, log_ratio));
          if (TRACE && td && tid == 0) {
            td[6] = ml_cur; td[7] = ml_star;
          }
          if (TRACE && ti && tid == 0) ti[2] = acc;
          __syncthreads();
          if (acc) {
            S = Ss;
            if (tid < SMAX) s_tau[tid] = s_tas[tid];
          }
          __syncthreads();
        }
        if (SEQM && valid) {
          // changepointsSetMove with method 'seq-coup' (moves.py:184-232): two sequentially coupled marginal likelihoods
          const double ml_cur = fp_score_seq(a.tb, node, tau_cur, S, lam, dlt, mu, mu_star, z, A, W, ld, k, g, t, x, tmp, rhs, red, sc + 3,
                                             stage, a.c0, a.T_half_plus_a, a.two_b);
          const double ml_star = fp_score_seq(a.tb, node, tau_star, Ss, lam, dlt, mu, mu_star, z, A, W, ld, k, g, t, x, tmp, rhs, red,
                                              sc + 3, stage, a.c0, a.T_half_plus_a, a.two_b);
          n_segs += S + Ss;
          n_evals += 2;
          const double lprior = log_cp_prior(tau_star, Ss, a.log_p, a.log_1mp) - log_cp_prior(tau_cur, S, a.log_p, a.log_1mp);
          const double log_ratio = ml_star - ml_cur + lprior + log((double)s_int[3] / (double)s_int[4]);
          const bool acc = sc[2] < fmin(1.0, exp(log_ratio));
          if (TRACE && td && tid == 0) {
            td[6] = ml_cur; td[7] = ml_star;
          }
          if (TRACE && ti && tid == 0) ti[2] = acc;
          __syncthreads();
          if (acc) {
            S = Ss;
            if (tid < SMAX) s_tau[tid] = s_tas[tid];
          }
          __syncthreads();
        }
        if (FPM == FPM_PLAIN && valid) {
          // moves.py:184-232: two plain marginal likelihoods (prior mean 0), u < min(1, exp(delta))
          const FpSummary sc_ = fp_sweep_plain(a.tb, node, tau_cur, S, lam, A, ld, k, g, t, x, red, sc + 3, stage);
          const double ml_cur = a.c0 - 0.5 * sc_.ld - a.T_half_plus_a * log(a.two_b + sc_.s - lam * sc_.c);
          const FpSummary ss_ = fp_sweep_plain(a.tb, node, tau_star, Ss, lam, A, ld, k, g, t, x, red, sc + 3, stage);
          const double ml_star = a.c0 - 0.5 * ss_.ld - a.T_half_plus_a * log(a.two_b + ss_.s - lam * ss_.c);
          n_segs += S + Ss;
          n_evals += 2;
          const double lprior = log_cp_prior(tau_star, Ss, a.log_p, a.log_1mp) - log_cp_prior(tau_cur, S, a.log_p, a.log_1mp);
          const double log_ratio = ml_star - ml_cur + lprior + log((double)s_int[3] / (double)s_int[4]);
          const bool acc = sc[2] < fmin(1.0, exp(log_ratio));
          if (TRACE && td && tid == 0) {
            td[6] = ml_cur; td[7] = ml_star;
          }
          if (TRACE && ti && tid == 0) ti[2] = acc;
          __syncthreads();
          if (acc) {
            S = Ss;
            if (tid < SMAX) s_tau[tid] = s_tas[tid];
          }
          __syncthreads();
        }
        if (GLOBM && valid) {
          const double ils = 1.0 / (lam * sig2);
          // ---- current changepoints: log-ML at mu and the density of mu under muSampler(mu | tau) (moves.py:66-73).  The
          // segments tau shares with tau* go to (Mc, a_cur), the one or two the proposal changes to (Ms, a_star) ...
          for (int i = tid; i < k; i += FP_BLOCK) a_cur[i] = a_star[i] = 0.0;
          for (int i = tid; i < k * ld; i += FP_BLOCK) Mc[i] = Ms[i] = 0.0;
          __syncthreads();
          FpSummary s_com = {0.0, 0.0, 0.0}, s_old = {0.0, 0.0, 0.0};
          {
            const int ne = fp_sweep_split(a.tb, node, tau_cur, S, tau_star, Ss, false, lam, A, W, ld, k, Mc, a_cur, s_com, Ms, a_star, s_old, g, t, x,
                                          red, sc + 3, stage);
            n_segs += ne;
            n_inv += ne;
            n_qfac += 2;
          }
          // ... then (Ms, a_star) become the sums of tau, (Mc, a_cur) stay the shared part, to which the second pass adds tau*'s new segments
          for (int i = tid; i < k * ld; i += FP_BLOCK) {   // ... and Q = I + M / (lam s2) of tau in the same pass (diagonal added below)
            const double m = Ms[i] + Mc[i];
            Ms[i] = m;
            Q[i] = m * ils;
          }
          for (int i = tid; i < k; i += FP_BLOCK) a_star[i] += a_cur[i];
          __syncthreads();
          const FpSummary sc_ = {s_com.c + s_old.c, s_com.s + s_old.s, s_com.ld + s_old.ld};
          double* const M_cur = Ms, *const M_star = Mc, *const av_cur = a_star, *const av_star = a_cur;
          const double q_cur = fp_quad_at(sc_, M_cur, ld, k, av_cur, mu, lam, tmp, red);
          const double ml_cur = a.c0 - 0.5 * sc_.ld - a.T_half_plus_a * log(a.two_b + q_cur);
          ++n_evals;
          // Q = I + M / (lam s2); mean = Q^-1 (a / s2 + mu)   (samplers.py:324, :336: the input mu is added)
          for (int i = tid; i < k; i += FP_BLOCK) {
            Q[i * ld + i] += 1.0;
            rhs[i] = av_cur[i] / sig2 + mu[i];
          }
          __syncthreads();
          double ldq_cur;
          if (BIG) {   // solve form (see the Gibbs pass): mean = Q^-1 rhs by two blocked solves, no inverse factor
            ldq_cur = fpb_factor(Q, W, ld, k, t, stage, false);
            for (int i = tid; i < k; i += FP_BLOCK) z[i] = 0.0;
            __syncthreads();
            fpb_solve_fwd2(Q, W, ld, k, rhs, z, stage);
            fpb_solve_bwd2(Q, W, ld, k, rhs, z, stage);
            for (int i = tid; i < k; i += FP_BLOCK) mean[i] = rhs[i];
            __syncthreads();
          } else {
            ldq_cur = fp_factor_inv(Q, W, ld, k, t, stage);
            fp_trmv(W, ld, k, rhs, t, stage);
            fp_trmv_t(W, ld, k, t, mean, stage, red);
          }
          if (TRACE && td)
            for (int i = tid; i < k; i += FP_BLOCK) td[10 + k + i] = mean[i];
          for (int i = tid; i < k; i += FP_BLOCK) x[i] = mu[i] - mean[i];
          __syncthreads();
          fp_symv(M_cur, ld, k, x, tmp);
          const double dQd_cur = fp_dot(x, x, k, red) + fp_dot(x, tmp, k, red) * ils;
          const double ld_cur = -0.5 * ((double)k * 1.8378770664093454836 - ldq_cur + dQd_cur);
          // ---- proposed changepoints: mu* ~ muSampler(0 | tau*), its density, log-ML at mu* (moves.py:86-93)
          FpSummary ss_ = s_com;
          {
            const int ne = fp_sweep_split(a.tb, node, tau_star, Ss, tau_cur, S, true, lam, A, W, ld, k, M_star, av_star, ss_, M_star, av_star, ss_, g, t,
                                          x, red, sc + 3, stage);
            n_segs += ne;
            n_inv += ne;
          }
          for (int i = tid; i < k * ld; i += FP_BLOCK) Q[i] = M_star[i] * ils;
          __syncthreads();
          for (int i = tid; i < k; i += FP_BLOCK) {
            Q[i * ld + i] += 1.0;
            rhs[i] = av_star[i] / sig2;
          }
          __syncthreads();
          double ldq_star;
          if (BIG) {
            ldq_star = fpb_factor(Q, W, ld, k, t, stage, false);
            for (int i = tid; i < k; i += FP_BLOCK) z[i] = 0.0;
            __syncthreads();
            fpb_solve_fwd2(Q, W, ld, k, rhs, z, stage);
            if (!replay) fp_normals(z, k, gchain, unode, uit, 8192u, key);
            fpb_solve_bwd2(Q, W, ld, k, rhs, z, stage);   // rhs <- Q^-1 rhs = mean, z <- L^-T z (cov = Q^-1)
            for (int i = tid; i < k; i += FP_BLOCK) {
              mean[i] = rhs[i];
              mu_star[i] = replay ? rd[3 + i] : rhs[i] + z[i];
            }
            __syncthreads();
          } else {
          ldq_star = fp_factor_inv(Q, W, ld, k, t, stage);
          fp_trmv(W, ld, k, rhs, t, stage);
          fp_trmv_t(W, ld, k, t, mean, stage, red);
          if (replay) {
            for (int i = tid; i < k; i += FP_BLOCK) mu_star[i] = rd[3 + i];
            __syncthreads();
          } else {
            fp_normals(z, k, gchain, unode, uit, 8192u, key);
            fp_trmv_t(W, ld, k, z, e, stage, red);  // cov = Q^-1 = W^T W
            for (int i = tid; i < k; i += FP_BLOCK) mu_star[i] = mean[i] + e[i];
            __syncthreads();
          }
          }
          if (TRACE && td)
            for (int i = tid; i < k; i += FP_BLOCK) td[10 + 2 * k + i] = mean[i];
          for (int i = tid; i < k; i += FP_BLOCK) x[i] = mu_star[i] - mean[i];
          __syncthreads();
          fp_symv(M_star, ld, k, x, tmp);
          const double dQd_star = fp_dot(x, x, k, red) + fp_dot(x, tmp, k, red) * ils;
          const double ld_star = -0.5 * ((double)k * 1.8378770664093454836 - ldq_star + dQd_star);
          const double q_star = fp_quad_at(ss_, M_star, ld, k, av_star, mu_star, lam, tmp, red);
          const double ml_star = a.c0 - 0.5 * ss_.ld - a.T_half_plus_a * log(a.two_b + q_star);
          ++n_evals;
          const double lp_star = -0.5 * ((double)k * 1.8378770664093454836 + fp_dot(mu_star, mu_star, k, red));
          const double lp_cur = -0.5 * ((double)k * 1.8378770664093454836 + fp_dot(mu, mu, k, red));
          const double lprior = log_cp_prior(tau_star, Ss, a.log_p, a.log_1mp) - log_cp_prior(tau_cur, S, a.log_p, a.log_1mp);
          // moves.py:112-120: u < exp(min(1, delta))
          const double log_ratio = ml_star - ml_cur + lprior + lp_star - lp_cur + ld_cur - ld_star +
                                   log((double)s_int[3] / (double)s_int[4]);
          const bool acc = sc[2] < exp(fmin(1.0, log_ratio));
          if (TRACE && td && tid == 0) {
            td[6] = ml_cur; td[7] = ml_star; td[8] = ld_cur; td[9] = ld_star;
          }
          if (TRACE && ti && tid == 0) ti[2] = acc;
          __syncthreads();
          if (acc) {
            S = Ss;
            if (tid < SMAX) s_tau[tid] = s_tas[tid];
            for (int i = tid; i < k; i += FP_BLOCK) mu[i] = mu_star[i];
          }
          __syncthreads();
        }
        // ---- counts: tau_vector[it] is kept iff it >= burn_in and (it - burn_in) % thin == 0 (network.py:321-322)
        if (it >= a.burn_in && ((it - a.burn_in) % a.thin) == 0 && tid == 0) {
          atomicAdd(&a.cp_kept[node], 1ull);
          for (int j = 0; j < S - 1; ++j) atomicAdd(&a.cp_hist[(long long)node * (N + 3) + s_tau[j]], 1ull);
        }
      }
      if constexpr (VVM) {   // the mean drawn before the changepoint move takes effect now (fpvvGlobCoup.py:103)
        for (int i = tid; i < k; i += FP_BLOCK) mu[i] = mu_star[i];
        __syncthreads();
      }
      // ---- counts: mu_vector[it+1] is kept iff idx >= burn_in and (idx - burn_in) % thin == 0 (network.py:309-311)
      if constexpr (GLOBM) {
        const int idx = it + 1;
        if (idx >= a.burn_in && ((idx - a.burn_in) % a.thin) == 0) {
          if (tid == 0) atomicAdd(&a.kept[node], 1ull);
          for (int j = 1 + tid; j < k; j += FP_BLOCK) {
            const int gene = fp_col(j, node) - 1;
            if (mu[j] >= 0.0) atomicAdd(&a.pos[(long long)node * n + gene], 1ull);
            if (mu[j] <= 0.0) atomicAdd(&a.neg[(long long)node * n + gene], 1ull);
          }
        }
      }
      if (TRACE && td)
        for (int i = tid; i < k; i += FP_BLOCK) td[10 + i] = mu[i];
      if (TRACE && ti && tid == 0) {
        ti[4] = S;
        for (int j = 0; j < SMAX; ++j) ti[5 + j] = (j < S) ? s_tau[j] : -1;
      }
      __syncthreads();
    }

    // ---- store state
    if (tid == 0) {
      a.st_S[u] = S;
      a.st_sig[u] = sig2;
      a.st_lam[u] = lam;
      if (SEQM) a.st_del[u] = dlt;
      atomicAdd(&a.counters[0], (unsigned long long)a.n_iter);
      atomicAdd(&a.counters[1], n_evals);
      atomicAdd(&a.counters[2], n_segs);
      // algorithmic bytes / flops of the executed segment factorisations (SURVEY.md 8d with k = n)
      atomicAdd(&a.counters[3], n_segs * 16ull * (unsigned long long)(k * (k + 1) / 2 + k + 1));
      atomicAdd(&a.counters[4], n_segs * (unsigned long long)((long long)k * k * k / 3 + 4ll * k * k + 5 * k));
      // what the level-3 steps execute: k^3 / 3 per Cholesky factor, another 2 k^3 / 3 where the inverse factor and I - A^-1 are
      // formed (the muSampler precision needs the explicit k x k matrix).  Slots 5 / 6 (unused by the full-parents methods).
      atomicAdd(&a.counters[5], SEQM ? n_segs : n_inv);
      atomicAdd(&a.counters[6], n_segs + n_qfac);
    }
    if (tid < SMAX) a.st_tau[tid * a.U + u] = s_tau[tid];
    for (int i = tid; i < k; i += FP_BLOCK) a.st_mu[(long long)i * a.U + u] = mu[i];
  }
}

inline size_t fp_smem_bytes(int k, bool mats_in_smem) {
  size_t b = (size_t)(FP_NVEC * k + FP_BLOCK_MAX + 32 + 8 + DBN_SMAX) * sizeof(double) + (2 * DBN_SMAX + 8) * sizeof(int);
  if (mats_in_smem) b += (size_t)FP_NMAT * fp_ld(k) * k * sizeof(double);
  else b += (size_t)FPB_SMEM_DOUBLES * sizeof(double) + 16;
  return b;
}
inline size_t fp_ws_doubles_per_block(int k) { return (size_t)FP_NMAT * fp_ld_big(k) * k; }

}  // namespace dbn
