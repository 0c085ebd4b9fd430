// Device building blocks of the scorer: prefix-table gather, k x k LDL^T, and the per-state passes.
//
// One thread evaluates one (node, chain) state: every matrix is at most KM x KM (KM = 1 + fan-in <= 5) and is
// held in registers as a packed lower triangle with all loops unrolled.  A state with k < KM design columns is
// padded with identity rows (G = 0 there), which leaves every real entry's arithmetic untouched, so threads of
// a warp with different parent-set sizes execute the same instruction stream.
//
// Math = SURVEY.md Appendix A: with A_h = I + w G_h = L D L^T,
//   ln det C_h = sum ln d_j,   r^T C_h^-1 r = rho - w (L^-1 v)^T D^-1 (L^-1 v),  v = g - G mu, rho = s - 2 mu.g + mu.G.mu
// replacing the T_h x T_h det/inv of marginalLikelihood.py:104-144 and samplers.py:112-116.
#pragma once
#include <cstdint>
#include <cmath>

namespace dbn {

#define DBN_TRI(i, j) ((i) * ((i) + 1) / 2 + (j))

// ----- prefix table geometry (one row per t = 0..N) -------------------------------------------------------
// row = [ ZZ packed lower triangle over z = (1, x_0..x_{n-1}) : (n+1)(n+2)/2 doubles |
//         for each node i (of the handle's node range): ZY[i][0..n] (sum z_a * y_i), YY[i] (sum y_i^2) : (n + 2) doubles each ]
struct Table {
  const double* base;
  long long row_stride;  // doubles per row
  int n;
  int N;        // regression rows; table has N+1 rows, row 0 is all zeros
  int zz_size;  // (n+1)(n+2)/2
  int zy_base;  // offset of node i's ZY | YY block = zy_base + i (n + 2): zz_size when the table holds every node's block,
                // zz_size - node0 (n + 2) when it holds only the blocks of nodes node0.. (a handle restricted to a unit range)
};

template <int KM>
struct Cols {
  int zz[KM * (KM + 1) / 2];
  int zy[KM];
  int yy;
  int k;
};

// f[0] = 0 (intercept); f[1..k-1] = sorted parent gene ids + 1 (utils.py:202-204: columns sorted by label)
template <int KM>
__device__ __forceinline__ void make_cols(Cols<KM>& c, const Table& t, int node, const int (&f)[KM], int k) {
  c.k = k;
  const int nb = t.zy_base + node * (t.n + 2);
#pragma unroll
  for (int i = 0; i < KM; ++i) {
    const int fi = f[i];
    c.zy[i] = nb + fi;
#pragma unroll
    for (int j = 0; j <= i; ++j) c.zz[DBN_TRI(i, j)] = fi * (fi + 1) / 2 + f[j];
  }
  c.yy = nb + t.n + 1;
}

// segment statistics of rows [lo, hi): difference of two table rows restricted to the state's columns
template <int KM>
__device__ __forceinline__ void load_segment(const Table& t, const Cols<KM>& c, int lo, int hi,
                                             double (&G)[KM * (KM + 1) / 2], double (&g)[KM], double& s) {
  const double* __restrict__ ph = t.base + (long long)hi * t.row_stride;
  const double* __restrict__ pl = t.base + (long long)lo * t.row_stride;
#pragma unroll
  for (int i = 0; i < KM; ++i) {
    const bool on = i < c.k;
    g[i] = on ? (__ldg(ph + c.zy[i]) - __ldg(pl + c.zy[i])) : 0.0;
#pragma unroll
    for (int j = 0; j <= i; ++j) G[DBN_TRI(i, j)] = on ? (__ldg(ph + c.zz[DBN_TRI(i, j)]) - __ldg(pl + c.zz[DBN_TRI(i, j)])) : 0.0;
  }
  s = __ldg(ph + c.yy) - __ldg(pl + c.yy);
}

// reciprocal of a pivot.  Pivots of I + w G (G positive semi-definite) are finite and >= 1, so the special-case
// handling of the IEEE division (~2x the instructions) is not needed: hardware seed (rcp.approx.ftz.f64, ~2^-20
// relative) and two Newton steps; the result is within 1 ulp of 1/d.
__device__ __forceinline__ double pivot_rcp(double d) {
#ifdef DBN_EXACT_DIV
  return 1.0 / d;
#else
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  return r;
#endif
}

// 1 / sqrt(d) of a pivot (same domain as pivot_rcp): hardware seed (rsqrt.approx.ftz.f64, ~2^-20 relative) and one
// third-order step r += r e (1/2 + 3/8 e), e = 1 - d r^2: relative error ~e^3, i.e. below 1 ulp.
__device__ __forceinline__ double pivot_rsqrt(double d) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  const double e = fma(-d * r, r, 1.0);
  return fma(r * e, fma(0.375, e, 0.5), r);
}

// ----- KM x KM symmetric positive definite algebra on packed lower triangles ------------------------------
// In place A -> unit-lower L (strict part) ; returns prod_j d_j and fills dinv[j] = 1/d_j.
template <int KM>
__device__ __forceinline__ double ldl_factor(double (&A)[KM * (KM + 1) / 2], double (&dinv)[KM]) {
  double prod = 1.0;
#pragma unroll
  for (int j = 0; j < KM; ++j) {
    const double d = A[DBN_TRI(j, j)];
    prod *= d;
    const double di = pivot_rcp(d);
    dinv[j] = di;
#pragma unroll
    for (int i = j + 1; i < KM; ++i) {
      const double l = A[DBN_TRI(i, j)] * di;
#pragma unroll
      for (int m = j + 1; m <= i; ++m) A[DBN_TRI(i, m)] -= l * A[DBN_TRI(m, j)];
    }
#pragma unroll
    for (int i = j + 1; i < KM; ++i) A[DBN_TRI(i, j)] *= di;
  }
  return prod;
}

// Same factorisation with the pivots returned as rs[j] = 1 / sqrt(d_j) and dinv[j] = rs[j]^2 (one SFU seed per pivot
// serves both): the Gibbs sweep needs sqrt(1 / d_j) to scale its normals (samplers.py:214-216).
template <int KM>
__device__ __forceinline__ void ldl_factor_rs(double (&A)[KM * (KM + 1) / 2], double (&dinv)[KM], double (&rs)[KM]) {
#pragma unroll
  for (int j = 0; j < KM; ++j) {
    const double r = pivot_rsqrt(A[DBN_TRI(j, j)]);
    const double di = r * r;
    rs[j] = r;
    dinv[j] = di;
#pragma unroll
    for (int i = j + 1; i < KM; ++i) {
      const double l = A[DBN_TRI(i, j)] * di;
#pragma unroll
      for (int m = j + 1; m <= i; ++m) A[DBN_TRI(i, m)] -= l * A[DBN_TRI(m, j)];
    }
#pragma unroll
    for (int i = j + 1; i < KM; ++i) A[DBN_TRI(i, j)] *= di;
  }
}

// w <- L^-1 w (unit lower)
template <int KM>
__device__ __forceinline__ void solve_lower(const double (&L)[KM * (KM + 1) / 2], double (&w)[KM]) {
#pragma unroll
  for (int i = 1; i < KM; ++i)
#pragma unroll
    for (int p = 0; p < i; ++p) w[i] -= L[DBN_TRI(i, p)] * w[p];
}

// x <- L^-T x (unit lower transposed)
template <int KM>
__device__ __forceinline__ void solve_upper(const double (&L)[KM * (KM + 1) / 2], double (&x)[KM]) {
#pragma unroll
  for (int i = KM - 2; i >= 0; --i)
#pragma unroll
    for (int p = i + 1; p < KM; ++p) x[i] -= L[DBN_TRI(p, i)] * x[p];
}

// x <- (L D L^T)^-1 x
template <int KM>
__device__ __forceinline__ void solve_ldl(const double (&L)[KM * (KM + 1) / 2], const double (&dinv)[KM], double (&x)[KM]) {
  solve_lower<KM>(L, x);
#pragma unroll
  for (int i = 0; i < KM; ++i) x[i] *= dinv[i];
  solve_upper<KM>(L, x);
}

// y = G x for packed symmetric G
template <int KM>
__device__ __forceinline__ void symv(const double (&G)[KM * (KM + 1) / 2], const double (&x)[KM], double (&y)[KM]) {
#pragma unroll
  for (int i = 0; i < KM; ++i) {
    double a = 0.0;
#pragma unroll
    for (int j = 0; j < KM; ++j) a += G[(i >= j) ? DBN_TRI(i, j) : DBN_TRI(j, i)] * x[j];
    y[i] = a;
  }
}

template <int KM>
__device__ __forceinline__ double dot(const double (&a)[KM], const double (&b)[KM]) {
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < KM; ++i) s += a[i] * b[i];
  return s;
}

// A = I + w G (packed)
template <int KM>
__device__ __forceinline__ void form_a(const double (&G)[KM * (KM + 1) / 2], double w, double (&A)[KM * (KM + 1) / 2]) {
#pragma unroll
  for (int i = 0; i < KM; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) A[DBN_TRI(i, j)] = (i == j ? 1.0 : 0.0) + w * G[DBN_TRI(i, j)];
}

// quadratic form r^T C^-1 r of one segment given the factorisation of A = I + w G
template <int KM, bool HAS_MU>
__device__ __forceinline__ double segment_q(const double (&G)[KM * (KM + 1) / 2], const double (&g)[KM], double s,
                                            const double (&mu)[KM], double w, const double (&L)[KM * (KM + 1) / 2],
                                            const double (&dinv)[KM], double (&wv)[KM]) {
  double rho = s;
  if (HAS_MU) {
    double Gmu[KM];
    symv<KM>(G, mu, Gmu);
#pragma unroll
    for (int i = 0; i < KM; ++i) wv[i] = g[i] - Gmu[i];
    rho = s - 2.0 * dot<KM>(mu, g) + dot<KM>(mu, Gmu);
  } else {
#pragma unroll
    for (int i = 0; i < KM; ++i) wv[i] = g[i];
  }
  solve_lower<KM>(L, wv);
  double acc = 0.0;
#pragma unroll
  for (int i = 0; i < KM; ++i) acc += wv[i] * wv[i] * dinv[i];
  return rho - w * acc;
}

// work accounting: algorithmic bytes and FP64 flops of one segment evaluation with k columns
struct Work {
  unsigned long long evals, segs, bytes, flops;
  __device__ __forceinline__ void segment(int k, bool has_mu) {
    segs += 1;
    bytes += 16ull * (unsigned)(k * (k + 1) / 2 + k + 1);
    flops += (unsigned)((k * k * k) / 3 + 2 * k * k + 5 * k + (has_mu ? 2 * k * k : 0));
  }
  __device__ __forceinline__ void eval() {
    evals += 1;
    flops += 12;
  }
};

// end row (exclusive) of segment h: tau[h] - 1, N for the last one (utils.py:179,262)
#define DBN_SEG_END(tau_h, h, S, N) (((h) == (S)-1) ? (N) : ((tau_h)-1))

// ----- segment sources: seg.begin() starts a pass, seg(h, G, g, s) loads the statistics of segment h (h = 0, 1, .. in order)
// and returns its length.  The coupled chain kernel has a second source that reads cached boundary rows (dbn_chain.cuh).
// From the prefix table (two gathered rows per segment):
template <int KM, class TauFn>
struct SegFromTable {
  const Table& t;
  const Cols<KM>& c;
  TauFn tau;
  int S;
  __device__ __forceinline__ void begin() const {}
  __device__ __forceinline__ int operator()(int h, double (&G)[KM * (KM + 1) / 2], double (&g)[KM], double& s) const {
    const int lo = h ? DBN_SEG_END(tau(h - 1), h - 1, S, t.N) : 0;
    const int hi = DBN_SEG_END(tau(h), h, S, t.N);
    load_segment<KM>(t, c, lo, hi, G, g, s);
    return hi - lo;
  }
};
template <int KM, class TauFn>
__device__ __forceinline__ SegFromTable<KM, TauFn> seg_from_table(const Table& t, const Cols<KM>& c, TauFn tau, int S) {
  return SegFromTable<KM, TauFn>{t, c, tau, S};
}

// ----- a2/a3: log marginal likelihood, one prior mean for all segments (nh/h: zeros; glob: mu) ------------
// tau(h) is a callable returning the h-th changepoint.  c0 = lgamma(T/2+a) - lgamma(a) - (T/2) ln(pi) + a ln(2b).
template <int KM, bool HAS_MU, class SegFn>
__device__ __forceinline__ double score_plain_seg(SegFn seg, int S, int k, double lam, const double (&mu)[KM], double c0, double T_half_plus_a,
                                                  double two_b, Work& wk) {
  double acc_ld = 0.0, acc_q = 0.0;
  seg.begin();
  for (int h = 0; h < S; ++h) {
    double G[KM * (KM + 1) / 2], g[KM], s, A[KM * (KM + 1) / 2], dinv[KM], wv[KM];
    seg(h, G, g, s);
    form_a<KM>(G, lam, A);
    acc_ld += log(ldl_factor<KM>(A, dinv));
    acc_q += segment_q<KM, HAS_MU>(G, g, s, mu, lam, A, dinv, wv);
    wk.segment(k, HAS_MU);
  }
  wk.eval();
  return c0 - 0.5 * acc_ld - T_half_plus_a * log(two_b + acc_q);
}
template <int KM, bool HAS_MU, class TauFn>
__device__ __forceinline__ double score_plain(const Table& t, const Cols<KM>& c, TauFn tau, int S, double lam,
                                              const double (&mu)[KM], double c0, double T_half_plus_a, double two_b,
                                              Work& wk) {
  return score_plain_seg<KM, HAS_MU>(seg_from_table<KM>(t, c, tau, S), S, c.k, lam, mu, c0, T_half_plus_a, two_b, wk);
}

// ----- vvLogMargLikelihood (marginalLikelihood.py:5-43): segment-specific variances ---------------------------
// The sum over the segments of single-segment log marginal likelihoods, each with its own length T_h in the Gamma, pi
// and exponent terms (:29-30, :39), lambda^2 and the prior mean mu shared (:13, :32).
// c_seg = a ln(2b) - lgamma(a); ln(pi) = 1.1447298858494002.
template <int KM, class SegFn>
__device__ __forceinline__ double score_vv_seg(SegFn seg, int S, int k, double lam, const double (&mu)[KM], double a_sig, double c_seg,
                                               double two_b, Work& wk) {
  double acc = 0.0;
  seg.begin();
  for (int h = 0; h < S; ++h) {
    double G[KM * (KM + 1) / 2], g[KM], s, A[KM * (KM + 1) / 2], dinv[KM], wv[KM];
    const int len = seg(h, G, g, s);
    form_a<KM>(G, lam, A);
    const double ld = log(ldl_factor<KM>(A, dinv));
    const double q = segment_q<KM, true>(G, g, s, mu, lam, A, dinv, wv);
    const double Th_half = 0.5 * (double)len;
    acc += lgamma(Th_half + a_sig) + c_seg - Th_half * 1.1447298858494002 - 0.5 * ld - (Th_half + a_sig) * log(two_b + q);
    wk.segment(k, true);
    wk.flops += 12;
  }
  wk.eval();
  return acc;
}
template <int KM, class TauFn>
__device__ __forceinline__ double score_vv(const Table& t, const Cols<KM>& c, TauFn tau, int S, double lam,
                                           const double (&mu)[KM], double a_sig, double c_seg, double two_b, Work& wk) {
  return score_vv_seg<KM>(seg_from_table<KM>(t, c, tau, S), S, c.k, lam, mu, a_sig, c_seg, two_b, wk);
}
// ----- a2 + a8: sequentially coupled log marginal likelihood ------------------------------------------------
// Prior means are betaTildeSampler's (samplers.py:4-54): bt[0] = 0, bt[1] = (I/lam + G_1)^-1 g_1,
// bt[h] = (I/dlt + G_{h-1})^-1 (bt[h-2]/dlt + g_{h-1}); C_h uses dlt for h > 0 (marginalLikelihood.py:104-107).
template <int KM, class SegFn>
__device__ __forceinline__ double score_seq_seg(SegFn seg, int S, int k, double lam, double dlt, double c0, double T_half_plus_a, double two_b,
                                                Work& wk) {
  double acc_ld = 0.0, acc_q = 0.0;
  double bt_cur[KM], bt_prev[KM];
#pragma unroll
  for (int i = 0; i < KM; ++i) bt_cur[i] = bt_prev[i] = 0.0;
  seg.begin();
  for (int h = 0; h < S; ++h) {
    double G[KM * (KM + 1) / 2], g[KM], s, A[KM * (KM + 1) / 2], dinv[KM], wv[KM];
    seg(h, G, g, s);
    if (h == 1) {  // bt[1] from this segment's own data with lambda^2
      form_a<KM>(G, lam, A);
      ldl_factor<KM>(A, dinv);
#pragma unroll
      for (int i = 0; i < KM; ++i) bt_cur[i] = lam * g[i];
      solve_ldl<KM>(A, dinv, bt_cur);
      wk.segment(k, false);
    }
    const double w = (h == 0) ? lam : dlt;
    form_a<KM>(G, w, A);
    acc_ld += log(ldl_factor<KM>(A, dinv));
    acc_q += segment_q<KM, true>(G, g, s, bt_cur, w, A, dinv, wv);
    wk.segment(k, true);
    if (h >= 1) {  // bt[h+1] = A_dlt(h)^-1 (bt[h-1] + dlt g_h)
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
  }
  wk.eval();
  return c0 - 0.5 * acc_ld - T_half_plus_a * log(two_b + acc_q);
}
template <int KM, class TauFn>
__device__ __forceinline__ double score_seq(const Table& t, const Cols<KM>& c, TauFn tau, int S, double lam, double dlt,
                                            double c0, double T_half_plus_a, double two_b, Work& wk) {
  return score_seq_seg<KM>(seg_from_table<KM>(t, c, tau, S), S, c.k, lam, dlt, c0, T_half_plus_a, two_b, wk);
}

// ----- a9: muSampler's Gaussian (samplers.py:307-351) on sufficient statistics ----------------------------
// Q = I + sum_h A_h^-1 G_h / s2 (precision), m0 = sum_h A_h^-1 g_h / s2 (mean numerator WITHOUT the mu_in term).
// sig2(h) is a callable: the same sigma^2 for every segment (glob), or the segment-specific sigma^2_h of vvMuSampler
// (samplers.py:353-394).
// With SUMS the same sweep also returns ld = sum ln det A_h, s = sum y_h^T y_h, c = sum g_h^T A_h^-1 g_h: together with Q
// and m0 they give the marginal likelihood at ANY mean (ml_from_sums), so the globally coupled moves need one sweep per
// (parent set, changepoint list) instead of two (muSampler sums, then the score at the mean drawn from them).
struct MuSums {
  double ld, s, c;
};
struct NoSegmentHook {
  template <class... T>
  __device__ __forceinline__ void operator()(T&&...) const {}
};
// `hook(len, G, g, s, A, dinv, prod)` sees every segment's statistics and factorisation (the var-glob kernel scores the
// current mean with it in the same sweep: its per-segment log(2b + q_h) terms cannot be recovered from the sums)
template <int KM, bool SUMS = false, class SegFn, class SigFn, class Hook = NoSegmentHook>
__device__ __forceinline__ MuSums mu_posterior_seg(SegFn seg, int S, int k, double lam, SigFn sig2, double (&Q)[KM * (KM + 1) / 2],
                                                   double (&m0)[KM], Work& wk, Hook hook = Hook()) {
  MuSums r = {0.0, 0.0, 0.0};
#pragma unroll
  for (int i = 0; i < KM; ++i) {
    m0[i] = 0.0;
#pragma unroll
    for (int j = 0; j <= i; ++j) Q[DBN_TRI(i, j)] = (i == j) ? 1.0 : 0.0;
  }
  seg.begin();
  for (int h = 0; h < S; ++h) {
    double G[KM * (KM + 1) / 2], g[KM], s, A[KM * (KM + 1) / 2], dinv[KM];
    const int len = seg(h, G, g, s);
    form_a<KM>(G, lam, A);
    const double prod = ldl_factor<KM>(A, dinv);
    hook(len, G, g, s, A, dinv, prod);
    const double is2 = 1.0 / sig2(h);
    double x[KM];
#pragma unroll
    for (int i = 0; i < KM; ++i) x[i] = g[i];
    solve_ldl<KM>(A, dinv, x);
    if (SUMS) {
      r.ld += log(prod);
      r.s += s;
      r.c += dot<KM>(g, x);
    }
#pragma unroll
    for (int i = 0; i < KM; ++i) m0[i] += x[i] * is2;
    // columns of A^-1 G (symmetric): only the lower triangle is accumulated
#pragma unroll
    for (int col = 0; col < KM; ++col) {
#pragma unroll
      for (int i = 0; i < KM; ++i) x[i] = G[(i >= col) ? DBN_TRI(i, col) : DBN_TRI(col, i)];
      solve_ldl<KM>(A, dinv, x);
#pragma unroll
      for (int i = col; i < KM; ++i) Q[DBN_TRI(i, col)] += x[i] * is2;
    }
    wk.segment(k, true);
  }
  return r;
}
// log marginal likelihood at prior mean mu from the sums of one sweep (one variance s2 for all segments):
// sum_h r_h^T C_h^-1 r_h = s - lam c - 2 mu.a + mu^T B mu with a = sum A_h^-1 g_h = s2 m0, B = sum A_h^-1 G_h = s2 (Q - I)
// (lam G A^-1 = I - A^-1; the same identity the full-parents kernel uses, dbn_fp.cuh)
template <int KM>
__device__ __forceinline__ double ml_from_sums(const MuSums& r, const double (&Q)[KM * (KM + 1) / 2], const double (&m0)[KM], double s2,
                                               double lam, const double (&mu)[KM], double c0, double T_half_plus_a, double two_b, Work& wk) {
  double Qmu[KM];
  symv<KM>(Q, mu, Qmu);
  const double q = r.s - lam * r.c - 2.0 * s2 * dot<KM>(mu, m0) + s2 * (dot<KM>(mu, Qmu) - dot<KM>(mu, mu));
  wk.eval();
  return c0 - 0.5 * r.ld - T_half_plus_a * log(two_b + q);
}
template <int KM, class TauFn, class SigFn>
__device__ __forceinline__ void mu_posterior(const Table& t, const Cols<KM>& c, TauFn tau, int S, double lam, SigFn sig2,
                                             double (&Q)[KM * (KM + 1) / 2], double (&m0)[KM], Work& wk) {
  mu_posterior_seg<KM>(seg_from_table<KM>(t, c, tau, S), S, c.k, lam, sig2, Q, m0, wk);
}

}  // namespace dbn
