// Kernel 3, h-dbn / nh-dbn variant: the regressor loops of bayesianLinearRegression.py:114-147 and
// bayesianPwLinearRegression.py:89-132 with the segment statistics of the CURRENT state held in a per-unit,
// coalesced cache, so that an iteration gathers from the prefix table only what a proposal changes.
//
// Why (profiles/r1_v0): with every pass re-gathering 2 x 21 scattered doubles per segment, the v0 kernel was bound
// by the L1 miss-request path (one 32-byte sector request per gathered double, ~1 request/cycle/SM) at 12 warps/SM.
// Here
//   * cache[slot][entry][unit] holds the prefix-table row of each of the unit's <= 9 segment boundaries, restricted
//     to the unit's design columns (15 + 5 + 1 doubles at KM = 5).  Threads of a warp read the same (slot, entry) of
//     consecutive units: fully coalesced, 4 sectors per request instead of 1.
//   * the parent-set move scores pi and pi* in ONE sweep: both column sets share a base of <= KM-1 columns, which is
//     factorised once per segment (A_B = L D L^T); the column only pi has and the column only pi* has are two
//     alternative last rows of that factorisation.  Only the new column (<= KM+1 values per boundary) is gathered.
//   * the changepoint move re-evaluates only the segments a birth / death / reallocation changes (<= 2 old, <= 2 new)
//     and gathers one table row (the new boundary) into a scratch slot of the cache.
// Columns are kept in cache-slot order, not sorted by gene id: for mu = 0 (h, nh) the marginal likelihood and the
// Gibbs parameters are invariant to the column order (only rounding differs, ~1e-15 relative); traces are written
// in the reference's sorted order.
#pragma once
#include "dbn_chain.cuh"

namespace dbn {

constexpr int FAST_BLOCK = 128;
#ifndef DBN_FAST_MINB
#define DBN_FAST_MINB 1
#endif
constexpr int BSLOTS = DBN_SMAX;  // cached rows per unit: <= 9 boundaries (8 changepoints + the final row) + 1 scratch

template <int KM>
struct FastDims {
  static constexpr int KB = KM - 1;
  static constexpr int KT = KM * (KM + 1) / 2;
  static constexpr int NE = KT + KM + 1;  // doubles per cached row: ZZ packed by slot | ZY by slot | YY
};

__host__ __device__ inline int fast_cache_row_doubles(int km) { return km * (km + 1) / 2 + km + 1; }

template <int METHOD, int KM, bool TRACE>
__global__ void __launch_bounds__(FAST_BLOCK, DBN_FAST_MINB) chain_fast_kernel(const ChainArgs a) {
  constexpr bool HOMOG = METHOD == DBN_H_DBN;
  constexpr int KB = FastDims<KM>::KB, KT = FastDims<KM>::KT, NE = FastDims<KM>::NE;
  constexpr int KBT = KB * (KB + 1) / 2;

  __shared__ int s_tau[SMAX * FAST_BLOCK];
  __shared__ int s_tas[SMAX * FAST_BLOCK];
  __shared__ unsigned char s_bs[SMAX * FAST_BLOCK];  // cache slot of boundary j+1 (rows tau_j - 1, last = N)
  const int tid = threadIdx.x;
  const int u_raw = blockIdx.x * FAST_BLOCK + tid;
  const bool active = u_raw < a.U;
  const int u = active ? u_raw : 0;
  const int n_it = active ? a.n_iter : 0;
  const int node = u / a.C;
  const int chain = u - node * a.C;
  const int n = a.tb.n, N = a.tb.N, F = n - 1;
  const long long U = a.U;
  auto TAU = [&](int j) -> int& { return s_tau[j * FAST_BLOCK + tid]; };
  auto TAS = [&](int j) -> int& { return s_tas[j * FAST_BLOCK + tid]; };
  auto BS = [&](int j) -> unsigned char& { return s_bs[j * FAST_BLOCK + tid]; };
  auto tau_cur = [&](int j) { return s_tau[j * FAST_BLOCK + tid]; };
  auto tau_star = [&](int j) { return s_tas[j * FAST_BLOCK + tid]; };
  double* const cbase = a.cache + u;
  auto crow = [&](int slot) -> double* { return cbase + (long long)slot * NE * U; };
  const int nb = a.tb.zz_size + node * (n + 2);  // table offset of this node's ZY / YY block

  // ---- load state
  int npi = a.st_npi[u];
  int pi[PMAX];
#pragma unroll
  for (int j = 0; j < PMAX; ++j) pi[j] = a.st_pi[j * a.U + u];
  int S = a.st_S[u];
  for (int j = 0; j < SMAX; ++j) TAU(j) = a.st_tau[j * a.U + u];
  double sig2 = a.st_sig[u], lam = a.st_lam[u];
  int cg[KM];  // gene id held by column slot s (>= 1), -1 = empty; slot 0 is the intercept
  cg[0] = -1;

  // table row `t` restricted to the slots' columns -> cache row `slot` (zeros for empty slots)
  auto gather_row = [&](int t, int slot) {
    const double* __restrict__ row = a.tb.base + (long long)t * a.tb.row_stride;
    double* c = crow(slot);
#pragma unroll
    for (int i = 0; i < KM; ++i) {
      const bool oi = (i == 0) || cg[i] >= 0;
      const int fi = (i == 0) ? 0 : cg[i] + 1;
#pragma unroll
      for (int j = 0; j <= i; ++j) {
        const bool oj = (j == 0) || cg[j] >= 0;
        const int fj = (j == 0) ? 0 : cg[j] + 1;
        const int hi = max(fi, fj), lo = min(fi, fj);
        c[(long long)DBN_TRI(i, j) * U] = (oi && oj) ? __ldg(row + hi * (hi + 1) / 2 + lo) : 0.0;
      }
      c[(long long)(KT + i) * U] = oi ? __ldg(row + nb + fi) : 0.0;
    }
    c[(long long)(KT + KM) * U] = __ldg(row + nb + n + 1);
  };

  if (a.refresh) {
    // (re)build the cache from the table: canonical slots = parent positions, boundary j -> slot j
#pragma unroll
    for (int s = 1; s < KM; ++s) cg[s] = (s - 1 < npi) ? pi[s - 1] : -1;
    for (int j = 0; j < S; ++j) {
      BS(j) = (unsigned char)j;
      if (active) gather_row(DBN_SEG_END(TAU(j), j, S, N), j);
    }
  } else {
#pragma unroll
    for (int s = 1; s < KM; ++s) cg[s] = a.st_cg[(s - 1) * a.U + u];
    for (int j = 0; j < SMAX; ++j) BS(j) = (unsigned char)a.st_bslot[j * a.U + u];
  }

  Rng rng;
  rng.init(a.seed, a.chain_offset + (unsigned)chain, (unsigned)node);
  Work wk = {0, 0, 0, 0};
  unsigned long long n_fact = 0;  // segment factorisations actually executed
  const bool replay = TRACE && a.rd != nullptr;

  // statistics of one segment in slot order: cache row `hi` minus cache row `lo` (lo < 0: table row 0 = zeros)
  auto load_seg = [&](int lo, int hi, double(&G)[KT], double(&g)[KM], double& s) {
    const double* ph = crow(hi);
    const double* pl = crow(lo < 0 ? hi : lo);
    const bool has_lo = lo >= 0;
#pragma unroll
    for (int e = 0; e < KT; ++e) G[e] = ph[(long long)e * U] - (has_lo ? pl[(long long)e * U] : 0.0);
#pragma unroll
    for (int i = 0; i < KM; ++i) g[i] = ph[(long long)(KT + i) * U] - (has_lo ? pl[(long long)(KT + i) * U] : 0.0);
    s = ph[(long long)(KT + KM) * U] - (has_lo ? pl[(long long)(KT + KM) * U] : 0.0);
  };
  // pivot product and quadratic form r^T C^-1 r of one segment (mu = 0), all KM slots
  auto seg_eval = [&](int lo, int hi, double& prod, double& q) {
    double G[KT], g[KM], s, A[KT], dinv[KM], wv[KM];
    load_seg(lo, hi, G, g, s);
    form_a<KM>(G, lam, A);
    prod = ldl_factor<KM>(A, dinv);
    q = segment_q<KM, false>(G, g, s, g, lam, A, dinv, wv);
    ++n_fact;
  };

  for (int li = 0; li < n_it; ++li) {
    const int it = a.it0 + li;
    rng.start_iteration((unsigned)it);
    const double* rd = replay ? a.rd + ((long long)u * a.n_iter + li) * DBN_RD_STRIDE : nullptr;
    const int* ri = replay ? a.ri + ((long long)u * a.n_iter + li) * DBN_RI_STRIDE : nullptr;
    double* td = (TRACE && a.td) ? a.td + ((long long)u * a.n_iter + li) * DBN_TD_STRIDE : nullptr;
    int* ti = (TRACE && a.ti) ? a.ti + ((long long)u * a.n_iter + li) * DBN_TI_STRIDE : nullptr;
    if (TRACE && ti) ti[DBN_TI_S_BEFORE] = S;
    int k = npi + 1;

    // rank of each slot's column in the reference's sorted design matrix (trace only; utils.py:202-204)
    int rank[KM];
    if (TRACE) {
      rank[0] = 0;
#pragma unroll
      for (int i = 1; i < KM; ++i) {
        int r = 1;
#pragma unroll
        for (int j = 1; j < KM; ++j) r += (cg[j] >= 0 && cg[j] < cg[i]) ? 1 : 0;
        rank[i] = (cg[i] >= 0) ? r : KM - 1;
      }
    }

    // =========================== Gibbs pass: sigma^2, beta, lambda^2 ===========================================
    // One sweep with the OLD lambda^2 (samplers.py:103-127, :189-218, :265-305).  beta_h = mean_h + sigma e_h with
    // e_h independent of sigma^2, so sum_h |beta_h|^2 is assembled from (|m|^2, m.e, |e|^2) after sigma^2 is drawn.
    double q_sum = 0.0, acc_dd = 0.0, acc_de = 0.0, acc_ee = 0.0;
    {
      int lo = -1;
      for (int h = 0; h < S; ++h) {
        const int hi = BS(h);
        double G[KT], g[KM], s, A[KT], dinv[KM], wv[KM];
        load_seg(lo, hi, G, g, s);
        form_a<KM>(G, lam, A);
        ldl_factor<KM>(A, dinv);
        ++n_fact;
        q_sum += segment_q<KM, false>(G, g, s, g, lam, A, dinv, wv);
        wk.segment(k, false);
        // posterior mean (I/lam + G)^-1 g = A^-1 (lam g); wv = L^-1 g is already there (samplers.py:210-212)
        double mean[KM];
#pragma unroll
        for (int i = 0; i < KM; ++i) mean[i] = lam * wv[i] * dinv[i];
        solve_upper<KM>(A, mean);
        if (!replay) {
          double e[KM];
          const double sw = sqrt(lam);
#pragma unroll
          for (int i = 0; i < KM; i += 2) {
            double z0, z1;
            rng.normal2(z0, z1);
            e[i] = ((i == 0) || cg[i] >= 0) ? sw * z0 * sqrt(dinv[i]) : 0.0;
            if (i + 1 < KM) e[i + 1] = (cg[i + 1] >= 0) ? sw * z1 * sqrt(dinv[i + 1]) : 0.0;
          }
          solve_upper<KM>(A, e);  // e = sqrt(lam) L^-T D^-1/2 z: cov = sigma^2 lam A^-1 (samplers.py:214-216)
          acc_dd += dot<KM>(mean, mean);
          acc_de += dot<KM>(mean, e);
          acc_ee += dot<KM>(e, e);
          if (TRACE && td) {
#pragma unroll
            for (int i = 0; i < KM; ++i)
              if (i == 0 || cg[i] >= 0) td[DBN_TD_BETA + h * KMAX + rank[i]] = e[i];  // completed below
          }
        } else {
          double ss = 0.0;
          for (int i = 0; i < k; ++i) {
            const double b = rd[DBN_RD_BETA + h * KMAX + i];  // injected beta_h (sorted order; only |beta|^2 is used)
            ss += b * b;
            if (TRACE && td) td[DBN_TD_BETA + h * KMAX + i] = b;
          }
          acc_dd += ss;
        }
        if (TRACE && td) {
          double Aorig[KT], covp[KT];
          form_a<KM>(G, lam, Aorig);
          invert_spd<KM>(Aorig, lam, covp);  // cov / sigma^2 = lam A^-1
#pragma unroll
          for (int i = 0; i < KM; ++i) {
            if (i == 0 || cg[i] >= 0) {
              td[DBN_TD_BETA_MEAN + h * KMAX + rank[i]] = mean[i];
#pragma unroll
              for (int j = 0; j <= i; ++j) {
                if (j == 0 || cg[j] >= 0) {
                  const int ri_ = max(rank[i], rank[j]), rj_ = min(rank[i], rank[j]);
                  td[DBN_TD_BETA_COV + h * DBN_KTRI + DBN_TRI(ri_, rj_)] = covp[DBN_TRI(i, j)];
                }
              }
            }
          }
        }
        lo = hi;
      }
    }
    {
      const double shape = a.a_sig + a.T_sigma_half;
      const double rate = a.b_sig + 0.5 * q_sum;
      sig2 = replay ? rd[DBN_RD_SIGMA2] : rate / rng.gamma(shape);
      if (TRACE && td) {
        td[DBN_TD_SIG_SHAPE] = shape;
        td[DBN_TD_SIG_RATE] = rate;
        td[DBN_TD_SIGMA2] = sig2;
      }
    }
    {
      const double sg = sqrt(sig2);
      if (TRACE && td) {
        for (int h = 0; h < S; ++h) {
          if (!replay)
            for (int i = 0; i < k; ++i)
              td[DBN_TD_BETA + h * KMAX + i] = td[DBN_TD_BETA_MEAN + h * KMAX + i] + sg * td[DBN_TD_BETA + h * KMAX + i];
          for (int i = 0; i < k; ++i)
            for (int j = 0; j <= i; ++j) td[DBN_TD_BETA_COV + h * DBN_KTRI + DBN_TRI(i, j)] *= sig2;
        }
      }
      const double ss = replay ? acc_dd : (acc_dd + 2.0 * sg * acc_de + sig2 * acc_ee);
      // nh: samplers.py:285-289 shape a + S k/2 ; h: :301 shape a + k/2 ; rate b + ss / (2 sigma^2)
      const double shape = a.a_lam + 0.5 * (double)(HOMOG ? k : S * k);
      const double rate = a.b_lam + 0.5 * ss / sig2;
      lam = replay ? rd[DBN_RD_LAMBDA2] : rate / rng.gamma(shape);
      if (TRACE && td) {
        td[DBN_TD_LAM_SHAPE] = shape;
        td[DBN_TD_LAM_RATE] = rate;
        td[DBN_TD_LAMBDA2] = lam;
        td[DBN_TD_DEL_SHAPE] = td[DBN_TD_DEL_RATE] = td[DBN_TD_DELTA2] = NAN;
        td[DBN_TD_LOGML + 0] = td[DBN_TD_LOGML + 1] = td[DBN_TD_LOGML + 2] = td[DBN_TD_LOGML + 3] = NAN;
        for (int e = DBN_TD_MUS_LOGDENS; e < DBN_TD_BETA_MEAN; ++e) td[e] = NAN;
        for (int i = 0; i < KMAX; ++i) td[DBN_TD_MU + i] = NAN;
      }
    }

    // ============================ proposals (no scoring yet) ==================================================
    // ---- parent-set proposal: addMove / deleteMove / exchangeMove (utils.py:271-323)
    const int pmove = replay ? ri[DBN_RI_PI_MOVE] : rng.below(3);
    bool pvalid = true;
    int star[PMAX], nstar = npi;
    int el = -1, cnew = -1;  // gene leaving / entering the parent set
    double p_log_hr = 0.0;
    {
      int spi[PMAX];
#pragma unroll
      for (int j = 0; j < PMAX; ++j) spi[j] = star[j] = pi[j];
      sort4(spi, npi);
      if (pmove == 0) {
        const int ncand = F - npi;
        if (npi > a.fan_in - 1 || ncand <= 0) {
          pvalid = false;
        } else {
          const int p = replay ? ri[DBN_RI_PI_PICK] : rng.below(ncand);
          cnew = nth_candidate(p, node, spi, npi);
#pragma unroll
          for (int j = 0; j < PMAX; ++j)
            if (j == npi) star[j] = cnew;
          nstar = npi + 1;
          p_log_hr = log((double)(F - npi) / (double)nstar);  // moves.py:612
        }
      } else if (npi < 1) {
        pvalid = false;  // utils.py:285-286, :308-309
      } else {
        const int idx = replay ? ri[DBN_RI_PI_PICK] : rng.below(npi);
        el = pi[0];
#pragma unroll
        for (int j = 1; j < PMAX; ++j)
          if (j == idx) el = pi[j];
        int w_ = 0;  // sorted(pi \ el)  (np.setdiff1d, utils.py:290, :318)
#pragma unroll
        for (int j = 0; j < PMAX; ++j) {
          if (j < npi && spi[j] != el) {
#pragma unroll
            for (int q = 0; q < PMAX; ++q)
              if (q == w_) star[q] = spi[j];
            ++w_;
          }
        }
        if (pmove == 1) {
          nstar = npi - 1;
          p_log_hr = log((double)npi / (double)(F - nstar));  // moves.py:615
        } else {
          const int ncand = F - npi;
          if (ncand <= 0) {
            pvalid = false;
          } else {
            const int p = replay ? ri[DBN_RI_PI_PICK + 1] : rng.below(ncand);
            cnew = nth_candidate(p, node, spi, npi);
#pragma unroll
            for (int j = 0; j < PMAX; ++j)
              if (j == npi - 1) star[j] = cnew;
            nstar = npi;
          }
        }
      }
    }
    // ---- changepoint proposal: cpBirthMove / cpDeathMove / cpRellocationMove (changepointMoves.py)
    const bool do_tau = !HOMOG && a.with_tau;
    int tmove = 0, Ss = S;
    bool tvalid = false;
    int tj = 0;      // boundary index (1-based among b_1..b_S) that is inserted / removed / moved
    int trow = 0;    // table row of the new boundary (birth / reallocation)
    double t_log_hr = 0.0;
    if (do_tau) {
      tmove = replay ? ri[DBN_RI_TAU_MOVE] : rng.below(3);
      const int ns = a.num_samples;
      tvalid = true;
      if (tmove == 0) {
        int inside = 0;
        for (int j = 0; j < S - 1; ++j) inside += (TAU(j) >= 2 && TAU(j) <= ns) ? 1 : 0;
        const int ncand = ns - 1 - inside;
        if (ncand <= 0) {
          tvalid = false;
        } else {
          const int p = replay ? ri[DBN_RI_TAU_PICK] : rng.below(ncand);
          int c = 2 + p;
          for (int j = 0; j < S - 1; ++j)
            if (TAU(j) <= c) ++c;
          int w_ = 0;
          bool placed = false;
          for (int j = 0; j < S; ++j) {
            if (!placed && c < TAU(j)) {
              tj = w_ + 1;
              TAS(w_++) = c;
              placed = true;
            }
            TAS(w_++) = TAU(j);
          }
          trow = c - 1;
          Ss = S + 1;
          if (Ss > 9 || !placed || trow >= N) tvalid = false;  // moves.py:156
          t_log_hr = log((double)(ns - S) / (double)(Ss - 1));  // moves.py:159
        }
      } else if (S < 2) {
        tvalid = false;  // changepointMoves.py:81-82, :21-22
      } else if (tmove == 1) {
        const int idx = replay ? ri[DBN_RI_TAU_PICK] : rng.below(S - 1);
        int w_ = 0;
        for (int j = 0; j < S; ++j)
          if (j != idx) TAS(w_++) = TAU(j);
        tj = idx + 1;
        Ss = S - 1;
        t_log_hr = log((double)(S - 1) / (double)(ns - Ss));  // moves.py:163
      } else {
        const int idx = replay ? ri[DBN_RI_TAU_PICK] : rng.below(S - 1);
        const int last = TAU(S - 1) - 1;
        const int left = (idx == 0) ? 1 : TAU(idx - 1);
        const int right = (idx + 1 == S - 1) ? last : TAU(idx + 1);
        const int ncand = right - left - 2;
        if (ncand <= 0) {
          tvalid = false;
        } else {
          const int p = replay ? ri[DBN_RI_TAU_PICK + 1] : rng.below(ncand);
          int c = left + 1 + p;
          if (c >= TAU(idx)) ++c;
          for (int j = 0; j < S; ++j) TAS(j) = (j == idx) ? c : TAU(j);
          tj = idx + 1;
          trow = c - 1;
          if (trow >= N) tvalid = false;
        }
      }
    }
    if (TRACE && ti) {
      ti[DBN_TI_PI_MOVE] = pmove;
      ti[DBN_TI_PI_VALID] = pvalid;
      ti[DBN_TI_PI_ACCEPT] = 0;
      if (do_tau) {
        ti[DBN_TI_TAU_MOVE] = tmove;
        ti[DBN_TI_TAU_VALID] = tvalid;
        ti[DBN_TI_TAU_ACCEPT] = 0;
      }
    }

    // ============================ parent-set move: pi and pi* in one sweep (a4) ===============================
    // position p of the sweep holds slot perm[p]; the last position is the slot that leaves (delete / exchange) or
    // the free slot that would be filled (add).  current = base + last position, star = base + new column.
    double ld_now = 0.0, q_now = 0.0, ml_now = NAN;
    if (pvalid || tvalid) {
      int tslot = KM - 1;
      if (pvalid) {
        if (pmove == 0) {
#pragma unroll
          for (int s = KM - 1; s >= 1; --s)
            if (cg[s] < 0) tslot = s;  // lowest free slot
        } else {
#pragma unroll
          for (int s = 1; s < KM; ++s)
            if (cg[s] == el) tslot = s;
        }
      }
      int perm[KM];
      perm[0] = 0;
#pragma unroll
      for (int p = 1; p < KB; ++p) perm[p] = (p < tslot) ? p : p + 1;
      perm[KB] = tslot;
      const bool has_new = pvalid && cnew >= 0;
      // cache entries of the sweep's positions, table offsets of the new column
      int eb[KBT], et[KB], ezy[KM], o2[KB];
      const int cf = cnew + 1;
#pragma unroll
      for (int p = 0; p < KB; ++p) {
#pragma unroll
        for (int q = 0; q <= p; ++q) eb[DBN_TRI(p, q)] = DBN_TRI(perm[p], perm[q]);  // perm is increasing on the base
        et[p] = (tslot > perm[p]) ? DBN_TRI(tslot, perm[p]) : DBN_TRI(perm[p], tslot);
        ezy[p] = KT + perm[p];
        const int fp = (p == 0) ? 0 : cg[perm[p]] + 1;  // 0 when the slot is empty; masked by o2 < 0
        const bool occ = (p == 0) || cg[perm[p]] >= 0;
        const int hi = max(cf, fp), lo = min(cf, fp);
        o2[p] = (has_new && occ) ? hi * (hi + 1) / 2 + lo : -1;
      }
      ezy[KB] = KT + tslot;
      const int ett = DBN_TRI(tslot, tslot);
      const int o2d = cf * (cf + 1) / 2 + cf, o2y = nb + cf;

      double ldB = 0.0, R1 = 1.0, R2 = 1.0, q1 = 0.0, q2 = 0.0;
      double n_prev[KB + 2];  // new column at the previous boundary: cross terms, diagonal, ZY
#pragma unroll
      for (int i = 0; i < KB + 2; ++i) n_prev[i] = 0.0;
      int lo = -1;
      for (int h = 0; h < S; ++h) {
        const int hi = BS(h);
        const double* ph = crow(hi);
        const double* pl = crow(lo < 0 ? hi : lo);
        const bool has_lo = lo >= 0;
        auto diff = [&](int e) { return ph[(long long)e * U] - (has_lo ? pl[(long long)e * U] : 0.0); };
        double A[KBT], gB[KB], t1[KB], t2[KB], dinv[KB];
#pragma unroll
        for (int e = 0; e < KBT; ++e) A[e] = diff(eb[e]);
#pragma unroll
        for (int p = 0; p < KB; ++p) {
          gB[p] = diff(ezy[p]);
          t1[p] = diff(et[p]);
        }
        const double t1d = diff(ett), t1y = diff(ezy[KB]);
        const double s = diff(KT + KM);
        double t2d = 0.0, t2y = 0.0;
        if (has_new) {
          const double* __restrict__ row = a.tb.base + (long long)DBN_SEG_END(tau_cur(h), h, S, N) * a.tb.row_stride;
          double cur[KB + 2];
#pragma unroll
          for (int p = 0; p < KB; ++p) cur[p] = (o2[p] >= 0) ? __ldg(row + o2[p]) : 0.0;
          cur[KB] = __ldg(row + o2d);
          cur[KB + 1] = __ldg(row + o2y);
#pragma unroll
          for (int p = 0; p < KB; ++p) t2[p] = cur[p] - n_prev[p];
          t2d = cur[KB] - n_prev[KB];
          t2y = cur[KB + 1] - n_prev[KB + 1];
#pragma unroll
          for (int i = 0; i < KB + 2; ++i) n_prev[i] = cur[i];
        } else {
#pragma unroll
          for (int p = 0; p < KB; ++p) t2[p] = 0.0;
        }
        // base factorisation of A_B = I + lam G_B
#pragma unroll
        for (int p = 0; p < KB; ++p)
#pragma unroll
          for (int q = 0; q <= p; ++q) A[DBN_TRI(p, q)] = ((p == q) ? 1.0 : 0.0) + lam * A[DBN_TRI(p, q)];
        const double prodB = ldl_factor<KB>(A, dinv);
        ++n_fact;
        solve_lower<KB>(A, gB);  // w_B = L_B^-1 g_B
        double accB = 0.0;
#pragma unroll
        for (int p = 0; p < KB; ++p) accB += gB[p] * gB[p] * dinv[p];
        // alternative last rows: y = L_B^-1 (lam G_tB); d_t = 1 + lam G_tt - sum y^2 / d; w_t = g_t - sum y w_B / d
        auto tail = [&](double(&y)[KB], double gtt, double gty, double& dt, double& contrib) {
#pragma unroll
          for (int p = 0; p < KB; ++p) y[p] *= lam;
          solve_lower<KB>(A, y);
          double d = 1.0 + lam * gtt, w = gty;
#pragma unroll
          for (int p = 0; p < KB; ++p) {
            const double yd = y[p] * dinv[p];
            d -= y[p] * yd;
            w -= yd * gB[p];
          }
          dt = d;
          contrib = w * w / d;
        };
        double d1, c1, d2, c2;
        tail(t1, t1d, t1y, d1, c1);
        tail(t2, t2d, t2y, d2, c2);
        ldB += log(prodB);
        R1 *= d1;
        R2 *= d2;
        q1 += s - lam * (accB + c1);
        q2 += s - lam * (accB + c2);
        lo = hi;
      }
      const double ld_cur = ldB + log(R1), ld_star = ldB + log(R2);
      const double ml_cur = a.c0 - 0.5 * ld_cur - a.T_half_plus_a * log(a.two_b + q1);  // moves.py:599-600 / :681
      wk.eval();
      for (int h = 0; h < S; ++h) wk.segment(k, false);
      ld_now = ld_cur;
      q_now = q1;
      ml_now = ml_cur;
      if (pvalid) {
        const int ks = nstar + 1;
        const double ml_star = a.c0 - 0.5 * ld_star - a.T_half_plus_a * log(a.two_b + q2);  // moves.py:574-576 / :672
        wk.eval();
        for (int h = 0; h < S; ++h) wk.segment(ks, false);
        const double log_ratio = ml_star - ml_cur + p_log_hr;  // uniform parent-set prior cancels (priors.py:48-68)
        const double uacc = replay ? rd[DBN_RD_U_PI] : rng.uniform();
        const bool acc = uacc < fmin(1.0, exp(log_ratio));  // moves.py:620-633, :699-702
        if (TRACE && td) {
          td[DBN_TD_LOGML + 0] = ml_star;
          td[DBN_TD_LOGML + 1] = ml_cur;
        }
        if (TRACE && ti) ti[DBN_TI_PI_ACCEPT] = acc;
        if (acc) {
          npi = nstar;
          k = ks;
#pragma unroll
          for (int j = 0; j < PMAX; ++j) pi[j] = star[j];
          ld_now = ld_star;
          q_now = q2;
          ml_now = ml_star;
          // the new column (or zeros, for a delete) takes slot `tslot` in every cached boundary row
          for (int h = 0; h < S; ++h) {
            double* c = crow(BS(h));
            const double* __restrict__ row = a.tb.base + (long long)DBN_SEG_END(tau_cur(h), h, S, N) * a.tb.row_stride;
#pragma unroll
            for (int p = 0; p < KB; ++p) c[(long long)et[p] * U] = (o2[p] >= 0) ? __ldg(row + o2[p]) : 0.0;
            c[(long long)ett * U] = has_new ? __ldg(row + o2d) : 0.0;
            c[(long long)ezy[KB] * U] = has_new ? __ldg(row + o2y) : 0.0;
          }
#pragma unroll
          for (int s = 1; s < KM; ++s)
            if (s == tslot) cg[s] = has_new ? cnew : -1;
        }
      }
    }
    // ---- counts: pi_vector[it+1] is kept iff idx >= burn_in and (idx - burn_in) % thin != 0 (network.py:358-359)
    {
      const int idx = it + 1;
      if (idx >= a.burn_in && ((idx - a.burn_in) % a.thin) != 0 && npi > 0) {
        atomicAdd(&a.nonempty[node], 1ull);
#pragma unroll
        for (int j = 0; j < PMAX; ++j)
          if (j < npi) atomicAdd(&a.edge[(long long)node * n + pi[j]], 1ull);
      }
      if (TRACE && ti) {
        ti[DBN_TI_NPI] = npi;
        for (int j = 0; j < PMAX; ++j) ti[DBN_TI_PI + j] = (j < npi) ? pi[j] : -1;
      }
    }

    // ============================ changepoint move: only the changed segments (a5) ============================
    if (do_tau) {
      if (tvalid) {
        // boundaries around the change: left = b_{tj-1}, mid = b_tj (absent for a birth), right = the next one
        unsigned used = 0;
        for (int j = 0; j < S; ++j) used |= 1u << BS(j);
        const int scratch = __ffs(~used) - 1;
        const bool birth = tmove == 0, death = tmove == 1;
        const int left = (tj >= 2) ? (int)BS(tj - 2) : -1;
        const int mid = birth ? -1 : (int)BS(tj - 1);
        const int right = birth ? (int)BS(tj - 1) : (int)BS(tj);
        if (!death) gather_row(trow, scratch);
        double prod_old = 1.0, q_old = 0.0, prod_new = 1.0, q_new = 0.0, pr, qq;
        if (birth) {
          seg_eval(left, right, pr, qq); prod_old *= pr; q_old += qq;
        } else {
          seg_eval(left, mid, pr, qq); prod_old *= pr; q_old += qq;
          seg_eval(mid, right, pr, qq); prod_old *= pr; q_old += qq;
        }
        if (death) {
          seg_eval(left, right, pr, qq); prod_new *= pr; q_new += qq;
        } else {
          seg_eval(left, scratch, pr, qq); prod_new *= pr; q_new += qq;
          seg_eval(scratch, right, pr, qq); prod_new *= pr; q_new += qq;
        }
        const double ld_star = ld_now + log(prod_new / prod_old);
        const double q_star = q_now - q_old + q_new;
        const double ml_star = a.c0 - 0.5 * ld_star - a.T_half_plus_a * log(a.two_b + q_star);  // moves.py:209-216
        wk.eval();
        for (int h = 0; h < Ss; ++h) wk.segment(k, false);
        const double lprior = log_cp_prior(tau_star, Ss, a.log_p, a.log_1mp) - log_cp_prior(tau_cur, S, a.log_p, a.log_1mp);
        const double log_ratio = ml_star - ml_now + lprior + t_log_hr;  // moves.py:225
        const double uacc = replay ? rd[DBN_RD_U_TAU] : rng.uniform();
        const bool acc = uacc < fmin(1.0, exp(log_ratio));
        if (TRACE && td) {
          td[DBN_TD_LOGML + 2] = ml_now;
          td[DBN_TD_LOGML + 3] = ml_star;
        }
        if (TRACE && ti) ti[DBN_TI_TAU_ACCEPT] = acc;
        if (acc) {
          if (birth) {
            for (int j = S; j >= tj; --j) BS(j) = BS(j - 1);
            BS(tj - 1) = (unsigned char)scratch;
          } else if (death) {
            for (int j = tj - 1; j < S - 1; ++j) BS(j) = BS(j + 1);
          } else {
            BS(tj - 1) = (unsigned char)scratch;
          }
          S = Ss;
          for (int j = 0; j < Ss; ++j) TAU(j) = TAS(j);
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
#pragma unroll
    for (int s = 1; s < KM; ++s) a.st_cg[(s - 1) * a.U + u] = cg[s];
    for (int j = 0; j < SMAX; ++j) a.st_bslot[j * a.U + u] = BS(j);
  }
  // ---- work counters: one atomic per warp
  unsigned long long v[6] = {(unsigned long long)n_it, wk.evals, wk.segs, wk.bytes, wk.flops, n_fact};
  __syncwarp();
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    unsigned long long x = v[c];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
    if ((threadIdx.x & 31) == 0 && x) atomicAdd(&a.counters[c == 5 ? 6 : c], x);
  }
}

}  // namespace dbn
