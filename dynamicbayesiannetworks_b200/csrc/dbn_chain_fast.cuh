// Kernel 3, h-dbn / nh-dbn variant: the regressor loops of bayesianLinearRegression.py:114-147 and
// bayesianPwLinearRegression.py:89-132 with the segment statistics of the CURRENT state held in a per-unit,
// coalesced cache, so that an iteration gathers from the prefix table only what a proposal changes.
//
// Why (profiles/r1_v0): with every pass re-gathering 2 x 21 scattered doubles per segment, the v0 kernel was bound
// by the L1 miss-request path (one 32-byte sector request per gathered double, ~1 request/cycle/SM) at 12 warps/SM.
// Here
//   * cache[j][pair][unit], j < 9, holds prefix-table row b_j (the end boundary of the unit's j-th segment, in time
//     order) restricted to the unit's design columns, as 16-byte pairs: 14 + 5 + 1 doubles at KM = 5 (the (0, 0) entry
//     of ZZ is not stored: sum 1 * 1 over the rows before b is b itself, so a segment's G[0][0] is its length).  Threads
//     of a warp read the same (j, pair) of consecutive units: fully coalesced, 10 x 16 bytes per thread, segment and
//     sweep.  A segment's statistics (G_h, g_h, s_h) are the difference of two consecutive rows, taken between two
//     shared-memory staging buffers.  Rows are kept physically ordered, because a per-unit row indirection scatters
//     the warp's addresses again (measured: profiles/r1_v1).  Caching boundary rows rather than their differences
//     (v6/v7) makes every update a copy: no table row has to be gathered again when a move is accepted.
//   * round 2: everything a sweep step needs is requested one step ahead by cp.async and nothing is gathered with a
//     dependent load inside a sweep any more (profiles/r1_v10: 37 % of all long-scoreboard stalls were the candidate
//     column's six scattered loads per segment, 12 % the gather of the changepoint move's new boundary row).  A step
//     first moves its operands from the staging buffers into registers and only then issues the next step's copies,
//     so two row buffers suffice; the third holds the changepoint proposal's new boundary row, requested before the
//     parent-set sweep starts, and a six-double slot holds the candidate column at the next boundary.
//   * the parent-set move scores pi and pi* in ONE sweep: both column sets share a base of <= KM-1 columns, which is
//     factorised once per segment (A_B = L D L^T); the column only pi has and the column only pi* has are two
//     alternative last rows of that factorisation.  Only the new column (<= KM+1 values per boundary) is gathered.
//   * the changepoint move re-evaluates only the segments a birth / death / reallocation changes (<= 2 old, <= 2 new)
//     from the <= 3 cached boundary rows around the change and ONE gathered table row, the new boundary (none for a
//     death).
//   * accepted moves rewrite the owner's rows: the new column in every row (parent-set move), or one row inserted /
//     removed / replaced (changepoint move), as 16-byte pairs.  Only the thread that owns a unit ever writes that
//     unit's rows: handing this work to the other lanes of the warp (v4/v5) needs votes and shuffles, and those
//     were measured to run on a split warp on rare occasions (profiles/r1_notes.md).
// Every cached row is an exact copy of table entries (no arithmetic), so the state does not depend on the history of
// moves; dbn_verify_cache checks that bit for bit.
// Columns are kept in cache-slot order, not sorted by gene id: for mu = 0 (h, nh) the marginal likelihood and the
// Gibbs parameters are invariant to the column order (only rounding differs, ~1e-15 relative); traces are written
// in the reference's sorted order.
#pragma once
#include "dbn_chain.cuh"

namespace dbn {

// Block shape = register budget.  The step needs 236 registers per thread (255 with two rows in flight in the Gibbs sweep); 2 x 192 threads per SM (rounds 1 and 2a: 12 warps,
// every shape tried there kept ~12 warps) caps it at 168 and spills.  2 x 128 threads (8 warps per SM, no spills) is 16 % faster
// on the bench workload with one handle and 17 % with two (7.2e8 / 8.0e8 against 6.2e8 / 6.8e8 unit-iterations/s), and the
// end-to-end step gains 13 %: profiles/r2_small/block_shape_handles_ab2.txt (64 x 4, 96 x 2, 96 x 3, 112 x 2, 144 x 2, 160 x 2,
// 176 x 2, 224 x 2 measured beside it).
#ifndef DBN_FAST_BLOCK
#define DBN_FAST_BLOCK 128
#endif
#ifndef DBN_FAST_MINB
#define DBN_FAST_MINB 2
#endif
constexpr int FAST_BLOCK = DBN_FAST_BLOCK;
constexpr int SEG_ROWS = DBN_SMAX - 1;  // <= 9 segments
constexpr int BSLOTS = SEG_ROWS;        // cached rows per unit

template <int KM>
struct FastDims {
  static constexpr int KB = KM - 1;
  static constexpr int KT = KM * (KM + 1) / 2;
  static constexpr int NE = KT + KM;      // entries per cached row: ZZ packed by slot without (0, 0) | ZY by slot | YY
  static constexpr int NP = NE / 2;       // 16-byte pairs per cached row
  static constexpr int NC = KB + 2;       // candidate column at one boundary: cross terms with the base | diagonal | ZY
  static_assert(NE % 2 == 0, "a cached row is a whole number of 16-byte pairs (KM = 4: 14 entries, KM = 5: 20)");
};

// Cache layout: row r of unit u is NP pairs of doubles, pair p at doubles ((r * NP + p) * U + u) * 2: the two
// entries of a pair are adjacent, so a row moves as NP 16-byte copies per thread, consecutive threads 16 bytes apart.
// Entry ids: ZZ(i, j) (i >= j, not both 0) -> DBN_TRI(i, j) - 1; ZY(i) -> KT - 1 + i; YY -> KT - 1 + KM.
__host__ __device__ inline int fast_cache_row_doubles(int km) { return km * (km + 1) / 2 + km; }
// The candidate column of the parent-set sweep (<= 6 scattered 8-byte table reads per boundary, the line with 45 % of all
// long-scoreboard samples in profiles/r2_c/stalls.txt) is requested TWO steps ahead into two alternating slots; the cached rows
// stay one step ahead (deeper row staging was measured twice and lost).  1.2614 -> 1.2426 ms per step with two handles, 1.4188 ->
// 1.3815 with one (profiles/r2_small/gibbs_two_rows_in_flight.txt); 78 848 bytes of shared memory per block, still the 164 KB carve-out.
#ifndef DBN_CAND_SHALLOW
constexpr int FAST_CAND_SLOTS = 2;
#else
constexpr int FAST_CAND_SLOTS = 1;
#endif
inline size_t fast_smem_bytes(int km) {
  return (size_t)(3 * fast_cache_row_doubles(km) + FAST_CAND_SLOTS * (km + 1)) * FAST_BLOCK * sizeof(double) +
         (size_t)2 * DBN_SMAX * FAST_BLOCK * sizeof(unsigned short);
}
#define DBN_EZZ(i, j) (DBN_TRI(i, j) - 1)
// entry e of a row whose pair 0 is at `p` (global: pair stride = 2 * U doubles; shared staging: 2 * FAST_BLOCK doubles)
#define DBN_CE(p, e) (p)[(long long)((e) >> 1) * (2 * U) + ((e) & 1)]
#define DBN_SE(p, e) (p)[((e) >> 1) * (2 * FAST_BLOCK) + ((e) & 1)]

// 16-byte asynchronous global -> shared copy (LDGSTS): the staged row of the next boundary is in flight while the
// current segment is factorised; no registers are held for it.  .cg = through L2 only (the rows are streamed, never reused
// from L1).
__device__ __forceinline__ void cp_async16(double* smem, const double* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
// 8-byte variant (through L1: used for single prefix-table entries, which are read-only)
__device__ __forceinline__ void cp_async8(double* smem, const double* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// Block barrier that does NOT assume the warp arrives converged.  `__syncthreads()` is the .aligned form: ptxas emits
// the WARPSYNC.ALL in front of it only where its own analysis sees possible divergence, and in the h-dbn instantiations
// it saw none, although warps do reach the barrier in two parts (lanes that needed a second try in a rejection loop of
// a called function; measured with __activemask probes, profiles/r2_notes.md).  Each part then counted as a whole warp,
// the barrier opened before every warp had arrived, and the late warps ran one phase behind for the rest of the
// launch: their work counters were added after the block's totals had been published.  The non-aligned form counts
// threads and rejoins the warp.
__device__ __forceinline__ void block_barrier() {
#if defined(DBN_BARRIER_NONALIGNED)
  asm volatile("barrier.sync 0;" ::: "memory");
#elif defined(DBN_BARRIER_NONE)
#else
  __syncthreads();
#endif
}
// Rejoin the lanes of the warp.  Reconvergence after divergent code is opportunistic on this hardware (the
// BSYNC.RECONVERGENT ptxas places after a loop or a branch lets the lanes that arrive first run on), and ptxas drops a
// plain __syncwarp() wherever ITS analysis says the warp is converged.  Lanes whose segment count is smaller left the
// sweeps early and ran the rest of the iteration as a warp of their own: a third of all lanes were in split warps at
// every probe of the nh kernel (profiles/r2_notes.md), i.e. the same instructions were issued two or three times.  The
// sync is therefore placed under a condition that always holds but that the compiler cannot see through (`opaque` is a
// per-thread run-time value >= 0), which makes it keep the WARPSYNC.
#ifndef DBN_REJOIN
#define DBN_REJOIN 1   // 0: never, 1: in front of the block barriers, 2: also after the sweeps and the accepted-move updates
#endif
template <int LEVEL>
__device__ __forceinline__ void warp_rejoin(int opaque) {
  if (LEVEL <= DBN_REJOIN && opaque >= 0) __syncwarp();
}

#ifdef DBN_DIAG_SPLIT2   // diagnostic build: range checks, the largest violated code lands in counters[23]
#define DBN_CHECK(cond, code) do { if (!(cond)) atomicMax(&a.counters[23], (unsigned long long)(code)); } while (0)
#else
#define DBN_CHECK(cond, code) do { } while (0)
#endif

template <int METHOD, int KM, bool TRACE>
__global__ void __launch_bounds__(FAST_BLOCK, DBN_FAST_MINB) chain_fast_kernel(const ChainArgs a) {
  constexpr bool HOMOG = METHOD == DBN_H_DBN;
  constexpr int KB = FastDims<KM>::KB, KT = FastDims<KM>::KT, NP = FastDims<KM>::NP, NC = FastDims<KM>::NC;
  constexpr int KBT = KB * (KB + 1) / 2;
  constexpr int E_ZY = KT - 1, E_YY = KT - 1 + KM;  // entry ids of ZY(0) and YY

  // dynamic shared memory: [3][NP][BLOCK][2] doubles (boundary rows staged by cp.async) | [NC][BLOCK] candidate column
  // | tau | tau* (16-bit: the host refuses series with N + 2 > 65535 for this kernel)
  extern __shared__ __align__(16) double s_dyn[];
  double* const s_row = s_dyn;
  constexpr int BUF = (2 * NP) * FAST_BLOCK;  // doubles per staging buffer
  double* const s_cand = s_dyn + 3 * BUF;
  unsigned short* const s_tau = reinterpret_cast<unsigned short*>(s_cand + FAST_CAND_SLOTS * NC * FAST_BLOCK);
  unsigned short* const s_tas = s_tau + SMAX * FAST_BLOCK;
  const int tid = threadIdx.x;
  __shared__ unsigned long long s_tot[8];  // work counters of the block
  __shared__ unsigned int s_done;          // warps of the block that have added theirs
  if (tid < 8) s_tot[tid] = 0ull;
  if (tid == 8) s_done = 0u;
  __syncthreads();  // kernel entry: every warp is converged here
  const int u_raw = a.lanes >= 32 ? blockIdx.x * FAST_BLOCK + tid   // see dbn_chain.cuh
                                  : ((tid & 31) < a.lanes ? (blockIdx.x * (FAST_BLOCK / 32) + (tid >> 5)) * a.lanes + (tid & 31) : a.U);
  const bool active = u_raw < a.U;
  const int u = active ? u_raw : 0;
  const int node = (a.unit0 + u) / a.C;
  const int chain = (a.unit0 + u) - node * a.C;
  const int n = a.tb.n, N = a.tb.N, F = n - a.lag;   // n feature columns (genes x lag), the node's own copies are no candidates
  const long long U = a.U;
  auto TAU = [&](int j) -> unsigned short& { return s_tau[j * FAST_BLOCK + tid]; };
  auto TAS = [&](int j) -> unsigned short& { return s_tas[j * FAST_BLOCK + tid]; };
  auto tau_cur = [&](int j) { return (int)s_tau[j * FAST_BLOCK + tid]; };
  auto tau_star = [&](int j) { return (int)s_tas[j * FAST_BLOCK + tid]; };
  double* const sbuf = s_row + 2 * tid;  // this thread's slice of staging buffer 0; buffer b at sbuf + b * BUF
  double* const cand = s_cand + tid;     // this thread's candidate-column slot: value i at cand[i * FAST_BLOCK]
  double* const cbase = a.cache + 2ll * u;
  auto crow = [&](int row) -> double* { return cbase + (long long)row * NP * (2 * U); };  // pair 0 of cached row `row`
  const int nb = a.tb.zy_base + node * (n + 2);  // table offset of this node's ZY / YY block

  // ---- load state
  int npi = a.st_npi[u];
  int pi[PMAX];
#pragma unroll
  for (int j = 0; j < PMAX; ++j) pi[j] = a.st_pi[j * a.U + u];
  int S = active ? a.st_S[u] : 0;  // lanes past the end run the barriers of the loop on an empty state
  for (int j = 0; j < SMAX; ++j) TAU(j) = (unsigned short)a.st_tau[j * a.U + u];
  double sig2 = a.st_sig[u], lam = a.st_lam[u];
  int cg[KM];  // gene id held by column slot s (>= 1), -1 = empty; slot 0 is the intercept
  cg[0] = -1;
  auto bnd = [&](int j) { return DBN_SEG_END(tau_cur(j), j, S, N); };  // table row of the end boundary of segment j

  // request prefix-table row `t`, restricted to the slots' columns, into the staging buffer `dst` (8-byte cp.async per
  // entry; zeros are stored directly for empty slots).  The caller commits and waits.
  auto request_row = [&](int t, double* dst) {
    DBN_CHECK(t >= 0 && t <= N, 1);
    const double* __restrict__ row = a.tb.base + (long long)t * a.tb.row_stride;
#pragma unroll
    for (int i = 0; i < KM; ++i) {
      const bool oi = (i == 0) || cg[i] >= 0;
      const int fi = (i == 0) ? 0 : cg[i] + 1;
#pragma unroll
      for (int j = 0; j <= i; ++j) {
        if (i == 0) continue;  // ZZ(0, 0) = t is not stored
        const bool oj = (j == 0) || cg[j] >= 0;
        const int fj = (j == 0) ? 0 : cg[j] + 1;
        const int hi = max(fi, fj), lo = min(fi, fj);
        if (oi && oj) cp_async8(&DBN_SE(dst, DBN_EZZ(i, j)), row + hi * (hi + 1) / 2 + lo);
        else DBN_SE(dst, DBN_EZZ(i, j)) = 0.0;
      }
      if (oi) cp_async8(&DBN_SE(dst, E_ZY + i), row + nb + fi);
      else DBN_SE(dst, E_ZY + i) = 0.0;
    }
    cp_async8(&DBN_SE(dst, E_YY), row + nb + n + 1);
  };

  if (a.refresh) {
    // (re)build the cache from the table: canonical slots = parent positions, end boundary of segment j -> row j;
    // three rows per round trip through the three staging buffers, written out as the rows' 16-byte pairs
#pragma unroll
    for (int s = 1; s < KM; ++s) cg[s] = (s - 1 < npi) ? pi[s - 1] : -1;
    for (int j0 = 0; j0 < S; j0 += 3) {
#pragma unroll
      for (int b = 0; b < 3; ++b)
        if (j0 + b < S) request_row(bnd(j0 + b), sbuf + b * BUF);
      cp_async_commit();
      cp_async_wait<0>();
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const int j = j0 + b;
        if (j < S) {
          double2* cu = reinterpret_cast<double2*>(a.cache) + u;
          const double* src = sbuf + b * BUF;
#pragma unroll
          for (int p = 0; p < NP; ++p)
            cu[((long long)j * NP + p) * U] = *reinterpret_cast<const double2*>(src + p * (2 * FAST_BLOCK));
        }
      }
    }
  } else {
#pragma unroll
    for (int s = 1; s < KM; ++s) cg[s] = a.st_cg[(s - 1) * a.U + u];
  }

  // Philox counter = (global chain, node, iteration, slot); slots: Gibbs normals h (four per block) and 32 + h / 4 (the
  // fifth of four consecutive segments), discrete picks and accept uniforms 48..50, sigma^2 gamma 64.., lambda^2
  // gamma 256..
  const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
  const uint32_t gchain = a.chain_offset + (unsigned)chain, unode = (unsigned)node;
  Work wk = {0, 0, 0, 0};
  unsigned long long n_fact = 0;  // segment factorisations actually executed
  unsigned n_pacc = 0, n_tacc = 0;  // accepted parent-set / changepoint moves
  const bool replay = TRACE && a.rd != nullptr;

  // start the asynchronous copy of cached row j into staging buffer `buf` (the caller commits)
  auto issue_row = [&](int j, int buf) {
    DBN_CHECK(j >= 0 && j < BSLOTS && buf >= 0 && buf < 3, 2);
    const double* src = crow(j);
    double* dst = sbuf + buf * BUF;
#pragma unroll
    for (int p = 0; p < NP; ++p) cp_async16(dst + p * (2 * FAST_BLOCK), src + (long long)p * (2 * U));
  };
  // pivot product and quadratic form r^T C^-1 r (mu = 0, all KM slots) of the segment of `len` rows between the
  // boundary rows in the staging buffers `lo` (null: table row 0 = zeros) and `hi`
  auto seg_eval = [&](const double* lo, const double* hi, int len, double& prod, double& q) {
    double G[KT], g[KM], s, A[KT], dinv[KM], wv[KM];
    const bool hl = lo != nullptr;
    const double* l_ = hl ? lo : hi;
    G[0] = (double)len;
#pragma unroll
    for (int e = 1; e < KT; ++e) G[e] = DBN_SE(hi, e - 1) - (hl ? DBN_SE(l_, e - 1) : 0.0);
#pragma unroll
    for (int i = 0; i < KM; ++i) g[i] = DBN_SE(hi, E_ZY + i) - (hl ? DBN_SE(l_, E_ZY + i) : 0.0);
    s = DBN_SE(hi, E_YY) - (hl ? DBN_SE(l_, E_YY) : 0.0);
    form_a<KM>(G, lam, A);
    prod = ldl_factor<KM>(A, dinv);
    q = segment_q<KM, false>(G, g, s, g, lam, A, dinv, wv);
    ++n_fact;
  };

  // rows 0 and 1 of the first Gibbs sweep, one copy group each (every later iteration requests them at its end).  The Gibbs sweep
  // keeps TWO rows in flight (the third staging buffer is free until the proposals are drawn): one row ahead, a step waited
  // for most of an HBM round trip at its top (6.9 % of all stall samples on that line, profiles/r2_b/stalls.txt); measured
  // 1.271 -> 1.254 ms per step (profiles/r2_small/gibbs_two_rows_in_flight.txt).  The same depth in the parent-set sweep needs
  // a fourth row buffer and a second candidate slot (99 KB of shared memory per block) and was 19 % SLOWER.  What changes with it
  // is the carve-out (164 -> 228 KB, L1 92 -> 28 KB); the kernel's L1 hit rate is only 1.5 % (profiles/r2_b), so it is not cache
  // hits that are lost -- the mechanism was not investigated further.
  if (S > 0) issue_row(0, 0);
  cp_async_commit();
#ifndef DBN_GIBBS_SHALLOW
  if (S > 1) issue_row(1, 1);
  cp_async_commit();
#endif

  for (int li = 0; li < a.n_iter; ++li) {
    const int it = a.it0 + li;
    const uint32_t uit = (uint32_t)it;
    // The step is ~8k instructions of mostly straight-line code per thread, far more than the instruction cache
    // holds; warps that drift apart each stream it from L2 on their own (stall_no_instruction, profiles/r1_v4).
    // Block-wide barriers keep the block's warps on the same stretch of code so one fetch serves all of them.
    warp_rejoin<1>(S);
#ifdef DBN_DIAG_SPLIT2   // diagnostic build: lanes that reach probe k in a split warp -> counters[8 + k]
    if (__activemask() != 0xffffffffu) atomicAdd(&a.counters[8], 1ull);
#endif
    block_barrier();
    const double* rd = replay ? a.rd + ((long long)u * a.n_iter + li) * DBN_RD_STRIDE : nullptr;
    const int* ri = replay ? a.ri + ((long long)u * a.n_iter + li) * DBN_RI_STRIDE : nullptr;
    double* td = (TRACE && a.td && active) ? a.td + ((long long)u * a.n_iter + li) * DBN_TD_STRIDE : nullptr;
    int* ti = (TRACE && a.ti && active) ? a.ti + ((long long)u * a.n_iter + li) * DBN_TI_STRIDE : nullptr;
    if (TRACE && ti) ti[DBN_TI_S_BEFORE] = S;
    int k = npi + 1;
    DBN_CHECK(npi >= 0 && npi <= PMAX && S >= 0 && S <= SEG_ROWS, 8);
#pragma unroll
    for (int j = 0; j < PMAX; ++j) DBN_CHECK(j >= npi || (pi[j] >= 0 && pi[j] < n && pi[j] != node), 9);

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
    // Row h arrives in staging buffer h & 1; once a step has its operands in registers it requests row h + 1 into the
    // buffer of row h - 1.
    double q_sum = 0.0, acc_dd = 0.0, acc_de = 0.0, acc_ee = 0.0;
    {
      float4 z5 = make_float4(0.f, 0.f, 0.f, 0.f);  // KM = 5: the fifth normal of four consecutive segments
      int bprev = 0;
#ifndef DBN_GIBBS_SHALLOW
      int bi = 0;  // h mod 3: row h sits in staging buffer bi, row h - 1 in the one before it (cyclically)
#endif
      for (int h = 0; h < S; ++h) {
#ifndef DBN_GIBBS_SHALLOW
        cp_async_wait<1>();   // row h has landed; row h + 1 may still be on its way
        const int bp = (bi == 0) ? 2 : bi - 1;
        const double* sr = sbuf + bi * BUF;
        const double* sp = sbuf + bp * BUF;
#else
        cp_async_wait<0>();
        const double* sr = sbuf + (h & 1) * BUF;
        const double* sp = sbuf + ((h & 1) ^ 1) * BUF;
#endif
        const bool hp = h > 0;
        const int bcur = bnd(h);
        double G[KT], g[KM], s, A[KT], dinv[KM], rs[KM], wv[KM];
        G[0] = (double)(bcur - bprev);
        bprev = bcur;
#pragma unroll
        for (int e = 1; e < KT; ++e) G[e] = DBN_SE(sr, e - 1) - (hp ? DBN_SE(sp, e - 1) : 0.0);
#pragma unroll
        for (int i = 0; i < KM; ++i) g[i] = DBN_SE(sr, E_ZY + i) - (hp ? DBN_SE(sp, E_ZY + i) : 0.0);
        s = DBN_SE(sr, E_YY) - (hp ? DBN_SE(sp, E_YY) : 0.0);
#ifndef DBN_GIBBS_SHALLOW
        if (h + 2 < S) issue_row(h + 2, bp);   // into the buffer of row h - 1, whose values are in registers now
        bi = (bi == 2) ? 0 : bi + 1;
#else
        if (h + 1 < S) issue_row(h + 1, (h + 1) & 1);
#endif
        cp_async_commit();
        form_a<KM>(G, lam, A);
        ldl_factor_rs<KM>(A, dinv, rs);
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
          const float4 z4 = normal4_fast(gchain, unode, uit, (uint32_t)h, key);
          float zl = 0.f;
          if (KM == 5) {
            if ((h & 3) == 0) z5 = normal4_fast(gchain, unode, uit, 32u + (uint32_t)(h >> 2), key);
            const int w_ = h & 3;
            zl = (w_ == 0) ? z5.x : ((w_ == 1) ? z5.y : ((w_ == 2) ? z5.z : z5.w));
          }
          const float zf[5] = {z4.x, z4.y, z4.z, z4.w, zl};
#pragma unroll
          for (int i = 0; i < KM; ++i) e[i] = ((i == 0) || cg[i] >= 0) ? sw * (double)zf[i] * rs[i] : 0.0;
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
      }
    }
    warp_rejoin<2>(S);
#ifdef DBN_DIAG_SPLIT2
    if (__activemask() != 0xffffffffu) atomicAdd(&a.counters[9], 1ull);
#endif
    // row 0 of the parent-set sweep: on its way while sigma^2, lambda^2 and the proposals are drawn
    if (S > 0) issue_row(0, 0);
    cp_async_commit();
    // sigma^2 and lambda^2 draws (after the Gibbs sweep's sums; before anything that uses the new lambda^2)
    auto draw_sigma_lambda = [&]() {
    {
      const double shape = a.a_sig + a.T_sigma_half;
      const double rate = a.b_sig + 0.5 * q_sum;
      sig2 = replay ? rd[DBN_RD_SIGMA2] : rate / gamma_draw_fast(gchain, unode, uit, 64u, key, shape);
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
      lam = replay ? rd[DBN_RD_LAMBDA2] : rate / gamma_draw_fast(gchain, unode, uit, 256u, key, shape);
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
    };
#ifndef DBN_LATE_DRAWS
    draw_sigma_lambda();
#endif

#ifndef DBN_NO_MID_BAR
    warp_rejoin<1>(S);
    block_barrier();  // the Gibbs sweep's trip count differs per lane: regroup the block's warps (measured +1.3 %)
#endif
#ifdef DBN_DIAG_SPLIT2
    if (__activemask() != 0xffffffffu) atomicAdd(&a.counters[12], 1ull);
#endif
    // ============================ proposals (no scoring yet) ==================================================
    // all discrete picks and both accept uniforms of the iteration come from three Philox blocks
    uint4 wa = make_uint4(0, 0, 0, 0), wb = wa, wc = wa;
    if (!replay) {
      wa = philox_block(gchain, unode, uit, 48u, key);
      wb = philox_block(gchain, unode, uit, 49u, key);
      wc = philox_block(gchain, unode, uit, 50u, key);
#ifdef DBN_DIAG_SPLIT2
    if (__activemask() != 0xffffffffu) atomicAdd(&a.counters[13], 1ull);
#endif
    }
    auto pick = [](uint32_t word, int m) { return (int)__umulhi(word, (uint32_t)m); };
    auto unif = [](uint32_t hi, uint32_t lo) { return (double)((((uint64_t)hi << 32) | lo) >> 11) * 0x1.0p-53; };
    // ---- parent-set proposal: addMove / deleteMove / exchangeMove (utils.py:271-323)
    const int pmove = replay ? ri[DBN_RI_PI_MOVE] : pick(wa.x, 3);
    bool pvalid = active;
    int star[PMAX], nstar = npi;
    int el = -1, cnew = -1;  // gene leaving / entering the parent set
    int hr_num = 1, hr_den = 1;  // Hastings ratio of the parent-set move (moves.py:612, :615)
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
          const int p = replay ? ri[DBN_RI_PI_PICK] : pick(wa.y, ncand);
          cnew = nth_candidate(p, node, spi, npi, a.n_genes, a.lag);
#pragma unroll
          for (int j = 0; j < PMAX; ++j)
            if (j == npi) star[j] = cnew;
          nstar = npi + 1;
          hr_num = F - npi, hr_den = nstar;  // moves.py:612
        }
      } else if (npi < 1) {
        pvalid = false;  // utils.py:285-286, :308-309
      } else {
        const int idx = replay ? ri[DBN_RI_PI_PICK] : pick(wa.y, npi);
        el = pi[0];
#pragma unroll
        for (int j = 1; j < PMAX; ++j)
          if (j == idx) el = pi[j];
        int w_ = 0;  // sorted(pi \ el)  (np.setdiff1d, utils.py:290, :318)
#pragma unroll
        for (int j = 0; j < PMAX; ++j) {  // selects only: a switch on w_ here becomes an indirect branch (BRX)
          const bool keep = (j < npi) && (spi[j] != el);
#pragma unroll
          for (int q = 0; q < PMAX; ++q) star[q] = (keep && q == w_) ? spi[j] : star[q];
          w_ += keep ? 1 : 0;
        }
        if (pmove == 1) {
          nstar = npi - 1;
          hr_num = npi, hr_den = F - nstar;  // moves.py:615
        } else {
          const int ncand = F - npi;
          if (ncand <= 0) {
            pvalid = false;
          } else {
            const int p = replay ? ri[DBN_RI_PI_PICK + 1] : pick(wa.z, ncand);
            cnew = nth_candidate(p, node, spi, npi, a.n_genes, a.lag);
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
    int t_num = 1, t_den = 1;  // Hastings ratio of the changepoint move (moves.py:159, :163)
    if (do_tau) {
      tmove = replay ? ri[DBN_RI_TAU_MOVE] : pick(wa.w, 3);
      const int ns = a.num_samples;
      tvalid = active;
      if (tmove == 0) {
        int inside = 0;
        for (int j = 0; j < S - 1; ++j) inside += (TAU(j) >= 2 && TAU(j) <= ns) ? 1 : 0;
        const int ncand = ns - 1 - inside;
        if (ncand <= 0) {
          tvalid = false;
        } else {
          const int p = replay ? ri[DBN_RI_TAU_PICK] : pick(wb.x, ncand);
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
          t_num = ns - S, t_den = Ss - 1;  // moves.py:159
        }
      } else if (S < 2) {
        tvalid = false;  // changepointMoves.py:81-82, :21-22
      } else if (tmove == 1) {
        const int idx = replay ? ri[DBN_RI_TAU_PICK] : pick(wb.x, S - 1);
        int w_ = 0;
        for (int j = 0; j < S; ++j)
          if (j != idx) TAS(w_++) = TAU(j);
        tj = idx + 1;
        Ss = S - 1;
        t_num = S - 1, t_den = ns - Ss;  // moves.py:163
      } else {
        const int idx = replay ? ri[DBN_RI_TAU_PICK] : pick(wb.x, S - 1);
        const int last = TAU(S - 1) - 1;
        const int left = (idx == 0) ? 1 : TAU(idx - 1);
        const int right = (idx + 1 == S - 1) ? last : TAU(idx + 1);
        const int ncand = right - left - 2;
        if (ncand <= 0) {
          tvalid = false;
        } else {
          const int p = replay ? ri[DBN_RI_TAU_PICK + 1] : pick(wb.y, ncand);
          int c = left + 1 + p;
          if (c >= TAU(idx)) ++c;
          for (int j = 0; j < S; ++j) TAS(j) = (j == idx) ? c : TAU(j);
          tj = idx + 1;
          trow = c - 1;
          if (trow >= N) tvalid = false;
        }
      }
    }
    warp_rejoin<2>(S);
#ifdef DBN_DIAG_SPLIT2
    if (__activemask() != 0xffffffffu) atomicAdd(&a.counters[14], 1ull);
#endif
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
    const bool birth = tmove == 0, death = tmove == 1;
    const bool t_new = tvalid && !death;  // the changepoint proposal has a new boundary row (table row `trow`)
    double* const b0 = sbuf;
    double* const b1 = sbuf + BUF;
    double* const b2 = sbuf + 2 * BUF;

#ifdef DBN_BAR_PI
    warp_rejoin<1>(S);
    block_barrier();
#endif
    // ============================ parent-set move: pi and pi* in one sweep (a4) ===============================
    // position p of the sweep holds slot perm[p]; the last position is the slot that leaves (delete / exchange) or
    // the free slot that would be filled (add).  current = base + last position, star = base + new column.
    double ld_now = 0.0, q_now = 0.0, ml_now = NAN;
    bool p_acc = false;
    int tslot = KM - 1;
    int et[KB], o2[KB], ett = 0, ezyt = 0, o2d = -1, o2y = -1;  // cache entries / table offsets of the changing column
#pragma unroll
    for (int p = 0; p < KB; ++p) et[p] = o2[p] = 0;
    if (pvalid || tvalid) {
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
      int eb[KBT], ezy[KB];
      const int cf = cnew + 1;
#pragma unroll
      for (int p = 0; p < KB; ++p) {
#pragma unroll
        for (int q = 0; q <= p; ++q) eb[DBN_TRI(p, q)] = DBN_EZZ(perm[p], perm[q]);  // perm is increasing on the base; (0, 0) unused
        et[p] = (tslot > perm[p]) ? DBN_EZZ(tslot, perm[p]) : DBN_EZZ(perm[p], tslot);
        ezy[p] = E_ZY + perm[p];
        // perm[p] is p or p + 1: selects keep cg in registers (a run-time index sends the array to local memory)
        const int cgp = (p == 0) ? -1 : ((p < tslot) ? cg[p] : cg[p + 1 < KM ? p + 1 : p]);
        const int fp = (p == 0) ? 0 : cgp + 1;  // 0 when the slot is empty; masked by o2 < 0
        const bool occ = (p == 0) || cgp >= 0;
        const int hi = max(cf, fp), lo = min(cf, fp);
        o2[p] = (has_new && occ) ? hi * (hi + 1) / 2 + lo : -1;
      }
      ezyt = E_ZY + tslot;
      ett = DBN_EZZ(tslot, tslot);
      o2d = has_new ? cf * (cf + 1) / 2 + cf : -1;
      o2y = has_new ? nb + cf : -1;
      // request the candidate column (<= KM + 1 table entries) at table row t into the slot `cand`
      auto request_cand = [&](int t, int slot) {
        double* const cand = s_cand + tid + slot * (NC * FAST_BLOCK);
        DBN_CHECK(t >= 0 && t <= N, 3);
        DBN_CHECK(o2d >= 0 && o2d < a.tb.zz_size && o2y >= a.tb.zz_size && o2y < a.tb.row_stride, 4);
#pragma unroll
        for (int p = 0; p < KB; ++p) DBN_CHECK(o2[p] < a.tb.zz_size, 5);
        const double* __restrict__ row = a.tb.base + (long long)t * a.tb.row_stride;
#pragma unroll
        for (int p = 0; p < KB; ++p)
          if (o2[p] >= 0) cp_async8(cand + p * FAST_BLOCK, row + o2[p]);
        cp_async8(cand + KB * FAST_BLOCK, row + o2d);
        cp_async8(cand + (KB + 1) * FAST_BLOCK, row + o2y);
      };
      // with row 0 (requested after the Gibbs sweep): the candidate column at the first boundary and the new boundary
      // row of the changepoint proposal (into the third staging buffer, where it stays until the changepoint move)
      if (has_new && S > 0) request_cand(bnd(0), 0);
      if (t_new) request_row(trow, b2);
      cp_async_commit();
#ifndef DBN_CAND_SHALLOW
      if (has_new && S > 1) request_cand(bnd(1), 1);
      cp_async_commit();
#endif
#ifdef DBN_LATE_DRAWS   // the scattered requests above are in flight while the two gamma variates are drawn
      draw_sigma_lambda();
#endif

      // ln det of the base blocks: the pivot products are accumulated as mantissa x 2^exponent (one log per sweep)
      double mB = 1.0, R1 = 1.0, R2 = 1.0, q1 = 0.0, q2 = 0.0;
      int eB = 0;
      double n_prev[KB + 2];  // new column at the previous boundary: cross terms, diagonal, ZY
#pragma unroll
      for (int i = 0; i < KB + 2; ++i) n_prev[i] = 0.0;
      int bprev = 0;
      for (int h = 0; h < S; ++h) {
#ifndef DBN_CAND_SHALLOW
        cp_async_wait<1>();   // row h and candidate column h have landed; candidate column h + 1 (the newest group) may not have
        const double* const cand = s_cand + tid + (h & 1) * (NC * FAST_BLOCK);
#else
        cp_async_wait<0>();
#endif
        const double* sr = sbuf + (h & 1) * BUF;
        const double* sp = sbuf + ((h & 1) ^ 1) * BUF;
        const bool hp = h > 0;
        const int bcur = bnd(h);
        auto seg = [&](int e) { return DBN_SE(sr, e) - (hp ? DBN_SE(sp, e) : 0.0); };  // entry e of this segment's statistics
        double A[KBT], gB[KB], t1[KB], t2[KB], dinv[KB];
        A[0] = (double)(bcur - bprev);
        bprev = bcur;
#pragma unroll
        for (int e = 1; e < KBT; ++e) A[e] = seg(eb[e]);
#pragma unroll
        for (int p = 0; p < KB; ++p) {
          gB[p] = seg(ezy[p]);
          t1[p] = seg(et[p]);
        }
        const double t1d = seg(ett), t1y = seg(ezyt);
        const double s = seg(E_YY);
        double t2d = 0.0, t2y = 0.0;
        if (has_new) {
          double cur[KB + 2];
#pragma unroll
          for (int p = 0; p < KB; ++p) cur[p] = (o2[p] >= 0) ? cand[p * FAST_BLOCK] : 0.0;
          cur[KB] = cand[KB * FAST_BLOCK];
          cur[KB + 1] = cand[(KB + 1) * FAST_BLOCK];
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
        // this step's operands are in registers: request the next step's (row h + 1 into the buffer of row h - 1, the
        // candidate column at boundary h + 1 into its slot)
#ifndef DBN_CAND_SHALLOW
        // two groups per step, the row first: wait<1> at the top of step h + 1 then covers row h + 1 and every older group
        // (candidate column h + 1 among them) and leaves only the newest one, candidate column h + 2, in flight
        if (h + 1 < S) issue_row(h + 1, (h + 1) & 1);
        cp_async_commit();
        if (has_new && h + 2 < S) request_cand(bnd(h + 2), h & 1);   // this step's slot is free: its values are in registers
        cp_async_commit();
#else
        if (h + 1 < S) {
          issue_row(h + 1, (h + 1) & 1);
          if (has_new) request_cand(bnd(h + 1), 0);
        }
        cp_async_commit();
#endif
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
          contrib = w * w * pivot_rcp(d);
        };
        double d1, c1, d2, c2;
        tail(t1, t1d, t1y, d1, c1);
        tail(t2, t2d, t2y, d2, c2);
        {
          int ex;
          mB = frexp(mB * prodB, &ex);
          eB += ex;
        }
        R1 *= d1;
        R2 *= d2;
        q1 += s - lam * (accB + c1);
        q2 += s - lam * (accB + c2);
      }
      const double ldB = (double)eB * 0.69314718055994530942;
      const double ld_cur = ldB + dlog(mB * R1), ld_star = ldB + dlog(mB * R2);
      const double ml_cur = a.c0 - 0.5 * ld_cur - a.T_half_plus_a * dlog(a.two_b + q1);  // moves.py:599-600 / :681
      wk.eval();
      for (int h = 0; h < S; ++h) wk.segment(k, false);
      ld_now = ld_cur;
      q_now = q1;
      ml_now = ml_cur;
      if (pvalid) {
        const int ks = nstar + 1;
        const double ml_star = a.c0 - 0.5 * ld_star - a.T_half_plus_a * dlog(a.two_b + q2);  // moves.py:574-576 / :672
        wk.eval();
        for (int h = 0; h < S; ++h) wk.segment(ks, false);
        const double log_ratio = ml_star - ml_cur + dlog((double)hr_num / (double)hr_den);  // uniform parent-set prior cancels (priors.py:48-68)
        const double uacc = replay ? rd[DBN_RD_U_PI] : unif(wb.z, wb.w);
        p_acc = uacc < fmin(1.0, dexp(log_ratio));  // moves.py:620-633, :699-702
        if (TRACE && td) {
          td[DBN_TD_LOGML + 0] = ml_star;
          td[DBN_TD_LOGML + 1] = ml_cur;
        }
        if (TRACE && ti) ti[DBN_TI_PI_ACCEPT] = p_acc;
        if (p_acc) {
          npi = nstar;
          k = ks;
#pragma unroll
          for (int j = 0; j < PMAX; ++j) pi[j] = star[j];
          ld_now = ld_star;
          q_now = q2;
          ml_now = ml_star;
#pragma unroll
          for (int s = 1; s < KM; ++s)
            if (s == tslot) cg[s] = has_new ? cnew : -1;
          ++n_pacc;
        }
      }
    } else {
#ifdef DBN_LATE_DRAWS
      draw_sigma_lambda();
#endif
      cp_async_wait<0>();  // row 0 was requested for a sweep that does not run
    }
    warp_rejoin<2>(S);
#ifdef DBN_DIAG_SPLIT2
    if (__activemask() != 0xffffffffu) atomicAdd(&a.counters[16], 1ull);
#endif
    // ---- accepted parent-set move: the new column (zeros for a delete) takes slot `tslot` in every cached row and in
    // the changepoint proposal's new boundary row (staging buffer 2).  Its table entries at those <= S + 1 boundaries
    // are fetched by cp.async into the two row buffers (free between the sweeps), RB boundaries per round trip.
    if (p_acc) {
      DBN_CHECK(tslot >= 1 && tslot < KM && ett >= 0 && ett < FastDims<KM>::NE && ezyt >= 0 && ezyt < FastDims<KM>::NE, 6);
#pragma unroll
      for (int p = 0; p < KB; ++p) DBN_CHECK(et[p] >= 0 && et[p] < FastDims<KM>::NE, 7);
      constexpr int RB = (2 * 2 * NP) / NC;  // boundaries per round: value v = j * NC + p lives in buffers 0 / 1
      auto stash = [&](int v) -> double* {
        const int b_ = v / (2 * NP), w_ = v - b_ * (2 * NP);
        return sbuf + b_ * BUF + (w_ >> 1) * (2 * FAST_BLOCK) + (w_ & 1);
      };
      const bool fetch = o2d >= 0;
      const int nbnd = S + (t_new ? 1 : 0);
      for (int j0 = 0; j0 < nbnd; j0 += RB) {
        const int j1 = min(j0 + RB, nbnd);
        if (fetch) {
          for (int j = j0; j < j1; ++j) {
            const double* __restrict__ row = a.tb.base + (long long)((j < S) ? bnd(j) : trow) * a.tb.row_stride;
            const int v0 = (j - j0) * NC;
#pragma unroll
            for (int p = 0; p < KB; ++p)
              if (o2[p] >= 0) cp_async8(stash(v0 + p), row + o2[p]);
            cp_async8(stash(v0 + KB), row + o2d);
            cp_async8(stash(v0 + KB + 1), row + o2y);
          }
        }
        cp_async_commit();
        cp_async_wait<0>();
        for (int j = j0; j < j1; ++j) {
          const int v0 = (j - j0) * NC;
          if (j < S) {
            double* c = crow(j);
#pragma unroll
            for (int p = 0; p < KB; ++p) DBN_CE(c, et[p]) = (fetch && o2[p] >= 0) ? *stash(v0 + p) : 0.0;
            DBN_CE(c, ett) = fetch ? *stash(v0 + KB) : 0.0;
            DBN_CE(c, ezyt) = fetch ? *stash(v0 + KB + 1) : 0.0;
          } else {
#pragma unroll
            for (int p = 0; p < KB; ++p) DBN_SE(b2, et[p]) = (fetch && o2[p] >= 0) ? *stash(v0 + p) : 0.0;
            DBN_SE(b2, ett) = fetch ? *stash(v0 + KB) : 0.0;
            DBN_SE(b2, ezyt) = fetch ? *stash(v0 + KB + 1) : 0.0;
          }
        }
      }
    }
    warp_rejoin<2>(S);
#ifdef DBN_DIAG_SPLIT2
    if (__activemask() != 0xffffffffu) atomicAdd(&a.counters[17], 1ull);
#endif
    // ---- counts: pi_vector[it+1] is kept iff idx >= burn_in and (idx - burn_in) % thin != 0 (network.py:358-359)
    {
      const int idx = it + 1;
      if (active && idx >= a.burn_in && ((idx - a.burn_in) % a.thin) != 0 && npi > 0) {
        atomicAdd(&a.nonempty[node], 1ull);
        count_edges(a.edge + (long long)node * a.n_genes, pi, npi, a.n_genes);
      }
      if (TRACE && ti) {
        ti[DBN_TI_NPI] = npi;
        for (int j = 0; j < PMAX; ++j) ti[DBN_TI_PI + j] = (j < npi) ? pi[j] : -1;
      }
    }

#ifdef DBN_BAR_TAU
    warp_rejoin<1>(S);
    block_barrier();
#endif
    // ============================ changepoint move: only the changed segments (a5) ============================
    if (do_tau) {
      bool t_acc = false;
      const int ja = tj - 1;  // cached row the move inserts before (birth), removes (death) or replaces (reallocation)
      if (tvalid) {
        // The boundaries around the change: left = row ja-1 (table row 0 when ja == 0), mid = row ja (death,
        // reallocation), right = row ja (birth) or ja+1; they are cached rows.  The new boundary (birth, reallocation)
        // is the only table row that had to be gathered: it has been in staging buffer 2 since the parent-set sweep.
        //   old segments: (left, mid), (mid, right)   or, for a birth, (left, right)
        //   new segments: (left, new), (new, right)   or, for a death, (left, right)
        // Buffers: left -> 0, mid -> 1, new -> 2; right -> 1 (birth), 2 (death), or, for a reallocation, 0 once the
        // two segments that start at the left boundary are done.
        const bool has_l = ja >= 1;
        const bool realloc = !birth && !death;
        auto stage = [&](int row, double* dst) {
          const double* src = crow(row);
#pragma unroll
          for (int p = 0; p < NP; ++p) cp_async16(dst + p * (2 * FAST_BLOCK), src + (long long)p * (2 * U));
        };
        if (has_l) stage(ja - 1, b0);
        stage(ja, b1);
        if (death) stage(ja + 1, b2);
        cp_async_commit();
        const int t_l = has_l ? bnd(ja - 1) : 0;
        const int t_m = bnd(ja);                              // birth: this is the right boundary
        const int t_r = birth ? t_m : bnd(ja + 1);
        cp_async_wait<0>();
        const double* lo0 = has_l ? b0 : nullptr;
        const double* right = birth ? b1 : (death ? b2 : b0);
        double prod_old = 1.0, q_old = 0.0, prod_new = 1.0, q_new = 0.0;
        // steps 0, 1: the segments that start at the left boundary (old, new); 2, 3: those that end at the right one
        for (int step = 0; step < 4; ++step) {
          const bool is_new = step & 1;
          if (step >= 2 && (is_new ? death : birth)) continue;  // that side has one segment only
          if (step == 2 && realloc) {
            stage(ja + 1, b0);
            cp_async_commit();
            cp_async_wait<0>();
          }
          const double* lo;
          const double* hi;
          int len;
          if (step < 2) {
            lo = lo0;
            const bool to_right = is_new ? death : birth;     // the only segment of its side spans left .. right
            hi = to_right ? right : (is_new ? b2 : b1);
            len = (to_right ? t_r : (is_new ? trow : t_m)) - t_l;
          } else {
            lo = is_new ? b2 : b1;
            hi = right;
            len = t_r - (is_new ? trow : t_m);
          }
          double pr, qq;
          seg_eval(lo, hi, len, pr, qq);
          if (is_new) {
            prod_new *= pr;
            q_new += qq;
          } else {
            prod_old *= pr;
            q_old += qq;
          }
        }
        const double ld_star = ld_now + dlog(prod_new / prod_old);
        const double q_star = q_now - q_old + q_new;
        const double ml_star = a.c0 - 0.5 * ld_star - a.T_half_plus_a * dlog(a.two_b + q_star);  // moves.py:209-216
        wk.eval();
        for (int h = 0; h < Ss; ++h) wk.segment(k, false);
        const double lprior = log_cp_prior(tau_star, Ss, a.log_p, a.log_1mp) - log_cp_prior(tau_cur, S, a.log_p, a.log_1mp);
        const double log_ratio = ml_star - ml_now + lprior + dlog((double)t_num / (double)t_den);  // moves.py:225
        const double uacc = replay ? rd[DBN_RD_U_TAU] : unif(wc.x, wc.y);
        t_acc = uacc < fmin(1.0, dexp(log_ratio));
        if (TRACE && td) {
          td[DBN_TD_LOGML + 2] = ml_now;
          td[DBN_TD_LOGML + 3] = ml_star;
        }
        if (TRACE && ti) ti[DBN_TI_TAU_ACCEPT] = t_acc;
      }
      // ---- accepted changepoint move: a birth inserts the new boundary's row before row ja (rows ja .. S-1 move up by
      // one), a death removes row ja (rows ja+1 .. S-1 move down by one), a reallocation replaces row ja.  The new
      // boundary's row is in staging buffer 2.  A row moves as NP 16-byte pairs.
      if (t_acc) {
        double2* cu = reinterpret_cast<double2*>(a.cache) + u;  // pair p of row r at cu[(r * NP + p) * U]
        auto move_row = [&](int dst, int src) {
          double2 t[NP];
#pragma unroll
          for (int p = 0; p < NP; ++p) t[p] = __ldcg(cu + ((long long)src * NP + p) * U);
#pragma unroll
          for (int p = 0; p < NP; ++p) cu[((long long)dst * NP + p) * U] = t[p];
        };
        // two rows per round trip (both loaded before either is stored; a pair never overlaps its own targets)
        auto move_two = [&](int src_a, int src_b, int dlt_) {
          double2 ta[NP], tb_[NP];
#pragma unroll
          for (int p = 0; p < NP; ++p) ta[p] = __ldcg(cu + ((long long)src_a * NP + p) * U);
#pragma unroll
          for (int p = 0; p < NP; ++p) tb_[p] = __ldcg(cu + ((long long)src_b * NP + p) * U);
#pragma unroll
          for (int p = 0; p < NP; ++p) cu[((long long)(src_a + dlt_) * NP + p) * U] = ta[p];
#pragma unroll
          for (int p = 0; p < NP; ++p) cu[((long long)(src_b + dlt_) * NP + p) * U] = tb_[p];
        };
        if (birth) {
          int r = S - 1;
          for (; r - 1 >= ja; r -= 2) move_two(r, r - 1, 1);
          if (r >= ja) move_row(r + 1, r);
        } else if (death) {
          int r = ja + 1;
          for (; r + 1 < S; r += 2) move_two(r, r + 1, -1);
          if (r < S) move_row(r - 1, r);
        }
        if (!death) {
#pragma unroll
          for (int p = 0; p < NP; ++p)
            cu[((long long)ja * NP + p) * U] = *reinterpret_cast<const double2*>(b2 + p * (2 * FAST_BLOCK));
        }
      }
      if (t_acc) {
        S = Ss;
        for (int j = 0; j < Ss; ++j) TAU(j) = TAS(j);
        ++n_tacc;
      }
      // ---- counts: tau_vector[it] is kept iff it >= burn_in and (it - burn_in) % thin == 0 (network.py:388-389)
      if (active && it >= a.burn_in && ((it - a.burn_in) % a.thin) == 0) {
        atomicAdd(&a.cp_kept[node], 1ull);
        for (int j = 0; j < S - 1; ++j) atomicAdd(&a.cp_hist[(long long)node * (N + 3) + TAU(j)], 1ull);
      }
    }
    if (TRACE && ti) {
      ti[DBN_TI_S] = S;
      for (int j = 0; j < SMAX; ++j) ti[DBN_TI_TAU + j] = (j < S) ? TAU(j) : -1;
    }
    // row 0 of the next iteration's Gibbs sweep (after this iteration's own updates of the cached rows)
    if (S > 0 && li + 1 < a.n_iter) issue_row(0, 0);
    cp_async_commit();
#ifndef DBN_GIBBS_SHALLOW
    if (S > 1 && li + 1 < a.n_iter) issue_row(1, 1);
    cp_async_commit();
#endif
  }
  cp_async_wait<0>();

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
  }
  // ---- work counters: summed per warp with shuffles (behind a WARPSYNC that the compiler cannot drop, see warp_rejoin),
  // then per block with one shared-memory atomic per counter and warp; the warp that finishes LAST publishes the
  // block's totals (one global atomic per counter and block).  No block barrier here: the loop's barriers are a
  // scheduling aid only, and a warp that ran them in two parts may be a phase behind the rest of its block.
  // (192 threads x 8 64-bit shared-memory atomics on 8 addresses were 5.8 % of the launch, profiles/r2_a.)
  unsigned long long v[8] = {active ? (unsigned long long)a.n_iter : 0ull, wk.evals, wk.segs, wk.bytes, wk.flops,
                             (unsigned long long)n_pacc, n_fact, (unsigned long long)n_tacc};
  warp_rejoin<0>(S);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v[c] += __shfl_xor_sync(0xffffffffu, v[c], off);
  }
  if ((tid & 31) == 0) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (v[c]) atomicAdd(&s_tot[c], v[c]);
    __threadfence_block();
    if (atomicAdd(&s_done, 1u) + 1u == (unsigned)(FAST_BLOCK / 32)) {
      __threadfence_block();
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const unsigned long long t = *(volatile unsigned long long*)&s_tot[c];
        if (t) atomicAdd(&a.counters[c], t);
      }
    }
  }
}

// Invariant check (test hook behind dbn_verify_cache): every cached row of every unit must equal the prefix-table row
// at the end boundary of the corresponding segment, restricted to the unit's column slots, bit for bit.
// counts[0] += units with at least one differing entry, counts[1] += differing entries.
// (Re)build the boundary-row cache from the table with one thread per (unit, cached row): canonical slots = parent positions,
// end boundary of segment j -> row j.  The chain kernel can do this itself (`refresh`), three rows per round trip of its 384
// threads per SM; as a kernel of its own every row of every unit is in flight at once (the end-to-end step rebuilds table and
// cache for every new series: 0.33 ms of its 2.2 ms were this gather, scripts/e2e_breakdown.py).
template <int KM>
__global__ void __launch_bounds__(256) cache_fill_kernel(const ChainArgs a) {
  constexpr int KT = FastDims<KM>::KT, NP = FastDims<KM>::NP;
  const long long i = blockIdx.x * 256ll + threadIdx.x;
  const long long U = a.U;
  if (i >= U * SEG_ROWS) return;
  const int j = (int)(i / U), u = (int)(i - (long long)j * U);
  const int S = a.st_S[u];
  const int npi = a.st_npi[u];
  int cg[KM];
  cg[0] = -1;
#pragma unroll
  for (int s = 1; s < KM; ++s) cg[s] = (s - 1 < npi) ? a.st_pi[(s - 1) * a.U + u] : -1;
  if (j == 0) {
#pragma unroll
    for (int s = 1; s < KM; ++s) a.st_cg[(s - 1) * a.U + u] = cg[s];
  }
  if (j >= S) return;
  const int node = (a.unit0 + u) / a.C, n = a.tb.n, N = a.tb.N;
  const int nb = a.tb.zy_base + node * (n + 2);
  const int t = DBN_SEG_END(a.st_tau[j * a.U + u], j, S, N);
  const double* __restrict__ row = a.tb.base + (long long)t * a.tb.row_stride;
  double v[2 * NP];
#pragma unroll
  for (int q = 0; q < KM; ++q) {
    const bool oi = (q == 0) || cg[q] >= 0;
    const int fi = (q == 0) ? 0 : cg[q] + 1;
#pragma unroll
    for (int jj = 0; jj <= q; ++jj) {
      if (q == 0) continue;   // ZZ(0, 0) = t is not stored
      const bool oj = (jj == 0) || cg[jj] >= 0;
      const int fj = (jj == 0) ? 0 : cg[jj] + 1;
      const int hi = max(fi, fj), lo = min(fi, fj);
      v[DBN_EZZ(q, jj)] = (oi && oj) ? __ldg(row + hi * (hi + 1) / 2 + lo) : 0.0;
    }
    v[KT - 1 + q] = oi ? __ldg(row + nb + fi) : 0.0;
  }
  v[KT - 1 + KM] = __ldg(row + nb + n + 1);
  double2* cu = reinterpret_cast<double2*>(a.cache) + u;
#pragma unroll
  for (int p = 0; p < NP; ++p) cu[((long long)j * NP + p) * U] = make_double2(v[2 * p], v[2 * p + 1]);
}

template <int KM>
__global__ void cache_verify_kernel(const ChainArgs a, unsigned long long* counts) {
  constexpr int KT = FastDims<KM>::KT, NP = FastDims<KM>::NP;
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= a.U) return;
  const long long U = a.U;
  const int node = (a.unit0 + u) / a.C, n = a.tb.n, N = a.tb.N;
  const int nb = a.tb.zy_base + node * (n + 2);
  const int S = a.st_S[u];
  int cg[KM];
  cg[0] = -1;
  for (int s = 1; s < KM; ++s) cg[s] = a.st_cg[(s - 1) * a.U + u];
  unsigned long long bad = 0;
  for (int j = 0; j < S; ++j) {
    const int t = DBN_SEG_END(a.st_tau[j * a.U + u], j, S, N);
    const double* row = a.tb.base + (long long)t * a.tb.row_stride;
    const double* c = a.cache + 2ll * u + (long long)j * NP * (2 * U);
    bad += (row[0] != (double)t) ? 1 : 0;  // the entry the cache does not store: ZZ(0, 0) at boundary t is t
    for (int i = 0; i < KM; ++i) {
      const bool oi = (i == 0) || cg[i] >= 0;
      const int fi = (i == 0) ? 0 : cg[i] + 1;
      for (int jj = 0; jj <= i; ++jj) {
        if (i == 0) continue;
        const bool oj = (jj == 0) || cg[jj] >= 0;
        const int fj = (jj == 0) ? 0 : cg[jj] + 1;
        const int hi = max(fi, fj), lo = min(fi, fj), off = hi * (hi + 1) / 2 + lo;
        const double ex = (oi && oj) ? row[off] : 0.0;
        bad += (ex != DBN_CE(c, DBN_EZZ(i, jj))) ? 1 : 0;
      }
      const double ex = oi ? row[nb + fi] : 0.0;
      bad += (ex != DBN_CE(c, KT - 1 + i)) ? 1 : 0;
    }
    bad += (row[nb + n + 1] != DBN_CE(c, KT - 1 + KM)) ? 1 : 0;
  }
  if (bad) {
    atomicAdd(&counts[0], 1ull);
    atomicAdd(&counts[1], bad);
  }
}

}  // namespace dbn
