// Kernels 1 and 2 (prefix statistics, batched scorer), the FP64 peak probe, and the C ABI of include/dbn_b200.h.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>
#include <algorithm>
#include <cstdlib>

#include "dbn_chain_fast.cuh"
#include "dbn_fp.cuh"
#include "dbn_launch.h"

namespace dbn {

// =====================================================================================================
// Kernel 1 — time-prefix sufficient statistics (replaces utils.py:128-269 re-slicing)
// =====================================================================================================
// Every table column p is the running sum over regression rows of a product of two entries of
//   v_t = [1, x_t(0..n-1), y_t(0..n-1)]        (x_t = data row t, y_t = data row t+1 within a series segment)
// colA[p], colB[p] name the two entries.  The table is written once, row-major in t, so the kernel is a pure
// HBM write stream: 8 * (N+1) * row_doubles bytes out, 16 * N * n bytes in.

constexpr int PREFIX_BLOCK = 256;

// gather the lagged rows: V[t] = [1, x_t, x_{t-1}, .., x_{t-lag+1}, y_t] with x_{t-l} = raw[xrow[t] - l] (zeros for the first l
// rows of a series segment: the reference pads every lagged copy with zeros, network.py:106-113) and y_t = raw[xrow[t] + 1]
__global__ void build_v_kernel(const double* __restrict__ raw, const int* __restrict__ xrow, const int* __restrict__ xpos, int N, int n,
                               int lag, double* __restrict__ V, unsigned* __restrict__ ticket) {
  if (ticket && blockIdx.x == 0 && threadIdx.x == 0) *ticket = 0u;   // block numbering of prefix_onepass_kernel, which runs next
  const int nv = n * lag;          // feature columns: block l = lag-(l+1) copies of the n genes
  const int W = 1 + nv + n;
  const long long total = (long long)N * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i / W), c = (int)(i - (long long)t * W);
    double v = 1.0;
    if (c >= 1 && c <= nv) {
      const int l = (c - 1) / n, g = (c - 1) - l * n;
      v = (xpos[t] >= l) ? raw[(long long)(xrow[t] - l) * n + g] : 0.0;
    }
    if (c > nv) v = raw[(long long)(xrow[t] + 1) * n + (c - nv - 1)];
    V[i] = v;
  }
}

// WRITE == false: per-(chunk, column) totals.  WRITE == true: running sums, starting from the sum of the totals of the
// earlier chunks.  Every thread serves PREFIX_CPT columns (PREFIX_BLOCK apart, so each store instruction is coalesced):
// the chunk's `v` rows are staged once per PREFIX_BLOCK * PREFIX_CPT columns instead of once per PREFIX_BLOCK.
#ifndef DBN_PREFIX_CPT
#define DBN_PREFIX_CPT 4
#endif
constexpr int PREFIX_CPT = DBN_PREFIX_CPT;
// Stage the chunk's rows v_t = [1, x_t, (lagged copies,) y_t], t0 <= t < t0 + len, in shared memory with cp.async, every copy
// of a thread in flight at once; the caller waits and synchronises.  Round 2 measured where the build's time went
// (profiles/r2_prefix/where_the_time_goes.txt): without ANY table store it still took 0.067 of 0.088 ms -- synchronous staging
// loops (25 dependent round trips to L2 per block), the chunk-offset sums and a shared-memory-bound totals pass, not the store
// stream.  (Staging straight from the series instead of the gathered copy V was measured too: no gain at n = 100, slower at
// n = 500 -- twice the copy instructions at 8 bytes each.)
__device__ __forceinline__ void prefix_stage(double* sv, const double* __restrict__ V, int t0, int len, int W) {
  const double* src = V + (long long)t0 * W;
  const int total = len * W;
  const bool aligned = (((long long)t0 * W) & 1) == 0;   // the chunk starts on 16 bytes
  const int n16 = aligned ? total >> 1 : 0;
  for (int i = threadIdx.x; i < n16; i += PREFIX_BLOCK) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(sv + 2 * i);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src + 2 * i) : "memory");
  }
  for (int i = 2 * n16 + threadIdx.x; i < total; i += PREFIX_BLOCK) {   // the odd tail, or everything when not aligned
    const unsigned sa = (unsigned)__cvta_generic_to_shared(sv + i);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(src + i) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// With many chunks (long series, wide rows: BASELINE config 5 has 910) summing the earlier chunks' totals inside the write
// pass is quadratic in the chunk count; chunk_scan_kernel turns the totals into exclusive running sums once
// (thread per column, coalesced over columns) and the write pass reads a single offset row (SCANNED).
__global__ void __launch_bounds__(256) chunk_scan_kernel(double* __restrict__ chunk, long long R, int n_chunks) {
  const long long p = blockIdx.x * 256ll + threadIdx.x;
  if (p >= R) return;
  double acc = 0.0;
  // 64 loads in flight per thread: a load-add-store chain per chunk is one L2 round trip per chunk (19 us for 63 chunks),
  // batches of 16 still four (8.9 us, profiles/r2_prefix/launches_before.csv)
  for (int c0 = 0; c0 < n_chunks; c0 += 64) {
    double v[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) v[j] = (c0 + j < n_chunks) ? __ldcg(chunk + (long long)(c0 + j) * R + p) : 0.0;
#pragma unroll
    for (int j = 0; j < 64; ++j) {
      if (c0 + j < n_chunks) chunk[(long long)(c0 + j) * R + p] = acc;
      acc += v[j];
    }
  }
}

// Per-(chunk, column) totals as a register-tiled product.  A table column is (row entity r, column c): ZZ rows a = 0..nv
// (value index a, c <= a, offset a (a + 1) / 2 + c) and one ZY | YY row per node i = node0 + il of the table's `n` nodes (value
// index nv + 1 + i, c <= nv + 1, offset zz_size + il (nv + 2) + c); its total over the chunk is sum_t v_t[aidx(r)] v_t[c] (v_t[aidx(r)]^2 for the YY entry c = nv + 1).
// A warp takes four row entities, a lane the columns lane + 32 j: per time step 4 broadcast + 4 conflict-free shared loads
// feed 16 FMAs (the write pass's column mapping, used here in round 1, needs 2 loads per FMA and was bound by the
// shared-memory pipe: 19 us of the 86 us build).
__global__ void __launch_bounds__(PREFIX_BLOCK) chunk_totals_kernel(const double* __restrict__ V, int N, int W, int nv, int n, int node0,
                                                                    int zz_size, long long R, int Lc, double* __restrict__ chunk) {
  extern __shared__ __align__(16) double sv[];  // [Lc][W]
  const int t0 = blockIdx.y * Lc;
  const int len = min(Lc, N - t0);
  prefix_stage(sv, V, t0, len, W);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int E = nv + 1 + n;                       // row entities
  const int e0 = (blockIdx.x * (PREFIX_BLOCK / 32) + warp) * 4;
  int aidx[4], ncols[4];
  long long base[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int e = e0 + r;
    const bool zz = e <= nv;
    aidx[r] = e < E ? (zz ? e : e + node0) : 0;    // ZZ row a: v index a; ZY row of node i = node0 + e - (nv + 1): v index nv + 1 + i
    ncols[r] = e < E ? (zz ? e + 1 : nv + 2) : 0;
    base[r] = zz ? (long long)e * (e + 1) / 2 : (long long)zz_size + (long long)(e - nv - 1) * (nv + 2);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (e0 >= E) return;
  const int maxc = max(max(ncols[0], ncols[1]), max(ncols[2], ncols[3]));
  double* out = chunk + (long long)blockIdx.y * R;
  for (int c0 = 0; c0 < maxc; c0 += 128) {
    double acc[4][4], yy[4];
    int col[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) col[j] = c0 + lane + 32 * j;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      yy[r] = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[r][j] = 0.0;
    }
    for (int t = 0; t < len; ++t) {
      const double* row = sv + t * W;
      double a[4], b[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) a[r] = row[aidx[r]];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = col[j] <= nv ? row[col[j]] : 0.0;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        yy[r] = fma(a[r], a[r], yy[r]);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[r][j] = fma(a[r], b[j], acc[r][j]);
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (col[j] < ncols[r]) out[base[r] + col[j]] = (col[j] == nv + 1) ? yy[r] : acc[r][j];
  }
}

template <bool WRITE, bool SCANNED = false>
__global__ void __launch_bounds__(PREFIX_BLOCK) prefix_kernel(const double* __restrict__ V, const unsigned short* __restrict__ colA,
                                                              const unsigned short* __restrict__ colB, int N, int W, long long R,
                                                              int Lc, double* __restrict__ chunk, double* __restrict__ table) {
  extern __shared__ __align__(16) double sv[];  // [Lc][W]
  const int t0 = blockIdx.y * Lc;
  const int len = min(Lc, N - t0);
  prefix_stage(sv, V, t0, len, W);   // in flight while the column maps and the chunk offsets are loaded
  const long long p0 = (long long)blockIdx.x * (PREFIX_BLOCK * PREFIX_CPT) + threadIdx.x;
  int ca[PREFIX_CPT], cb[PREFIX_CPT];
  bool live[PREFIX_CPT];
  double acc[PREFIX_CPT];
#pragma unroll
  for (int j = 0; j < PREFIX_CPT; ++j) {
    const long long p = p0 + (long long)j * PREFIX_BLOCK;
    live[j] = p < R;
    ca[j] = live[j] ? colA[p] : 0;
    cb[j] = live[j] ? colB[p] : 0;
    acc[j] = 0.0;
  }
  if (WRITE) {
    // offset of this chunk = sum of the totals of the chunks before it, added in chunk order (coalesced, L2-resident)
    if constexpr (SCANNED) {
#pragma unroll
      for (int j = 0; j < PREFIX_CPT; ++j)
        if (live[j]) acc[j] = __ldcg(chunk + (long long)blockIdx.y * R + p0 + (long long)j * PREFIX_BLOCK);
    } else {
      for (int c = 0; c < (int)blockIdx.y; ++c) {
#pragma unroll
        for (int j = 0; j < PREFIX_CPT; ++j)
          if (live[j]) acc[j] += chunk[(long long)c * R + p0 + (long long)j * PREFIX_BLOCK];
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (!WRITE) {
    for (int t = 0; t < len; ++t) {
#pragma unroll
      for (int j = 0; j < PREFIX_CPT; ++j) acc[j] = fma(sv[t * W + ca[j]], sv[t * W + cb[j]], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < PREFIX_CPT; ++j)
      if (live[j]) chunk[(long long)blockIdx.y * R + p0 + (long long)j * PREFIX_BLOCK] = acc[j];
  } else {
    if (blockIdx.y == 0) {
#pragma unroll
      for (int j = 0; j < PREFIX_CPT; ++j)
        if (live[j]) __stcs(table + p0 + (long long)j * PREFIX_BLOCK, 0.0);  // row 0
    }
    double* out = table + (long long)(t0 + 1) * R + p0;
    for (int t = 0; t < len; ++t) {
#pragma unroll
      for (int j = 0; j < PREFIX_CPT; ++j) {
#if defined(DBN_PREFIX_EXP) && DBN_PREFIX_EXP == 1   // experiment: no shared-memory loads (results wrong, timing only)
        acc[j] += 1.0;
#else
        acc[j] = fma(sv[t * W + ca[j]], sv[t * W + cb[j]], acc[j]);
#endif
#if defined(DBN_PREFIX_EXP) && DBN_PREFIX_EXP == 2   // experiment: only the chunk's last row is stored
        if (live[j] && t == len - 1) __stcs(out + (long long)j * PREFIX_BLOCK, acc[j]);
#else
        if (live[j]) __stcs(out + (long long)j * PREFIX_BLOCK, acc[j]);  // streaming store: the table is far larger than L2
#endif
      }
      out += R;
    }
  }
}

// ---- one-pass build: chunk totals, chunk offsets by decoupled look-back, running sums ------------------------------------
// The three launches in front of the write pass (totals, scan, and their gaps: 28 of the build's 77 us on configs[3],
// profiles/r2_prefix/where_the_time_goes.txt) become the prologue of the write pass itself.  A block takes a ticket (blocks are
// numbered in the order they START, chunk-major, so a block only ever waits for blocks that are already running), stages its
// chunk's rows, sums its PREFIX_BLOCK * PREFIX_CPT columns over the chunk (the same fma chain from 0 as chunk_totals_kernel:
// same bits), publishes them as the chunk's AGGREGATE, and finds its offset by looking back over the earlier chunks of its column
// tile: warp 0 reads 32 predecessor flags per round; the nearest chunk that has published an INCLUSIVE prefix ends the walk,
// every chunk in between contributes its aggregate.  offset = incl[c] + agg[c + 1] + ... + agg[y - 1] added in chunk order,
// and incl[c] is itself that left fold, so the offset has the same bits wherever the walk ends -- and the same bits as
// chunk_scan_kernel's running sums.  The block then publishes its own inclusive prefix and streams its rows out.
// Flags carry the launch's generation (gen << 2 | state), so nothing is cleared between builds; the ticket counter is zeroed by
// build_v_kernel, which always runs in front.
constexpr unsigned PFX_AGG = 1u, PFX_INCL = 2u;

__global__ void __launch_bounds__(PREFIX_BLOCK, 4) prefix_onepass_kernel(const double* __restrict__ V, const unsigned short* __restrict__ colA,
                                                                      const unsigned short* __restrict__ colB, int N, int W, long long R,
                                                                      int Lc, int ntx, int n_chunks, unsigned gen, unsigned* __restrict__ ticket,
                                                                      unsigned* flags, double* agg, double* incl, double* __restrict__ table) {
  extern __shared__ __align__(16) double sv[];  // [Lc][W]
  __shared__ unsigned s_ticket;
  __shared__ int s_incl_c;
  if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1u);
  __syncthreads();
  const int y = (int)(s_ticket / (unsigned)ntx), x = (int)(s_ticket - (unsigned)y * (unsigned)ntx);
  const int t0 = y * Lc;
  const int len = min(Lc, N - t0);
  prefix_stage(sv, V, t0, len, W);   // in flight while the column maps are loaded
  const long long p0 = (long long)x * (PREFIX_BLOCK * PREFIX_CPT) + threadIdx.x;
  int ca[PREFIX_CPT], cb[PREFIX_CPT];
  bool live[PREFIX_CPT];
  double tot[PREFIX_CPT], acc[PREFIX_CPT];
#pragma unroll
  for (int j = 0; j < PREFIX_CPT; ++j) {
    const long long p = p0 + (long long)j * PREFIX_BLOCK;
    live[j] = p < R;
    ca[j] = live[j] ? colA[p] : 0;
    cb[j] = live[j] ? colB[p] : 0;
    tot[j] = 0.0;
    acc[j] = 0.0;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  for (int t = 0; t < len; ++t) {
#pragma unroll
    for (int j = 0; j < PREFIX_CPT; ++j) tot[j] = fma(sv[t * W + ca[j]], sv[t * W + cb[j]], tot[j]);
  }
  volatile unsigned* fl = flags + (long long)x * n_chunks;   // this column tile's flags, one per chunk
  const unsigned f_agg = (gen << 2) | PFX_AGG, f_incl = (gen << 2) | PFX_INCL;
  const bool last_chunk = y == n_chunks - 1;
  if (y > 0) {
    if (!last_chunk) {
#pragma unroll
      for (int j = 0; j < PREFIX_CPT; ++j)
        if (live[j]) __stcg(agg + (long long)y * R + p0 + (long long)j * PREFIX_BLOCK, tot[j]);
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) fl[y] = f_agg;
    }
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x;
      int base = y - 1, found = -2;
      while (found == -2) {
        const int c = base - lane;
        unsigned st;
        while (true) {
          st = PFX_INCL;                              // chunks before the first: an inclusive prefix of zero
          if (c >= 0) {
            const unsigned f = fl[c];
            st = (f >> 2) == gen ? (f & 3u) : 0u;
          }
          const unsigned ready = __ballot_sync(0xffffffffu, st != 0u);
          const unsigned inclm = __ballot_sync(0xffffffffu, st == PFX_INCL);
          const int n_ready = (ready == 0xffffffffu) ? 32 : (__ffs((int)~ready) - 1);   // leading run of published chunks
          const int first_incl = inclm ? (__ffs((int)inclm) - 1) : 32;
          if (first_incl < n_ready) {                 // the walk ends at chunk base - first_incl
            found = base - first_incl;
            break;
          }
          if (n_ready == 32) {                        // 32 aggregates: next window
            base -= 32;
            break;
          }
        }
      }
      if (lane == 0) s_incl_c = found < 0 ? -1 : found;
      __threadfence();
    }
    __syncthreads();
    const int ci = s_incl_c;
    if (ci >= 0) {
#pragma unroll
      for (int j = 0; j < PREFIX_CPT; ++j)
        if (live[j]) acc[j] = __ldcg(incl + (long long)ci * R + p0 + (long long)j * PREFIX_BLOCK);
    }
    for (int c0 = ci + 1; c0 < y; c0 += 4) {          // 4 x PREFIX_CPT loads in flight, added in chunk order
      double v[4][PREFIX_CPT];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < PREFIX_CPT; ++j)
          v[i][j] = (c0 + i < y && live[j]) ? __ldcg(agg + (long long)(c0 + i) * R + p0 + (long long)j * PREFIX_BLOCK) : 0.0;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (c0 + i < y) {
#pragma unroll
          for (int j = 0; j < PREFIX_CPT; ++j) acc[j] += v[i][j];
        }
    }
  }
  if (!last_chunk) {
#pragma unroll
    for (int j = 0; j < PREFIX_CPT; ++j)
      if (live[j]) __stcg(incl + (long long)y * R + p0 + (long long)j * PREFIX_BLOCK, acc[j] + tot[j]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) fl[y] = f_incl;
  }
  if (y == 0) {
#pragma unroll
    for (int j = 0; j < PREFIX_CPT; ++j)
      if (live[j]) __stcs(table + p0 + (long long)j * PREFIX_BLOCK, 0.0);  // row 0
  }
  double* out = table + (long long)(t0 + 1) * R + p0;
  for (int t = 0; t < len; ++t) {
#pragma unroll
    for (int j = 0; j < PREFIX_CPT; ++j) {
      acc[j] = fma(sv[t * W + ca[j]], sv[t * W + cb[j]], acc[j]);
      if (live[j]) __stcs(out + (long long)j * PREFIX_BLOCK, acc[j]);  // streaming store: the table is far larger than L2
    }
    out += R;
  }
}

// =====================================================================================================
// Kernel 2 — batched scorer (replaces marginalLikelihood.py:91-166)
// =====================================================================================================
__global__ void __launch_bounds__(128) score_batch_kernel(Table tb, int mode, double c0, double T_half_plus_a, double two_b, int B,
                                                          const int* __restrict__ node, const int* __restrict__ pi,
                                                          const int* __restrict__ pi_len, const int* __restrict__ tau,
                                                          const int* __restrict__ tau_len, const double* __restrict__ lam,
                                                          const double* __restrict__ dlt, const double* __restrict__ mu,
                                                          double* __restrict__ out, unsigned long long* counters) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int p[PMAX];
  const int npi = pi_len[b];
#pragma unroll
  for (int j = 0; j < PMAX; ++j) p[j] = pi[b * PMAX + j];
  int f[KMAX];
  design_cols(p, npi, f);
  Cols<KMAX> cols;
  make_cols<KMAX>(cols, tb, node[b], f, npi + 1);
  const int S = tau_len[b];
  const int* tb_ = tau + (long long)b * SMAX;
  auto tau_fn = [&](int j) { return __ldg(tb_ + j); };
  double m[KMAX];
#pragma unroll
  for (int i = 0; i < KMAX; ++i) m[i] = (i < npi + 1) ? mu[b * KMAX + i] : 0.0;
  Work wk = {0, 0, 0, 0};
  double r;
  if (mode == 1)
    r = score_seq<KMAX>(tb, cols, tau_fn, S, lam[b], dlt[b], c0, T_half_plus_a, two_b, wk);
  else
    r = score_plain<KMAX, true>(tb, cols, tau_fn, S, lam[b], m, c0, T_half_plus_a, two_b, wk);
  out[b] = r;
  atomicAdd(&counters[1], wk.evals);
  atomicAdd(&counters[2], wk.segs);
  atomicAdd(&counters[3], wk.bytes);
  atomicAdd(&counters[4], wk.flops);
}

// =====================================================================================================
// FP64 peak probe: 8 independent DFMA chains per thread
// =====================================================================================================
__global__ void fp64_peak_kernel(double* out, int iters, double seed) {
  double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
  const double m = 0.999999, c = 1e-9 * (threadIdx.x + 1);
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace dbn

// =====================================================================================================
// C ABI
// =====================================================================================================
using namespace dbn;

struct dbn_handle {
  dbn_config cfg;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  std::string err;
  // data / table
  int n = 0, N = 0, W = 0, n_chunks = 0, Lc = 0;
  int lag = 1, nv = 0;        // nv = n * lag feature columns (lag > 1: zero-padded lagged copies, network.py:92-122)
  int* d_xpos = nullptr;      // position of regression row t inside its series segment
  long long raw_rows = 0;
  long long R = 0;
  int zz_size = 0;
  int tn0 = 0, tnn = 0;       // the table holds the ZY | YY blocks of nodes tn0 .. tn0 + tnn - 1 (the nodes of the handle's unit range)
  double* d_raw = nullptr;
  int* d_xrow = nullptr;
  std::vector<int> seg_len;  // series-segment lengths the row map on the device was built for
  double* d_V = nullptr;
  unsigned short *d_colA = nullptr, *d_colB = nullptr;
  double* d_chunk = nullptr;
  double* d_incl = nullptr;      // one-pass build: inclusive prefixes per (chunk, column), beside the aggregates in d_chunk
  unsigned* d_flags = nullptr;   // [column tiles][chunks], generation-stamped
  unsigned* d_ticket = nullptr;
  unsigned prefix_gen = 0;
  double* d_table = nullptr;

  bool have_stats = false;
  void* pinned = nullptr;  // host bounce buffer of dbn_read_counts
  size_t pinned_bytes = 0;
  // chain
  bool have_chain = false;
  dbn_run_params rp;
  int U = 0;          // units of this handle
  long long unit0 = 0, unit_count = -1;   // dbn_set_unit_range (count < 0: the whole n * n_chains grid)
  int it_done = 0;
  int *st_npi = nullptr, *st_pi = nullptr, *st_S = nullptr, *st_tau = nullptr;
  double *st_sig = nullptr, *st_lam = nullptr, *st_del = nullptr, *st_mu = nullptr;
  // h / nh fast path: boundary-row cache, column slots, boundary slots
  double* d_cache = nullptr;
  int* st_cg = nullptr;
  bool cache_valid = false;
  // full-parents globally coupled chain: global mean [n][U], per-block matrix workspace (large n only)
  double *st_mu_fp = nullptr, *d_fp_ws = nullptr, *d_fp_bsc = nullptr;
  int fp_grid = 0;
  int sm_count = 0;   // multiprocessors of the device (read once in dbn_chain_init)
  // dbn_score_batch: grow-only device scratch (round 1 did nine cudaMalloc / cudaFree per call) and the events around its kernel
  char* d_sb = nullptr;
  size_t sb_bytes = 0;
  cudaEvent_t sb_e0 = nullptr, sb_e1 = nullptr;
  float sb_ms = -1.f;
  int lanes_env = 0;  // DBN_LANES override (0: automatic)
  unsigned long long* d_counts = nullptr;
  long long counts_elems = 0;
  unsigned long long* d_counters = nullptr;
};

// ---- chain ------------------------------------------------------------------------------------------
__global__ void init_state_kernel(int U, int N, int* npi, int* pi, int* S, int* tau, double* sig, double* lam, double* del, double* mu) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= U) return;
  npi[u] = 0;
  for (int j = 0; j < PMAX; ++j) pi[j * U + u] = 0;
  S[u] = 1;
  tau[u] = N + 2;  // network.py:162 — only the pseudo changepoint
  for (int j = 1; j < SMAX; ++j) tau[j * U + u] = 0;
  sig[u] = lam[u] = del[u] = 1.0;  // bayesianPwLinearRegression.py:84-85
  for (int i = 0; i < KMAX; ++i) mu[i * U + u] = 0.0;
}


struct TauList {
  int len;
  int c[DBN_SMAX];
};
__global__ void set_tau_kernel(int U, TauList t, int* S, int* tau) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= U) return;
  S[u] = t.len;
  for (int j = 0; j < DBN_SMAX; ++j) tau[j * U + u] = (j < t.len) ? t.c[j] : 0;
}

static inline bool is_fp_method(int m) { return m == DBN_FP_GLOB_DBN || m == DBN_FP_NH_DBN || m == DBN_FP_H_DBN || m == DBN_FP_VV_DBN || m == DBN_FP_SEQ_DBN; }

static std::string g_create_error;
static std::mutex g_mu;

#define DBN_FAIL(h, code, ...)                       \
  do {                                               \
    char buf_[512];                                  \
    snprintf(buf_, sizeof(buf_), __VA_ARGS__);       \
    (h)->err = buf_;                                 \
    return (code);                                   \
  } while (0)

#define CU(h, expr)                                                                                      \
  do {                                                                                                   \
    cudaError_t e_ = (expr);                                                                             \
    if (e_ != cudaSuccess) DBN_FAIL(h, DBN_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

template <class T>
static void free_dev(T*& p) {
  if (p) cudaFree(p);
  p = nullptr;
}

extern "C" {

int dbn_abi_version(void) { return DBN_ABI_VERSION; }

const char* dbn_last_error(const dbn_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int dbn_create(const dbn_config* cfg, dbn_handle** out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!cfg || !out) {
    g_create_error = "dbn_create: null argument";
    return DBN_ERR_INVALID;
  }
  if (cfg->n_chains < 1) {
    g_create_error = "dbn_create: n_chains must be >= 1";
    return DBN_ERR_INVALID;
  }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g_create_error = std::string("dbn_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU fallback";
    return DBN_ERR_CUDA;
  }
  if (cfg->device < 0 || cfg->device >= count) {
    g_create_error = "dbn_create: device ordinal out of range";
    return DBN_ERR_INVALID;
  }
  e = cudaSetDevice(cfg->device);
  if (e != cudaSuccess) {
    g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
    return DBN_ERR_CUDA;
  }
  dbn_handle* h = new dbn_handle();
  h->cfg = *cfg;
  e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    g_create_error = std::string("cudaStreamCreate: ") + cudaGetErrorString(e);
    delete h;
    return DBN_ERR_CUDA;
  }
  h->stream = h->own_stream;
  e = cudaMalloc(&h->d_counters, 32 * sizeof(unsigned long long));   // [8] work counters (+ 16 probe counters of diagnostic builds)
  if (e == cudaSuccess) e = cudaMemset(h->d_counters, 0, 32 * sizeof(unsigned long long));
  if (e != cudaSuccess) {
    g_create_error = std::string("cudaMalloc: ") + cudaGetErrorString(e);
    cudaStreamDestroy(h->own_stream);
    delete h;
    return DBN_ERR_CUDA;
  }
  *out = h;
  return DBN_OK;
}

static void free_stats(dbn_handle* h) {
  free_dev(h->d_raw);
  free_dev(h->d_xrow);
  free_dev(h->d_xpos);
  free_dev(h->d_V);
  free_dev(h->d_colA);
  free_dev(h->d_colB);
  free_dev(h->d_chunk);
  free_dev(h->d_incl);
  free_dev(h->d_flags);
  free_dev(h->d_ticket);
  free_dev(h->d_table);
  h->have_stats = false;
}

static void free_chain(dbn_handle* h) {
  free_dev(h->st_npi);
  free_dev(h->st_pi);
  free_dev(h->st_S);
  free_dev(h->st_tau);
  free_dev(h->st_sig);
  free_dev(h->st_lam);
  free_dev(h->st_del);
  free_dev(h->st_mu);
  free_dev(h->d_cache);
  free_dev(h->st_cg);
  free_dev(h->st_mu_fp);
  free_dev(h->d_fp_ws);
  free_dev(h->d_fp_bsc);
  h->cache_valid = false;
  free_dev(h->d_counts);
  h->have_chain = false;
}

void dbn_destroy(dbn_handle* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  cudaDeviceSynchronize();
  free_stats(h);
  free_chain(h);
  free_dev(h->d_counters);
  free_dev(h->d_sb);
  if (h->sb_e0) cudaEventDestroy(h->sb_e0);
  if (h->sb_e1) cudaEventDestroy(h->sb_e1);
  if (h->pinned) cudaFreeHost(h->pinned);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
}

int dbn_set_stream(dbn_handle* h, void* cuda_stream) {
  if (!h) return DBN_ERR_INVALID;
  cudaStream_t next = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
  if (next == h->stream) return DBN_OK;
  // work already queued on the previous stream (prefix kernels, the memsets and init kernel of dbn_chain_init, earlier
  // dbn_run windows) must be ordered before whatever the next calls launch on the new one
  CU(h, cudaSetDevice(h->cfg.device));
  cudaEvent_t ev;
  CU(h, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  cudaError_t e = cudaEventRecord(ev, h->stream);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(next, ev, 0);
  cudaEventDestroy(ev);
  if (e != cudaSuccess) DBN_FAIL(h, DBN_ERR_CUDA, "dbn_set_stream: %s", cudaGetErrorString(e));
  h->stream = next;
  return DBN_OK;
}

int dbn_sync(dbn_handle* h) {
  if (!h) return DBN_ERR_INVALID;
  CU(h, cudaSetDevice(h->cfg.device));
  CU(h, cudaStreamSynchronize(h->stream));
  return DBN_OK;
}

// Nodes whose ZY | YY blocks the handle's table needs: every node (no unit range) or the nodes its unit range touches
// (u = node * n_chains + chain).  A rank / handle of a sharded run then builds and keeps only its own columns: for BASELINE
// configs[4] at 8 ranks 12.6 GB instead of the 30 GB of the full table (the ZZ block is shared by all nodes).
static void wanted_nodes(const dbn_handle* h, int n, int* tn0, int* tnn) {
  *tn0 = 0;
  *tnn = n;
  if (h->unit_count < 0 || h->cfg.n_chains < 1) return;
  const long long C = h->cfg.n_chains;
  long long first = h->unit0 / C, last = (h->unit0 + h->unit_count - 1) / C;
  if (first > n - 1) first = n - 1;   // a range past the grid is refused by dbn_chain_init; keep the table well-formed
  if (last > n - 1) last = n - 1;
  *tn0 = (int)first;
  *tnn = (int)(last - first + 1);
}

// (Re)allocate everything whose size depends on the table's column count for the node range h->tn0, h->tnn and upload the
// column map: ZZ(a,b) = v[a] v[b]; ZY(i,a) = v[a] v[nv+1+i]; YY(i) = v[nv+1+i]^2.  h->n, nv, N, W, zz_size, Lc, n_chunks are set.
static int layout_columns(dbn_handle* h) {
  free_dev(h->d_colA);
  free_dev(h->d_colB);
  free_dev(h->d_chunk);
  free_dev(h->d_incl);
  free_dev(h->d_flags);
  free_dev(h->d_ticket);
  free_dev(h->d_table);
  const int nv = h->nv, N = h->N;
  h->R = (long long)h->zz_size + (long long)h->tnn * (nv + 2);
  std::vector<unsigned short> ca((size_t)h->R), cb((size_t)h->R);
  {
    size_t p = 0;
    for (int a = 0; a <= nv; ++a)
      for (int b = 0; b <= a; ++b) {
        ca[p] = (unsigned short)a;
        cb[p] = (unsigned short)b;
        ++p;
      }
    for (int i = h->tn0; i < h->tn0 + h->tnn; ++i) {
      for (int a = 0; a <= nv; ++a) {
        ca[p] = (unsigned short)a;
        cb[p] = (unsigned short)(nv + 1 + i);
        ++p;
      }
      ca[p] = cb[p] = (unsigned short)(nv + 1 + i);
      ++p;
    }
  }
  CU(h, cudaMalloc(&h->d_colA, (size_t)h->R * sizeof(unsigned short)));
  CU(h, cudaMalloc(&h->d_colB, (size_t)h->R * sizeof(unsigned short)));
  CU(h, cudaMalloc(&h->d_chunk, (size_t)h->n_chunks * h->R * sizeof(double)));
  // The one-pass build (prefix_onepass_kernel) is kept as a tested alternative, not the default: on one B200 it takes 0.083 ms
  // against 0.0765 ms for the three-launch build on configs[3] and 12.0 against 8.3 ms at configs[4] size
  // (profiles/r2_prefix/onepass_ab.txt): every block of a wave is in its prologue (totals, flags, look-back: ~8 us of dependent L2
  // round trips) at the same time, so the prologues are not hidden behind other blocks' stores, and at n = 500 only two blocks fit an SM.
  if (getenv("DBN_PREFIX_ONEPASS")) {
    CU(h, cudaMalloc(&h->d_incl, (size_t)h->n_chunks * h->R * sizeof(double)));
    const size_t ntx = (size_t)((h->R + PREFIX_BLOCK * PREFIX_CPT - 1) / (PREFIX_BLOCK * PREFIX_CPT));
    CU(h, cudaMalloc(&h->d_flags, ntx * h->n_chunks * sizeof(unsigned)));
    CU(h, cudaMemsetAsync(h->d_flags, 0, ntx * h->n_chunks * sizeof(unsigned), h->stream));
    CU(h, cudaMalloc(&h->d_ticket, sizeof(unsigned)));
    h->prefix_gen = 0;
  }
  cudaError_t e = cudaMalloc(&h->d_table, (size_t)(N + 1) * h->R * sizeof(double));
  if (e != cudaSuccess) DBN_FAIL(h, DBN_ERR_NOMEM, "prefix table of %.2f GB does not fit: %s", (double)(N + 1) * h->R * 8 / 1e9, cudaGetErrorString(e));
  CU(h, cudaMemcpyAsync(h->d_colA, ca.data(), (size_t)h->R * sizeof(unsigned short), cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(h->d_colB, cb.data(), (size_t)h->R * sizeof(unsigned short), cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));  // ca/cb go out of scope
  return DBN_OK;
}

static int launch_prefix(dbn_handle* h) {
  const int n = h->n, N = h->N, W = h->W;
  {
    // one element per thread up to 16 blocks per SM: the gather is two dependent round trips per element, and 296 blocks
    // walked five elements each in sequence (8.5 us for 400 k elements, now 5.9)
    const long long total = (long long)N * W;
    const unsigned gv = (unsigned)std::min<long long>((total + 255) / 256, 148ll * 16);
    build_v_kernel<<<gv, 256, 0, h->stream>>>(h->d_raw, h->d_xrow, h->d_xpos, N, n, h->lag, h->d_V, h->d_ticket);
    CU(h, cudaGetLastError());
  }
  const size_t smem = (size_t)h->Lc * W * sizeof(double);
  dim3 grid((unsigned)((h->R + PREFIX_BLOCK * PREFIX_CPT - 1) / (PREFIX_BLOCK * PREFIX_CPT)), (unsigned)h->n_chunks);
  if (h->d_incl) {   // DBN_PREFIX_ONEPASS: one launch for totals, look-back offsets and the write stream (measured slower, see below)
    h->prefix_gen = (h->prefix_gen + 1) & 0x3fffffffu;
    if (h->prefix_gen == 0) {          // the generation wrapped: old stamps could be mistaken for new ones
      CU(h, cudaMemsetAsync(h->d_flags, 0, (size_t)grid.x * grid.y * sizeof(unsigned), h->stream));
      h->prefix_gen = 1;
    }
    prefix_onepass_kernel<<<grid.x * grid.y, PREFIX_BLOCK, smem, h->stream>>>(h->d_V, h->d_colA, h->d_colB, N, W, h->R, h->Lc, (int)grid.x,
                                                                               h->n_chunks, h->prefix_gen, h->d_ticket, h->d_flags, h->d_chunk,
                                                                               h->d_incl, h->d_table);
    CU(h, cudaGetLastError());
    return DBN_OK;
  }
  if (getenv("DBN_PREFIX_OLD_TOTALS")) {   // round-1 totals pass (A/B)
    prefix_kernel<false><<<grid, PREFIX_BLOCK, smem, h->stream>>>(h->d_V, h->d_colA, h->d_colB, N, W, h->R, h->Lc, h->d_chunk, h->d_table);
  } else {
    const int row_tiles = (h->nv + 1 + h->tnn + 3) / 4;
    dim3 gt((unsigned)((row_tiles + PREFIX_BLOCK / 32 - 1) / (PREFIX_BLOCK / 32)), (unsigned)h->n_chunks);
    chunk_totals_kernel<<<gt, PREFIX_BLOCK, smem, h->stream>>>(h->d_V, N, W, h->nv, h->tnn, h->tn0, h->zz_size, h->R, h->Lc, h->d_chunk);
  }
  CU(h, cudaGetLastError());
  if (!getenv("DBN_PREFIX_NO_SCAN")) {   // the 64-way scan costs less than the in-pass sums of the write kernel (0.079 vs 0.082 ms)
    chunk_scan_kernel<<<(unsigned)((h->R + 255) / 256), 256, 0, h->stream>>>(h->d_chunk, h->R, h->n_chunks);
    CU(h, cudaGetLastError());
    prefix_kernel<true, true><<<grid, PREFIX_BLOCK, smem, h->stream>>>(h->d_V, h->d_colA, h->d_colB, N, W, h->R, h->Lc, h->d_chunk, h->d_table);
  } else {
    prefix_kernel<true><<<grid, PREFIX_BLOCK, smem, h->stream>>>(h->d_V, h->d_colA, h->d_colB, N, W, h->R, h->Lc, h->d_chunk, h->d_table);
  }
  CU(h, cudaGetLastError());
  return DBN_OK;
}

int dbn_build_stats(dbn_handle* h, const double* data, const int32_t* seg_len, int32_t n_seg, int32_t n, int32_t lag) {
  if (!h) return DBN_ERR_INVALID;
  if (!data || !seg_len || n_seg < 1 || n < 2) DBN_FAIL(h, DBN_ERR_INVALID, "dbn_build_stats: bad arguments (need n >= 2, n_seg >= 1)");
  if (lag < 1 || lag > 2)
    DBN_FAIL(h, DBN_ERR_INVALID, "dbn_build_stats: lag must be 1 or 2 (the reference scores lag-1 and highest-lag copies only, scores.py:172-178)");
  if ((lag + 1) * n + 1 > 65535) DBN_FAIL(h, DBN_ERR_INVALID, "dbn_build_stats: n too large for the 16-bit column map");
  CU(h, cudaSetDevice(h->cfg.device));
  long long rows = 0, N = 0;
  for (int s = 0; s < n_seg; ++s) {
    if (seg_len[s] < 2) DBN_FAIL(h, DBN_ERR_INVALID, "dbn_build_stats: every series segment needs >= 2 rows");
    rows += seg_len[s];
    N += seg_len[s] - 1;
  }
  // regression row t -> raw row of x_t (network.py:105-118: rows [:len-1] of each segment, concatenated)
  std::vector<int> xrow((size_t)N), xpos((size_t)N);
  {
    long long base = 0, t = 0;
    for (int s = 0; s < n_seg; ++s) {
      for (int j = 0; j + 1 < seg_len[s]; ++j) {
        xpos[(size_t)t] = j;
        xrow[(size_t)t++] = (int)(base + j);
      }
      base += seg_len[s];
    }
  }
  int tn0, tnn;
  wanted_nodes(h, n, &tn0, &tnn);
  const bool same_geometry = h->have_stats && h->n == n && h->lag == lag && h->N == (int)N && h->raw_rows == rows && h->tn0 == tn0 && h->tnn == tnn;
  if (!same_geometry) {
    free_stats(h);
    free_chain(h);
    h->n = n;
    h->lag = lag;
    const int nv = n * lag;      // feature columns z = (1, x_t, x_{t-1}, ..): every kernel sees `nv` candidate columns
    h->nv = nv;
    h->N = (int)N;
    h->raw_rows = rows;
    h->W = 1 + nv + n;
    h->zz_size = (nv + 1) * (nv + 2) / 2;
    h->tn0 = tn0;
    h->tnn = tnn;
    const size_t row_bytes = (size_t)h->W * sizeof(double);
    int Lc = (int)(96 * 1024 / row_bytes);
    const int Lmax = Lc;
    if (Lc > 32) Lc = 32;  // measured on B200 (scripts/prefix_tune.py)
    if (const char* e = getenv("DBN_PREFIX_LC")) Lc = std::min(Lmax, std::max(1, atoi(e)));  // tuning knob
    if (Lc < 1) Lc = 1;
    h->Lc = Lc;
    h->n_chunks = (int)((N + Lc - 1) / Lc);
    CU(h, cudaFuncSetAttribute(prefix_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(Lc * row_bytes)));
    CU(h, cudaFuncSetAttribute(prefix_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(Lc * row_bytes)));
    CU(h, cudaFuncSetAttribute(prefix_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(Lc * row_bytes)));
    CU(h, cudaFuncSetAttribute(chunk_totals_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(Lc * row_bytes)));
    CU(h, cudaFuncSetAttribute(prefix_onepass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(Lc * row_bytes)));
    CU(h, cudaMalloc(&h->d_raw, (size_t)rows * n * sizeof(double)));
    CU(h, cudaMalloc(&h->d_xrow, (size_t)N * sizeof(int)));
    CU(h, cudaMalloc(&h->d_xpos, (size_t)N * sizeof(int)));
    CU(h, cudaMalloc(&h->d_V, (size_t)N * h->W * sizeof(double)));
    int rc = layout_columns(h);
    if (rc != DBN_OK) return rc;
  }
  // same geometry: the table is rebuilt in place and the chain state (if any) carries on over the new data
  CU(h, cudaMemcpyAsync(h->d_raw, data, (size_t)rows * n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  bool must_sync = false;
  if (!same_geometry || h->seg_len != std::vector<int>(seg_len, seg_len + n_seg)) {
    CU(h, cudaMemcpyAsync(h->d_xrow, xrow.data(), (size_t)N * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(h->d_xpos, xpos.data(), (size_t)N * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    h->seg_len.assign(seg_len, seg_len + n_seg);
    must_sync = true;                        // the staging vector goes out of scope
  }
  {
    // a pageable caller buffer has been staged by the time cudaMemcpyAsync returns; a PINNED one is read by the copy
    // engine later, and the caller keeps it unchanged until the next synchronising call (dbn_sync, dbn_read_*): no host
    // synchronisation on this path then (include/dbn_b200.h)
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, data) != cudaSuccess) {
      cudaGetLastError();
      must_sync = true;
    }
  }
  if (must_sync) CU(h, cudaStreamSynchronize(h->stream));
  int rc = launch_prefix(h);                 // asynchronous: overlaps whatever the host does next
  if (rc != DBN_OK) return rc;
  h->have_stats = true;
  h->cache_valid = false;  // the chain (if any) re-gathers its boundary rows from the new table
  return DBN_OK;
}

int dbn_rebuild_stats(dbn_handle* h) {
  if (!h) return DBN_ERR_INVALID;
  if (!h->have_stats) DBN_FAIL(h, DBN_ERR_STATE, "dbn_rebuild_stats: call dbn_build_stats first");
  CU(h, cudaSetDevice(h->cfg.device));
  h->cache_valid = false;
  return launch_prefix(h);
}

int dbn_stats_info(const dbn_handle* h, int32_t* N, int64_t* row_doubles, int64_t* algorithmic_bytes) {
  if (!h || !h->have_stats) return DBN_ERR_STATE;
  if (N) *N = h->N;
  if (row_doubles) *row_doubles = h->R;
  // 8 * [ N*n*2 (read x_t, y_t) + (N+1) * ((n+1)(n+2)/2 + (n+1) n + n) ]
  if (algorithmic_bytes) *algorithmic_bytes = 8ll * ((long long)h->N * h->n * 2 + (long long)(h->N + 1) * h->R);
  return DBN_OK;
}

int dbn_read_stats_row(dbn_handle* h, int32_t t, double* out_row) {
  if (!h || !out_row) return DBN_ERR_INVALID;
  if (!h->have_stats) DBN_FAIL(h, DBN_ERR_STATE, "dbn_read_stats_row: call dbn_build_stats first");
  if (t < 0 || t > h->N) DBN_FAIL(h, DBN_ERR_INVALID, "dbn_read_stats_row: t out of range");
  CU(h, cudaSetDevice(h->cfg.device));
  CU(h, cudaMemcpyAsync(out_row, h->d_table + (long long)t * h->R, (size_t)h->R * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return DBN_OK;
}

static Table make_table(const dbn_handle* h) {
  Table t;
  t.base = h->d_table;
  t.row_stride = h->R;
  t.n = h->nv;      // candidate feature columns (n genes x lag)
  t.N = h->N;
  t.zz_size = h->zz_size;
  t.zy_base = h->zz_size - h->tn0 * (h->nv + 2);
  return t;
}

static double ml_const(double T, double alpha, double beta) {
  return lgamma(T / 2 + alpha) - lgamma(alpha) - (T / 2) * log(M_PI) + alpha * log(2 * beta);
}

int dbn_score_batch(dbn_handle* h, int32_t mode, double alpha, double beta, double T, int32_t B, const int32_t* node,
                    const int32_t* pi, const int32_t* pi_len, const int32_t* tau, const int32_t* tau_len,
                    const double* lambda2, const double* delta2, const double* mu, double* out_logml) {
  if (!h) return DBN_ERR_INVALID;
  if (!h->have_stats) DBN_FAIL(h, DBN_ERR_STATE, "dbn_score_batch: call dbn_build_stats first");
  if (B < 0 || !node || !pi || !pi_len || !tau || !tau_len || !lambda2 || !delta2 || !mu || !out_logml)
    DBN_FAIL(h, DBN_ERR_INVALID, "dbn_score_batch: null argument");
  if (mode != 0 && mode != 1) DBN_FAIL(h, DBN_ERR_INVALID, "dbn_score_batch: mode must be 0 (shared prior mean) or 1 (seq-coup)");
  if (B == 0) return DBN_OK;
  for (int b = 0; b < B; ++b) {
    if (node[b] < 0 || node[b] >= h->n || pi_len[b] < 0 || pi_len[b] > PMAX || tau_len[b] < 1 || tau_len[b] > SMAX)
      DBN_FAIL(h, DBN_ERR_INVALID, "dbn_score_batch: item %d out of range", b);
    if (node[b] < h->tn0 || node[b] >= h->tn0 + h->tnn)
      DBN_FAIL(h, DBN_ERR_INVALID, "dbn_score_batch: item %d scores node %d, but this handle's table holds nodes %d..%d only (dbn_set_unit_range)", b,
               node[b], h->tn0, h->tn0 + h->tnn - 1);
    for (int j = 0; j < pi_len[b]; ++j)
      if (pi[b * PMAX + j] < 0 || pi[b * PMAX + j] >= h->nv || pi[b * PMAX + j] % h->n == node[b])
        DBN_FAIL(h, DBN_ERR_INVALID, "dbn_score_batch: item %d has an invalid parent", b);
    int prev = 1;
    for (int j = 0; j + 1 < tau_len[b]; ++j) {
      const int c = tau[b * SMAX + j];
      if (c <= prev || c > h->N) DBN_FAIL(h, DBN_ERR_INVALID, "dbn_score_batch: item %d has an invalid changepoint list", b);
      prev = c;
    }
  }
  CU(h, cudaSetDevice(h->cfg.device));
  const size_t b4 = ((size_t)B * sizeof(int) + 255) & ~(size_t)255, b8 = ((size_t)B * sizeof(double) + 255) & ~(size_t)255;
  const size_t need = b4 * (3 + PMAX + SMAX) + b8 * (3 + KMAX);
  if (need > h->sb_bytes) {
    CU(h, cudaStreamSynchronize(h->stream));
    free_dev(h->d_sb);
    h->sb_bytes = 0;
    cudaError_t e = cudaMalloc(&h->d_sb, need);
    if (e != cudaSuccess) DBN_FAIL(h, DBN_ERR_NOMEM, "dbn_score_batch: %.1f MB of scratch do not fit: %s", need / 1e6, cudaGetErrorString(e));
    h->sb_bytes = need;
  }
  if (!h->sb_e0) {
    CU(h, cudaEventCreate(&h->sb_e0));
    CU(h, cudaEventCreate(&h->sb_e1));
  }
  char* q = h->d_sb;
  auto take = [&](size_t bytes) { char* r = q; q += bytes; return r; };
  int* d_node = (int*)take(b4); int* d_pi = (int*)take(b4 * PMAX); int* d_pl = (int*)take(b4);
  int* d_tau = (int*)take(b4 * SMAX); int* d_tl = (int*)take(b4);
  double* d_lam = (double*)take(b8); double* d_dlt = (double*)take(b8); double* d_mu = (double*)take(b8 * KMAX); double* d_out = (double*)take(b8);
  const size_t n4 = (size_t)B * sizeof(int), n8 = (size_t)B * sizeof(double);
  CU(h, cudaMemcpyAsync(d_node, node, n4, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(d_pi, pi, n4 * PMAX, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(d_pl, pi_len, n4, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(d_tau, tau, n4 * SMAX, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(d_tl, tau_len, n4, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(d_lam, lambda2, n8, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(d_dlt, delta2, n8, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(d_mu, mu, n8 * KMAX, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaEventRecord(h->sb_e0, h->stream));
  score_batch_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(make_table(h), mode, ml_const(T, alpha, beta), T / 2 + alpha, 2 * beta, B,
                                                            d_node, d_pi, d_pl, d_tau, d_tl, d_lam, d_dlt, d_mu, d_out, h->d_counters);
  CU(h, cudaGetLastError());
  CU(h, cudaEventRecord(h->sb_e1, h->stream));
  CU(h, cudaMemcpyAsync(out_logml, d_out, n8, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  CU(h, cudaEventElapsedTime(&h->sb_ms, h->sb_e0, h->sb_e1));
  return DBN_OK;
}

int dbn_score_batch_time(dbn_handle* h, double* ms) {
  if (!h || !ms) return DBN_ERR_INVALID;
  if (h->sb_ms < 0.f) DBN_FAIL(h, DBN_ERR_STATE, "dbn_score_batch_time: no dbn_score_batch call yet");
  *ms = h->sb_ms;
  return DBN_OK;
}

int dbn_chain_init(dbn_handle* h, const dbn_run_params* p) {
  if (!h || !p) return DBN_ERR_INVALID;
  if (!h->have_stats) DBN_FAIL(h, DBN_ERR_STATE, "dbn_chain_init: call dbn_build_stats first");
  if (p->method < DBN_H_DBN || p->method > DBN_FP_SEQ_DBN) DBN_FAIL(h, DBN_ERR_INVALID, "dbn_chain_init: unknown method %d", p->method);
  if (!is_fp_method(p->method) && (p->fan_in < 1 || p->fan_in > PMAX)) DBN_FAIL(h, DBN_ERR_INVALID, "dbn_chain_init: fan_in must be in 1..%d", PMAX);
  if (p->thin < 1 || p->burn_in < 0) DBN_FAIL(h, DBN_ERR_INVALID, "dbn_chain_init: bad burn_in/thin");
  if (p->num_samples < 2 || p->num_samples > h->N) DBN_FAIL(h, DBN_ERR_INVALID, "dbn_chain_init: num_samples must be in 2..N");
  const long long grid_units = (long long)h->n * h->cfg.n_chains;
  if (grid_units > 0x7fffffffll) DBN_FAIL(h, DBN_ERR_INVALID, "dbn_chain_init: too many units");
  if (h->unit_count >= 0 && (h->unit0 < 0 || h->unit0 + h->unit_count > grid_units || h->unit_count < 1))
    DBN_FAIL(h, DBN_ERR_INVALID, "dbn_chain_init: unit range [%lld, %lld) outside the %lld units of the grid", h->unit0, h->unit0 + h->unit_count, grid_units);
  const long long U = h->unit_count >= 0 ? h->unit_count : grid_units;
  if (h->unit_count < 0) h->unit0 = 0;
  CU(h, cudaSetDevice(h->cfg.device));
  free_chain(h);
  CU(h, cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->cfg.device));
  {
    const char* ev = getenv("DBN_LANES");
    h->lanes_env = (ev && atoi(ev) >= 1 && atoi(ev) <= 32) ? atoi(ev) : 0;
  }
  h->rp = *p;
  h->U = (int)U;
  h->it_done = 0;
  CU(h, cudaMalloc(&h->st_npi, U * sizeof(int)));
  CU(h, cudaMalloc(&h->st_pi, U * PMAX * sizeof(int)));
  CU(h, cudaMalloc(&h->st_S, U * sizeof(int)));
  CU(h, cudaMalloc(&h->st_tau, U * SMAX * sizeof(int)));
  CU(h, cudaMalloc(&h->st_sig, U * sizeof(double)));
  CU(h, cudaMalloc(&h->st_lam, U * sizeof(double)));
  CU(h, cudaMalloc(&h->st_del, U * sizeof(double)));
  CU(h, cudaMalloc(&h->st_mu, U * KMAX * sizeof(double)));
  if (p->method == DBN_H_DBN || p->method == DBN_NH_DBN) {
    if (h->N + 2 > 65535) DBN_FAIL(h, DBN_ERR_INVALID, "dbn_chain_init: the h / nh kernel keeps changepoints as 16-bit values (N + 2 <= 65535), N = %d", h->N);
    const int km = p->fan_in <= 3 ? 4 : 5;
    const size_t cache_bytes = (size_t)BSLOTS * fast_cache_row_doubles(km) * U * sizeof(double);
    cudaError_t e = cudaMalloc(&h->d_cache, cache_bytes);
    if (e != cudaSuccess) DBN_FAIL(h, DBN_ERR_NOMEM, "boundary-row cache of %.2f GB does not fit: %s", cache_bytes / 1e9, cudaGetErrorString(e));
    CU(h, cudaMalloc(&h->st_cg, U * PMAX * sizeof(int)));
  } else if (!is_fp_method(p->method)) {
    // coupled methods (seq / glob / var-glob): boundary rows of the current state + the proposal's new row + a zero row (dbn_score.cuh)
    const int km = p->fan_in <= 3 ? 4 : 5;
    const size_t cache_bytes = chain_cache_doubles_per_unit(km) * U * sizeof(double);
    cudaError_t e = cudaMalloc(&h->d_cache, cache_bytes);
    if (e != cudaSuccess) DBN_FAIL(h, DBN_ERR_NOMEM, "boundary-row cache of %.2f GB does not fit: %s", cache_bytes / 1e9, cudaGetErrorString(e));
  }
  h->counts_elems = (long long)h->n * h->n + h->n + (long long)h->n * (h->N + 3) + h->n;
  if (is_fp_method(p->method)) {
    if (h->lag != 1) DBN_FAIL(h, DBN_ERR_INVALID, "dbn_chain_init: the full-parents kernels take lag == 1 only");
    h->counts_elems += (long long)h->n * h->n;  // neg[n][n]
    const int k = h->n;
    CU(h, cudaMalloc(&h->st_mu_fp, (size_t)U * k * sizeof(double)));
    CU(h, cudaMemsetAsync(h->st_mu_fp, 0, (size_t)U * k * sizeof(double), h->stream));
    cudaDeviceProp prop;
    CU(h, cudaGetDeviceProperties(&prop, h->cfg.device));
    const bool in_smem = fp_smem_bytes(k, true) <= 160 * 1024;
    if (!in_smem && fp_smem_bytes(k, false) > (size_t)prop.sharedMemPerBlockOptin)
      DBN_FAIL(h, DBN_ERR_INVALID, "dbn_chain_init: n = %d is too large for the full-parents kernel (%zu bytes of shared memory per block, %zu available)",
               k, fp_smem_bytes(k, false), (size_t)prop.sharedMemPerBlockOptin);
    // large k: one block of FP_BIG_BLOCK threads per SM (fewer workspaces in flight, dbn_fp_big.cuh)
    h->fp_grid = (int)(in_smem ? U : (U < (long long)prop.multiProcessorCount ? U : (long long)prop.multiProcessorCount));
    if (!in_smem) {
      const size_t ws = (size_t)h->fp_grid * fp_ws_doubles_per_block(k) * sizeof(double);
      cudaError_t e = cudaMalloc(&h->d_fp_ws, ws);
      if (e != cudaSuccess) DBN_FAIL(h, DBN_ERR_NOMEM, "full-parents workspace of %.2f GB does not fit: %s", ws / 1e9, cudaGetErrorString(e));
    }
    if (p->method != DBN_FP_GLOB_DBN) CU(h, cudaMalloc(&h->d_fp_bsc, (size_t)h->fp_grid * 2 * SMAX * k * sizeof(double)));
  }
  CU(h, cudaMalloc(&h->d_counts, (size_t)h->counts_elems * sizeof(unsigned long long)));
  CU(h, cudaMemsetAsync(h->d_counts, 0, (size_t)h->counts_elems * sizeof(unsigned long long), h->stream));
  CU(h, cudaMemsetAsync(h->d_counters, 0, 8 * sizeof(unsigned long long), h->stream));
  init_state_kernel<<<(unsigned)((U + 255) / 256), 256, 0, h->stream>>>((int)U, h->N, h->st_npi, h->st_pi, h->st_S, h->st_tau, h->st_sig,
                                                                       h->st_lam, h->st_del, h->st_mu);
  CU(h, cudaGetLastError());
  h->have_chain = true;
  return DBN_OK;
}

int dbn_set_unit_range(dbn_handle* h, int64_t first_unit, int64_t n_units) {
  if (!h) return DBN_ERR_INVALID;
  if (first_unit < 0 || n_units == 0) DBN_FAIL(h, DBN_ERR_INVALID, "dbn_set_unit_range: first_unit >= 0 and n_units != 0 (negative: the whole grid)");
  h->unit0 = n_units < 0 ? 0 : first_unit;
  h->unit_count = n_units < 0 ? -1 : n_units;
  if (h->have_chain) {   // the range takes effect with the next dbn_chain_init
    CU(h, cudaSetDevice(h->cfg.device));
    CU(h, cudaStreamSynchronize(h->stream));
    free_chain(h);
  }
  if (h->have_stats) {   // the table follows the range's nodes: rebuilt from the series already on the device
    int tn0, tnn;
    wanted_nodes(h, h->n, &tn0, &tnn);
    if (tn0 != h->tn0 || tnn != h->tnn) {
      CU(h, cudaSetDevice(h->cfg.device));
      CU(h, cudaStreamSynchronize(h->stream));
      h->tn0 = tn0;
      h->tnn = tnn;
      h->have_stats = false;
      int rc = layout_columns(h);
      if (rc != DBN_OK) return rc;
      rc = launch_prefix(h);
      if (rc != DBN_OK) return rc;
      h->have_stats = true;
      h->cache_valid = false;
    }
  }
  return DBN_OK;
}

int dbn_set_changepoints(dbn_handle* h, const int32_t* tau, int32_t len) {
  if (!h || !tau) return DBN_ERR_INVALID;
  if (!h->have_chain) DBN_FAIL(h, DBN_ERR_STATE, "dbn_set_changepoints: call dbn_chain_init first");
  if (h->it_done != 0) DBN_FAIL(h, DBN_ERR_STATE, "dbn_set_changepoints: the chain has already been advanced");
  if (h->rp.method == DBN_H_DBN) DBN_FAIL(h, DBN_ERR_INVALID, "dbn_set_changepoints: the homogeneous model has no changepoints");
  if (len < 1 || len > DBN_SMAX - 1) DBN_FAIL(h, DBN_ERR_INVALID, "dbn_set_changepoints: 1..%d changepoints (with the pseudo changepoint)", DBN_SMAX - 1);
  if (tau[len - 1] != h->N + 2) DBN_FAIL(h, DBN_ERR_INVALID, "dbn_set_changepoints: the last entry must be the pseudo changepoint N + 2 = %d (network.py:162)", h->N + 2);
  for (int j = 0; j + 1 < len; ++j)
    if (tau[j] < 2 || tau[j] > h->N || (j > 0 && tau[j] <= tau[j - 1]))
      DBN_FAIL(h, DBN_ERR_INVALID, "dbn_set_changepoints: changepoints must be increasing values in 2..N");
  CU(h, cudaSetDevice(h->cfg.device));
  TauList t;
  t.len = len;
  for (int j = 0; j < DBN_SMAX; ++j) t.c[j] = (j < len) ? tau[j] : 0;
  set_tau_kernel<<<(unsigned)((h->U + 255) / 256), 256, 0, h->stream>>>(h->U, t, h->st_S, h->st_tau);
  CU(h, cudaGetLastError());
  h->cache_valid = false;
  return DBN_OK;
}

int dbn_run(dbn_handle* h, int32_t n_iter, const double* replay_d, const int32_t* replay_i, double* trace_d, int32_t* trace_i) {
  if (!h) return DBN_ERR_INVALID;
  if (!h->have_chain) DBN_FAIL(h, DBN_ERR_STATE, "dbn_run: call dbn_chain_init first");
  if (is_fp_method(h->rp.method)) DBN_FAIL(h, DBN_ERR_STATE, "dbn_run: the full-parents methods run with dbn_run_fp");
  if (n_iter < 0) DBN_FAIL(h, DBN_ERR_INVALID, "dbn_run: n_iter < 0");
  if ((replay_d == nullptr) != (replay_i == nullptr)) DBN_FAIL(h, DBN_ERR_INVALID, "dbn_run: replay_d and replay_i must be given together");
  if ((replay_d != nullptr) && !(trace_d && trace_i)) DBN_FAIL(h, DBN_ERR_INVALID, "dbn_run: replay mode requires trace buffers");
  if ((trace_d == nullptr) != (trace_i == nullptr)) DBN_FAIL(h, DBN_ERR_INVALID, "dbn_run: trace_d and trace_i must be given together");
  if (n_iter == 0) return DBN_OK;
  CU(h, cudaSetDevice(h->cfg.device));
  const dbn_run_params& p = h->rp;
  ChainArgs a;
  memset(&a, 0, sizeof(a));
  a.tb = make_table(h);
  a.method = p.method;
  a.fan_in = p.fan_in;
  a.with_tau = p.with_tau;
  a.num_samples = p.num_samples;
  a.burn_in = p.burn_in;
  a.thin = p.thin;
  a.C = h->cfg.n_chains;
  a.U = h->U;
  a.unit0 = (int)h->unit0;
  a.n_genes = h->n;
  a.lag = h->lag;
  a.it0 = h->it_done;
  a.n_iter = n_iter;
  a.chain_offset = (unsigned)h->cfg.chain_offset;
  a.seed = h->cfg.seed;
  a.T_sigma_half = p.T_sigma / 2;
  a.a_sig = p.a_sigma; a.b_sig = p.b_sigma;
  a.a_lam = p.a_lambda; a.b_lam = p.b_lambda;
  a.a_del = p.a_delta; a.b_del = p.b_delta;
  a.log_p = log(p.cp_p);
  a.log_1mp = log(1.0 - p.cp_p);
  a.c0 = ml_const(p.T_ml, p.a_sigma, p.b_sigma);
  a.T_half_plus_a = p.T_ml / 2 + p.a_sigma;
  a.two_b = 2 * p.b_sigma;
  a.st_npi = h->st_npi; a.st_pi = h->st_pi; a.st_S = h->st_S; a.st_tau = h->st_tau;
  a.st_sig = h->st_sig; a.st_lam = h->st_lam; a.st_del = h->st_del; a.st_mu = h->st_mu;
  a.edge = h->d_counts;
  a.nonempty = a.edge + (long long)h->n * h->n;
  a.cp_hist = a.nonempty + h->n;
  a.cp_kept = a.cp_hist + (long long)h->n * (h->N + 3);
  a.counters = h->d_counters;
  a.cache = h->d_cache;
  a.st_cg = h->st_cg;
  a.refresh = h->cache_valid ? 0 : 1;

  const bool trace = trace_d != nullptr;
  const bool replay = replay_d != nullptr;
  double *d_rd = nullptr, *d_td = nullptr;
  int *d_ri = nullptr, *d_ti = nullptr;
  auto cleanup = [&]() { free_dev(d_rd); free_dev(d_td); free_dev(d_ri); free_dev(d_ti); };
#define CUX(expr)                                                                \
  do {                                                                           \
    cudaError_t e_ = (expr);                                                     \
    if (e_ != cudaSuccess) {                                                     \
      cleanup();                                                                 \
      DBN_FAIL(h, DBN_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); \
    }                                                                            \
  } while (0)
  const size_t rec = (size_t)h->U * n_iter;
  if (trace) {
    CUX(cudaMalloc(&d_td, rec * DBN_TD_STRIDE * sizeof(double)));
    CUX(cudaMalloc(&d_ti, rec * DBN_TI_STRIDE * sizeof(int)));
    CUX(cudaMemsetAsync(d_td, 0xFF, rec * DBN_TD_STRIDE * sizeof(double), h->stream));  // NaN
    CUX(cudaMemsetAsync(d_ti, 0xFF, rec * DBN_TI_STRIDE * sizeof(int), h->stream));     // -1
    a.td = d_td;
    a.ti = d_ti;
  }
  if (replay) {
    CUX(cudaMalloc(&d_rd, rec * DBN_RD_STRIDE * sizeof(double)));
    CUX(cudaMalloc(&d_ri, rec * DBN_RI_STRIDE * sizeof(int)));
    CUX(cudaMemcpyAsync(d_rd, replay_d, rec * DBN_RD_STRIDE * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUX(cudaMemcpyAsync(d_ri, replay_i, rec * DBN_RI_STRIDE * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    a.rd = d_rd;
    a.ri = d_ri;
  }
  const int km = p.fan_in <= 3 ? 4 : 5;
  const bool fast = p.method == DBN_H_DBN || p.method == DBN_NH_DBN;
  const int block = fast ? FAST_BLOCK : CHAIN_BLOCK;
  const size_t smem = fast ? fast_smem_bytes(km) : chain_smem_bytes(km);
  // small unit counts are spread over more warps: `lanes` units per warp, chosen so that the warps about fill the device
  // (DBN_LANES overrides: A/B measurements in profiles/r2_small)
  int lanes = 32;
  {
    // measured (profiles/r2_small/lanes_ab.txt): the h / nh kernel gains down to one unit per warp (c2, 320 units: 8.7e6 ->
    // 1.4e7 unit-iterations/s); the coupled kernel is fastest at 16 and loses below (c3, 2816 units: 4.0e7 / 4.3e7 / 1.8e7 at
    // 32 / 16 / 4 -- the gathered table no longer stays in the L1 of a few SMs)
    const long long slots = (long long)h->sm_count * 8;   // resident warps of either kernel (2 blocks of 128 threads per SM)
    const int min_lanes = fast ? 1 : 16;
    // halve while the warps fill at most half a wave (c4 data, profiles/r2_small/lanes_c4.txt: 12 800 units are fastest at 16 per
    // warp, 25 600 and more at 32)
    while (lanes > min_lanes && (h->U + lanes / 2 - 1) / (lanes / 2) <= slots / 2) lanes /= 2;
    if (h->lanes_env) lanes = h->lanes_env;
  }
  a.lanes = lanes;
  const int upb = lanes >= 32 ? block : lanes * (block / 32);   // units per block
  const unsigned grid = (unsigned)((h->U + upb - 1) / upb);
  if (fast && a.refresh && !getenv("DBN_REFRESH_IN_KERNEL")) {   // rebuild the boundary-row cache with every row of every unit in flight
    const unsigned gf = (unsigned)(((long long)h->U * SEG_ROWS + 255) / 256);
    if (km == 4) cache_fill_kernel<4><<<gf, 256, 0, h->stream>>>(a);
    else cache_fill_kernel<5><<<gf, 256, 0, h->stream>>>(a);
    CUX(cudaGetLastError());
    a.refresh = 0;
  }
  if (fast) CUX((trace ? launch_fast_trace : launch_fast_plain)(p.method, km, grid, block, smem, h->stream, a));
  else CUX((trace ? launch_chain_trace : launch_chain_plain)(p.method, km, grid, block, smem, h->stream, a));
  h->cache_valid = true;
  h->it_done += n_iter;
  if (trace) {
    CUX(cudaMemcpyAsync(trace_d, d_td, rec * DBN_TD_STRIDE * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUX(cudaMemcpyAsync(trace_i, d_ti, rec * DBN_TI_STRIDE * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUX(cudaStreamSynchronize(h->stream));
  }
  cleanup();
#undef CUX
  return DBN_OK;
}

int dbn_run_fp(dbn_handle* h, int32_t n_iter, const double* replay_d, const int32_t* replay_i, double* trace_d, int32_t* trace_i) {
  if (!h) return DBN_ERR_INVALID;
  if (!h->have_chain || !is_fp_method(h->rp.method)) DBN_FAIL(h, DBN_ERR_STATE, "dbn_run_fp: call dbn_chain_init with a full-parents method first");
  if (n_iter < 0) DBN_FAIL(h, DBN_ERR_INVALID, "dbn_run_fp: n_iter < 0");
  if ((replay_d == nullptr) != (replay_i == nullptr)) DBN_FAIL(h, DBN_ERR_INVALID, "dbn_run_fp: replay_d and replay_i must be given together");
  if ((replay_d != nullptr) && !(trace_d && trace_i)) DBN_FAIL(h, DBN_ERR_INVALID, "dbn_run_fp: replay mode requires trace buffers");
  if ((trace_d == nullptr) != (trace_i == nullptr)) DBN_FAIL(h, DBN_ERR_INVALID, "dbn_run_fp: trace_d and trace_i must be given together");
  if (n_iter == 0) return DBN_OK;
  CU(h, cudaSetDevice(h->cfg.device));
  const dbn_run_params& p = h->rp;
  const int k = h->n;
  FpArgs a;
  memset(&a, 0, sizeof(a));
  a.tb = make_table(h);
  a.num_samples = p.num_samples; a.burn_in = p.burn_in; a.thin = p.thin; a.with_tau = p.with_tau;
  a.C = h->cfg.n_chains; a.U = h->U; a.unit0 = (int)h->unit0; a.it0 = h->it_done; a.n_iter = n_iter;
  a.chain_offset = (unsigned)h->cfg.chain_offset;
  a.seed = h->cfg.seed;
  a.T_sigma_half = p.T_sigma / 2;
  a.a_sig = p.a_sigma; a.b_sig = p.b_sigma; a.a_lam = p.a_lambda; a.b_lam = p.b_lambda;
  a.log_p = log(p.cp_p); a.log_1mp = log(1.0 - p.cp_p);
  a.c0 = ml_const(p.T_ml, p.a_sigma, p.b_sigma);
  a.T_half_plus_a = p.T_ml / 2 + p.a_sigma;
  a.two_b = 2 * p.b_sigma;
  a.st_S = h->st_S; a.st_tau = h->st_tau; a.st_sig = h->st_sig; a.st_lam = h->st_lam; a.st_mu = h->st_mu_fp;
  a.st_del = h->st_del; a.a_del = p.a_delta; a.b_del = p.b_delta;
  a.pos = h->d_counts;
  a.kept = a.pos + (long long)h->n * h->n;
  a.cp_hist = a.kept + h->n;
  a.cp_kept = a.cp_hist + (long long)h->n * (h->N + 3);
  a.neg = a.cp_kept + h->n;
  a.counters = h->d_counters;
  a.ws = h->d_fp_ws;
  a.bsc = h->d_fp_bsc;
  const bool trace = trace_d != nullptr, replay = replay_d != nullptr;
  double *d_rd = nullptr, *d_td = nullptr;
  int *d_ri = nullptr, *d_ti = nullptr;
  auto cleanup = [&]() { free_dev(d_rd); free_dev(d_td); free_dev(d_ri); free_dev(d_ti); };
#define CUX(expr)                                                                \
  do {                                                                           \
    cudaError_t e_ = (expr);                                                     \
    if (e_ != cudaSuccess) {                                                     \
      cleanup();                                                                 \
      DBN_FAIL(h, DBN_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); \
    }                                                                            \
  } while (0)
  const size_t rec = (size_t)h->U * n_iter;
  const size_t rds = fp_rd_stride(k), tds = fp_td_stride(k);
  if (trace) {
    CUX(cudaMalloc(&d_td, rec * tds * sizeof(double)));
    CUX(cudaMalloc(&d_ti, rec * 16 * sizeof(int)));
    CUX(cudaMemsetAsync(d_td, 0xFF, rec * tds * sizeof(double), h->stream));
    CUX(cudaMemsetAsync(d_ti, 0xFF, rec * 16 * sizeof(int), h->stream));
    a.td = d_td;
    a.ti = d_ti;
  }
  if (replay) {
    CUX(cudaMalloc(&d_rd, rec * rds * sizeof(double)));
    CUX(cudaMalloc(&d_ri, rec * 4 * sizeof(int)));
    CUX(cudaMemcpyAsync(d_rd, replay_d, rec * rds * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUX(cudaMemcpyAsync(d_ri, replay_i, rec * 4 * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    a.rd = d_rd;
    a.ri = d_ri;
  }
  const size_t smem = fp_smem_bytes(k, h->d_fp_ws == nullptr);
  const int fpm = p.method == DBN_FP_GLOB_DBN ? FPM_GLOB : (p.method == DBN_FP_VV_DBN ? FPM_VV : (p.method == DBN_FP_SEQ_DBN ? FPM_SEQ : FPM_PLAIN));
  const bool big = h->d_fp_ws != nullptr;
  const int fp_block = fp_block_threads(k, !big);
  CUX((trace ? (big ? launch_fp_trace_big : launch_fp_trace_small) : (big ? launch_fp_plain_big : launch_fp_plain_small))(
      fpm, (unsigned)h->fp_grid, fp_block, smem, h->stream, a));
  h->it_done += n_iter;
  if (trace) {
    CUX(cudaMemcpyAsync(trace_d, d_td, rec * tds * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUX(cudaMemcpyAsync(trace_i, d_ti, rec * 16 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUX(cudaStreamSynchronize(h->stream));
  }
  cleanup();
#undef CUX
  return DBN_OK;
}

int dbn_counts_size(const dbn_handle* h, int64_t* n_elems) {
  if (!h || !n_elems) return DBN_ERR_INVALID;
  if (!h->have_chain) return DBN_ERR_STATE;
  *n_elems = h->counts_elems;
  return DBN_OK;
}

int dbn_read_counts(dbn_handle* h, int64_t* out) {
  if (!h || !out) return DBN_ERR_INVALID;
  if (!h->have_chain) DBN_FAIL(h, DBN_ERR_STATE, "dbn_read_counts: call dbn_chain_init first");
  CU(h, cudaSetDevice(h->cfg.device));
  const size_t bytes = (size_t)h->counts_elems * sizeof(int64_t);
  {
    // a PINNED (or registered) destination is written by the copy engine directly: one transfer, no host memcpy
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, out) == cudaSuccess && at.type == cudaMemoryTypeHost) {
      CU(h, cudaMemcpyAsync(out, h->d_counts, bytes, cudaMemcpyDeviceToHost, h->stream));
      CU(h, cudaStreamSynchronize(h->stream));
      return DBN_OK;
    }
    cudaGetLastError();
  }
  // pageable destination: through a pinned bounce buffer owned by the handle (a device -> pageable copy runs at a fraction of
  // the link rate)
  if (h->pinned_bytes < bytes) {
    if (h->pinned) cudaFreeHost(h->pinned);
    h->pinned = nullptr;
    h->pinned_bytes = 0;
    if (cudaMallocHost(&h->pinned, bytes) == cudaSuccess) h->pinned_bytes = bytes;
    else cudaGetLastError();  // no pinned memory: fall back to the direct copy below
  }
  if (h->pinned_bytes >= bytes) {
    CU(h, cudaMemcpyAsync(h->pinned, h->d_counts, bytes, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    memcpy(out, h->pinned, bytes);
  } else {
    CU(h, cudaMemcpyAsync(out, h->d_counts, bytes, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
  }
  return DBN_OK;
}

int dbn_device_counts(dbn_handle* h, void** device_ptr) {
  if (!h || !device_ptr) return DBN_ERR_INVALID;
  if (!h->have_chain) DBN_FAIL(h, DBN_ERR_STATE, "dbn_device_counts: call dbn_chain_init first");
  *device_ptr = h->d_counts;
  return DBN_OK;
}

int dbn_read_state(dbn_handle* h, int32_t* pi, int32_t* pi_len, int32_t* tau, int32_t* tau_len, double* scalars, double* mu) {
  if (!h) return DBN_ERR_INVALID;
  if (!h->have_chain) DBN_FAIL(h, DBN_ERR_STATE, "dbn_read_state: call dbn_chain_init first");
  CU(h, cudaSetDevice(h->cfg.device));
  const size_t U = (size_t)h->U;
  CU(h, cudaStreamSynchronize(h->stream));
  std::vector<int> ti(U * SMAX);
  std::vector<double> td(U * KMAX);
  if (pi) {
    CU(h, cudaMemcpy(ti.data(), h->st_pi, U * PMAX * sizeof(int), cudaMemcpyDeviceToHost));
    for (size_t u = 0; u < U; ++u)
      for (int j = 0; j < PMAX; ++j) pi[u * PMAX + j] = ti[j * U + u];
  }
  if (pi_len) CU(h, cudaMemcpy(pi_len, h->st_npi, U * sizeof(int), cudaMemcpyDeviceToHost));
  if (tau) {
    CU(h, cudaMemcpy(ti.data(), h->st_tau, U * SMAX * sizeof(int), cudaMemcpyDeviceToHost));
    for (size_t u = 0; u < U; ++u)
      for (int j = 0; j < SMAX; ++j) tau[u * SMAX + j] = ti[j * U + u];
  }
  if (tau_len) CU(h, cudaMemcpy(tau_len, h->st_S, U * sizeof(int), cudaMemcpyDeviceToHost));
  if (scalars) {
    std::vector<double> s(U);
    double* src[3] = {h->st_sig, h->st_lam, h->st_del};
    for (int c = 0; c < 3; ++c) {
      CU(h, cudaMemcpy(s.data(), src[c], U * sizeof(double), cudaMemcpyDeviceToHost));
      for (size_t u = 0; u < U; ++u) scalars[u * 3 + c] = s[u];
    }
  }
  if (mu) {
    CU(h, cudaMemcpy(td.data(), h->st_mu, U * KMAX * sizeof(double), cudaMemcpyDeviceToHost));
    for (size_t u = 0; u < U; ++u)
      for (int i = 0; i < KMAX; ++i) mu[u * KMAX + i] = td[i * U + u];
  }
  return DBN_OK;
}

int dbn_read_counters(dbn_handle* h, uint64_t out[8]) {
  if (!h || !out) return DBN_ERR_INVALID;
  CU(h, cudaSetDevice(h->cfg.device));
  CU(h, cudaMemcpyAsync(out, h->d_counters, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
#ifdef DBN_DIAG_SPLIT2
  {
    unsigned long long d[16];
    CU(h, cudaMemcpy(d, h->d_counters + 8, sizeof(d), cudaMemcpyDeviceToHost));
    fprintf(stderr, "split-warp probes:");
    for (int i = 0; i < 10; ++i) fprintf(stderr, " %llu", d[i]);
    fprintf(stderr, "  check code %llu", d[15]);
    fprintf(stderr, "\n");
    CU(h, cudaMemset(h->d_counters + 8, 0, sizeof(d)));
  }
#endif
  return DBN_OK;
}

int dbn_verify_cache(dbn_handle* h, uint64_t out[2]) {
  if (!h || !out) return DBN_ERR_INVALID;
  if (!h->have_chain) DBN_FAIL(h, DBN_ERR_STATE, "dbn_verify_cache: call dbn_chain_init first");
  out[0] = out[1] = 0;
  // only the h / nh kernel keeps a cache across launches (the coupled kernel refills its rows at every launch start);
  // nothing is cached before the first dbn_run
  if (!h->d_cache || !h->cache_valid || !(h->rp.method == DBN_H_DBN || h->rp.method == DBN_NH_DBN)) return DBN_OK;
  CU(h, cudaSetDevice(h->cfg.device));
  ChainArgs a;
  memset(&a, 0, sizeof(a));
  a.tb = make_table(h);
  a.C = h->cfg.n_chains;
  a.U = h->U;
  a.unit0 = (int)h->unit0;
  a.n_genes = h->n;
  a.lag = h->lag;
  a.st_S = h->st_S; a.st_tau = h->st_tau; a.st_cg = h->st_cg; a.cache = h->d_cache;
  unsigned long long* d = nullptr;
  CU(h, cudaMalloc(&d, 2 * sizeof(unsigned long long)));
  cudaError_t e = cudaMemsetAsync(d, 0, 2 * sizeof(unsigned long long), h->stream);
  if (e == cudaSuccess) {
    const unsigned grid = (unsigned)((h->U + 127) / 128);
    if (h->rp.fan_in <= 3) cache_verify_kernel<4><<<grid, 128, 0, h->stream>>>(a, d);
    else cache_verify_kernel<5><<<grid, 128, 0, h->stream>>>(a, d);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d);
  if (e != cudaSuccess) DBN_FAIL(h, DBN_ERR_CUDA, "dbn_verify_cache failed: %s", cudaGetErrorString(e));
  return DBN_OK;
}

int dbn_reset_counters(dbn_handle* h) {
  if (!h) return DBN_ERR_INVALID;
  CU(h, cudaSetDevice(h->cfg.device));
  CU(h, cudaMemsetAsync(h->d_counters, 0, 8 * sizeof(uint64_t), h->stream));
  return DBN_OK;
}

int dbn_measure_fp64_peak(dbn_handle* h, double* gflops) {
  if (!h || !gflops) return DBN_ERR_INVALID;
  CU(h, cudaSetDevice(h->cfg.device));
  cudaDeviceProp prop;
  CU(h, cudaGetDeviceProperties(&prop, h->cfg.device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
  double* d_out = nullptr;
  CU(h, cudaMalloc(&d_out, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, h->stream);
    fp64_peak_kernel<<<blocks, threads, 0, h->stream>>>(d_out, iters, 1.0 + rep);
    cudaEventRecord(e1, h->stream);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double gf = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e9;
    if (rep > 0 && gf > best) best = gf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  CU(h, cudaGetLastError());
  *gflops = best;
  return DBN_OK;
}

}  // extern "C"
