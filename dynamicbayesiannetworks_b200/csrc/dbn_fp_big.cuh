// Large-k routines of kernel 4 (full-parents chains with the k x k matrices in a per-block global workspace; BASELINE
// config 5, k = 500): Cholesky factor, inverse factor and the rank-k update M += I - A^-1 as TILED level-3 steps on the
// FP64 tensor cores (mma.sync.m8n8k4.f64, SASS DMMA.8x8x4).
//
// Round-1 version (dbn_fp.cuh history): panels of 16 factor rows with the operands loaded into registers behind block
// barriers; ncu (profiles/r1_v13/c5_fp_glob_ncu_full.csv): FP64 pipe 14 % active, 5.5 long-scoreboard + 4.3 barrier stall
// cycles per issued instruction -- every step waited for its own loads.  Here every level-3 step is the same building
// block, a 64 x 128 output tile held by 512 threads (16 warps x 16 x 32), whose two operand streams are staged in shared
// memory by cp.async through a four-deep ring (16 factor rows per stage), so the loads of three stages are in flight while
// the fourth is consumed.  Why DMMA and not DFMA (profiles/r2_c5): the two have the same peak on this part (36.9 vs 36.4
// TFLOP/s measured, scripts/dmma_peak.cu), but a 4 x 4 register block fed from shared memory needs four 128-bit loads per
// 16 FMAs and a 128-bit load occupies the shared-memory pipe for four quarter-warp phases: 16 warps x 24 cycles of
// shared-memory time against 128 cycles of FP64 time per factor row, i.e. bound by shared memory at 25 % of the FP64 peak
// -- which is what the first tiled version measured (fpb_check_v0_dfma_4x4.txt: 24 %).  One DMMA does 256 FMAs from two
// 64-bit loads per lane, and six loads feed eight of them.
//
// Storage (row-major, leading dimension ld = k rounded up to 8 doubles so that every tile row starts on 16 bytes):
//   A  full symmetric input; on exit the factor U = L^T (A = U^T U) in the UPPER triangle, the lower triangle untouched
//   W  = L^-1, LOWER triangular (W[i][a], a <= i).  Both are "K-slow": for a fixed factor row the 64 / 128 operand
//      entries of a tile are contiguous, so the two operands of every product below are rows of a row-major matrix:
//   potrf   R[a][c]  = A[p0+a][c]  - sum_{j < p0}        U[j][p0+a] U[j][c]        (block row p0 of the factor)
//   trinv   C[a][b]  =               sum_{b0 <= j < p0}  U[j][p0+a] W[j][b]        (block row p0 of L^-1 = -Li C)
//   syrk    M[a][b] += delta_ab    - sum_{j >= max(a0,b0)} W[j][a]  W[j][b]        (I - A^-1 = I - W^T W)
//   followed, for the first two, by the 64 x 64 triangular solve with the panel's diagonal block, done as a product
//   with the inverted block (both operands in shared memory, the same DMMA loop).  The diagonal block itself is
//   factorised and inverted by sub-panels of 16 columns (see fpb_factor).
// Entries of W right of the diagonal inside a tile read as zero: the diagonal block is written with its zeros and the
// block right of an even-numbered diagonal block is cleared, so no K range has to be cut per column.
#pragma once

namespace dbn {

constexpr int FPB_THREADS = 512;
#ifndef DBN_FPB_KD
#define DBN_FPB_KD 32   // 32 rows x 2 stages measured 7 % faster than 16 x 4 on the factorisation (half the barriers), profiles/r2_c5/ring_ab.txt
#endif
#ifndef DBN_FPB_NSTAGE
#define DBN_FPB_NSTAGE 2
#endif
constexpr int FPB_TM = 64, FPB_TN = 128, FPB_KD = DBN_FPB_KD, FPB_NSTAGE = DBN_FPB_NSTAGE;
static_assert(FPB_KD % 16 == 0 && FPB_KD * FPB_NSTAGE == 64, "ring = 64 factor rows (it also holds the staged tile and the diagonal block)");
// row strides in shared memory: = 4 mod 16 doubles, so that the 4 (factor rows) x 8 (entries) operand fragment of a DMMA
// is read by each half-warp from 32 different banks
constexpr int FPB_LDX = FPB_TM + 4, FPB_LDY = FPB_TN + 4;
constexpr int FPB_STAGE = FPB_KD * (FPB_LDX + FPB_LDY);   // doubles per ring stage: X[KD][68] | Y[KD][132]
constexpr int FPB_PIPE = FPB_NSTAGE * FPB_STAGE;          // 12800 doubles; between products it holds the tile R[64][132] and U_pp[64][64]
constexpr int FPB_OFF_US = 64 * FPB_LDY;                  // 8448
static_assert(FPB_OFF_US + 64 * FPB_LDX <= FPB_PIPE, "tile staging and the diagonal block fit the ring");
constexpr int FPB_OFF_LI = FPB_PIPE;                      // inverted diagonal block li[c][a] = Li[a][c], row stride FPB_LDX
constexpr int FPB_OFF_COL = FPB_OFF_LI + 64 * FPB_LDX;    // two slots of (pivot column of D | row of X)
constexpr int FPB_OFF_DINV = FPB_OFF_COL + 256 + 16 * 17;   // (after the 16 x 16 inverse of a sub-panel's diagonal block) reciprocal pivots of the block
constexpr int FPB_OFF_RED = FPB_OFF_DINV + 64;            // reduction scratch of the level-2 routines
constexpr int FPB_SMEM_DOUBLES = FPB_OFF_RED + FPB_THREADS + 32;

__host__ __device__ inline int fp_ld_big(int k) { return (k + 7) & ~7; }

__device__ __forceinline__ void fpb_cp16(double* smem, const double* gmem, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int nbytes = valid ? 16 : 0;   // 0: nothing is read, the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void fpb_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void fpb_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// D(8 x 8) += A(8 x 4) B(4 x 8): lane holds A[lane / 4][lane % 4], B[lane % 4][lane / 4], D[lane / 4][2 (lane % 4) + {0, 1}]
__device__ __forceinline__ void fpb_dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// The warp's 16 x 32 part of the 64 x 128 tile as 2 x 4 DMMA accumulators; element (i, j, e) is tile row
// a + 8 i, column b + 8 j + e.
struct FpbCoord {
  int a, b;     // first row / column of this LANE's elements: warp origin + lane / 4, warp origin + 2 (lane % 4)
  int fa, fb;   // operand fragment offsets: (lane % 4) rows down, warp origin + lane / 4 along
  int kr;       // lane % 4
  int wa, wb;   // first row / column of the warp
};
__device__ __forceinline__ FpbCoord fpb_coord() {
  // warp w: rows 16 (w % 4), columns 32 (w / 4): the four warps of a scheduler (w % 4 fixed) cover all four column groups, so
  // tiles whose right half is masked (fpb_gemm's `ncols`) halve the work of every scheduler
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  FpbCoord c;
  c.wa = (w & 3) * 16;
  c.wb = (w >> 2) * 32;
  c.a = c.wa + (lane >> 2);
  c.b = c.wb + 2 * (lane & 3);
  c.fa = c.wa + (lane >> 2);
  c.fb = c.wb + (lane >> 2);
  c.kr = lane & 3;
  return c;
}

typedef double FpbAcc[2][4][2];
__device__ __forceinline__ void fpb_zero(FpbAcc& acc) {
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
}

// acc += X^T Y over `nk4` groups of four factor rows, both operands in shared memory (row strides ldx, ldy)
__device__ __forceinline__ void fpb_mma_rows(FpbAcc& acc, const FpbCoord& co, const double* xs, int ldx, const double* ys, int ldy, int nk4) {
  const double* xp = xs + co.kr * ldx + co.fa;
  const double* yp = ys + co.kr * ldy + co.fb;
#pragma unroll 4
  for (int s = 0; s < nk4; ++s) {
    double a[2], b[4];
#pragma unroll
    for (int i = 0; i < 2; ++i) a[i] = xp[s * 4 * ldx + 8 * i];
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = yp[s * 4 * ldy + 8 * j];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) fpb_dmma(acc[i][j], a[i], b[j]);
  }
}

// acc[a][b] += sum_{j_lo <= j < j_hi} X[j][xc0 + a] * Y[j][yc0 + b]; columns >= ncol (even) read as zero
__device__ inline void fpb_gemm(FpbAcc& acc, const FpbCoord& co, const double* __restrict__ X, int xc0, const double* __restrict__ Y, int yc0,
                                int j_lo, int j_hi, int ld, int ncol, double* pipe, int ncols = FPB_TN) {   // ncols: tile columns that are needed (64 or 128)
  const int tid = threadIdx.x;
  const int nst = (j_hi - j_lo + FPB_KD - 1) / FPB_KD;
  const int xr = tid >> 5, xc = (tid & 31) * 2;   // X stage: 16 rows x 32 chunks of 16 bytes
  const int yr = tid >> 6, yc = (tid & 63) * 2;   // Y stage: 16 rows x 64 chunks, two per thread (rows yr and yr + 8)
  const bool xv = xc0 + xc < ncol, yv = yc0 + yc < ncol;
  const double* xsrc = X + (xv ? xc0 + xc : 0);
  const double* ysrc = Y + (yv ? yc0 + yc : 0);
  auto load = [&](int s) {
    double* buf = pipe + (s % FPB_NSTAGE) * FPB_STAGE;
    const int j = j_lo + s * FPB_KD;
#pragma unroll
    for (int h = 0; h < FPB_KD / 16; ++h) {
      const int r = xr + 16 * h;
      const bool v = xv && j + r < j_hi;
      fpb_cp16(buf + r * FPB_LDX + xc, xsrc + (long long)(v ? j + r : j_lo) * ld, v);
    }
#pragma unroll
    for (int h = 0; h < FPB_KD / 8; ++h) {
      const int r = yr + 8 * h;
      const bool v = yv && j + r < j_hi;
      fpb_cp16(buf + FPB_KD * FPB_LDX + r * FPB_LDY + yc, ysrc + (long long)(v ? j + r : j_lo) * ld, v);
    }
  };
  __syncthreads();   // the ring is free (previous product / tile staging of every thread is complete)
#pragma unroll
  for (int s = 0; s < FPB_NSTAGE - 1; ++s) {
    if (s < nst) load(s);
    fpb_commit();
  }
  for (int s = 0; s < nst; ++s) {
    fpb_wait<FPB_NSTAGE - 2>();
    __syncthreads();   // stage s has landed for every thread; stage s - 1 has been consumed by every thread
    if (s + FPB_NSTAGE - 1 < nst) load(s + FPB_NSTAGE - 1);
    fpb_commit();
    const double* buf = pipe + (s % FPB_NSTAGE) * FPB_STAGE;
    if (co.wb < ncols) fpb_mma_rows(acc, co, buf, FPB_LDX, buf + FPB_KD * FPB_LDX, FPB_LDY, FPB_KD / 4);   // warp-uniform
  }
}

// stage the tile held in `acc` as R[64][FPB_LDY] at the start of the ring (the caller synchronises before it is read)
__device__ __forceinline__ void fpb_stage_tile(const FpbAcc& acc, const FpbCoord& co, double* pipe) {
  __syncthreads();   // the ring is no longer read by any product
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<double2*>(pipe + (co.a + 8 * i) * FPB_LDY + co.b + 8 * j) = make_double2(acc[i][j][0], acc[i][j][1]);
}

// acc = Li R for the staged tile R; Li = inverse of the panel's lower diagonal block, li[c * LDX + a] = Li[a][c] (zero for c > a)
__device__ __forceinline__ void fpb_apply_li(FpbAcc& acc, const FpbCoord& co, const double* li, const double* pipe) {
  fpb_zero(acc);
  fpb_mma_rows(acc, co, li, FPB_LDX, pipe, FPB_LDY, (co.wa + 16) / 4);   // warp-uniform: Li[a][c] = 0 beyond the warp's last row
}

// A = U^T U in place (see the header), idiag[j] = 1 / U[j][j]; W != nullptr: also W = L^-1.  Returns ln det A.
// full_w == false: only the 64 x 64 diagonal blocks of W (the inverted diagonal blocks of L) are written -- what the blocked
// triangular solves below need; the k^3 / 3 of the full inverse is skipped.
static __device__ __noinline__ double fpb_factor(double* A, double* W, int ld, int k, double* idiag, double* sm, bool full_w = true) {
  __builtin_assume(__isShared(sm));   // generic pointer across the call: without the hint every access below is a generic LD / ST
  __builtin_assume(__isShared(idiag));
  double* const pipe = sm;
  double* const us = sm + FPB_OFF_US;    // U_pp (row c, column a), alive between the diagonal step and the next product
  double* const li = sm + FPB_OFF_LI;
  double* const colbuf = sm + FPB_OFF_COL;
  double* const dinv = sm + FPB_OFF_DINV;
  const int tid = threadIdx.x;
  const FpbCoord co = fpb_coord();
  const int ncol = (k + 1) & ~1;
  double mant = 1.0;
  int expo = 0;
  for (int p0 = 0; p0 < k; p0 += 64) {
    for (int c0 = p0; c0 < k; c0 += FPB_TN) {
      FpbAcc acc, ain;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int row = p0 + co.a + 8 * i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = c0 + co.b + 8 * j;   // even; ld is even: the pair is inside the row and 16-byte aligned
          double2 v = make_double2(0.0, 0.0);
          if (row < k && col < ncol) v = *reinterpret_cast<const double2*>(A + (long long)row * ld + col);
          ain[i][j][0] = v.x;
          ain[i][j][1] = v.y;
        }
      }
      fpb_zero(acc);
      fpb_gemm(acc, co, A, p0, A, c0, 0, p0, ld, ncol, pipe, c0 + 64 >= k ? 64 : FPB_TN);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e)   // entries past k are cleared: the pad row / column of the workspace holds anything, and 0 x NaN must not reach the solve
            acc[i][j][e] = (p0 + co.a + 8 * i < k && c0 + co.b + 8 * j + e < k) ? ain[i][j][e] - acc[i][j][e] : 0.0;
      fpb_stage_tile(acc, co, pipe);
      __syncthreads();
      if (c0 == p0) {
        // ---- diagonal block D (staged, read-only) -> U_pp in `us` (us[c][a] = L[a][c]) and Li = L_pp^-1 in `li` (li[c][a] =
        // Li[a][c]).  Left-looking by sub-panels of 16 columns: thread (r, g) = (tid / 8, tid % 8) holds row r, columns
        // 16 q + g + 8 e of the sub-panel in registers; a column step broadcasts the pivot column through a double-buffered
        // slot (one barrier per column, two elements per thread).  The four warps that hold the sub-panel's diagonal 16 x 16
        // block apply the same column operations to an identity block, which leaves T = L_qq^-1; X = Li then follows from
        // X_q <- T X_q, X_i <- X_i - L_iq X_q (i > q), 16-term sums on at most two entries per thread.
        const int r = tid >> 3, g = tid & 7;
        double* const tq = colbuf + 256;   // [16][17]
        for (int e = tid; e < 64 * 64; e += FPB_THREADS) li[(e >> 6) * FPB_LDX + (e & 63)] = ((e >> 6) == (e & 63)) ? 1.0 : 0.0;
        double prod = 1.0;
        for (int q = 0; q < 4; ++q) {
          const int q0 = 16 * q;
          const bool inq = (r >> 4) == q;   // warp-uniform (a warp holds four rows)
          double d[2], t2[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int cc = q0 + g + 8 * e;
            double v = (p0 + r < k && p0 + cc < k) ? pipe[r * FPB_LDY + cc] : (r == cc ? 1.0 : 0.0);
            double v1 = 0.0, v2 = 0.0, v3 = 0.0;   // four partial sums: the 16 q products are not one dependent chain
            for (int m = 0; m < q0; m += 4) {
              v = fma(-us[m * FPB_LDX + r], us[m * FPB_LDX + cc], v);
              v1 = fma(-us[(m + 1) * FPB_LDX + r], us[(m + 1) * FPB_LDX + cc], v1);
              v2 = fma(-us[(m + 2) * FPB_LDX + r], us[(m + 2) * FPB_LDX + cc], v2);
              v3 = fma(-us[(m + 3) * FPB_LDX + r], us[(m + 3) * FPB_LDX + cc], v3);
            }
            v = (v + v1) + (v2 + v3);
            d[e] = v;
            t2[e] = (r == cc) ? 1.0 : 0.0;
          }
#pragma unroll 1
          for (int s_ = 0; s_ < 16; ++s_) {
            const int c = q0 + s_;
            double* cb = colbuf + (c & 1) * 128;
            double* xb = cb + 64;
            if ((c & 7) == g) cb[r] = (s_ & 8) ? d[1] : d[0];
            if (r == c) {
              xb[g] = t2[0];
              xb[g + 8] = t2[1];
            }
            __syncthreads();
            const double dp = cb[c];
            const double rinv = rsqrt(dp);
            prod *= rinv;
            if ((c & 7) == 7) {   // ln det A = -2 ln(product of the reciprocal pivots), kept as mantissa x 2^expo
              int ex;
              mant = frexp(mant * prod, &ex);
              expo += ex;
              prod = 1.0;
            }
            const double la = (r > c) ? cb[r] * rinv : 0.0;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int cc = q0 + g + 8 * e;
              const double lb = (cc > c) ? cb[cc] * rinv : 0.0;
              d[e] = fma(-la, lb, d[e]);
            }
            if (inq) {
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const double xs = xb[g + 8 * e] * rinv;
                t2[e] = (r == c) ? xs : fma(-la, xs, t2[e]);
              }
            }
            if (tid < 64) us[c * FPB_LDX + tid] = tid > c ? cb[tid] * rinv : (tid == c ? dp * rinv : 0.0);
            if (tid == 64) dinv[c] = rinv;
          }
          if (inq) {
            tq[(r - q0) * 17 + g] = t2[0];
            tq[(r - q0) * 17 + g + 8] = t2[1];
          }
          __syncthreads();
          // X_q <- T X_q over the columns j < 16 (q + 1) (X_qq was the identity); entry (i, j) at li[j][q0 + i]
          double xn[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int o = tid + h * FPB_THREADS, i = o & 15, j = o >> 4;
            double v = 0.0;
            if (j < q0 + 16)
              for (int m = 0; m < 16; ++m) v = fma(tq[i * 17 + m], li[j * FPB_LDX + q0 + m], v);
            xn[h] = v;
          }
          __syncthreads();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int o = tid + h * FPB_THREADS, i = o & 15, j = o >> 4;
            if (j < q0 + 16) li[j * FPB_LDX + q0 + i] = xn[h];
          }
          __syncthreads();
          // X_i <- X_i - L_iq X_q for the rows below the sub-panel
          const int nr = 48 - q0;   // rows below
          for (int o = tid; o < nr * (q0 + 16); o += FPB_THREADS) {
            const int i = q0 + 16 + o % nr, j = o / nr;
            double v = li[j * FPB_LDX + i];
            for (int m = 0; m < 16; ++m) v = fma(-us[(q0 + m) * FPB_LDX + i], li[j * FPB_LDX + q0 + m], v);
            li[j * FPB_LDX + i] = v;
          }
        }
        __syncthreads();
        for (int e = tid; e < 64 * 64; e += FPB_THREADS) {
          const int c = e >> 6, a = e & 63;
          if (p0 + c < k && p0 + a < k) {
            if (a >= c) A[(long long)(p0 + c) * ld + p0 + a] = us[c * FPB_LDX + a];
            if (W) {
              W[(long long)(p0 + c) * ld + p0 + a] = li[a * FPB_LDX + c];   // row p0 + c of L^-1, zeros right of the diagonal
              if ((p0 & 64) == 0 && p0 + 64 + a < ld) W[(long long)(p0 + c) * ld + p0 + 64 + a] = 0.0;
            }
          }
        }
        if (tid < 64 && p0 + tid < k) idiag[p0 + tid] = dinv[tid];
      }
      // the solve, skipped (warp-uniform, no barrier inside) by the warps whose columns are the diagonal block or lie past k
      const bool keep = (c0 != p0 || co.wb >= 64) && c0 + co.wb < k;
      if (keep) {
        fpb_apply_li(acc, co, li, pipe);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int row = p0 + co.a + 8 * i;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int col = c0 + co.b + 8 * j;
            if (row < k && col + 1 < k) *reinterpret_cast<double2*>(A + (long long)row * ld + col) = make_double2(acc[i][j][0], acc[i][j][1]);
            else if (row < k && col < k) A[(long long)row * ld + col] = acc[i][j][0];
          }
        }
      }
    }
    if (W && full_w) {
      // ---- block row p0 of L^-1 left of the diagonal block: -Li (U[b0..p0, p0..]^T W[b0..p0, b0..])
      for (int b0 = 0; b0 < p0; b0 += FPB_TN) {
        FpbAcc acc;
        fpb_zero(acc);
        fpb_gemm(acc, co, A, p0, W, b0, b0, p0, ld, ncol, pipe, b0 + 64 >= p0 ? 64 : FPB_TN);   // columns >= p0 are the diagonal block
#pragma unroll
        for (int i = 0; i < 2; ++i)
          if (p0 + co.a + 8 * i >= k) {
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
          }
        fpb_stage_tile(acc, co, pipe);
        __syncthreads();
        if (b0 + co.wb < p0) {   // warp-uniform
          fpb_apply_li(acc, co, li, pipe);
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int row = p0 + co.a + 8 * i;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int col = b0 + co.b + 8 * j;   // p0 is even: a pair is never cut by the diagonal block
              if (row < k && col < p0) *reinterpret_cast<double2*>(W + (long long)row * ld + col) = make_double2(-acc[i][j][0], -acc[i][j][1]);
            }
          }
        }
      }
    }
  }
  __syncthreads();
  return -2.0 * (log(mant) + (double)expo * 0.69314718055994530942);
}

// Blocked triangular solves with two right-hand sides at once (x0, x1 in shared memory, overwritten by the solutions), from the
// factor U (upper triangle of A) and the inverted diagonal blocks Li_pp (diagonal blocks of W): one pass over the factor
// serves both vectors.
//   forward:  x <- L^-1 x.  Panel by panel: x_p <- Li_pp x_p, then every later entry i loses sum_{j in p} U[j][i] x_j (thread per
//   entry: the rows of U are read coalesced).
static __device__ __noinline__ void fpb_solve_fwd2(const double* A, const double* W, int ld, int k, double* x0, double* x1, double* sm) {
  __builtin_assume(__isShared(sm));
  __builtin_assume(__isShared(x0));
  __builtin_assume(__isShared(x1));
  double* red = sm + FPB_OFF_RED;   // [512] partial sums of x0, [.. + 512) of x1 live in the pivot-column slots + li tail: use the ring instead
  double* red1 = sm;                // the ring is free between products
  const int tid = threadIdx.x;
  for (int p0 = 0; p0 < k; p0 += 64) {
    // diagonal solve: row a = tid / 8, the eight threads of a row split its entries
    {
      const int a = tid >> 3, g = tid & 7;
      double s0 = 0.0, s1 = 0.0;
      if (p0 + a < k) {
        const double* w = W + (long long)(p0 + a) * ld + p0;
        for (int c = g; c <= a; c += 8) {
          const double l = w[c];
          s0 = fma(l, x0[p0 + c], s0);
          s1 = fma(l, x1[p0 + c], s1);
        }
      }
      __syncthreads();   // every thread has read x_p of the previous update
      red[tid] = s0;
      red1[tid] = s1;
      __syncthreads();
      if (tid < 128) {
        const int a2 = tid & 63;
        const double* r = (tid < 64 ? red : red1) + a2 * 8;
        const double v = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        if (p0 + a2 < k) (tid < 64 ? x0 : x1)[p0 + a2] = v;
      }
      __syncthreads();
    }
    // update of the entries behind the panel
    const int nb = min(64, k - p0);
    for (int i = p0 + 64 + tid; i < k; i += FPB_THREADS) {
      const double* u = A + (long long)p0 * ld + i;
      double s0[2] = {0.0, 0.0}, s1[2] = {0.0, 0.0};
      int j = 0;
      for (; j + 8 <= nb; j += 8) {
        double v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = u[(long long)(j + q) * ld];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          s0[q & 1] = fma(v[q], x0[p0 + j + q], s0[q & 1]);
          s1[q & 1] = fma(v[q], x1[p0 + j + q], s1[q & 1]);
        }
      }
      for (; j < nb; ++j) {
        const double v = u[(long long)j * ld];
        s0[0] = fma(v, x0[p0 + j], s0[0]);
        s1[0] = fma(v, x1[p0 + j], s1[0]);
      }
      x0[i] -= s0[0] + s0[1];
      x1[i] -= s1[0] + s1[1];
    }
    __syncthreads();   // the next panel's entries are final
  }
}

//   backward: x <- L^-T x = U^-1 x.  Panels bottom-up: x_p loses U[p, behind p] x_behind (16 threads per row), then x_p <- Li_pp^T x_p.
static __device__ __noinline__ void fpb_solve_bwd2(const double* A, const double* W, int ld, int k, double* x0, double* x1, double* sm) {
  __builtin_assume(__isShared(sm));
  __builtin_assume(__isShared(x0));
  __builtin_assume(__isShared(x1));
  double* red = sm + FPB_OFF_RED;
  double* red1 = sm;
  const int tid = threadIdx.x;
  for (int p0 = ((k - 1) / 64) * 64; p0 >= 0; p0 -= 64) {
    // row sums over the columns behind the panel: 16 threads per row, 32 rows per pass
    for (int r0 = 0; r0 < 64; r0 += FPB_THREADS / 16) {
      const int a = r0 + (tid >> 4), sl = tid & 15;
      double s0 = 0.0, s1 = 0.0;
      if (p0 + a < k) {
        const double* u = A + (long long)(p0 + a) * ld;
        for (int j = p0 + 64 + sl; j < k; j += 16) {
          const double v = u[j];
          s0 = fma(v, x0[j], s0);
          s1 = fma(v, x1[j], s1);
        }
      }
      __syncthreads();
      red[tid] = s0;
      red1[tid] = s1;
      __syncthreads();
      if (tid < 64) {
        const int a2 = r0 + (tid & 31);
        const double* r = (tid < 32 ? red : red1) + (tid & 31) * 16;
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < 16; ++q) v += r[q];
        if (p0 + a2 < k) (tid < 32 ? x0 : x1)[p0 + a2] -= v;
      }
    }
    __syncthreads();
    // diagonal solve with Li_pp^T: entry a = tid % 64, the eight groups tid / 64 split the rows c >= a
    {
      const int a = tid & 63, g = tid >> 6;
      double s0 = 0.0, s1 = 0.0;
      if (p0 + a < k) {
        for (int c = a + g; c < 64 && p0 + c < k; c += 8) {
          const double l = W[(long long)(p0 + c) * ld + p0 + a];
          s0 = fma(l, x0[p0 + c], s0);
          s1 = fma(l, x1[p0 + c], s1);
        }
      }
      __syncthreads();
      red[tid] = s0;
      red1[tid] = s1;
      __syncthreads();
      if (tid < 128) {
        const int a2 = tid & 63;
        const double* r = (tid < 64 ? red : red1) + a2;
        const double v = ((r[0] + r[64]) + (r[128] + r[192])) + ((r[256] + r[320]) + (r[384] + r[448]));
        if (p0 + a2 < k) (tid < 64 ? x0 : x1)[p0 + a2] = v;
      }
      __syncthreads();
    }
  }
}

// M += scale (I - W^T W) = scale (I - A^-1), full symmetric M
static __device__ __noinline__ void fpb_syrk(const double* W, double* M, int ld, int k, double* sm, double scale) {
  __builtin_assume(__isShared(sm));
  const FpbCoord co = fpb_coord();
  const int ncol = (k + 1) & ~1;
  for (int a0 = 0; a0 < k; a0 += FPB_TM) {
    for (int b0 = 0; b0 <= a0 + FPB_TM - 1 && b0 < k; b0 += FPB_TN) {
      FpbAcc acc, mold;   // the old M entries are requested before the product, so their latency is hidden by it
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int ra = a0 + co.a + 8 * i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int rb = b0 + co.b + 8 * j;
          double2 v = make_double2(0.0, 0.0);
          if (ra < k && rb <= ra) v = *reinterpret_cast<const double2*>(M + (long long)ra * ld + rb);
          mold[i][j][0] = v.x;
          mold[i][j][1] = v.y;
        }
      }
      fpb_zero(acc);
      fpb_gemm(acc, co, W, a0, W, b0, max(a0, b0), k, ld, ncol, sm, (b0 == a0 || b0 + 64 >= k) ? 64 : FPB_TN);   // b <= a only
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int ra = a0 + co.a + 8 * i;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int rb = b0 + co.b + 8 * j + e;
            if (ra < k && rb <= ra) {
              const double v = mold[i][j][e] + scale * (((ra == rb) ? 1.0 : 0.0) - acc[i][j][e]);
              M[(long long)ra * ld + rb] = v;
              if (ra != rb) M[(long long)rb * ld + ra] = v;   // M is symmetric: the mirrored entry needs no read
            }
          }
      }
    }
  }
  __syncthreads();
}

// t = W x = L^-1 x: 16 threads per row, partial sums through the reduction scratch
static __device__ __noinline__ void fpb_trmv(const double* W, int ld, int k, const double* x, double* t, double* sm) {
  __builtin_assume(__isShared(sm));
  __builtin_assume(__isShared(x));
  __builtin_assume(__isShared(t));
  double* red = sm + FPB_OFF_RED;
  const int tid = threadIdx.x, rr = tid >> 4, sl = tid & 15;
  for (int r0 = 0; r0 < k; r0 += FPB_THREADS / 16) {
    const int i = r0 + rr;
    double acc0 = 0.0, acc1 = 0.0;
    if (i < k) {
      const double* w = W + (long long)i * ld;
      int a = sl;
      for (; a + 16 <= i; a += 32) {
        acc0 = fma(w[a], x[a], acc0);
        acc1 = fma(w[a + 16], x[a + 16], acc1);
      }
      if (a <= i) acc0 = fma(w[a], x[a], acc0);
    }
    __syncthreads();
    red[tid] = acc0 + acc1;
    __syncthreads();
    if (tid < FPB_THREADS / 16 && r0 + tid < k) {
      double z = 0.0;
#pragma unroll
      for (int q = 0; q < 16; ++q) z += red[tid * 16 + q];
      t[r0 + tid] = z;
    }
  }
  __syncthreads();
}

// y = W^T t = L^-T t: thread per entry, coalesced over the entries, eight loads in flight
static __device__ __noinline__ void fpb_trmv_t(const double* W, int ld, int k, const double* t, double* y) {
  __builtin_assume(__isShared(t));
  __builtin_assume(__isShared(y));
  for (int a = threadIdx.x; a < k; a += FPB_THREADS) {
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    const double* w = W + a;
    int i = a;
    for (; i + 8 <= k; i += 8) {
      double v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = w[(long long)(i + q) * ld];
#pragma unroll
      for (int q = 0; q < 8; ++q) s[q & 3] = fma(v[q], t[i + q], s[q & 3]);
    }
    for (; i < k; ++i) s[0] = fma(w[(long long)i * ld], t[i], s[0]);
    y[a] = (s[0] + s[1]) + (s[2] + s[3]);
  }
  __syncthreads();
}

}  // namespace dbn
