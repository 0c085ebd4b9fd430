// Stand-alone check and timing of the tiled large-k routines (csrc/dbn_fp_big.cuh): every block factorises its own copy of
// an SPD matrix, inverts the factor and accumulates I - A^-1, as the full-parents kernel does per segment.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I dynamicbayesiannetworks_b200/csrc scripts/fp_big_check.cu -o build/fp_big_check
//   build/fp_big_check [k=500] [blocks=148] [repeats=4]
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "dbn_fp_big.cuh"

using namespace dbn;

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));      \
      exit(1);                                                                                 \
    }                                                                                          \
  } while (0)

__global__ void __launch_bounds__(FPB_THREADS, 1)
    check_kernel(const double* __restrict__ A0, double* ws, int k, int ld, int repeats, double* logdet, double* xin, double* tout, double* yout,
                 long long* cycles, double* s0out, double* s1out, long long* solve_cycles) {
  extern __shared__ __align__(16) double sm[];
  double* A = ws + (long long)blockIdx.x * 3 * ld * k;
  double* W = A + (long long)ld * k;
  double* M = W + (long long)ld * k;
  double* vec = sm + FPB_SMEM_DOUBLES;   // x | t | y | idiag | s0 | s1
  double* x = vec, *t = vec + ld, *y = vec + 2 * ld, *idiag = vec + 3 * ld, *s0 = vec + 4 * ld, *s1 = vec + 5 * ld;
  const int tid = threadIdx.x;
  for (int i = tid; i < k; i += FPB_THREADS) x[i] = xin[i];
  for (long long i = tid; i < (long long)k * ld; i += FPB_THREADS) M[i] = 0.0;
  long long c_f = 0, c_s = 0, c_v = 0, c_v2 = 0;
  double ldet = 0.0;
  for (int r = 0; r < repeats; ++r) {
    for (long long i = tid; i < (long long)k * ld; i += FPB_THREADS) A[i] = A0[i];
    __syncthreads();
    long long t0 = clock64();
    ldet = fpb_factor(A, W, ld, k, idiag, sm);
    long long t1 = clock64();
    fpb_syrk(W, M, ld, k, sm, 1.0);
    long long t2 = clock64();
    fpb_trmv(W, ld, k, x, t, sm);
    fpb_trmv_t(W, ld, k, t, y);
    long long t3 = clock64();
    // the blocked solves against the products with the inverse factor: s0 = L^-T L^-1 x (= y), s1 = L^-T L^-1 (2 x)
    for (int i = tid; i < k; i += FPB_THREADS) {
      s0[i] = x[i];
      s1[i] = 2.0 * x[i];
    }
    __syncthreads();
    fpb_solve_fwd2(A, W, ld, k, s0, s1, sm);
    fpb_solve_bwd2(A, W, ld, k, s0, s1, sm);
    long long t4 = clock64();
    c_v2 += t4 - t3;
    c_f += t1 - t0;
    c_s += t2 - t1;
    c_v += t3 - t2;
  }
  if (tid == 0) {
    logdet[blockIdx.x] = ldet;
    cycles[blockIdx.x * 3 + 0] = c_f / repeats;
    cycles[blockIdx.x * 3 + 1] = c_s / repeats;
    cycles[blockIdx.x * 3 + 2] = c_v / repeats;
    solve_cycles[blockIdx.x] = c_v2 / repeats;
  }
  if (blockIdx.x == gridDim.x - 1)
    for (int i = tid; i < k; i += FPB_THREADS) {
      tout[i] = t[i];
      yout[i] = y[i];
      s0out[i] = s0[i];
      s1out[i] = s1[i];
    }
}

int main(int argc, char** argv) {
  const int k = argc > 1 ? atoi(argv[1]) : 500;
  const int blocks = argc > 2 ? atoi(argv[2]) : 148;
  const int repeats = argc > 3 ? atoi(argv[3]) : 4;
  const int ld = fp_ld_big(k);
  // SPD test matrix: A = I + lam B^T B / m
  const int m = k + 50;
  std::vector<double> B((size_t)m * k), A((size_t)k * ld, 0.0), x(k);
  srand(12345);
  for (auto& v : B) v = (rand() / (double)RAND_MAX) - 0.5;
  for (auto& v : x) v = (rand() / (double)RAND_MAX) - 0.5;
  for (int a = 0; a < k; ++a)
    for (int b = 0; b <= a; ++b) {
      double s = 0.0;
      for (int i = 0; i < m; ++i) s += B[(size_t)i * k + a] * B[(size_t)i * k + b];
      s = 3.0 * s / m * 12.0 + (a == b ? 1.0 : 0.0);
      A[(size_t)a * ld + b] = s;
      A[(size_t)b * ld + a] = s;
    }
  // host reference: L (lower), W = L^-1, Ainv = W^T W
  std::vector<double> L((size_t)k * k, 0.0), Wh((size_t)k * k, 0.0);
  double ldet_ref = 0.0;
  for (int j = 0; j < k; ++j) {
    double d = A[(size_t)j * ld + j];
    for (int q = 0; q < j; ++q) d -= L[(size_t)j * k + q] * L[(size_t)j * k + q];
    const double l = sqrt(d);
    ldet_ref += 2.0 * log(l);
    L[(size_t)j * k + j] = l;
    for (int i = j + 1; i < k; ++i) {
      double s = A[(size_t)i * ld + j];
      for (int q = 0; q < j; ++q) s -= L[(size_t)i * k + q] * L[(size_t)j * k + q];
      L[(size_t)i * k + j] = s / l;
    }
  }
  for (int c = 0; c < k; ++c)
    for (int i = c; i < k; ++i) {
      double s = (i == c) ? 1.0 : 0.0;
      for (int j = c; j < i; ++j) s -= L[(size_t)i * k + j] * Wh[(size_t)j * k + c];
      Wh[(size_t)i * k + c] = s / L[(size_t)i * k + i];
    }
  double *dA0, *dws, *dld, *dx, *dt, *dy;
  long long* dcyc, *dcyc2;
  double *ds0, *ds1;
  CK(cudaMalloc(&dA0, A.size() * 8));
  CK(cudaMalloc(&dws, (size_t)blocks * 3 * ld * k * 8));
  CK(cudaMemset(dws, 0xFF, (size_t)blocks * 3 * ld * k * 8));   // NaN garbage: nothing may depend on untouched workspace
  CK(cudaMalloc(&dld, blocks * 8));
  CK(cudaMalloc(&dx, k * 8));
  CK(cudaMalloc(&dt, k * 8));
  CK(cudaMalloc(&dy, k * 8));
  CK(cudaMalloc(&dcyc, blocks * 3 * 8));
  CK(cudaMalloc(&dcyc2, blocks * 8));
  CK(cudaMalloc(&ds0, k * 8));
  CK(cudaMalloc(&ds1, k * 8));
  CK(cudaMemcpy(dA0, A.data(), A.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dx, x.data(), k * 8, cudaMemcpyHostToDevice));
  const size_t smem = (size_t)(FPB_SMEM_DOUBLES + 6 * ld) * 8;
  CK(cudaFuncSetAttribute(check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  check_kernel<<<blocks, FPB_THREADS, smem>>>(dA0, dws, k, ld, 1, dld, dx, dt, dy, dcyc, ds0, ds1, dcyc2);
  CK(cudaDeviceSynchronize());
  // ---- correctness (last block, one repeat)
  std::vector<double> hA((size_t)k * ld), hW((size_t)k * ld), hM((size_t)k * ld), ht(k), hy(k), hld(blocks);
  const size_t off = (size_t)(blocks - 1) * 3 * ld * k;
  CK(cudaMemcpy(hA.data(), dws + off, hA.size() * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hW.data(), dws + off + (size_t)ld * k, hW.size() * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hM.data(), dws + off + 2 * (size_t)ld * k, hM.size() * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ht.data(), dt, k * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hy.data(), dy, k * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hld.data(), dld, blocks * 8, cudaMemcpyDeviceToHost));
  double eU = 0, eW = 0, eM = 0, et = 0, ey = 0, eL = 0;
  for (int j = 0; j < k; ++j)
    for (int i = j; i < k; ++i) {
      eU = fmax(eU, fabs(hA[(size_t)j * ld + i] - L[(size_t)i * k + j]));
      eW = fmax(eW, fabs(hW[(size_t)i * ld + j] - Wh[(size_t)i * k + j]));
    }
  for (int a = 0; a < k; ++a)
    for (int b = 0; b < k; ++b) {
      double s = 0.0;
      for (int i = (a > b ? a : b); i < k; ++i) s += Wh[(size_t)i * k + a] * Wh[(size_t)i * k + b];
      eM = fmax(eM, fabs(hM[(size_t)a * ld + b] - ((a == b ? 1.0 : 0.0) - s)));
    }
  for (int i = 0; i < k; ++i) {
    double s = 0.0;
    for (int a = 0; a <= i; ++a) s += Wh[(size_t)i * k + a] * x[a];
    et = fmax(et, fabs(ht[i] - s));
  }
  {
    std::vector<double> tr(k);
    for (int i = 0; i < k; ++i) {
      double s = 0.0;
      for (int a = 0; a <= i; ++a) s += Wh[(size_t)i * k + a] * x[a];
      tr[i] = s;
    }
    for (int a = 0; a < k; ++a) {
      double s = 0.0;
      for (int i = a; i < k; ++i) s += Wh[(size_t)i * k + a] * tr[i];
      ey = fmax(ey, fabs(hy[a] - s));
    }
  }
  double es = 0;
  {
    std::vector<double> h0(k), h1(k);
    CK(cudaMemcpy(h0.data(), ds0, k * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h1.data(), ds1, k * 8, cudaMemcpyDeviceToHost));
    for (int i = 0; i < k; ++i) es = fmax(es, fmax(fabs(h0[i] - hy[i]), fabs(h1[i] - 2.0 * hy[i])));
  }
  for (int b = 0; b < blocks; ++b) eL = fmax(eL, fabs(hld[b] - ldet_ref));
  printf("k=%d ld=%d blocks=%d  max abs err: U %.3e  W %.3e  M %.3e  t %.3e  y %.3e  solves %.3e  logdet %.3e (ref %.6f)\n", k, ld, blocks, eU, eW, eM, et, ey, es, eL,
         ldet_ref);
  const bool ok = eU < 1e-10 && eW < 1e-10 && eM < 1e-10 && et < 1e-10 && ey < 1e-10 && es < 1e-10 && eL < 1e-9;
  // ---- timing
  CK(cudaMemset(dws, 0, (size_t)blocks * 3 * ld * k * 8));
  CK(cudaEventRecord(e0));
  check_kernel<<<blocks, FPB_THREADS, smem>>>(dA0, dws, k, ld, repeats, dld, dx, dt, dy, dcyc, ds0, ds1, dcyc2);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> cyc(blocks * 3);
  CK(cudaMemcpy(cyc.data(), dcyc, blocks * 3 * 8, cudaMemcpyDeviceToHost));
  double cf = 0, cs = 0, cv = 0;
  for (int b = 0; b < blocks; ++b) {
    cf += cyc[b * 3];
    cs += cyc[b * 3 + 1];
    cv += cyc[b * 3 + 2];
  }
  cf /= blocks; cs /= blocks; cv /= blocks;
  const double k3 = (double)k * k * k;
  int dev, clk;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev));
  const double ghz = clk * 1e-6;
  printf("kernel %.3f ms for %d x %d segments: %.2f TFLOP/s executed (k^3 per segment)\n", ms, blocks, repeats, blocks * (double)repeats * k3 * 2 / 2 / (ms * 1e-3) / 1e12);
  printf("cycles per segment and block: factor+inverse %.0f (%.1f us, %.1f %% of the SM's DFMA peak), syrk %.0f (%.1f us, %.1f %%), trmv pair %.0f (%.1f us)\n", cf,
         cf / ghz * 1e-3, 100.0 * (2.0 * k3 / 3.0) / (cf * 128.0), cs, cs / ghz * 1e-3, 100.0 * (k3 / 3.0) / (cs * 128.0), cv, cv / ghz * 1e-3);
  {
    std::vector<long long> c2(blocks);
    CK(cudaMemcpy(c2.data(), dcyc2, blocks * 8, cudaMemcpyDeviceToHost));
    double a = 0;
    for (int b = 0; b < blocks; ++b) a += c2[b];
    printf("two-vector forward + backward solve: %.0f cycles (%.1f us)\n", a / blocks, a / blocks / ghz * 1e-3);
  }
  printf("%s\n", ok ? "OK" : "FAILED");
  return ok ? 0 : 1;
}
