// FP64 tensor-core (mma.sync.m8n8k4.f64) against DFMA throughput on one GPU: decides whether the level-3 steps of the
// large-k full-parents kernel should use DMMA.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/dmma_peak.cu -o build/dmma_peak
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NCH>
__global__ void dmma_kernel(double* out, int iters, double seed) {
  double c[NCH][2];
  for (int i = 0; i < NCH; ++i) c[i][0] = c[i][1] = 0.0;
  double a = seed + threadIdx.x * 1e-9, b = 1.0 - seed;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0;
  for (int i = 0; i < NCH; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dfma_kernel(double* out, int iters, double seed) {
  double a[8];
  for (int i = 0; i < 8; ++i) a[i] = seed + i;
  const double m = 0.999999, c = 1e-9 * (threadIdx.x + 1);
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, c);
  double s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F>
float timeit(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  double* out; cudaMalloc(&out, 148 * 4 * 1024 * 8);
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    const int thr = warps * 32, blocks = 148;
    float ms = timeit([&] { dmma_kernel<8><<<blocks, thr>>>(out, iters, 0.5); });
    printf("DMMA m8n8k4, %2d warps/SM x 8 chains: %.2f TFLOP/s\n", warps, (double)blocks * warps * iters * 8 * 512 / (ms * 1e-3) / 1e12);
    ms = timeit([&] { dmma_kernel<2><<<blocks, thr>>>(out, iters, 0.5); });
    printf("DMMA m8n8k4, %2d warps/SM x 2 chains: %.2f TFLOP/s\n", warps, (double)blocks * warps * iters * 2 * 512 / (ms * 1e-3) / 1e12);
    ms = timeit([&] { dfma_kernel<<<blocks, thr>>>(out, iters, 0.5); });
    printf("DFMA,         %2d warps/SM x 8 chains: %.2f TFLOP/s\n", warps, (double)blocks * thr * iters * 8 * 2 / (ms * 1e-3) / 1e12);
  }
  cudaError_t e = cudaGetLastError();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
