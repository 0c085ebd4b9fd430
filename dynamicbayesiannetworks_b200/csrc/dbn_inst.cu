// One kernel family per object file: compiled several times by build.py with
//   -DDBN_INST_FAMILY=1 (h / nh fast kernel) | 2 (seq / glob / var-glob) | 3 (full parents)
//   -DDBN_INST_TRACE=0|1      (TRACE template flag: 1 = the replay / trace build the parity fixtures pin)
//   -DDBN_INST_BIG=0|1        (family 3 only: matrices in the global workspace)
#include "dbn_launch.h"

#ifndef DBN_INST_FAMILY
#error "dbn_inst.cu: define DBN_INST_FAMILY"
#endif
#ifndef DBN_INST_TRACE
#define DBN_INST_TRACE 0
#endif
#ifndef DBN_INST_BIG
#define DBN_INST_BIG 0
#endif

#if DBN_INST_FAMILY == 1
#include "dbn_chain_fast.cuh"
#elif DBN_INST_FAMILY == 2
#include "dbn_chain.cuh"
#else
#include "dbn_fp.cuh"
#endif

#include <cstdlib>
namespace dbn {

template <class Fn, class Args>
static cudaError_t launch_(Fn fn, unsigned grid, int block, size_t smem, cudaStream_t s, const Args& a) {
  if (smem > 16 * 1024) {   // static + dynamic shared memory above 48 KB needs the opt-in as well (the var-glob kernel has 20 KB static)
    cudaError_t e = cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  if (const char* c = getenv("DBN_CARVEOUT")) {   // experiment knob: preferred shared-memory carve-out in percent (profiles/r2_small)
    cudaError_t e = cudaFuncSetAttribute((const void*)fn, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(c));
    if (e != cudaSuccess) return e;
  }
  fn<<<grid, block, smem, s>>>(a);
  return cudaGetLastError();
}

constexpr bool TR = DBN_INST_TRACE != 0;

#if DBN_INST_FAMILY == 1
#if DBN_INST_TRACE
cudaError_t launch_fast_trace
#else
cudaError_t launch_fast_plain
#endif
    (int method, int km, unsigned grid, int block, size_t smem, cudaStream_t s, const ChainArgs& a) {
  typedef void (*fn_t)(const ChainArgs);
  fn_t fn;
  if (method == DBN_H_DBN) fn = km == 4 ? chain_fast_kernel<DBN_H_DBN, 4, TR> : chain_fast_kernel<DBN_H_DBN, 5, TR>;
  else fn = km == 4 ? chain_fast_kernel<DBN_NH_DBN, 4, TR> : chain_fast_kernel<DBN_NH_DBN, 5, TR>;
  return launch_(fn, grid, block, smem, s, a);
}
#elif DBN_INST_FAMILY == 2
#if DBN_INST_TRACE
cudaError_t launch_chain_trace
#else
cudaError_t launch_chain_plain
#endif
    (int method, int km, unsigned grid, int block, size_t smem, cudaStream_t s, const ChainArgs& a) {
  typedef void (*fn_t)(const ChainArgs);
  fn_t fn;
  if (method == DBN_SEQ_DBN) fn = km == 4 ? chain_kernel<DBN_SEQ_DBN, 4, TR> : chain_kernel<DBN_SEQ_DBN, 5, TR>;
  else if (method == DBN_VV_GLOB_DBN) fn = km == 4 ? chain_kernel<DBN_VV_GLOB_DBN, 4, TR> : chain_kernel<DBN_VV_GLOB_DBN, 5, TR>;
  else fn = km == 4 ? chain_kernel<DBN_GLOB_DBN, 4, TR> : chain_kernel<DBN_GLOB_DBN, 5, TR>;
  return launch_(fn, grid, block, smem, s, a);
}
#else
constexpr bool BG = DBN_INST_BIG != 0;
#if DBN_INST_TRACE && DBN_INST_BIG
cudaError_t launch_fp_trace_big
#elif DBN_INST_TRACE
cudaError_t launch_fp_trace_small
#elif DBN_INST_BIG
cudaError_t launch_fp_plain_big
#else
cudaError_t launch_fp_plain_small
#endif
    (int fpm, unsigned grid, int block, size_t smem, cudaStream_t s, const FpArgs& a) {
  typedef void (*fn_t)(const FpArgs);
  const fn_t fn = fpm == FPM_GLOB ? fp_glob_kernel<TR, FPM_GLOB, BG>
                  : fpm == FPM_VV ? fp_glob_kernel<TR, FPM_VV, BG>
                  : fpm == FPM_SEQ ? fp_glob_kernel<TR, FPM_SEQ, BG>
                                   : fp_glob_kernel<TR, FPM_PLAIN, BG>;
  return launch_(fn, grid, block, smem, s, a);
}
#endif

}  // namespace dbn
