// Launchers of the chain kernels.  Each kernel family is instantiated in its own translation unit (dbn_inst.cu compiled
// with -DDBN_INST_FAMILY / -DDBN_INST_TRACE / -DDBN_INST_BIG), so the library builds in parallel and a change to one
// family recompiles only its objects; dbn_kernels.cu (the C ABI) calls these instead of naming a kernel.
#pragma once
#include <cuda_runtime.h>

namespace dbn {

struct ChainArgs;
struct FpArgs;

// h / nh: chain_fast_kernel<METHOD, KM, TRACE> (dbn_chain_fast.cuh)
cudaError_t launch_fast_plain(int method, int km, unsigned grid, int block, size_t smem, cudaStream_t s, const ChainArgs& a);
cudaError_t launch_fast_trace(int method, int km, unsigned grid, int block, size_t smem, cudaStream_t s, const ChainArgs& a);
// seq / glob / var-glob: chain_coupled_kernel<METHOD, KM, TRACE> (dbn_chain.cuh)
cudaError_t launch_chain_plain(int method, int km, unsigned grid, int block, size_t smem, cudaStream_t s, const ChainArgs& a);
cudaError_t launch_chain_trace(int method, int km, unsigned grid, int block, size_t smem, cudaStream_t s, const ChainArgs& a);
// full parents: fp_glob_kernel<TRACE, FPM, BIG> (dbn_fp.cuh)
cudaError_t launch_fp_plain_small(int fpm, unsigned grid, int block, size_t smem, cudaStream_t s, const FpArgs& a);
cudaError_t launch_fp_plain_big(int fpm, unsigned grid, int block, size_t smem, cudaStream_t s, const FpArgs& a);
cudaError_t launch_fp_trace_small(int fpm, unsigned grid, int block, size_t smem, cudaStream_t s, const FpArgs& a);
cudaError_t launch_fp_trace_big(int fpm, unsigned grid, int block, size_t smem, cudaStream_t s, const FpArgs& a);

}  // namespace dbn
