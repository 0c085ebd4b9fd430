// Development aid for kernel A/B experiments (build.py --slim): a library with only the h / nh kernel family; the
// launchers of the other families fail loudly.  Never part of the shipped libdbn_b200.so.
#include "dbn_launch.h"

namespace dbn {
cudaError_t launch_chain_plain(int, int, unsigned, int, size_t, cudaStream_t, const ChainArgs&) { return cudaErrorNotSupported; }
cudaError_t launch_chain_trace(int, int, unsigned, int, size_t, cudaStream_t, const ChainArgs&) { return cudaErrorNotSupported; }
cudaError_t launch_fp_plain_small(int, unsigned, int, size_t, cudaStream_t, const FpArgs&) { return cudaErrorNotSupported; }
cudaError_t launch_fp_plain_big(int, unsigned, int, size_t, cudaStream_t, const FpArgs&) { return cudaErrorNotSupported; }
cudaError_t launch_fp_trace_small(int, unsigned, int, size_t, cudaStream_t, const FpArgs&) { return cudaErrorNotSupported; }
cudaError_t launch_fp_trace_big(int, unsigned, int, size_t, cudaStream_t, const FpArgs&) { return cudaErrorNotSupported; }
}  // namespace dbn
