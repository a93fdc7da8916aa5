// Instantiates the mode-pass kernels with a middle bond (HAS_X = true), including the segment-GEMM form.
#include "ttsk_sparse_pass.cuh"

namespace ttsk {

int launch_pass_with_x(ttsk_ctx* ctx, PassParams& P, cudaStream_t st) { return launch_pass_x<true>(ctx, P, st); }

}  // namespace ttsk
