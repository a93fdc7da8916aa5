// Instantiates the mode-pass kernels without a middle bond (HAS_X = false) and the unbucketed last-mode form.
#include "ttsk_sparse_pass.cuh"

namespace ttsk {

int launch_pass_without_x(ttsk_ctx* ctx, PassParams& P, cudaStream_t st) { return launch_pass_x<false>(ctx, P, st); }

int launch_last_mode_unbucketed(ttsk_ctx* ctx, PassParams& P, cudaStream_t st, bool* used) {
    const int mi = (P.rA + 7) / 8;
    *used = false;
    if (mi <= 1) return try_launch_sg_flat<1>(ctx, P, st, used);
    if (mi <= 3) return try_launch_sg_flat<3>(ctx, P, st, used);
    if (mi <= 5) return try_launch_sg_flat<5>(ctx, P, st, used);
    if (mi <= 8) return try_launch_sg_flat<8>(ctx, P, st, used);
    return TTSK_OK;
}

}  // namespace ttsk
