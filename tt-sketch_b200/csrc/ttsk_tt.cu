// Streaming sketch of ONE TensorTrain summand with TT DRMs in a single C call: the DRM
// contractions (tensor_train_drm.py:71-85, reference), Omega (tensor_train_sketch.py:8-13) and Psi
// (tensor_train_sketch.py:16-35) as a fixed sequence of strided FP64 GEMMs on the caller's stream,
// accumulated into the packed sketch.  The Python mirror issues the same GEMMs one ctypes call at a
// time (about 4 ms of host overhead per summand); a TensorSum of a hundred TT summands (BASELINE
// config 5) is launch-bound there, not here.
#include <algorithm>
#include <vector>

#include "ttsk_common.cuh"

using namespace ttsk;

namespace {

struct Mat {  // (rows, cols) view: element (i, j) at p[i * rs + j * cs]
    double* p;
    int64_t rows, cols, rs, cs;
};

int mm(ttsk_ctx* ctx, const Mat& A, const Mat& B, double beta, const Mat& C, cudaStream_t st) {
    return gemm_launch(ctx, A.rows, B.cols, A.cols, 1.0, A.p, A.rs, A.cs, B.p, B.rs, B.cs, beta, C.p, C.rs, C.cs, 1, 0, 0,
                       0, st);
}

}  // namespace

extern "C" int ttsk_tt_sketch(ttsk_ctx* ctx, int d, const int64_t* h_shape, const int32_t* h_tt_rank,
                              const double* const* h_core_ptrs, const ttsk_drm* left, const ttsk_drm* right,
                              double* d_out, void* stream) {
    TTSK_ARG(ctx != nullptr && h_shape && h_tt_rank && h_core_ptrs && left && right && d_out, "NULL argument");
    TTSK_ARG(d >= 2 && d <= TTSK_MAX_ORDER, "tensor order must be in [2, 16]");
    TTSK_ARG(left->kind == TTSK_DRM_TT && right->kind == TTSK_DRM_TT, "ttsk_tt_sketch needs TT DRMs on both sides");
    TTSK_ARG(!left->right && right->right, "left/right DRM orientation mismatch");
    TTSK_ARG(h_tt_rank[0] == 1 && h_tt_rank[d] == 1, "boundary TT ranks must be 1");
    cudaStream_t st = (cudaStream_t)stream;
    TTSK_CUDA(cudaSetDevice(ctx->device));
    const int32_t* r = h_tt_rank;  // core k is (r[k], n_k, r[k+1])
    int rL[TTSK_MAX_ORDER], rR[TTSK_MAX_ORDER];
    for (int mu = 0; mu < d - 1; mu++) {
        rL[mu] = left->rank_max[mu] - left->rank_min[mu];
        rR[mu] = right->rank_max[mu] - right->rank_min[mu];
        TTSK_ARG(rL[mu] >= 1 && rR[mu] >= 1, "empty rank slice");
        TTSK_ARG(left->rank_max[mu] <= left->core_r1[mu] && right->rank_max[mu] <= right->core_r1[d - 2 - mu],
                 "rank slice exceeds TT-DRM core rank");
    }
    // packed layout (same as ttsk_sparse_sketch)
    int64_t psi_off[TTSK_MAX_ORDER], omega_off[TTSK_MAX_ORDER], off = 0;
    for (int mu = 0; mu < d; mu++) {
        psi_off[mu] = off;
        off += (int64_t)(mu == 0 ? 1 : rL[mu - 1]) * h_shape[mu] * (mu == d - 1 ? 1 : rR[mu]);
    }
    for (int mu = 0; mu < d - 1; mu++) {
        omega_off[mu] = off;
        off += (int64_t)rL[mu] * rR[mu];
    }
    // workspace: the chain matrices of every bond (r_T x core rank), one scratch for the widest intermediate
    int64_t bytes = 0, scratch = 0;
    auto add = [&](int64_t n) { const int64_t at = bytes; bytes += align_up(n * 8, 256); return at; };
    int64_t l_at[TTSK_MAX_ORDER], r_at[TTSK_MAX_ORDER];
    for (int mu = 0; mu < d - 1; mu++) {
        l_at[mu] = add((int64_t)r[mu + 1] * left->core_r1[mu]);
        r_at[mu] = add((int64_t)r[mu + 1] * right->core_r1[d - 2 - mu]);
    }
    for (int k = 0; k < d; k++) {
        const int64_t n = h_shape[k];
        if (k >= 1 && k <= d - 2) scratch = std::max<int64_t>(scratch, (int64_t)left->core_r0[k] * n * r[k + 1]);
        if (k >= 1 && k <= d - 2) scratch = std::max<int64_t>(scratch, (int64_t)r[k] * right->core_r0[d - 1 - k] * n);
        if (k >= 1 && k <= d - 2) scratch = std::max<int64_t>(scratch, (int64_t)rL[k - 1] * n * r[k + 1]);
    }
    const int64_t scratch_at = add(std::max<int64_t>(scratch, 1));
    TTSK_TRY(ctx->ws_reserve(bytes + 4096));
    ctx->ws_reset();
    char* ws = (char*)ctx->ws_alloc(bytes);
    if (!ws) { set_error("workspace carve failed (TT sketch)"); return TTSK_E_NOMEM; }
    double* tmp = (double*)(ws + scratch_at);

    // ---- left DRM: lr_mu (r[mu+1], core_r1[mu]) = DRM contracted with the first mu+1 cores
    for (int mu = 0; mu < d - 1; mu++) {
        const int64_t n = h_shape[mu];
        double* c = const_cast<double*>(h_core_ptrs[mu]);
        double* g = const_cast<double*>(left->d_cores[mu]);
        const int64_t rT0 = r[mu], rT1 = r[mu + 1], rD0 = left->core_r0[mu], rD1 = left->core_r1[mu];
        Mat lr{(double*)(ws + l_at[mu]), rT1, rD1, rD1, 1};
        if (mu == 0) {  // c (n, rT1)^T @ g (n, rD1)
            TTSK_TRY(mm(ctx, Mat{c, rT1, n, 1, rT1}, Mat{g, n, rD1, rD1, 1}, 0.0, lr, st));
        } else {
            Mat prev{(double*)(ws + l_at[mu - 1]), rT0, rD0, rD0, 1};
            // w (rD0, n*rT1) = prev^T @ c (rT0, n*rT1);   lr = w.reshape(rD0*n, rT1)^T @ g.reshape(rD0*n, rD1)
            TTSK_TRY(mm(ctx, Mat{prev.p, rD0, rT0, 1, rD0}, Mat{c, rT0, n * rT1, n * rT1, 1}, 0.0, Mat{tmp, rD0, n * rT1, n * rT1, 1}, st));
            TTSK_TRY(mm(ctx, Mat{tmp, rT1, rD0 * n, 1, rT1}, Mat{g, rD0 * n, rD1, rD1, 1}, 0.0, lr, st));
        }
    }
    // ---- right DRM (operates on the mode-reversed tensor): rr of bond mu (r[mu+1], core_r1[d-2-mu])
    for (int k = 0; k < d - 1; k++) {
        const int m = d - 1 - k, mu = d - 2 - k;  // tensor mode of this level, bond it belongs to
        const int64_t n = h_shape[m];
        double* c = const_cast<double*>(h_core_ptrs[m]);   // (ra, n, rb)
        double* g = const_cast<double*>(right->d_cores[k]);  // (rD0, n, rD1)
        const int64_t ra = r[m], rb = r[m + 1], rD0 = right->core_r0[k], rD1 = right->core_r1[k];
        Mat rr{(double*)(ws + r_at[mu]), ra, rD1, rD1, 1};
        if (k == 0) {  // c (ra, n) @ g (n, rD1)
            TTSK_TRY(mm(ctx, Mat{c, ra, n, n, 1}, Mat{g, n, rD1, rD1, 1}, 0.0, rr, st));
        } else {
            Mat prev{(double*)(ws + r_at[mu + 1]), rb, rD0, rD0, 1};
            // W3[a] (rD0, n) = prev^T (rD0, rb) @ c[a]^T (rb, n), batched over a;   rr = W3.reshape(ra, rD0*n) @ g.reshape(rD0*n, rD1)
            TTSK_TRY(gemm_launch(ctx, rD0, n, rb, 1.0, prev.p, 1, rD0, c, 1, rb, 0.0, tmp, n, 1, ra, 0, n * rb, rD0 * n, st));
            TTSK_TRY(mm(ctx, Mat{tmp, ra, rD0 * n, rD0 * n, 1}, Mat{g, rD0 * n, rD1, rD1, 1}, 0.0, rr, st));
        }
    }
    auto Lm = [&](int mu) { return Mat{(double*)(ws + l_at[mu]) + left->rank_min[mu], r[mu + 1], rL[mu], left->core_r1[mu], 1}; };
    auto Rm = [&](int mu) { return Mat{(double*)(ws + r_at[mu]) + right->rank_min[mu], r[mu + 1], rR[mu], right->core_r1[d - 2 - mu], 1}; };
    // ---- Omega_mu += L_mu^T R_mu
    for (int mu = 0; mu < d - 1; mu++) {
        const Mat L = Lm(mu), R = Rm(mu);
        TTSK_TRY(mm(ctx, Mat{L.p, L.cols, L.rows, L.cs, L.rs}, R, 1.0, Mat{d_out + omega_off[mu], rL[mu], rR[mu], rR[mu], 1}, st));
    }
    // ---- Psi_mu
    for (int mu = 0; mu < d; mu++) {
        const int64_t n = h_shape[mu], r0 = r[mu], r1 = r[mu + 1];
        double* c = const_cast<double*>(h_core_ptrs[mu]);
        double* out = d_out + psi_off[mu];
        if (mu == 0) {  // (n, r1) @ R_0
            const Mat R = Rm(0);
            TTSK_TRY(mm(ctx, Mat{c, n, r1, r1, 1}, R, 1.0, Mat{out, n, rR[0], rR[0], 1}, st));
        } else if (mu == d - 1) {  // L^T @ (r0, n)
            const Mat L = Lm(d - 2);
            TTSK_TRY(mm(ctx, Mat{L.p, L.cols, L.rows, L.cs, L.rs}, Mat{c, r0, n, n, 1}, 1.0, Mat{out, rL[d - 2], n, n, 1}, st));
        } else {  // t1 (rL, n*r1) = L^T @ c (r0, n*r1);   out (rL*n, rR) += t1.reshape(rL*n, r1) @ R
            const Mat L = Lm(mu - 1), R = Rm(mu);
            TTSK_TRY(mm(ctx, Mat{L.p, L.cols, L.rows, L.cs, L.rs}, Mat{c, r0, n * r1, n * r1, 1}, 0.0,
                        Mat{tmp, rL[mu - 1], n * r1, n * r1, 1}, st));
            TTSK_TRY(mm(ctx, Mat{tmp, rL[mu - 1] * n, r1, r1, 1}, R, 1.0, Mat{out, rL[mu - 1] * n, rR[mu], rR[mu], 1}, st));
        }
    }
    return TTSK_OK;
}
