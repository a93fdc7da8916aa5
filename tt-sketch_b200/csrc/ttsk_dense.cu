// Streaming sketch of ONE DenseTensor with TensorTrainDRMs in a single C call.
//
// Replaces (reference, for DenseTensor input and method = streaming):
//   TensorTrainDRM.sketch_dense          tt_sketch/drm/tensor_train_drm.py:109-122
//   sketch_omega_dense / sketch_psi_dense tt_sketch/sketching_methods/dense_sketch.py:7-52
//
// The reference streams X once per Omega and once per Psi (about 2 d - 1 passes) against materialised DRM
// unfoldings.  Here X is read from HBM ONCE:
//   * the LEFT DRM is a chain, so its partial contractions are swept left to right,
//         XL_0 = G_0^T X_(0),    XL_mu = G_mu^T XL_{mu-1}   (g1_mu x prod_{m > mu} n_m),
//     each a 20x smaller operand than the one before; every Omega_mu / Psi_mu is then a small product of an XL
//     with a right-DRM unfolding:  Omega_mu = XL_mu Rq_mu^T,  Psi_mu = XL_{mu-1} as ((a, i_mu) x rest) Rq_mu^T;
//   * the RIGHT unfoldings Rq_mu are used as flat arrays against the C-order unfolding of X exactly like the
//     reference does (its reversed-mode column order), so they cannot be swept; they are
//     materialised once per call like the reference (`sketch_dense`), the largest being (rR x prod_{m >= 1} n_m);
//   * the only two products that touch X -- XL_0 and Psi_0 = X_(0) Rq_0^T -- are ONE fused kernel
//     (dense_first_pass_kernel): tiles of X ([n_0 rows] x 72 columns) and of Rq_0 are staged by TMA
//     (cp.async.bulk.tensor.2d, mbarrier complete_tx, multi-stage ring) and both products run on FP64 tensor-core
//     MMAs from the staged tile.  Everything after it works on operands that are at most half the size of X and
//     L2-resident.
// All launches are issued back to back on the caller's stream (no host round trip in between).
#include <algorithm>
#include <cuda.h>
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

// ------------------------------------------------------------------ the fused first pass
constexpr int kFpCols = 72;      // columns of a tile: 9 MMA column tiles, pitch == 8 (mod 16) doubles -> conflict-free B fragments
constexpr int kFpWarps = 9;      // one MMA column tile of XL per warp
constexpr int kFpThreads = 32 * kFpWarps;
constexpr int kFpStages = 6;
constexpr int kFpMaxRows = 32;   // n_0 (rows of a tile) <= 32
constexpr int kFpMaxG = 32;      // g1_0 (left rank of the first bond) <= 32
constexpr int kFpMaxR = 32;      // right rank of the first bond <= 32

__device__ __forceinline__ void fp_dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ unsigned fp_smem(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

struct FirstPassArgs {
    long long n0, F;          // X is (n0, F) row-major
    int g1, rR, rRp;          // rRp: row pitch of Rq_0 in doubles (even: TMA rows are multiples of 16 bytes)
    const double* G0;         // (n0, g1) row-major: first core of the left DRM
    double* XL0;              // (g1, F) row-major
    double* Y0;               // (n0, rR) row-major, zero on entry: X Rq_0^T (accumulated with atomics)
    long long n_tiles;
};

// One CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... of 72 columns.  Thread 0 issues the TMA loads of
// both operand tiles of a stage (X: n0 x 72 box, Rq_0 viewed as (F, rR): 72 x rR box) and the stage's mbarrier
// completes on the byte count; stages are recycled after a CTA barrier.
//   XL0[a, col]  = sum_row G0[row, a] X[row, col]        M = g1 (MI tiles), N = 72 (warp w: tile w), K = n0
//   Y0[row, b]  += sum_col X[row, col] Rq[col, b]        M = n0, N = rR, K = 72: k-steps split over the warps
template <int MI /* ceil(g1 / 8) */, int MR /* ceil(n0 / 8) */, int NJ /* ceil(rR / 8) */>
__global__ void __launch_bounds__(kFpThreads, 1)
    dense_first_pass_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_r,
                            const FirstPassArgs A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int n0 = (int)A.n0, n0p = 8 * MR;
    const int x_bytes = n0 * kFpCols * 8;                 // what TMA writes per stage for X
    const int x_stage = ((n0p * kFpCols * 8) + 127) & ~127;  // rows up to the MMA padding are zeroed once and never written
    const int r_bytes = kFpCols * A.rRp * 8;
    const int r_stage = ((kFpCols * 8 * NJ * 8) + 127) & ~127;
    unsigned char* xs = smem_raw;
    unsigned char* rs = xs + (size_t)kFpStages * x_stage;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(rs + (size_t)kFpStages * r_stage);
    double* g0s = reinterpret_cast<double*>(bars + kFpStages);  // [4 * ceil(n0 / 4)][8 * MI + 1]  (row k, column a), zero padded
    const int GP = 8 * MI + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int ksteps_x = (n0 + 3) >> 2;

    for (int i = tid; i < (kFpStages * x_stage + kFpStages * r_stage) / 8; i += kFpThreads) reinterpret_cast<double*>(smem_raw)[i] = 0.0;
    for (int i = tid; i < 4 * ksteps_x * GP; i += kFpThreads) {
        const int k = i / GP, a = i - k * GP;
        g0s[i] = (k < n0 && a < A.g1) ? A.G0[(long long)k * A.g1 + a] : 0.0;
    }
    if (tid == 0) {
        for (int s = 0; s < kFpStages; s++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fp_smem(&bars[s])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // make the generic-proxy zero fill visible to the async proxy before TMA writes into the same buffers
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    auto issue = [&](long long t, int s) {  // thread 0
        const unsigned bar = fp_smem(&bars[s]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(x_bytes + r_bytes) : "memory");
        const int c0 = (int)(t * kFpCols);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(fp_smem(xs + (size_t)s * x_stage)), "l"(&tm_x), "r"(c0), "r"(0), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(fp_smem(rs + (size_t)s * r_stage)), "l"(&tm_r), "r"(0), "r"(c0), "r"(bar) : "memory");
    };
    const long long t0 = blockIdx.x, dt = gridDim.x;
    if (tid == 0)
        for (int s = 0; s < kFpStages; s++)
            if (t0 + s * dt < A.n_tiles) issue(t0 + s * dt, s);

    // A fragments of G0^T stay in registers for the whole kernel
    double ga[MI][(kFpMaxRows + 3) / 4];
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
        for (int s = 0; s < (kFpMaxRows + 3) / 4; s++) ga[i][s] = (s < ksteps_x) ? g0s[(4 * s + q) * GP + 8 * i + g] : 0.0;

    double yacc[MR][NJ][2];
#pragma unroll
    for (int i = 0; i < MR; i++)
#pragma unroll
        for (int j = 0; j < NJ; j++) yacc[i][j][0] = yacc[i][j][1] = 0.0;

    int s = 0;
    unsigned ph = 0;
    const int rRp = A.rRp;  // row pitch of the staged Rq tile (doubles): dense TMA box (72 rows x rRp)
    for (long long t = t0; t < A.n_tiles; t += dt) {
        {   // wait for the stage
            const unsigned bar = fp_smem(&bars[s]);
            unsigned done;
            do {
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
                             : "=r"(done) : "r"(bar), "r"(ph), "r"(100000u) : "memory");
            } while (!done);
        }
        const double* X = reinterpret_cast<const double*>(xs + (size_t)s * x_stage);   // [n0p][72]
        const double* R = reinterpret_cast<const double*>(rs + (size_t)s * r_stage);   // [72][rR]
        // ---- XL0 tile: this warp's 8 columns
        {
            double acc[MI][2];
#pragma unroll
            for (int i = 0; i < MI; i++) acc[i][0] = acc[i][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < (kFpMaxRows + 3) / 4; ks++) {
                if (ks < ksteps_x) {
                    const double b = X[(4 * ks + q) * kFpCols + 8 * warp + g];  // rows >= n0 of the stage are zero
#pragma unroll
                    for (int i = 0; i < MI; i++) fp_dmma(acc[i][0], acc[i][1], ga[i][ks], b);
                }
            }
            const long long col = t * kFpCols + 8 * warp + 2 * q;
#pragma unroll
            for (int i = 0; i < MI; i++) {
                const int a = 8 * i + g;
                if (a < A.g1) {
                    double* dst = A.XL0 + (long long)a * A.F + col;
                    if (col + 1 < A.F && ((reinterpret_cast<unsigned long long>(dst) & 15ull) == 0)) {
                        *reinterpret_cast<double2*>(dst) = make_double2(acc[i][0], acc[i][1]);
                    } else {
                        if (col < A.F) dst[0] = acc[i][0];
                        if (col + 1 < A.F) dst[1] = acc[i][1];
                    }
                }
            }
        }
        // ---- Y0 += X tile (n0 x 72) @ R tile (72 x rR): k-steps 2 * warp, 2 * warp + 1 of the 18
#pragma unroll
        for (int kk = 0; kk < 2; kk++) {
            const int ks = 2 * warp + kk;
            double a[MR], b[NJ];
#pragma unroll
            for (int i = 0; i < MR; i++) a[i] = X[(8 * i + g) * kFpCols + 4 * ks + q];
#pragma unroll
            for (int j = 0; j < NJ; j++) b[j] = (8 * j + g < A.rR) ? R[(4 * ks + q) * rRp + 8 * j + g] : 0.0;  // columns past F are zero-filled by TMA
#pragma unroll
            for (int i = 0; i < MR; i++)
#pragma unroll
                for (int j = 0; j < NJ; j++) fp_dmma(yacc[i][j][0], yacc[i][j][1], a[i], b[j]);
        }
        __syncthreads();  // every warp is done with the stage
        if (tid == 0 && t + kFpStages * dt < A.n_tiles) issue(t + kFpStages * dt, s);
        if (++s == kFpStages) { s = 0; ph ^= 1u; }
    }
    // Y0: the warps' partial sums (each warp covered its own k-steps) are added up in shared memory, so a CTA sends
    // ONE atomic per element of the small (n0 x rR) matrix (all CTAs add into the same few hundred addresses)
    __syncthreads();  // every TMA load has been consumed: the stage buffers are free
    double* ysum = reinterpret_cast<double*>(xs);  // [kFpWarps][8 MR][8 NJ]
    constexpr int YP = 8 * NJ;
#pragma unroll
    for (int i = 0; i < MR; i++)
#pragma unroll
        for (int j = 0; j < NJ; j++) {
            double* dst = ysum + ((size_t)warp * 8 * MR + 8 * i + g) * YP + 8 * j + 2 * q;
            dst[0] = yacc[i][j][0];
            dst[1] = yacc[i][j][1];
        }
    __syncthreads();
    for (int e = tid; e < n0 * A.rR; e += kFpThreads) {
        const int row = e / A.rR, col = e - row * A.rR;
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < kFpWarps; w++) sum += ysum[((size_t)w * 8 * MR + row) * YP + col];
        if (sum != 0.0) atomicAdd(A.Y0 + (long long)row * A.rR + col, sum);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}

// (rows, cols) FP64 matrix with row pitch `pitch` doubles, box (box_rows, box_cols); out-of-bounds elements read as zero
bool make_map(CUtensorMap* m, const double* base, int64_t rows, int64_t cols, int64_t pitch, int box_rows, int box_cols) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch * 8};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int MI, int MR, int NJ>
int first_pass_launch_t(ttsk_ctx* ctx, const CUtensorMap& tx, const CUtensorMap& tr, const FirstPassArgs& A, cudaStream_t st) {
    auto kern = dense_first_pass_kernel<MI, MR, NJ>;
    const int x_stage = ((8 * MR * kFpCols * 8) + 127) & ~127;
    const int r_stage = ((kFpCols * 8 * NJ * 8) + 127) & ~127;
    const int ksteps_x = ((int)A.n0 + 3) >> 2;
    const size_t smem = (size_t)kFpStages * (x_stage + r_stage) + kFpStages * 8 + (size_t)4 * ksteps_x * (8 * MI + 1) * 8 + 128;
    TTSK_ARG(smem <= 227 * 1024, "dense first pass: shared-memory plan exceeds 227 KB");
    TTSK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long grid = std::min<long long>(ctx->sm_count, A.n_tiles);
    kern<<<(unsigned)grid, kFpThreads, smem, st>>>(tx, tr, A);
    TTSK_LAUNCHED(ctx);
    return TTSK_OK;
}

// returns TTSK_OK and sets *used when the fused kernel ran
int first_pass(ttsk_ctx* ctx, const double* X, int64_t n0, int64_t F, const double* G0, int g1, const double* Rq0, int rR,
               int rRp, double* XL0, double* Y0, cudaStream_t st, bool* used) {
    *used = false;
    static const int off = getenv("TTSK_NO_TMA") ? atoi(getenv("TTSK_NO_TMA")) : 0;
    if (off || n0 > kFpMaxRows || g1 > kFpMaxG || rR > kFpMaxR || F < kFpCols) return TTSK_OK;
    // TMA: 16-byte aligned bases and row strides
    if ((F & 1) || (rRp & 1) || (reinterpret_cast<uintptr_t>(X) & 15) || (reinterpret_cast<uintptr_t>(Rq0) & 15)) return TTSK_OK;
    CUtensorMap tx, tr;
    if (!make_map(&tx, X, n0, F, F, (int)n0, kFpCols)) return TTSK_OK;
    if (!make_map(&tr, Rq0, F, rRp, rRp, kFpCols, rRp)) return TTSK_OK;  // the pad column is staged too and never read
    FirstPassArgs A;
    A.n0 = n0; A.F = F; A.g1 = g1; A.rR = rR; A.rRp = rRp; A.G0 = G0; A.XL0 = XL0; A.Y0 = Y0;
    A.n_tiles = (F + kFpCols - 1) / kFpCols;
    const int mi = (g1 + 7) / 8, mr = (int)((n0 + 7) / 8), nj = (rR + 7) / 8;
    const int MIr = mi <= 2 ? 2 : 4, MRr = mr <= 3 ? 3 : 4, NJr = nj <= 2 ? 2 : 4;
    int rc = TTSK_OK;
#define TTSK_FP(a, b, c) if (MIr == a && MRr == b && NJr == c) rc = first_pass_launch_t<a, b, c>(ctx, tx, tr, A, st)
    TTSK_FP(2, 3, 2); TTSK_FP(2, 3, 4); TTSK_FP(4, 3, 2); TTSK_FP(4, 3, 4);
    TTSK_FP(2, 4, 2); TTSK_FP(2, 4, 4); TTSK_FP(4, 4, 2); TTSK_FP(4, 4, 4);
#undef TTSK_FP
    if (rc == TTSK_OK) *used = true;
    return rc;
}

}  // namespace

extern "C" int ttsk_dense_sketch(ttsk_ctx* ctx, int d, const int64_t* h_shape, const double* d_X, const ttsk_drm* left,
                                 const ttsk_drm* right, double* d_out, void* stream) {
    TTSK_ARG(ctx != nullptr && h_shape && d_X && left && right && d_out, "NULL argument");
    TTSK_ARG(d >= 2 && d <= TTSK_MAX_ORDER, "tensor order must be in [2, 16]");
    TTSK_ARG(left->kind == TTSK_DRM_TT && right->kind == TTSK_DRM_TT, "ttsk_dense_sketch needs TT DRMs on both sides");
    TTSK_ARG(!left->right && right->right, "left/right DRM orientation mismatch");
    cudaStream_t st = (cudaStream_t)stream;
    TTSK_CUDA(cudaSetDevice(ctx->device));
    // the reference's dense sketch ignores rank slices (tensor_train_drm.py:109-122): whole cores only
    int rL[TTSK_MAX_ORDER], rR[TTSK_MAX_ORDER];
    for (int mu = 0; mu < d - 1; mu++) {
        TTSK_ARG(left->rank_min[mu] == 0 && left->rank_max[mu] == left->core_r1[mu], "dense sketch: the left DRM must not be sliced");
        TTSK_ARG(right->rank_min[mu] == 0 && right->rank_max[mu] == right->core_r1[d - 2 - mu], "dense sketch: the right DRM must not be sliced");
        rL[mu] = left->core_r1[mu];
        rR[mu] = right->core_r1[d - 2 - mu];
    }
    for (int k = 0; k < d - 1; k++) {
        TTSK_ARG(left->d_cores[k] && right->d_cores[k], "TT DRM core pointer is NULL");
        TTSK_ARG(left->core_r0[k] == (k == 0 ? 1 : left->core_r1[k - 1]), "left TT-DRM core ranks do not chain");
        TTSK_ARG(right->core_r0[k] == (k == 0 ? 1 : right->core_r1[k - 1]), "right TT-DRM core ranks do not chain");
    }
    int64_t psi_off[TTSK_MAX_ORDER], omega_off[TTSK_MAX_ORDER], off = 0;
    for (int mu = 0; mu < d; mu++) {
        psi_off[mu] = off;
        off += (int64_t)(mu == 0 ? 1 : rL[mu - 1]) * h_shape[mu] * (mu == d - 1 ? 1 : rR[mu]);
    }
    for (int mu = 0; mu < d - 1; mu++) {
        omega_off[mu] = off;
        off += (int64_t)rL[mu] * rR[mu];
    }
    // F[mu] = prod_{m > mu} n_m
    int64_t F[TTSK_MAX_ORDER];
    F[d - 1] = 1;
    for (int mu = d - 2; mu >= 0; mu--) F[mu] = F[mu + 1] * h_shape[mu + 1];
    // workspace: right unfoldings pc_k (F[d-2-k] x h1_k), left sweeps XL_mu (g1_mu x F[mu]), Y0 (n_0 x rR_0)
    int64_t bytes = 0;
    auto add = [&](int64_t n) { const int64_t at = bytes; bytes += align_up(n * 8, 256); return at; };
    int64_t pc_at[TTSK_MAX_ORDER], xl_at[TTSK_MAX_ORDER];
    const int rRp0 = (rR[0] + 1) & ~1;  // Rq_0 = pc_{d-2} is stored with an even row pitch (TMA rows are multiples of 16 bytes)
    for (int k = 0; k < d - 1; k++) pc_at[k] = add(F[d - 2 - k] * (k == d - 2 ? rRp0 : right->core_r1[k]));
    for (int mu = 0; mu < d - 1; mu++) xl_at[mu] = add((int64_t)rL[mu] * F[mu]);
    const int64_t y0_at = add(h_shape[0] * rR[0]);
    TTSK_TRY(ctx->ws_reserve(bytes + 4096));
    ctx->ws_reset();
    char* ws = (char*)ctx->ws_alloc(bytes);
    if (!ws) { set_error("workspace carve failed (dense sketch)"); return TTSK_E_NOMEM; }
    auto PC = [&](int k) { return (double*)(ws + pc_at[k]); };   // unfolding of right level k = bond d-2-k
    auto XL = [&](int mu) { return (double*)(ws + xl_at[mu]); };
    double* Y0 = (double*)(ws + y0_at);

    // ---- right DRM unfoldings (sketch_dense on the reversed shape): pc_0 = H_0 (n, h1), pc_k = (pc_{k-1} @ H_k).reshape(-1, h1_k)
    for (int k = 0; k < d - 1; k++) {
        const int64_t n = h_shape[d - 1 - k];
        const int64_t h0 = right->core_r0[k], h1 = right->core_r1[k];
        double* H = const_cast<double*>(right->d_cores[k]);
        const int64_t pitch = (k == d - 2) ? rRp0 : h1;
        if (k == 0) {
            TTSK_CUDA(cudaMemcpy2DAsync(PC(0), (size_t)pitch * 8, H, (size_t)h1 * 8, (size_t)h1 * 8, (size_t)n, cudaMemcpyDeviceToDevice, st));
        } else if (pitch == h1) {
            const int64_t rows = F[d - 1 - k];  // rows of pc_{k-1}
            TTSK_TRY(mm(ctx, Mat{PC(k - 1), rows, h0, h0, 1}, Mat{H, h0, n * h1, n * h1, 1}, 0.0, Mat{PC(k), rows, n * h1, n * h1, 1}, st));
        } else {  // padded rows: one GEMM per mode index j, C_j[row, b] at ((row n + j) pitch + b)
            const int64_t rows = F[d - 1 - k];
            TTSK_TRY(gemm_launch(ctx, rows, h1, h0, 1.0, PC(k - 1), h0, 1, H, n * h1, 1, 0.0, PC(k), n * pitch, 1, n, 0, h1, pitch, st));
        }
    }
    // ---- first pass over X: XL_0 (g1_0 x F_0) and Y0 = X (n_0 x F_0) @ Rq_0^T, Rq_0 = pc_{d-2} viewed (F_0 x rR_0)
    const int64_t n0 = h_shape[0];
    double* G0 = const_cast<double*>(left->d_cores[0]);  // (1, n_0, g1_0)
    double* X = const_cast<double*>(d_X);
    TTSK_CUDA(cudaMemsetAsync(Y0, 0, (size_t)n0 * rR[0] * 8, st));
    bool fused = false;
    TTSK_TRY(first_pass(ctx, X, n0, F[0], G0, rL[0], PC(d - 2), rR[0], rRp0, XL(0), Y0, st, &fused));
    if (!fused) {
        TTSK_TRY(mm(ctx, Mat{G0, rL[0], n0, 1, rL[0]}, Mat{X, n0, F[0], F[0], 1}, 0.0, Mat{XL(0), rL[0], F[0], F[0], 1}, st));
        TTSK_TRY(mm(ctx, Mat{X, n0, F[0], F[0], 1}, Mat{PC(d - 2), F[0], rR[0], rRp0, 1}, 0.0, Mat{Y0, n0, rR[0], rR[0], 1}, st));
    }
    // Psi_0 += Y0,  Omega_0 += G_0^T Y0
    TTSK_TRY(axpy_launch(ctx, n0 * rR[0], 1.0, Y0, d_out + psi_off[0], st));
    TTSK_TRY(mm(ctx, Mat{G0, rL[0], n0, 1, rL[0]}, Mat{Y0, n0, rR[0], rR[0], 1}, 1.0, Mat{d_out + omega_off[0], rL[0], rR[0], rR[0], 1}, st));
    // ---- left sweeps and the remaining sketches
    for (int mu = 1; mu < d; mu++) {
        const int64_t n = h_shape[mu];
        // Psi_mu: XL_{mu-1} as ((a, i_mu) x F[mu]) @ Rq_mu^T;   last mode: Psi_{d-1} = XL_{d-2}
        if (mu < d - 1) {
            const int k = d - 2 - mu;
            TTSK_TRY(mm(ctx, Mat{XL(mu - 1), rL[mu - 1] * n, F[mu], F[mu], 1}, Mat{PC(k), F[mu], rR[mu], rR[mu], 1}, 1.0,
                        Mat{d_out + psi_off[mu], rL[mu - 1] * n, rR[mu], rR[mu], 1}, st));
            // XL_mu = G_mu (as (g0 n) x g1)^T @ XL_{mu-1} (as (g0 n) x F[mu])
            double* G = const_cast<double*>(left->d_cores[mu]);
            const int64_t g0 = left->core_r0[mu];
            TTSK_TRY(mm(ctx, Mat{G, rL[mu], g0 * n, 1, rL[mu]}, Mat{XL(mu - 1), g0 * n, F[mu], F[mu], 1}, 0.0,
                        Mat{XL(mu), rL[mu], F[mu], F[mu], 1}, st));
            // Omega_mu += XL_mu @ Rq_mu^T
            TTSK_TRY(mm(ctx, Mat{XL(mu), rL[mu], F[mu], F[mu], 1}, Mat{PC(k), F[mu], rR[mu], rR[mu], 1}, 1.0,
                        Mat{d_out + omega_off[mu], rL[mu], rR[mu], rR[mu], 1}, st));
        } else {
            TTSK_TRY(axpy_launch(ctx, (int64_t)rL[d - 2] * n, 1.0, XL(d - 2), d_out + psi_off[d - 1], st));
        }
    }
    return TTSK_OK;
}
