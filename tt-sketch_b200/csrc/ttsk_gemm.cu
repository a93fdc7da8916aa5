// Strided, batched FP64 GEMM with split-K for the skinny/latency-bound contractions of the
// TT / CP / dense paths and the edge-Omega products of the sparse path.
// Replaces the NumPy einsum / matmul calls of tt_sketch/drm/tensor_train_drm.py:71-122,
// tt_sketch/sketching_methods/{tensor_train,cp,dense}_sketch.py.
//
// Shapes here are (r x n*r)-like with r = 10..100: far too small for tcgen05 (no FP64 kind
// anyway) -- the work is bound by launch latency and by streaming the one large operand
// once, so the kernel is a plain shared-memory tiled DFMA kernel whose grid is widened by
// split-K until it covers the 148 SMs.
#include "ttsk_common.cuh"

namespace ttsk {

constexpr int BM = 32, BN = 32, BK = 32;

struct GemmArgs {
    long long M, N, K;
    double alpha;
    const double* A; long long a_rs, a_cs, a_bs;
    const double* B; long long b_rs, b_cs, b_bs;
    double* C; long long c_rs, c_cs, c_bs;
    int splits; long long k_per_split;
    int use_atomic;  // accumulate into C with atomics (C already holds beta*C)
    double beta;
};

__global__ void scale_kernel(double* C, long long M, long long N, long long c_rs, long long c_cs, long long c_bs,
                             long long batch, double beta) {
    const long long total = M * N * batch;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long b = e / (M * N);
        const long long rem = e - b * M * N;
        const long long m = rem / N, n = rem - m * N;
        double* p = C + b * c_bs + m * c_rs + n * c_cs;
        *p = (beta == 0.0) ? 0.0 : (*p) * beta;
    }
}

__global__ void __launch_bounds__(256) gemm_kernel(GemmArgs g) {
    __shared__ double As[BK][BM + 1];
    __shared__ double Bs[BK][BN + 1];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 2 x 2 micro-tile each
    const long long m0 = (long long)blockIdx.y * BM, n0 = (long long)blockIdx.x * BN;
    const int split = blockIdx.z % g.splits;
    const long long batch = blockIdx.z / g.splits;
    const double* A = g.A + batch * g.a_bs;
    const double* B = g.B + batch * g.b_bs;
    double* C = g.C + batch * g.c_bs;
    const long long k_begin = (long long)split * g.k_per_split;
    long long k_end = k_begin + g.k_per_split;
    if (k_end > g.K) k_end = g.K;

    const bool a_kfast = (g.a_cs == 1);   // A row-major: k contiguous
    const bool b_nfast = (g.b_cs == 1);   // B row-major: n contiguous
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};

    for (long long k0 = k_begin; k0 < k_end; k0 += BK) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int mm, kk;
            if (a_kfast) { kk = tid & 31; mm = (tid >> 5) + 8 * i; }
            else         { mm = tid & 31; kk = (tid >> 5) + 8 * i; }
            const long long m = m0 + mm, k = k0 + kk;
            As[kk][mm] = (m < g.M && k < k_end) ? A[m * g.a_rs + k * g.a_cs] : 0.0;
            int nn, kb;
            if (b_nfast) { nn = tid & 31; kb = (tid >> 5) + 8 * i; }
            else         { kb = tid & 31; nn = (tid >> 5) + 8 * i; }
            const long long n = n0 + nn, k2 = k0 + kb;
            Bs[kb][nn] = (n < g.N && k2 < k_end) ? B[k2 * g.b_rs + n * g.b_cs] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            const double a0 = As[kk][ty], a1 = As[kk][ty + 16];
            const double b0 = Bs[kk][tx], b1 = Bs[kk][tx + 16];
            acc[0][0] = fma(a0, b0, acc[0][0]);
            acc[0][1] = fma(a0, b1, acc[0][1]);
            acc[1][0] = fma(a1, b0, acc[1][0]);
            acc[1][1] = fma(a1, b1, acc[1][1]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const long long m = m0 + ty + 16 * i, n = n0 + tx + 16 * j;
            if (m < g.M && n < g.N) {
                double* p = C + m * g.c_rs + n * g.c_cs;
                const double v = g.alpha * acc[i][j];
                if (g.use_atomic) atomicAdd(p, v);
                else *p = (g.beta == 0.0) ? v : v + g.beta * (*p);
            }
        }
}

int gemm_launch(ttsk_ctx* ctx, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t a_rs,
                int64_t a_cs, const double* B, int64_t b_rs, int64_t b_cs, double beta, double* C, int64_t c_rs,
                int64_t c_cs, int64_t batch, int64_t a_bs, int64_t b_bs, int64_t c_bs, cudaStream_t st) {
    if (M <= 0 || N <= 0 || batch <= 0) return TTSK_OK;
    const long long tm = (M + BM - 1) / BM, tn = (N + BN - 1) / BN;
    TTSK_ARG(tm <= 65535, "gemm: M too large for grid.y (reshape the problem)");
    GemmArgs g;
    g.M = M; g.N = N; g.K = K; g.alpha = alpha; g.beta = beta;
    g.A = A; g.a_rs = a_rs; g.a_cs = a_cs; g.a_bs = a_bs;
    g.B = B; g.b_rs = b_rs; g.b_cs = b_cs; g.b_bs = b_bs;
    g.C = C; g.c_rs = c_rs; g.c_cs = c_cs; g.c_bs = c_bs;
    // widen the grid with split-K until ~2 CTAs per SM
    long long tiles = tm * tn * batch;
    long long splits = 1;
    const long long target = 2LL * ctx->sm_count;
    if (tiles < target && K > 4 * BK) {
        splits = (target + tiles - 1) / tiles;
        const long long max_splits = (K + 4 * BK - 1) / (4 * BK);
        if (splits > max_splits) splits = max_splits;
        if (splits < 1) splits = 1;
    }
    long long kps = (K + splits - 1) / splits;
    kps = (kps + BK - 1) / BK * BK;
    splits = K > 0 ? (K + kps - 1) / kps : 1;
    if (splits < 1) splits = 1;
    TTSK_ARG(batch * splits <= 65535, "gemm: batch*splits too large for grid.z");
    g.splits = (int)splits;
    g.k_per_split = kps;
    g.use_atomic = splits > 1;
    if (g.use_atomic) {
        if (beta != 1.0) {
            long long total = M * N * batch;
            long long blocks = (total + 255) / 256;
            if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
            scale_kernel<<<(unsigned)blocks, 256, 0, st>>>(C, M, N, c_rs, c_cs, c_bs, batch, beta);
            TTSK_LAUNCHED(ctx);
        }
    }
    dim3 grid((unsigned)tn, (unsigned)tm, (unsigned)(batch * splits));
    gemm_kernel<<<grid, 256, 0, st>>>(g);
    TTSK_LAUNCHED(ctx);
    return TTSK_OK;
}

// out[j, k, m] = A[k, j] * Rm[j, m]
__global__ void khatri_rao_kernel(long long n, long long R, long long r, const double* __restrict__ A,
                                  const double* __restrict__ Rm, long long r_rs, double* __restrict__ out) {
    const long long total = R * n * r;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long j = e / (n * r);
        const long long rem = e - j * n * r;
        const long long k = rem / r, m = rem - k * r;
        out[e] = A[k * R + j] * Rm[j * r_rs + m];
    }
}

}  // namespace ttsk

extern "C" int ttsk_gemm(ttsk_ctx* ctx, int64_t M, int64_t N, int64_t K, double alpha, const double* d_A,
                         int64_t a_rs, int64_t a_cs, const double* d_B, int64_t b_rs, int64_t b_cs, double beta,
                         double* d_C, int64_t c_rs, int64_t c_cs, int64_t batch, int64_t a_bs, int64_t b_bs,
                         int64_t c_bs, void* stream) {
    TTSK_ARG(ctx != nullptr, "ctx is NULL");
    TTSK_ARG(M >= 0 && N >= 0 && K >= 0 && batch >= 0, "negative dimension");
    TTSK_ARG(d_C != nullptr || M * N * batch == 0, "C is NULL");
    return ttsk::gemm_launch(ctx, M, N, K, alpha, d_A, a_rs, a_cs, d_B, b_rs, b_cs, beta, d_C, c_rs, c_cs, batch,
                             a_bs, b_bs, c_bs, (cudaStream_t)stream);
}

extern "C" int ttsk_khatri_rao(ttsk_ctx* ctx, int64_t n, int64_t R, int64_t r, const double* d_A, const double* d_Rm,
                               int64_t r_rs, double* d_out, void* stream) {
    TTSK_ARG(ctx != nullptr, "ctx is NULL");
    const long long total = (long long)R * n * r;
    if (total <= 0) return TTSK_OK;
    long long blocks = (total + 255) / 256;
    if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
    ttsk::khatri_rao_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(n, R, r, d_A, d_Rm, r_rs, d_out);
    TTSK_LAUNCHED(ctx);
    return TTSK_OK;
}
