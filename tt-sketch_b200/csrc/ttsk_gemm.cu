// Strided, batched FP64 GEMM with split-K for the skinny/latency-bound contractions of the
// TT / CP / dense paths and the edge-Omega products of the sparse path.
// Replaces the NumPy einsum / matmul calls of tt_sketch/drm/tensor_train_drm.py:71-122,
// tt_sketch/sketching_methods/{tensor_train,cp,dense}_sketch.py.
//
// Shapes here are (r x n*r)-like with r = 10..100 in FP64: tcgen05 has no FP64 kind, the FP64 tensor
// instruction is mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4).  The work is bound by launch latency and by
// streaming the one large operand once, so the kernel is a shared-memory tiled DMMA kernel (a CTA
// computes a 64 x 32 tile, 16 x 32 for skinny M; tile pitches == 4 (mod 16) so the fragment loads are
// conflict free; arbitrary element strides, so transposes / slices / unfoldings are views) whose grid
// is widened by split-K until it covers the 148 SMs.  Chains of these launches are replayed as CUDA
// graphs by the host side (sketch_dispatch._graphed).
#include <algorithm>

#include "ttsk_common.cuh"

namespace ttsk {

constexpr int BN = 32, BK = 32;  // the tile height BM is a template parameter: 64, or 16 for skinny M
constexpr int PA = BK + 4;  // pitch of the A tile [m][k]: == 4 (mod 16) -> conflict-free a-fragment loads
constexpr int PB = BN + 4;  // pitch of the B tile [k][n]: == 4 (mod 16) -> conflict-free b-fragment loads

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

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// A CTA computes a BM x 32 tile of C over its K range with FP64 tensor-core MMAs (DMMA.8x8x4).  BM = 64: warp w
// owns rows 8w..8w+7 and the four 8-column tiles; BM = 16 (M <= 16: the DRM-rank-sized reductions over a huge K,
// where a 64-row tile would spend 4/5 of the FP64 pipe on padding): warp w owns row tile w & 1, column tile w >> 1.
// (The first version was a 2 x 2 register-tile DFMA kernel: one shared-memory load per FMA made it shared-memory
// bound at ~0.7 TB/s of operand streaming.)
template <int BM>
__global__ void __launch_bounds__(256) gemm_kernel(GemmArgs g) {
    constexpr int TPW = BM / 16;  // 8x8 output tiles per warp
    __shared__ double As[BM * PA];
    __shared__ double Bs[BK * PB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gq = lane >> 2, q = lane & 3;
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
    const int rt = BM == 64 ? warp : (warp & 1);   // row tile of this warp
    const int jc = BM == 64 ? 0 : (warp >> 1);     // its first column tile
    double acc[TPW][2];
#pragma unroll
    for (int j = 0; j < TPW; j++) acc[j][0] = acc[j][1] = 0.0;

    for (long long k0 = k_begin; k0 < k_end; k0 += BK) {
        const int kmax = (int)((k_end - k0 < BK) ? k_end - k0 : BK);
        // all global loads of the step are issued before the first shared-memory store (A / B are generic pointers:
        // interleaved, every load would have to wait for the store before it)
        constexpr int NA = BM / 8;  // A elements per thread
        double ra[NA], rb[4];
        // A tile: BM x BK elements; lanes run along the contiguous dimension
#pragma unroll
        for (int i = 0; i < NA; i++) {
            int mm, kk;
            if (a_kfast) { kk = tid & 31; mm = (tid >> 5) + 8 * i; }
            else if (BM == 64) { mm = (tid & 31) + 32 * (i & 1); kk = (tid >> 5) + 8 * (i >> 1); }
            else { mm = tid & 15; kk = (tid >> 4) + 16 * i; }
            const long long m = m0 + mm, k = k0 + kk;
            ra[i] = (m < g.M && kk < kmax) ? __ldg(A + m * g.a_rs + k * g.a_cs) : 0.0;
        }
        // B tile: BK x BN = 1024 elements, 4 per thread
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int nn, kb;
            if (b_nfast) { nn = tid & 31; kb = (tid >> 5) + 8 * i; }
            else         { kb = tid & 31; nn = (tid >> 5) + 8 * i; }
            const long long n = n0 + nn, k2 = k0 + kb;
            rb[i] = (n < g.N && kb < kmax) ? __ldg(B + k2 * g.b_rs + n * g.b_cs) : 0.0;
        }
#pragma unroll
        for (int i = 0; i < NA; i++) {
            int mm, kk;
            if (a_kfast) { kk = tid & 31; mm = (tid >> 5) + 8 * i; }
            else if (BM == 64) { mm = (tid & 31) + 32 * (i & 1); kk = (tid >> 5) + 8 * (i >> 1); }
            else { mm = tid & 15; kk = (tid >> 4) + 16 * i; }
            As[mm * PA + kk] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int nn, kb;
            if (b_nfast) { nn = tid & 31; kb = (tid >> 5) + 8 * i; }
            else         { kb = tid & 31; nn = (tid >> 5) + 8 * i; }
            Bs[kb * PB + nn] = rb[i];
        }
        __syncthreads();
        const int ksteps = (kmax + 3) >> 2;
        for (int s = 0; s < ksteps; s++) {
            const double a = As[(8 * rt + gq) * PA + 4 * s + q];
#pragma unroll
            for (int j = 0; j < TPW; j++) dmma884(acc[j][0], acc[j][1], a, Bs[(4 * s + q) * PB + 8 * (jc + j) + gq]);
        }
        __syncthreads();
    }
    const long long m = m0 + 8 * rt + gq;
    if (m < g.M) {
        // the two accumulators of a lane are adjacent columns: one 16-byte read-modify-write when C allows it
        const bool vec = !g.use_atomic && g.c_cs == 1 && (g.c_rs & 1) == 0 && (g.c_bs & 1) == 0 &&
                         (reinterpret_cast<unsigned long long>(g.C) & 15ull) == 0;
#pragma unroll
        for (int j = 0; j < TPW; j++) {
            const long long n = n0 + 8 * (jc + j) + 2 * q;
            if (vec && n + 1 < g.N) {
                double2* p = reinterpret_cast<double2*>(C + m * g.c_rs + n);
                double2 v = make_double2(g.alpha * acc[j][0], g.alpha * acc[j][1]);
                if (g.beta != 0.0) {
                    const double2 old = *p;
                    v.x += g.beta * old.x;
                    v.y += g.beta * old.y;
                }
                *p = v;
                continue;
            }
#pragma unroll
            for (int e = 0; e < 2; e++) {
                if (n + e < g.N) {
                    double* p = C + m * g.c_rs + (n + e) * g.c_cs;
                    const double v = g.alpha * acc[j][e];
                    if (g.use_atomic) atomicAdd(p, v);
                    else *p = (g.beta == 0.0) ? v : v + g.beta * (*p);
                }
            }
        }
    }
}

int gemm_launch(ttsk_ctx* ctx, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t a_rs,
                int64_t a_cs, const double* B, int64_t b_rs, int64_t b_cs, double beta, double* C, int64_t c_rs,
                int64_t c_cs, int64_t batch, int64_t a_bs, int64_t b_bs, int64_t c_bs, cudaStream_t st) {
    if (M <= 0 || N <= 0 || batch <= 0) return TTSK_OK;
    const int BM = M <= 32 ? 16 : 64;  // rank-sized M: short tiles waste less of the FP64 pipe on padding
    const long long tm = (M + BM - 1) / BM, tn = (N + BN - 1) / BN;
    TTSK_ARG(tm <= 65535, "gemm: M too large for grid.y (reshape the problem)");
    GemmArgs g;
    g.M = M; g.N = N; g.K = K; g.alpha = alpha; g.beta = beta;
    g.A = A; g.a_rs = a_rs; g.a_cs = a_cs; g.a_bs = a_bs;
    g.B = B; g.b_rs = b_rs; g.b_cs = b_cs; g.b_bs = b_bs;
    g.C = C; g.c_rs = c_rs; g.c_cs = c_cs; g.c_bs = c_bs;
    // widen the grid with split-K until ~4 CTAs per SM
    long long tiles = tm * tn * batch;
    long long splits = 1;
    const long long target = 4LL * ctx->sm_count;
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
    if (BM == 16) gemm_kernel<16><<<grid, 256, 0, st>>>(g);
    else gemm_kernel<64><<<grid, 256, 0, st>>>(g);
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
