// Stand-alone lazy Gaussian DRM kernels (operator-level entry point + prefix tables).
// Replaces inds_to_normal, tt_sketch/drm/fast_lazy_gaussian.pyx:183-201.
#include "ttsk_common.cuh"
#include "ttsk_gauss.cuh"

namespace ttsk {

struct GaussIdx {
    const long long* rows[TTSK_MAX_ORDER];
    long long strides[TTSK_MAX_ORDER];  // int32-wrapped, sign-extended (pyx:60-71)
    int k;
};

constexpr int kMaxSalts = 2048;

// one thread per (nonzero p, column a); out is (nnz, rank) row-major so stores coalesce and
// the k index loads of a warp hit one or two addresses.
__global__ void __launch_bounds__(256) lazy_gaussian_kernel(GaussIdx gi, long long nnz, int rank_min, int rank,
                                                           unsigned long long seed, double* __restrict__ out) {
    __shared__ double2 s_tab[kGaussTabEntries];
    __shared__ unsigned long long s_salt[kMaxSalts];
    load_logtab(s_tab);
    for (int a = threadIdx.x; a < rank && a < kMaxSalts; a += blockDim.x)
        s_salt[a] = hash64((unsigned long long)(rank_min + a)) + seed;
    __syncthreads();
    const long long total = nnz * (long long)rank;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const long long p = e / rank;
        const int a = (int)(e - p * rank);
        unsigned long long flat = 0;
        if (gi.k > 0) {
            flat = (unsigned long long)gi.rows[0][p];
            for (int i = 1; i < gi.k; i++)
                flat += (unsigned long long)gi.rows[i][p] * (unsigned long long)gi.strides[i];
        } else {
            flat = (unsigned long long)p;  // table mode: the flat index IS the row number
        }
        const unsigned long long salt =
            a < kMaxSalts ? s_salt[a] : hash64((unsigned long long)(rank_min + a)) + seed;
        out[e] = ndtri_any(uniform_from_hash(hash64(flat + salt)), s_tab);
    }
}

// strides with the reference's C `int prod` wrap: truncated to 32 bits, sign-extended.
void wrapped_strides(const int64_t* shape, int k, long long* strides) {
    int32_t prod = (int32_t)(uint32_t)(uint64_t)shape[0];
    strides[0] = 1;
    for (int i = 1; i < k; i++) {
        strides[i] = (long long)prod;
        prod = (int32_t)((uint32_t)prod * (uint32_t)(uint64_t)shape[i]);
    }
}

int gauss_rows_launch(ttsk_ctx* ctx, const GaussIdx& gi, int64_t nnz, int rank_min, int rank, uint64_t seed,
                      double* d_out, cudaStream_t st) {
    if (nnz <= 0 || rank <= 0) return TTSK_OK;
    const long long total = (long long)nnz * rank;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    lazy_gaussian_kernel<<<(unsigned)blocks, 256, 0, st>>>(gi, nnz, rank_min, rank, seed, d_out);
    TTSK_LAUNCHED(ctx);
    return TTSK_OK;
}

// table of the first `rows` flat indices: out (rows, rank)
int gauss_table_launch(ttsk_ctx* ctx, int64_t rows, int rank_min, int rank, uint64_t seed, double* d_out,
                       cudaStream_t st) {
    GaussIdx gi;
    gi.k = 0;
    return gauss_rows_launch(ctx, gi, rows, rank_min, rank, seed, d_out, st);
}

// ------------------------------------------------------------------ sparse sign DRM
// Replaces inds_to_sparse_sign / _inds_to_sparse_sign, tt_sketch/drm/fast_lazy_gaussian.pyx:121-180: per nonzero,
// nnz_row hashed doubles (the Gaussian DRM's hash with the top bits forced to 001, columns 0..nnz_row-1); frexp
// splits each into an exponent whose parity is the sign (Python-style e % 2: hash bit 52, entries -1 / +1) and a
// mantissa m * 2 - 1 (the low 52 hash bits as a uniform) that drives a partial Fisher-Yates shuffle of the row.
// One thread per nonzero; the row lives in local memory.  out is (nnz, rank_max - rank_min) row-major FP64 (what
// the sketching operators consume).
constexpr int kMaxSignRank = 512;

__global__ void __launch_bounds__(128) lazy_sparse_sign_kernel(GaussIdx gi, long long nnz, int rank, int rank_min, int width,
                                                              int nnz_row, unsigned long long seed, double* __restrict__ out) {
    short row[kMaxSignRank];
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < nnz; p += (long long)gridDim.x * blockDim.x) {
        unsigned long long flat = (unsigned long long)gi.rows[0][p];
        for (int i = 1; i < gi.k; i++) flat += (unsigned long long)gi.rows[i][p] * (unsigned long long)gi.strides[i];
        for (int j = 0; j < rank; j++) row[j] = 0;
        for (int j = 0; j < nnz_row; j++) {
            const unsigned long long h = hash64(flat + hash64((unsigned long long)j) + seed);
            row[j] = (short)((int)((h >> 52) & 1ull) * 2 - 1);  // parity of the forced-001 double's frexp exponent
        }
        for (int j = 0; j < nnz_row; j++) {
            const double m = uniform_from_hash(hash64(flat + hash64((unsigned long long)j) + seed));  // frexp mantissa * 2 - 1
            int rn = __double2int_rz(__dadd_rn(__dmul_rn(m, (double)(rank - j)), (double)j));
            rn = rn < 0 ? 0 : (rn >= rank ? rank - 1 : rn);
            const short t = row[j];
            row[j] = row[rn];
            row[rn] = t;
        }
        for (int a = 0; a < width; a++) out[p * width + a] = (double)row[rank_min + a];
    }
}

// self-test: div_rn_safe vs __ddiv_rn on pseudo-random operands in the ranges ndtri uses
__global__ void selftest_div_kernel(long long n, unsigned long long seed, unsigned long long* mismatches) {
    unsigned long long bad = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long h1 = hash64(seed + 2 * i), h2 = hash64(seed + 2 * i + 1);
        const double ua = uniform_from_hash(h1), ub = uniform_from_hash(h2);
        // numerators 2^-120..2^10, denominators 2^-30..2^10 (log-uniform), random signs, some exact zeros
        const int ea = (int)((h1 >> 52) % 131) - 120, eb = (int)((h2 >> 52) % 41) - 30;
        double a = ldexp(1.0 + ua, ea), b = ldexp(1.0 + ub, eb);
        if (h1 >> 63) a = -a;
        if (h2 >> 63) b = -b;
        if ((i & 1023) == 0) a = 0.0;
        const double want = __ddiv_rn(a, b), got = div_rn_safe(a, b);
        if (__double_as_longlong(want) != __double_as_longlong(got)) bad++;
    }
    if (bad) atomicAdd(mismatches, bad);
}

// self-test: sqrt_rn_safe vs __dsqrt_rn on pseudo-random operands 2^-8 .. 2^12 (ndtri uses (4, 80))
__global__ void selftest_sqrt_kernel(long long n, unsigned long long seed, unsigned long long* mismatches) {
    unsigned long long bad = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long h1 = hash64(seed + i);
        const int ea = (int)((h1 >> 52) % 21) - 8;
        const double a = ldexp(1.0 + uniform_from_hash(h1), ea);
        const double want = __dsqrt_rn(a), got = sqrt_rn_safe(a);
        if (__double_as_longlong(want) != __double_as_longlong(got)) bad++;
    }
    if (bad) atomicAdd(mismatches, bad);
}

}  // namespace ttsk

extern "C" int ttsk_selftest_sqrt(ttsk_ctx* ctx, int64_t n, uint64_t seed, uint64_t* h_mismatches) {
    TTSK_ARG(ctx != nullptr && h_mismatches != nullptr && n >= 0, "selftest_sqrt");
    unsigned long long* d = nullptr;
    TTSK_CUDA(cudaMalloc((void**)&d, 8));
    TTSK_CUDA(cudaMemset(d, 0, 8));
    ttsk::selftest_sqrt_kernel<<<ctx->sm_count * 8, 256>>>(n, seed, d);
    TTSK_LAUNCHED(ctx);
    TTSK_CUDA(cudaMemcpy(h_mismatches, d, 8, cudaMemcpyDeviceToHost));
    TTSK_CUDA(cudaFree(d));
    return TTSK_OK;
}

extern "C" int ttsk_selftest_div(ttsk_ctx* ctx, int64_t n, uint64_t seed, uint64_t* h_mismatches) {
    TTSK_ARG(ctx != nullptr && h_mismatches != nullptr && n >= 0, "selftest_div");
    unsigned long long* d = nullptr;
    TTSK_CUDA(cudaMalloc((void**)&d, 8));
    TTSK_CUDA(cudaMemset(d, 0, 8));
    ttsk::selftest_div_kernel<<<ctx->sm_count * 8, 256>>>(n, seed, d);
    TTSK_LAUNCHED(ctx);
    TTSK_CUDA(cudaMemcpy(h_mismatches, d, 8, cudaMemcpyDeviceToHost));
    TTSK_CUDA(cudaFree(d));
    return TTSK_OK;
}

extern "C" int ttsk_lazy_sparse_sign(ttsk_ctx* ctx, const int64_t* d_idx, int64_t idx_row_stride, int k, int64_t nnz,
                                     const int64_t* h_shape, int rank, int rank_min, int rank_max, int nnz_row,
                                     uint64_t seed, double* d_out, void* stream) {
    TTSK_ARG(ctx != nullptr, "ctx is NULL");
    TTSK_ARG(k >= 1 && k <= TTSK_MAX_ORDER, "k out of range");
    TTSK_ARG(nnz >= 0 && rank_max >= rank_min && rank_min >= 0 && rank_max <= rank, "nnz/rank range");
    TTSK_ARG(rank >= 1 && rank <= ttsk::kMaxSignRank, "sparse sign DRM: rank must be in [1, 512]");
    TTSK_ARG(nnz_row >= 1 && nnz_row <= rank, "sparse sign DRM: non-zeros per row must be in [1, rank]");
    TTSK_ARG(h_shape != nullptr && (nnz == 0 || (d_idx && d_out)), "NULL pointer");
    if (nnz == 0 || rank_max == rank_min) return TTSK_OK;
    ttsk::GaussIdx gi;
    gi.k = k;
    for (int i = 0; i < k; i++) gi.rows[i] = (const long long*)(d_idx + i * idx_row_stride);
    ttsk::wrapped_strides(h_shape, k, gi.strides);
    long long blocks = (nnz + 127) / 128;
    if (blocks > (long long)ctx->sm_count * 16) blocks = (long long)ctx->sm_count * 16;
    ttsk::lazy_sparse_sign_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(gi, nnz, rank, rank_min, rank_max - rank_min,
                                                                                     nnz_row, seed, d_out);
    TTSK_LAUNCHED(ctx);
    return TTSK_OK;
}

extern "C" int ttsk_lazy_gaussian(ttsk_ctx* ctx, const int64_t* d_idx, int64_t idx_row_stride, int k, int64_t nnz,
                                  const int64_t* h_shape, int rank_min, int rank_max, uint64_t seed, double* d_out,
                                  void* stream) {
    TTSK_ARG(ctx != nullptr, "ctx is NULL");
    TTSK_ARG(k >= 1 && k <= TTSK_MAX_ORDER, "k out of range");
    TTSK_ARG(nnz >= 0 && rank_max >= rank_min && rank_min >= 0, "nnz/rank range");
    TTSK_ARG(h_shape != nullptr && (nnz == 0 || (d_idx && d_out)), "NULL pointer");
    ttsk::GaussIdx gi;
    gi.k = k;
    for (int i = 0; i < k; i++) gi.rows[i] = (const long long*)(d_idx + i * idx_row_stride);
    ttsk::wrapped_strides(h_shape, k, gi.strides);
    return ttsk::gauss_rows_launch(ctx, gi, nnz, rank_min, rank_max - rank_min, seed, d_out, (cudaStream_t)stream);
}
