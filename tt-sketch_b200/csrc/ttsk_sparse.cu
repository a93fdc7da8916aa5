// Sparse (COO) input: fused gather -> lazy-Gaussian / table / chain-row sources -> segment
// accumulation of Psi_mu (and the middle Omega_mu) with FP64 tensor-core MMAs.
//
// Replaces (reference): general_sketch for SparseTensor (tt_sketch/sketch_dispatch.py:202-275),
// SparseGaussianDRM.sketch_sparse (drm/sparse_gaussian_drm.py:29-44), TensorTrainDRM.sketch_sparse
// (drm/tensor_train_drm.py:60-69), sketch_omega_sparse / sketch_psi_sparse
// (sketching_methods/sparse_sketch.py:8-69).
//
// Design:
//   * the reference materialises (r x nnz) DRM rows for every bond and scans all nonzeros once
//     per slice j of every mode (O(n_mu * nnz)).  Here nonzeros are bucketed once per mode
//     (counting sort -> permutation + sorted keys) and a mode pass walks the buckets: a CTA
//     takes a contiguous piece of the sorted order, forms for TN nonzeros at a time the tiles
//         At (rA x TN) = v_p * L_{mu-1}(p),  Bt (rB x TN) = R_mu(p),  Xt (rX x TN) = R_{mu-1}(p)
//     in shared memory and accumulates  Psi_mu[:, j, :] += At Bt^T  (and Omega_{mu-1} += At Xt^T)
//     in registers with mma.sync.m8n8k4.f64, writing each slice once per piece.
//   * tile entries come from a "source": hash-seeded Gaussian generated on the fly (never
//     stored in HBM), a small prefix table of the same generator gathered through L2 when the
//     prefix index space is much smaller than nnz, or rows of a per-nonzero buffer (TT-DRM
//     chain products / user-supplied rows).
//   * the tail branch of ndtri (27% of draws) is deferred to a dense second phase so the
//     central branch runs without the tail's divergence.
//   * edge bonds need no per-nonzero work:  Omega_0 = L_0^T Psi_0,  Omega_{d-2} = Psi_{d-1} R_{d-2}^T.
#include <algorithm>
#include <cstring>

#include "ttsk_common.cuh"
#include "ttsk_gauss.cuh"

namespace ttsk {

int gauss_table_launch(ttsk_ctx* ctx, int64_t rows, int rank_min, int rank, uint64_t seed, double* d_out,
                       cudaStream_t st);
void wrapped_strides(const int64_t* shape, int k, long long* strides);

enum { SRC_NONE = 0, SRC_GAUSS = 1, SRC_ROWS = 2, SRC_TABLE = 3 };

struct Source {
    int kind;
    int r;         // columns produced
    int k;         // modes in the flat index (GAUSS / TABLE)
    int rank_min;  // GAUSS: first column of the infinite matrix
    int modes[TTSK_MAX_ORDER];
    long long strides[TTSK_MAX_ORDER];
    unsigned long long seed;
    const double* base;  // ROWS / TABLE: element (q, a) at base[q*row_stride + a*col_stride]
    long long row_stride, col_stride;
};

struct PassParams {
    int d;
    long long nnz;
    const long long* idx[TTSK_MAX_ORDER];
    const double* val;
    const int* perm;  // sorted order -> nonzero id (nullptr: identity)
    const int* skey;  // sorted keys (nullptr: all zero)
    long long n_mu;
    Source A, B, X;
    int rA, rB, rX;  // logical tile heights (1 for SRC_NONE)
    double* psi;     // (rA, n_mu, rB)
    double* omega;   // (rA, rX), only with X
    int piece;
};

// ------------------------------------------------------------------ bucketing (counting sort)
__global__ void hist_kernel(const long long* __restrict__ idx, long long nnz, int* __restrict__ hist) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < nnz;
         p += (long long)gridDim.x * blockDim.x)
        atomicAdd(&hist[idx[p]], 1);
}

// exclusive scan of hist[0..n) into offs[0..n] (offs[n] = total) and a copy into cursor; one CTA.
__global__ void __launch_bounds__(1024) scan_kernel(const int* __restrict__ hist, long long n, int* __restrict__ offs,
                                                    int* __restrict__ cursor) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long base = 0; base < n; base += 1024) {
        const long long i = base + threadIdx.x;
        const int v = (i < n) ? hist[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const int carry = s_carry;
        const int excl = carry + (warp > 0 ? s_warp[warp - 1] : 0) + x - v;
        if (i < n) { offs[i] = excl; cursor[i] = excl; }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + s_warp[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) offs[n] = s_carry;
}

__global__ void scatter_kernel(const long long* __restrict__ idx, long long nnz, int* __restrict__ cursor,
                               int* __restrict__ perm, int* __restrict__ skey) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < nnz;
         p += (long long)gridDim.x * blockDim.x) {
        const int key = (int)idx[p];
        const int pos = atomicAdd(&cursor[key], 1);
        perm[pos] = (int)p;
        skey[pos] = key;
    }
}

// ------------------------------------------------------------------ TT-DRM chain step
// v_out[p, b] = sum_a v_in[p, a] * core[a, idx[p], b]   (first core: v_out[p, b] = core[0, idx[p], b])
__global__ void __launch_bounds__(256) ttdrm_step_kernel(long long nnz, const long long* __restrict__ idx,
                                                        const double* __restrict__ v_in, int r_in,
                                                        const double* __restrict__ core, long long n, int r_out,
                                                        double* __restrict__ v_out) {
    const long long total = nnz * (long long)r_out;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long p = e / r_out;
        const int b = (int)(e - p * r_out);
        const long long j = idx[p];
        double s;
        if (v_in == nullptr) {
            s = core[j * r_out + b];
        } else {
            s = 0.0;
            const double* c = core + j * r_out + b;
            const double* v = v_in + p * r_in;
            for (int a = 0; a < r_in; a++) s = fma(v[a], c[(long long)a * n * r_out], s);
        }
        v_out[e] = s;
    }
}

// ------------------------------------------------------------------ the mode pass
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

constexpr int kPassThreads = 256;
constexpr int kPassWarps = kPassThreads / 32;

template <int TN>
__device__ __forceinline__ void fill_source(const Source& S, int src_id, double* __restrict__ tile, int len,
                                            const unsigned long long* __restrict__ s_flat,
                                            const unsigned long long* __restrict__ s_salt,
                                            const double* __restrict__ s_val, bool scale, int* __restrict__ s_queue,
                                            int* __restrict__ s_qcount) {
    constexpr int TNP = TN + 4;
    const int tid = threadIdx.x, lane = tid & 31;
    if (S.kind == SRC_GAUSS) {
        const int total = S.r * TN;  // multiple of 32: whole warps stay together
        for (int e0 = 0; e0 < total; e0 += kPassThreads) {
            const int e = e0 + tid;
            const bool in = e < total;
            const int a = in ? e / TN : 0, p = e % TN;
            bool tail = false;
            int enc = 0;
            if (in) {
                double out = 0.0;
                if (p < len) {
                    const double u = uniform_from_hash(hash64(s_flat[p] + s_salt[a]));
                    const int cls = ndtri_class(u);
                    if (cls == 0) {
                        out = ndtri_central(u);
                        if (scale) out *= s_val[p];
                    } else {
                        out = u;
                        tail = true;
                        enc = (src_id << 28) | (cls << 26) | (a * TNP + p);
                    }
                }
                tile[a * TNP + p] = out;
            }
            const unsigned m = __ballot_sync(0xffffffffu, tail);
            if (m) {
                int base = 0;
                if (lane == 0) base = atomicAdd(s_qcount, __popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (tail) s_queue[base + __popc(m & ((1u << lane) - 1u))] = enc;
            }
        }
    } else if (S.kind == SRC_ROWS || S.kind == SRC_TABLE) {
        const int total = S.r * TN;
        for (int e = tid; e < total; e += kPassThreads) {
            const int p = e / S.r, a = e - p * S.r;  // column fastest: row reads coalesce
            double out = 0.0;
            if (p < len) {
                out = S.base[(long long)s_flat[p] * S.row_stride + (long long)a * S.col_stride];
                if (scale) out *= s_val[p];
            }
            tile[a * TNP + p] = out;
        }
    }
}

// MI/NJ: 8x8 MMA tiles covering rA / max(rB, rX).  HAS_X: also accumulate Omega = At Xt^T.
template <int MI, int NJ, bool HAS_X, int TN>
__global__ void __launch_bounds__(kPassThreads) sparse_pass_kernel(const PassParams P) {
    constexpr int TNP = TN + 4;
    constexpr int RAP = 8 * MI, RBP = 8 * NJ, RXP = HAS_X ? 8 * NJ : 0;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* s_tab = reinterpret_cast<double2*>(smem_raw);                       // 128 x 16 B
    double* At = reinterpret_cast<double*>(s_tab + 128);                         // [RAP][TNP]
    double* Bt = At + RAP * TNP;                                                 // [RBP][TNP]
    double* Xt = Bt + RBP * TNP;                                                 // [RXP][TNP]
    unsigned long long* s_salt = reinterpret_cast<unsigned long long*>(Xt + RXP * TNP);  // [RAP+RBP+RXP]
    unsigned long long* s_flat = s_salt + (RAP + RBP + RXP);                     // [3][TN]
    double* s_val = reinterpret_cast<double*>(s_flat + 3 * TN);                  // [TN]
    int* s_queue = reinterpret_cast<int*>(s_val + TN);                           // [(RAP+RBP+RXP)*TN]
    int* s_misc = s_queue + (RAP + RBP + RXP) * TN;                              // len, key, qcount

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;

    load_logtab(s_tab);
    for (int i = tid; i < (RAP + RBP + RXP) * TNP; i += kPassThreads) At[i] = 0.0;
    if (P.A.kind == SRC_GAUSS)
        for (int a = tid; a < P.A.r; a += kPassThreads)
            s_salt[a] = hash64((unsigned long long)(P.A.rank_min + a)) + P.A.seed;
    if (P.B.kind == SRC_GAUSS)
        for (int a = tid; a < P.B.r; a += kPassThreads)
            s_salt[RAP + a] = hash64((unsigned long long)(P.B.rank_min + a)) + P.B.seed;
    if (HAS_X && P.X.kind == SRC_GAUSS)
        for (int a = tid; a < P.X.r; a += kPassThreads)
            s_salt[RAP + RBP + a] = hash64((unsigned long long)(P.X.rank_min + a)) + P.X.seed;

    double acc[MI][NJ][2];
    double acc_o[HAS_X ? MI : 1][HAS_X ? NJ : 1][2];
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NJ; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    if (HAS_X) {
#pragma unroll
        for (int i = 0; i < (HAS_X ? MI : 1); i++)
#pragma unroll
            for (int j = 0; j < (HAS_X ? NJ : 1); j++) acc_o[i][j][0] = acc_o[i][j][1] = 0.0;
    }
    __syncthreads();

    auto flush_psi = [&](long long key) {
#pragma unroll
        for (int i = 0; i < MI; i++)
#pragma unroll
            for (int j = 0; j < NJ; j++) {
                const int row = 8 * i + g, col = 8 * j + 2 * q;
                if (row < P.rA) {
                    double* dst = P.psi + ((long long)row * P.n_mu + key) * P.rB + col;
                    if (col < P.rB && acc[i][j][0] != 0.0) atomicAdd(dst, acc[i][j][0]);
                    if (col + 1 < P.rB && acc[i][j][1] != 0.0) atomicAdd(dst + 1, acc[i][j][1]);
                }
                acc[i][j][0] = acc[i][j][1] = 0.0;
            }
    };

    const long long n_pieces = (P.nnz + P.piece - 1) / P.piece;
    for (long long piece = blockIdx.x; piece < n_pieces; piece += gridDim.x) {
        const long long s = piece * P.piece;
        const long long e = (s + P.piece < P.nnz) ? s + P.piece : P.nnz;
        long long cur_key = -1;
        long long c = s;
        while (c < e) {
            // ---- tile header: run of equal keys starting at c, at most TN long
            if (warp == 0) {
                int len = 0;
                long long key0 = 0;
                if (P.skey == nullptr) {
                    len = (int)((e - c < TN) ? e - c : TN);
                } else {
                    key0 = P.skey[c];
                    bool open = true;
#pragma unroll
                    for (int t = 0; t < TN / 32; t++) {
                        const long long pos = c + t * 32 + lane;
                        const bool same = (pos < e) && (P.skey[pos] == (int)key0);
                        const unsigned m = __ballot_sync(0xffffffffu, same);
                        if (open) {
                            if (m == 0xffffffffu) len += 32;
                            else { len += __ffs(~m) - 1; open = false; }
                        }
                    }
                }
                if (lane == 0) { s_misc[0] = len; s_misc[1] = (int)key0; s_misc[2] = 0; }
            }
            __syncthreads();
            const int len = s_misc[0];
            const long long key = s_misc[1];
            if (key != cur_key) {
                if (cur_key >= 0) flush_psi(cur_key);
                cur_key = key;
            }
            // ---- per-nonzero metadata
            if (tid < TN) {
                double v = 0.0;
                unsigned long long fa = 0, fb = 0, fx = 0;
                if (tid < len) {
                    const long long id = P.perm ? (long long)P.perm[c + tid] : c + tid;
                    v = P.val[id];
                    if (P.A.kind == SRC_ROWS) fa = (unsigned long long)id;
                    else if (P.A.kind != SRC_NONE)
                        for (int i = 0; i < P.A.k; i++)
                            fa += (unsigned long long)P.idx[P.A.modes[i]][id] * (unsigned long long)P.A.strides[i];
                    if (P.B.kind == SRC_ROWS) fb = (unsigned long long)id;
                    else if (P.B.kind != SRC_NONE)
                        for (int i = 0; i < P.B.k; i++)
                            fb += (unsigned long long)P.idx[P.B.modes[i]][id] * (unsigned long long)P.B.strides[i];
                    if (HAS_X) {
                        if (P.X.kind == SRC_ROWS) fx = (unsigned long long)id;
                        else
                            for (int i = 0; i < P.X.k; i++)
                                fx += (unsigned long long)P.idx[P.X.modes[i]][id] * (unsigned long long)P.X.strides[i];
                    }
                }
                s_val[tid] = v;
                s_flat[tid] = fa;
                s_flat[TN + tid] = fb;
                s_flat[2 * TN + tid] = fx;
                if (P.A.kind == SRC_NONE) At[tid] = v;                      // Psi_0: 1 x rB, scaled by v
                if (P.B.kind == SRC_NONE) Bt[tid] = (tid < len) ? 1.0 : 0.0;  // Psi_{d-1}: rA x 1
            }
            __syncthreads();
            // ---- sources -> tiles (central branch inline, tails queued)
            fill_source<TN>(P.A, 0, At, len, s_flat, s_salt, s_val, true, s_queue, &s_misc[2]);
            fill_source<TN>(P.B, 1, Bt, len, s_flat + TN, s_salt + RAP, s_val, false, s_queue, &s_misc[2]);
            if (HAS_X)
                fill_source<TN>(P.X, 2, Xt, len, s_flat + 2 * TN, s_salt + RAP + RBP, s_val, false, s_queue,
                                &s_misc[2]);
            __syncthreads();
            // ---- deferred tails, dense over the queue
            {
                const int nq = s_misc[2];
                for (int qi = tid; qi < nq; qi += kPassThreads) {
                    const int enc = s_queue[qi];
                    const int src = enc >> 28, cls = (enc >> 26) & 3, off = enc & 0x3ffffff;
                    double* tile = src == 0 ? At : (src == 1 ? Bt : Xt);
                    double out = ndtri_tail(tile[off], cls, s_tab);
                    if (src == 0) out *= s_val[off % TNP];
                    tile[off] = out;
                }
            }
            __syncthreads();
            // ---- accumulate: each warp takes k-chunks of 4 nonzeros
            for (int ch = warp; ch * 4 < len; ch += kPassWarps) {
                const int p0 = ch * 4 + q;
                double a[MI], b[NJ];
#pragma unroll
                for (int i = 0; i < MI; i++) a[i] = At[(8 * i + g) * TNP + p0];
#pragma unroll
                for (int j = 0; j < NJ; j++) b[j] = Bt[(8 * j + g) * TNP + p0];
#pragma unroll
                for (int i = 0; i < MI; i++)
#pragma unroll
                    for (int j = 0; j < NJ; j++) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
                if (HAS_X) {
#pragma unroll
                    for (int j = 0; j < NJ; j++) b[j] = Xt[(8 * j + g) * TNP + p0];
#pragma unroll
                    for (int i = 0; i < (HAS_X ? MI : 1); i++)
#pragma unroll
                        for (int j = 0; j < (HAS_X ? NJ : 1); j++)
                            dmma(acc_o[i][j][0], acc_o[i][j][1], a[i], b[j]);
                }
            }
            c += len;
        }
        if (cur_key >= 0) flush_psi(cur_key);
    }
    if (HAS_X) {
#pragma unroll
        for (int i = 0; i < (HAS_X ? MI : 1); i++)
#pragma unroll
            for (int j = 0; j < (HAS_X ? NJ : 1); j++) {
                const int row = 8 * i + g, col = 8 * j + 2 * q;
                if (row < P.rA) {
                    double* dst = P.omega + (long long)row * P.rX + col;
                    if (col < P.rX && acc_o[i][j][0] != 0.0) atomicAdd(dst, acc_o[i][j][0]);
                    if (col + 1 < P.rX && acc_o[i][j][1] != 0.0) atomicAdd(dst + 1, acc_o[i][j][1]);
                }
            }
    }
}

template <int MI, int NJ, bool HAS_X, int TN>
static int launch_pass_t(ttsk_ctx* ctx, const PassParams& P, cudaStream_t st) {
    constexpr int TNP = TN + 4;
    constexpr int R = 8 * MI + 8 * NJ + (HAS_X ? 8 * NJ : 0);
    const size_t smem = 128 * 16 + (size_t)R * TNP * 8 + (size_t)R * 8 + 3 * TN * 8 + TN * 8 +
                        (size_t)R * TN * 4 + 16;
    auto kern = sparse_pass_kernel<MI, NJ, HAS_X, TN>;
    TTSK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    TTSK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kPassThreads, smem));
    if (per_sm < 1) per_sm = 1;
    const long long n_pieces = (P.nnz + P.piece - 1) / P.piece;
    long long grid = (long long)ctx->sm_count * per_sm;
    if (grid > n_pieces) grid = n_pieces;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, kPassThreads, smem, st>>>(P);
    TTSK_LAUNCHED(ctx);
    return TTSK_OK;
}

template <bool HAS_X>
static int launch_pass_x(ttsk_ctx* ctx, const PassParams& P, cudaStream_t st) {
    const int mi = (P.rA + 7) / 8;
    const int nj = (std::max(P.rB, HAS_X ? P.rX : 1) + 7) / 8;
    TTSK_ARG(mi <= 8 && nj <= 8, "sparse pass: DRM rank above 64 is not supported by the fused kernel");
#define TTSK_PASS(MI_, NJ_) return launch_pass_t<MI_, NJ_, HAS_X, 32>(ctx, P, st)
    const int MIr = mi <= 1 ? 1 : (mi <= 3 ? 3 : (mi <= 5 ? 5 : 8));
    const int NJr = nj <= 1 ? 1 : (nj <= 3 ? 3 : (nj <= 5 ? 5 : 8));
    switch (MIr * 10 + NJr) {
        case 11: TTSK_PASS(1, 1);
        case 13: TTSK_PASS(1, 3);
        case 15: TTSK_PASS(1, 5);
        case 18: TTSK_PASS(1, 8);
        case 31: TTSK_PASS(3, 1);
        case 33: TTSK_PASS(3, 3);
        case 35: TTSK_PASS(3, 5);
        case 38: TTSK_PASS(3, 8);
        case 51: TTSK_PASS(5, 1);
        case 53: TTSK_PASS(5, 3);
        case 55: TTSK_PASS(5, 5);
        case 58: TTSK_PASS(5, 8);
        case 81: TTSK_PASS(8, 1);
        case 83: TTSK_PASS(8, 3);
        case 85: TTSK_PASS(8, 5);
        case 88: TTSK_PASS(8, 8);
    }
#undef TTSK_PASS
    set_error("sparse pass: no kernel variant");
    return TTSK_E_ARG;
}

static int launch_pass(ttsk_ctx* ctx, const PassParams& P, bool has_x, cudaStream_t st) {
    if (P.nnz <= 0) return TTSK_OK;
    return has_x ? launch_pass_x<true>(ctx, P, st) : launch_pass_x<false>(ctx, P, st);
}

// ------------------------------------------------------------------ host-side planning
struct SketchLayout {
    int d;
    int64_t shape[TTSK_MAX_ORDER];
    int rL[TTSK_MAX_ORDER], rR[TTSK_MAX_ORDER];
    int64_t psi_off[TTSK_MAX_ORDER], omega_off[TTSK_MAX_ORDER];
    int64_t total;
    int r1(int mu) const { return mu == 0 ? 1 : rL[mu - 1]; }
    int r2(int mu) const { return mu == d - 1 ? 1 : rR[mu]; }
};

static void make_layout(SketchLayout& L, int d, const int64_t* shape, const int32_t* rL, const int32_t* rR) {
    L.d = d;
    int64_t off = 0;
    for (int mu = 0; mu < d; mu++) L.shape[mu] = shape[mu];
    for (int mu = 0; mu < d - 1; mu++) { L.rL[mu] = rL[mu]; L.rR[mu] = rR[mu]; }
    for (int mu = 0; mu < d; mu++) {
        L.psi_off[mu] = off;
        off += (int64_t)L.r1(mu) * shape[mu] * L.r2(mu);
    }
    for (int mu = 0; mu < d - 1; mu++) {
        L.omega_off[mu] = off;
        off += (int64_t)rL[mu] * rR[mu];
    }
    L.total = off;
}

// Gaussian source for level `lvl` (0-based bond of the DRM's own orientation).
static void gauss_source(Source& S, const ttsk_drm& drm, int d, const int64_t* shape, int bond) {
    std::memset(&S, 0, sizeof(S));
    S.kind = SRC_GAUSS;
    S.r = drm.rank_max[bond] - drm.rank_min[bond];
    S.rank_min = drm.rank_min[bond];
    int64_t shp[TTSK_MAX_ORDER];
    if (!drm.right) {  // modes 0..bond, first fastest; seed_mu = mu + seed (sparse_gaussian_drm.py:32-36)
        S.k = bond + 1;
        for (int i = 0; i < S.k; i++) { S.modes[i] = i; shp[i] = shape[i]; }
        S.seed = (uint64_t)bond + drm.seed;
    } else {  // operates on tensor.T: modes d-1, d-2, ..., bond+1; loop index mu' = d-2-bond
        const int mup = d - 2 - bond;
        S.k = mup + 1;
        for (int i = 0; i < S.k; i++) { S.modes[i] = d - 1 - i; shp[i] = shape[d - 1 - i]; }
        S.seed = (uint64_t)mup + drm.seed;
    }
    wrapped_strides(shp, S.k, S.strides);
}

// number of distinct flat indices of a Gaussian source, or -1 if it overflows 2^31 (then the
// reference's int32 stride wraps and a table indexed by the true index is not equivalent)
static int64_t gauss_prefix_size(const Source& S, const int64_t* shape) {
    int64_t p = 1;
    for (int i = 0; i < S.k; i++) {
        p *= shape[S.modes[i]];
        if (p >= ((int64_t)1 << 31)) return -1;
    }
    return p;
}

struct SideState {
    Source src[TTSK_MAX_ORDER];        // per bond
    double* chain[TTSK_MAX_ORDER];     // TT: per-level chain buffers (chunk x true rank), DRM orientation
};

static int bucket_mode(ttsk_ctx* ctx, const long long* d_idx_mu, int64_t nnz, int64_t n_mu, int* d_hist,
                       int* d_offs, int* d_cursor, int* d_perm, int* d_skey, cudaStream_t st) {
    TTSK_CUDA(cudaMemsetAsync(d_hist, 0, (size_t)n_mu * sizeof(int), st));
    long long blocks = (nnz + 255) / 256;
    if (blocks > (long long)ctx->sm_count * 16) blocks = (long long)ctx->sm_count * 16;
    hist_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_idx_mu, nnz, d_hist);
    TTSK_LAUNCHED(ctx);
    scan_kernel<<<1, 1024, 0, st>>>(d_hist, n_mu, d_offs, d_cursor);
    TTSK_LAUNCHED(ctx);
    scatter_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_idx_mu, nnz, d_cursor, d_perm, d_skey);
    TTSK_LAUNCHED(ctx);
    return TTSK_OK;
}

static int ttdrm_step(ttsk_ctx* ctx, int64_t nnz, const long long* idx_mu, const double* v_in, int r_in,
                      const double* core, int64_t n, int r_out, double* v_out, cudaStream_t st) {
    if (nnz <= 0) return TTSK_OK;
    const long long total = (long long)nnz * r_out;
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)ctx->sm_count * 16) blocks = (long long)ctx->sm_count * 16;
    ttdrm_step_kernel<<<(unsigned)blocks, 256, 0, st>>>(nnz, idx_mu, v_in, r_in, core, n, r_out, v_out);
    TTSK_LAUNCHED(ctx);
    return TTSK_OK;
}

static int validate_drm(const ttsk_drm* drm, int d, int want_right) {
    TTSK_ARG(drm != nullptr, "DRM descriptor is NULL");
    TTSK_ARG(drm->kind == TTSK_DRM_GAUSS || drm->kind == TTSK_DRM_TT, "unknown DRM kind");
    TTSK_ARG((drm->right != 0) == (want_right != 0), "left/right DRM orientation mismatch");
    for (int mu = 0; mu < d - 1; mu++) {
        TTSK_ARG(drm->rank_min[mu] >= 0 && drm->rank_max[mu] > drm->rank_min[mu], "empty or negative rank slice");
        if (drm->kind == TTSK_DRM_TT) TTSK_ARG(drm->d_cores[mu] != nullptr, "TT DRM core pointer is NULL");
    }
    if (drm->kind == TTSK_DRM_TT) {
        TTSK_ARG(drm->core_r0[0] == 1, "first TT-DRM core must have left rank 1");
        for (int k = 0; k < d - 1; k++) {
            if (k > 0) TTSK_ARG(drm->core_r0[k] == drm->core_r1[k - 1], "TT-DRM core ranks do not chain");
            const int bond = drm->right ? d - 2 - k : k;
            TTSK_ARG(drm->rank_max[bond] <= drm->core_r1[k], "rank slice exceeds TT-DRM core rank");
        }
    }
    return TTSK_OK;
}

// Core of ttsk_sparse_sketch on device-resident COO data (one chunk = whole input here;
// the host entry point calls it per staged chunk).  `out` must be zero on entry for the
// Psi_0 / Psi_{d-1} blocks when edge==true is requested by the caller afterwards.
struct SparsePlan {
    SketchLayout lay;
    SideState left, right;
    int64_t n_max;
    bool use_table_L[TTSK_MAX_ORDER], use_table_R[TTSK_MAX_ORDER];
    double* table_L[TTSK_MAX_ORDER];
    double* table_R[TTSK_MAX_ORDER];
    double* edge_L0;   // (n_0, rL[0]) table for Omega_0 (Gaussian) or nullptr (TT: view of core)
    double* edge_R;    // (n_{d-1}, rR[d-2])
};

static int64_t chunk_bytes_per_nnz(int d, const ttsk_drm* left, const ttsk_drm* right) {
    int64_t b = 8;  // perm + skey
    if (left->kind == TTSK_DRM_TT)
        for (int k = 0; k < d - 1; k++) b += 8LL * left->core_r1[k];
    if (right->kind == TTSK_DRM_TT)
        for (int k = 0; k < d - 1; k++) b += 8LL * right->core_r1[k];
    return b;
}

static int sparse_chunk(ttsk_ctx* ctx, SparsePlan& pl, int d, const int64_t* shape, int64_t nnz,
                        const int64_t* d_idx, int64_t idx_row_stride, const double* d_val, const ttsk_drm* left,
                        const ttsk_drm* right, double* out, int* d_hist, int* d_offs, int* d_cursor, int* d_perm,
                        int* d_skey, cudaStream_t st) {
    if (nnz <= 0) return TTSK_OK;
    const SketchLayout& lay = pl.lay;
    // TT-DRM chains for this chunk (tensor_train_drm.py:60-69)
    if (left->kind == TTSK_DRM_TT)
        for (int k = 0; k < d - 1; k++)
            TTSK_TRY(ttdrm_step(ctx, nnz, (const long long*)(d_idx + k * idx_row_stride),
                                k == 0 ? nullptr : pl.left.chain[k - 1], left->core_r0[k], left->d_cores[k],
                                shape[k], left->core_r1[k], pl.left.chain[k], st));
    if (right->kind == TTSK_DRM_TT)
        for (int k = 0; k < d - 1; k++)
            TTSK_TRY(ttdrm_step(ctx, nnz, (const long long*)(d_idx + (d - 1 - k) * idx_row_stride),
                                k == 0 ? nullptr : pl.right.chain[k - 1], right->core_r0[k], right->d_cores[k],
                                shape[d - 1 - k], right->core_r1[k], pl.right.chain[k], st));
    for (int mu = 0; mu < d; mu++) {
        PassParams P;
        std::memset(&P, 0, sizeof(P));
        P.d = d;
        P.nnz = nnz;
        for (int m = 0; m < d; m++) P.idx[m] = (const long long*)(d_idx + m * idx_row_stride);
        P.val = d_val;
        P.n_mu = shape[mu];
        TTSK_TRY(bucket_mode(ctx, P.idx[mu], nnz, shape[mu], d_hist, d_offs, d_cursor, d_perm, d_skey, st));
        P.perm = d_perm;
        P.skey = d_skey;
        P.piece = 2048;
        P.A.kind = SRC_NONE; P.B.kind = SRC_NONE; P.X.kind = SRC_NONE;
        P.rA = lay.r1(mu);
        P.rB = lay.r2(mu);
        P.rX = 0;
        if (mu > 0) P.A = pl.left.src[mu - 1];
        if (mu < d - 1) P.B = pl.right.src[mu];
        const bool has_x = (mu >= 2 && mu <= d - 2);  // middle bond mu-1: Omega_{mu-1} += (v L_{mu-1})^T R_{mu-1}
        if (has_x) {
            P.X = pl.right.src[mu - 1];
            P.rX = lay.rR[mu - 1];
            P.omega = out + lay.omega_off[mu - 1];
        }
        P.psi = out + lay.psi_off[mu];
        if (ctx->timing) {
            while ((int)ctx->ev_pass.size() < 2 * (ctx->n_pass_events + 1)) {
                cudaEvent_t ev;
                TTSK_CUDA(cudaEventCreate(&ev));
                ctx->ev_pass.push_back(ev);
            }
            TTSK_CUDA(cudaEventRecord(ctx->ev_pass[2 * ctx->n_pass_events], st));
        }
        TTSK_TRY(launch_pass(ctx, P, has_x, st));
        if (ctx->timing) {
            TTSK_CUDA(cudaEventRecord(ctx->ev_pass[2 * ctx->n_pass_events + 1], st));
            ctx->n_pass_events++;
        }
    }
    return TTSK_OK;
}

// Build the per-bond sources (tables are generated here; chain buffers are carved per chunk)
static int build_plan(ttsk_ctx* ctx, SparsePlan& pl, int d, const int64_t* shape, int64_t nnz_total,
                      int64_t chunk, const ttsk_drm* left, const ttsk_drm* right, cudaStream_t st) {
    int32_t rL[TTSK_MAX_ORDER], rR[TTSK_MAX_ORDER];
    for (int mu = 0; mu < d - 1; mu++) {
        rL[mu] = left->rank_max[mu] - left->rank_min[mu];
        rR[mu] = right->rank_max[mu] - right->rank_min[mu];
    }
    make_layout(pl.lay, d, shape, rL, rR);
    pl.n_max = 0;
    for (int mu = 0; mu < d; mu++) pl.n_max = std::max<int64_t>(pl.n_max, shape[mu]);
    pl.edge_L0 = nullptr;
    pl.edge_R = nullptr;
    const int64_t table_rows_cap = std::max<int64_t>(nnz_total / 4, 1);
    const int64_t table_bytes_cap = (int64_t)4 << 30;
    for (int side = 0; side < 2; side++) {
        const ttsk_drm* drm = side == 0 ? left : right;
        SideState& ss = side == 0 ? pl.left : pl.right;
        for (int bond = 0; bond < d - 1; bond++) {
            Source& S = ss.src[bond];
            if (drm->kind == TTSK_DRM_GAUSS) {
                gauss_source(S, *drm, d, shape, bond);
                const int64_t rows = gauss_prefix_size(S, shape);
                const bool edge = (side == 0 && bond == 0) || (side == 1 && bond == d - 2);
                const bool small = rows > 0 && rows <= table_rows_cap && rows * S.r * 8 <= table_bytes_cap;
                if (small || (edge && rows > 0)) {
                    double* tab = (double*)ctx->ws_alloc(rows * (int64_t)S.r * 8);
                    if (!tab) { set_error("workspace too small for DRM table"); return TTSK_E_NOMEM; }
                    TTSK_TRY(gauss_table_launch(ctx, rows, S.rank_min, S.r, S.seed, tab, st));
                    if (edge && side == 0) pl.edge_L0 = tab;
                    if (edge && side == 1) pl.edge_R = tab;
                    if (small) {
                        // true (unwrapped) strides equal the wrapped ones because rows < 2^31
                        S.kind = SRC_TABLE;
                        S.base = tab;
                        S.row_stride = S.r;
                        S.col_stride = 1;
                    }
                }
            } else {
                std::memset(&S, 0, sizeof(S));
                S.kind = SRC_ROWS;
                S.r = drm->rank_max[bond] - drm->rank_min[bond];
                const int k = drm->right ? d - 2 - bond : bond;  // chain level in DRM orientation
                ss.chain[k] = (double*)ctx->ws_alloc(chunk * (int64_t)drm->core_r1[k] * 8);
                if (!ss.chain[k]) { set_error("workspace too small for TT-DRM chain"); return TTSK_E_NOMEM; }
                S.base = ss.chain[k] + drm->rank_min[bond];
                S.row_stride = drm->core_r1[k];
                S.col_stride = 1;
            }
        }
    }
    return TTSK_OK;
}

static int64_t plan_workspace_bytes(int d, const int64_t* shape, int64_t nnz_total, int64_t chunk,
                                    const ttsk_drm* left, const ttsk_drm* right, int64_t sketch_elems) {
    int64_t n_max = 0;
    for (int mu = 0; mu < d; mu++) n_max = std::max<int64_t>(n_max, shape[mu]);
    int64_t bytes = 0;
    auto add = [&](int64_t b) { bytes = align_up(bytes, 256) + b; };
    add(sketch_elems * 8);                  // temp sketch when accumulating
    add((n_max + 1) * 4); add((n_max + 1) * 4); add((n_max + 1) * 4);  // hist, offs, cursor
    add(chunk * 4); add(chunk * 4);         // perm, skey
    const int64_t table_rows_cap = std::max<int64_t>(nnz_total / 4, 1);
    for (int side = 0; side < 2; side++) {
        const ttsk_drm* drm = side == 0 ? left : right;
        for (int bond = 0; bond < d - 1; bond++) {
            const int r = drm->rank_max[bond] - drm->rank_min[bond];
            if (drm->kind == TTSK_DRM_GAUSS) {
                int64_t rows = 1;
                bool ok = true;
                if (!drm->right) { for (int i = 0; i <= bond; i++) { rows *= shape[i]; if (rows >= ((int64_t)1 << 31)) { ok = false; break; } } }
                else { for (int i = d - 1; i > bond; i--) { rows *= shape[i]; if (rows >= ((int64_t)1 << 31)) { ok = false; break; } } }
                const bool edge = (side == 0 && bond == 0) || (side == 1 && bond == d - 2);
                if (ok && ((rows <= table_rows_cap && rows * r * 8 <= ((int64_t)4 << 30)) || edge)) add(rows * r * 8);
            } else {
                const int k = drm->right ? d - 2 - bond : bond;
                add(chunk * (int64_t)drm->core_r1[k] * 8);
            }
        }
    }
    return bytes + 4096;
}

// edge bonds from the (this-call-only) Psi_0 / Psi_{d-1}
static int edge_omegas(ttsk_ctx* ctx, const SparsePlan& pl, int d, const int64_t* shape, const ttsk_drm* left,
                       const ttsk_drm* right, double* sk, cudaStream_t st) {
    const SketchLayout& lay = pl.lay;
    {   // Omega_0[a, b] = sum_j L_0[j, a] Psi_0[0, j, b]
        const double* Ltab; int64_t l_rs;
        if (left->kind == TTSK_DRM_GAUSS) { Ltab = pl.edge_L0; l_rs = lay.rL[0]; }
        else { Ltab = left->d_cores[0] + left->rank_min[0]; l_rs = left->core_r1[0]; }
        TTSK_TRY(gemm_launch(ctx, lay.rL[0], lay.rR[0], shape[0], 1.0, Ltab, 1, l_rs, sk + lay.psi_off[0],
                             lay.rR[0], 1, 1.0, sk + lay.omega_off[0], lay.rR[0], 1, 1, 0, 0, 0, st));
    }
    if (d >= 3) {  // Omega_{d-2}[a, b] = sum_j Psi_{d-1}[a, j, 0] R_{d-2}[j, b]
        const int bond = d - 2;
        const double* Rtab; int64_t r_rs;
        if (right->kind == TTSK_DRM_GAUSS) { Rtab = pl.edge_R; r_rs = lay.rR[bond]; }
        else { Rtab = right->d_cores[0] + right->rank_min[bond]; r_rs = right->core_r1[0]; }
        TTSK_TRY(gemm_launch(ctx, lay.rL[bond], lay.rR[bond], shape[d - 1], 1.0, sk + lay.psi_off[d - 1],
                             shape[d - 1], 1, Rtab, r_rs, 1, 1.0, sk + lay.omega_off[bond], lay.rR[bond], 1, 1, 0,
                             0, 0, st));
    }
    return TTSK_OK;
}

static int validate_common(ttsk_ctx* ctx, int d, const int64_t* h_shape, int64_t nnz, const ttsk_drm* left,
                           const ttsk_drm* right) {
    TTSK_ARG(ctx != nullptr, "ctx is NULL");
    TTSK_ARG(d >= 2 && d <= TTSK_MAX_ORDER, "tensor order must be in [2, 16]");
    TTSK_ARG(h_shape != nullptr && nnz >= 0, "shape/nnz");
    for (int m = 0; m < d; m++) TTSK_ARG(h_shape[m] >= 1 && h_shape[m] < ((int64_t)1 << 31), "mode size out of range");
    TTSK_TRY(validate_drm(left, d, 0));
    TTSK_TRY(validate_drm(right, d, 1));
    return TTSK_OK;
}

static int64_t pick_chunk(int d, int64_t nnz, const ttsk_drm* left, const ttsk_drm* right, int64_t budget_bytes) {
    const int64_t per = chunk_bytes_per_nnz(d, left, right);
    int64_t chunk = budget_bytes / per;
    if (chunk > nnz) chunk = nnz;
    if (chunk > ((int64_t)1 << 30)) chunk = (int64_t)1 << 30;
    if (chunk < 1) chunk = 1;
    return chunk;
}

}  // namespace ttsk

using namespace ttsk;

extern "C" int64_t ttsk_sketch_size(int d, const int64_t* h_shape, const int32_t* rL, const int32_t* rR) {
    if (d < 1 || d > TTSK_MAX_ORDER || !h_shape) return -1;
    SketchLayout L;
    make_layout(L, d, h_shape, rL, rR);
    return L.total;
}

extern "C" int ttsk_sparse_sketch(ttsk_ctx* ctx, int d, const int64_t* h_shape, int64_t nnz, const int64_t* d_idx,
                                  int64_t idx_row_stride, const double* d_val, const ttsk_drm* left,
                                  const ttsk_drm* right, double* d_out, int accumulate, void* stream) {
    TTSK_TRY(validate_common(ctx, d, h_shape, nnz, left, right));
    TTSK_ARG(d_out != nullptr && (nnz == 0 || (d_idx && d_val)), "NULL device pointer");
    cudaStream_t st = (cudaStream_t)stream;
    TTSK_CUDA(cudaSetDevice(ctx->device));
    SparsePlan pl;
    int32_t rL[TTSK_MAX_ORDER], rR[TTSK_MAX_ORDER];
    for (int mu = 0; mu < d - 1; mu++) {
        rL[mu] = left->rank_max[mu] - left->rank_min[mu];
        rR[mu] = right->rank_max[mu] - right->rank_min[mu];
    }
    const int64_t total = ttsk_sketch_size(d, h_shape, rL, rR);
    const int64_t chunk = pick_chunk(d, std::max<int64_t>(nnz, 1), left, right, (int64_t)24 << 30);
    TTSK_TRY(ctx->ws_reserve(plan_workspace_bytes(d, h_shape, nnz, chunk, left, right, total)));
    ctx->ws_reset();
    ctx->n_pass_events = 0;
    double* tmp = (double*)ctx->ws_alloc(total * 8);
    int64_t n_max = 0;
    for (int mu = 0; mu < d; mu++) n_max = std::max<int64_t>(n_max, h_shape[mu]);
    int* d_hist = (int*)ctx->ws_alloc((n_max + 1) * 4);
    int* d_offs = (int*)ctx->ws_alloc((n_max + 1) * 4);
    int* d_cursor = (int*)ctx->ws_alloc((n_max + 1) * 4);
    int* d_perm = (int*)ctx->ws_alloc(chunk * 4);
    int* d_skey = (int*)ctx->ws_alloc(chunk * 4);
    if (!tmp || !d_hist || !d_offs || !d_cursor || !d_perm || !d_skey) {
        set_error("workspace carve failed");
        return TTSK_E_NOMEM;
    }
    if (ctx->timing) TTSK_CUDA(cudaEventRecord(ctx->ev_t0, st));
    TTSK_TRY(build_plan(ctx, pl, d, h_shape, nnz, chunk, left, right, st));
    double* sk = accumulate ? tmp : d_out;
    TTSK_CUDA(cudaMemsetAsync(sk, 0, (size_t)total * 8, st));
    for (int64_t c0 = 0; c0 < nnz; c0 += chunk) {
        const int64_t n = std::min<int64_t>(chunk, nnz - c0);
        TTSK_TRY(sparse_chunk(ctx, pl, d, h_shape, n, d_idx + c0, idx_row_stride, d_val + c0, left, right, sk, d_hist,
                              d_offs, d_cursor, d_perm, d_skey, st));
    }
    TTSK_TRY(edge_omegas(ctx, pl, d, h_shape, left, right, sk, st));
    if (accumulate) TTSK_TRY(axpy_launch(ctx, total, 1.0, tmp, d_out, st));
    if (ctx->timing) TTSK_CUDA(cudaEventRecord(ctx->ev_t1, st));
    return TTSK_OK;
}

extern "C" int ttsk_last_kernel_ms(ttsk_ctx* ctx, double* ms_total, double* ms_dominant) {
    TTSK_ARG(ctx != nullptr, "ctx is NULL");
    TTSK_CUDA(cudaEventSynchronize(ctx->ev_t1));
    float t = 0.f;
    TTSK_CUDA(cudaEventElapsedTime(&t, ctx->ev_t0, ctx->ev_t1));
    if (ms_total) *ms_total = t;
    double sum = 0.0;
    for (int i = 0; i < ctx->n_pass_events; i++) {
        float p = 0.f;
        TTSK_CUDA(cudaEventElapsedTime(&p, ctx->ev_pass[2 * i], ctx->ev_pass[2 * i + 1]));
        sum += p;
    }
    if (ms_dominant) *ms_dominant = sum;
    return TTSK_OK;
}

// Host-buffer entry points: the COO arrays live in (ideally pinned) host memory; chunks are
// copied host->device on the context's copy stream into two staging buffers while the previous
// chunk is sketched on the compute stream.  The packed sketch ends in d_out (device, for a
// following all-reduce) and/or h_out (host).
static int sparse_sketch_from_host(ttsk_ctx* ctx, int d, const int64_t* h_shape, int64_t nnz, const int64_t* h_idx,
                                   int64_t idx_row_stride, const double* h_val, const ttsk_drm* left,
                                   const ttsk_drm* right, double* d_out, double* h_out, int accumulate) {
    TTSK_TRY(validate_common(ctx, d, h_shape, nnz, left, right));
    TTSK_ARG((h_out != nullptr || d_out != nullptr) && (nnz == 0 || (h_idx && h_val)), "NULL pointer");
    TTSK_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->compute_stream, cs = ctx->copy_stream;
    int32_t rL[TTSK_MAX_ORDER], rR[TTSK_MAX_ORDER];
    for (int mu = 0; mu < d - 1; mu++) {
        rL[mu] = left->rank_max[mu] - left->rank_min[mu];
        rR[mu] = right->rank_max[mu] - right->rank_min[mu];
    }
    const int64_t total = ttsk_sketch_size(d, h_shape, rL, rR);
    // chunk: bounded by the chain workspace and by a staging size that overlaps well
    int64_t chunk = pick_chunk(d, std::max<int64_t>(nnz, 1), left, right, (int64_t)8 << 30);
    const int64_t stage_cap = (int64_t)1 << 23;  // 8M nonzeros = 320 MB per staging buffer at d=4
    if (chunk > stage_cap) chunk = stage_cap;
    const int64_t rec = (int64_t)(d + 1) * 8;
    const int64_t stage_bytes = align_up(chunk * rec, 256);
    const int64_t plan_bytes = plan_workspace_bytes(d, h_shape, nnz, chunk, left, right, total);
    TTSK_TRY(ctx->ws_reserve(plan_bytes + 2 * stage_bytes + 1024));
    ctx->ws_reset();
    ctx->n_pass_events = 0;
    char* d_stage[2];
    d_stage[0] = (char*)ctx->ws_alloc(stage_bytes);
    d_stage[1] = (char*)ctx->ws_alloc(stage_bytes);
    double* sk = (double*)ctx->ws_alloc(total * 8);
    int64_t n_max = 0;
    for (int mu = 0; mu < d; mu++) n_max = std::max<int64_t>(n_max, h_shape[mu]);
    int* d_hist = (int*)ctx->ws_alloc((n_max + 1) * 4);
    int* d_offs = (int*)ctx->ws_alloc((n_max + 1) * 4);
    int* d_cursor = (int*)ctx->ws_alloc((n_max + 1) * 4);
    int* d_perm = (int*)ctx->ws_alloc(chunk * 4);
    int* d_skey = (int*)ctx->ws_alloc(chunk * 4);
    if (!d_stage[0] || !d_stage[1] || !sk || !d_hist || !d_offs || !d_cursor || !d_perm || !d_skey) {
        set_error("workspace carve failed");
        return TTSK_E_NOMEM;
    }
    SparsePlan pl;
    if (ctx->timing) TTSK_CUDA(cudaEventRecord(ctx->ev_t0, st));
    TTSK_TRY(build_plan(ctx, pl, d, h_shape, nnz, chunk, left, right, st));
    TTSK_CUDA(cudaMemsetAsync(sk, 0, (size_t)total * 8, st));
    int buf = 0;
    for (int64_t c0 = 0; c0 < nnz; c0 += chunk, buf ^= 1) {
        const int64_t n = std::min<int64_t>(chunk, nnz - c0);
        // the staging buffer may be overwritten once the kernels that read it (two chunks ago) are done
        TTSK_CUDA(cudaStreamWaitEvent(cs, ctx->ev_done[buf], 0));
        long long* di = (long long*)d_stage[buf];
        double* dv = (double*)(d_stage[buf] + (size_t)d * chunk * 8);
        for (int m = 0; m < d; m++)
            TTSK_CUDA(cudaMemcpyAsync(di + (size_t)m * chunk, h_idx + m * idx_row_stride + c0, (size_t)n * 8,
                                      cudaMemcpyHostToDevice, cs));
        TTSK_CUDA(cudaMemcpyAsync(dv, h_val + c0, (size_t)n * 8, cudaMemcpyHostToDevice, cs));
        TTSK_CUDA(cudaEventRecord(ctx->ev_copy[buf], cs));
        TTSK_CUDA(cudaStreamWaitEvent(st, ctx->ev_copy[buf], 0));
        TTSK_TRY(sparse_chunk(ctx, pl, d, h_shape, n, (const int64_t*)di, chunk, dv, left, right, sk, d_hist, d_offs,
                              d_cursor, d_perm, d_skey, st));
        TTSK_CUDA(cudaEventRecord(ctx->ev_done[buf], st));
    }
    TTSK_TRY(edge_omegas(ctx, pl, d, h_shape, left, right, sk, st));
    if (ctx->timing) TTSK_CUDA(cudaEventRecord(ctx->ev_t1, st));
    if (d_out) {
        if (accumulate) TTSK_TRY(axpy_launch(ctx, total, 1.0, sk, d_out, st));
        else TTSK_CUDA(cudaMemcpyAsync(d_out, sk, (size_t)total * 8, cudaMemcpyDeviceToDevice, st));
    }
    if (h_out) {
        if (!accumulate) {
            TTSK_CUDA(cudaMemcpyAsync(h_out, sk, (size_t)total * 8, cudaMemcpyDeviceToHost, st));
            TTSK_CUDA(cudaStreamSynchronize(st));
        } else {
            std::vector<double> part((size_t)total);
            TTSK_CUDA(cudaMemcpyAsync(part.data(), sk, (size_t)total * 8, cudaMemcpyDeviceToHost, st));
            TTSK_CUDA(cudaStreamSynchronize(st));
            for (int64_t i = 0; i < total; i++) h_out[i] += part[(size_t)i];
        }
    }
    TTSK_CUDA(cudaStreamSynchronize(st));
    TTSK_CUDA(cudaStreamSynchronize(cs));
    return TTSK_OK;
}

extern "C" int ttsk_sparse_sketch_host(ttsk_ctx* ctx, int d, const int64_t* h_shape, int64_t nnz,
                                       const int64_t* h_idx, int64_t idx_row_stride, const double* h_val,
                                       const ttsk_drm* left, const ttsk_drm* right, double* h_out, int accumulate) {
    TTSK_ARG(h_out != nullptr, "h_out is NULL");
    return sparse_sketch_from_host(ctx, d, h_shape, nnz, h_idx, idx_row_stride, h_val, left, right, nullptr, h_out,
                                   accumulate);
}

extern "C" int ttsk_sparse_sketch_stream(ttsk_ctx* ctx, int d, const int64_t* h_shape, int64_t nnz,
                                         const int64_t* h_idx, int64_t idx_row_stride, const double* h_val,
                                         const ttsk_drm* left, const ttsk_drm* right, double* d_out, int accumulate) {
    TTSK_ARG(d_out != nullptr, "d_out is NULL");
    return sparse_sketch_from_host(ctx, d, h_shape, nnz, h_idx, idx_row_stride, h_val, left, right, d_out, nullptr,
                                   accumulate);
}

// ------------------------------------------------------------------ operator-level entry points
extern "C" int ttsk_ttdrm_sparse_step(ttsk_ctx* ctx, int64_t nnz, const int64_t* d_idx_mu, const double* d_v_in,
                                      int r_in, const double* d_core, int64_t n, int r_out, double* d_v_out,
                                      void* stream) {
    TTSK_ARG(ctx != nullptr, "ctx is NULL");
    TTSK_ARG(nnz >= 0 && r_out >= 1 && n >= 1 && (d_v_in == nullptr || r_in >= 1), "ttdrm step dims");
    TTSK_ARG(nnz == 0 || (d_idx_mu && d_core && d_v_out), "NULL pointer");
    return ttdrm_step(ctx, nnz, (const long long*)d_idx_mu, d_v_in, r_in, d_core, n, r_out, d_v_out,
                      (cudaStream_t)stream);
}

static void rows_source(Source& S, const double* base, int r, int64_t ps, int64_t cs) {
    std::memset(&S, 0, sizeof(S));
    if (base == nullptr) { S.kind = SRC_NONE; S.r = 1; return; }
    S.kind = SRC_ROWS;
    S.r = r;
    S.base = base;  // element (p, a) at base[p*ps + a*cs]
    S.row_stride = ps;
    S.col_stride = cs;
}

extern "C" int ttsk_sparse_omega(ttsk_ctx* ctx, int64_t nnz, const double* d_val, const double* d_left, int rL,
                                 int64_t l_ps, int64_t l_cs, const double* d_right, int rR, int64_t r_ps,
                                 int64_t r_cs, double* d_omega, void* stream) {
    TTSK_ARG(ctx != nullptr, "ctx is NULL");
    TTSK_ARG(nnz >= 0 && rL >= 1 && rR >= 1 && nnz < ((int64_t)1 << 31), "omega dims");
    TTSK_ARG(nnz == 0 || (d_val && d_left && d_right && d_omega), "NULL pointer");
    PassParams P;
    std::memset(&P, 0, sizeof(P));
    P.d = 0; P.nnz = nnz; P.val = d_val; P.n_mu = 1; P.piece = 2048;
    rows_source(P.A, d_left, rL, l_ps, l_cs);
    rows_source(P.B, d_right, rR, r_ps, r_cs);
    P.X.kind = SRC_NONE;
    P.rA = rL; P.rB = rR; P.rX = 0;
    P.psi = d_omega;  // a Psi with a single slice IS Omega
    return launch_pass(ctx, P, false, (cudaStream_t)stream);
}

extern "C" int ttsk_sparse_psi(ttsk_ctx* ctx, int64_t nnz, const int64_t* d_idx_mu, int64_t n_mu,
                               const double* d_val, const double* d_left, int rL, int64_t l_ps, int64_t l_cs,
                               const double* d_right, int rR, int64_t r_ps, int64_t r_cs, double* d_psi,
                               void* stream) {
    TTSK_ARG(ctx != nullptr, "ctx is NULL");
    TTSK_ARG(nnz >= 0 && n_mu >= 1 && nnz < ((int64_t)1 << 31) && n_mu < ((int64_t)1 << 31), "psi dims");
    TTSK_ARG(d_left != nullptr || d_right != nullptr, "sketch_psi_sparse needs at least one side (sparse_sketch.py:21-32)");
    TTSK_ARG(nnz == 0 || (d_val && d_idx_mu && d_psi), "NULL pointer");
    if (nnz == 0) return TTSK_OK;
    cudaStream_t st = (cudaStream_t)stream;
    TTSK_TRY(ctx->ws_reserve(3 * (n_mu + 1) * 4 + 2 * nnz * 4 + 4096));
    ctx->ws_reset();
    int* d_hist = (int*)ctx->ws_alloc((n_mu + 1) * 4);
    int* d_offs = (int*)ctx->ws_alloc((n_mu + 1) * 4);
    int* d_cursor = (int*)ctx->ws_alloc((n_mu + 1) * 4);
    int* d_perm = (int*)ctx->ws_alloc(nnz * 4);
    int* d_skey = (int*)ctx->ws_alloc(nnz * 4);
    TTSK_TRY(bucket_mode(ctx, (const long long*)d_idx_mu, nnz, n_mu, d_hist, d_offs, d_cursor, d_perm, d_skey, st));
    PassParams P;
    std::memset(&P, 0, sizeof(P));
    P.d = 0; P.nnz = nnz; P.val = d_val; P.n_mu = n_mu; P.piece = 2048;
    P.perm = d_perm; P.skey = d_skey;
    rows_source(P.A, d_left, rL, l_ps, l_cs);
    rows_source(P.B, d_right, rR, r_ps, r_cs);
    P.X.kind = SRC_NONE;
    P.rA = d_left ? rL : 1; P.rB = d_right ? rR : 1; P.rX = 0;
    P.psi = d_psi;
    return launch_pass(ctx, P, false, st);
}
