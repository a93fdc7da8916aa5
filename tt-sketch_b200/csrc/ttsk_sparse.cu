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
//     per slice j of every mode (O(n_mu * nnz)).  Here the chunk is packed into 32-byte records,
//     bucketed once per mode (counting sort -> sorted (key, id) words) and a mode pass walks the
//     buckets tile by tile, forming
//         At (rA x TN) = v_p * L_{mu-1}(p),  Bt (rB x TN) = R_mu(p),  Xt (rX x TN) = R_{mu-1}(p)
//     in shared memory and accumulating  Psi_mu[:, j, :] += At Bt^T  (and Omega_{mu-1} += At Xt^T)
//     in registers with mma.sync.m8n8k4.f64, flushing a slice when the key changes.
//   * the pass kernels (ttsk_sparse_pass.cuh, instantiated in ttsk_sparse_pass_x0/x1.cu) are
//     warp-specialised and persistent: producer warps stage tiles and launch asynchronous row
//     copies into an mbarrier ring, consumer warps generate / multiply their own rows without CTA
//     barriers; a segment-GEMM form and an unbucketed last-mode form replace the per-nonzero MMAs
//     when the right factor is a small table or absent.
//   * tile entries come from a "source": hash-seeded Gaussian generated on the fly (never
//     stored in HBM), a prefix table of the same generator (or the first core of a TT DRM) gathered
//     through L2 when the prefix index space is not larger than nnz, or rows of a per-nonzero
//     buffer (TT-DRM chain products / user-supplied rows).
//   * the tail branch of ndtri (27% of draws) is deferred to a dense second phase so the
//     central branch runs without the tail's divergence.
//   * edge bonds need no per-nonzero work:  Omega_0 = L_0^T Psi_0,  Omega_{d-2} = Psi_{d-1} R_{d-2}^T.
//
// This file: bucketing kernels, TT-DRM chain kernels, host-side planning, the C entry points.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "ttsk_common.cuh"
#include "ttsk_gauss.cuh"

namespace ttsk {

int gauss_table_launch(ttsk_ctx* ctx, int64_t rows, int rank_min, int rank, uint64_t seed, double* d_out,
                       cudaStream_t st);
void wrapped_strides(const int64_t* shape, int k, long long* strides);

}  // namespace ttsk
#include "ttsk_sparse_pass.cuh"
namespace ttsk {

// ------------------------------------------------------------------ bucketing (counting sort)
__global__ void hist_kernel(const long long* __restrict__ idx, long long nnz, int* __restrict__ hist) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < nnz;
         p += (long long)gridDim.x * blockDim.x)
        atomicAdd(&hist[idx[p]], 1);
}

// shared-memory privatised histogram: each CTA counts a contiguous block of nonzeros in shared
// memory and adds only its non-empty bins to the global histogram (n_mu <= kLocalBins)
constexpr int kLocalBins = 26 * 1024;  // two int arrays of this size fit the 227 KB of one SM
__global__ void __launch_bounds__(1024) hist_local_kernel(const long long* __restrict__ idx, long long nnz,
                                                          long long block_len, int n_mu, int* __restrict__ hist,
                                                          int* __restrict__ cta_cnt) {
    extern __shared__ int s_cnt[];
    for (int k = threadIdx.x; k < n_mu; k += blockDim.x) s_cnt[k] = 0;
    __syncthreads();
    const long long b0 = (long long)blockIdx.x * block_len;
    const long long b1 = (b0 + block_len < nnz) ? b0 + block_len : nnz;
    for (long long p = b0 + threadIdx.x; p < b1; p += blockDim.x) atomicAdd(&s_cnt[idx[p]], 1);
    __syncthreads();
    for (int k = threadIdx.x; k < n_mu; k += blockDim.x) {
        const int c = s_cnt[k];
        if (c) atomicAdd(&hist[k], c);
        cta_cnt[(long long)blockIdx.x * n_mu + k] = c;  // kept for the scatter: no second counting pass
    }
}

// per-CTA counts -> per-CTA start of every key's range: base[c][k] = offs[k] + sum_{c' < c} cnt[c'][k] (in place)
__global__ void cta_base_kernel(int* __restrict__ cta_cnt, const int* __restrict__ offs, int n_mu, int n_cta) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_mu) return;
    int run = offs[k];
    for (int c = 0; c < n_cta; c++) {
        const long long at = (long long)c * n_mu + k;
        const int t = cta_cnt[at];
        cta_cnt[at] = run;
        run += t;
    }
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


// Pack the chunk's COO arrays (SoA, int64 indices) into one sector-sized record per nonzero so a
// mode pass fetches a nonzero with ONE random 32-byte access instead of d+1 of them.
__global__ void pack_records_kernel(int d, long long nnz, ScatterIdx rows, const double* __restrict__ val,
                                    unsigned* __restrict__ recs, int rec_words) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < nnz;
         p += (long long)gridDim.x * blockDim.x) {
        unsigned* r = recs + p * rec_words;
        const unsigned long long v = (unsigned long long)__double_as_longlong(val[p]);
        unsigned w[8];
        w[0] = (unsigned)v; w[1] = (unsigned)(v >> 32);
        for (int base = 0; base < rec_words; base += 8) {
#pragma unroll
            for (int j = (base == 0 ? 2 : 0); j < 8; j++) {
                const int m = base + j - 2;
                w[j] = (m < d) ? (unsigned)rows.idx[m][p] : 0u;
            }
            reinterpret_cast<uint4*>(r + base)[0] = make_uint4(w[0], w[1], w[2], w[3]);
            reinterpret_cast<uint4*>(r + base)[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
    }
}

__global__ void scatter_kernel(const ScatterParams S) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < S.nnz;
         p += (long long)gridDim.x * blockDim.x) {
        const int key = (int)S.key_idx[p];
        const int pos = atomicAdd(&S.cursor[key], 1);
        S.keyid[pos] = ((unsigned long long)(unsigned)key << 32) | (unsigned long long)(unsigned)p;
    }
}

// Two-level scatter for n_mu <= kLocalBins: the start of every (CTA, key) range was derived from the
// histogram kernel's per-CTA counts (deterministic bucket layout, no global atomics); a CTA only
// ranks the nonzeros of its contiguous block within their key with shared-memory atomics.
// The words of one (CTA, key) range are written 8 bytes at a time over the whole lifetime of the CTA; with
// grid * n_mu open ranges the partially written sectors do not fit L2 and go to DRAM half empty.  The block is
// therefore scattered in `rounds` sweeps over disjoint key ranges (re-reading the keys, which stream), so only
// grid * n_mu / rounds ranges are open at a time.
__global__ void __launch_bounds__(1024) scatter_local_kernel(const ScatterParams S, long long block_len, int n_mu,
                                                             const int* __restrict__ cta_base, int rounds) {
    extern __shared__ int s_bins[];
    int* s_cnt = s_bins;
    int* s_base = s_bins + n_mu;
    for (int k = threadIdx.x; k < n_mu; k += blockDim.x) {
        s_cnt[k] = 0;
        s_base[k] = cta_base[(long long)blockIdx.x * n_mu + k];
    }
    __syncthreads();
    const long long b0 = (long long)blockIdx.x * block_len;
    const long long b1 = (b0 + block_len < S.nnz) ? b0 + block_len : S.nnz;
    const int span = (n_mu + rounds - 1) / rounds;
    for (int r = 0; r < rounds; r++) {
        const unsigned klo = (unsigned)(r * span);
        for (long long p = b0 + threadIdx.x; p < b1; p += blockDim.x) {
            const int key = (int)__ldcs(S.key_idx + p);
            if ((unsigned)key - klo >= (unsigned)span) continue;
            const int pos = s_base[key] + atomicAdd(&s_cnt[key], 1);
            S.keyid[pos] = ((unsigned long long)(unsigned)key << 32) | (unsigned long long)(unsigned)p;
        }
    }
}

// ------------------------------------------------------------------ two-level partition (n_mu <= 16384)
// Bucketing as two partition sweeps: first by key >> 7 (at most 128 coarse buckets), then, tile by tile inside a coarse
// bucket, by the full key.  A tile of 4096 elements is ranked per digit with shared-memory counters, reserves its
// space in every destination bucket with one global atomic per non-empty digit, and writes runs of consecutive
// elements (about 50 words per coarse bucket in the first sweep, 32 per key in the second) -- against single
// 8-byte words into 10^4 x CTAs open ranges in the one-sweep scatter above.  Traffic: keys read twice (histogram,
// sweep 1), words written and read once more: 4 GB at nnz = 1e8.
constexpr int kPartTile = 4096, kPartThreads = 512, kPartBins = 128, kPartPer = kPartTile / kPartThreads;

// the tile map of the second sweep: coarse bucket b owns keys [128 b, 128 b + 128)
__global__ void __launch_bounds__(kPartBins) part_setup_kernel(const int* __restrict__ offs, int n_mu, int* __restrict__ tstart) {
    __shared__ int s_t[kPartBins];
    const int b = threadIdx.x;
    const int k0 = min(b * kPartBins, n_mu), k1 = min((b + 1) * kPartBins, n_mu);
    const int lo = offs[k0], hi = offs[k1];
    s_t[b] = (hi - lo + kPartTile - 1) / kPartTile;
    __syncthreads();
    if (b == 0) {
        int run = 0;
        for (int i = 0; i < kPartBins; i++) { const int t = s_t[i]; tstart[i] = run; run += t; }
        tstart[kPartBins] = run;
    }
}

// start of (CTA c, coarse bucket b) in the first sweep's output, from the per-(CTA, key) range starts of the final
// layout (cta_base_kernel): bucket b begins at offs[128 b] and CTA c's share of it follows the shares of the CTAs
// before it.  With these the first sweep needs no global atomics (79 hot cursors hit by every tile cost 1 ms).
__global__ void __launch_bounds__(kPartBins) part_l1base_kernel(const int* __restrict__ cta_base, const int* __restrict__ offs,
                                                                int n_mu, int* __restrict__ l1base) {
    const int c = blockIdx.x, b = threadIdx.x;
    const int k0 = min(b * kPartBins, n_mu), k1 = min((b + 1) * kPartBins, n_mu);
    int before = 0;
    for (int k = k0; k < k1; k++) before += cta_base[(long long)c * n_mu + k] - offs[k];
    l1base[c * kPartBins + b] = offs[k0] + before;
}

// LEVEL 1: CTA c walks ITS block of the nonzeros (the block the histogram kernel counted for it) in input order,
//          digit = key >> 7, destination = the CTA's own cursor of the coarse bucket (shared memory, no atomics).
// LEVEL 2: elements are the words of the first sweep, walked per coarse bucket (tile map `tstart`), digit = key & 127,
//          destination cursor cursor[key].
//
// PAY (payload partition, for passes that only gather table rows): the element is a PAIR of words -- w = key << kshift |
// row of table A << bshift | row of table B, and the nonzero's value -- formed in the first sweep from the packed
// records (read in input order, i.e. coalesced) instead of (key << 32 | id): the pass then streams everything it
// needs per nonzero in sorted order and never touches the records (one random 32-byte read per nonzero, a 128-byte
// DRAM access, in the id form).

template <int LEVEL, bool PAY>
__global__ void __launch_bounds__(kPartThreads) partition_kernel(const long long* __restrict__ keys,
                                                                const unsigned long long* __restrict__ in_words, long long n,
                                                                int n_mu, const int* __restrict__ offs,
                                                                const int* __restrict__ tstart, int* __restrict__ cursors,
                                                                unsigned long long* __restrict__ out_words, long long block_len,
                                                                int kshift, const PayPlan pp) {
    // every warp ranks its elements in its OWN row of counters (one address hit by all 16 warps serialises at the
    // bank: the first sweep has only ~80 live digits); a prefix over the warps then places the rows of a digit
    __shared__ int s_wcnt[kPartThreads / 32][kPartBins];
    __shared__ int s_base[kPartBins];
    __shared__ int s_cur[kPartBins];
    __shared__ int s_tot[kPartBins];
    __shared__ int s_loc[kPartBins];
    __shared__ unsigned long long s_out[kPartTile];
    extern __shared__ unsigned long long s_pay[];  // PAY: the values' tile buffer (kPartTile words of dynamic shared memory)
    __shared__ long long s_range[2];
    __shared__ int s_bucket;
    const int tid = threadIdx.x, warp = tid >> 5;
    // LEVEL 1: tiles of this CTA's block; LEVEL 2: tiles of the whole array, strided over the grid
    const long long blk_lo = (long long)blockIdx.x * block_len, blk_hi = (blk_lo + block_len < n) ? blk_lo + block_len : n;
    const long long n_tiles = (LEVEL == 1) ? (blk_hi > blk_lo ? (blk_hi - blk_lo + kPartTile - 1) / kPartTile : 0) : (long long)tstart[kPartBins];
    if (LEVEL == 1 && tid < kPartBins) s_cur[tid] = cursors[blockIdx.x * kPartBins + tid];
    for (long long t = (LEVEL == 1) ? 0 : blockIdx.x; t < n_tiles; t += (LEVEL == 1) ? 1 : gridDim.x) {
        for (int i = tid; i < (kPartThreads / 32) * kPartBins; i += kPartThreads) (&s_wcnt[0][0])[i] = 0;
        if (tid == 0) {
            if (LEVEL == 1) {
                s_range[0] = blk_lo + t * kPartTile;
                s_range[1] = (s_range[0] + kPartTile < blk_hi) ? s_range[0] + kPartTile : blk_hi;
                s_bucket = 0;
            } else {
                int lo = 0, hi = kPartBins - 1;  // last bucket b with tstart[b] <= t
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if ((long long)tstart[mid] <= t) lo = mid; else hi = mid - 1;
                }
                const int k0 = min(lo * kPartBins, n_mu), k1 = min((lo + 1) * kPartBins, n_mu);
                const long long b_lo = offs[k0], b_hi = offs[k1];
                s_range[0] = b_lo + (t - tstart[lo]) * kPartTile;
                s_range[1] = (s_range[0] + kPartTile < b_hi) ? s_range[0] + kPartTile : b_hi;
                s_bucket = lo;
            }
        }
        __syncthreads();
        const long long lo = s_range[0], hi = s_range[1];
        unsigned long long w[kPartPer], pay[PAY ? kPartPer : 1];
        int rank[kPartPer];
#pragma unroll
        for (int u = 0; u < kPartPer; u++) {
            const long long p = lo + tid + (long long)u * kPartThreads;
            if (p < hi) {
                if (LEVEL == 1 && PAY) {
                    const uint4 r0 = __ldg(reinterpret_cast<const uint4*>(pp.recs + p * 8));
                    const uint4 r1 = __ldg(reinterpret_cast<const uint4*>(pp.recs + p * 8) + 1);
                    const unsigned wd[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
                    unsigned long long a = 0, b = 0;
#pragma unroll
                    for (int m = 0; m < 6; m++) {
                        a += (unsigned long long)wd[2 + m] * pp.smul_a[m];
                        b += (unsigned long long)wd[2 + m] * pp.smul_b[m];
                    }
                    unsigned key = 0;
#pragma unroll
                    for (int m = 2; m < 8; m++) key = (m == pp.key_word) ? wd[m] : key;
                    w[u] = ((unsigned long long)key << kshift) | (a << pp.bshift) | b;
                    pay[u] = ((unsigned long long)r0.y << 32) | r0.x;
                } else if (LEVEL == 1) {
                    w[u] = ((unsigned long long)(unsigned)__ldcs(keys + p) << 32) | (unsigned long long)(unsigned)p;
                } else {
                    w[u] = __ldcs(in_words + p);
                    if (PAY) pay[u] = __ldcs(pp.in_pay + p);
                }
            } else {
                w[u] = ~0ull;
            }
        }
#pragma unroll
        for (int u = 0; u < kPartPer; u++) {
            if (w[u] != ~0ull) {
                const int key = (int)(w[u] >> kshift);
                const int dgt = (LEVEL == 1) ? (key >> 7) : (key & (kPartBins - 1));
                rank[u] = atomicAdd(&s_wcnt[warp][dgt], 1);
            }
        }
        __syncthreads();
        if (tid < kPartBins) {
            int total = 0;
#pragma unroll
            for (int ww = 0; ww < kPartThreads / 32; ww++) {
                const int c = s_wcnt[ww][tid];
                s_wcnt[ww][tid] = total;
                total += c;
            }
            s_tot[tid] = total;
            if (LEVEL == 1) {
                s_base[tid] = s_cur[tid];
                s_cur[tid] += total;
            } else if (total) {
                s_base[tid] = atomicAdd(&cursors[s_bucket * kPartBins + tid], total);
            }
        }
        __syncthreads();
        if (tid < 32) {  // exclusive scan of the 128 digit totals: where a digit's run starts inside the tile
            int v[kPartBins / 32], run = 0;
#pragma unroll
            for (int k = 0; k < kPartBins / 32; k++) { v[k] = s_tot[4 * tid + k]; run += v[k]; }
            int incl = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, incl, o);
                if (tid >= o) incl += y;
            }
            int excl = incl - run;
#pragma unroll
            for (int k = 0; k < kPartBins / 32; k++) { s_loc[4 * tid + k] = excl; excl += v[k]; }
        }
        __syncthreads();
        // the tile in digit order in shared memory, then out in runs: consecutive threads write consecutive words of
        // a run (a lane-per-element scatter costs one L2 sector write per word, the limit of these sweeps)
#pragma unroll
        for (int u = 0; u < kPartPer; u++) {
            if (w[u] != ~0ull) {
                const int key = (int)(w[u] >> kshift);
                const int dgt = (LEVEL == 1) ? (key >> 7) : (key & (kPartBins - 1));
                const int at = s_loc[dgt] + s_wcnt[warp][dgt] + rank[u];
                s_out[at] = w[u];
                if (PAY) s_pay[at] = pay[u];  // the values take the same route through their own tile buffer
            }
        }
        __syncthreads();
        const int n_here = (int)(hi - lo);
#pragma unroll
        for (int u = 0; u < kPartPer; u++) {
            const int i = tid + u * kPartThreads;
            if (i < n_here) {
                const unsigned long long x = s_out[i];
                const int key = (int)(x >> kshift);
                const int dgt = (LEVEL == 1) ? (key >> 7) : (key & (kPartBins - 1));
                const int at = s_base[dgt] + (i - s_loc[dgt]);
                out_words[at] = x;
                if (PAY) pp.out_pay[at] = s_pay[i];
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ TT-DRM chain step
// v_out[p, b] = sum_a v_in[p, a] * core[a, idx[p], b]   (first core: v_out[p, b] = core[0, idx[p], b])
__global__ void __launch_bounds__(256) ttdrm_step_kernel(long long nnz, const long long* __restrict__ idx,
                                                        const double* __restrict__ v_in, int r_in,
                                                        const long long* __restrict__ vin_idx, long long vin_stride,
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
            const double* v = v_in + (vin_idx ? vin_idx[p] : p) * vin_stride;  // level 1 reads rows of the first core directly
            for (int a = 0; a < r_in; a++) s = fma(v[a], c[(long long)a * n * r_out], s);
        }
        v_out[e] = s;
    }
}

// ------------------------------------------------------------------ TT-DRM chain step, bucketed
// v_out[id, :] = v_in[id, :] @ core[:, i_m(id), :] for every nonzero, walked in the order sorted by
// i_m: all nonzeros of a segment share the r_in x r_out core slice, which is staged once in shared
// memory, and a tile of 64 nonzeros is a (64 x r_in) @ (r_in x r_out) product on DMMA.  (The
// reference gathers an (r_in, nnz, r_out) array and einsums it: tensor_train_drm.py:60-69.)
struct ChainParams {
    long long nnz, n_mu;
    const unsigned long long* keyid;
    const int* offs;
    long long work_items, item_len;
    const double* core;   // (r_in, n_mu, r_out)
    const double* v_in;   // row of nonzero id at v_in + (vin_idx ? vin_idx[id] : id) * vin_stride
    const long long* vin_idx;  // level 1: index row of the level-0 mode (v_in is the first core itself)
    long long vin_stride;
    double* v_out;        // (chunk, r_out) by nonzero id
    int r_in, r_out, pitch;
};

template <int NJ>
__global__ void __launch_bounds__(256) ttdrm_chain_kernel(const ChainParams C) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* Gs = reinterpret_cast<double*>(smem_raw);                    // [r_in_pad][pitch]
    const int r_in_pad = (C.r_in + 3) & ~3;
    unsigned long long* s_w = reinterpret_cast<unsigned long long*>(Gs + r_in_pad * C.pitch);  // [64]
    long long* s_bounds = reinterpret_cast<long long*>(s_w + 64);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int ksteps = r_in_pad >> 2;
    for (long long item = blockIdx.x; item < C.work_items; item += gridDim.x) {
        __syncthreads();
        if (tid == 0) {
            long long lo = item * C.item_len, hi = (item + 1) * C.item_len;
            if (hi > C.nnz || item == C.work_items - 1) hi = C.nnz;
            if (item > 0) lo = snap_to_segment(C.offs, C.n_mu, lo, C.item_len / 2);
            if (hi < C.nnz) hi = snap_to_segment(C.offs, C.n_mu, hi, C.item_len / 2);
            s_bounds[0] = lo;
            s_bounds[1] = hi;
        }
        __syncthreads();
        const long long item_lo = s_bounds[0], item_hi = s_bounds[1];
        long long cur_key = -1;
        long long c = item_lo;
        while (c < item_hi) {
            __syncthreads();  // previous tile done with s_w / Gs
            if (tid < 64) s_w[tid] = (c + tid < item_hi) ? C.keyid[c + tid] : ~0ull;
            __syncthreads();
            const int key0 = (int)(s_w[0] >> 32);
            int len = 0;
            {
                bool open = true;
#pragma unroll
                for (int t = 0; t < 2; t++) {
                    const unsigned long long w = s_w[t * 32 + lane];
                    const bool same = (w != ~0ull) && ((int)(w >> 32) == key0);
                    const unsigned m = __ballot_sync(0xffffffffu, same);
                    if (open) {
                        if (m == 0xffffffffu) len += 32;
                        else { len += __ffs(~m) - 1; open = false; }
                    }
                }
            }
            if ((long long)key0 != cur_key) {  // stage this segment's core slice
                cur_key = key0;
                for (int e = tid; e < r_in_pad * C.pitch; e += 256) {
                    const int a = e / C.pitch, b = e - a * C.pitch;
                    Gs[e] = (a < C.r_in && b < C.r_out) ? C.core[((long long)a * C.n_mu + key0) * C.r_out + b] : 0.0;
                }
                __syncthreads();
            }
            const int row = 8 * warp + g;
            if (8 * warp < len) {
                const bool valid = row < len;
                const long long id = valid ? (long long)(s_w[row] & 0xffffffffull) : 0;
                const double* vin = C.v_in + (C.vin_idx ? C.vin_idx[id] : id) * C.vin_stride;
                double acc[NJ][2];
#pragma unroll
                for (int j = 0; j < NJ; j++) acc[j][0] = acc[j][1] = 0.0;
                // input-row fragments four k-steps at a time (independent loads in flight), then their MMAs
                for (int k0 = 0; k0 < ksteps; k0 += 4) {
                    double a[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int k = 4 * (k0 + u) + q;
                        a[u] = (k0 + u < ksteps && valid && k < C.r_in) ? __ldg(vin + k) : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        if (k0 + u < ksteps) {
                            const int k = 4 * (k0 + u) + q;
#pragma unroll
                            for (int j = 0; j < NJ; j++) dmma(acc[j][0], acc[j][1], a[u], Gs[k * C.pitch + 8 * j + g]);
                        }
                    }
                }
                if (valid) {
                    double* vout = C.v_out + id * C.r_out;
#pragma unroll
                    for (int j = 0; j < NJ; j++) {
                        const int col = 8 * j + 2 * q;
                        if (col < C.r_out) vout[col] = acc[j][0];
                        if (col + 1 < C.r_out) vout[col + 1] = acc[j][1];
                    }
                }
            }
            c += len;
        }
    }
}

template <int NJ>
static int launch_chain_t(ttsk_ctx* ctx, ChainParams& C, cudaStream_t st) {
    auto kern = ttdrm_chain_kernel<NJ>;
    C.pitch = tile_pitch(NJ);
    const int r_in_pad = (C.r_in + 3) & ~3;
    const size_t smem = (size_t)r_in_pad * C.pitch * 8 + 64 * 8 + 64;
    TTSK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    TTSK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)ctx->sm_count * per_sm;
    long long items = grid * 8;
    const long long min_len = 2048;
    if (items * min_len > C.nnz) items = (C.nnz + min_len - 1) / min_len;
    if (items < 1) items = 1;
    C.work_items = items;
    C.item_len = (C.nnz + items - 1) / items;
    if (grid > items) grid = items;
    kern<<<(unsigned)grid, 256, smem, st>>>(C);
    TTSK_LAUNCHED(ctx);
    return TTSK_OK;
}

static int launch_chain(ttsk_ctx* ctx, ChainParams& C, cudaStream_t st) {
    const int nj = (C.r_out + 7) / 8;
    TTSK_ARG(nj <= 8 && C.r_in <= 64, "TT-DRM chain: core rank above 64 is not supported by the bucketed kernel");
    switch (nj) {
        case 1: return launch_chain_t<1>(ctx, C, st);
        case 2: return launch_chain_t<2>(ctx, C, st);
        case 3: return launch_chain_t<3>(ctx, C, st);
        case 4: return launch_chain_t<4>(ctx, C, st);
        case 5: return launch_chain_t<5>(ctx, C, st);
        case 6: return launch_chain_t<6>(ctx, C, st);
        case 7: return launch_chain_t<7>(ctx, C, st);
        default: return launch_chain_t<8>(ctx, C, st);
    }
}

static int launch_pass(ttsk_ctx* ctx, PassParams& P, bool has_x, cudaStream_t st) {
    if (P.nnz <= 0) return TTSK_OK;
    return has_x ? launch_pass_with_x(ctx, P, st) : launch_pass_without_x(ctx, P, st);
}

// ---- sort buffers (workspace views) and the bucket pass
struct SortBufs {
    int* hist; int* offs; int* cursor;
    int* cta_cnt;  // per-CTA key counts / range starts of the two-level scatter
    unsigned long long* part_tmp;  // words after the first sweep of the two-level partition
    unsigned long long* pay_tmp;   // payload partition: values after the first sweep
    unsigned long long* payv;      // payload partition: values in sorted order (next to keyid)
    int* part_aux;                 // tile map of the second sweep [129], then the (CTA, coarse bucket) starts of the first [CTAs][128]
    // TT DRMs bucket a mode up to three times (chain level of either side + the mode pass): with n_modes > 0 every
    // mode keeps its own sorted words / segment starts for the chunk and is bucketed once
    int n_modes = 0;
    unsigned long long* keyid_m[TTSK_MAX_ORDER];
    int* offs_m[TTSK_MAX_ORDER];
    bool sorted_m[TTSK_MAX_ORDER];
    unsigned long long* keyid;
    unsigned* recs;
    int rec_words;
};
static int rec_words_for(int d) { return d <= 0 ? 0 : (int)align_up(2 + d, 8); }
constexpr int kMaxSortCtas = 2 * 160;  // the two-level scatter runs two CTAs per SM
static int64_t cta_cnt_bytes(int64_t n_max) { return n_max <= kLocalBins ? (int64_t)kMaxSortCtas * n_max * 4 : 256; }
static int64_t sortbufs_bytes(int64_t n_max, int64_t chunk, int d, int per_mode = 0) {
    return 3 * align_up((n_max + 1) * 4, 256) + 4 * align_up(chunk * 8, 256) + align_up(chunk * 4 * rec_words_for(d), 256) +
           align_up(cta_cnt_bytes(n_max), 256) + 4096 + align_up((128 + 8 + (int64_t)kMaxSortCtas * 128) * 4, 256) +
           (per_mode ? (int64_t)d * (align_up(chunk * 8, 256) + align_up((n_max + 1) * 4, 256)) : 0);
}
static int carve_sortbufs(ttsk_ctx* ctx, SortBufs& sb, int64_t n_max, int64_t chunk, int d, int per_mode = 0) {
    sb.n_modes = per_mode ? d : 0;
    for (int m = 0; m < sb.n_modes; m++) {
        sb.keyid_m[m] = (unsigned long long*)ctx->ws_alloc(chunk * 8);
        sb.offs_m[m] = (int*)ctx->ws_alloc((n_max + 1) * 4);
        sb.sorted_m[m] = false;
        if (!sb.keyid_m[m] || !sb.offs_m[m]) {
            set_error("workspace carve failed (per-mode sort buffers)");
            return TTSK_E_NOMEM;
        }
    }
    sb.rec_words = rec_words_for(d);
    sb.recs = sb.rec_words ? (unsigned*)ctx->ws_alloc(chunk * 4 * sb.rec_words) : nullptr;
    if (sb.rec_words && !sb.recs) {
        set_error("workspace carve failed (packed records)");
        return TTSK_E_NOMEM;
    }
    sb.hist = (int*)ctx->ws_alloc((n_max + 1) * 4);
    sb.offs = (int*)ctx->ws_alloc((n_max + 1) * 4);
    sb.cursor = (int*)ctx->ws_alloc((n_max + 1) * 4);
    sb.keyid = (unsigned long long*)ctx->ws_alloc(chunk * 8);
    sb.cta_cnt = (int*)ctx->ws_alloc(cta_cnt_bytes(n_max));
    sb.part_tmp = (unsigned long long*)ctx->ws_alloc(chunk * 8);
    sb.pay_tmp = (unsigned long long*)ctx->ws_alloc(chunk * 8);
    sb.payv = (unsigned long long*)ctx->ws_alloc(chunk * 8);
    sb.part_aux = (int*)ctx->ws_alloc((kPartBins + 8 + (int64_t)kMaxSortCtas * kPartBins) * 4);
    if (!sb.hist || !sb.offs || !sb.cursor || !sb.keyid || !sb.cta_cnt || !sb.part_tmp || !sb.part_aux || !sb.pay_tmp || !sb.payv) {
        set_error("workspace carve failed (sort buffers)");
        return TTSK_E_NOMEM;
    }
    return TTSK_OK;
}

// bucket the chunk by the index row `key_idx`: sb.keyid receives (key << 32 | id) in sorted order
// With `pay` (see PayPlan) and the two-level partition available, the sorted elements are (word, value) pairs in
// sb.keyid / sb.payv and *pay_done is set; otherwise the id form is produced as usual.
static int sort_keys(ttsk_ctx* ctx, int64_t nnz, const long long* key_idx, int64_t n_mu, SortBufs& sb,
                     cudaStream_t st, int mode = -1, const PayPlan* pay = nullptr, int pay_kshift = 32,
                     bool* pay_done = nullptr) {
    if (pay_done) *pay_done = false;
    if (mode >= 0 && mode < sb.n_modes) {
        sb.keyid = sb.keyid_m[mode];
        sb.offs = sb.offs_m[mode];
        if (sb.sorted_m[mode]) return TTSK_OK;  // this chunk is already bucketed by this mode
        sb.sorted_m[mode] = true;
    }
    long long blocks = (nnz + 255) / 256;
    if (blocks > (long long)ctx->sm_count * 16) blocks = (long long)ctx->sm_count * 16;
    long long block_len = (nnz + 2LL * ctx->sm_count - 1) / (2LL * ctx->sm_count);
    if (block_len < 65536) block_len = 65536;
    const long long local_grid = (nnz + block_len - 1) / block_len;
    const bool local = n_mu <= kLocalBins && nnz >= 65536 && local_grid <= kMaxSortCtas;
    static const int part_off = getenv("TTSK_SORT_ONE_SWEEP") ? atoi(getenv("TTSK_SORT_ONE_SWEEP")) : 0;
    const bool two_level = local && !part_off && n_mu <= kPartBins * kPartBins && nnz < ((int64_t)1 << 31);
    TTSK_CUDA(cudaMemsetAsync(sb.hist, 0, (size_t)n_mu * sizeof(int), st));
    if (local) {
        const size_t smem = (size_t)n_mu * 4;
        TTSK_CUDA(cudaFuncSetAttribute(hist_local_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        hist_local_kernel<<<(unsigned)local_grid, 1024, smem, st>>>(key_idx, nnz, block_len, (int)n_mu, sb.hist, sb.cta_cnt);
    } else {
        hist_kernel<<<(unsigned)blocks, 256, 0, st>>>(key_idx, nnz, sb.hist);
    }
    TTSK_LAUNCHED(ctx);
    scan_kernel<<<1, 1024, 0, st>>>(sb.hist, n_mu, sb.offs, sb.cursor);
    TTSK_LAUNCHED(ctx);
    ScatterParams S;
    S.nnz = nnz;
    S.key_idx = key_idx;
    S.cursor = sb.cursor;
    S.keyid = sb.keyid;
    if (two_level) {
        int* tstart = sb.part_aux;
        int* l1base = sb.part_aux + kPartBins + 8;
        cta_base_kernel<<<(unsigned)((n_mu + 255) / 256), 256, 0, st>>>(sb.cta_cnt, sb.offs, (int)n_mu, (int)local_grid);
        TTSK_LAUNCHED(ctx);
        part_l1base_kernel<<<(unsigned)local_grid, kPartBins, 0, st>>>(sb.cta_cnt, sb.offs, (int)n_mu, l1base);
        TTSK_LAUNCHED(ctx);
        part_setup_kernel<<<1, kPartBins, 0, st>>>(sb.offs, (int)n_mu, tstart);
        TTSK_LAUNCHED(ctx);
        const unsigned grid2 = (unsigned)std::min<long long>((nnz + kPartTile - 1) / kPartTile, 4LL * ctx->sm_count);
        if (pay && sb.pay_tmp && sb.payv) {
            PayPlan pp = *pay;
            pp.in_pay = nullptr;
            pp.out_pay = sb.pay_tmp;
            const size_t pay_smem = (size_t)kPartTile * 8;
            TTSK_CUDA(cudaFuncSetAttribute(partition_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pay_smem));
            TTSK_CUDA(cudaFuncSetAttribute(partition_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pay_smem));
            partition_kernel<1, true><<<(unsigned)local_grid, kPartThreads, pay_smem, st>>>(key_idx, nullptr, nnz, (int)n_mu, sb.offs, tstart,
                                                                                    l1base, sb.part_tmp, block_len, pay_kshift, pp);
            TTSK_LAUNCHED(ctx);
            pp.in_pay = sb.pay_tmp;
            pp.out_pay = sb.payv;
            partition_kernel<2, true><<<grid2, kPartThreads, pay_smem, st>>>(nullptr, sb.part_tmp, nnz, (int)n_mu, sb.offs, tstart, sb.cursor,
                                                                     sb.keyid, 0, pay_kshift, pp);
            TTSK_LAUNCHED(ctx);
            if (pay_done) *pay_done = true;
            return TTSK_OK;
        }
        PayPlan none;
        std::memset(&none, 0, sizeof(none));
        partition_kernel<1, false><<<(unsigned)local_grid, kPartThreads, 0, st>>>(key_idx, nullptr, nnz, (int)n_mu, sb.offs, tstart, l1base,
                                                                                 sb.part_tmp, block_len, 32, none);
        TTSK_LAUNCHED(ctx);
        const unsigned grid = (unsigned)std::min<long long>((nnz + kPartTile - 1) / kPartTile, 4LL * ctx->sm_count);
        partition_kernel<2, false><<<grid, kPartThreads, 0, st>>>(nullptr, sb.part_tmp, nnz, (int)n_mu, sb.offs, tstart, sb.cursor, sb.keyid, 0,
                                                                 32, none);
        TTSK_LAUNCHED(ctx);
        return TTSK_OK;
    }
    if (local) {
        const size_t smem = (size_t)n_mu * 8;
        TTSK_CUDA(cudaFuncSetAttribute(scatter_local_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cta_base_kernel<<<(unsigned)((n_mu + 255) / 256), 256, 0, st>>>(sb.cta_cnt, sb.offs, (int)n_mu, (int)local_grid);
        TTSK_LAUNCHED(ctx);
        static const int rounds_env = getenv("TTSK_SORT_ROUNDS") ? atoi(getenv("TTSK_SORT_ROUNDS")) : 0;
        int rounds = (int)((local_grid * n_mu * 128 + ((int64_t)48 << 20) - 1) / ((int64_t)48 << 20));  // open ranges x one line each
        if (rounds > 8) rounds = 8;
        if (rounds_env > 0) rounds = rounds_env;
        if (rounds < 1) rounds = 1;
        scatter_local_kernel<<<(unsigned)local_grid, 1024, smem, st>>>(S, block_len, (int)n_mu, sb.cta_cnt, rounds);
    } else {
        scatter_kernel<<<(unsigned)blocks, 256, 0, st>>>(S);
    }
    TTSK_LAUNCHED(ctx);
    return TTSK_OK;
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

static int ttdrm_step(ttsk_ctx* ctx, int64_t nnz, const long long* idx_mu, const double* v_in, int r_in,
                      const long long* vin_idx, int64_t vin_stride, const double* core, int64_t n, int r_out,
                      double* v_out, cudaStream_t st) {
    if (nnz <= 0) return TTSK_OK;
    const long long total = (long long)nnz * r_out;
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)ctx->sm_count * 16) blocks = (long long)ctx->sm_count * 16;
    ttdrm_step_kernel<<<(unsigned)blocks, 256, 0, st>>>(nnz, idx_mu, v_in, r_in, vin_idx, vin_stride, core, n, r_out, v_out);
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
    int64_t b = 8 + 4 * rec_words_for(d);  // sorted (key, id) words + packed records
    if (left->kind == TTSK_DRM_TT || right->kind == TTSK_DRM_TT) b += 8LL * d;  // per-mode sorted words
    if (left->kind == TTSK_DRM_TT)
        for (int k = 1; k < d - 1; k++) b += 8LL * left->core_r1[k];
    if (right->kind == TTSK_DRM_TT)
        for (int k = 1; k < d - 1; k++) b += 8LL * right->core_r1[k];
    return b;
}

static int sparse_chunk(ttsk_ctx* ctx, SparsePlan& pl, int d, const int64_t* shape, int64_t nnz,
                        const int64_t* d_idx, int64_t idx_row_stride, const double* d_val, const ttsk_drm* left,
                        const ttsk_drm* right, double* out, SortBufs& sb, cudaStream_t st) {
    if (nnz <= 0) return TTSK_OK;
    const SketchLayout& lay = pl.lay;
    for (int m = 0; m < sb.n_modes; m++) sb.sorted_m[m] = false;
    // TT-DRM chains for this chunk (tensor_train_drm.py:60-69): level 0 is a row gather of the first
    // core; every further level is one bucketed pass over the nonzeros sorted by that level's mode
    for (int side = 0; side < 2; side++) {
        const ttsk_drm* drm = side == 0 ? left : right;
        SideState& ss = side == 0 ? pl.left : pl.right;
        if (drm->kind != TTSK_DRM_TT) continue;
        // level 0 is the first core itself (a table indexed by that mode: no buffer, see build_plan)
        const int mode0 = side == 0 ? 0 : d - 1;
        const long long* idx_0 = (const long long*)(d_idx + mode0 * idx_row_stride);
        for (int k = 1; k < d - 1; k++) {
            const int mode = side == 0 ? k : d - 1 - k;
            const long long* idx_m = (const long long*)(d_idx + mode * idx_row_stride);
            const double* v_in = k == 1 ? drm->d_cores[0] : ss.chain[k - 1];
            const long long* vin_idx = k == 1 ? idx_0 : nullptr;
            const int64_t vin_stride = k == 1 ? drm->core_r1[0] : drm->core_r0[k];
            const bool bucketed = nnz >= 4096 && drm->core_r1[k] <= 64 && drm->core_r0[k] <= 64;
            if (!bucketed) {
                TTSK_TRY(ttdrm_step(ctx, nnz, idx_m, v_in, drm->core_r0[k], vin_idx, vin_stride, drm->d_cores[k],
                                    shape[mode], drm->core_r1[k], ss.chain[k], st));
                continue;
            }
            TTSK_TRY(sort_keys(ctx, nnz, idx_m, shape[mode], sb, st, mode));
            ChainParams C;
            std::memset(&C, 0, sizeof(C));
            C.nnz = nnz; C.n_mu = shape[mode]; C.keyid = sb.keyid; C.offs = sb.offs;
            C.core = drm->d_cores[k]; C.v_in = v_in; C.vin_idx = vin_idx; C.vin_stride = vin_stride; C.v_out = ss.chain[k];
            C.r_in = drm->core_r0[k]; C.r_out = drm->core_r1[k];
            TTSK_TRY(launch_chain(ctx, C, st));
        }
    }
    const long long* idx_rows[TTSK_MAX_ORDER];
    for (int m = 0; m < d; m++) idx_rows[m] = (const long long*)(d_idx + m * idx_row_stride);
    {
        ScatterIdx rows;
        for (int m = 0; m < TTSK_MAX_ORDER; m++) rows.idx[m] = m < d ? idx_rows[m] : nullptr;
        long long blocks = (nnz + 255) / 256;
        if (blocks > (long long)ctx->sm_count * 16) blocks = (long long)ctx->sm_count * 16;
        pack_records_kernel<<<(unsigned)blocks, 256, 0, st>>>(d, nnz, rows, d_val, sb.recs, sb.rec_words);
        TTSK_LAUNCHED(ctx);
    }
    for (int mu = 0; mu < d; mu++) {
        PassParams P;
        std::memset(&P, 0, sizeof(P));
        P.nnz = nnz;
        P.n_mu = shape[mu];
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
        P.recs = sb.recs;
        P.rec_words = sb.rec_words;
        P.val = d_val;
        P.psi = out + lay.psi_off[mu];
        P.sg_mode = mu;
        bool flat_done = false;
        const bool time_it = ctx->timing;
        auto mark = [&](int which) -> int {
            if (!time_it) return TTSK_OK;
            while ((int)ctx->ev_pass.size() < 2 * (ctx->n_pass_events + 1)) {
                cudaEvent_t ev;
                TTSK_CUDA(cudaEventCreate(&ev));
                ctx->ev_pass.push_back(ev);
            }
            TTSK_CUDA(cudaEventRecord(ctx->ev_pass[2 * ctx->n_pass_events + which], st));
            if (which == 1) ctx->n_pass_events++;
            return TTSK_OK;
        };
        // last mode: no bucketing needed when the mode fits shared memory -- but the unbucketed form adds every
        // generated entry into the CTA's T with a CAS loop (no native FP64 shared-memory atomic: ~1.7 ps per entry
        // measured), the sorted form keeps per-column sums in registers and costs one bucketing (~12 ps per nonzero
        // since the two-level partition): sorted wins from about 8 columns on
        static const int last_env = getenv("TTSK_LAST_MODE_FLAT") ? atoi(getenv("TTSK_LAST_MODE_FLAT")) : -1;
        const bool prefer_sorted = last_env >= 0 ? last_env == 0 : (P.rA >= 8 && nnz >= 65536 && shape[mu] <= 16384);
        if (mu == d - 1 && !has_x && !prefer_sorted) {
            TTSK_TRY(mark(0));
            TTSK_TRY(try_launch_gw_flat(ctx, P, st, &flat_done));
            if (!flat_done) TTSK_TRY(launch_last_mode_unbucketed(ctx, P, st, &flat_done));
            if (flat_done) { TTSK_TRY(mark(1)); continue; }
        }
        // both factors tabulated: the payload partition + the bulk-copy gather pass (ttsk_sparse_gather.cu)
        PayPlan pay;
        int pay_kshift = 32;
        bool pay_done = false;
        const bool want_pay = sb.n_modes == 0 && gt_plan(P, has_x, &pay, &pay_kshift);
        TTSK_TRY(sort_keys(ctx, nnz, idx_rows[mu], shape[mu], sb, st, mu, want_pay ? &pay : nullptr, pay_kshift, &pay_done));
        P.keyid = sb.keyid;
        P.offs = sb.offs;
        TTSK_TRY(mark(0));
        if (pay_done) {
            TTSK_TRY(launch_gt(ctx, P, sb.keyid, sb.payv, pay, pay_kshift, st));
            TTSK_TRY(mark(1));
            continue;
        }
        bool gw_done = false;  // warp-autonomous forms (ttsk_sparse_gen.cu) when exactly one source is generated
        if (!has_x) TTSK_TRY(try_launch_gw_direct(ctx, P, st, &gw_done));
        if (!gw_done) TTSK_TRY(try_launch_gw_seg(ctx, P, has_x, st, &gw_done));
        if (!gw_done) TTSK_TRY(launch_pass(ctx, P, has_x, st));
        TTSK_TRY(mark(1));
    }
    return TTSK_OK;
}

// Prefix tables are a function of the DRM alone (seed, column range, number of rows), so they are
// generated once per context and reused by later sketches with the same DRM (streaming updates,
// blocked sketches, repeated calls).  The cache is capped (ctx->table_cap, 6 GB by default); least
// recently used entries are dropped first, but NEVER an entry the plan being built already uses
// (entries are pinned with the plan's generation id).  When the cap cannot be met without that,
// `*out` stays nullptr unless `must` is set (edge tables, which are small and always needed), and
// the caller keeps generating that bond on the fly.
static int cached_gauss_table(ttsk_ctx* ctx, int64_t rows, int rank_min, int r, uint64_t seed, double** out,
                              cudaStream_t st, bool must) {
    *out = nullptr;
    for (size_t i = 0; i < ctx->tables.size(); i++) {
        auto& t = ctx->tables[i];
        if (t.seed == seed && t.rank_min == rank_min && t.r == r && t.rows == rows) {
            if (t.stream != st) TTSK_CUDA(cudaStreamSynchronize(t.stream));  // filled on another stream
            t.stream = st;
            t.pin_gen = ctx->plan_gen;
            *out = t.ptr;
            if (i + 1 != ctx->tables.size()) {  // most recently used entries live at the back
                auto e = t;
                ctx->tables.erase(ctx->tables.begin() + (long)i);
                ctx->tables.push_back(e);
            }
            return TTSK_OK;
        }
    }
    const int64_t bytes = rows * (int64_t)r * 8;
    bool synced = false;
    for (size_t i = 0; i < ctx->tables.size() && ctx->table_bytes + bytes > ctx->table_cap;) {
        if (ctx->tables[i].pin_gen == ctx->plan_gen) { i++; continue; }  // in use by the plan being built
        if (!synced) { TTSK_CUDA(cudaDeviceSynchronize()); synced = true; }
        TTSK_CUDA(cudaFree(ctx->tables[i].ptr));
        ctx->table_bytes -= ctx->tables[i].bytes;
        ctx->tables.erase(ctx->tables.begin() + (long)i);
    }
    if (ctx->table_bytes + bytes > ctx->table_cap && !must) return TTSK_OK;  // caller generates on the fly
    double* p = nullptr;
    cudaError_t e = cudaMalloc((void**)&p, (size_t)bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (!must) return TTSK_OK;
        set_error("DRM table allocation of %lld bytes failed: %s", (long long)bytes, cudaGetErrorString(e));
        return TTSK_E_NOMEM;
    }
    TTSK_TRY(gauss_table_launch(ctx, rows, rank_min, r, seed, p, st));
    ctx->tables.push_back({seed, rank_min, r, rows, p, bytes, st, ctx->plan_gen});
    ctx->table_bytes += bytes;
    *out = p;
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
    ctx->plan_gen++;  // tables fetched from here on are pinned until the next plan
    pl.n_max = 0;
    for (int mu = 0; mu < d; mu++) pl.n_max = std::max<int64_t>(pl.n_max, shape[mu]);
    pl.edge_L0 = nullptr;
    pl.edge_R = nullptr;
    const int64_t table_rows_cap = std::max<int64_t>(nnz_total, 1);  // a table row is cheaper than two on-the-fly rows and is cached
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
                    double* tab = nullptr;
                    TTSK_TRY(cached_gauss_table(ctx, rows, S.rank_min, S.r, S.seed, &tab, st, edge));
                    if (edge && side == 0) pl.edge_L0 = tab;
                    if (edge && side == 1) pl.edge_R = tab;
                    if (small && tab) {
                        // true (unwrapped) strides equal the wrapped ones because rows < 2^31
                        S.kind = SRC_TABLE;
                        S.base = tab;
                        S.row_stride = S.r;
                        S.col_stride = 1;
                        S.span_bytes = rows * S.r * 8;
                    }
                }
            } else {
                std::memset(&S, 0, sizeof(S));
                S.r = drm->rank_max[bond] - drm->rank_min[bond];
                const int k = drm->right ? d - 2 - bond : bond;  // chain level in DRM orientation
                if (k == 0) {  // the first core (1, n, r) IS the table of this bond, indexed by its mode
                    const int mode0 = drm->right ? d - 1 : 0;
                    S.kind = SRC_TABLE;
                    S.k = 1;
                    S.modes[0] = mode0;
                    S.strides[0] = 1;
                    S.base = drm->d_cores[0] + drm->rank_min[bond];
                    S.row_stride = drm->core_r1[0];
                    S.col_stride = 1;
                    S.span_bytes = shape[mode0] * (int64_t)drm->core_r1[0] * 8;
                    ss.chain[0] = nullptr;
                    continue;
                }
                S.kind = SRC_ROWS;
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
    add(sortbufs_bytes(n_max, chunk, d, left->kind == TTSK_DRM_TT || right->kind == TTSK_DRM_TT));   // hist/offs/cursor + sorted (key, id) words + packed records
    const int64_t table_rows_cap = std::max<int64_t>(nnz_total, 1);  // a table row is cheaper than two on-the-fly rows and is cached
    for (int side = 0; side < 2; side++) {
        const ttsk_drm* drm = side == 0 ? left : right;
        for (int bond = 0; bond < d - 1; bond++) {
            const int r = drm->rank_max[bond] - drm->rank_min[bond];
            if (drm->kind == TTSK_DRM_GAUSS) {
                int64_t rows = 1;
                bool ok = true;
                if (!drm->right) { for (int i = 0; i <= bond; i++) { rows *= shape[i]; if (rows >= ((int64_t)1 << 31)) { ok = false; break; } } }
                else { for (int i = d - 1; i > bond; i--) { rows *= shape[i]; if (rows >= ((int64_t)1 << 31)) { ok = false; break; } } }
                (void)ok; (void)r; (void)table_rows_cap;  // prefix tables live in the context's table cache, not in the arena
            } else {
                const int k = drm->right ? d - 2 - bond : bond;
                if (k > 0) add(chunk * (int64_t)drm->core_r1[k] * 8);
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
    SortBufs sb;
    TTSK_TRY(carve_sortbufs(ctx, sb, n_max, chunk, d, left->kind == TTSK_DRM_TT || right->kind == TTSK_DRM_TT));
    if (!tmp) {
        set_error("workspace carve failed");
        return TTSK_E_NOMEM;
    }
    if (ctx->timing) TTSK_CUDA(cudaEventRecord(ctx->ev_t0, st));
    TTSK_TRY(build_plan(ctx, pl, d, h_shape, nnz, chunk, left, right, st));
    double* sk = accumulate ? tmp : d_out;
    TTSK_CUDA(cudaMemsetAsync(sk, 0, (size_t)total * 8, st));
    for (int64_t c0 = 0; c0 < nnz; c0 += chunk) {
        const int64_t n = std::min<int64_t>(chunk, nnz - c0);
        TTSK_TRY(sparse_chunk(ctx, pl, d, h_shape, n, d_idx + c0, idx_row_stride, d_val + c0, left, right, sk, sb, st));
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

extern "C" int64_t ttsk_sg_pass_count(ttsk_ctx* ctx) { return ctx ? ctx->sg_passes : -1; }

extern "C" int ttsk_last_pass_ms(ttsk_ctx* ctx, double* ms, int cap, int* n_out) {
    TTSK_ARG(ctx != nullptr && n_out != nullptr && (cap == 0 || ms != nullptr), "last_pass_ms");
    TTSK_CUDA(cudaEventSynchronize(ctx->ev_t1));
    *n_out = ctx->n_pass_events;
    for (int i = 0; i < ctx->n_pass_events && i < cap; i++) {
        float p = 0.f;
        TTSK_CUDA(cudaEventElapsedTime(&p, ctx->ev_pass[2 * i], ctx->ev_pass[2 * i + 1]));
        ms[i] = p;
    }
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
    const int64_t stage_cap = ctx->stage_nnz;  // default 16M nonzeros = 640 MB per staging buffer at d=4
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
    SortBufs sb;
    TTSK_TRY(carve_sortbufs(ctx, sb, n_max, chunk, d, left->kind == TTSK_DRM_TT || right->kind == TTSK_DRM_TT));
    if (!d_stage[0] || !d_stage[1] || !sk) {
        set_error("workspace carve failed");
        return TTSK_E_NOMEM;
    }
    SparsePlan pl;
    if (ctx->timing) TTSK_CUDA(cudaEventRecord(ctx->ev_t0, st));
    TTSK_TRY(build_plan(ctx, pl, d, h_shape, nnz, chunk, left, right, st));
    TTSK_CUDA(cudaMemsetAsync(sk, 0, (size_t)total * 8, st));
    int buf = 0;
    // the first chunks are short (1/8, 3/8 of a staging buffer) so the first kernels start after a short copy
    int64_t step = std::max<int64_t>(chunk / 8, std::min<int64_t>(chunk, 1 << 20));
    for (int64_t c0 = 0, n = 0; c0 < nnz; c0 += n, buf ^= 1, step = std::min<int64_t>(chunk, step * 3)) {
        n = std::min<int64_t>(step, nnz - c0);
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
        TTSK_TRY(sparse_chunk(ctx, pl, d, h_shape, n, (const int64_t*)di, chunk, dv, left, right, sk, sb, st));
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
    return ttdrm_step(ctx, nnz, (const long long*)d_idx_mu, d_v_in, r_in, nullptr, r_in, d_core, n, r_out, d_v_out,
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

static int operator_pass(ttsk_ctx* ctx, int64_t nnz, const int64_t* d_idx_mu, int64_t n_mu, const double* d_val,
                         const double* d_left, int rL, int64_t l_ps, int64_t l_cs, const double* d_right, int rR,
                         int64_t r_ps, int64_t r_cs, double* d_out, cudaStream_t st) {
    if (nnz == 0) return TTSK_OK;
    const int64_t n_b = d_idx_mu ? n_mu : 1;
    TTSK_TRY(ctx->ws_reserve(sortbufs_bytes(n_b, nnz, 0) + 4096));
    ctx->ws_reset();
    SortBufs sb;
    TTSK_TRY(carve_sortbufs(ctx, sb, n_b, nnz, 0));
    PassParams P;
    std::memset(&P, 0, sizeof(P));
    P.nnz = nnz;
    P.n_mu = n_b;
    rows_source(P.A, d_left, rL, l_ps, l_cs);
    rows_source(P.B, d_right, rR, r_ps, r_cs);
    P.X.kind = SRC_NONE;
    P.rA = d_left ? rL : 1;
    P.rB = d_right ? rR : 1;
    P.rX = 0;
    P.psi = d_out;  // Omega is a Psi with a single slice
    if (d_idx_mu) {
        TTSK_TRY(sort_keys(ctx, nnz, (const long long*)d_idx_mu, n_b, sb, st));
        P.keyid = sb.keyid;
        P.offs = sb.offs;
    }
    P.val = d_val;
    return launch_pass(ctx, P, false, st);
}

extern "C" int ttsk_sparse_omega(ttsk_ctx* ctx, int64_t nnz, const double* d_val, const double* d_left, int rL,
                                 int64_t l_ps, int64_t l_cs, const double* d_right, int rR, int64_t r_ps,
                                 int64_t r_cs, double* d_omega, void* stream) {
    TTSK_ARG(ctx != nullptr, "ctx is NULL");
    TTSK_ARG(nnz >= 0 && rL >= 1 && rR >= 1 && nnz < ((int64_t)1 << 31), "omega dims");
    TTSK_ARG(nnz == 0 || (d_val && d_left && d_right && d_omega), "NULL pointer");
    return operator_pass(ctx, nnz, nullptr, 1, d_val, d_left, rL, l_ps, l_cs, d_right, rR, r_ps, r_cs, d_omega,
                         (cudaStream_t)stream);
}

extern "C" int ttsk_sparse_psi(ttsk_ctx* ctx, int64_t nnz, const int64_t* d_idx_mu, int64_t n_mu,
                               const double* d_val, const double* d_left, int rL, int64_t l_ps, int64_t l_cs,
                               const double* d_right, int rR, int64_t r_ps, int64_t r_cs, double* d_psi,
                               void* stream) {
    TTSK_ARG(ctx != nullptr, "ctx is NULL");
    TTSK_ARG(nnz >= 0 && n_mu >= 1 && nnz < ((int64_t)1 << 31) && n_mu < ((int64_t)1 << 31), "psi dims");
    TTSK_ARG(d_left != nullptr || d_right != nullptr, "sketch_psi_sparse needs at least one side (sparse_sketch.py:21-32)");
    TTSK_ARG(nnz == 0 || (d_val && d_idx_mu && d_psi), "NULL pointer");
    return operator_pass(ctx, nnz, d_idx_mu, n_mu, d_val, d_left, rL, l_ps, l_cs, d_right, rR, r_ps, r_cs, d_psi,
                         (cudaStream_t)stream);
}
