// Mode passes whose only per-nonzero source is ONE on-the-fly Gaussian DRM matrix: warp-autonomous form.
//
// Replaces (reference): SparseGaussianDRM.sketch_sparse + inds_to_normal (drm/sparse_gaussian_drm.py:29-44,
// drm/fast_lazy_gaussian.pyx:52-105,183-201) fused with sketch_psi_sparse / sketch_omega_sparse
// (sketching_methods/sparse_sketch.py:39-69) for the passes of a sparse sketch in which one side is generated
// and the other side is absent (first / last mode) or a small prefix table (segment-GEMM form, see
// ttsk_sparse_pass.cuh).  At BASELINE config 4 these are three of the four passes and 80 variates per nonzero.
//
// Why another form: the producer/consumer kernels of ttsk_sparse_pass.cuh spend ~60 % of their issue slots on
// instructions that are not FP64 arithmetic (ncu, profiles/r02a_*): a per-element lane-pattern load, a flat-index
// load and address arithmetic (the element -> lane map changes every step), producer warps, ring polling
// (10 % of all issued instructions), CAS-loop shared-memory atomics at ~29 instructions per 32 elements.  Here
//   * every warp is autonomous: it fetches the (key, id) words and packed records of 32 nonzeros itself (words two
//     tiles ahead in registers, records requested into L2 one tile ahead), so there is no producer warp, no
//     mbarrier ring and no polling;
//   * LANE = ROW: lane i owns nonzero i of the tile, its flat index lives in a register, the column salt is a
//     warp-uniform shared-memory broadcast, and a variate is stored column-major ([col][lane], conflict free) at an
//     immediate offset -- per variate only the 64-bit hash, the ndtri arithmetic, one store and the tail-queue
//     push remain;
//   * the tail branch of ndtri (27 % of draws) is still deferred to a per-warp queue and evaluated densely, four
//     entries per lane in flight;
//   * accumulation: first / last mode in sorted order keep per-column sums in registers (lane = column) and flush
//     them when the key changes; the table forms add v * row into the CTA's T with lane = row, so value and
//     table row index stay in registers and the only shared-memory traffic is one load and one atomic per element.
// Results are the same sums in another order (the generated entries are bit-identical by construction: same
// device functions as ttsk_lazy_gaussian).
#include "ttsk_sparse_pass.cuh"

namespace ttsk {

constexpr int kGwWarps = 16;
constexpr int kGwThreads = 32 * kGwWarps;
constexpr int kGwPB = 34;     // column pitch (doubles) of a warp's generated block: [col][lane], 16-byte aligned columns
constexpr int kGwIL = 4;      // independent variates per lane in flight in the main loop
constexpr int kGwTailIL = 4;  // ... and in the tail drain

enum { GW_SEG_DIRECT = 0, GW_FLAT_T = 1, GW_SEG_T = 2 };

struct GwParams {
    long long nnz, n_mu;
    const unsigned long long* keyid;  // sorted (key << 32 | id) words; nullptr for GW_FLAT_T (original order)
    const unsigned* recs;             // 32-byte records [val | int32 idx[6]]
    const int* offs;                  // segment starts (n_mu + 1), GW_SEG_T
    int r, rank_min;                  // the generated source: columns, first column of the infinite matrix
    unsigned long long seed;
    long long smul_g[6];              // stride of every mode in the generated source's flat index (0: unused)
    long long smul_s[6];              // T forms: stride of every mode in the T row index
    int S_rows, pitch_t;              // rows of T, its row pitch in doubles
    // GW_SEG_DIRECT / GW_FLAT_T: out[key * key_stride + col * col_stride] += sum
    double* out;
    long long key_stride, col_stride;
    // GW_SEG_T: Psi[:, key, :] += T^T B,  Omega += T^T X[key S : (key + 1) S]
    const double* Btab; long long b_rs; int rB;
    const double* Xtab; long long x_rs; int rX;
    double* psi; double* omega;
    long long work_items, item_len;
};

__device__ __forceinline__ void ld_shared_v2u64(unsigned addr, unsigned long long& a, unsigned long long& b) {
    asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}
__device__ __forceinline__ double2 ld_shared_v2f64(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

// hash -> b = 1 + u (two words) of IL columns of the lane's row; `salt_addr` -> the columns' salts (u64, warp-uniform)
template <int IL>
__device__ __forceinline__ void gw_hash(unsigned long long flat, unsigned salt_addr, unsigned (&hi)[IL], unsigned (&lo)[IL]) {
    if constexpr (IL % 2 == 0) {
#pragma unroll
        for (int c = 0; c < IL; c += 2) {
            unsigned long long s0, s1;
            ld_shared_v2u64(salt_addr + 8u * c, s0, s1);
            hash_to_b(flat + s0, hi[c], lo[c]);
            hash_to_b(flat + s1, hi[c + 1], lo[c + 1]);
        }
    } else {
#pragma unroll
        for (int c = 0; c < IL; c++) hash_to_b(flat + ld_shared_u64(salt_addr + 8u * c), hi[c], lo[c]);
    }
}

// Classify IL hashed columns and queue them.  Every element stores b = 1 + u (the uniform with the exponent of 1.0) in
// its slot and pushes the slot (address >> 3, 16 bits) on one of the warp's two queues, which share one array of
// r * 32 entries: elements that are surely in the central branch of ndtri grow a queue from the front, elements
// whose high word says they MAY be in a tail grow one from the back.  Both branches are then evaluated densely --
// no lane computes a central branch whose result a tail would overwrite (27 % of the lanes did before).
// `cq_lane` = central front + 2 * lane (per lane), `tq` = tail back (warp-uniform).  One predicate per variate.
template <int IL>
__device__ __forceinline__ void gw_classify(const unsigned (&hi)[IL], const unsigned (&lo)[IL], unsigned slot0, unsigned& cq_lane,
                                            unsigned& tq, unsigned lt) {
    const unsigned e0 = slot0 >> 3;
#pragma unroll
    for (int c = 0; c < IL; c++) {
        const unsigned bh = hi[c] | 0x3FF00000u;
        const double b = __hiloint2double((int)bh, (int)lo[c]);
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            ".reg .b32 m, t, n2, c2, at, ac;\n"
            "sub.u32 t, %2, %3;\n"
            "setp.ge.u32 p, t, %4;\n"            // may be a tail
            "st.shared.f64 [%6], %5;\n"
            "vote.sync.ballot.b32 m, p, 0xffffffff;\n"
            "popc.b32 n2, m;\n"
            "shl.b32 n2, n2, 1;\n"
            "and.b32 c2, m, %7;\n"
            "popc.b32 c2, c2;\n"
            "shl.b32 c2, c2, 1;\n"               // 2 * (tails before this lane)
            "sub.u32 %1, %1, n2;\n"              // the tail queue grows down by the step's tails
            "add.u32 at, %1, c2;\n"
            "sub.u32 ac, %0, c2;\n"              // central slot: front + 2 * (lane - tails before this lane)
            "selp.u32 at, at, ac, p;\n"
            "st.shared.u16 [at], %8;\n"
            "add.u32 %0, %0, 64;\n"
            "sub.u32 %0, %0, n2;\n"              // the central queue grows by 32 - tails
            "}"
            : "+r"(cq_lane), "+r"(tq)
            : "r"(bh), "n"(kCentralLo + 0x3FF00000u), "n"(kCentralSpan), "d"(b), "r"(slot0 + (unsigned)(kGwPB * 8 * c)), "r"(lt),
              "h"((unsigned short)(e0 + (unsigned)(kGwPB * c)))
            : "memory");
    }
}

// central branch of the queue entries [q0, q0 + 32 IL), IL per lane in flight; a lane without a further entry repeats
// its first one (same value written twice)
template <int IL>
__device__ __forceinline__ void gw_drain_central(unsigned q_addr, int q0, int count, int lane) {
    const int qi = q0 + lane;
    if (qi >= count) return;
    unsigned a[IL];
    double cc[IL];
#pragma unroll
    for (int c = 0; c < IL; c++) {
        const int qc = (qi + 32 * c < count) ? qi + 32 * c : qi;
        a[c] = ld_shared_u16(q_addr + 2u * qc) << 3;
    }
#pragma unroll
    for (int c = 0; c < IL; c++) cc[c] = ndtri_central_b(ld_shared_f64(a[c]));
#pragma unroll
    for (int c = 0; c < IL; c++) st_shared_f64(a[c], cc[c]);
}

// the warp's deferred tails [q0, q0 + 32 IL) of its queue, IL entries per lane in flight (one entry alone is a chain of
// ~100 dependent FP64 operations).  The generator pre-filters on the high word of the uniform only, so a (rare) entry
// may belong to the central branch after all; a lane without a further entry repeats its first one.
template <int IL>
__device__ __forceinline__ void gw_drain(unsigned wq_base, int q0, int wcount, unsigned tab_addr, int lane) {
    const int qi = q0 + lane;
    if (qi >= wcount) return;
    unsigned a[IL];
    double u[IL], out[IL];
    int cls[IL];
#pragma unroll
    for (int c = 0; c < IL; c++) {
        const int qc = (qi + 32 * c < wcount) ? qi + 32 * c : qi;
        a[c] = ld_shared_u16(wq_base + 2u * qc) << 3;
    }
#pragma unroll
    for (int c = 0; c < IL; c++) {
        u[c] = __dadd_rn(ld_shared_f64(a[c]), -1.0);
        cls[c] = ndtri_class(u[c]);
    }
    ndtri_tail_n<IL>(u, cls, SmemTab{tab_addr}, out);
#pragma unroll
    for (int c = 0; c < IL; c++) {
        if (cls[c] == 0) out[c] = ndtri_central(u[c]);  // rare, divergent
        st_shared_f64(a[c], out[c]);
    }
}

// the (r x 32) block of one tile: column `col` of lane's row at buf_lane + col * kGwPB * 8.  Three dense phases:
// hash + classify every element, central branch over the front queue, tail branch over the back queue.  (Measured on
// B200: an FP64 instruction costs two issue cycles and every other instruction one, with no overlap in this mix
// however the streams are interleaved, so what counts is the number of instructions per variate -- and that no
// lane evaluates a branch it will not keep.)
__device__ __forceinline__ void gw_generate(unsigned long long flat, unsigned salt_base, int r, unsigned buf_lane,
                                            unsigned wq_base, unsigned lt, unsigned tab_addr, int lane) {
    constexpr unsigned kColB = (unsigned)(kGwPB * 8);
    const unsigned wq_end = wq_base + 64u * (unsigned)r;  // r * 32 entries of two bytes
    unsigned cq_lane = wq_base + 2u * (unsigned)lane, tq = wq_end;
    int col = 0;
#pragma unroll 1
    for (; col + kGwIL <= r; col += kGwIL) {
        unsigned hi[kGwIL], lo[kGwIL];
        gw_hash<kGwIL>(flat, salt_base + 8u * col, hi, lo);
        gw_classify<kGwIL>(hi, lo, buf_lane + kColB * col, cq_lane, tq, lt);
    }
#pragma unroll 1
    for (; col < r; col++) {
        unsigned hi[1], lo[1];
        gw_hash<1>(flat, salt_base + 8u * col, hi, lo);
        gw_classify<1>(hi, lo, buf_lane + kColB * col, cq_lane, tq, lt);
    }
    __syncwarp();
    const int n_central = (int)((cq_lane - 2u * (unsigned)lane - wq_base) >> 1);
    const int n_tail = (int)((wq_end - tq) >> 1);
    int q0 = 0;
#pragma unroll 1
    for (; n_central - q0 > 64; q0 += 32 * kGwIL) gw_drain_central<kGwIL>(wq_base, q0, n_central, lane);
    if (n_central - q0 > 32) gw_drain_central<2>(wq_base, q0, n_central, lane);
    else if (n_central - q0 > 0) gw_drain_central<1>(wq_base, q0, n_central, lane);
    // deferred tails, dense over the back queue: four entries per lane while that keeps most lanes busy, then two, one
    q0 = 0;
#pragma unroll 1
    for (; n_tail - q0 > 64; q0 += 32 * kGwTailIL) gw_drain<kGwTailIL>(tq, q0, n_tail, tab_addr, lane);
    if (n_tail - q0 > 32) gw_drain<2>(tq, q0, n_tail, tab_addr, lane);
    else if (n_tail - q0 > 0) gw_drain<1>(tq, q0, n_tail, tab_addr, lane);
    __syncwarp();
}

// first key whose segment ends after position x (the key of the nonzero at sorted position x)
__device__ __forceinline__ int key_at(const int* __restrict__ offs, long long n_mu, long long x) {
    long long lo = 0, hi = n_mu - 1;  // smallest k with offs[k + 1] > x
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if ((long long)offs[mid + 1] > x) hi = mid; else lo = mid + 1;
    }
    return (int)lo;
}

template <int FORM, int MI, int NJ, bool HAS_X>
__global__ void __launch_bounds__(kGwThreads, 1) gw_kernel(const GwParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* s_tab = reinterpret_cast<double2*>(smem_raw);
    unsigned long long* s_salt = reinterpret_cast<unsigned long long*>(s_tab + kGaussTabEntries);
    const int r = P.r;
    const int r_pad = (r + 1) & ~1;
    const int PT = P.pitch_t;
    const int S_rows = P.S_rows, S_pad = (S_rows + 3) & ~3;
    double* T = reinterpret_cast<double*>(s_salt + r_pad);  // [S_pad][PT] (T forms)
    const int t_doubles = (FORM == GW_SEG_DIRECT) ? 0 : S_pad * PT;
    // per-warp area: generated block [r][kGwPB], tail queue (r * 32 slots), values and keys of the tile (direct form)
    const int q_bytes = ((r * 32 * 2 + 15) & ~15);
    const int warp_bytes = r * kGwPB * 8 + q_bytes + (FORM == GW_SEG_DIRECT ? 32 * 8 + 32 * 4 : 0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned char* my = reinterpret_cast<unsigned char*>(T + t_doubles) + (size_t)warp * warp_bytes;
    double* buf = reinterpret_cast<double*>(my);
    const unsigned buf_addr = smem_addr(buf);
    const unsigned wq_base = buf_addr + (unsigned)(r * kGwPB * 8);
    double* s_v = reinterpret_cast<double*>(my + r * kGwPB * 8 + q_bytes);  // direct form only
    int* s_key = reinterpret_cast<int*>(s_v + 32);

    load_logtab(s_tab);
    for (int c = tid; c < r_pad; c += kGwThreads)
        s_salt[c] = hash64((unsigned long long)(P.rank_min + (c < r ? c : 0))) + P.seed + kHashAdd;
    for (int i = tid; i < t_doubles; i += kGwThreads) T[i] = 0.0;
    __syncthreads();

    unsigned lt;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt));
    const unsigned tab_addr = smem_addr(s_tab);
    const unsigned salt_base = smem_addr(s_salt);
    const unsigned buf_lane = buf_addr + 8u * lane;

    // value, flat index of the generated source and T row index of nonzero `id`
    auto decode = [&](bool in, long long id, double& v, unsigned long long& flat, int& srow) {
        unsigned w[8];
        if (in) {
            const unsigned* rec = P.recs + id * 8;
            const uint4 r0 = __ldg(reinterpret_cast<const uint4*>(rec));
            const uint4 r1 = __ldg(reinterpret_cast<const uint4*>(rec) + 1);
            w[0] = r0.x; w[1] = r0.y; w[2] = r0.z; w[3] = r0.w;
            w[4] = r1.x; w[5] = r1.y; w[6] = r1.z; w[7] = r1.w;
        } else {
#pragma unroll
            for (int m = 0; m < 8; m++) w[m] = 0u;
        }
        v = __hiloint2double((int)w[1], (int)w[0]);  // 0 past the tile
        unsigned long long f = 0, s = 0;
#pragma unroll
        for (int m = 0; m < 6; m++) {
            f += (unsigned long long)w[2 + m] * (unsigned long long)P.smul_g[m];
            if (FORM != GW_SEG_DIRECT) s += (unsigned long long)w[2 + m] * (unsigned long long)P.smul_s[m];
        }
        flat = f;
        srow = (int)s;
    };
    auto prefetch_rec = [&](long long id) { asm volatile("prefetch.global.L2 [%0];" ::"l"(P.recs + id * 8)); };

    // T[srow, :] += v * (generated row of this lane) for the lanes in [ra, rb)
    auto t_accumulate = [&](int ra, int rb, double v, int srow) {
        if (lane >= ra && lane < rb) {
            double* trow = T + (size_t)srow * PT;
            int col = 0;
            for (; col + 4 <= r; col += 4) {
                double x[4];
#pragma unroll
                for (int c = 0; c < 4; c++) x[c] = buf[(col + c) * kGwPB + lane];
#pragma unroll
                for (int c = 0; c < 4; c++) atomicAdd(trow + col + c, __dmul_rn(x[c], v));
            }
            for (; col < r; col++) atomicAdd(trow + col, __dmul_rn(buf[col * kGwPB + lane], v));
        }
        __syncwarp();
    };

    if constexpr (FORM == GW_SEG_DIRECT) {
        // ------------------------------------------------------------ sorted order, per-column sums in registers
        const long long n_warps = (long long)gridDim.x * kGwWarps;
        const bool second = lane + 32 < r;
        const unsigned xa0 = buf_addr + (unsigned)(kGwPB * 8) * (lane < r ? lane : 0);
        const unsigned xa1 = buf_addr + (unsigned)(kGwPB * 8) * (second ? lane + 32 : 0);
        const unsigned v_addr = smem_addr(s_v);
        double d0 = 0.0, d1 = 0.0;
        int cur_key = -1;
        auto flush = [&]() {
            if (cur_key >= 0) {
                double* dst = P.out + (long long)cur_key * P.key_stride;
                if (lane < r && d0 != 0.0) atomicAdd(dst + (long long)lane * P.col_stride, d0);
                if (second && d1 != 0.0) atomicAdd(dst + (long long)(lane + 32) * P.col_stride, d1);
            }
            d0 = d1 = 0.0;
        };
        for (long long item = (long long)blockIdx.x * kGwWarps + warp; item < P.work_items; item += n_warps) {
            const long long item_lo = item * P.item_len;
            long long item_hi = item_lo + P.item_len;
            if (item_hi > P.nnz || item == P.work_items - 1) item_hi = P.nnz;
            auto word = [&](long long q) -> unsigned long long { return (q + lane < item_hi) ? P.keyid[q + lane] : 0ull; };
            unsigned long long w_cur = word(item_lo), w_nxt = word(item_lo + 32);
            for (long long q = item_lo; q < item_hi; q += 32) {
                const int n_rows = (int)((item_hi - q < 32) ? item_hi - q : 32);
                if (q + 32 + lane < item_hi) prefetch_rec((long long)(w_nxt & 0xffffffffull));
                const unsigned long long w_nn = word(q + 64);
                const bool in = lane < n_rows;
                double v;
                unsigned long long flat;
                int srow;
                decode(in, (long long)(w_cur & 0xffffffffull), v, flat, srow);
                const int key = in ? (int)(w_cur >> 32) : -1;
                s_v[lane] = v;
                s_key[lane] = key;
                gw_generate(flat, salt_base, r, buf_lane, wq_base, lt, tab_addr, lane);  // ends with __syncwarp
                if (__all_sync(0xffffffffu, key == cur_key)) {
                    // a whole tile of the current segment: two rows per step, all loads of a step first
#pragma unroll 4
                    for (int p = 0; p < 32; p += 2) {
                        const double2 vv = ld_shared_v2f64(v_addr + 8u * p);
                        const double2 x0 = ld_shared_v2f64(xa0 + 8u * p);
                        d0 = fma(vv.x, x0.x, d0);
                        d0 = fma(vv.y, x0.y, d0);
                        if (second) {
                            const double2 x1 = ld_shared_v2f64(xa1 + 8u * p);
                            d1 = fma(vv.x, x1.x, d1);
                            d1 = fma(vv.y, x1.y, d1);
                        }
                    }
                } else {
                    for (int p = 0; p < n_rows; p++) {
                        const int k = s_key[p];
                        if (k != cur_key) {
                            flush();
                            cur_key = k;
                        }
                        const double vp = s_v[p];
                        d0 = fma(vp, ld_shared_f64(xa0 + 8u * p), d0);
                        if (second) d1 = fma(vp, ld_shared_f64(xa1 + 8u * p), d1);
                    }
                }
                __syncwarp();
                w_cur = w_nxt;
                w_nxt = w_nn;
            }
            flush();
            cur_key = -1;
        }
    } else if constexpr (FORM == GW_FLAT_T) {
        // ------------------------------------------------------------ original order, T[i_mu, :] shared by the CTA
        const long long n_tiles = (P.nnz + 31) / 32;
        const long long stride = (long long)gridDim.x * kGwWarps;
        for (long long t = (long long)blockIdx.x * kGwWarps + warp; t < n_tiles; t += stride) {
            const long long q = 32 * t;
            if (32 * (t + stride) + lane < P.nnz) prefetch_rec(32 * (t + stride) + lane);
            const bool in = q + lane < P.nnz;
            double v;
            unsigned long long flat;
            int srow;
            decode(in, q + lane, v, flat, srow);
            gw_generate(flat, salt_base, r, buf_lane, wq_base, lt, tab_addr, lane);
            const int n_rows = (int)((P.nnz - q < 32) ? P.nnz - q : 32);
            t_accumulate(0, n_rows, v, srow);
        }
        __syncthreads();
        for (int e = tid; e < S_rows * r; e += kGwThreads) {
            const int srow = e / r, a = e - srow * r;
            const double x = T[srow * PT + a];
            if (x != 0.0) atomicAdd(P.out + (long long)srow * P.key_stride + (long long)a * P.col_stride, x);
        }
    } else {
        // ------------------------------------------------------------ sorted order, T_j shared by the CTA, one pair of
        // small GEMMs per segment (per CTA that holds a part of it)
        const int g = lane >> 2, qq = lane & 3;
        auto segment_gemm = [&](int key) {
            named_barrier(1, kGwThreads);  // every warp has finished adding to T
            const bool omega_role = HAS_X && warp >= kGwWarps / 2;
            const int role_warps = HAS_X ? kGwWarps / 2 : kGwWarps;
            const int role_rank = HAS_X ? (warp & (kGwWarps / 2 - 1)) : warp;
            const int rR = omega_role ? P.rX : P.rB;
            const long long rs = omega_role ? P.x_rs : P.b_rs;
            const double* tab = omega_role ? P.Xtab + (long long)key * S_rows * rs : P.Btab;
            double acc[MI][NJ][2];
#pragma unroll
            for (int i = 0; i < MI; i++)
#pragma unroll
                for (int j = 0; j < NJ; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
            for (int ch = role_rank; 4 * ch < S_pad; ch += role_warps) {
                const int srow = 4 * ch + qq;
                double a[MI], b[NJ];
#pragma unroll
                for (int i = 0; i < MI; i++) a[i] = (8 * i + g < PT) ? T[srow * PT + 8 * i + g] : 0.0;  // columns >= r of T stay zero
                const double* brow = tab + (long long)srow * rs;
#pragma unroll
                for (int j = 0; j < NJ; j++) b[j] = (srow < S_rows && 8 * j + g < rR) ? __ldg(brow + 8 * j + g) : 0.0;
#pragma unroll
                for (int i = 0; i < MI; i++)
#pragma unroll
                    for (int j = 0; j < NJ; j++) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
            double* dst_base = omega_role ? P.omega : P.psi + (long long)key * P.rB;
            const long long row_pitch = omega_role ? (long long)P.rX : (long long)P.n_mu * P.rB;
#pragma unroll
            for (int i = 0; i < MI; i++)
#pragma unroll
                for (int j = 0; j < NJ; j++) {
                    const int row = 8 * i + g, col = 8 * j + 2 * qq;
                    if (row < r) {
                        double* dst = dst_base + (long long)row * row_pitch + col;
                        if (col < rR && acc[i][j][0] != 0.0) atomicAdd(dst, acc[i][j][0]);
                        if (col + 1 < rR && acc[i][j][1] != 0.0) atomicAdd(dst + 1, acc[i][j][1]);
                    }
                }
            named_barrier(1, kGwThreads);  // all products read T
            for (int i = tid; i < S_pad * PT; i += kGwThreads) T[i] = 0.0;
            named_barrier(1, kGwThreads);
        };
        for (long long item = blockIdx.x; item < P.work_items; item += gridDim.x) {
            long long lo = 0, hi = 0;
            int k0 = 0;
            if (lane == 0) {
                lo = item * P.item_len;
                hi = (item + 1) * P.item_len;
                if (hi > P.nnz || item == P.work_items - 1) hi = P.nnz;
                if (item > 0) lo = snap_to_segment(P.offs, P.n_mu, lo, P.item_len / 2);
                if (hi < P.nnz) hi = snap_to_segment(P.offs, P.n_mu, hi, P.item_len / 2);
                if (lo < hi) k0 = key_at(P.offs, P.n_mu, lo);
            }
            const long long item_lo = __shfl_sync(0xffffffffu, lo, 0), item_hi = __shfl_sync(0xffffffffu, hi, 0);
            if (item_lo >= item_hi) continue;  // (uniform over the CTA)
            int seg = __shfl_sync(0xffffffffu, k0, 0);
            long long seg_end = P.offs[seg + 1];
            // close the current segment and move to the one that holds sorted position seg_end
            auto next_segment = [&]() {
                segment_gemm(seg);
                do { seg++; } while ((long long)P.offs[seg + 1] <= seg_end);
                seg_end = P.offs[seg + 1];
            };
            constexpr long long kStep = 32 * kGwWarps;
            auto word = [&](long long q) -> unsigned long long { return (q + lane < item_hi) ? P.keyid[q + lane] : 0ull; };
            const long long q0 = item_lo + 32 * warp;
            unsigned long long w_cur = word(q0), w_nxt = word(q0 + kStep);
            for (long long q = q0; q < item_hi; q += kStep) {
                while (seg_end <= q) next_segment();  // segments that end before my tile are complete for me
                const long long tile_hi = (q + 32 < item_hi) ? q + 32 : item_hi;
                if (q + kStep + lane < item_hi) prefetch_rec((long long)(w_nxt & 0xffffffffull));
                const unsigned long long w_nn = word(q + 2 * kStep);
                const bool in = q + lane < tile_hi;
                double v;
                unsigned long long flat;
                int srow;
                decode(in, (long long)(w_cur & 0xffffffffull), v, flat, srow);
                gw_generate(flat, salt_base, r, buf_lane, wq_base, lt, tab_addr, lane);
                long long pos = q;
                while (true) {
                    const long long run_hi = (tile_hi < seg_end) ? tile_hi : seg_end;
                    t_accumulate((int)(pos - q), (int)(run_hi - q), v, srow);
                    pos = run_hi;
                    if (pos >= tile_hi) break;
                    next_segment();
                }
                w_cur = w_nxt;
                w_nxt = w_nn;
            }
            while (seg_end < item_hi) next_segment();
            segment_gemm(seg);
        }
    }
}

// ------------------------------------------------------------------ host side
static bool gw_strides(const Source& S, long long* smul) {
    for (int m = 0; m < 6; m++) smul[m] = 0;
    for (int i = 0; i < S.k; i++) {
        if (S.modes[i] >= 6) return false;
        smul[S.modes[i]] += S.strides[i];
    }
    return true;
}

static size_t gw_smem(int form, int r, int S_rows, int pitch_t) {
    const int r_pad = (r + 1) & ~1, S_pad = (S_rows + 3) & ~3;
    const size_t q_bytes = (size_t)((r * 32 * 2 + 15) & ~15);
    const size_t warp_bytes = (size_t)r * kGwPB * 8 + q_bytes + (form == GW_SEG_DIRECT ? 32 * 8 + 32 * 4 : 0);
    return (size_t)kGaussTabEntries * 16 + (size_t)r_pad * 8 + (form == GW_SEG_DIRECT ? 0 : (size_t)S_pad * pitch_t * 8) +
           (size_t)kGwWarps * warp_bytes;
}

// T row pitch: >= r (and >= 8 * MI so the MMA fragments of the segment GEMM stay inside a row), odd, so the
// rows touched by the 32 lanes of an update spread over the banks
static int gw_pitch_t(int r, int mi) {
    int p = std::max(r, 8 * mi);
    return p | 1;
}

template <int FORM, int MI, int NJ, bool HAS_X>
static int gw_launch(ttsk_ctx* ctx, GwParams& G, cudaStream_t st) {
    auto kern = gw_kernel<FORM, MI, NJ, HAS_X>;
    const size_t smem = gw_smem(FORM, G.r, G.S_rows, G.pitch_t);
    TTSK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long grid = ctx->sm_count;
    const long long tiles = (G.nnz + 31) / 32;
    if (FORM == GW_SEG_DIRECT) {
        // one item per warp and round: long enough that a warp sees few segment changes, short enough to balance
        long long items = grid * kGwWarps * 4;
        const long long min_len = 1024;
        if (items * min_len > G.nnz) items = (G.nnz + min_len - 1) / min_len;
        if (items < 1) items = 1;
        G.work_items = items;
        G.item_len = (G.nnz + items - 1) / items;
        G.item_len = (G.item_len + 31) / 32 * 32;
        G.work_items = (G.nnz + G.item_len - 1) / G.item_len;
        grid = std::min<long long>(grid, (G.work_items + kGwWarps - 1) / kGwWarps);
    } else if (FORM == GW_FLAT_T) {
        grid = std::min<long long>(grid, (tiles + kGwWarps - 1) / kGwWarps);
    } else {
        long long items = grid * 8;
        const long long min_len = 8192;
        if (items * min_len > G.nnz) items = (G.nnz + min_len - 1) / min_len;
        if (items < 1) items = 1;
        G.work_items = items;
        G.item_len = (G.nnz + items - 1) / items;
        grid = std::min<long long>(grid, items);
    }
    if (grid < 1) grid = 1;
    if (getenv("TTSK_DEBUG"))
        fprintf(stderr, "[ttsk] warp-autonomous generator pass form=%d r=%d S=%d MI=%d NJ=%d X=%d smem=%zu grid=%lld items=%lld x %lld\n",
                FORM, G.r, G.S_rows, MI, NJ, (int)HAS_X, smem, grid, G.work_items, G.item_len);
    kern<<<(unsigned)grid, kGwThreads, smem, st>>>(G);
    TTSK_LAUNCHED(ctx);
    return TTSK_OK;
}

static bool gw_enabled() {
    static const int off = getenv("TTSK_NO_GW") ? atoi(getenv("TTSK_NO_GW")) : 0;
    return !off;
}

static void gw_common(GwParams& G, const PassParams& P, const Source& S) {
    std::memset(&G, 0, sizeof(G));
    G.nnz = P.nnz;
    G.n_mu = P.n_mu;
    G.keyid = P.keyid;
    G.recs = P.recs;
    G.offs = P.offs;
    G.r = S.r;
    G.rank_min = S.rank_min;
    G.seed = S.seed;
    G.pitch_t = 1;
}

// first mode (A absent, B generated) or last mode in sorted order (A generated, B absent)
int try_launch_gw_direct(ttsk_ctx* ctx, PassParams& P, cudaStream_t st, bool* used) {
    *used = false;
    if (!gw_enabled() || !P.recs || P.rec_words != 8 || !P.keyid) return TTSK_OK;
    const bool first = P.A.kind == SRC_NONE && P.B.kind == SRC_GAUSS;
    const bool last = P.A.kind == SRC_GAUSS && P.B.kind == SRC_NONE;
    if (!first && !last) return TTSK_OK;
    const Source& S = first ? P.B : P.A;
    if (S.r < 1 || S.r > 64) return TTSK_OK;
    GwParams G;
    gw_common(G, P, S);
    if (!gw_strides(S, G.smul_g)) return TTSK_OK;
    if (gw_smem(GW_SEG_DIRECT, S.r, 0, 1) > 227 * 1024) return TTSK_OK;
    G.out = P.psi;
    if (first) { G.key_stride = P.rB; G.col_stride = 1; }       // Psi_0 is (1, n, rB)
    else { G.key_stride = 1; G.col_stride = P.n_mu; }           // Psi_{d-1} is (rA, n, 1)
    TTSK_TRY((gw_launch<GW_SEG_DIRECT, 1, 1, false>(ctx, G, st)));
    *used = true;
    return TTSK_OK;
}

// last mode without bucketing: T[i_mu, :] in shared memory
int try_launch_gw_flat(ttsk_ctx* ctx, PassParams& P, cudaStream_t st, bool* used) {
    *used = false;
    if (!gw_enabled() || !P.recs || P.rec_words != 8) return TTSK_OK;
    if (P.A.kind != SRC_GAUSS || P.B.kind != SRC_NONE || P.sg_mode < 0 || P.sg_mode >= 6) return TTSK_OK;
    if (P.nnz < 65536 || P.A.r < 1 || P.A.r > 64) return TTSK_OK;
    GwParams G;
    gw_common(G, P, P.A);
    if (!gw_strides(P.A, G.smul_g)) return TTSK_OK;
    G.smul_s[P.sg_mode] = 1;
    G.S_rows = (int)P.n_mu;
    G.pitch_t = gw_pitch_t(P.A.r, 0);
    if (P.n_mu > 4096 || gw_smem(GW_FLAT_T, P.A.r, G.S_rows, G.pitch_t) > 227 * 1024) return TTSK_OK;
    G.keyid = nullptr;
    G.out = P.psi;
    G.key_stride = 1;        // Psi_{d-1} is (rA, n, 1)
    G.col_stride = P.n_mu;
    TTSK_TRY((gw_launch<GW_FLAT_T, 1, 1, false>(ctx, G, st)));
    ctx->sg_passes++;
    *used = true;
    return TTSK_OK;
}

template <int MI, int NJ, bool HAS_X>
static int try_gw_seg_t(ttsk_ctx* ctx, PassParams& P, cudaStream_t st, bool* used) {
    GwParams G;
    gw_common(G, P, P.A);
    if (!gw_strides(P.A, G.smul_g)) return TTSK_OK;
    const long long S_rows = P.B.span_bytes / (8 * P.B.row_stride);
    if (S_rows < 1 || S_rows > 4096 || P.nnz < 512 * P.n_mu) return TTSK_OK;  // long segments only
    long long sb[6], sx[6];
    if (!gw_strides(P.B, sb)) return TTSK_OK;
    if (sb[P.sg_mode] != 0) return TTSK_OK;  // B must not depend on the pass mode
    if (HAS_X) {
        if (!gw_strides(P.X, sx)) return TTSK_OK;
        for (int m = 0; m < 6; m++)
            if (sx[m] != (m == P.sg_mode ? S_rows : sb[m])) return TTSK_OK;  // X row = key * S + B row
        if (P.X.span_bytes != P.n_mu * S_rows * 8 * P.X.row_stride) return TTSK_OK;
    }
    for (int m = 0; m < 6; m++) G.smul_s[m] = sb[m];
    G.S_rows = (int)S_rows;
    G.pitch_t = gw_pitch_t(P.A.r, MI);
    if (gw_smem(GW_SEG_T, P.A.r, G.S_rows, G.pitch_t) > 227 * 1024) return TTSK_OK;
    G.Btab = P.B.base; G.b_rs = P.B.row_stride; G.rB = P.rB;
    G.Xtab = HAS_X ? P.X.base : nullptr; G.x_rs = HAS_X ? P.X.row_stride : 0; G.rX = HAS_X ? P.rX : 0;
    G.psi = P.psi;
    G.omega = P.omega;
    TTSK_TRY((gw_launch<GW_SEG_T, MI, NJ, HAS_X>(ctx, G, st)));
    ctx->sg_passes++;
    *used = true;
    return TTSK_OK;
}

// segment-GEMM form: A generated, B (and X) small prefix tables
int try_launch_gw_seg(ttsk_ctx* ctx, PassParams& P, bool has_x, cudaStream_t st, bool* used) {
    *used = false;
    if (!gw_enabled() || !P.recs || P.rec_words != 8 || !P.keyid || !P.offs) return TTSK_OK;
    if (P.A.kind != SRC_GAUSS || P.B.kind != SRC_TABLE || P.B.col_stride != 1 || P.sg_mode < 0 || P.sg_mode >= 6) return TTSK_OK;
    if (has_x && (P.X.kind != SRC_TABLE || P.X.col_stride != 1)) return TTSK_OK;
    if (P.A.r < 1 || P.A.r > 64 || P.n_mu >= ((long long)1 << 31) - 1) return TTSK_OK;
    const int mi = (P.rA + 7) / 8;
    const int nj = (std::max(P.rB, has_x ? P.rX : 1) + 7) / 8;
    if (mi > 8 || nj > 8) return TTSK_OK;
    const int MIr = mi <= 1 ? 1 : (mi <= 3 ? 3 : (mi <= 5 ? 5 : 8));
    const int NJr = nj <= 1 ? 1 : (nj <= 3 ? 3 : (nj <= 5 ? 5 : 8));
    if (MIr * NJr > 25) return TTSK_OK;  // accumulators of the segment GEMM must stay in registers (no spills)
#define TTSK_GW(MI_, NJ_)                                                              \
    case MI_ * 10 + NJ_:                                                               \
        return has_x ? try_gw_seg_t<MI_, NJ_, true>(ctx, P, st, used) : try_gw_seg_t<MI_, NJ_, false>(ctx, P, st, used)
    switch (MIr * 10 + NJr) {
        TTSK_GW(1, 1); TTSK_GW(1, 3); TTSK_GW(1, 5); TTSK_GW(1, 8);
        TTSK_GW(3, 1); TTSK_GW(3, 3); TTSK_GW(3, 5); TTSK_GW(3, 8);
        TTSK_GW(5, 1); TTSK_GW(5, 3); TTSK_GW(5, 5);
        TTSK_GW(8, 1); TTSK_GW(8, 3);
    }
#undef TTSK_GW
    return TTSK_OK;
}

}  // namespace ttsk

