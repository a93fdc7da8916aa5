// Gather-only mode pass: both factors of Psi_mu[:, j, :] = sum_{p: i_mu(p) = j} v_p A[a_p, :] (x) B[b_p, :] are rows of
// prefix tables (sketch_psi_sparse of the reference, sparse_sketch.py:47-68, with both DRM contractions tabulated;
// BASELINE config 4's mode 1: A = L_0, 1e4 x 20, L2 resident; B = R_1, 5e6 x 40 = 1.6 GB, one random 320-byte row
// per nonzero -- the pass is bound by that gather).
//
// Round 1/2's general kernel (sparse_pass_kernel, NP = 8) ran this pass at 4.0 TB/s of DRAM traffic, 14.2 ms: per
// nonzero it chased sorted word -> packed record (a random 32-byte read = one 128-byte DRAM access) -> table rows,
// and moved every row with twenty 16-byte cp.async instructions.  Here
//   * the payload partition (partition_kernel<LEVEL, true>) delivers, in sorted order, everything a nonzero needs:
//     one word (key | row of A | row of B) and its value, streamed coalesced -- no record reads at all;
//   * a row is ONE bulk copy (cp.async.bulk.shared::cluster.global, completion counted in bytes on the stage's
//     mbarrier -- the TMA engine's linear mode) for the long rows of B;
//   * eight producer warps share the rows of every tile (see the producer branch); six stages of 64 rows keep ~190 KB of
//     row copies in flight per SM.
// Consumers are the eight MMA warps of the general kernel: each owns 8 rows of a tile, scales the A fragment by the
// value, accumulates (rA x rB) on mma.sync.m8n8k4.f64 and flushes with FP64 atomics when the key changes.
#include "ttsk_sparse_pass.cuh"

namespace ttsk {

constexpr int kGtTN = 96, kGtConsumers = 12, kGtProducers = 4, kGtThreads = 32 * (kGtConsumers + kGtProducers), kGtMaxStages = 8;

struct GtParams {
    long long nnz, n_mu;
    const unsigned long long* words;  // key << kshift | row of A << bshift | row of B, sorted by key
    const double* vals;               // value of the nonzero, same order
    const double* A;
    const double* B;
    long long a_stride, b_stride;     // doubles between rows
    int rA, rB, pa, pb;               // columns; shared-memory pitches (doubles)
    int kshift, bshift;
    unsigned long long amask, bmask;
    double* psi;                      // (rA, n_mu, rB)
    long long item_len;               // sorted positions per CTA (multiple of the tile height)
    int nstages, stage_bytes, off_a, off_b;
    int debug;  // profiling builds only (TTSK_ABLATE): 1 no MMA, 2 no B copies, 4 no A copies
};

__device__ __forceinline__ void gt_expect_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gt_bulk_row(unsigned dst, const void* src, unsigned bytes, unsigned long long* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(smem_addr(b))
                 : "memory");
}

template <int MI, int NJ>
__global__ void __launch_bounds__(kGtThreads, 1) gt_kernel(const GtParams P) {
    extern __shared__ __align__(128) unsigned char gt_smem[];
    unsigned long long* full = reinterpret_cast<unsigned long long*>(gt_smem);  // B rows of the stage have landed
    unsigned long long* empty = full + kGtMaxStages;                            // the consumers are done with the stage
    unsigned long long* ids = empty + kGtMaxStages;                             // keys / values / A row numbers are written
    unsigned char* stages = gt_smem + 256;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NST = P.nstages;
    for (int i = tid; i < NST * P.stage_bytes / 8; i += kGtThreads) reinterpret_cast<double*>(stages)[i] = 0.0;
    if (tid == 0) {
        for (int s = 0; s < NST; s++) {
            mbar_init(&full[s], kGtProducers);   // one arrive.expect_tx per producer warp (+ the bytes of its bulk copies)
            mbar_init(&ids[s], kGtProducers);    // one plain arrive per producer warp
            mbar_init(&empty[s], kGtConsumers);  // one arrive per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the zero fill precedes the bulk copies into the stages
    __syncthreads();
    const long long lo = (long long)blockIdx.x * P.item_len;
    const long long hi = (lo + P.item_len < P.nnz) ? lo + P.item_len : P.nnz;
    const long long n_tiles = hi > lo ? (hi - lo + kGtTN - 1) / kGtTN : 0;
    constexpr int kOffKey = 16, kOffRa = 16 + kGtTN * 4, kOffVal = 16 + kGtTN * 8;

    if (warp >= kGtConsumers) {
        // =========================================================== producers: warp pw stages rows [8 pw, 8 pw + 8) of
        // every tile.  A bulk copy is a warp-uniform instruction (UBLKCP: the compiler serialises the lanes of a
        // divergent cp.async.bulk through ELECT / R2UR, ~60 cycles per copy and warp), so the B rows -- the HBM stream --
        // are spread over all producer warps, one lane per row.  The short A rows (an L2-resident table) are NOT
        // staged: 16-byte cp.async chunks sustained only ~0.6 chunks per cycle and SM here (5.8 ms for the 1e9 chunks
        // of C4's mode 1); the consumers read their A fragments straight from L2, one tile ahead.
        constexpr int RP = kGtTN / kGtProducers;  // rows per producer warp
        const int pw = warp - kGtConsumers;
        unsigned long long w0 = 0, w1 = 0, w2 = 0;
        double v0 = 0.0, v1 = 0.0, v2 = 0.0;
        auto fetch = [&](long long t, unsigned long long& w, double& v) {
            const long long pos = lo + t * kGtTN + pw * RP + lane;
            const bool in = lane < RP && t < n_tiles && pos < hi;
            w = in ? __ldcs(P.words + pos) : 0ull;
            v = in ? __ldcs(P.vals + pos) : 0.0;
        };
        fetch(0, w0, v0);
        fetch(1, w1, v1);
        const unsigned b_bytes = (unsigned)P.rB * 8u;
        for (long long t = 0; t <= n_tiles; t++) {  // tile n_tiles is the end-of-stream marker
            fetch(t + 2, w2, v2);
            const int s = (int)(t % NST);
            const long long round = t / NST;
            if (round > 0) mbar_wait(&empty[s], (unsigned)((round - 1) & 1));
            unsigned char* st = stages + (size_t)s * P.stage_bytes;
            const int n_rows = t == n_tiles ? 0 : (int)((lo + (t + 1) * kGtTN <= hi) ? kGtTN : hi - (lo + t * kGtTN));
            int mine = n_rows - pw * RP;  // valid rows among this warp's
            mine = mine < 0 ? 0 : (mine > RP ? RP : mine);
            if (lane < RP) {
                const int i = pw * RP + lane;
                reinterpret_cast<int*>(st + kOffKey)[i] = lane < mine ? (int)(w0 >> P.kshift) : -1;
                reinterpret_cast<int*>(st + kOffRa)[i] = lane < mine ? (int)((w0 >> P.bshift) & P.amask) : 0;
                reinterpret_cast<double*>(st + kOffVal)[i] = lane < mine ? v0 : 0.0;
            }
            if (pw == 0 && lane == 0) *reinterpret_cast<int*>(st) = n_rows;
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&ids[s]);
                gt_expect_tx(&full[s], (P.debug & 2) ? 0u : (unsigned)mine * b_bytes);
            }
            __syncwarp();
            if (lane < mine && !(P.debug & 2)) {
                const unsigned long long rb = w0 & P.bmask;
                gt_bulk_row(smem_addr(st + P.off_b + (size_t)(pw * RP + lane) * P.pb * 8), P.B + rb * P.b_stride, b_bytes, &full[s]);
            }
            w0 = w1; v0 = v1;
            w1 = w2; v1 = v2;
        }
        return;
    }

    // =============================================================== consumers
    const int g = lane >> 2, q = lane & 3;
    constexpr int RG = kGtTN / kGtConsumers, kCh = RG / 4;
    const int row0 = warp * RG;
    const int PB = P.pb;
    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NJ; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    auto flush_psi = [&](long long key) {
        double* base = P.psi + key * P.rB;
        const long long row_pitch = P.n_mu * P.rB;
#pragma unroll
        for (int i = 0; i < MI; i++)
#pragma unroll
            for (int j = 0; j < NJ; j++) {
                const int row = 8 * i + g, col = 8 * j + 2 * q;
                if (row < P.rA) {
                    double* dst = base + (long long)row * row_pitch + col;
                    if (col < P.rB && acc[i][j][0] != 0.0) atomicAdd(dst, acc[i][j][0]);
                    if (col + 1 < P.rB && acc[i][j][1] != 0.0) atomicAdd(dst + 1, acc[i][j][1]);
                }
                acc[i][j][0] = acc[i][j][1] = 0.0;
            }
    };
    // A fragments of this warp's rows of a stage: element (row rbase + q, column 8 i + g) of the m8n8k4 A operand
    double a_cur[kCh][MI], a_nxt[kCh][MI];
    auto load_a = [&](const unsigned char* st, double (*a)[MI]) {
        const int* s_ra = reinterpret_cast<const int*>(st + kOffRa);
#pragma unroll
        for (int ch = 0; ch < kCh; ch++) {
            const double* row = P.A + (long long)s_ra[row0 + 4 * ch + q] * P.a_stride;
#pragma unroll
            for (int i = 0; i < MI; i++) a[ch][i] = (8 * i + g < P.rA && !(P.debug & 4)) ? __ldg(row + 8 * i + g) : 0.0;
        }
    };
    int cur_key = -1, s = 0;
    unsigned ph = 0;
    mbar_wait(&ids[0], 0u);
    load_a(stages, a_nxt);
    while (true) {
        const unsigned char* st = stages + (size_t)s * P.stage_bytes;
        const int n_rows = *reinterpret_cast<const int*>(st);  // written before the stage's `ids` barrier completed
        if (n_rows == 0) break;
#pragma unroll
        for (int ch = 0; ch < kCh; ch++)
#pragma unroll
            for (int i = 0; i < MI; i++) a_cur[ch][i] = a_nxt[ch][i];
        {   // the next tile's A fragments are requested before this tile's B rows are waited for
            const int sn = (s + 1 == NST) ? 0 : s + 1;
            mbar_wait(&ids[sn], (s + 1 == NST) ? (ph ^ 1u) : ph);
            load_a(stages + (size_t)sn * P.stage_bytes, a_nxt);
        }
        mbar_wait(&full[s], ph);
        if (row0 < n_rows && !(P.debug & 1)) {
            const int* s_key = reinterpret_cast<const int*>(st + kOffKey);
            const double* s_val = reinterpret_cast<const double*>(st + kOffVal);
            const double* Bt = reinterpret_cast<const double*>(st + P.off_b);
#pragma unroll
            for (int ch = 0; ch < kCh; ch++) {
                const int rbase = row0 + 4 * ch;
                if (rbase >= n_rows) break;  // warp-uniform
                const int p = rbase + q;
                const double v = s_val[p];  // 0 past the tile
                const int key = s_key[p];   // -1 past the tile
                double a[MI], b[NJ];
#pragma unroll
                for (int i = 0; i < MI; i++) a[i] = a_cur[ch][i] * v;
#pragma unroll
                for (int j = 0; j < NJ; j++) b[j] = Bt[p * PB + 8 * j + g];
                if (__all_sync(0xffffffffu, key == cur_key)) {
#pragma unroll
                    for (int i = 0; i < MI; i++)
#pragma unroll
                        for (int j = 0; j < NJ; j++) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
                } else {
                    // the chunk starts a new segment or straddles boundaries: one MMA round per run of equal keys
                    int start = 0;
                    while (start < 4) {
                        const int kcur = __shfl_sync(0xffffffffu, key, start);  // lane `start` holds row rbase + start
                        if (kcur < 0) break;
                        if (kcur != cur_key) {
                            if (cur_key >= 0) flush_psi(cur_key);
                            cur_key = kcur;
                        }
                        const bool mine = (q >= start) && (key == kcur);
                        const unsigned diff = __ballot_sync(0xffffffffu, (q > start) && (key != kcur)) & 0xFu;
#pragma unroll
                        for (int i = 0; i < MI; i++) {
                            const double am = mine ? a[i] : 0.0;
#pragma unroll
                            for (int j = 0; j < NJ; j++) dmma(acc[i][j][0], acc[i][j][1], am, b[j]);
                        }
                        start = diff ? (__ffs(diff) - 1) : 4;
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (++s == NST) { s = 0; ph ^= 1u; }
    }
    if (cur_key >= 0) flush_psi(cur_key);
}

static int bits_for(long long n) {  // bits that hold 0 .. n - 1
    int b = 0;
    while (((long long)1 << b) < n) b++;
    return b;
}

// Can the pass described by P take the payload form?  Fills the sweep-1 plan and the word layout.
bool gt_plan(const PassParams& P, bool has_x, PayPlan* pay, int* kshift) {
    static const int off = getenv("TTSK_NO_GT") ? atoi(getenv("TTSK_NO_GT")) : 0;
    if (off || has_x || P.A.kind != SRC_TABLE || P.B.kind != SRC_TABLE || !P.recs || P.rec_words != 8) return false;
    if (P.rA > 64 || P.rB > 64 || (P.rA & 1) || (P.rB & 1) || P.nnz < 65536 || P.n_mu > 16384) return false;
    if (P.A.col_stride != 1 || P.B.col_stride != 1 || (P.A.row_stride & 1) || (P.B.row_stride & 1)) return false;
    if (((uintptr_t)P.A.base & 15) || ((uintptr_t)P.B.base & 15) || P.A.span_bytes <= 0 || P.B.span_bytes <= 0) return false;
    if (P.sg_mode < 0 || P.sg_mode >= 6) return false;
    std::memset(pay, 0, sizeof(*pay));
    const Source* src[2] = {&P.A, &P.B};
    for (int k = 0; k < 2; k++)
        for (int i = 0; i < src[k]->k; i++) {
            if (src[k]->modes[i] >= 6 || src[k]->strides[i] < 0 || src[k]->strides[i] >= ((long long)1 << 32)) return false;
            (k == 0 ? pay->smul_a : pay->smul_b)[src[k]->modes[i]] += (unsigned)src[k]->strides[i];
        }
    const int abits = bits_for(P.A.span_bytes / (8 * P.A.row_stride)), bbits = bits_for(P.B.span_bytes / (8 * P.B.row_stride));
    const int kbits = bits_for(P.n_mu);
    if (abits + bbits + kbits > 63) return false;
    pay->recs = P.recs;
    pay->key_word = 2 + P.sg_mode;
    pay->bshift = bbits;
    *kshift = abits + bbits;
    return true;
}

template <int MI, int NJ>
static int launch_gt_t(ttsk_ctx* ctx, const PassParams& P, const unsigned long long* words, const unsigned long long* vals,
                       const PayPlan& pay, int kshift, cudaStream_t st) {
    GtParams G;
    std::memset(&G, 0, sizeof(G));
    G.nnz = P.nnz; G.n_mu = P.n_mu;
    G.words = words; G.vals = reinterpret_cast<const double*>(vals);
    G.A = P.A.base; G.B = P.B.base; G.a_stride = P.A.row_stride; G.b_stride = P.B.row_stride;
    G.rA = P.rA; G.rB = P.rB; G.pa = tile_pitch(MI);
    // B fragment loads: lane (g, q) reads row q, column 8 j + g; a half warp (g < 4, all q) is conflict free when the
    // row pitch is 8 or 24 words (mod 32), i.e. 4 or 12 doubles (mod 16) -- tile_pitch's 8 (mod 16) puts rows q and
    // q + 2 on the same banks (ncu: 38 % of this kernel's shared wavefronts were conflicts)
    G.pb = 8 * NJ;
    while (G.pb % 16 != 4 && G.pb % 16 != 12) G.pb += 2;
    G.kshift = kshift; G.bshift = pay.bshift;
    G.bmask = (1ull << pay.bshift) - 1ull;
    G.amask = (1ull << (kshift - pay.bshift)) - 1ull;
    G.psi = P.psi;
    G.off_a = 0;  // A rows are not staged
    G.off_b = 16 + kGtTN * 4 + kGtTN * 4 + kGtTN * 8;
    G.stage_bytes = (int)align_up(G.off_b + kGtTN * G.pb * 8, 128);
    static const int st_env = getenv("TTSK_GT_STAGES") ? atoi(getenv("TTSK_GT_STAGES")) : 0;
    static const int cta_env = getenv("TTSK_GT_CTAS") ? atoi(getenv("TTSK_GT_CTAS")) : 0;
    int per_sm = cta_env > 0 ? cta_env : 1;
    const int budget = (per_sm >= 2 ? 113 : 226) * 1024 - 256 - 1024;
    int nst = budget / G.stage_bytes;
    if (nst < 2 && per_sm >= 2) { per_sm = 1; nst = (226 * 1024 - 256) / G.stage_bytes; }
    if (nst > kGtMaxStages) nst = kGtMaxStages;
    if (st_env > 1 && st_env <= nst) nst = st_env;
    TTSK_ARG(nst >= 2, "gather pass: a stage does not fit shared memory");
    G.nstages = nst;
    const size_t smem = 256 + (size_t)nst * G.stage_bytes;
    auto kern = gt_kernel<MI, NJ>;
    TTSK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TTSK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    long long grid = (long long)ctx->sm_count * per_sm;
    long long item = (P.nnz + grid - 1) / grid;
    item = (item + kGtTN - 1) / kGtTN * kGtTN;
    grid = (P.nnz + item - 1) / item;
    G.item_len = item;
    G.debug = ablate_bits();
    if (getenv("TTSK_DEBUG"))
        fprintf(stderr, "[ttsk] gather pass MI=%d NJ=%d stages=%d x %d B smem=%zu ctas/sm=%d grid=%lld kshift=%d bshift=%d\n", MI, NJ,
                nst, G.stage_bytes, smem, per_sm, grid, kshift, pay.bshift);
    kern<<<(unsigned)grid, kGtThreads, smem, st>>>(G);
    TTSK_LAUNCHED(ctx);
    ctx->sg_passes++;  // counted with the other specialised pass forms (ttsk_sg_pass_count): tests assert the form ran
    return TTSK_OK;
}

int launch_gt(ttsk_ctx* ctx, const PassParams& P, const unsigned long long* words, const unsigned long long* vals,
              const PayPlan& pay, int kshift, cudaStream_t st) {
    const int mi = (P.rA + 7) / 8, nj = (P.rB + 7) / 8;
    const int MIr = mi <= 1 ? 1 : (mi <= 3 ? 3 : (mi <= 5 ? 5 : 8));
    const int NJr = nj <= 1 ? 1 : (nj <= 3 ? 3 : (nj <= 5 ? 5 : 8));
#define TTSK_GT(MI_, NJ_) return launch_gt_t<MI_, NJ_>(ctx, P, words, vals, pay, kshift, st)
    switch (MIr * 10 + NJr) {
        case 11: TTSK_GT(1, 1);
        case 13: TTSK_GT(1, 3);
        case 15: TTSK_GT(1, 5);
        case 18: TTSK_GT(1, 8);
        case 31: TTSK_GT(3, 1);
        case 33: TTSK_GT(3, 3);
        case 35: TTSK_GT(3, 5);
        case 38: TTSK_GT(3, 8);
        case 51: TTSK_GT(5, 1);
        case 53: TTSK_GT(5, 3);
        case 55: TTSK_GT(5, 5);
        case 58: TTSK_GT(5, 8);
        case 81: TTSK_GT(8, 1);
        case 83: TTSK_GT(8, 3);
        case 85: TTSK_GT(8, 5);
        case 88: TTSK_GT(8, 8);
    }
#undef TTSK_GT
    set_error("gather pass: no kernel variant");
    return TTSK_E_ARG;
}

}  // namespace ttsk
