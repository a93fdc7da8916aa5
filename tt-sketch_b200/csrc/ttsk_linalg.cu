// Small dense linear algebra for TT assembly and the orthogonalisation step.
// Replaces scipy.linalg.lstsq (LAPACK gelsd, SVD based) in tt_sketch/utils.py:98-109 as used by
// assemble_sketched_tt (tt_sketch/sketch.py:400-443) and orth_step
// (tt_sketch/sketch_dispatch.py:160-174), and scipy.linalg.qr(mode="economic") there (:172).
//
// The matrices are tiny (Omega is at most ~64 x 128; the QR panel is (r*n) x r), so each
// factorisation runs in ONE thread block out of L2-resident global memory: the work is bound
// by dependency latency, not by bandwidth or flops.  The pseudo-inverse is formed explicitly
// (one-sided Jacobi SVD, singular values below rcond*s_max dropped like gelsd) and applied to
// the (r*n) right-hand sides with ttsk_gemm.
#include <cfloat>
#include <cstdlib>

#include "ttsk_common.cuh"

namespace ttsk {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// W: (rows, c) stored COLUMN-major (W[j*rows + i]); V: (c, c) column-major.
// One-sided Jacobi with a round-robin tournament: in each round c/2 disjoint column pairs are
// rotated concurrently, one warp per pair.
// With `pinv == nullptr` the factors are written instead (ttsk_svd): singular values in descending order to
// svd_S (c), U (m, c) row-major to svd_U -- its columns multiplied by the singular values when u_times_s -- and
// V^T (c, n) row-major to svd_Vt.
__global__ void __launch_bounds__(1024) jacobi_pinv_kernel(const double* __restrict__ A, int m, int n, double rcond,
                                                          double* __restrict__ W, double* __restrict__ V,
                                                          double* __restrict__ sig, double* __restrict__ pinv,
                                                          double* __restrict__ svd_U, double* __restrict__ svd_S,
                                                          double* __restrict__ svd_Vt, int u_times_s,
                                                          int* __restrict__ perm) {
    const bool tall = m >= n;
    const int rows = tall ? m : n, c = tall ? n : m;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    __shared__ int s_rot;
    __shared__ double s_max;
    // W = A (tall) or A^T
    for (int e = tid; e < rows * c; e += blockDim.x) {
        const int j = e / rows, i = e - j * rows;
        W[e] = tall ? A[(long long)i * n + j] : A[(long long)j * n + i];
    }
    for (int e = tid; e < c * c; e += blockDim.x) V[e] = (e / c == e % c) ? 1.0 : 0.0;
    __syncthreads();
    const int cc = (c + 1) & ~1;  // players in the tournament (one dummy if c is odd)
    for (int sweep = 0; sweep < 60; sweep++) {
        if (tid == 0) s_rot = 0;
        __syncthreads();
        for (int round = 0; round < cc - 1; round++) {
            for (int k = warp; k < cc / 2; k += nwarps) {
                // circle method: player cc-1 fixed, others rotate
                int p = (k == 0) ? cc - 1 : (round + k) % (cc - 1);
                int q = (round + cc - 1 - k) % (cc - 1);
                if (p > q) { const int t = p; p = q; q = t; }
                if (q >= c) continue;  // dummy
                double* wp = W + (long long)p * rows;
                double* wq = W + (long long)q * rows;
                double a = 0.0, b = 0.0, g = 0.0;
                for (int i = lane; i < rows; i += 32) {
                    const double x = wp[i], y = wq[i];
                    a = fma(x, x, a); b = fma(y, y, b); g = fma(x, y, g);
                }
                a = warp_sum(a); b = warp_sum(b); g = warp_sum(g);
                if (fabs(g) > 1e-15 * sqrt(a * b) && g != 0.0) {
                    const double zeta = (b - a) / (2.0 * g);
                    const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
                    for (int i = lane; i < rows; i += 32) {
                        const double x = wp[i], y = wq[i];
                        wp[i] = cs * x - sn * y;
                        wq[i] = sn * x + cs * y;
                    }
                    double* vp = V + (long long)p * c;
                    double* vq = V + (long long)q * c;
                    for (int i = lane; i < c; i += 32) {
                        const double x = vp[i], y = vq[i];
                        vp[i] = cs * x - sn * y;
                        vq[i] = sn * x + cs * y;
                    }
                    if (lane == 0) s_rot = 1;
                }
            }
            __syncthreads();
        }
        const int rot = s_rot;
        __syncthreads();
        if (!rot) break;
    }
    // singular values = column norms
    if (tid == 0) s_max = 0.0;
    __syncthreads();
    for (int j = warp; j < c; j += nwarps) {
        double a = 0.0;
        for (int i = lane; i < rows; i += 32) a = fma(W[(long long)j * rows + i], W[(long long)j * rows + i], a);
        a = warp_sum(a);
        if (lane == 0) sig[j] = sqrt(a);
    }
    __syncthreads();
    if (tid == 0) {
        double mx = 0.0;
        for (int j = 0; j < c; j++) mx = fmax(mx, sig[j]);
        s_max = mx;
    }
    __syncthreads();
    if (pinv == nullptr) {
        if (tid == 0) {  // descending order of the singular values (c <= 256: a selection sort by one thread)
            for (int j = 0; j < c; j++) perm[j] = j;
            for (int a = 0; a < c; a++) {
                int best = a;
                for (int b = a + 1; b < c; b++)
                    if (sig[perm[b]] > sig[perm[best]]) best = b;
                const int t = perm[a]; perm[a] = perm[best]; perm[best] = t;
            }
        }
        __syncthreads();
        for (int j = tid; j < c; j += blockDim.x) svd_S[j] = sig[perm[j]];
        // tall: A = (W / s) S V^T;  wide: A^T = (W / s) S V^T  =>  A = V S (W / s)^T
        for (int e = tid; e < m * c; e += blockDim.x) {
            const int i = e / c, jj = e - i * c, j = perm[jj];
            const double sj = sig[j];
            double u = tall ? (sj > 0.0 ? W[(long long)j * rows + i] / sj : 0.0) : V[(long long)j * c + i];
            if (u_times_s) u = tall ? W[(long long)j * rows + i] : u * sj;
            svd_U[e] = u;
        }
        for (long long e = tid; e < (long long)c * n; e += blockDim.x) {
            const int jj = (int)(e / n), i = (int)(e - (long long)jj * n), j = perm[jj];
            const double sj = sig[j];
            svd_Vt[e] = tall ? V[(long long)j * c + i] : (sj > 0.0 ? W[(long long)j * rows + i] / sj : 0.0);
        }
        return;
    }
    const double cut = (rcond < 0.0 ? DBL_EPSILON : rcond) * s_max;
    // pinv (n, m) row-major
    for (int e = tid; e < n * m; e += blockDim.x) {
        const int i = e / m, k = e - i * m;  // pinv[i][k]
        double s = 0.0;
        for (int j = 0; j < c; j++) {
            const double sj = sig[j];
            if (sj > cut) {
                // tall:  pinv = V S^-2 W^T   -> V[i][j] W[k][j];   wide: pinv = W S^-2 V^T -> W[i][j] V[k][j]
                const double x = tall ? V[(long long)j * c + i] * W[(long long)j * rows + k]
                                      : W[(long long)j * rows + i] * V[(long long)j * c + k];
                s += x / (sj * sj);
            }
        }
        pinv[e] = s;
    }
}

// Householder QR of A (m, n) row-major in place, Q (economic) returned in A.  LAPACK dgeqr2 /
// dorg2r conventions (beta = -sign(alpha) * norm), so Q matches scipy.linalg.qr.
__global__ void __launch_bounds__(1024) householder_q_kernel(double* __restrict__ A, long long m, int n,
                                                            double* __restrict__ tau_g) {
    __shared__ double s_red[32][33];
    __shared__ double s_w[64];
    __shared__ double s_scal[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int j = 0; j < n; j++) {
        // ---- dlarfg on column j, rows j..m-1
        double part = 0.0;
        for (long long i = j + 1 + tid; i < m; i += blockDim.x) {
            const double x = A[i * n + j];
            part = fma(x, x, part);
        }
        part = warp_sum(part);
        if (lane == 0) s_red[0][warp] = part;
        __syncthreads();
        if (tid == 0) {
            double ss = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); w++) ss += s_red[0][w];
            const double xnorm = sqrt(ss);
            const double alpha = A[(long long)j * n + j];
            double tau = 0.0, beta = alpha, scale = 0.0;
            if (xnorm != 0.0) {
                beta = -copysign(hypot(alpha, xnorm), alpha);
                tau = (beta - alpha) / beta;
                scale = 1.0 / (alpha - beta);
            }
            s_scal[0] = tau; s_scal[1] = beta; s_scal[2] = scale;
            tau_g[j] = tau;
            A[(long long)j * n + j] = beta;
        }
        __syncthreads();
        const double tau = s_scal[0], scale = s_scal[2];
        if (tau != 0.0) {
            for (long long i = j + 1 + tid; i < m; i += blockDim.x) A[i * n + j] *= scale;
            __syncthreads();
            // ---- apply H_j = I - tau v v^T to columns j+1..n-1 (v_j = 1, v_i = A[i][j])
            for (int c0 = j + 1; c0 < n; c0 += 32) {
                const int col = c0 + lane;
                double w = 0.0;
                if (col < n)
                    for (long long i = j + warp; i < m; i += 32) {
                        const double v = (i == j) ? 1.0 : A[i * n + j];
                        w = fma(v, A[i * n + col], w);
                    }
                s_red[warp][lane] = w;
                __syncthreads();
                if (warp == 0) {
                    double t = 0.0;
                    for (int r = 0; r < 32; r++) t += s_red[r][lane];
                    s_w[lane] = t * tau;
                }
                __syncthreads();
                if (col < n) {
                    const double tw = s_w[lane];
                    for (long long i = j + warp; i < m; i += 32) {
                        const double v = (i == j) ? 1.0 : A[i * n + j];
                        A[i * n + col] -= v * tw;
                    }
                }
                __syncthreads();
            }
        }
        __syncthreads();
    }
    // ---- dorg2r: Q = H_0 ... H_{n-1} applied to the first n columns of I, built backwards
    for (int j = n - 1; j >= 0; j--) {
        const double tau = tau_g[j];
        // columns j+1..n-1 of Q (already formed for rows >= j+1; row j of them is 0) get H_j applied
        for (int c0 = j + 1; c0 < n; c0 += 32) {
            const int col = c0 + lane;
            double w = 0.0;
            if (col < n)
                for (long long i = j + 1 + warp; i < m; i += 32) w = fma(A[i * n + j], A[i * n + col], w);
            s_red[warp][lane] = w;
            __syncthreads();
            if (warp == 0) {
                double t = 0.0;
                for (int r = 0; r < 32; r++) t += s_red[r][lane];
                s_w[lane] = t * tau;
            }
            __syncthreads();
            if (col < n) {
                const double tw = s_w[lane];
                if (warp == 0) A[(long long)j * n + col] = -tw;  // row j: 0 - 1*tw
                for (long long i = j + 1 + warp; i < m; i += 32) A[i * n + col] -= A[i * n + j] * tw;
            }
            __syncthreads();
        }
        // column j itself: Q[:, j] = e_j - tau v
        for (long long i = j + 1 + tid; i < m; i += blockDim.x) A[i * n + j] *= -tau;
        if (tid == 0) A[(long long)j * n + j] = 1.0 - tau;
        for (int r = tid; r < j; r += blockDim.x) A[(long long)r * n + j] = 0.0;
        __syncthreads();
    }
}

// ------------------------------------------------------------------ the same QR over the whole GPU
// Tall panels (the (r n) x r unfoldings of orth_step at n = 10^4 ... have 10^5 rows) on ONE SM stream the panel
// through a single CTA 2 n times.  Here every CTA owns a block of rows (L2-resident) and the n Householder steps run
// in lock step: per column one grid-wide reduction for the norm and one for the n - j - 1 products v^T A[:, c]
// (FP64 atomics into per-column slots, then a grid barrier), and one more per column when Q is formed backwards.
// The arithmetic is dgeqr2 / dorg2r's (same reflectors, same signs: Q matches scipy.linalg.qr), only the
// summation order of the reductions differs.  Launched cooperatively (all CTAs co-resident).
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& epoch, unsigned n_cta) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        epoch += n_cta;
        atomicAdd(counter, 1u);
        while (*reinterpret_cast<volatile unsigned*>(counter) < epoch) {}
        __threadfence();
    }
    __syncthreads();
}

constexpr int kQrThreads = 512;

// red: [n] squared norms, [n][n] forward products, [n][n] backward products (zero on entry); bar: zero on entry
__global__ void __launch_bounds__(kQrThreads, 1) householder_q_grid_kernel(double* __restrict__ A, long long m, int n,
                                                                         long long rows_per_cta, double* __restrict__ red,
                                                                         unsigned* __restrict__ bar) {
    __shared__ double s_part[kQrThreads / 32][65];
    __shared__ double s_tw[64];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kQrThreads / 32;
    const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = (r0 + rows_per_cta < m) ? r0 + rows_per_cta : m;
    double* nrm = red;
    double* fwd = red + n;
    double* bwd = fwd + (long long)n * n;
    double* taus = bwd + (long long)n * n;  // written redundantly by every CTA (same values)
    unsigned epoch = 0;
    const unsigned n_cta = gridDim.x;
    // products of column j with columns c in (j, n) over this CTA's rows i >= lo: lane = column (two columns per lane
    // for n <= 64), warps stride the rows; v_i = A[i][j] * scale (and 1 for the pivot row when unit_pivot)
    auto column_products = [&](int j, long long lo, double scale, bool unit_pivot, bool scale_in_place, double* out) {
        const int c0 = j + 1 + lane, c1 = j + 33 + lane;
        double w0 = 0.0, w1 = 0.0;
        for (long long i = (lo > r0 ? lo : r0) + warp; i < r1; i += nwarps) {
            double v = A[i * n + j];
            if (unit_pivot && i == j) v = 1.0;
            else {
                v *= scale;
                if (scale_in_place && lane == 0) A[i * n + j] = v;
            }
            if (c0 < n) w0 = fma(v, A[i * n + c0], w0);
            if (c1 < n) w1 = fma(v, A[i * n + c1], w1);
        }
        s_part[warp][lane] = w0;
        s_part[warp][32 + lane] = w1;
        __syncthreads();
        if (tid < 64 && j + 1 + tid < n) {
            double t = 0.0;
            for (int w = 0; w < nwarps; w++) t += s_part[w][tid];
            if (t != 0.0) atomicAdd(out + j + 1 + tid, t);
        }
        __syncthreads();
    };
    for (int j = 0; j < n; j++) {
        // ---- dlarfg on column j, rows j .. m-1
        double part = 0.0;
        for (long long i = (j + 1 > r0 ? j + 1 : r0) + tid; i < r1; i += kQrThreads) {
            const double x = A[i * n + j];
            part = fma(x, x, part);
        }
        part = warp_sum(part);
        if (lane == 0) s_part[warp][0] = part;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < nwarps; w++) t += s_part[w][0];
            if (t != 0.0) atomicAdd(nrm + j, t);
        }
        grid_barrier(bar, epoch, n_cta);
        const double xnorm = sqrt(__ldcg(nrm + j));
        const double alpha = __ldcg(A + (long long)j * n + j);  // (the pivot is not overwritten: R is not needed)
        double tau = 0.0, scale = 0.0;
        if (xnorm != 0.0) {
            const double beta = -copysign(hypot(alpha, xnorm), alpha);
            tau = (beta - alpha) / beta;
            scale = 1.0 / (alpha - beta);
        }
        if (blockIdx.x == 0 && tid == 0) taus[j] = tau;
        if (tau != 0.0) {
            // ---- w[c] = sum_i v_i A[i][c]; the column is scaled to the reflector on the way
            column_products(j, j, scale, true, true, fwd + (long long)j * n);
            grid_barrier(bar, epoch, n_cta);
            if (tid < 64) s_tw[tid] = (j + 1 + tid < n) ? tau * __ldcg(fwd + (long long)j * n + j + 1 + tid) : 0.0;
            __syncthreads();
            const int c0 = j + 1 + lane, c1 = j + 33 + lane;
            for (long long i = (j > r0 ? j : r0) + warp; i < r1; i += nwarps) {
                const double v = (i == j) ? 1.0 : A[i * n + j];
                if (c0 < n) A[i * n + c0] -= v * s_tw[lane];
                if (c1 < n) A[i * n + c1] -= v * s_tw[32 + lane];
            }
            __syncthreads();
        } else {
            grid_barrier(bar, epoch, n_cta);  // keep the barrier count uniform
        }
    }
    // ---- dorg2r: Q = H_0 ... H_{n-1} applied to the first n columns of I, built backwards
    grid_barrier(bar, epoch, n_cta);
    for (int j = n - 1; j >= 0; j--) {
        const double tau = __ldcg(taus + j);
        column_products(j, j + 1, 1.0, false, false, bwd + (long long)j * n);
        grid_barrier(bar, epoch, n_cta);
        if (tid < 64) s_tw[tid] = (j + 1 + tid < n) ? tau * __ldcg(bwd + (long long)j * n + j + 1 + tid) : 0.0;
        __syncthreads();
        const int c0 = j + 1 + lane, c1 = j + 33 + lane;
        if (j >= r0 && j < r1 && warp == 0) {  // row j: 0 - 1 * tw
            if (c0 < n) A[(long long)j * n + c0] = -s_tw[lane];
            if (c1 < n) A[(long long)j * n + c1] = -s_tw[32 + lane];
        }
        for (long long i = (j + 1 > r0 ? j + 1 : r0) + warp; i < r1; i += nwarps) {
            const double v = A[i * n + j];
            if (c0 < n) A[i * n + c0] -= v * s_tw[lane];
            if (c1 < n) A[i * n + c1] -= v * s_tw[32 + lane];
        }
        __syncthreads();
        // column j itself: Q[:, j] = e_j - tau v
        for (long long i = (j + 1 > r0 ? j + 1 : r0) + tid; i < r1; i += kQrThreads) A[i * n + j] *= -tau;
        if (j >= r0 && j < r1 && tid == 0) A[(long long)j * n + j] = 1.0 - tau;
        for (long long r = r0 + tid; r < j && r < r1; r += kQrThreads) A[r * n + j] = 0.0;
        __syncthreads();
    }
}

}  // namespace ttsk

extern "C" int ttsk_pinv(ttsk_ctx* ctx, const double* d_A, int m, int n, double rcond, double* d_pinv, void* stream) {
    TTSK_ARG(ctx != nullptr, "ctx is NULL");
    TTSK_ARG(m >= 1 && n >= 1 && m <= 4096 && n <= 4096 && (m <= 256 || n <= 256), "pinv: min(m, n) must be <= 256");
    TTSK_ARG(d_A && d_pinv, "NULL pointer");
    const int rows = m >= n ? m : n, c = m >= n ? n : m;
    const int64_t need = ((int64_t)rows * c + (int64_t)c * c + c) * 8 + 1024;
    TTSK_TRY(ctx->ws_reserve(need));
    ctx->ws_reset();
    double* W = (double*)ctx->ws_alloc((int64_t)rows * c * 8);
    double* V = (double*)ctx->ws_alloc((int64_t)c * c * 8);
    double* sig = (double*)ctx->ws_alloc((int64_t)c * 8);
    ttsk::jacobi_pinv_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(d_A, m, n, rcond, W, V, sig, d_pinv, nullptr, nullptr, nullptr, 0,
                                                                   nullptr);
    TTSK_LAUNCHED(ctx);
    return TTSK_OK;
}

// Thin SVD A (m, n) = U diag(S) V^T with k = min(m, n) <= 256 columns (one-sided Jacobi, one CTA; S descending).
// Replaces np.linalg.svd in TensorTrain.round / svdvals (reference tensor.py:446-506): the validation step after a
// sketch, not a throughput kernel.
extern "C" int ttsk_svd(ttsk_ctx* ctx, const double* d_A, int m, int n, double* d_U, double* d_S, double* d_Vt,
                        int u_times_s, void* stream) {
    TTSK_ARG(ctx != nullptr, "ctx is NULL");
    TTSK_ARG(m >= 1 && n >= 1 && m <= 65536 && n <= 65536 && (m <= 256 || n <= 256), "svd: min(m, n) must be <= 256");
    TTSK_ARG(d_A && d_U && d_S && d_Vt, "NULL pointer");
    const int rows = m >= n ? m : n, c = m >= n ? n : m;
    const int64_t need = ((int64_t)rows * c + (int64_t)c * c + 2 * c) * 8 + 2048;
    TTSK_TRY(ctx->ws_reserve(need));
    ctx->ws_reset();
    double* W = (double*)ctx->ws_alloc((int64_t)rows * c * 8);
    double* V = (double*)ctx->ws_alloc((int64_t)c * c * 8);
    double* sig = (double*)ctx->ws_alloc((int64_t)c * 8);
    int* perm = (int*)ctx->ws_alloc((int64_t)c * 4);
    TTSK_ARG(W && V && sig && perm, "svd: workspace");
    ttsk::jacobi_pinv_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(d_A, m, n, -1.0, W, V, sig, nullptr, d_U, d_S, d_Vt, u_times_s, perm);
    TTSK_LAUNCHED(ctx);
    return TTSK_OK;
}

extern "C" int ttsk_qr_q(ttsk_ctx* ctx, double* d_A, int64_t m, int n, void* stream) {
    TTSK_ARG(ctx != nullptr, "ctx is NULL");
    TTSK_ARG(m >= n && n >= 1 && n <= 1024, "qr: need m >= n >= 1");
    TTSK_ARG(d_A != nullptr, "NULL pointer");
    static const int one_cta = getenv("TTSK_QR_ONE_CTA") ? atoi(getenv("TTSK_QR_ONE_CTA")) : 0;
    if (m >= 8192 && n <= 64 && !one_cta) {
        // tall panel: every SM takes a block of rows, the Householder steps run in lock step (cooperative launch)
        int per_sm = 0;
        TTSK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ttsk::householder_q_grid_kernel, ttsk::kQrThreads, 0));
        if (per_sm >= 1) {
            long long grid = ctx->sm_count;
            long long rows = (m + grid - 1) / grid;
            if (rows < 256) rows = 256;
            grid = (m + rows - 1) / rows;
            const int64_t red_doubles = (int64_t)n + 2LL * n * n + n;
            TTSK_TRY(ctx->ws_reserve(red_doubles * 8 + 1024));
            ctx->ws_reset();
            double* red = (double*)ctx->ws_alloc(red_doubles * 8 + 256);
            unsigned* bar = (unsigned*)(red + red_doubles);
            TTSK_CUDA(cudaMemsetAsync(red, 0, (size_t)red_doubles * 8 + 256, (cudaStream_t)stream));
            long long m_arg = m;
            void* args[] = {(void*)&d_A, (void*)&m_arg, (void*)&n, (void*)&rows, (void*)&red, (void*)&bar};
            TTSK_CUDA(cudaLaunchCooperativeKernel((const void*)ttsk::householder_q_grid_kernel, dim3((unsigned)grid), dim3(ttsk::kQrThreads),
                                                  args, 0, (cudaStream_t)stream));
            ctx->launches++;
            return TTSK_OK;
        }
    }
    TTSK_TRY(ctx->ws_reserve((int64_t)n * 8 + 1024));
    ctx->ws_reset();
    double* tau = (double*)ctx->ws_alloc((int64_t)n * 8);
    ttsk::householder_q_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(d_A, m, n, tau);
    TTSK_LAUNCHED(ctx);
    return TTSK_OK;
}
