/*
 * ttsk.h -- C ABI of libttsk.so, the B200 (sm_100a) implementation of tt-sketch's sketching
 * hot path.  This is the drop-in boundary: plain pointers and sizes, no C++/torch types.
 * A host binding (ctypes, cgo, JNI ...) needs nothing but this file; the Python package in
 * tt-sketch_b200/tt_sketch/ binds it with ctypes.
 *
 * Conventions
 *   - every function returns 0 on success, a negative TTSK_E_* code otherwise and never
 *     throws; ttsk_last_error() returns a thread-local message for the last failure.
 *   - pointers named d_* are DEVICE pointers, h_* are HOST pointers; the caller owns all
 *     memory it passes; the library owns only its context-local workspace.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls are
 *     asynchronous with respect to the host unless stated otherwise.
 *   - all arithmetic is IEEE float64; indices are int64 (the reference's dtypes).
 *   - matrices are row-major ("C order") unless explicit strides are given.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the
 * reference checkout, RikVoorhaar/tt-sketch v1.1).
 */
#ifndef TTSK_H
#define TTSK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TTSK_VERSION 100
#define TTSK_MAX_ORDER 16 /* maximum tensor order d */

enum {
    TTSK_OK = 0,
    TTSK_E_ARG = -1,     /* bad argument (shape/rank/order out of range, NULL pointer ...) */
    TTSK_E_CUDA = -2,    /* CUDA runtime error, see ttsk_last_error() */
    TTSK_E_NOMEM = -3,   /* device or pinned-host allocation failed */
    TTSK_E_NODEVICE = -4 /* no CUDA device / wrong architecture */
};

typedef struct ttsk_ctx ttsk_ctx;

/* ---------------------------------------------------------------- context & memory */
int ttsk_version(void);
const char *ttsk_last_error(void);
int ttsk_device_count(int *count);
/* Creates a context bound to CUDA device `device` (fails with TTSK_E_NODEVICE if there is
 * none: there is no CPU fallback). */
int ttsk_create(int device, ttsk_ctx **out);
int ttsk_destroy(ttsk_ctx *ctx);
int ttsk_malloc(ttsk_ctx *ctx, int64_t bytes, void **d_ptr);
int ttsk_free(ttsk_ctx *ctx, void *d_ptr);
int ttsk_malloc_host(ttsk_ctx *ctx, int64_t bytes, void **h_ptr); /* pinned */
int ttsk_free_host(ttsk_ctx *ctx, void *h_ptr);
int ttsk_memcpy_h2d(ttsk_ctx *ctx, void *d_dst, const void *h_src, int64_t bytes, void *stream);
int ttsk_memcpy_d2h(ttsk_ctx *ctx, void *h_dst, const void *d_src, int64_t bytes, void *stream);
int ttsk_memset_zero(ttsk_ctx *ctx, void *d_ptr, int64_t bytes, void *stream);
int ttsk_sync(ttsk_ctx *ctx, void *stream);
/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
int64_t ttsk_launch_count(ttsk_ctx *ctx);
/* Account for kernels of this library launched outside its entry points: the replay of a CUDA graph captured around
 * ttsk_* calls launches the captured kernels again without passing through the counter. */
int ttsk_note_replayed_launches(ttsk_ctx *ctx, int64_t n);
/* The library keeps two grow-only device allocations per context: a workspace arena and a cache of
 * Gaussian-DRM prefix tables (DRM state: reused by later sketches with the same DRM, like the
 * reference keeps TT-DRM cores, tt_sketch/drm/tensor_train_drm.py:52-56).  ttsk_set_table_cache_cap
 * bounds the cache (default 6 GiB; least recently used tables are dropped, never one the running
 * call uses; bonds whose table does not fit are generated on the fly instead);
 * ttsk_trim releases both allocations (synchronises). */
int ttsk_set_table_cache_cap(ttsk_ctx *ctx, int64_t bytes);
int64_t ttsk_table_cache_bytes(ttsk_ctx *ctx);
/* Counter that changes whenever the context's workspace arena is allocated, grown or released: a CUDA graph captured
 * around ttsk_* calls holds arena pointers and must be re-captured when it changes. */
int64_t ttsk_workspace_generation(ttsk_ctx *ctx);
/* Nonzeros per staging buffer of the host-buffer entry points below (default 2^24; two buffers of
 * 8 (d + 1) bytes per nonzero each).  Inputs longer than this are streamed in chunks. */
int ttsk_set_stage_nnz(ttsk_ctx *ctx, int64_t nnz);
int ttsk_trim(ttsk_ctx *ctx);
/* Milliseconds spent in the sketch kernels of the most recent ttsk_sparse_sketch* call,
 * measured with CUDA events on the launching stream (synchronises). */
int ttsk_last_kernel_ms(ttsk_ctx *ctx, double *ms_total, double *ms_dominant);
/* Per-launch milliseconds of the dominant (mode pass) kernel in that call: fills ms[0..min(cap, n)) and returns n in *n_out. */
int ttsk_last_pass_ms(ttsk_ctx *ctx, double *ms, int cap, int *n_out);
/* Number of sparse mode passes this context ran in the segment-GEMM form (small trailing prefix tables). */
int64_t ttsk_sg_pass_count(ttsk_ctx *ctx);

/* ---------------------------------------------------------------- lazy Gaussian DRM
 * Replaces inds_to_normal(indices, shape, rank_min, rank_max, seed)
 *   tt_sketch/drm/fast_lazy_gaussian.pyx:183-201 (and :13-105 underneath).
 * d_idx: k rows of nnz int64, row i at d_idx + i*idx_row_stride.  h_shape: the k dims.
 * d_out: (nnz, rank_max-rank_min) row-major.  Entries are bit-identical to the reference. */
int ttsk_lazy_gaussian(ttsk_ctx *ctx, const int64_t *d_idx, int64_t idx_row_stride, int k,
                       int64_t nnz, const int64_t *h_shape, int rank_min, int rank_max,
                       uint64_t seed, double *d_out, void *stream);
/* Sparse sign DRM rows (entries 0 / -1 / +1) for the first k index rows:
 * replaces inds_to_sparse_sign, tt_sketch/drm/fast_lazy_gaussian.pyx:121-180 (called from
 * SparseSignDRM.sketch_sparse, tt_sketch/drm/sparse_sign_drm.py:34-51).  d_out is (nnz, rank_max - rank_min)
 * row-major FP64: columns [rank_min, rank_max) of the (nnz, rank) matrix whose rows hold nnz_row non-zeros. */
int ttsk_lazy_sparse_sign(ttsk_ctx *ctx, const int64_t *d_idx, int64_t idx_row_stride, int k, int64_t nnz,
                          const int64_t *h_shape, int rank, int rank_min, int rank_max, int nnz_row,
                          uint64_t seed, double *d_out, void *stream);

/* Self-test of the straight-line FP64 division used inside the generator: counts operand
 * pairs (n pseudo-random pairs from `seed`, magnitudes 2^-30..2^10 and zero numerators) whose
 * quotient differs from CUDA's IEEE __ddiv_rn.  Must report 0.  Synchronous. */
int ttsk_selftest_div(ttsk_ctx *ctx, int64_t n, uint64_t seed, uint64_t *h_mismatches);
/* Same for the straight-line FP64 square root of the tail branch (operands 2^-8..2^12) against CUDA's
 * IEEE __dsqrt_rn.  Must report 0.  Synchronous. */
int ttsk_selftest_sqrt(ttsk_ctx *ctx, int64_t n, uint64_t seed, uint64_t *h_mismatches);

/* ---------------------------------------------------------------- DRM descriptors
 * A dimension-reduction map in USER orientation (bond mu = 0..d-2).
 * Replaces the state of tt_sketch/drm_base.py:14-63 (DRM), sparse_gaussian_drm.py:11-27 and
 * tensor_train_drm.py:23-58. */
enum { TTSK_DRM_GAUSS = 1, TTSK_DRM_TT = 2 };

typedef struct {
    int32_t kind;                       /* TTSK_DRM_GAUSS | TTSK_DRM_TT */
    int32_t right;                      /* 0: left DRM (transpose=False), 1: right DRM */
    uint64_t seed;                      /* DRM.seed (already reduced mod 2^32-1) */
    int32_t rank_min[TTSK_MAX_ORDER];   /* per bond, as passed by the user (bond order) */
    int32_t rank_max[TTSK_MAX_ORDER];
    /* TT kind: d-1 device core pointers in the DRM's OWN orientation (for a right DRM core
     * k belongs to mode d-1-k), core k is (core_r0[k], n, core_r1[k]) row-major. */
    const double *d_cores[TTSK_MAX_ORDER];
    int32_t core_r0[TTSK_MAX_ORDER];
    int32_t core_r1[TTSK_MAX_ORDER];
} ttsk_drm;

/* ---------------------------------------------------------------- sparse input (COO)
 * Replaces, for SparseTensor input and method=streaming, the whole of
 *   general_sketch                         tt_sketch/sketch_dispatch.py:202-275
 *   SparseGaussianDRM.sketch_sparse        tt_sketch/drm/sparse_gaussian_drm.py:29-44
 *   TensorTrainDRM.sketch_sparse           tt_sketch/drm/tensor_train_drm.py:60-69
 *   sketch_omega_sparse / sketch_psi_sparse tt_sketch/sketching_methods/sparse_sketch.py:39-69
 * in one call.  d_idx: d rows of nnz int64 (row m at d_idx + m*idx_row_stride), d_val: nnz
 * doubles.  d_out is the packed sketch
 *     [Psi_0 | ... | Psi_{d-1} | Omega_0 | ... | Omega_{d-2}]
 * with Psi_mu of shape (rL[mu-1] or 1, n_mu, rR[mu] or 1) and Omega_mu (rL[mu], rR[mu]), all
 * row-major, rL/rR = rank_max-rank_min of the left/right DRM.  If accumulate == 0 the buffer
 * is zeroed first, otherwise the sketch is ADDED to it (TensorSum, streaming updates,
 * multi-GPU partial sketches: sketch_dispatch.py:85-136). */
int64_t ttsk_sketch_size(int d, const int64_t *h_shape, const int32_t *rL, const int32_t *rR);
int ttsk_sparse_sketch(ttsk_ctx *ctx, int d, const int64_t *h_shape, int64_t nnz,
                       const int64_t *d_idx, int64_t idx_row_stride, const double *d_val,
                       const ttsk_drm *left, const ttsk_drm *right, double *d_out,
                       int accumulate, void *stream);
/* Same with HOST buffers (h_idx/h_val/h_out): chunks are staged through pinned memory and
 * copied host->device on a copy stream overlapped with the kernels; the packed sketch is
 * copied back.  Synchronous.  This is the call bench.py times for the e2e number. */
int ttsk_sparse_sketch_host(ttsk_ctx *ctx, int d, const int64_t *h_shape, int64_t nnz,
                            const int64_t *h_idx, int64_t idx_row_stride, const double *h_val,
                            const ttsk_drm *left, const ttsk_drm *right, double *h_out,
                            int accumulate);
/* Host COO buffers in, DEVICE packed sketch out (so partial sketches of several GPUs can be
 * all-reduced before one device->host copy).  Synchronous; runs on the context's own streams. */
int ttsk_sparse_sketch_stream(ttsk_ctx *ctx, int d, const int64_t *h_shape, int64_t nnz,
                              const int64_t *h_idx, int64_t idx_row_stride, const double *h_val,
                              const ttsk_drm *left, const ttsk_drm *right, double *d_out,
                              int accumulate);

/* Operator-level sparse kernels on explicit per-nonzero DRM rows (the reference's per-mu
 * plug-in signatures).  Element (nonzero p, column a) of the left rows is
 * d_left[p*l_ps + a*l_cs] (so both the reference's (rL, nnz) layout, l_ps=1 l_cs=nnz, and
 * (nnz, rL) chain buffers, l_ps=rL l_cs=1, are views); same for the right rows.  NULL = that
 * side is absent (first / last core).
 *   ttsk_sparse_omega  replaces sketch_omega_sparse  sparse_sketch.py:39-46
 *   ttsk_sparse_psi    replaces sketch_psi_sparse    sparse_sketch.py:49-69 (+ :8-36)
 * Outputs are ADDED to d_omega (rL, rR) / d_psi (rL or 1, n_mu, rR or 1). */
int ttsk_sparse_omega(ttsk_ctx *ctx, int64_t nnz, const double *d_val,
                      const double *d_left, int rL, int64_t l_ps, int64_t l_cs,
                      const double *d_right, int rR, int64_t r_ps, int64_t r_cs,
                      double *d_omega, void *stream);
int ttsk_sparse_psi(ttsk_ctx *ctx, int64_t nnz, const int64_t *d_idx_mu, int64_t n_mu,
                    const double *d_val,
                    const double *d_left, int rL, int64_t l_ps, int64_t l_cs,
                    const double *d_right, int rR, int64_t r_ps, int64_t r_cs,
                    double *d_psi, void *stream);
/* One step of the per-nonzero TT-DRM chain (tensor_train_drm.py:60-69):
 *   v_out[p, :] = v_in[p, :] @ core[:, idx_mu[p], :]      (v_in NULL for the first core)
 * v_in (nnz, r_in), core (r_in, n, r_out), v_out (nnz, r_out), all row-major. */
int ttsk_ttdrm_sparse_step(ttsk_ctx *ctx, int64_t nnz, const int64_t *d_idx_mu, const double *d_v_in,
                           int r_in, const double *d_core, int64_t n, int r_out, double *d_v_out,
                           void *stream);

/* ---------------------------------------------------------------- TensorTrain input
 * Streaming sketch of ONE TensorTrain summand with TensorTrainDRMs, ADDED to the packed sketch d_out
 * (layout of ttsk_sparse_sketch).  Replaces, for TensorTrain input and method=streaming,
 *   TensorTrainDRM.sketch_tt            tt_sketch/drm/tensor_train_drm.py:71-85
 *   sketch_omega_tt / sketch_psi_tt      tt_sketch/sketching_methods/tensor_train_sketch.py:8-35
 * as one fixed sequence of strided GEMMs (no per-GEMM host round trip).  h_tt_rank has d+1 entries
 * (1, r_1, ..., r_{d-1}, 1); h_core_ptrs[k] is the DEVICE pointer of core k, (r_k, n_k, r_{k+1}) row-major. */
int ttsk_tt_sketch(ttsk_ctx *ctx, int d, const int64_t *h_shape, const int32_t *h_tt_rank,
                   const double *const *h_core_ptrs, const ttsk_drm *left, const ttsk_drm *right,
                   double *d_out, void *stream);

/* ---------------------------------------------------------------- DenseTensor input
 * Streaming sketch of ONE DenseTensor (d_X: C-order, prod(h_shape) doubles on the device) with TensorTrainDRMs,
 * ADDED to the packed sketch d_out (layout of ttsk_sparse_sketch).  Replaces, for DenseTensor input and
 * method=streaming,
 *   TensorTrainDRM.sketch_dense           tt_sketch/drm/tensor_train_drm.py:109-122
 *   sketch_omega_dense / sketch_psi_dense  tt_sketch/sketching_methods/dense_sketch.py:7-52
 * X is read from HBM once (TMA-staged fused first pass: XL_0 = G_0^T X_(0) and Psi_0), the left DRM is swept
 * through ever smaller partial contractions and every Omega_mu / Psi_mu is a small product with a right-DRM
 * unfolding used as a flat array like the reference does (its reversed-mode column order included).  Like the
 * reference the dense sketch ignores rank slices: both DRMs must span their whole cores. */
int ttsk_dense_sketch(ttsk_ctx *ctx, int d, const int64_t *h_shape, const double *d_X, const ttsk_drm *left,
                      const ttsk_drm *right, double *d_out, void *stream);

/* ---------------------------------------------------------------- dense building blocks
 * C[b] (M,N) = alpha * A[b] (M,K) @ B[b] (K,N) + beta * C[b], arbitrary element strides
 * (row stride, column stride) so transposes/slices are views.  Backs the TT / CP / dense
 * paths: tensor_train_drm.py:71-122, tensor_train_sketch.py:8-35, cp_sketch.py:6-36,
 * dense_sketch.py:7-52 (all NumPy einsum/matmul calls there). */
int ttsk_gemm(ttsk_ctx *ctx, int64_t M, int64_t N, int64_t K, double alpha,
              const double *d_A, int64_t a_rs, int64_t a_cs,
              const double *d_B, int64_t b_rs, int64_t b_cs, double beta,
              double *d_C, int64_t c_rs, int64_t c_cs,
              int64_t batch, int64_t a_bs, int64_t b_bs, int64_t c_bs, void *stream);
/* out[j, k, m] = A[k, j] * R[j, m]   (CP Khatri-Rao operand, cp_sketch.py:29-35 and
 * tensor_train_drm.py:98-104).  A (n, R) row-major, Rm (R, r) with row stride r_rs,
 * out (R, n, r) row-major. */
int ttsk_khatri_rao(ttsk_ctx *ctx, int64_t n, int64_t R, int64_t r, const double *d_A,
                    const double *d_Rm, int64_t r_rs, double *d_out, void *stream);

/* ---------------------------------------------------------------- small dense LA (assembly)
 * d_pinv (n, m) = pseudo-inverse of d_A (m, n) by one-sided Jacobi SVD, singular values
 * below rcond * s_max dropped (rcond < 0: machine epsilon, LAPACK gelsd's default, as used by
 * scipy.linalg.lstsq in tt_sketch/utils.py:98-109).  m, n <= 256. */
int ttsk_pinv(ttsk_ctx *ctx, const double *d_A, int m, int n, double rcond, double *d_pinv,
              void *stream);
/* Thin SVD A (m, n) = U diag(S) V^T, k = min(m, n) <= 256, max(m, n) <= 65536: U (m, k) row-major (columns times the
 * singular values when u_times_s), S (k) descending, V^T (k, n) row-major.  Replaces np.linalg.svd in
 * TensorTrain.round / svdvals, tt_sketch/tensor.py:446-506 (the step after the sketching path). */
int ttsk_svd(ttsk_ctx *ctx, const double *d_A, int m, int n, double *d_U, double *d_S, double *d_Vt, int u_times_s,
             void *stream);
/* In-place economic QR of d_A (m, n) row-major, m >= n, n <= 256: on return d_A holds Q with
 * LAPACK's Householder sign convention (scipy.linalg.qr(mode="economic") in
 * tt_sketch/sketch_dispatch.py:172). */
int ttsk_qr_q(ttsk_ctx *ctx, double *d_A, int64_t m, int n, void *stream);

/* ---- FROSTT ".tns" text -> COO (host code; the on-disk format in front of the sketching path).  Replaces the
 * per-line Python loop of scripts/frostt.py:51-66 of the reference.  One nonzero per line: d 1-based integer coordinates and
 * a value, blank separated; empty lines and lines starting with '#' are skipped.
 * ttsk_tns_count: number of data lines in buf[0, len) (or -1) and, in *d_out, the order read off the first one.
 * ttsk_tns_parse: fills idx (d rows of nnz int64, 0-based: the layout of ttsk_sparse_sketch_host), val (nnz) and
 * max_idx (d: largest 0-based coordinate per mode); nnz must equal ttsk_tns_count.  Multi-threaded. */
int64_t ttsk_tns_count(const char *buf, int64_t len, int *d_out);
int ttsk_tns_parse(const char *buf, int64_t len, int d, int64_t nnz, int64_t *idx, double *val, int64_t *max_idx);

#ifdef __cplusplus
}
#endif
#endif /* TTSK_H */
