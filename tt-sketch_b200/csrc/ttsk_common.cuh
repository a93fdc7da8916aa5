// Shared internals of libttsk.so (not part of the public ABI; see include/ttsk.h).
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

#include "../../include/ttsk.h"

namespace ttsk {

void set_error(const char* fmt, ...);

#define TTSK_CUDA(call)                                                                      \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            ttsk::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                    \
                            cudaGetErrorString(e__));                                        \
            return TTSK_E_CUDA;                                                              \
        }                                                                                    \
    } while (0)

#define TTSK_ARG(cond, msg)                                                                  \
    do {                                                                                     \
        if (!(cond)) {                                                                       \
            ttsk::set_error("%s:%d: bad argument: %s", __FILE__, __LINE__, msg);             \
            return TTSK_E_ARG;                                                               \
        }                                                                                    \
    } while (0)

#define TTSK_TRY(call)                                                                       \
    do {                                                                                     \
        int rc__ = (call);                                                                   \
        if (rc__ != TTSK_OK) return rc__;                                                    \
    } while (0)

// check the launch that was just issued and count it
#define TTSK_LAUNCHED(ctx)                                                                   \
    do {                                                                                     \
        (ctx)->launches++;                                                                   \
        TTSK_CUDA(cudaGetLastError());                                                       \
    } while (0)

inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

}  // namespace ttsk

// Context: device binding, launch counter, a grow-only device workspace arena that is
// bump-allocated per API call (all work of one context is issued on the caller's stream, so
// reuse is stream-ordered), pinned staging buffers for the host-buffer entry points.
struct ttsk_ctx {
    int device = 0;
    int sm_count = 148;
    int64_t launches = 0;
    int64_t sg_passes = 0;  // mode passes that ran in the segment-GEMM form
    // workspace arena
    char* ws = nullptr;
    int64_t ws_bytes = 0;
    int64_t ws_used = 0;
    int64_t ws_gen = 0;  // bumped whenever the arena is (re)allocated or freed: captured CUDA graphs hold arena pointers
    // pinned staging + copy stream for *_host entry points
    void* pinned[2] = {nullptr, nullptr};
    int64_t pinned_bytes = 0;
    cudaStream_t copy_stream = nullptr;
    cudaStream_t compute_stream = nullptr;
    cudaEvent_t ev_copy[2] = {nullptr, nullptr};
    cudaEvent_t ev_done[2] = {nullptr, nullptr};
    // kernel timing of the last sparse sketch
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
    std::vector<cudaEvent_t> ev_pass;  // pairs (start, stop) around every pass kernel
    int n_pass_events = 0;
    bool timing = true;
    // prefix tables of Gaussian DRMs depend only on (seed, column range, rows): kept across calls
    struct TableEntry { uint64_t seed; int rank_min, r; int64_t rows; double* ptr; int64_t bytes; cudaStream_t stream; int64_t pin_gen; };
    std::vector<TableEntry> tables;  // least recently used first
    int64_t table_bytes = 0;
    int64_t table_cap = (int64_t)6 << 30;
    int64_t stage_nnz = (int64_t)1 << 24;  // nonzeros per staging buffer of the host-buffer entry points
    int64_t plan_gen = 0;  // generation id of the sparse plan being built: its tables are never evicted

    int ws_reserve(int64_t bytes);              // make the arena at least this large (may sync)
    void ws_reset() { ws_used = 0; }
    void* ws_alloc(int64_t bytes);              // nullptr if the arena is too small
};

namespace ttsk {
// internal launchers shared between translation units
int gemm_launch(ttsk_ctx* ctx, int64_t M, int64_t N, int64_t K, double alpha, const double* A,
                int64_t a_rs, int64_t a_cs, const double* B, int64_t b_rs, int64_t b_cs,
                double beta, double* C, int64_t c_rs, int64_t c_cs, int64_t batch, int64_t a_bs,
                int64_t b_bs, int64_t c_bs, cudaStream_t st);
int axpy_launch(ttsk_ctx* ctx, int64_t n, double alpha, const double* x, double* y, cudaStream_t st);
}  // namespace ttsk
