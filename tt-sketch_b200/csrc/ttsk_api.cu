// Context, memory and error plumbing of libttsk.so (C ABI in include/ttsk.h).
#include <cstdlib>
#include <cstring>

#include "ttsk_common.cuh"

namespace ttsk {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

__global__ void axpy_kernel(int64_t n, double alpha, const double* __restrict__ x, double* __restrict__ y) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) y[i] += alpha * x[i];
}
int axpy_launch(ttsk_ctx* ctx, int64_t n, double alpha, const double* x, double* y, cudaStream_t st) {
    if (n <= 0) return TTSK_OK;
    int64_t blocks = (n + 255) / 256;
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    axpy_kernel<<<(unsigned)blocks, 256, 0, st>>>(n, alpha, x, y);
    TTSK_LAUNCHED(ctx);
    return TTSK_OK;
}
}  // namespace ttsk

int ttsk_ctx::ws_reserve(int64_t bytes) {
    if (bytes <= ws_bytes) return TTSK_OK;
    ws_gen++;
    if (ws) {
        TTSK_CUDA(cudaDeviceSynchronize());
        TTSK_CUDA(cudaFree(ws));
        ws = nullptr;
        ws_bytes = 0;
    }
    bytes = ttsk::align_up(bytes, 1 << 20);
    cudaError_t e = cudaMalloc((void**)&ws, bytes);
    if (e != cudaSuccess) {
        ttsk::set_error("workspace allocation of %lld bytes failed: %s", (long long)bytes, cudaGetErrorString(e));
        cudaGetLastError();
        return TTSK_E_NOMEM;
    }
    ws_bytes = bytes;
    return TTSK_OK;
}

void* ttsk_ctx::ws_alloc(int64_t bytes) {
    int64_t off = ttsk::align_up(ws_used, 256);
    if (off + bytes > ws_bytes) return nullptr;
    ws_used = off + bytes;
    return ws + off;
}

extern "C" {

int ttsk_version(void) { return TTSK_VERSION; }
const char* ttsk_last_error(void) { return ttsk::g_err; }

int ttsk_device_count(int* count) {
    TTSK_ARG(count != nullptr, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    *count = n;
    return TTSK_OK;
}

int ttsk_create(int device, ttsk_ctx** out) {
    TTSK_ARG(out != nullptr, "out is NULL");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        ttsk::set_error("no CUDA device visible: libttsk has no CPU fallback");
        return TTSK_E_NODEVICE;
    }
    TTSK_ARG(device >= 0 && device < n, "device index out of range");
    TTSK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    TTSK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        ttsk::set_error("device %d is sm_%d%d; libttsk is built for sm_100a only", device, prop.major, prop.minor);
        return TTSK_E_NODEVICE;
    }
    // cache-policy knob (does not change results): bytes the L2 fetches from DRAM around a missing sector
    if (getenv("TTSK_L2_FETCH")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(getenv("TTSK_L2_FETCH")));
    ttsk_ctx* c = new ttsk_ctx();
    if (getenv("TTSK_STAGE_NNZ") && atoll(getenv("TTSK_STAGE_NNZ")) > 0) c->stage_nnz = atoll(getenv("TTSK_STAGE_NNZ"));
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    TTSK_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    TTSK_CUDA(cudaStreamCreateWithFlags(&c->compute_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        TTSK_CUDA(cudaEventCreateWithFlags(&c->ev_copy[i], cudaEventDisableTiming));
        TTSK_CUDA(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
    }
    TTSK_CUDA(cudaEventCreate(&c->ev_t0));
    TTSK_CUDA(cudaEventCreate(&c->ev_t1));
    *out = c;
    return TTSK_OK;
}

int ttsk_destroy(ttsk_ctx* ctx) {
    if (!ctx) return TTSK_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    if (ctx->ws) cudaFree(ctx->ws);
    for (auto& t : ctx->tables) cudaFree(t.ptr);
    for (int i = 0; i < 2; i++) {
        if (ctx->pinned[i]) cudaFreeHost(ctx->pinned[i]);
        if (ctx->ev_copy[i]) cudaEventDestroy(ctx->ev_copy[i]);
        if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]);
    }
    for (auto e : ctx->ev_pass) cudaEventDestroy(e);
    if (ctx->ev_t0) cudaEventDestroy(ctx->ev_t0);
    if (ctx->ev_t1) cudaEventDestroy(ctx->ev_t1);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->compute_stream) cudaStreamDestroy(ctx->compute_stream);
    delete ctx;
    return TTSK_OK;
}

int ttsk_malloc(ttsk_ctx* ctx, int64_t bytes, void** d_ptr) {
    TTSK_ARG(ctx && d_ptr && bytes >= 0, "ttsk_malloc");
    TTSK_CUDA(cudaSetDevice(ctx->device));
    *d_ptr = nullptr;
    if (bytes == 0) return TTSK_OK;
    cudaError_t e = cudaMalloc(d_ptr, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        ttsk::set_error("cudaMalloc(%lld) failed: %s", (long long)bytes, cudaGetErrorString(e));
        return TTSK_E_NOMEM;
    }
    return TTSK_OK;
}
int ttsk_free(ttsk_ctx* ctx, void* d_ptr) {
    TTSK_ARG(ctx != nullptr, "ctx is NULL");
    if (d_ptr) TTSK_CUDA(cudaFree(d_ptr));
    return TTSK_OK;
}
int ttsk_malloc_host(ttsk_ctx* ctx, int64_t bytes, void** h_ptr) {
    TTSK_ARG(ctx && h_ptr && bytes >= 0, "ttsk_malloc_host");
    *h_ptr = nullptr;
    if (bytes == 0) return TTSK_OK;
    cudaError_t e = cudaMallocHost(h_ptr, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        ttsk::set_error("cudaMallocHost(%lld) failed: %s", (long long)bytes, cudaGetErrorString(e));
        return TTSK_E_NOMEM;
    }
    return TTSK_OK;
}
int ttsk_free_host(ttsk_ctx* ctx, void* h_ptr) {
    TTSK_ARG(ctx != nullptr, "ctx is NULL");
    if (h_ptr) TTSK_CUDA(cudaFreeHost(h_ptr));
    return TTSK_OK;
}
int ttsk_memcpy_h2d(ttsk_ctx* ctx, void* d_dst, const void* h_src, int64_t bytes, void* stream) {
    TTSK_ARG(ctx != nullptr, "ctx is NULL");
    if (bytes > 0) TTSK_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return TTSK_OK;
}
int ttsk_memcpy_d2h(ttsk_ctx* ctx, void* h_dst, const void* d_src, int64_t bytes, void* stream) {
    TTSK_ARG(ctx != nullptr, "ctx is NULL");
    if (bytes > 0) TTSK_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return TTSK_OK;
}
int ttsk_memset_zero(ttsk_ctx* ctx, void* d_ptr, int64_t bytes, void* stream) {
    TTSK_ARG(ctx != nullptr, "ctx is NULL");
    if (bytes > 0) TTSK_CUDA(cudaMemsetAsync(d_ptr, 0, bytes, (cudaStream_t)stream));
    return TTSK_OK;
}
int ttsk_sync(ttsk_ctx* ctx, void* stream) {
    TTSK_ARG(ctx != nullptr, "ctx is NULL");
    TTSK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return TTSK_OK;
}
int64_t ttsk_launch_count(ttsk_ctx* ctx) { return ctx ? ctx->launches : -1; }
int ttsk_note_replayed_launches(ttsk_ctx* ctx, int64_t n) {
    TTSK_ARG(ctx != nullptr && n >= 0, "ttsk_note_replayed_launches");
    ctx->launches += n;
    return TTSK_OK;
}

int ttsk_set_table_cache_cap(ttsk_ctx* ctx, int64_t bytes) {
    TTSK_ARG(ctx != nullptr && bytes >= 0, "ttsk_set_table_cache_cap");
    ctx->table_cap = bytes;
    return TTSK_OK;
}
int64_t ttsk_table_cache_bytes(ttsk_ctx* ctx) { return ctx ? ctx->table_bytes : -1; }
int64_t ttsk_workspace_generation(ttsk_ctx* ctx) { return ctx ? ctx->ws_gen : -1; }

int ttsk_set_stage_nnz(ttsk_ctx* ctx, int64_t nnz) {
    TTSK_ARG(ctx != nullptr && nnz >= 1, "ttsk_set_stage_nnz");
    ctx->stage_nnz = nnz;
    return TTSK_OK;
}

int ttsk_trim(ttsk_ctx* ctx) {
    TTSK_ARG(ctx != nullptr, "ctx is NULL");
    TTSK_CUDA(cudaSetDevice(ctx->device));
    TTSK_CUDA(cudaDeviceSynchronize());
    if (ctx->ws) TTSK_CUDA(cudaFree(ctx->ws));
    ctx->ws = nullptr;
    ctx->ws_bytes = ctx->ws_used = 0;
    ctx->ws_gen++;
    for (auto& t : ctx->tables) TTSK_CUDA(cudaFree(t.ptr));
    ctx->tables.clear();
    ctx->table_bytes = 0;
    return TTSK_OK;
}

}  // extern "C"
