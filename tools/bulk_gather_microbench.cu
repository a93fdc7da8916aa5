// Microbenchmark: random gather of ROW_BYTES-byte table rows into shared memory, one
// cp.async.bulk per row (mbarrier complete_tx), double-buffered tiles of TN rows, vs. the same
// rows read with plain 16-byte LDGs.  Decides whether the pass kernel stages table rows with the
// bulk-copy engine.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_gather_microbench bulk_gather_microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

template <int TN, int STAGES>
__global__ void __launch_bounds__(256) bulk_kernel(const char* __restrict__ table, const int* __restrict__ ids,
                                                   long long n, int row_bytes, double* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    char* bufs = reinterpret_cast<char*>(smem + 128);
    const int tid = threadIdx.x;
    const size_t stage_bytes = (size_t)TN * row_bytes;
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long tiles = n / TN;
    const long long per = (tiles + gridDim.x - 1) / gridDim.x;
    const long long t0 = (long long)blockIdx.x * per, t1 = (t0 + per < tiles) ? t0 + per : tiles;
    double acc = 0.0;
    auto issue = [&](long long t, int s) {
        if (tid == 0) mbar_expect_tx(&bars[s], (uint32_t)stage_bytes);
        __syncwarp();
        if (tid < TN) {
            const int id = ids[t * TN + tid];
            bulk_g2s(bufs + s * stage_bytes + (size_t)tid * row_bytes, table + (size_t)id * row_bytes, row_bytes, &bars[s]);
        }
    };
    for (int s = 0; s < STAGES - 1; s++)
        if (t0 + s < t1) issue(t0 + s, s);
    for (long long t = t0; t < t1; t++) {
        const int s = (int)((t - t0) % STAGES);
        const uint32_t parity = (uint32_t)(((t - t0) / STAGES) & 1);
        if (t + STAGES - 1 < t1) issue(t + STAGES - 1, (int)((t - t0 + STAGES - 1) % STAGES));
        mbar_wait(&bars[s], parity);
        const double* d = reinterpret_cast<const double*>(bufs + s * stage_bytes);
        const int words = (int)(stage_bytes / 8);
        for (int i = tid; i < words; i += 256) acc += d[i];
        __syncthreads();
    }
    if (acc == 12345.678) out[0] = acc;
}

template <int TN>
__global__ void __launch_bounds__(256) ldg_kernel(const char* __restrict__ table, const int* __restrict__ ids,
                                                  long long n, int row_bytes, double* __restrict__ out) {
    const int tid = threadIdx.x;
    const long long tiles = n / TN;
    const long long per = (tiles + gridDim.x - 1) / gridDim.x;
    const long long t0 = (long long)blockIdx.x * per, t1 = (t0 + per < tiles) ? t0 + per : tiles;
    const int chunks = row_bytes / 16;
    double acc = 0.0;
    for (long long t = t0; t < t1; t++) {
        for (int e = tid; e < TN * chunks; e += 256) {
            const int r = e / chunks, c = e - r * chunks;
            const int id = ids[t * TN + r];
            const double2 v = __ldg(reinterpret_cast<const double2*>(table + (size_t)id * row_bytes) + c);
            acc += v.x + v.y;
        }
    }
    if (acc == 12345.678) out[0] = acc;
}

int main(int argc, char** argv) {
    const int row_bytes = argc > 1 ? atoi(argv[1]) : 320;
    const long long rows = argc > 2 ? atoll(argv[2]) : 5000000;
    const long long n = 1LL << 26;
    char* table; int* ids; double* out;
    CK(cudaMalloc(&table, (size_t)rows * row_bytes));
    CK(cudaMemset(table, 0, (size_t)rows * row_bytes));
    CK(cudaMalloc(&ids, n * 4));
    CK(cudaMalloc(&out, 8));
    int* h = (int*)malloc(n * 4);
    uint64_t s = 88172645463325252ULL;
    for (long long i = 0; i < n; i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % (uint64_t)rows); }
    CK(cudaMemcpy(ids, h, n * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto report = [&](const char* name, float ms) {
        printf("%-34s row=%dB rows=%lld: %.3f ms  %.1f GB/s  %.2f Grows/s\n", name, row_bytes, rows, ms,
               (double)n * row_bytes / ms / 1e6, (double)n / ms / 1e6);
    };
#define RUN_BULK(TN, ST, CTAS)                                                                         \
    {                                                                                                  \
        auto k = bulk_kernel<TN, ST>;                                                                  \
        size_t sm = 128 + (size_t)ST * TN * row_bytes;                                                 \
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));             \
        k<<<148 * CTAS, 256, sm>>>(table, ids, n, row_bytes, out);                                     \
        CK(cudaDeviceSynchronize());                                                                   \
        cudaEventRecord(e0); k<<<148 * CTAS, 256, sm>>>(table, ids, n, row_bytes, out); cudaEventRecord(e1); \
        CK(cudaDeviceSynchronize());                                                                   \
        float ms; cudaEventElapsedTime(&ms, e0, e1);                                                   \
        report("bulk TN=" #TN " stages=" #ST " ctas/sm=" #CTAS, ms);                                   \
    }
    RUN_BULK(128, 2, 1)
    RUN_BULK(128, 2, 2)
    RUN_BULK(128, 3, 1)
    RUN_BULK(64, 2, 2)
    RUN_BULK(64, 3, 2)
    RUN_BULK(64, 4, 2)
    RUN_BULK(64, 2, 4)
    {
        auto k = ldg_kernel<128>;
        for (int ctas = 2; ctas <= 8; ctas *= 2) {
            k<<<148 * ctas, 256>>>(table, ids, n, row_bytes, out);
            CK(cudaDeviceSynchronize());
            cudaEventRecord(e0); k<<<148 * ctas, 256>>>(table, ids, n, row_bytes, out); cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            char nm[64]; snprintf(nm, 64, "ldg16 TN=128 ctas/sm=%d", ctas);
            report(nm, ms);
        }
    }
    return 0;
}
