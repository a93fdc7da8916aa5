// FP64 pipe microbenchmark for the roofline denominators of the sparse kernel (run on the B200):
//   dfma : dependent-chain-free DFMA throughput (vector FP64 pipe)
//   dmma : mma.sync.m8n8k4.f64 throughput (FP64 tensor path)
//   both : the two interleaved in one warp -- tells whether DMMA has its own pipe.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_microbench tools/fp64_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, int iters, double x, double y) {
    double f[8], c[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) { f[i] = threadIdx.x * 1e-9 + i; c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0 || MODE == 2) f[i] = fma(f[i], x, y);
            if (MODE == 1 || MODE == 2) dmma(c[i][0], c[i][1], x, y);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += f[i] + c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double* d, int sms) {
    const int iters = 20000, blocks = sms * 8;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<blocks, 256>>>(d, 100, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<MODE><<<blocks, 256>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double threads = (double)blocks * 256, warps = threads / 32;
    const double dfma = (MODE != 1) ? threads * iters * 8.0 : 0.0;              // thread-level DFMA
    const double mma_fma = (MODE != 0) ? warps * iters * 8.0 * 256.0 : 0.0;     // FMAs inside DMMAs
    printf("%-5s %8.3f ms  DFMA %.3e /s (%.2f TFLOP/s)  DMMA-FMA %.3e /s (%.2f TFLOP/s)\n", name, ms,
           dfma / (ms * 1e-3), 2 * dfma / (ms * 1e-3) / 1e12, mma_fma / (ms * 1e-3), 2 * mma_fma / (ms * 1e-3) / 1e12);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    double* d; cudaMalloc(&d, (size_t)p.multiProcessorCount * 8 * 256 * 8);
    run<0>("dfma", d, p.multiProcessorCount);
    run<1>("dmma", d, p.multiProcessorCount);
    run<2>("both", d, p.multiProcessorCount);
    run<0>("dfma", d, p.multiProcessorCount);
    return 0;
}
