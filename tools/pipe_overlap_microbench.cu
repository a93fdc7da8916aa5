// Which integer instruction classes overlap with FP64 work on a B200 SM sub-partition?
// Each thread runs 4 independent {DMUL, DADD} chains and, per FP64 pair, N extra instructions of one class
// on independent integer chains.  If the class overlaps with the FP64 pipe, time stays flat while
// 2*N <= 4 (the FP64 pair takes 4 issue cycles at 16 lanes/clk); otherwise it grows with N.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_overlap_microbench tools/pipe_overlap_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

enum { LOP3 = 0, SHF = 1, IADD3 = 2, IMAD = 3, IMADW = 4, LDS = 5, NONE = 6 };

template <int CLS>
__device__ __forceinline__ void extra(unsigned& a, unsigned b, unsigned long long& w, const unsigned* s) {
    if (CLS == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(b + 1));
    if (CLS == SHF) asm volatile("shf.r.wrap.b32 %0, %0, %1, 13;" : "+r"(a) : "r"(b));
    if (CLS == IADD3) asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(a) : "r"(b), "r"(b + 7));
    if (CLS == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(b + 3));
    if (CLS == IMADW) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w) : "r"(a), "r"(b));
    if (CLS == LDS) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"((a & 0xffcu))); a ^= v; }
}

template <int CLS, int N>
__global__ void __launch_bounds__(256) k(double* out, int iters, double x, double y, unsigned seed) {
    __shared__ unsigned s[1024];
    for (int i = threadIdx.x; i < 1024; i += 256) s[i] = i * 4;
    __syncthreads();
    double f[4];
    unsigned a[4][N > 0 ? N : 1];
    unsigned long long w[4][N > 0 ? N : 1];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        f[c] = threadIdx.x * 1e-9 + c;
#pragma unroll
        for (int n = 0; n < (N > 0 ? N : 1); n++) { a[c][n] = seed + threadIdx.x * 977u + c * 31u + n; w[c][n] = a[c][n]; }
    }
    const unsigned b = seed * 2654435761u + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            f[c] = __dadd_rn(__dmul_rn(f[c], x), y);
#pragma unroll
            for (int n = 0; n < N; n++) extra<CLS>(a[c][n], b, w[c][n], s);
        }
    }
    double acc = 0;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        acc += f[c];
#pragma unroll
        for (int n = 0; n < (N > 0 ? N : 1); n++) acc += (double)(a[c][n] + (unsigned)w[c][n]);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int CLS, int N>
void run(const char* name, double* d, int sms) {
    const int iters = 20000, blocks = sms * 4;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<CLS, N><<<blocks, 256>>>(d, 100, 1.0000001, 1e-9, 1); cudaDeviceSynchronize();
    cudaEventRecord(a); k<CLS, N><<<blocks, 256>>>(d, iters, 1.0000001, 1e-9, 1); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    // cycles per {DMUL, DADD} pair per scheduler: 8 warps per scheduler, 4 pairs per iteration per warp
    const double cyc = ms * 1e-3 * 1.965e9 / ((double)iters * 4 * 8);
    printf("%-6s extra/pair=%d  %7.3f ms  %.2f cycles per FP64 pair per scheduler\n", name, N, ms, cyc);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* d; cudaMalloc(&d, (size_t)p.multiProcessorCount * 4 * 256 * 8);
    const int sms = p.multiProcessorCount;
    run<NONE, 0>("none", d, sms);
#define SWEEP(C, nm) run<C, 1>(nm, d, sms); run<C, 2>(nm, d, sms); run<C, 3>(nm, d, sms); run<C, 4>(nm, d, sms);
    SWEEP(LOP3, "lop3") SWEEP(SHF, "shf") SWEEP(IADD3, "iadd") SWEEP(IMAD, "imad") SWEEP(IMADW, "imadw") SWEEP(LDS, "lds")
    return 0;
}
