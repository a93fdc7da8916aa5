// How do the FP64 pipe and the integer pipes share the issue port of a B200 scheduler on the generator's mix?
// Every thread runs CH independent Horner chains  f = add.rn(mul.rn(f, x), c[k])  (the unfused steps of the bit-exact
// ndtri polynomials), all written as volatile asm so program order is fixed, with NX extra integer instructions placed
// (a) interleaved one by one after the FP64 instructions, or (b) clustered in front of the chain's FP64 block.
//   cycles per FP64 instruction per scheduler = 2.0 when the FP64 pipe is the only limit.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/issue_model_microbench tools/issue_model_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

__constant__ double c_coef[16] = {1e-9, 2e-9, 3e-9, 4e-9, 5e-9, 6e-9, 7e-9, 8e-9, 9e-9, 1e-8, 1.1e-8, 1.2e-8, 1.3e-8, 1.4e-8, 1.5e-8, 1.6e-8};

#define DMUL(d, a, b) asm volatile("mul.rn.f64 %0, %1, %2;" : "=d"(d) : "d"(a), "d"(b))
#define DADD(d, a, b) asm volatile("add.rn.f64 %0, %1, %2;" : "=d"(d) : "d"(a), "d"(b))
#define XOP(a, b) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(b + 1))
#define SOP(a, b) asm volatile("shf.r.wrap.b32 %0, %0, %1, 13;" : "+r"(a) : "r"(b))

// STEPS Horner steps per chain and iteration (2 FP64 each); NX extras per Horner step and chain
template <int CH, int STEPS, int NX, bool CLUSTER, bool CONSTC>
__global__ void __launch_bounds__(256) k(double* out, int iters, double x, double y, unsigned seed) {
    double f[CH];
    unsigned a[CH];
#pragma unroll
    for (int c = 0; c < CH; c++) { f[c] = threadIdx.x * 1e-9 + c; a[c] = seed + threadIdx.x * 977u + c * 31u; }
    const unsigned b = seed * 2654435761u + threadIdx.x;
    for (int it = 0; it < iters; it++) {
        if (CLUSTER) {
#pragma unroll
            for (int c = 0; c < CH; c++)
#pragma unroll
                for (int n = 0; n < NX * STEPS; n++) { if (n & 1) XOP(a[c], b); else SOP(a[c], b); }
        }
#pragma unroll
        for (int s = 0; s < STEPS; s++) {
#pragma unroll
            for (int c = 0; c < CH; c++) {
                double t;
                DMUL(t, f[c], x);
                if (!CLUSTER && NX >= 1) SOP(a[c], b);
                if (CONSTC) DADD(f[c], t, c_coef[s & 15]); else DADD(f[c], t, y);
                if (!CLUSTER && NX >= 2) XOP(a[c], b);
                if (!CLUSTER && NX >= 3) SOP(a[c], b);
                if (!CLUSTER && NX >= 4) XOP(a[c], b);
            }
        }
    }
    double acc = 0;
#pragma unroll
    for (int c = 0; c < CH; c++) acc += f[c] + (double)a[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int CH, int STEPS, int NX, bool CLUSTER, bool CONSTC>
void run(double* d, int sms, int ctas) {
    const int iters = 2000, blocks = sms * ctas;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<CH, STEPS, NX, CLUSTER, CONSTC><<<blocks, 256>>>(d, 10, 1.0000001, 1e-9, 1); cudaDeviceSynchronize();
    cudaEventRecord(a); k<CH, STEPS, NX, CLUSTER, CONSTC><<<blocks, 256>>>(d, iters, 1.0000001, 1e-9, 1); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double warps_per_sched = ctas * 8 / 4.0;
    const double fp64 = (double)iters * STEPS * CH * 2 * warps_per_sched;  // FP64 warp-instructions per scheduler
    printf("chains=%d steps=%2d extras/step=%d %s %s warps/sched=%4.1f  %7.3f ms  %.2f cycles per FP64 instr per scheduler\n", CH, STEPS, NX,
           CLUSTER ? "clustered  " : "interleaved", CONSTC ? "const" : "reg  ", warps_per_sched, ms, ms * 1e-3 * 1.965e9 / fp64);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* d; cudaMalloc(&d, (size_t)p.multiProcessorCount * 8 * 256 * 8);
    const int sms = p.multiProcessorCount;
    for (int ctas : {2, 4}) {
        run<4, 12, 0, false, false>(d, sms, ctas);
        run<4, 12, 1, false, false>(d, sms, ctas);
        run<4, 12, 2, false, false>(d, sms, ctas);
        run<4, 12, 3, false, false>(d, sms, ctas);
        run<4, 12, 1, true, false>(d, sms, ctas);
        run<4, 12, 2, true, false>(d, sms, ctas);
        run<4, 12, 0, false, true>(d, sms, ctas);
        run<4, 12, 1, false, true>(d, sms, ctas);
        run<4, 12, 2, false, true>(d, sms, ctas);
        run<4, 12, 2, true, true>(d, sms, ctas);
        run<2, 12, 2, false, true>(d, sms, ctas);
        run<1, 12, 2, false, true>(d, sms, ctas);
    }
    return 0;
}
