// What the FP64 pipe of a B200 sustains on the instruction mix of the bit-exact ndtri (DMUL + DADD
// Horner steps that must NOT be fused, operands from registers / the constant bank), next to DFMA:
//   dfma      : f = fma(f, x, y)                       8 independent chains per thread
//   muladd_r  : f = dadd(dmul(f, x), y), x, y registers 8 chains
//   muladd_c  : same with y from the constant bank (what nvcc emits for polynomial coefficients)
//   mix_int   : muladd_r with 1 integer instruction per FP64 instruction interleaved (hash-like load)
// and the same at 1 / 2 / 4 / 8 CTAs of 256 threads per SM (2 .. 16 warps per scheduler).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_mix_microbench tools/fp64_mix_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

__constant__ double c_coef[8] = {1e-9, 2e-9, 3e-9, 4e-9, 5e-9, 6e-9, 7e-9, 8e-9};

template <int MODE, int CH>
__global__ void __launch_bounds__(256) k(double* out, int iters, double x, double y, unsigned seed) {
    double f[CH];
    unsigned h[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { f[i] = threadIdx.x * 1e-9 + i; h[i] = seed + threadIdx.x * 977u + i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (MODE == 0) f[i] = fma(f[i], x, y);
            if (MODE == 1) f[i] = __dadd_rn(__dmul_rn(f[i], x), y);
            if (MODE == 2) f[i] = __dadd_rn(__dmul_rn(f[i], x), c_coef[i & 7]);
            if (MODE == 3) {
                f[i] = __dadd_rn(__dmul_rn(f[i], x), y);
                h[i] = (h[i] ^ (h[i] >> 13)) + 0x9E3779B9u;   // 3 ALU-pipe instructions
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += f[i] + (double)h[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int CH>
void run(const char* name, double* d, int sms, int ctas) {
    const int iters = 20000, blocks = sms * ctas;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE, CH><<<blocks, 256>>>(d, 100, 1.0000001, 1e-9, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<MODE, CH><<<blocks, 256>>>(d, iters, 1.0000001, 1e-9, 1);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double fp64 = (double)blocks * 256 * iters * CH * (MODE == 0 ? 1.0 : 2.0);
    printf("%-9s chains=%d ctas/sm=%d %8.3f ms  FP64 thread-instr/s %.3e\n", name, CH, ctas, ms, fp64 / (ms * 1e-3));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    double* d; cudaMalloc(&d, (size_t)p.multiProcessorCount * 8 * 256 * 8);
    const int sms = p.multiProcessorCount;
    for (int ctas : {1, 2, 4, 8}) {
        run<0, 8>("dfma", d, sms, ctas);
        run<1, 8>("muladd_r", d, sms, ctas);
        run<2, 8>("muladd_c", d, sms, ctas);
        run<3, 8>("mix_int", d, sms, ctas);
        run<1, 2>("muladd_r", d, sms, ctas);
        run<1, 4>("muladd_r", d, sms, ctas);
    }
    return 0;
}
