// Ceiling of the lazy-Gaussian generator in isolation (registers only, no memory):
//   gen   : hash -> uniform -> branch-free central ndtri, two chains per thread (as in fill_gauss)
//   tail  : ndtri_tail on uniforms forced into the lower tail
//   hash  : hash + uniform only
// Reports variates/s and the FP64-instruction rate vs the DFMA peak of tools/fp64_microbench.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I tt-sketch_b200/csrc -o tools/ndtri_microbench tools/ndtri_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ttsk_gauss.cuh"
using namespace ttsk;

template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, int iters, unsigned long long seed) {
    __shared__ double2 s_tab[ttsk::kGaussTabEntries];
    load_logtab(s_tab);
    __syncthreads();
    unsigned long long f0 = seed + threadIdx.x + 977ull * blockIdx.x, f1 = f0 * 31 + 7;
    double acc = 0.0;
    for (int it = 0; it < iters; it++) {
        const double u0 = uniform_from_hash(hash64(f0 + it)), u1 = uniform_from_hash(hash64(f1 + it));
        if (MODE == 0) {
            const double c0 = ndtri_central(u0), c1 = ndtri_central(u1);
            acc += (ndtri_class(u0) ? u0 : c0) + (ndtri_class(u1) ? u1 : c1);
        } else if (MODE == 1) {
            acc += ndtri_tail(u0 * 0.13, 1, s_tab) + ndtri_tail(u1 * 0.13, 1, s_tab);
        } else {
            acc += u0 + u1;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, double* d, int sms, int ctas_per_sm, double fp64_per_variate) {
    const int iters = 4000, blocks = sms * ctas_per_sm;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<blocks, 256>>>(d, 10, 1); cudaDeviceSynchronize();
    cudaEventRecord(a); k<MODE><<<blocks, 256>>>(d, iters, 1); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double variates = (double)blocks * 256 * iters * 2;
    printf("%-5s ctas/sm=%d  %8.3f ms  %.3e variates/s  fp64 instr/s %.3e (%.0f%% of 1.68e13)\n", name, ctas_per_sm, ms,
           variates / (ms * 1e-3), variates * fp64_per_variate / (ms * 1e-3), 100 * variates * fp64_per_variate / (ms * 1e-3) / 1.68e13);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* d; cudaMalloc(&d, (size_t)p.multiProcessorCount * 8 * 256 * 8);
    for (int c : {1, 2, 3, 4, 8}) {
        if (c == 1) { run<0>("gen", d, p.multiProcessorCount, 1, 43); run<1>("tail", d, p.multiProcessorCount, 1, 112); run<2>("hash", d, p.multiProcessorCount, 1, 1); }
        if (c == 2) { run<0>("gen", d, p.multiProcessorCount, 2, 43); run<1>("tail", d, p.multiProcessorCount, 2, 112); run<2>("hash", d, p.multiProcessorCount, 2, 1); }
        if (c == 3) { run<0>("gen", d, p.multiProcessorCount, 3, 43); }
        if (c == 4) { run<0>("gen", d, p.multiProcessorCount, 4, 43); run<1>("tail", d, p.multiProcessorCount, 4, 112); run<2>("hash", d, p.multiProcessorCount, 4, 1); }
        if (c == 8) { run<0>("gen", d, p.multiProcessorCount, 8, 43); run<1>("tail", d, p.multiProcessorCount, 8, 112); run<2>("hash", d, p.multiProcessorCount, 8, 1); }
    }
    return 0;
}
