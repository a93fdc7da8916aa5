// Where the central branch of the bit-exact generator loses FP64-pipe time (registers only, no memory):
//   0 central only, uniforms from a 1-instruction LCG     1 the same without the division (MUFU + 8 FP64)
//   2 hash + central (the generator's main loop body)      3 hash only
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I tt-sketch_b200/csrc -o tools/ndtri_mix_microbench tools/ndtri_mix_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ttsk_gauss.cuh"
using namespace ttsk;

__device__ __forceinline__ double central_nodiv(double b) {
    const double y = __dadd_rn(b, -1.5);
    const double y2 = __dmul_rn(y, y);
    double p = c_nd[0];
    TTSK_HORNER(p, y2, c_nd[1]); TTSK_HORNER(p, y2, c_nd[2]); TTSK_HORNER(p, y2, c_nd[3]); TTSK_HORNER(p, y2, c_nd[4]);
    double q = __dadd_rn(y2, c_nd[5]);
    TTSK_HORNER(q, y2, c_nd[6]); TTSK_HORNER(q, y2, c_nd[7]); TTSK_HORNER(q, y2, c_nd[8]); TTSK_HORNER(q, y2, c_nd[9]);
    TTSK_HORNER(q, y2, c_nd[10]); TTSK_HORNER(q, y2, c_nd[11]); TTSK_HORNER(q, y2, c_nd[12]);
    const double t = __dmul_rn(__dmul_rn(y2, p), q);
    return __dmul_rn(__dadd_rn(y, __dmul_rn(y, t)), c_misc[0]);
}

template <int MODE, int CH>
__global__ void __launch_bounds__(256) k(double* out, int iters, unsigned long long seed) {
    unsigned kk[CH];
    unsigned long long f[CH];
#pragma unroll
    for (int c = 0; c < CH; c++) { kk[c] = (unsigned)seed + threadIdx.x * 977u + c * 131u; f[c] = seed * (c + 3) + threadIdx.x + 977ull * blockIdx.x; }
    double acc = 0.0;
    for (int it = 0; it < iters; it++) {
        double r[CH];
#pragma unroll
        for (int c = 0; c < CH; c++) {
            if (MODE == 0 || MODE == 1) {
                kk[c] = kk[c] * 1664525u + 1013904223u;
                const double b = __hiloint2double((int)(0x3FF40000u | (kk[c] >> 14)), (int)kk[c]);  // 1 + u, u in [0.25, 0.5)
                r[c] = MODE == 0 ? ndtri_central_b(b) : central_nodiv(b);
            } else {
                unsigned hi, lo;
                hash_to_b(f[c] + it, hi, lo);
                const double b = __hiloint2double((int)(hi | 0x3FF00000u), (int)lo);
                r[c] = MODE == 2 ? ndtri_central_b(b) : b;
            }
        }
#pragma unroll
        for (int c = 0; c < CH; c++) acc += r[c];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// central branch (LCG uniforms) + NA ALU-pipe instructions (xor-shift) + NM FMA-pipe instructions (IMAD) per variate
template <int NA, int NM, int NW>
__global__ void __launch_bounds__(256) kx(double* out, int iters, unsigned long long seed) {
    constexpr int CH = 4;
    unsigned kk[CH], a[CH], m[CH];
    unsigned long long w[CH];
#pragma unroll
    for (int c = 0; c < CH; c++) { kk[c] = (unsigned)seed + threadIdx.x * 977u + c * 131u; a[c] = kk[c] * 3u; m[c] = kk[c] * 5u; w[c] = kk[c]; }
    const unsigned bb = (unsigned)seed * 2654435761u + threadIdx.x;
    double acc = 0.0;
    for (int it = 0; it < iters; it++) {
        double r[CH];
#pragma unroll
        for (int c = 0; c < CH; c++) {
            kk[c] = kk[c] * 1664525u + 1013904223u;
            const double b = __hiloint2double((int)(0x3FF40000u | (kk[c] >> 14)), (int)kk[c]);
#pragma unroll
            for (int n = 0; n < NA / 2; n++) {
                asm volatile("shf.r.wrap.b32 %0, %0, %1, 13;" : "+r"(a[c]) : "r"(bb));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(bb), "r"(bb + 1));
            }
#pragma unroll
            for (int n = 0; n < NM; n++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(m[c]) : "r"(bb), "r"(bb + 3));
#pragma unroll
            for (int n = 0; n < NW; n++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(m[c]), "r"(bb));
            r[c] = ndtri_central_b(b);
        }
#pragma unroll
        for (int c = 0; c < CH; c++) acc += r[c];
    }
#pragma unroll
    for (int c = 0; c < CH; c++) acc += (double)(a[c] + m[c] + (unsigned)w[c]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int NA, int NM, int NW>
void runx(double* d, int sms, int ctas) {
    const int iters = 4000, blocks = sms * ctas;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    kx<NA, NM, NW><<<blocks, 256>>>(d, 10, 1); cudaDeviceSynchronize();
    cudaEventRecord(a); kx<NA, NM, NW><<<blocks, 256>>>(d, iters, 1); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double v = (double)blocks * 256 * iters * 4;
    printf("central + %2d alu + %2d imad + %d imad.wide  ctas/sm=%d  %.3e variates/s  %.1f cycles per warp-variate per scheduler\n", NA, NM, NW, ctas,
           v / (ms * 1e-3), 1.965e9 * 148 * 4 / (v / 32 / (ms * 1e-3)));
}

// the same with the 14 polynomial coefficients held in REGISTERS (loaded from global memory once)
__device__ __forceinline__ double central_regs(double b, const double (&k)[15]) {
    const double y = __dadd_rn(b, -1.5);
    const double y2 = __dmul_rn(y, y);
    double p = k[0];
    TTSK_HORNER(p, y2, k[1]); TTSK_HORNER(p, y2, k[2]); TTSK_HORNER(p, y2, k[3]); TTSK_HORNER(p, y2, k[4]);
    double q = __dadd_rn(y2, k[5]);
    TTSK_HORNER(q, y2, k[6]); TTSK_HORNER(q, y2, k[7]); TTSK_HORNER(q, y2, k[8]); TTSK_HORNER(q, y2, k[9]);
    TTSK_HORNER(q, y2, k[10]); TTSK_HORNER(q, y2, k[11]); TTSK_HORNER(q, y2, k[12]);
    const double t = div_rn_safe(__dmul_rn(y2, p), q);
    return __dmul_rn(__dadd_rn(y, __dmul_rn(y, t)), k[13]);
}
template <int NA>
__global__ void __launch_bounds__(256) kr(double* out, int iters, unsigned long long seed, const double* coef) {
    constexpr int CH = 4;
    double k[15];
#pragma unroll
    for (int i = 0; i < 15; i++) k[i] = coef[i];
    unsigned kk[CH], a[CH];
#pragma unroll
    for (int c = 0; c < CH; c++) { kk[c] = (unsigned)seed + threadIdx.x * 977u + c * 131u; a[c] = kk[c] * 3u; }
    const unsigned bb = (unsigned)seed * 2654435761u + threadIdx.x;
    double acc = 0.0;
    for (int it = 0; it < iters; it++) {
        double r[CH];
#pragma unroll
        for (int c = 0; c < CH; c++) {
            kk[c] = kk[c] * 1664525u + 1013904223u;
            const double b = __hiloint2double((int)(0x3FF40000u | (kk[c] >> 14)), (int)kk[c]);
#pragma unroll
            for (int n = 0; n < NA / 2; n++) {
                asm volatile("shf.r.wrap.b32 %0, %0, %1, 13;" : "+r"(a[c]) : "r"(bb));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(bb), "r"(bb + 1));
            }
            r[c] = central_regs(b, k);
        }
#pragma unroll
        for (int c = 0; c < CH; c++) acc += r[c];
    }
#pragma unroll
    for (int c = 0; c < CH; c++) acc += (double)a[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int NA>
void runr(double* d, int sms, int ctas) {
    const int iters = 4000, blocks = sms * ctas;
    double h[15] = {-5.99633501014107895267E1, 9.80010754185999661536E1, -5.66762857469070293439E1, 1.39312609387279679503E1,
                    -1.23916583867381258016E0, 1.95448858338141759834E0, 4.67627912898881538453E0, 8.63602421390890590575E1,
                    -2.25462687854119370527E2, 2.00260212380060660359E2, -8.20372256168333339912E1, 1.59056225126211695515E1,
                    -1.18331621121330003142E0, 2.50662827463100050242E0, 0};
    double* dc; cudaMalloc(&dc, sizeof(h)); cudaMemcpy(dc, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    kr<NA><<<blocks, 256>>>(d, 10, 1, dc); cudaDeviceSynchronize();
    cudaEventRecord(a); kr<NA><<<blocks, 256>>>(d, iters, 1, dc); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double v = (double)blocks * 256 * iters * 4;
    printf("central(register coefficients) + %2d alu  ctas/sm=%d  %.3e variates/s  %.1f cycles per warp-variate per scheduler\n", NA, ctas,
           v / (ms * 1e-3), 1.965e9 * 148 * 4 / (v / 32 / (ms * 1e-3)));
}

// software-pipelined: the hashes of iteration it+1 are independent of the central branches of iteration it, so
// one loop body holds both and the scheduler can fill the FP64 pipe's issue gaps with the integer work
template <int CH>
__global__ void __launch_bounds__(256) kp(double* out, int iters, unsigned long long seed) {
    unsigned long long f[CH];
    unsigned hi[CH], lo[CH];
#pragma unroll
    for (int c = 0; c < CH; c++) { f[c] = seed * (c + 3) + threadIdx.x + 977ull * blockIdx.x; hash_to_b(f[c], hi[c], lo[c]); }
    double acc = 0.0;
    for (int it = 1; it <= iters; it++) {
        double r[CH];
        unsigned nhi[CH], nlo[CH];
#pragma unroll
        for (int c = 0; c < CH; c++) {
            hash_to_b(f[c] + it, nhi[c], nlo[c]);
            r[c] = ndtri_central_b(__hiloint2double((int)(hi[c] | 0x3FF00000u), (int)lo[c]));
        }
#pragma unroll
        for (int c = 0; c < CH; c++) { acc += r[c]; hi[c] = nhi[c]; lo[c] = nlo[c]; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int CH>
void runp(const char* name, double* d, int sms, int ctas, double fp64) {
    const int iters = 4000, blocks = sms * ctas;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    kp<CH><<<blocks, 256>>>(d, 10, 1); cudaDeviceSynchronize();
    cudaEventRecord(a); kp<CH><<<blocks, 256>>>(d, iters, 1); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double v = (double)blocks * 256 * iters * CH;
    printf("%-12s chains=%d ctas/sm=%d %8.3f ms  %.3e variates/s  fp64/s %.3e (%.0f%% of 1.85e13)\n", name, CH, ctas, ms,
           v / (ms * 1e-3), v * fp64 / (ms * 1e-3), 100 * v * fp64 / (ms * 1e-3) / 1.85e13);
}

template <int MODE, int CH>
void run(const char* name, double* d, int sms, int ctas, double fp64) {
    const int iters = 4000, blocks = sms * ctas;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE, CH><<<blocks, 256>>>(d, 10, 1); cudaDeviceSynchronize();
    cudaEventRecord(a); k<MODE, CH><<<blocks, 256>>>(d, iters, 1); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double v = (double)blocks * 256 * iters * CH;
    printf("%-12s chains=%d ctas/sm=%d %8.3f ms  %.3e variates/s  fp64/s %.3e (%.0f%% of 1.85e13)\n", name, CH, ctas, ms,
           v / (ms * 1e-3), v * fp64 / (ms * 1e-3), 100 * v * fp64 / (ms * 1e-3) / 1.85e13);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* d; cudaMalloc(&d, (size_t)p.multiProcessorCount * 8 * 256 * 8);
    const int sms = p.multiProcessorCount;
    runx<0, 0, 0>(d, sms, 4); runx<8, 0, 0>(d, sms, 4); runx<16, 0, 0>(d, sms, 4); runx<24, 0, 0>(d, sms, 4);
    runx<0, 8, 0>(d, sms, 4); runx<0, 16, 0>(d, sms, 4); runx<0, 0, 2>(d, sms, 4); runx<0, 0, 4>(d, sms, 4); runx<14, 8, 2>(d, sms, 4);
    runr<0>(d, sms, 2); runr<8>(d, sms, 2); runr<16>(d, sms, 2); runr<24>(d, sms, 2);
    for (int ctas : {4}) {
        run<0, 4>("central", d, sms, ctas, 38);
        run<1, 4>("central-div", d, sms, ctas, 32);
        run<2, 4>("hash+central", d, sms, ctas, 38);
        run<3, 4>("hash", d, sms, ctas, 1);
        runp<4>("pipelined", d, sms, ctas, 38);
        runp<2>("pipelined", d, sms, ctas, 38);
        run<0, 2>("central", d, sms, ctas, 38);
        run<2, 2>("hash+central", d, sms, ctas, 38);
    }
    return 0;
}
