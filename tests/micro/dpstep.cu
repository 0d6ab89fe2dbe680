// Micro-benchmark of the double-precision range step (cr_rc.cuh): latency of one chain and throughput of K chains per thread with W warps
// per CTA, one CTA per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dpstep dpstep.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define MAGIC 4503599627370495.5
#define TWO52 4503599627370496.0
template <int K, int FUSED>
__global__ void k_step(const double2* __restrict__ rec, int nrec, int iters, double* out, long long* cycles) {
    extern __shared__ double2 srec[];
    for (int i = threadIdx.x; i < nrec; i += blockDim.x) srec[i] = rec[i];
    __syncthreads();
    double d[K];
    for (int k = 0; k < K; k++) d[k] = 8589934590.0 - 2.0 * (threadIdx.x * K + k);
    uint32_t c418 = 0x41800000u; asm volatile("" : "+r"(c418));
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        double2 a = srec[0];
#pragma unroll 4
        for (int s = 0; s < nrec - 1; s++) {
            const double2 na = srec[s + 1];
            const double nf = -(TWO52 * a.y);
#pragma unroll
            for (int k = 0; k < K; k++) {
                const double T = fma(d[k], a.x, MAGIC);
                const double C = fma(T, a.y, nf);
                uint32_t ch = (uint32_t)__double2hiint(C);
                if (FUSED) asm("lop3.b32 %0, %1, 0x007FFFFF, %2, 0xEA;" : "=r"(ch) : "r"(ch), "r"(c418));
                else ch = (ch & 0x007FFFFFu) | 0x41800000u;
                d[k] = __hiloint2double((int)ch, __double2loint(C));
            }
            a = na;
        }
    }
    long long t1 = clock64();
    double acc = 0; for (int k = 0; k < K; k++) acc += d[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
__global__ void k_dfma_lat(double a, double b, int n, double* out, long long* cycles) {
    double x = a;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; i++) x = fma(x, b, a);
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) cycles[0] = t1 - t0;
}
template <int K, int FUSED> void run(const double2* drec, int nrec, int warps, int ctas, double* dout, long long* dcyc) {
    const int iters = 40;
    k_step<K, FUSED><<<ctas, warps * 32, nrec * 16>>>(drec, nrec, iters, dout, dcyc);
    cudaDeviceSynchronize();
    k_step<K, FUSED><<<ctas, warps * 32, nrec * 16>>>(drec, nrec, iters, dout, dcyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, dcyc, 8, cudaMemcpyDeviceToHost);
    const double steps = (double)iters * (nrec - 1);
    printf("K=%d fused=%d warps/CTA=%2d CTAs/SM=%d: %.1f cycles per step, %.2f state-steps/clk/SM\n", K, FUSED, warps, ctas / 148, c / steps, (double)K * warps * 32 * (ctas / 148) / (c / steps));
}
int main() {
    const int nrec = 1025;
    double2* h = new double2[nrec];
    uint32_t seed = 12345;
    for (int i = 0; i < nrec; i++) {
        seed = seed * 1664525u + 1013904223u; uint32_t sum = 64 + (seed >> 8) % 300000;
        seed = seed * 1664525u + 1013904223u; uint32_t frq = 1 + (seed >> 8) % sum;
        double inv = 1.0 / sum; unsigned long long b; memcpy(&b, &inv, 8); b += 2; b -= 1ull << 52; memcpy(&inv, &b, 8);
        h[i].x = inv; h[i].y = 2.0 * frq;
    }
    double2* drec; double* dout; long long* dcyc;
    cudaMalloc(&drec, nrec * 16); cudaMalloc(&dout, 148 * 8 * 1024 * 8); cudaMalloc(&dcyc, 148 * 8 * 8);
    cudaMemcpy(drec, h, nrec * 16, cudaMemcpyHostToDevice);
    k_dfma_lat<<<1, 32>>>(1.0000001, 0.9999999, 4096, dout, dcyc); cudaDeviceSynchronize();
    k_dfma_lat<<<1, 32>>>(1.0000001, 0.9999999, 4096, dout, dcyc); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, dcyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent DFMA: %.2f cycles\n", c / 4096.0);
    run<1, 0>(drec, nrec, 1, 148, dout, dcyc); run<1, 1>(drec, nrec, 1, 148, dout, dcyc);
    run<2, 1>(drec, nrec, 1, 148, dout, dcyc); run<4, 1>(drec, nrec, 1, 148, dout, dcyc); run<8, 1>(drec, nrec, 1, 148, dout, dcyc);
    run<4, 1>(drec, nrec, 4, 148, dout, dcyc); run<8, 1>(drec, nrec, 4, 148, dout, dcyc); run<4, 1>(drec, nrec, 8, 148, dout, dcyc);
    run<8, 1>(drec, nrec, 8, 148, dout, dcyc); run<4, 1>(drec, nrec, 16, 148, dout, dcyc); run<8, 1>(drec, nrec, 16, 148, dout, dcyc);
    run<4, 1>(drec, nrec, 32, 148, dout, dcyc); run<2, 1>(drec, nrec, 32, 148, dout, dcyc); run<16, 1>(drec, nrec, 8, 148, dout, dcyc);
    run<4, 0>(drec, nrec, 16, 148, dout, dcyc);
    return 0;
}
