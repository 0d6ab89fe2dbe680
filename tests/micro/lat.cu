// Micro-benchmark (not a test, not shipped): dependent-issue latency of the integer ops the range chain is made of, one lane.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 4096
template <int OP> __global__ void lat(uint32_t* out, long long* cyc, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    uint32_t x = a + threadIdx.x, y = b + threadIdx.x;        // per-thread values: keeps the ops off the uniform datapath
    unsigned long long z = ((unsigned long long)(c + threadIdx.x) << 32) | d;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N / 16; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            if (OP == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
            if (OP == 1) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x) : "r"(b));
            if (OP == 2) asm volatile("{.reg .u32 l, h; mov.b64 {l, h}, %0; mad.wide.u32 %0, l, %1, %0;}" : "+l"(z) : "r"(b));
            if (OP == 3) asm volatile("{.reg .u32 l, h; mov.b64 {l, h}, %0; mad.wide.u32 %0, h, %1, %0;}" : "+l"(z) : "r"(b));
            if (OP == 4) asm volatile("shf.r.clamp.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
            if (OP == 5) asm volatile("shf.r.clamp.b32 %0, %1, %2, %0;" : "+r"(x) : "r"(b), "r"(c));      // dependent through the shift amount
            if (OP == 6) asm volatile("bfind.u32 %0, %0;" : "+r"(x));
            if (OP == 7) asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %2, %3, p;}" : "+r"(x) : "r"(b), "r"(c), "r"(d));
            if (OP == 8) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(b));
            if (OP == 9) asm volatile("{.reg .u32 t; mul.hi.u32 t, %0, %1; mad.lo.u32 %0, t, %2, %3;}" : "+r"(x) : "r"(b), "r"(c), "r"(d));   // hi -> lo
            if (OP == 10) asm volatile("{.reg .u32 t; shf.r.clamp.b32 t, %0, %1, %2; mad.lo.u32 %0, t, %2, %3;}" : "+r"(x) : "r"(b), "r"(c), "r"(d)); // alu -> fma
            if (OP == 11) asm volatile("popc.b32 %0, %0;" : "+r"(x));
            if (OP == 12) asm volatile("{.reg .u32 l, h; mov.b64 {l, h}, %0; mad.wide.u32 %0, l, %1, %0; }" : "+l"(z) : "r"(y));   // same as 2 with y
            if (OP == 13) asm volatile("{.reg .u32 l, h; mul.wide.u32 %0, %1, %2; mov.b64 {l, h}, %0; mov.u32 %1, h;}" : "+l"(z), "+r"(x) : "r"(b));  // wide, hi half feeds next
            if (OP == 14) asm volatile("{.reg .f32 f; cvt.rn.f32.u32 f, %0; mov.b32 %0, f;}" : "+r"(x));
            if (OP == 15) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
            if (OP == 16) asm volatile("{.reg .f32 f; mov.b32 f, %0; fma.rn.f32 f, f, f, f; mov.b32 %0, f;}" : "+r"(x));
            if (OP == 18) asm volatile("{.reg .u32 t; mul.hi.u32 t, %0, %1; shr.u32 t, t, %2; mul.lo.u32 %0, t, %3;}" : "+r"(x) : "r"(b), "r"(c), "r"(d));   // the v7 chain
            if (OP == 19) asm volatile("{.reg .u32 t; mul.hi.u32 t, %0, %1; shr.u32 %0, t, %2;}" : "+r"(x) : "r"(b), "r"(c));
            if (OP == 20) asm volatile("{.reg .u32 t; mul.lo.u32 t, %0, %1; shr.u32 %0, t, %2;}" : "+r"(x) : "r"(b), "r"(c));
            if (OP == 21) asm volatile("{.reg .u32 t; mul.lo.u32 t, %0, %1; mul.hi.u32 %0, t, %2;}" : "+r"(x) : "r"(b), "r"(c));
            if (OP == 22) asm volatile("{.reg .u32 t; or.b32 t, %0, %1; mul.hi.u32 %0, t, %2;}" : "+r"(x) : "r"(b), "r"(c));
            if (OP == 23) asm volatile("{.reg .u32 l, h; mul.wide.u32 %0, %1, %2; mov.b64 {l, h}, %0; mov.u32 %1, l;}" : "+l"(z), "+r"(x) : "r"(b));  // wide, lo half feeds next
            if (OP == 24) asm volatile("{.reg .u32 t; mul.hi.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(x) : "r"(b), "r"(c));
            if (OP == 25) asm volatile("{.reg .u32 t; mul24.lo.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(x) : "r"(b), "r"(c));
            if (OP == 26) asm volatile("{.reg .u32 t; mul.hi.u32 t, %0, %1; mul.hi.u32 %0, t, %2;}" : "+r"(x) : "r"(b), "r"(c));
            if (OP == 27) asm volatile("{.reg .f32 f, g; mov.b32 f, %0; mov.b32 g, %1; mul.rn.f32 f, f, g; mov.b32 %0, f;}" : "+r"(x) : "r"(b));
            if (OP == 28) asm volatile("{.reg .f32 f; cvt.rn.f32.u32 f, %0; cvt.rzi.u32.f32 %0, f;}" : "+r"(x));
            if (OP == 17) asm volatile("{.reg .f64 f; mov.b64 f, %0; fma.rn.f64 f, f, f, f; mov.b64 %0, f;}" : "+l"(z));
        }
    }
    long long t1 = clock64();
    out[0] = x + (uint32_t)z + (uint32_t)(z >> 32); cyc[0] = t1 - t0;
}
template <int OP> void run(const char* name, uint32_t* out, long long* cyc) {
    lat<OP><<<1, 32>>>(out, cyc, 0x12345678u, 0x9abcdef1u, 7, 3);
    lat<OP><<<1, 32>>>(out, cyc, 0x12345678u, 0x9abcdef1u, 7, 3);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s %.2f cycles\n", name, (double)h / N);
}
int main() {
    uint32_t* out; long long* cyc; cudaMalloc(&out, 64); cudaMalloc(&cyc, 64);
    run<8>("IADD", out, cyc); run<0>("IMAD lo", out, cyc); run<1>("IMAD.HI", out, cyc); run<2>("IMAD.WIDE (lo feeds)", out, cyc); run<3>("IMAD.WIDE (hi feeds)", out, cyc);
    run<13>("MUL.WIDE hi feeds", out, cyc);
    run<4>("SHF (data)", out, cyc); run<5>("SHF (amount)", out, cyc); run<6>("FLO/bfind", out, cyc); run<7>("ISETP+SEL", out, cyc);
    run<9>("IMAD.HI -> IMAD", out, cyc); run<10>("SHF -> IMAD", out, cyc); run<11>("POPC", out, cyc); run<14>("I2F", out, cyc); run<15>("PRMT", out, cyc);
    run<16>("FFMA", out, cyc); run<17>("DFMA", out, cyc);
    run<18>("IMAD.HI -> SHF -> IMAD", out, cyc); run<19>("IMAD.HI -> SHF", out, cyc); run<20>("IMAD -> SHF", out, cyc); run<21>("IMAD -> IMAD.HI", out, cyc);
    run<22>("LOP3 -> IMAD.HI", out, cyc); run<23>("MUL.WIDE lo feeds", out, cyc); run<24>("IMAD.HI -> IADD", out, cyc); run<25>("MUL24 -> IADD", out, cyc);
    run<26>("IMAD.HI -> IMAD.HI", out, cyc); run<27>("FMUL", out, cyc); run<28>("I2F -> F2I", out, cyc);
    return 0;
}
