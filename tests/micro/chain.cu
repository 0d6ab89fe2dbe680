// Micro-benchmark (not a test, not shipped): cycles per symbol of candidate range-chain formulations, one lane, records in
// shared memory as in k_range_chain.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chain chain.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define NREC 1024
#define REPS 64
struct Rec { uint32_t frq, sum; };

// ---- A: exact 64-bit reciprocal, FLO (VARIANT 4).  t = {frq, M_lo, M_hi, sum}
__device__ __forceinline__ void stepA(uint32_t& range, uint32_t& msb, const uint4 t, uint32_t& q) {
    const uint32_t ahi = __umulhi(range, t.y);
    const unsigned long long S2 = (unsigned long long)range * t.z + ahi;
    const uint32_t amt = (msb & 24u) | 7u;
    q = (uint32_t)(S2 >> amt);
    range = q * t.x;
    asm volatile("bfind.u32 %0, %1;" : "=r"(msb) : "r"(range));
}
// ---- D: single multiply, fast path only (k in {0,1}, c >= 8k assumed; WRONG otherwise: timing only). t = {frq, M', sum, c}
__device__ __forceinline__ void stepD(uint32_t& range, const uint4 t, uint32_t& q) {
    const uint32_t h = __umulhi(range, t.y);
    const uint32_t sh = range < (1u << 24) ? t.w - 8u : t.w;
    q = h >> (sh & 31u);
    range = q * t.x | 0x10000u;      // keep the walk alive although q is not exact
}
// ---- E: two multiplies side by side, predicated shifts, no branch.  t = {frq, M', (a1 << 8 | b1) << 8 | b0, frq << a1(next)}
//         rA = un-normalised range, rB = rA << a1 (a1 = max(8 - c, 0) of THIS symbol), produced by the previous symbol's two IMADs
__device__ __forceinline__ void stepE(uint32_t& rA, uint32_t& rB, const uint4 t, const uint32_t frqB_next, uint32_t& q) {
    uint32_t hA, hB;
    asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(hA) : "r"(rA), "r"(t.y));
    asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(hB) : "r"(rB), "r"(t.y));
    const uint32_t b0 = t.z & 255u, b1 = (t.z >> 8) & 255u;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %1, 16777216;\n\tshr.u32 %0, %2, %3;\n\t@p shr.u32 %0, %4, %5;\n\t}"
                 : "=&r"(q) : "r"(rA), "r"(hA), "r"(b0), "r"(hB), "r"(b1));
    rA = q * t.x | 0x10000u;
    rB = q * frqB_next | 0x10000u;
}
// ---- F: like D plus a deferred check that only accumulates a flag (branch once per 8 symbols)
__device__ __forceinline__ void stepF(uint32_t& range, const uint4 t, uint32_t& q, uint32_t& bad) {
    const uint32_t h = __umulhi(range, t.y);
    const bool p24 = range < (1u << 24);
    const uint32_t sh = p24 ? t.w - 8u : t.w;
    q = h >> (sh & 31u);
    const uint32_t rem = (range << (p24 ? 8u : 0u)) - q * t.z;
    bad |= (rem >= t.z) | (range < (1u << 16));
    range = q * t.x | 0x10000u;
}

// ---- D variants for attribution
__device__ __forceinline__ void stepD2(uint32_t& range, const uint4 t, uint32_t& q) {      // shift amount through one selp (no predicated add)
    uint32_t h, sh;
    asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(h) : "r"(range), "r"(t.y));
    asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %1, 16777216;\n\tselp.u32 %0, %2, %3, p;\n\t}" : "=r"(sh) : "r"(range), "r"(t.z), "r"(t.w));
    q = h >> sh;
    range = q * t.x | 0x10000u;
}
__device__ __forceinline__ void stepD3(uint32_t& range, const uint4 t, uint32_t& q) {      // no compare at all: static shift
    const uint32_t h = __umulhi(range, t.y);
    q = h >> t.w;
    range = q * t.x | 0x10000u;
}
__device__ __forceinline__ void stepD4(uint32_t& range, const uint4 t, uint32_t& q) {      // static shift, no OR
    const uint32_t h = __umulhi(range, t.y);
    q = h >> t.w;
    range = q * t.x;
}
__device__ __noinline__ uint32_t slow_fix(uint32_t r, uint32_t sum) { return r / sum; }
// D2 + a rarely taken branch per symbol
__device__ __forceinline__ void stepG1(uint32_t& range, const uint4 t, uint32_t& q, const uint32_t g_thr) {
    uint32_t h, sh;
    asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(h) : "r"(range), "r"(t.y));
    asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %1, 16777216;\n\tselp.u32 %0, %2, %3, p;\n\t}" : "=r"(sh) : "r"(range), "r"(t.z), "r"(t.w));
    q = h >> sh;
    if (__builtin_expect(range < g_thr, 0)) q = slow_fix(range, t.x);
    range = q * t.x | 0x10000u;
}
// D2 + flag accumulation (checked per group by the caller)
__device__ __forceinline__ void stepG3(uint32_t& range, const uint4 t, uint32_t& q, uint32_t& bad, const uint32_t g_thr) {
    uint32_t h, sh;
    asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(h) : "r"(range), "r"(t.y));
    asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %1, 16777216;\n\tselp.u32 %0, %2, %3, p;\n\t}" : "=r"(sh) : "r"(range), "r"(t.z), "r"(t.w));
    q = h >> sh;
    bad |= range < g_thr;
    range = q * t.x | 0x10000u;
}
// ---- X: exact 64-bit reciprocal with three independent multiplies, carry through a predicate, normalisation decided on the
//         quotient against precomputed thresholds (no compare on the product).  a = {frq, Ml, Mh, T1}, b = {T2, sum, frq8, frq16}
template <int MODE> __device__ __forceinline__ void stepX(uint32_t& R, const uint4 a, const uint4 b, uint32_t& q) {
    uint32_t Rn;
    if (MODE == 0)
    asm volatile("{\n\t.reg .pred P, p1, p2;\n\t.reg .u32 q0, l, ah, nl, fs;\n\t"
        "mul.hi.u32 q0, %2, %4;\n\tmul.lo.u32 l, %2, %4;\n\tmul.hi.u32 ah, %2, %3;\n\t"
        "not.b32 nl, l;\n\tsetp.gt.u32 P, ah, nl;\n\tsetp.lt.u32 p1, q0, %5;\n\tsetp.lt.u32 p2, q0, %6;\n\t"
        "selp.u32 fs, %8, %7, p1;\n\tmov.u32 %0, q0;\n\t@P add.u32 %0, q0, 1;\n\t"
        "mul.lo.u32 %1, %0, fs;\n\t@p2 mul.lo.u32 %1, %0, %9;\n\t}"
        : "=&r"(q), "=&r"(Rn) : "r"(R), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(a.x), "r"(b.z), "r"(b.w));
    if (MODE == 1)   // no k = 2 override
    asm volatile("{\n\t.reg .pred P, p1;\n\t.reg .u32 q0, l, ah, nl, fs;\n\t"
        "mul.hi.u32 q0, %2, %4;\n\tmul.lo.u32 l, %2, %4;\n\tmul.hi.u32 ah, %2, %3;\n\t"
        "not.b32 nl, l;\n\tsetp.gt.u32 P, ah, nl;\n\tsetp.lt.u32 p1, q0, %5;\n\t"
        "selp.u32 fs, %8, %7, p1;\n\tmov.u32 %0, q0;\n\t@P add.u32 %0, q0, 1;\n\t"
        "mul.lo.u32 %1, %0, fs;\n\t}"
        : "=&r"(q), "=&r"(Rn) : "r"(R), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(a.x), "r"(b.z), "r"(b.w));
    if (MODE == 2)   // no carry either
    asm volatile("{\n\t.reg .pred p1;\n\t.reg .u32 fs;\n\t"
        "mul.hi.u32 %0, %2, %4;\n\tsetp.lt.u32 p1, %0, %5;\n\t"
        "selp.u32 fs, %8, %7, p1;\n\t"
        "mul.lo.u32 %1, %0, fs;\n\t}"
        : "=&r"(q), "=&r"(Rn) : "r"(R), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(a.x), "r"(b.z), "r"(b.w));
    if (MODE == 3)   // carry, no normalisation select
    asm volatile("{\n\t.reg .pred P;\n\t.reg .u32 q0, l, ah, nl;\n\t"
        "mul.hi.u32 q0, %2, %4;\n\tmul.lo.u32 l, %2, %4;\n\tmul.hi.u32 ah, %2, %3;\n\t"
        "not.b32 nl, l;\n\tsetp.gt.u32 P, ah, nl;\n\tselp.u32 nl, 1, 0, P;\n\tadd.u32 %0, q0, nl;\n\t"
        "mul.lo.u32 %1, %0, %7;\n\t}"
        : "=&r"(q), "=&r"(Rn) : "r"(R), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(a.x), "r"(b.z), "r"(b.w));
    R = Rn | 0x01000000u;
}
// ---- T: exact multiply-add-shift (round-up magic with addend 0, or round-down magic with addend m: one of the two is exact for
//         every divisor), applied to two / three pre-normalised candidates of the range that the previous symbol produced side by
//         side (q * frq, q * frq << 8, q * frq << 16).  a = {frq, m, addend, c}, b = {frq8, frq16, -, -}
template <int NC> __device__ __forceinline__ void stepT(uint32_t& RA, uint32_t& RB, uint32_t& RC, const uint4 a, const uint4 b, uint32_t& q) {
    const unsigned long long add = a.z;
    const uint32_t hA = (uint32_t)(((unsigned long long)RA * a.y + add) >> 32);
    const uint32_t hB = (uint32_t)(((unsigned long long)RB * a.y + add) >> 32);
    uint32_t h;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %1, 16777216;\n\tselp.u32 %0, %2, %3, p;\n\t}" : "=r"(h) : "r"(RA), "r"(hB), "r"(hA));
    if (NC == 3) {
        const uint32_t hC = (uint32_t)(((unsigned long long)RC * a.y + add) >> 32);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %1, 65536;\n\tselp.u32 %0, %2, %0, p;\n\t}" : "+r"(h) : "r"(RA), "r"(hC));
    }
    q = h >> a.w;
    RA = q * a.x | 0x00010000u; RB = q * b.x | 0x01000000u; if (NC == 3) RC = q * b.y | 0x01000000u;
}
template <int V> __global__ void benchT(const uint4* recs, uint32_t* out, long long* cyc) {
    __shared__ uint4 stage[2 * (NREC + 8)];
    __shared__ uint32_t oq[NREC];
    for (int i = threadIdx.x; i < 2 * (NREC + 8); i += blockDim.x) stage[i] = recs[i % (2 * NREC)];
    __syncthreads();
    uint32_t RA = 0xFFFFFFFFu, RB = 0xFFFFFFFFu, RC = 0xFFFFFFFFu, acc = 0;
    long long t0 = clock64();
    if (threadIdx.x == 0) {
        for (int rep = 0; rep < REPS; rep++) {
            uint4 pa[8], pb[8];
#pragma unroll
            for (int u = 0; u < 8; u++) { pa[u] = stage[2 * u]; pb[u] = stage[2 * u + 1]; }
#pragma unroll 1
            for (uint32_t j0 = 0; j0 < NREC; j0 += 8) {
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const uint4 a = pa[u], b = pb[u];
                    pa[u] = stage[2 * (j0 + u + 8)]; pb[u] = stage[2 * (j0 + u + 8) + 1];
                    uint32_t q;
                    stepT<V>(RA, RB, RC, a, b, q);
                    oq[j0 + u] = q;
                }
            }
            acc += RA;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = acc + oq[5]; cyc[0] = t1 - t0; }
}
template <int V> __global__ void benchX(const uint4* recs, uint32_t* out, long long* cyc) {
    __shared__ uint4 stage[2 * (NREC + 8)];
    __shared__ uint32_t oq[NREC];
    for (int i = threadIdx.x; i < 2 * (NREC + 8); i += blockDim.x) stage[i] = recs[i % (2 * NREC)];
    __syncthreads();
    uint32_t R = 0xFFFFFFFFu, acc = 0;
    long long t0 = clock64();
    if (threadIdx.x == 0) {
        for (int rep = 0; rep < REPS; rep++) {
            uint4 pa[8], pb[8];
#pragma unroll
            for (int u = 0; u < 8; u++) { pa[u] = stage[2 * u]; pb[u] = stage[2 * u + 1]; }
#pragma unroll 1
            for (uint32_t j0 = 0; j0 < NREC; j0 += 8) {
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const uint4 a = pa[u], b = pb[u];
                    pa[u] = stage[2 * (j0 + u + 8)]; pb[u] = stage[2 * (j0 + u + 8) + 1];
                    uint32_t q;
                    stepX<V>(R, a, b, q);
                    oq[j0 + u] = q;
                }
            }
            acc += R;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = acc + oq[5]; cyc[0] = t1 - t0; }
}
template <int V> __global__ void bench(const uint4* recs, uint32_t* out, long long* cyc, int active_lanes, uint32_t thr = 4096u) {
    __shared__ uint4 stage[NREC + 8];
    __shared__ uint32_t oq[NREC];
    for (int i = threadIdx.x; i < NREC + 8; i += blockDim.x) stage[i] = recs[i % NREC];
    __syncthreads();
    uint32_t range = 0xFFFFFFFFu, msb = 31, rB = 0xFFFFFFFFu, acc = 0, badcount = 0;
    long long t0 = clock64();
    if ((int)threadIdx.x < active_lanes) {
        for (int rep = 0; rep < REPS; rep++) {
            uint4 p[8];
#pragma unroll
            for (int u = 0; u < 8; u++) p[u] = stage[u];
#pragma unroll 1
            for (uint32_t j0 = 0; j0 < NREC; j0 += 8) {
                uint32_t bad = 0;
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const uint4 t = p[u];
                    p[u] = stage[j0 + u + 8];
                    uint32_t q;
                    if (V == 0) stepA(range, msb, t, q);
                    if (V == 1) stepD(range, t, q);
                    if (V == 2) stepE(range, rB, t, p[(u + 1) & 7].w, q);
                    if (V == 3) stepF(range, t, q, bad);
                    if (V == 4) stepD2(range, t, q);
                    if (V == 5) stepD3(range, t, q);
                    if (V == 6) stepD4(range, t, q);
                    if (V == 7) stepG1(range, t, q, thr);
                    if (V == 8) { stepG3(range, t, q, bad, thr); if ((u & 3) == 3) { if (__builtin_expect(bad != 0, 0)) { range = slow_fix(range, t.x) | 0x10000u; badcount++; } bad = 0; } }
                    if (V == 9) { stepG3(range, t, q, bad, thr); if (u == 7) { if (__builtin_expect(bad != 0, 0)) { range = slow_fix(range, t.x) | 0x10000u; badcount++; } bad = 0; } }
                    if (V != 7) oq[j0 + u] = q;
                }
                if (V == 3 && bad) { badcount++; }
            }
            acc += range;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = acc + oq[5] + badcount; cyc[0] = t1 - t0; }
}
int main() {
    std::vector<uint4> rA(NREC), rD(NREC), rE(NREC);
    srand(1);
    for (int i = 0; i < NREC; i++) {
        const int c = 7 + rand() % 8;
        const uint32_t sum = (1u << c) + rand() % (1u << c);
        uint32_t frq = (rand() % 8 == 0) ? 1 + rand() % 16 : sum / 2 + rand() % (sum / 2);
        if (frq > sum) frq = sum;
        const unsigned long long M = (1ull << 63) / sum + 1;
        unsigned long long Mp = (1ull << (32 + c)) / sum + 1; if (Mp >> 32) Mp = 0xFFFFFFFFull;
        rA[i] = make_uint4(frq, (uint32_t)M, (uint32_t)(M >> 32), sum);
        rD[i] = make_uint4(frq, (uint32_t)Mp, sum, c);
        const uint32_t a1 = c < 8 ? 8 - c : 0, b0 = c, b1 = c >= 8 ? c - 8 : 0;
        rE[i] = make_uint4(frq, (uint32_t)Mp, b1 << 8 | b0, frq << a1);
    }
    std::vector<uint4> rX(2 * NREC);
    for (int i = 0; i < NREC; i++) {
        const uint32_t frq = rA[i].x, sum = rA[i].w;
        const unsigned long long M1 = 2 * ((1ull << 63) / sum + 1);
        rX[2 * i] = make_uint4(frq, (uint32_t)M1, (uint32_t)(M1 >> 32), ((1u << 24) - 1) / frq + 1);
        rX[2 * i + 1] = make_uint4(((1u << 16) - 1) / frq + 1, sum, frq << 8, frq << 16);
    }
    uint4* d; uint32_t* out; long long* cyc; cudaMalloc(&d, NREC * 32); cudaMalloc(&out, 64); cudaMalloc(&cyc, 64);
    const char* names[10] = { "A exact 64-bit + FLO (v4)", "D single mul, fast only", "E dual mul, predicated", "F single mul + flag check", "D2 selp shift", "D3 static shift", "D4 static shift, no OR", "G1 D2 + branch per symbol", "G3 D2 + branch per 4", "G4 D2 + branch per 8" };
    for (int lanes = 1; lanes <= 32; lanes *= 32)
    for (int v = 0; v < 10; v++) {
        std::vector<uint4> r4 = rD; for (auto& x : r4) x.z = x.w - 8;
        const std::vector<uint4>& src = v == 0 ? rA : v == 2 ? rE : (v == 4 || v >= 7) ? r4 : rD;
        cudaMemcpy(d, src.data(), NREC * 16, cudaMemcpyHostToDevice);
        for (int k = 0; k < 2; k++) {
            if (v == 0) bench<0><<<1, 32>>>(d, out, cyc, lanes);
            if (v == 1) bench<1><<<1, 32>>>(d, out, cyc, lanes);
            if (v == 2) bench<2><<<1, 32>>>(d, out, cyc, lanes);
            if (v == 3) bench<3><<<1, 32>>>(d, out, cyc, lanes);
            if (v == 4) bench<4><<<1, 32>>>(d, out, cyc, lanes);
            if (v == 5) bench<5><<<1, 32>>>(d, out, cyc, lanes);
            if (v == 6) bench<6><<<1, 32>>>(d, out, cyc, lanes);
            if (v == 7) bench<7><<<1, 32>>>(d, out, cyc, lanes);
            if (v == 8) bench<8><<<1, 32>>>(d, out, cyc, lanes);
            if (v == 9) bench<9><<<1, 32>>>(d, out, cyc, lanes);
        }
        cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-28s lanes=%2d  %.2f cycles/symbol (%s)\n", names[v], lanes, (double)h / ((double)NREC * REPS), cudaGetErrorString(cudaGetLastError()));
    }
    cudaMemcpy(d, rX.data(), NREC * 32, cudaMemcpyHostToDevice);
    for (int m = 0; m < 4; m++) {
    for (int k = 0; k < 2; k++) { if (m == 0) benchX<0><<<1, 32>>>(d, out, cyc); if (m == 1) benchX<1><<<1, 32>>>(d, out, cyc); if (m == 2) benchX<2><<<1, 32>>>(d, out, cyc); if (m == 3) benchX<3><<<1, 32>>>(d, out, cyc); }
    cudaDeviceSynchronize();
    { long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      // CPU replay of the same records with the plain recurrence, to see how often the optimistic step is right
      uint32_t R = 0xFFFFFFFFu; long long wrong = 0, k2 = 0;
      for (int i = 0; i < NREC; i++) { const uint32_t frq = rA[i].x, sum = rA[i].w; uint32_t q = R / sum, r = q * frq; int k = 0; while (r < (1u << 24)) { r <<= 8; k++; } if (k >= 2) k2++; R = r; }
      printf("%-28s lanes= 1  %.2f cycles/symbol (%s)  [cpu walk: k>=2 on %lld of %d]\n", m == 0 ? "X exact, 3 multiplies" : m == 1 ? "X1 no k=2 override" : m == 2 ? "X2 no carry, no override" : "X3 carry only", (double)h / ((double)NREC * REPS), cudaGetErrorString(cudaGetLastError()), k2, NREC); }
    }
    { std::vector<uint4> rT(2 * NREC);
      for (int i = 0; i < NREC; i++) { const uint32_t frq = rA[i].x, sum = rA[i].w; int c = 31 - __builtin_clz(sum);
        rT[2 * i] = make_uint4(frq, rD[i].y, rD[i].y, c); rT[2 * i + 1] = make_uint4(frq << 8, frq << 16, 0, 0); }
      cudaMemcpy(d, rT.data(), NREC * 32, cudaMemcpyHostToDevice);
      for (int m = 2; m <= 3; m++) {
        for (int k = 0; k < 2; k++) { if (m == 2) benchT<2><<<1, 32>>>(d, out, cyc); else benchT<3><<<1, 32>>>(d, out, cyc); }
        cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("T%d multiply-add-shift, %d candidates  %.2f cycles/symbol (%s)\n", m, m, (double)h / ((double)NREC * REPS), cudaGetErrorString(cudaGetLastError()));
      } }
    return 0;
}
