// Micro-benchmark + exactness check (not a test, not shipped): a DOUBLE-PRECISION formulation of the range recurrence
//     q = floor(r / sum);  r' = q * frq;  while (r' < 2^24) r' <<= 8          (src/cr-rangecoder.c:60-70, 44-58)
// with two dependent DFMAs per symbol instead of IMAD.HI -> IMAD.WIDE -> SHF -> IMAD (+ FLO) -- see
// profiles/round1_chain_latency.md: the shipping chain costs 44 cycles per symbol, DFMA has 9.3 cycles of latency on B200.
//
//   T   = fma(R, inv, 2^52 - 0.5)            inv = the double just ABOVE 1/sum  ->  T == 2^52 + floor(R / sum)  exactly
//   C_k = fma(T, frq * 256^k, -(2^52 * frq * 256^k))  == q * frq * 256^k         exactly (k = 0, 1, 2 side by side)
//   R'  = C_0 >= 2^24 ? C_0 : C_0 >= 2^16 ? C_1 : C_2                            (q >= 2^8 as long as sum <= 2^16, so two shifts at most)
// Variant 2 (no candidates, no floating-point compare): renormalise C_0 by bumping its exponent field with three integer operations on
// its high word:  hi' = hi + ((0x41EFFFFF - hi) & 0x01800000)   (adds 8 * ((31 - floor(log2 C_0)) / 8) to the exponent, i.e. 0, 8 or 16).
// q is the low word of T (no conversion), the shift count of every symbol is a function of q * frq and is recomputed in parallel
// afterwards, so the serial chain carries R only.  Why T is exact: R < 2^32 and inv - 1/sum < 2^-51 / sum give
// R * inv = R / sum + e with 0 < e < 2^-19 / sum; the fractional part of R / sum is j / sum with 0 <= j <= sum - 1, so the
// fractional part of R * inv lies in (0, 1) strictly and fma's single rounding of  q + frac - 0.5 + 2^52  (ulp 1) lands on 2^52 + q.
//
// build + run the CPU exactness check here:   g++ -O2 -DCHAIN_DP_HOST_ONLY -x c++ -o chain_dp_check chain_dp.cu && ./chain_dp_check
// build + time on the GPU box:                nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chain_dp chain_dp.cu && ./chain_dp
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

struct RecD { double inv, f0, f1, f2; };          // 32 bytes per symbol (the shipping chain input is 16)
static const double TWO52 = 4503599627370496.0;

static RecD make_rec(uint32_t frq, uint32_t sum) {
    RecD r;
    double inv = 1.0 / (double)sum;                // correctly rounded; bump to strictly above 1/sum (two ulps cover the power-of-two case)
    inv = std::nextafter(std::nextafter(inv, 2.0), 2.0);
    r.inv = inv; r.f0 = (double)frq; r.f1 = (double)frq * 256.0; r.f2 = (double)frq * 65536.0;
    return r;
}

static double bump_exponent(double c0) {             // variant 2: C_0 * 256^k by integer arithmetic on the high word
    uint64_t b; memcpy(&b, &c0, 8);
    uint32_t hi = (uint32_t)(b >> 32);
    hi += (0x41EFFFFFu - hi) & 0x01800000u;
    b = (uint64_t)hi << 32 | (uint32_t)b;
    memcpy(&c0, &b, 8);
    return c0;
}

#ifdef CHAIN_DP_HOST_ONLY
// ------------------------------------------------------------------ CPU: the formulation against the integer recurrence
int main() {
    uint64_t s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    long long checked = 0, bad = 0;
    // (1) random states and models, including the corners
    for (int it = 0; it < 40000000; it++) {
        uint32_t sum = (it & 7) == 0 ? 1u << (rnd() % 17) : (uint32_t)(1 + rnd() % 65536);
        if (sum > 65536) sum = 65536;
        uint32_t frq = (uint32_t)(1 + rnd() % sum);
        uint32_t r = (it & 15) == 0 ? (uint32_t)((rnd() % (4294967296ull / sum)) * sum) : (uint32_t)rnd();      // exact multiples of sum too
        if (r < (1u << 24)) r |= 1u << 24;
        if ((it & 1023) == 0) r = 0xFFFFFFFFu;
        const uint32_t q = r / sum;
        uint32_t rn = q * frq; int k = 0;
        while (rn < (1u << 24)) { rn <<= 8; k++; }
        const RecD t = make_rec(frq, sum);
        const double R = (double)r;
        const double T = std::fma(R, t.inv, TWO52 - 0.5);
        uint64_t bits; memcpy(&bits, &T, 8);
        const uint32_t qd = (uint32_t)bits;                                    // low word of the mantissa
        const double C0 = std::fma(T, t.f0, -(TWO52 * t.f0)), C1 = std::fma(T, t.f1, -(TWO52 * t.f1)), C2 = std::fma(T, t.f2, -(TWO52 * t.f2));
        const double Rn = C0 >= 16777216.0 ? C0 : C0 >= 65536.0 ? C1 : C2;
        checked++;
        if (qd != q || T != TWO52 + (double)q || (k <= 2 && (Rn != (double)rn || bump_exponent(C0) != (double)rn))) {
            if (bad < 10) printf("MISMATCH r=%u sum=%u frq=%u: q=%u qd=%u rn=%u Rn=%.1f k=%d\n", r, sum, frq, q, qd, rn, Rn, k);
            bad++;
        }
    }
    // (2) long walks of the recurrence itself
    for (int walk = 0; walk < 64; walk++) {
        uint32_t r = 0xFFFFFFFFu; double R = 4294967295.0;
        for (int i = 0; i < 1000000; i++) {
            const int c = 7 + (int)(rnd() % 9);
            uint32_t sum = (1u << c) + (uint32_t)(rnd() % (1u << c)); if (sum > 65536) sum = 65536;
            uint32_t frq = (rnd() % 8 == 0) ? (uint32_t)(1 + rnd() % 16) : sum / 2 + (uint32_t)(rnd() % (sum / 2));
            if (frq > sum) frq = sum;
            const uint32_t q = r / sum; r = q * frq; while (r < (1u << 24)) r <<= 8;
            const RecD t = make_rec(frq, sum);
            const double T = std::fma(R, t.inv, TWO52 - 0.5);
            const double C0 = std::fma(T, t.f0, -(TWO52 * t.f0)), C1 = std::fma(T, t.f1, -(TWO52 * t.f1)), C2 = std::fma(T, t.f2, -(TWO52 * t.f2));
            R = C0 >= 16777216.0 ? C0 : C0 >= 65536.0 ? C1 : C2;
            checked++;
            if (R != bump_exponent(C0)) { printf("VARIANT 2 MISMATCH at %d\n", i); bad++; break; }
            if (R != (double)r) { if (bad < 10) printf("WALK MISMATCH at %d: r=%u R=%.1f\n", i, r, R); bad++; break; }
        }
    }
    printf("%lld steps checked, %lld mismatches\n", checked, bad);
    return bad != 0;
}
#else
// ------------------------------------------------------------------ GPU: cycles per symbol, one lane, records staged in shared memory
#include <cuda_runtime.h>
#define NREC 1024
#define REPS 64
__global__ void bench_dp2(const RecD* recs, double* out, long long* cyc) {
    __shared__ RecD stage[NREC + 8];
    __shared__ double oq[NREC];
    for (int i = threadIdx.x; i < NREC + 8; i += blockDim.x) stage[i] = recs[i % NREC];
    __syncthreads();
    double R = 4294967295.0, acc = 0.0;
    const double m = TWO52 - 0.5;
    long long t0 = clock64();
    if (threadIdx.x == 0) {
        for (int rep = 0; rep < REPS; rep++) {
            RecD p[8];
#pragma unroll
            for (int u = 0; u < 8; u++) p[u] = stage[u];
#pragma unroll 1
            for (uint32_t j0 = 0; j0 < NREC; j0 += 8) {
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const RecD t = p[u];
                    p[u] = stage[j0 + u + 8];
                    const double T = fma(R, t.inv, m);
                    const double C0 = fma(T, t.f0, t.f1);                      // f1 holds -(2^52 * frq) for this variant
                    int hi = __double2hiint(C0);
                    hi += (0x41EFFFFF - hi) & 0x01800000;
                    R = __hiloint2double(hi, __double2loint(C0));
                    oq[j0 + u] = T;
                }
            }
            acc += R;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = acc + oq[5]; cyc[0] = t1 - t0; }
}
__global__ void bench_dp(const RecD* recs, double* out, long long* cyc) {
    __shared__ RecD stage[NREC + 8];
    __shared__ double oq[NREC];
    for (int i = threadIdx.x; i < NREC + 8; i += blockDim.x) stage[i] = recs[i % NREC];
    __syncthreads();
    double R = 4294967295.0, acc = 0.0;
    const double m = TWO52 - 0.5;
    long long t0 = clock64();
    if (threadIdx.x == 0) {
        for (int rep = 0; rep < REPS; rep++) {
            RecD p[8];
#pragma unroll
            for (int u = 0; u < 8; u++) p[u] = stage[u];
#pragma unroll 1
            for (uint32_t j0 = 0; j0 < NREC; j0 += 8) {
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const RecD t = p[u];
                    p[u] = stage[j0 + u + 8];
                    const double T = fma(R, t.inv, m);
                    const double C0 = fma(T, t.f0, -(TWO52 * t.f0)), C1 = fma(T, t.f1, -(TWO52 * t.f1)), C2 = fma(T, t.f2, -(TWO52 * t.f2));   // addends: precompute in the record for the real kernel
                    R = C0 >= 16777216.0 ? C0 : C0 >= 65536.0 ? C1 : C2;
                    oq[j0 + u] = T;
                }
            }
            acc += R;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = acc + oq[5]; cyc[0] = t1 - t0; }
}
int main() {
    std::vector<RecD> r(NREC);
    srand(1);
    for (int i = 0; i < NREC; i++) {
        const int c = 7 + rand() % 8;
        const uint32_t sum = (1u << c) + rand() % (1u << c);
        uint32_t frq = (rand() % 8 == 0) ? 1 + rand() % 16 : sum / 2 + rand() % (sum / 2);
        if (frq > sum) frq = sum;
        r[i] = make_rec(frq, sum);
    }
    RecD* d; double* out; long long* cyc; cudaMalloc(&d, NREC * sizeof(RecD)); cudaMalloc(&out, 64); cudaMalloc(&cyc, 64);
    cudaMemcpy(d, r.data(), NREC * sizeof(RecD), cudaMemcpyHostToDevice);
    for (int k = 0; k < 2; k++) bench_dp<<<1, 32>>>(d, out, cyc);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DP chain 1 (2 dependent DFMA + 2 DSETP + 2 select levels)   %.2f cycles/symbol (%s)   [shipping integer chain: 44.0]\n", (double)h / ((double)NREC * REPS), cudaGetErrorString(cudaGetLastError()));
    for (auto& x : r) x.f1 = -(TWO52 * x.f0);
    cudaMemcpy(d, r.data(), NREC * sizeof(RecD), cudaMemcpyHostToDevice);
    for (int k = 0; k < 2; k++) bench_dp2<<<1, 32>>>(d, out, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DP chain 2 (2 dependent DFMA + 3 integer ops on the exponent)  %.2f cycles/symbol (%s)\n", (double)h / ((double)NREC * REPS), cudaGetErrorString(cudaGetLastError()));
    return 0;
}
#endif
