// cr_rc.cuh -- the carry-propagating 32-bit range coder (src/cr-rangecoder.c:35-79) and payload assembly.
//
// One coder per (block, stream): the reference re-initialises `coder` and `idx_coder` for every block
// (src/rolzmain/cr-coder.c:181-182), so streams are independent chains; inside a stream every step depends on
// the previous range, which makes this the serial floor of the whole encoder.  The triples arrive fully
// resolved from the model passes, and the divisor's reciprocal is precomputed in parallel (k_expand), so the
// dependent chain per symbol is  umulhi -> mul/sub -> compare/add -> mul -> normalise.
#pragma once
#include "cr_common.cuh"
#include <cmath>
#include <cstring>
#include "cr_ppm.cuh"

struct Tri {             // one range_encoder_encode call
    uint32_t cum;
    uint32_t frq;        // bit 31: last triple of a token (the reference tests "cannot compress" there)
    uint32_t sum;
    uint32_t magic;      // floor(2^32 / sum), saturated: q = umulhi(range, magic) is range/sum or one less
};
#define TRI_TOKEND 0x80000000u

CR_HD uint32_t rc_magic(uint32_t sum) { return sum <= 1 ? 0xFFFFFFFFu : (uint32_t)(0x100000000ull / sum); }

__global__ void k_esc_flags(const uint64_t* __restrict__ T1, uint32_t n, uint32_t* __restrict__ flag) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e > n) return;
    flag[e] = e < n ? (uint32_t)(T1[e] >> 63) : 0;       // n+1 entries so that the scan also yields the total
}

// dense main-stream triples: event e lands at e + (#escapes before e); an escape is followed by its o1 triple.
// tokend[e] == 0 marks events that do not close a token (LZP codes a match as two events); nullptr = all close.
__global__ void k_expand_main(const uint64_t* __restrict__ T1, const uint64_t* __restrict__ T2, const uint32_t* __restrict__ escord,
                              const uint8_t* __restrict__ tokend, uint32_t n, Tri* __restrict__ dense) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    uint64_t t = T1[e];
    uint32_t o = escord[e], esc = (uint32_t)(t >> 63);
    uint32_t endflag = (!tokend || tokend[e]) ? TRI_TOKEND : 0;
    Tri a;
    a.cum = (uint32_t)t & 0xFFFFFF; a.frq = (uint32_t)(t >> 24) & 0xFFFF; a.sum = (uint32_t)(t >> 40) & 0x7FFFFF;
    a.magic = rc_magic(a.sum);
    if (!esc) a.frq |= endflag;
    dense[(size_t)e + o] = a;
    if (esc) {
        uint64_t u = T2[o];
        Tri b;
        b.cum = (uint32_t)u & 0xFFFFFF; b.frq = ((uint32_t)(u >> 24) & 0xFFFF) | endflag; b.sum = (uint32_t)(u >> 40) & 0x7FFFFF;
        b.magic = rc_magic(b.sum);
        dense[(size_t)e + o + 1] = b;
    }
}
__global__ void k_expand_side(const uint64_t* __restrict__ TS, uint32_t n, Tri* __restrict__ dense) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t t = TS[i];
    Tri a;
    a.cum = (uint32_t)t & 0xFFFFFF; a.frq = (uint32_t)(t >> 24) & 0xFFFF; a.sum = (uint32_t)(t >> 40) & 0x7FFFFF;
    a.magic = rc_magic(a.sum);
    dense[i] = a;
}

// ------------------------------------------------------------------ double-precision form of the range recurrence (k_range_chain<7>, cr_rcpar.cuh)
// q = floor(r / sum); r' = q * frq; byte renormalisation -- as two dependent FMAs and ONE integer operation.  The state is kept as
// R = 2 * range (a double with an exact integer value in [2^25, 2^33)); the records hold inv/2 and 2*frq, both exact scalings:
//   T = fma(R, inv/2, 2^52 - 0.5) == 2^52 + floor(range / sum) exactly, for inv = two ulps above the correctly rounded 1/sum:  range < 2^32
//       makes range * inv exceed range / sum by less than 2^-19 / sum, the fractional part of range / sum is a multiple of 1/sum below 1, so
//       the product's fractional part lies strictly inside (0, 1) and the single rounding of the FMA (ulp 1 at 2^52) lands on 2^52 + q;
//   C = fma(T, 2*frq, -(2^52 * 2*frq)) == 2 * q * frq exactly (below 2^49);
//   renormalisation: with e = top-bit index of q * frq, the coder shifts by whole bytes until the top bit sits at 24 + (e & 7).  The
//       exponent field of C is 1024 + e, whose low three bits ARE e & 7, so  hi' = (hi & 0x007FFFFF) | 0x41800000  (exponent field
//       1048 + (e & 7), i.e. 2 * range' with range' in [2^24, 2^32)) is the whole renormalisation: one LOP3.  That is why the state
//       carries the factor two: 1047 is not a multiple of eight, 1048 is.
// Measured on B200: see profiles/round2_summary.md (the chain with the three-operation renormalisation: 39.1 cycles per symbol,
// profiles/round1_chain_latency.md).
#define RC_DP_MAGIC 4503599627370495.5         /* 2^52 - 0.5 */
#define RC_DP_TWO52 4503599627370496.0
#define RC_DP_R0    8589934590.0               /* 2 * 0xFFFFFFFF: the coder's initial range (src/cr-rangecoder.c:36) */
CR_HD void rc_dp_split(double v, uint32_t& hi, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
    hi = (uint32_t)__double2hiint(v); lo = (uint32_t)__double2loint(v);
#else
    unsigned long long b; memcpy(&b, &v, 8); hi = (uint32_t)(b >> 32); lo = (uint32_t)b;
#endif
}
CR_HD double rc_dp_join(uint32_t hi, uint32_t lo) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double((int)hi, (int)lo);
#else
    unsigned long long b = (unsigned long long)hi << 32 | lo; double v; memcpy(&v, &b, 8); return v;
#endif
}
// the renormalisation on the high word: (hi & 0x007FFFFF) | 0x41800000.  The compiler emits two LOP3 for two immediates; with the
// second constant in a register it is one (measured: 27.3 instead of 31.1 cycles per symbol, tests/micro/dpstep.cu)
CR_HD uint32_t rc_dp_renorm(uint32_t ch) {
#if defined(__CUDA_ARCH__)
    uint32_t c = 0x41800000u, r;
    asm volatile("" : "+r"(c));
    asm("lop3.b32 %0, %1, 0x007FFFFF, %2, 0xEA;" : "=r"(r) : "r"(ch), "r"(c));
    return r;
#else
    return (ch & 0x007FFFFFu) | 0x41800000u;
#endif
}
// chain record of one symbol: {inv / 2, 2 * frq} as two doubles in a uint4 slot (x, y = lo, hi of the first; z, w = lo, hi of the second)
CR_HD uint4 rc_dp_record(uint32_t frq, uint32_t sum) {
    double inv = 1.0 / (double)sum;
    uint32_t ih, il; rc_dp_split(inv, ih, il);
    unsigned long long b = ((unsigned long long)ih << 32 | il) + 2ull;          // two ulps up: strictly above 1/sum, power-of-two sums included
    b -= 1ull << 52;                                                            // halve (sum < 2^23: far from the subnormals)
    uint32_t fh, fl; rc_dp_split(2.0 * (double)frq, fh, fl);
    uint4 r; r.x = (uint32_t)b; r.y = (uint32_t)(b >> 32); r.z = fl; r.w = fh;
    return r;
}
// one symbol: R = 2 * (normalised range) as a double; returns q, msb = top-bit index of the UN-normalised new range (as VARIANT 4 stores it)
CR_HD void rc_dp_step(double& R, const uint4 t, uint32_t& q, uint32_t& msb) {
    const double inv = rc_dp_join(t.y, t.x), f = rc_dp_join(t.w, t.z);
    const double T = fma(R, inv, RC_DP_MAGIC);
    const double C = fma(T, f, -(RC_DP_TWO52 * f));
    uint32_t th, tl, ch, cl;
    rc_dp_split(T, th, tl); rc_dp_split(C, ch, cl);
    q = tl;
    msb = (ch >> 20) - 1024u;
    R = rc_dp_join(rc_dp_renorm(ch), cl);
}

// the same with -(2^52 * 2 frq) supplied by the caller (k_rcp_emit computes it while staging, off the chain)
CR_HD void rc_dp_step_nf(double& R, const uint4 t, const double nf, uint32_t& q, uint32_t& msb) {
    const double inv = rc_dp_join(t.y, t.x), f = rc_dp_join(t.w, t.z);
    const double T = fma(R, inv, RC_DP_MAGIC);
    const double C = fma(T, f, nf);
    uint32_t th, tl, ch, cl;
    rc_dp_split(T, th, tl); rc_dp_split(C, ch, cl);
    q = tl;
    msb = (ch >> 20) - 1024u;
    R = rc_dp_join(rc_dp_renorm(ch), cl);
}

struct RcStream {
    uint32_t ev_begin, ev_end;   // main: event range (dense range = + escord); side: dense range directly
    uint32_t is_main;
    uint32_t limit;              // abort once this many bytes are out at a token end (main only; 0xFFFFFFFF = never)
    uint64_t out_off;            // into the rc output buffer
    uint32_t out_cap;
    uint32_t pad;
};
struct RcResult { uint32_t nbytes; uint32_t aborted; uint32_t abort_tri; uint32_t pad; };   // abort_tri: dense index of the token-end triple the abort was taken at

struct RcCoder {
    uint32_t low, range, follow, carry, cache, n, cap;
    uint8_t* out;
    CR_D void put(uint32_t b) { if (n < cap) out[n] = (uint8_t)b; n++; }
    CR_D void shift_out() {                                   // renormalize(), cr-rangecoder.c:44-58
        if (low < 0xFF000000u || carry) {
            put(cache + carry);
            for (; follow; follow--) put(carry - 1);
            cache = low >> 24;
            carry = 0;
        } else follow++;
        low <<= 8;
    }
};

__global__ void k_range_encode(const Tri* __restrict__ dense_main, const Tri* __restrict__ dense_side, const uint32_t* __restrict__ escord,
                               const RcStream* __restrict__ streams, uint32_t nstreams, uint8_t* __restrict__ outbuf, RcResult* __restrict__ res) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nstreams) return;
    const RcStream S = streams[s];
    const Tri* tri = S.is_main ? dense_main : dense_side;
    size_t i0 = S.ev_begin, i1 = S.ev_end;
    if (S.is_main) { i0 += escord[S.ev_begin]; i1 += escord[S.ev_end]; }
    RcCoder c;
    c.low = 0; c.range = 0xFFFFFFFFu; c.follow = 0; c.carry = 0; c.cache = 0; c.n = 0; c.cap = S.out_cap; c.out = outbuf + S.out_off;
    uint32_t aborted = 0, abort_tri = 0;
    for (size_t i = i0; i < i1; i++) {
        const Tri t = tri[i];
        uint32_t q = __umulhi(c.range, t.magic);
        uint32_t rem = c.range - q * t.sum;
        if (rem >= t.sum) q++;
        uint32_t nl = c.low + t.cum * q;
        c.carry += nl < c.low;
        c.low = nl;
        c.range = q * (t.frq & 0x7FFFFFFFu);
        while (c.range < (1u << 24)) { c.range <<= 8; c.shift_out(); }
        if ((t.frq & TRI_TOKEND) && c.n >= S.limit) { aborted = 1; abort_tri = (uint32_t)i; break; }     // cr-coder.c:231-233
    }
    if (!aborted) for (int k = 0; k < 5; k++) c.shift_out();                    // range_encoder_flush
    res[s].nbytes = c.n;
    res[s].aborted = aborted;
    res[s].abort_tri = abort_tri; res[s].pad = 0;
}

// ------------------------------------------------------------------ payload assembly
struct CopyDesc { uint64_t src, dst; uint32_t len; uint32_t src_buf; };   // src_buf: 0 = rc output, 1 = dictionary-coded data
__global__ void k_copy_segments(const CopyDesc* __restrict__ descs, const uint8_t* __restrict__ buf0, const uint8_t* __restrict__ buf1, uint8_t* __restrict__ dst) {
    const CopyDesc d = descs[blockIdx.y];
    const uint8_t* src = (d.src_buf ? buf1 : buf0) + d.src;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < d.len; i += gridDim.x * blockDim.x) dst[d.dst + i] = src[i];
}
struct HeaderDesc { uint64_t dst; uint32_t len; uint8_t bytes[44]; };   // 6-byte block prefix + the largest inner header (32, LZ77)
__global__ void k_write_headers(const HeaderDesc* __restrict__ h, uint32_t n, uint8_t* __restrict__ dst) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (uint32_t k = 0; k < h[i].len; k++) dst[h[i].dst + k] = h[i].bytes[k];
}
