// cr_warp.cuh -- warp-cooperative versions of the serial model / coder loops (GPU only).
//
// The scalar kernels in cr_ppm.cuh / cr_rc.cuh define the semantics (and are what the CPU kernel-logic
// simulation runs); these kernels compute exactly the same thing with one WARP per serial chain:
//   * frequency tables live in registers, 8 symbols per lane, so a cumulative frequency is one masked
//     byte-sum per lane (dp4a) + one REDUX, instead of a walk over local memory;
//   * the chain's inputs are staged 32 at a time with coalesced loads issued one batch ahead, so the
//     dependent chain never waits on DRAM/L2 latency;
//   * results are written back coalesced, 32 at a time.
// tests/test_gpu_*.py run both families against the oracle; crgpu_set_option(h, "scalar_models", 1) selects
// the scalar family on the GPU for A/B checks.
#pragma once
#ifndef CRGPU_SIM
#include "cr_common.cuh"
#include "cr_ppm.cuh"
#include "cr_rc.cuh"

#define FULLMASK 0xFFFFFFFFu

CR_D uint32_t wsum4(uint32_t v) { return __dp4a(v, 0x01010101u, 0u); }                 // sum of the 4 bytes of v
CR_D uint32_t byte_mask_below(uint32_t k) { return k >= 4 ? 0xFFFFFFFFu : ((1u << (8 * k)) - 1u); } // bytes [0,k) = 0xFF
CR_D uint32_t ones_below(uint32_t k) { return byte_mask_below(k) & 0x01010101u; }

// first index i in [0,n) with key(i) >= c, for a sorted key array (warp-uniform binary search)
template <class K, class F> CR_D uint32_t lower_bound_key(const K* __restrict__ keys, uint32_t n, uint32_t c, F keyof) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (keyof(keys[mid]) < c) lo = mid + 1; else hi = mid; }
    return lo;
}

// ------------------------------------------------------------------ o2 pass, one warp per ctx16
// lane l holds the frequencies of symbols 8l..8l+7 in (f0, f1); flags 256/257 and the body total are uniform.
__global__ void __launch_bounds__(128) k_o2_pass_warp(const uint32_t* __restrict__ K, const uint32_t* __restrict__ V, uint32_t n, PpmState st,
                                                       uint64_t* __restrict__ T1, EscRec* __restrict__ esc_rec, uint32_t* __restrict__ esc_count) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t c16 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c16 >= 65536) return;
    auto keyof = [](uint32_t k) { return k & 0xffffu; };
    const uint32_t r0 = lower_bound_key(K, n, c16, keyof);
    if (r0 >= n || (K[r0] & 0xffff) != c16) return;
    const uint32_t r1 = lower_bound_key(K, n, c16 + 1, keyof);

    uint8_t* row = st.o2 + (size_t)c16 * PPM_O2_STRIDE;
    uint2 fv = ((const uint2*)row)[lane];
    uint32_t f0 = fv.x, f1 = fv.y;
    uint32_t f256 = row[256], f257 = row[257];
    uint32_t body = __reduce_add_sync(FULLMASK, wsum4(f0) + wsum4(f1));

    uint32_t kn = 0, vn = 0;
    if (r0 + lane < r1) { kn = K[r0 + lane]; vn = V[r0 + lane]; }
    for (uint32_t base = r0; base < r1; base += 32) {
        const uint32_t kc = kn, vc = vn;
        if (base + 32 + lane < r1) { kn = K[base + 32 + lane]; vn = V[base + 32 + lane]; }     // next batch in flight
        const uint32_t cnt = r1 - base < 32 ? r1 - base : 32;
        uint64_t mine = 0;
        for (uint32_t j = 0; j < cnt; j++) {
            const uint32_t k = __shfl_sync(FULLMASK, kc, j);
            const uint32_t sym = k >> 24, pr = (k >> 16) & 255;
            // frequency of the predicted byte and of the symbol (owner lanes: pr>>3, sym>>3)
            const uint32_t my_pr = ((pr & 4 ? f1 : f0) >> (8 * (pr & 3))) & 255;
            const uint32_t my_sy = ((sym & 4 ? f1 : f0) >> (8 * (sym & 3))) & 255;
            const uint32_t pf = __shfl_sync(FULLMASK, my_pr, pr >> 3);
            const uint32_t fs = __shfl_sync(FULLMASK, my_sy, sym >> 3);
            const uint32_t sum = body + f256 + f257 - pf;
            uint64_t t;
            uint32_t bump = 0xFFFFFFFFu;           // symbol (0..255) whose count goes up by one, if any
            bool rescale = false;
            if (sym == pr) {                                                     // cr-ppm.c:118-125
                t = ppm_pack(body - pf, f256, sum, 0);
                f256 = (f256 + 1) & 255;
                rescale = f256 > 250;
            } else if (fs > 0) {                                                 // cr-ppm.c:128-138
                const uint32_t own = sym >> 3, within = sym & 7;
                uint32_t part = 0;
                if (lane < own) part = wsum4(f0) + wsum4(f1);
                else if (lane == own) part = __dp4a(f0, ones_below(within), 0u) + (within > 4 ? __dp4a(f1, ones_below(within - 4), 0u) : 0u);
                const uint32_t cum = __reduce_add_sync(FULLMASK, part);
                t = ppm_pack(cum - (sym >= pr ? pf : 0), fs, sum, 0);
                bump = sym;
                if (fs + 1 > 250) rescale = true;
                else if (fs + 1 == 2) {                                          // escape estimator: 257 goes down
                    f257 = (f257 - 1) & 255;
                    // a wrap to 255 rescales too (uint8 arithmetic of the reference, cr-o2model.c:49,54)
                    if (f257 > 250) {
                        // apply the pending +1 first, then rescale below with the "257" rule
                        if (lane == own) { if (within < 4) f0 += 1u << (8 * within); else f1 += 1u << (8 * (within - 4)); }
                        body += 1; bump = 0xFFFFFFFFu; rescale = true;
                    }
                }
            } else {                                                             // cr-ppm.c:140-162
                t = ppm_pack(body + f256 - pf, f257, sum, 1);
                f257 = (f257 + 1) & 255;
                const bool resc257 = f257 > 250;
                if (resc257) {          // rescale happens BEFORE the exclusion mask is taken (cr-ppm.c:146-151)
                    f0 = (f0 >> 1) & 0x7f7f7f7fu; f1 = (f1 >> 1) & 0x7f7f7f7fu;
                    const uint32_t ones = (__popc(__vcmpeq4(f0, 0x01010101u)) + __popc(__vcmpeq4(f1, 0x01010101u))) >> 3;
                    const uint32_t ee = 1 + __reduce_add_sync(FULLMASK, ones);
                    body = __reduce_add_sync(FULLMASK, wsum4(f0) + wsum4(f1));
                    f256 = (f256 + 1) >> 1; f257 = ee & 255;
                }
                // exclusion mask: bit set = o2 frequency is zero and the symbol is not the predicted byte
                uint32_t z0 = __vcmpeq4(f0, 0u), z1 = __vcmpeq4(f1, 0u);
                uint32_t bits = ((z0 & 1u) | (z0 >> 7 & 2u) | (z0 >> 14 & 4u) | (z0 >> 21 & 8u)) | (((z1 & 1u) | (z1 >> 7 & 2u) | (z1 >> 14 & 4u) | (z1 >> 21 & 8u)) << 4);
                if (lane == (pr >> 3)) bits &= ~(1u << (pr & 7));
                uint32_t slot = 0;
                if (lane == 0) slot = atomicAdd(esc_count, 1u);
                slot = __shfl_sync(FULLMASK, slot, 0);
                EscRec* rec = esc_rec + slot;
                ((uint8_t*)rec->incl)[lane] = (uint8_t)bits;
                const uint32_t ev = __shfl_sync(FULLMASK, vc, j);
                if (lane == 0) { rec->e = ev; rec->info = (c16 & 0xff) | sym << 8; }
                if (!resc257) bump = sym;                                        // new symbol enters with count 1
            }
            if (bump != 0xFFFFFFFFu) {
                if (lane == (bump >> 3)) { const uint32_t w = bump & 7; if (w < 4) f0 += 1u << (8 * w); else f1 += 1u << (8 * (w - 4)); }
                body += 1;
            }
            if (rescale) {                                                       // cr-o2model.c:54-69
                f0 = (f0 >> 1) & 0x7f7f7f7fu; f1 = (f1 >> 1) & 0x7f7f7f7fu;
                const uint32_t ones = (__popc(__vcmpeq4(f0, 0x01010101u)) + __popc(__vcmpeq4(f1, 0x01010101u))) >> 3;
                const uint32_t ee = 1 + __reduce_add_sync(FULLMASK, ones);
                body = __reduce_add_sync(FULLMASK, wsum4(f0) + wsum4(f1));
                f256 = (f256 + 1) >> 1; f257 = ee & 255;
            }
            if (lane == j) mine = t;
        }
        if (lane < cnt) T1[vc] = mine;
    }
    ((uint2*)row)[lane] = make_uint2(f0, f1);
    if (lane == 0) { row[256] = (uint8_t)f256; row[257] = (uint8_t)f257; }
}

// ------------------------------------------------------------------ o1 pass, one warp per ctx8
// lane l holds o1 counts of symbols 8l..8l+7 in (a0, a1).  Escapes arrive sorted by (ctx8, time).
__global__ void __launch_bounds__(128) k_o1_pass_warp(const uint64_t* __restrict__ K, const uint32_t* __restrict__ V, uint32_t n, const EscRec* __restrict__ rec,
                                                       const uint32_t* __restrict__ ord, PpmState st, uint64_t* __restrict__ T2) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t c8 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c8 >= 256) return;
    auto keyof = [](uint64_t k) { return (uint32_t)(k >> 32) & 0xffu; };
    const uint32_t r0 = lower_bound_key(K, n, c8, keyof);
    if (r0 >= n || keyof(K[r0]) != c8) return;
    const uint32_t r1 = lower_bound_key(K, n, c8 + 1, keyof);
    uint8_t* row = st.o1 + c8 * 256;
    uint2 av = ((const uint2*)row)[lane];
    uint32_t a0 = av.x, a1 = av.y;

    // software pipeline: the record fields of step i+1 are loaded while step i is computed
    uint32_t vi = V[r0];
    uint32_t n_info = rec[vi].info, n_bits = ((const uint8_t*)rec[vi].incl)[lane], n_ord = ord[vi];
    for (uint32_t i = r0; i < r1; i++) {
        const uint32_t info = n_info, bits = n_bits, o = n_ord;
        if (i + 1 < r1) { vi = V[i + 1]; n_info = rec[vi].info; n_bits = ((const uint8_t*)rec[vi].incl)[lane]; n_ord = ord[vi]; }
        const uint32_t sym = (info >> 8) & 255, own = sym >> 3, within = sym & 7;
        // frequencies o1*8-7 of the included symbols of this lane
        uint32_t sum = 0, cum = 0;
#pragma unroll
        for (uint32_t k = 0; k < 8; k++) {
            const uint32_t c = ((k < 4 ? a0 : a1) >> (8 * (k & 3))) & 255;
            const uint32_t fr = (bits >> k & 1u) ? c * 8 - 7 : 0u;
            sum += fr;
            if (lane < own || (lane == own && k < within)) cum += fr;
        }
        sum = __reduce_add_sync(FULLMASK, sum);
        cum = __reduce_add_sync(FULLMASK, cum);
        const uint32_t mine = ((within < 4 ? a0 : a1) >> (8 * (within & 3))) & 255;
        const uint32_t cs = __shfl_sync(FULLMASK, mine, own);
        if (lane == 0) T2[o] = ppm_pack(cum, cs * 8 - 7, sum, 0);
        // ppm_update_o1 (cr-ppm.c:90-97)
        if (lane == own) { if (within < 4) a0 += 1u << (8 * within); else a1 += 1u << (8 * (within - 4)); }
        if (cs + 1 >= 255) { a0 -= (a0 >> 1) & 0x7f7f7f7fu; a1 -= (a1 >> 1) & 0x7f7f7f7fu; }
    }
    ((uint2*)row)[lane] = make_uint2(a0, a1);
}

// ------------------------------------------------------------------ order-0 side models, one warp
// lane l holds 8 x u16 counts of symbols 8l..8l+7 for both models (len, idx).
CR_D uint32_t side_part(const uint32_t (&f)[4], uint32_t upto) {      // sum of the first `upto` (0..8) counts of this lane
    uint32_t s = 0;
#pragma unroll
    for (uint32_t k = 0; k < 8; k++) { uint32_t c = (f[k >> 1] >> (16 * (k & 1))) & 0xffff; if (k < upto) s += c; }
    return s;
}
__global__ void __launch_bounds__(32) k_side_models_warp(const uint16_t* __restrict__ side_sym, uint32_t n, PpmState st, uint64_t* __restrict__ TS) {
    if (blockIdx.x != 0) return;
    const uint32_t lane = threadIdx.x & 31;
    uint32_t fa[4], fb[4];          // len_model, idx_model
    {
        const uint4 v = ((const uint4*)(st.m0))[lane];       fa[0] = v.x; fa[1] = v.y; fa[2] = v.z; fa[3] = v.w;
        const uint4 u = ((const uint4*)(st.m0 + 256))[lane]; fb[0] = u.x; fb[1] = u.y; fb[2] = u.z; fb[3] = u.w;
    }
    uint32_t tota = __reduce_add_sync(FULLMASK, side_part(fa, 8));
    uint32_t totb = __reduce_add_sync(FULLMASK, side_part(fb, 8));
    uint32_t sn = lane < n ? side_sym[lane] : 0;
    for (uint32_t base = 0; base < n; base += 32) {
        const uint32_t sc = sn;
        if (base + 32 + lane < n) sn = side_sym[base + 32 + lane];
        const uint32_t cnt = n - base < 32 ? n - base : 32;
        uint64_t mine = 0;
        for (uint32_t j = 0; j < cnt; j++) {
            const uint32_t v = __shfl_sync(FULLMASK, sc, j);
            const uint32_t m = v >> 8, s = v & 255, own = s >> 3, within = s & 7;
            const uint32_t upto = lane < own ? 8u : (lane == own ? within : 0u);
            const uint32_t sh = 16 * (within & 1), wi = within >> 1;
            uint32_t part, me, tot;
            if (m == 0) { part = side_part(fa, upto); me = (fa[wi] >> sh) & 0xffff; tot = tota; }
            else        { part = side_part(fb, upto); me = (fb[wi] >> sh) & 0xffff; tot = totb; }
            const uint32_t cum = __reduce_add_sync(FULLMASK, part);
            const uint32_t fr = __shfl_sync(FULLMASK, me, own);
            const uint64_t t = ppm_pack(cum, fr, tot, 0);
            if (lane == j) mine = t;
            // model_update(+4), halve (rounding up) when the total passes 32000 (cr-model.c:55-77)
            if (m == 0) {
                if (lane == own) fa[wi] += 4u << sh;
                tota += 4;
                if (tota > 32000) { for (int k = 0; k < 4; k++) fa[k] = ((fa[k] + 0x00010001u) >> 1) & 0x7fff7fffu; tota = __reduce_add_sync(FULLMASK, side_part(fa, 8)); }
            } else {
                if (lane == own) fb[wi] += 4u << sh;
                totb += 4;
                if (totb > 32000) { for (int k = 0; k < 4; k++) fb[k] = ((fb[k] + 0x00010001u) >> 1) & 0x7fff7fffu; totb = __reduce_add_sync(FULLMASK, side_part(fb, 8)); }
            }
        }
        if (lane < cnt) TS[base + lane] = mine;
    }
    ((uint4*)(st.m0))[lane] = make_uint4(fa[0], fa[1], fa[2], fa[3]);
    ((uint4*)(st.m0 + 256))[lane] = make_uint4(fb[0], fb[1], fb[2], fb[3]);
}

// ------------------------------------------------------------------ range coder, one warp per stream
// All lanes stage triples through shared memory one batch ahead; lane 0 runs the serial recurrence.
#define RC_BATCH 64
__global__ void __launch_bounds__(128) k_range_encode_warp(const Tri* __restrict__ dense_main, const Tri* __restrict__ dense_side, const uint32_t* __restrict__ escord,
                                                            const RcStream* __restrict__ streams, uint32_t nstreams, uint8_t* __restrict__ outbuf, RcResult* __restrict__ res) {
    __shared__ uint4 stage[4][2][RC_BATCH];
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= nstreams) return;
    const RcStream S = streams[s];
    const uint4* tri = (const uint4*)(S.is_main ? dense_main : dense_side);
    size_t i0 = S.ev_begin, i1 = S.ev_end;
    if (S.is_main) { i0 += escord[S.ev_begin]; i1 += escord[S.ev_end]; }
    RcCoder c;
    c.low = 0; c.range = 0xFFFFFFFFu; c.follow = 0; c.carry = 0; c.cache = 0; c.n = 0; c.cap = S.out_cap; c.out = outbuf + S.out_off;
    uint32_t aborted = 0;
    uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
    if (i0 + lane < i1) r0 = tri[i0 + lane];
    if (i0 + 32 + lane < i1) r1 = tri[i0 + 32 + lane];
    uint32_t buf = 0;
    for (size_t base = i0; base < i1; base += RC_BATCH, buf ^= 1) {
        stage[w][buf][lane] = r0; stage[w][buf][lane + 32] = r1;
        __syncwarp();
        const size_t nb = base + RC_BATCH;
        if (nb + lane < i1) r0 = tri[nb + lane];                               // next batch in flight
        if (nb + 32 + lane < i1) r1 = tri[nb + 32 + lane];
        const uint32_t cnt = i1 - base < RC_BATCH ? (uint32_t)(i1 - base) : RC_BATCH;
        if (lane == 0 && !aborted) {
            uint4 t = stage[w][buf][0];
            for (uint32_t j = 0; j < cnt; j++) {
                const uint4 tn = stage[w][buf][j + 1 < RC_BATCH ? j + 1 : j];   // prefetch next from shared
                // Tri layout: x = cum, y = frq|flag, z = sum, w = magic
                uint32_t q = __umulhi(c.range, t.w);
                const uint32_t rem = c.range - q * t.z;
                if (rem >= t.z) q++;
                const uint32_t nl = c.low + t.x * q;
                c.carry += nl < c.low;
                c.low = nl;
                c.range = q * (t.y & 0x7FFFFFFFu);
                while (c.range < (1u << 24)) { c.range <<= 8; c.shift_out(); }
                if ((t.y & TRI_TOKEND) && c.n >= S.limit) { aborted = 1; break; }
                t = tn;
            }
        }
        aborted = __shfl_sync(FULLMASK, aborted, 0);
        if (aborted) break;
    }
    if (lane == 0) {
        if (!aborted) for (int k = 0; k < 5; k++) c.shift_out();
        res[s].nbytes = c.n;
        res[s].aborted = aborted;
    }
}
#endif  // !CRGPU_SIM
