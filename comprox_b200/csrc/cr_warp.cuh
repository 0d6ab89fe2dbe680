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
#include "cr_decode.cuh"
#include "cr_lz77.cuh"

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

// ------------------------------------------------------------------ o3 pass, balanced over slot segments
// k_o3_pass (cr_ppm.cuh) gives one thread to every sorted rank and lets segment heads walk their segment: a warp
// then runs as long as its longest segment.  Here the segments are listed, ordered by length (longest first) and
// given one thread each, so the lanes of a warp have equally long walks; segments of O3_HANDOVER events or more
// go to k_o3_hot.
__global__ void k_o3_headflags(const uint32_t* __restrict__ K, uint32_t n, uint32_t* __restrict__ flag) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n) return;
    flag[r] = (r < n && (r == 0 || (K[r - 1] & 0x3fffff) != (K[r] & 0x3fffff))) ? 1u : 0u;
}
__global__ void k_o3_segstarts(const uint32_t* __restrict__ flag, const uint32_t* __restrict__ hidx, uint32_t n, uint32_t* __restrict__ seg_start) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n) return;
    if (r == n) seg_start[hidx[n]] = n;                 // sentinel after the last segment
    else if (flag[r]) seg_start[hidx[r]] = r;
}
__global__ void k_o3_seglen_keys(const uint32_t* __restrict__ seg_start, uint32_t nseg, uint32_t* __restrict__ key, uint32_t* __restrict__ val) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nseg) return;
    key[i] = 0xFFFFFFFFu - (seg_start[i + 1] - seg_start[i]);
    val[i] = i;
}
__global__ void k_o3_pass_sorted(const uint32_t* __restrict__ K, const uint32_t* __restrict__ V, uint32_t n, const uint32_t* __restrict__ seg_start,
                                 const uint32_t* __restrict__ order, uint32_t nseg, PpmState st, uint8_t* __restrict__ pred,
                                 O3Hot* __restrict__ hot, uint32_t* __restrict__ hot_count, uint32_t hot_cap) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nseg) return;
    const uint32_t seg = order[t];
    const uint32_t r0 = seg_start[seg], r1 = seg_start[seg + 1];
    const uint32_t slot = K[r0] & 0x3fffff;
    uint32_t byte = st.o3_byte[slot], conf = st.o3_conf[slot];
    if (r1 - r0 >= O3_HANDOVER) {
        uint32_t h = atomicAdd(hot_count, 1u);
        if (h < hot_cap) { hot[h].slot = slot; hot[h].rank = r0; hot[h].byte = byte; hot[h].conf = conf; return; }
    }
    for (uint32_t i = r0; i < r1; i += 8) {
        uint32_t kk[8], vv[8];
#pragma unroll
        for (int u = 0; u < 8; u++) { const uint32_t x = i + u < r1 ? i + u : r1 - 1; kk[u] = K[x]; vv[u] = V[x]; }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (i + u >= r1) break;
            const uint32_t sym = kk[u] >> 24;
            pred[vv[u]] = (uint8_t)byte;
            if (sym == byte) conf += conf < 15;
            else { conf = (conf > 1) + (conf > 2) + (conf > 4) + (conf > 8); if (conf == 0) { byte = sym; conf = 1; } }
        }
    }
    st.o3_byte[slot] = (uint8_t)byte; st.o3_conf[slot] = (uint8_t)conf;
}

// ------------------------------------------------------------------ o3 pass, long slot segments
// k_o3_pass (one thread per slot) hands segments longer than O3_HANDOVER events to this kernel: one warp per
// segment, keys staged 64 at a time through shared memory one batch ahead, lane 0 runs the 12-bit state machine
// (ppm_update_o3, cr-ppm.c:69-88), all lanes scatter the predicted bytes.
__global__ void __launch_bounds__(128) k_o3_hot(const uint32_t* __restrict__ K, const uint32_t* __restrict__ V, uint32_t n, PpmState st, uint8_t* __restrict__ pred,
                                                 const O3Hot* __restrict__ hot, const uint32_t* __restrict__ hot_count, uint32_t hot_cap) {
    __shared__ uint32_t sk[4][64];
    __shared__ uint8_t sp[4][64];
    __shared__ uint32_t s_cnt[4];
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    uint32_t total = *hot_count; if (total > hot_cap) total = hot_cap;
    for (uint32_t e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < total; e += nwarps) {
        const O3Hot H = hot[e];
        uint32_t byte = H.byte, conf = H.conf;
        uint32_t k0 = 0xFFFFFFFFu, k1 = 0xFFFFFFFFu, v0 = 0, v1 = 0;
        if (H.rank + lane < n) { k0 = K[H.rank + lane]; v0 = V[H.rank + lane]; }
        if (H.rank + 32 + lane < n) { k1 = K[H.rank + 32 + lane]; v1 = V[H.rank + 32 + lane]; }
        for (uint32_t base = H.rank;; base += 64) {
            const uint32_t c0 = v0, c1 = v1;
            sk[w][lane] = k0; sk[w][lane + 32] = k1;
            __syncwarp();
            const uint32_t nb = base + 64;
            k0 = k1 = 0xFFFFFFFFu;
            if (nb + lane < n) { k0 = K[nb + lane]; v0 = V[nb + lane]; }
            if (nb + 32 + lane < n) { k1 = K[nb + 32 + lane]; v1 = V[nb + 32 + lane]; }
            if (lane == 0) {
                uint32_t j = 0;
                for (; j < 64; j++) {
                    const uint32_t k = sk[w][j];
                    if (base + j >= n || (k & 0x3fffff) != H.slot) break;
                    const uint32_t sym = k >> 24;
                    sp[w][j] = (uint8_t)byte;
                    if (sym == byte) conf += conf < 15;
                    else { conf = (conf > 1) + (conf > 2) + (conf > 4) + (conf > 8); if (conf == 0) { byte = sym; conf = 1; } }
                }
                s_cnt[w] = j;
            }
            __syncwarp();
            const uint32_t cnt = s_cnt[w];
            if (lane < cnt) pred[c0] = sp[w][lane];
            if (lane + 32 < cnt) pred[c1] = sp[w][lane + 32];
            __syncwarp();
            if (cnt < 64) break;
        }
        if (lane == 0) { st.o3_byte[H.slot] = (uint8_t)byte; st.o3_conf[H.slot] = (uint8_t)conf; }
    }
}

// ------------------------------------------------------------------ o3 pass, long slot segments, speculative
// The o3 slot automaton (12 bits of state) forgets its past quickly: four misses in a row, or a run of hits that
// saturates the confidence, bring any two states together.  So a long segment is cut into chunks of O3S_CHUNK
// events; every chunk is replayed by its own thread from a GUESSED entry state (obtained by warming up over the
// previous chunk), all chunks at once.  Thread 0 then walks the chunks in order: where the guess equals the true
// entry state the speculative results stand, otherwise that chunk alone is replayed from the true state.
// Exactness does not depend on the guess; only speed does.
#define O3S_CHUNK   256u
#define O3S_THREADS 256u
CR_D void o3_step(uint32_t& byte, uint32_t& conf, uint32_t sym) {          // ppm_update_o3, cr-ppm.c:69-88
    if (sym == byte) conf += conf < 15;
    else { conf = (conf > 1) + (conf > 2) + (conf > 4) + (conf > 8); if (conf == 0) { byte = sym; conf = 1; } }
}
__global__ void __launch_bounds__(O3S_THREADS) k_o3_hot_spec(const uint32_t* __restrict__ K, const uint32_t* __restrict__ V, uint32_t n, const uint32_t* __restrict__ seg_start,
                                                             PpmState st, uint8_t* __restrict__ pred, const O3Hot* __restrict__ hot, const uint32_t* __restrict__ hot_count, uint32_t hot_cap) {
    __shared__ uint32_t s_guess[O3S_THREADS], s_end[O3S_THREADS], s_entry[O3S_THREADS];          // byte | conf << 8
    __shared__ uint32_t s_true, s_chain, s_fail;
    uint32_t total = *hot_count; if (total > hot_cap) total = hot_cap;
    // the list comes longest segment first (k_o3_pass_sorted hands segments over in that order): CTAs take the next entry when they
    // are done with theirs (hot_count[1] is the queue), which is longest-processing-time-first scheduling; a fixed round robin gave
    // the first CTAs the longest segment of every round (ncu: sm__cycles_active avg 1.26 M, max 4.93 M)
    __shared__ uint32_t s_e;
    for (;;) {
        if (threadIdx.x == 0) s_e = atomicAdd(const_cast<uint32_t*>(hot_count) + 1, 1u);
        __syncthreads();
        const uint32_t e = s_e;
        if (e >= total) break;
        const O3Hot H = hot[e];
        const uint32_t r0 = H.rank;
        uint32_t r1 = r0;                                                  // end of the slot's segment
        {   // the segment ends where the slot changes; segments are sorted, so gallop + binary search
            uint32_t lo = r0, step = O3S_CHUNK;
            while (lo + step < n && (K[lo + step] & 0x3fffff) == H.slot) { lo += step; step <<= 1; }
            uint32_t hi = lo + step < n ? lo + step : n;                   // K[hi] differs (or hi == n)
            while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if ((K[mid] & 0x3fffff) == H.slot) lo = mid; else hi = mid; }
            r1 = hi;
        }
        if (threadIdx.x == 0) s_true = H.byte | H.conf << 8;
        __syncthreads();
        for (uint32_t pass0 = r0; pass0 < r1; pass0 += O3S_CHUNK * O3S_THREADS) {
            const uint32_t c0 = pass0 + threadIdx.x * O3S_CHUNK;
            const bool have = c0 < r1;
            const uint32_t c1 = have ? (c0 + O3S_CHUNK < r1 ? c0 + O3S_CHUNK : r1) : c0;
            uint32_t byte = 0, conf = 0;
            if (have) {
                if (threadIdx.x == 0) { byte = s_true & 255; conf = s_true >> 8; }           // exact entry of the pass
                else for (uint32_t i = c0 - O3S_CHUNK; i < c0; i++) o3_step(byte, conf, K[i] >> 24);   // warm-up guess
                s_guess[threadIdx.x] = byte | conf << 8;
                for (uint32_t i = c0; i < c1; i++) { pred[V[i]] = (uint8_t)byte; o3_step(byte, conf, K[i] >> 24); }
                s_end[threadIdx.x] = byte | conf << 8;
            }
            __syncthreads();
            // Verification.  A wrong guess used to be replayed by thread 0 on the spot -- 256 events one after the other while 255 threads
            // waited (ncu: 38 barrier-stall cycles per issue, the hottest slot's CTA 4.9 M cycles of a 2.5 ms kernel).  The automaton
            // forgets, so a chunk entered in the wrong state still ENDS in the state its speculative run ended in, nearly always: thread 0
            // walks the chain assuming just that, notes the true entry of every chunk whose guess was wrong, all those chunks are replayed
            // side by side, and each replay checks the assumption (its end against the end the chain used).  The first chunk whose end
            // differs restarts the walk behind it.  Every accepted chunk's last run started from its true entry state.
            const uint32_t nch = (r1 - pass0 + O3S_CHUNK - 1) / O3S_CHUNK < O3S_THREADS ? (r1 - pass0 + O3S_CHUNK - 1) / O3S_CHUNK : O3S_THREADS;
            uint32_t from = 0;
            for (;;) {
                if (threadIdx.x == 0) {
                    uint32_t cur = s_true;
                    for (uint32_t c = from; c < nch; c++) { s_entry[c] = s_guess[c] == cur ? 0xFFFFFFFFu : cur; cur = s_end[c]; }
                    s_chain = cur; s_fail = 0xFFFFFFFFu;
                }
                __syncthreads();
                const uint32_t c = threadIdx.x;
                if (c >= from && c < nch && s_entry[c] != 0xFFFFFFFFu) {
                    uint32_t b = s_entry[c] & 255, cf = s_entry[c] >> 8;
                    const uint32_t a0 = pass0 + c * O3S_CHUNK, a1 = a0 + O3S_CHUNK < r1 ? a0 + O3S_CHUNK : r1;
                    for (uint32_t i = a0; i < a1; i++) { pred[V[i]] = (uint8_t)b; o3_step(b, cf, K[i] >> 24); }
                    const uint32_t e = b | cf << 8;
                    if (e != s_end[c]) atomicMin(&s_fail, c);
                    s_guess[c] = s_entry[c]; s_end[c] = e;                 // the chunk's latest run: entry -> end
                }
                __syncthreads();
                const uint32_t f = s_fail;
                if (f == 0xFFFFFFFFu) break;
                // chunks up to f are settled (every replayed end before f matched); the chain behind f starts from f's true end
                if (threadIdx.x == 0) s_true = s_end[f];
                from = f + 1;
                __syncthreads();
                if (from >= nch) { if (threadIdx.x == 0) s_chain = s_true; __syncthreads(); break; }
            }
            if (threadIdx.x == 0) s_true = s_chain;
            __syncthreads();
        }
        if (threadIdx.x == 0) { st.o3_byte[H.slot] = (uint8_t)(s_true & 255); st.o3_conf[H.slot] = (uint8_t)(s_true >> 8); }
        __syncthreads();
    }
}

// bounds[c] = first sorted rank whose 16-bit context is >= c (c = 0..65536): one binary search per context, done once
__global__ void k_o2_bounds(const uint32_t* __restrict__ K, uint32_t n, uint32_t* __restrict__ bounds) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > 65536) return;
    bounds[c] = lower_bound_key(K, n, c, [](uint32_t k) { return k & 0xffffu; });
}

// ------------------------------------------------------------------ o2 pass, one warp per ctx16
// lane l holds the frequencies of symbols 8l..8l+7 in (f0, f1); flags 256/257 and the body total are uniform.
__global__ void __launch_bounds__(128) k_o2_pass_warp(const uint32_t* __restrict__ K, const uint32_t* __restrict__ V, uint32_t n, PpmState st,
                                                       uint64_t* __restrict__ T1, EscRec* __restrict__ esc_rec, uint32_t* __restrict__ esc_count, uint32_t hot_min, const uint32_t* __restrict__ bounds) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t c16 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c16 >= 65536) return;
    const uint32_t r0 = bounds[c16], r1 = bounds[c16 + 1];
    if (r0 == r1 || r1 - r0 >= hot_min) return;           // empty, or taken by k_o2_pass_cta                       // taken by k_o2_pass_cta

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

// ------------------------------------------------------------------ o2 pass for HOT contexts, one CTA per ctx16
// The warp kernel above spends ~200 ns per event of a context; the hottest context of a text holds 3-4 % of all
// events and was the longest serial chain of the encoder.  Between two rescales the o2 table only counts:
//     f_x(i)   = f_x(0) + #{non-hit events j < i with symbol x}          (hits bump flag 256 instead, cr-ppm.c:124)
//     body(i)  = body(0) + #non-hits before i,   f256(i) = f256(0) + #hits before i,
//     f257(i)  = f257(0) + #escapes before i - #{counts that went 1 -> 2 before i}   (cr-ppm.c:136-138,146)
// so every quantity of ppm_encode is a start-of-step value plus a rank, and 1024 events are evaluated at once
// (per-warp histograms, prefix over warps / symbols, 32-way compare inside the warp -- as in k_side_epochs).
// The first event whose update would rescale the table (cr-o2model.c:54) ends the step; everything after it is
// recomputed in the next step from the rescaled table.
// Two instantiations: 256-thread CTAs when steps are cut early (flag 256 grows with every o3 hit, so a context with hit
// rate h rescales about every 188/h events), 1024-thread CTAs otherwise.  The host picks one per window from the overall
// hit rate (k_o2_hits).
// total number of o3 hits among the events (one atomic per CTA)
__global__ void k_o2_hits(const uint32_t* __restrict__ K, uint32_t n, uint32_t* __restrict__ hits) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t k = i < n ? K[i] : 0x01000000u;
    const int c = __syncthreads_count((k >> 24) == ((k >> 16) & 255));
    if (threadIdx.x == 0 && c) atomicAdd(hits, (uint32_t)c);
}
#define O2C_MIN     3072          // contexts with at least this many events in the window take this path
template <int O2C_THREADS>
__global__ void __launch_bounds__(O2C_THREADS) k_o2_pass_cta(const uint32_t* __restrict__ K, const uint32_t* __restrict__ V, uint32_t n, PpmState st,
                                                             uint64_t* __restrict__ T1, EscRec* __restrict__ esc_rec, uint32_t* __restrict__ esc_count, const uint32_t* __restrict__ bounds) {
    constexpr int O2C_WARPS = O2C_THREADS / 32;
    const uint32_t c16 = blockIdx.x;
    const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint32_t r0 = bounds[c16], r1 = bounds[c16 + 1];
    if (r1 - r0 < O2C_MIN) return;

    __shared__ uint32_t cnt[256], cumt[256], zmask[8];
    __shared__ uint32_t s_f256, s_f257, s_body, s_first;
    __shared__ __align__(4) uint16_t hist[O2C_WARPS][256];
    __shared__ uint16_t below[O2C_WARPS][256];
    __shared__ uint32_t wtot[3][32];                      // per-warp totals: non-hits, escapes, 1->2 transitions -> exclusive prefixes
    __shared__ uint16_t ssym[O2C_THREADS];                // symbol of each non-hit event (0x100 for hits) for the rare trigger-mask path
    __shared__ uint8_t esc_sym[O2C_THREADS];
    uint8_t* row = st.o2 + (size_t)c16 * PPM_O2_STRIDE;
    for (uint32_t i = tid; i < 256; i += O2C_THREADS) cnt[i] = row[i];
    if (tid == 0) { s_f256 = row[256]; s_f257 = row[257]; }
    __syncthreads();

    uint32_t pos = r0, cap = O2C_THREADS;                           // events attempted per step; grows while steps complete, shrinks when rescales cut them
    for (;;) {
        // ---- derived tables from cnt: cumt, body, zero mask (warp 0); clear the histograms (all)
        if (w == 0) {
            uint32_t v[8], sum = 0, zb = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) { v[k] = cnt[lane * 8 + k]; sum += v[k]; zb |= (uint32_t)(v[k] == 0) << k; }
            uint32_t inc = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(FULLMASK, inc, d); if (lane >= d) inc += t; }
            uint32_t run = inc - sum;
#pragma unroll
            for (int k = 0; k < 8; k++) { cumt[lane * 8 + k] = run; run += v[k]; }
            if (lane == 31) s_body = inc;
            // zmask word q = symbols 32q..32q+31 = lanes 4q..4q+3
            uint32_t z = zb << (8 * (lane & 3));
            z |= __shfl_xor_sync(FULLMASK, z, 1); z |= __shfl_xor_sync(FULLMASK, z, 2);
            if ((lane & 3) == 0) zmask[lane >> 2] = z;
        }
        const uint32_t step = r1 - pos < cap ? r1 - pos : cap;
        const uint32_t nw = (step + 31) >> 5;                 // warps that hold events in this step
        if (w < nw) for (uint32_t i = lane; i < 128; i += 32) ((uint32_t*)&hist[w][0])[i] = 0;
        if (tid < 96) (&wtot[0][0])[tid] = 0;
        if (tid == 0) s_first = 0xFFFFFFFFu;
        __syncthreads();
        if (pos >= r1) break;
        const bool active = tid < step;
        uint32_t k = 0, ev = 0;
        if (active) { k = K[pos + tid]; ev = V[pos + tid]; }
        const uint32_t sym = k >> 24, pr = (k >> 16) & 255;
        const bool hit = active && sym == pr, nonhit = active && sym != pr;
        ssym[tid] = nonhit ? (uint16_t)sym : (uint16_t)0x100;
        if (nonhit) atomicAdd((uint32_t*)&hist[w][0] + (sym >> 1), (sym & 1u) ? 0x10000u : 1u);
        __syncthreads();
        if (tid < 256) { uint32_t run = 0; for (uint32_t q = 0; q < nw; q++) { uint32_t h = hist[q][tid]; hist[q][tid] = (uint16_t)run; run += h; } }
        __syncthreads();
        if (w < nw) {
            uint32_t v[8], sum = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) { v[q] = hist[w][lane * 8 + q]; sum += v[q]; }
            uint32_t inc = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(FULLMASK, inc, d); if (lane >= d) inc += t; }
            uint32_t run = inc - sum;
#pragma unroll
            for (int q = 0; q < 8; q++) { below[w][lane * 8 + q] = (uint16_t)run; run += v[q]; }
        }
        __syncwarp();
        // ---- ranks inside the warp: earlier non-hit lanes with a smaller / the same symbol, or the predicted byte
        const uint32_t code = nonhit ? sym : 0x1FFu;
        uint32_t lt = 0, eq = 0, peq = 0;
#pragma unroll 8
        for (uint32_t j = 0; j < 32; j++) {
            const uint32_t sj = __shfl_sync(FULLMASK, code, j);
            if (j < lane) { lt += sj < sym; eq += sj == sym; peq += sj == pr; }
        }
        const uint32_t lanes_before = (1u << lane) - 1u;
        const uint32_t b_nh = __ballot_sync(FULLMASK, nonhit);
        uint32_t fs = 0, pf = 0;
        if (active) { fs = cnt[sym] + hist[w][sym] + eq; pf = cnt[pr] + hist[w][pr] + peq; }
        const bool esc = nonhit && fs == 0, two = nonhit && fs == 1;
        const uint32_t b_es = __ballot_sync(FULLMASK, esc), b_tw = __ballot_sync(FULLMASK, two);
        if (lane == 0) { wtot[0][w] = __popc(b_nh); wtot[1][w] = __popc(b_es); wtot[2][w] = __popc(b_tw); }
        __syncthreads();
        if (w < 3) {                                       // exclusive prefix of the three per-warp totals
            uint32_t v = wtot[w][lane], inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(FULLMASK, inc, d); if (lane >= d) inc += t; }
            wtot[w][lane] = inc - v;
        }
        __syncthreads();
        const uint32_t A = wtot[0][w] + __popc(b_nh & lanes_before);          // non-hits before this event
        const uint32_t ES = wtot[1][w] + __popc(b_es & lanes_before);         // escapes before
        const uint32_t TW = wtot[2][w] + __popc(b_tw & lanes_before);         // 1->2 transitions before
        const uint32_t f256 = s_f256 + (tid - A), f257 = s_f257 + ES - TW, body = s_body + A;
        // ---- does this event's update rescale the table?
        bool trig = false;
        if (hit) trig = f256 + 1 > 250;
        else if (esc) trig = f257 + 1 > 250;
        else if (nonhit) trig = (fs + 1 > 250) || (fs == 1 && f257 == 0);
        if (trig) atomicMin(&s_first, tid);
        if (esc) esc_sym[ES] = (uint8_t)sym;
        __syncthreads();
        const uint32_t first = s_first;
        const bool valid = active && tid <= first;
        if (valid) {
            const uint32_t sum = body + f256 + f257 - pf;
            if (hit) T1[ev] = ppm_pack(body - pf, f256, sum, 0);
            else if (!esc) T1[ev] = ppm_pack(cumt[sym] + below[w][sym] + lt - (sym >= pr ? pf : 0), fs, sum, 0);
            else {
                T1[ev] = ppm_pack(body + f256 - pf, f257, sum, 1);
                uint32_t m[8];
                if (trig) {
                    // the escape update rescales BEFORE the mask is taken (cr-ppm.c:146-151): zero <=> count <= 1 now
#pragma unroll
                    for (int q = 0; q < 8; q++) m[q] = 0;
                    for (uint32_t x = 0; x < 256; x++) {
                        uint32_t c = cnt[x] + hist[w][x];
                        for (uint32_t j = w * 32; j < tid && c < 2; j++) c += ssym[j] == x;
                        if (c < 2) m[x >> 5] |= 1u << (x & 31);
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 8; q++) m[q] = zmask[q];
                    for (uint32_t e = 0; e < ES; e++) { const uint32_t x = esc_sym[e]; m[x >> 5] &= ~(1u << (x & 31)); }
                }
                m[pr >> 5] &= ~(1u << (pr & 31));
                EscRec* rec = esc_rec + atomicAdd(esc_count, 1u);
                rec->e = ev; rec->info = (c16 & 0xff) | sym << 8;
#pragma unroll
                for (int q = 0; q < 8; q++) rec->incl[q] = m[q];
            }
        }
        // ---- apply the step: counts of all valid non-hit events (an escape that rescaled does not enter its symbol)
        if (valid && nonhit && !(esc && trig)) atomicAdd(&cnt[sym], 1u);
        __syncthreads();
        const uint32_t applied = first == 0xFFFFFFFFu ? step : first + 1;
        // totals through the last applied event (thread `applied-1` owns them)
        if (tid == applied - 1) {
            const uint32_t hits_incl = (tid - A) + (hit ? 1 : 0);
            if (first == 0xFFFFFFFFu) { s_f256 = s_f256 + hits_incl; s_f257 = f257 + (esc ? 1 : 0) - (two ? 1 : 0); }
            else { s_f256 = (s_f256 + hits_incl + 1) >> 1; }                    // flag 256 -> (f+1)/2 (cr-o2model.c:67)
        }
        __syncthreads();
        if (first != 0xFFFFFFFFu) {                                             // rescale: halve, count the ones (cr-o2model.c:55-68)
            uint32_t one = 0;
            if (tid < 256) { const uint32_t c = cnt[tid] >> 1; cnt[tid] = c; one = c == 1; }
            const uint32_t ones = __syncthreads_count(one);
            if (tid == 0) s_f257 = (1 + ones) & 255;
        }
        pos += applied;
        cap = O2C_THREADS;
        __syncthreads();
    }
    if (tid < 256) row[tid] = (uint8_t)cnt[tid];
    if (tid == 0) { row[256] = (uint8_t)s_f256; row[257] = (uint8_t)s_f257; }
}

// ------------------------------------------------------------------ o2 pass for HOT contexts, second form (default)
// Same step as k_o2_pass_cta -- every quantity of ppm_encode is a start-of-step value plus a rank, the first event whose update would
// rescale ends the step -- with the per-step latency cut (the hottest context is one chain of ~events / 200 steps and sets the time of the
// whole pass, profiles/round2_summary.md section 4):
//   * the events of the next two steps sit in a shared-memory ring filled by cp.async two steps ahead: no global load on the chain;
//   * ranks inside a warp come from eight ballots of the symbol bits (lanes with a smaller / the same symbol, lanes whose symbol is my
//     predicted byte) instead of a 32-step shuffle loop;
//   * the per-warp histogram is written by one leader lane per distinct symbol (no shared-memory atomics, no bank conflicts), and the
//     counts are bumped once per (warp, symbol);
//   * warp 0 rebuilds the cumulative table while the other warps already rank the next step; six barriers per step instead of ten.
CR_D void cr_cp_async4(void* smem, const void* gmem) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(sa), "l"(gmem) : "memory");
}
CR_D void cr_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> CR_D void cr_cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

struct O2Step { uint32_t c16, pos, applied, f256, f257, pad[3]; };
struct O2SplitCtl { uint32_t hot_count, rec_count, rec_cap, overflowed; };
template <int TH> struct O2HotSmem {
    static constexpr int NW = TH / 32, RING = 2 * TH;
    uint32_t ringK[RING], ringV[RING];
    uint32_t cnt[256], cumt[256], zmask[8];
    uint32_t wtot[3][32];
    uint32_t s_f256, s_f257, s_body, s_first;
    alignas(16) uint16_t hexcl[NW][256];           // nonhit events of earlier warps of the step, per symbol
    alignas(16) uint16_t below[NW][256];           // the same summed over smaller symbols
    uint8_t hraw[NW][256];             // nonhit events of this warp, per symbol (written by the leader lanes, cleared by them)
    uint16_t ssym[TH];
    uint8_t esc_sym[TH];
};
template <int TH>
__global__ void __launch_bounds__(TH) k_o2_hot(const uint32_t* __restrict__ K, const uint32_t* __restrict__ V, uint32_t n, PpmState st,
                                               uint64_t* __restrict__ T1, EscRec* __restrict__ esc_rec, uint32_t* __restrict__ esc_count, const uint32_t* __restrict__ bounds,
                                               const uint32_t* __restrict__ hotlist, const O2SplitCtl* __restrict__ ctl, const uint8_t* __restrict__ todo) {
    extern __shared__ __align__(16) unsigned char o2hot_raw[];
    O2HotSmem<TH>& S = *reinterpret_cast<O2HotSmem<TH>*>(o2hot_raw);
    constexpr int NW = TH / 32, RING = 2 * TH;
    // hotlist == nullptr: one CTA per ctx16; otherwise: the contexts k_o2_skel has marked as not done (ran out of step records)
    if (hotlist && (blockIdx.x >= ctl->hot_count || !ctl->overflowed)) return;
    const uint32_t c16 = hotlist ? hotlist[blockIdx.x] : blockIdx.x;
    if (hotlist && !todo[c16]) return;
    const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint32_t r0 = bounds[c16], r1 = bounds[c16 + 1];
    if (r1 - r0 < O2C_MIN) return;
    uint8_t* row = st.o2 + (size_t)c16 * PPM_O2_STRIDE;
    for (uint32_t i = tid; i < 256; i += TH) S.cnt[i] = row[i];
    for (uint32_t i = tid; i < NW * 64; i += TH) ((uint32_t*)&S.hraw[0][0])[i] = 0;
    if (tid == 0) { S.s_f256 = row[256]; S.s_f257 = row[257]; S.s_first = 0xFFFFFFFFu; }
    for (uint32_t i = tid; i < RING; i += TH) if (r0 + i < r1) { cr_cp_async4(&S.ringK[i], K + r0 + i); cr_cp_async4(&S.ringV[i], V + r0 + i); }
    cr_cp_async_commit();
    cr_cp_async_wait<0>();
    __syncthreads();

    const uint32_t before = (1u << lane) - 1u;
    uint32_t pos = r0;
    // an escape's record is written one step late: its slot comes from a global atomic whose round trip (~1000 cycles) would
    // otherwise sit on the chain of steps.  The thread keeps the record in registers and stores it in the next step (or after the loop).
    bool pend = false; uint32_t pend_slot = 0, pend_e = 0, pend_info = 0, pm0 = 0, pm1 = 0, pm2 = 0, pm3 = 0, pm4 = 0, pm5 = 0, pm6 = 0, pm7 = 0;
    auto flush = [&]() {
        if (pend) {
            EscRec* rec = esc_rec + pend_slot;
            rec->e = pend_e; rec->info = pend_info;
            rec->incl[0] = pm0; rec->incl[1] = pm1; rec->incl[2] = pm2; rec->incl[3] = pm3; rec->incl[4] = pm4; rec->incl[5] = pm5; rec->incl[6] = pm6; rec->incl[7] = pm7;
            pend = false;
        }
    };
    while (pos < r1) {
        const uint32_t step = r1 - pos < (uint32_t)TH ? r1 - pos : (uint32_t)TH;
        const uint32_t nw = (step + 31) >> 5;
        // ---- 1: warp 0 derives cumt / body / zero mask from cnt; everybody ranks its event inside its warp
        if (w == 0) {
            uint32_t v[8], sum = 0, zb = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) { v[k] = S.cnt[lane * 8 + k]; sum += v[k]; zb |= (uint32_t)(v[k] == 0) << k; }
            uint32_t inc = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(FULLMASK, inc, d); if (lane >= d) inc += t; }
            uint32_t run = inc - sum;
#pragma unroll
            for (int k = 0; k < 8; k++) { S.cumt[lane * 8 + k] = run; run += v[k]; }
            if (lane == 31) S.s_body = inc;
            uint32_t z = zb << (8 * (lane & 3));
            z |= __shfl_xor_sync(FULLMASK, z, 1); z |= __shfl_xor_sync(FULLMASK, z, 2);
            if ((lane & 3) == 0) S.zmask[lane >> 2] = z;
        }
        const bool active = tid < step;
        const uint32_t slot = (pos - r0 + tid) % RING;
        uint32_t k = 0, ev = 0;
        if (active) { k = S.ringK[slot]; ev = S.ringV[slot]; }
        const uint32_t sym = k >> 24, pr = (k >> 16) & 255;
        const bool hit = active && sym == pr, nonhit = active && sym != pr;
        const uint32_t b_nh = __ballot_sync(FULLMASK, nonhit);
        uint32_t mL = 0, mE = b_nh, mP = b_nh;        // nonhit lanes with a smaller symbol / my symbol / my predicted byte as their symbol
#pragma unroll
        for (int b = 7; b >= 0; b--) {
            const uint32_t Bb = __ballot_sync(FULLMASK, (sym >> b) & 1u);
            if ((sym >> b) & 1u) { mL |= mE & ~Bb; mE &= Bb; } else mE &= ~Bb;
            mP &= ((pr >> b) & 1u) ? Bb : ~Bb;
        }
        const uint32_t lt = __popc(mL & before), eq = __popc(mE & before), peq = __popc(mP & before);
        const bool leader = nonhit && (mE & before) == 0;
        if (leader) S.hraw[w][sym] = (uint8_t)__popc(mE);
        S.ssym[tid] = nonhit ? (uint16_t)sym : (uint16_t)0x100;
        __syncthreads();                                                             // B1
        // ---- 2a: per symbol, exclusive prefix over the warps of the step
        if (tid < 256) { uint32_t run = 0; for (uint32_t q = 0; q < nw; q++) { const uint32_t h = S.hraw[q][tid]; S.hexcl[q][tid] = (uint16_t)run; run += h; } }
        __syncthreads();                                                             // B2
        // ---- 2b: per warp, prefix over the symbols; frequencies of my symbol and of my predicted byte as of my event
        if (w < nw) {
            if (leader) S.hraw[w][sym] = 0;
            uint32_t v[8], sum = 0;
            const uint4 hv = *(const uint4*)&S.hexcl[w][lane * 8];
            v[0] = hv.x & 0xffff; v[1] = hv.x >> 16; v[2] = hv.y & 0xffff; v[3] = hv.y >> 16; v[4] = hv.z & 0xffff; v[5] = hv.z >> 16; v[6] = hv.w & 0xffff; v[7] = hv.w >> 16;
#pragma unroll
            for (int q = 0; q < 8; q++) sum += v[q];
            uint32_t inc = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(FULLMASK, inc, d); if (lane >= d) inc += t; }
            uint32_t run = inc - sum, o[8];
#pragma unroll
            for (int q = 0; q < 8; q++) { o[q] = run; run += v[q]; }
            *(uint4*)&S.below[w][lane * 8] = make_uint4(o[0] | o[1] << 16, o[2] | o[3] << 16, o[4] | o[5] << 16, o[6] | o[7] << 16);
        }
        __syncwarp();
        uint32_t fs = 0, pf = 0;
        if (active) { fs = S.cnt[sym] + S.hexcl[w][sym] + eq; pf = S.cnt[pr] + S.hexcl[w][pr] + peq; }
        const bool esc = nonhit && fs == 0, two = nonhit && fs == 1;
        const uint32_t b_es = __ballot_sync(FULLMASK, esc), b_tw = __ballot_sync(FULLMASK, two);
        if (lane == 0) { S.wtot[0][w] = __popc(b_nh); S.wtot[1][w] = __popc(b_es); S.wtot[2][w] = __popc(b_tw); }
        __syncthreads();                                                             // B3
        // ---- 3: totals of the earlier warps; does my event's update rescale the table?
        uint32_t p0 = 0, p1 = 0, p2 = 0;
        if (lane < w) { p0 = S.wtot[0][lane]; p1 = S.wtot[1][lane]; p2 = S.wtot[2][lane]; }
        p0 = __reduce_add_sync(FULLMASK, p0); p1 = __reduce_add_sync(FULLMASK, p1); p2 = __reduce_add_sync(FULLMASK, p2);
        const uint32_t A = p0 + __popc(b_nh & before);            // non-hits before this event
        const uint32_t ES = p1 + __popc(b_es & before);           // escapes before
        const uint32_t TW = p2 + __popc(b_tw & before);           // 1->2 transitions before
        const uint32_t f256 = S.s_f256 + (tid - A), f257 = S.s_f257 + ES - TW, body = S.s_body + A;
        bool trig = false;
        if (hit) trig = f256 + 1 > 250;
        else if (esc) trig = f257 + 1 > 250;
        else if (nonhit) trig = (fs + 1 > 250) || (fs == 1 && f257 == 0);
        const uint32_t b_tr = __ballot_sync(FULLMASK, trig);
        if (lane == 0 && b_tr) atomicMin(&S.s_first, w * 32 + __ffs(b_tr) - 1);
        if (esc) S.esc_sym[ES] = (uint8_t)sym;
        __syncthreads();                                                             // B4
        // ---- 4: results of the events up to and including the first trigger; their counts
        const uint32_t first = S.s_first;
        const bool valid = active && tid <= first;
        flush();                                                                     // the record of an earlier step: its slot has arrived
        if (valid) {
            const uint32_t sum = body + f256 + f257 - pf;
            if (hit) T1[ev] = ppm_pack(body - pf, f256, sum, 0);
            else if (!esc) T1[ev] = ppm_pack(S.cumt[sym] + S.below[w][sym] + lt - (sym >= pr ? pf : 0), fs, sum, 0);
            else {
                T1[ev] = ppm_pack(body + f256 - pf, f257, sum, 1);
                uint32_t m0, m1, m2, m3, m4, m5, m6, m7;
                if (trig) {
                    // the escape update rescales BEFORE the mask is taken (cr-ppm.c:146-151): zero <=> count <= 1 now
                    auto word = [&](uint32_t q) {
                        uint32_t r = 0;
                        for (uint32_t bb = 0; bb < 32; bb++) {
                            const uint32_t x = q * 32 + bb;
                            uint32_t c = S.cnt[x] + S.hexcl[w][x];
                            for (uint32_t j = w * 32; j < tid && c < 2; j++) c += S.ssym[j] == x;
                            if (c < 2) r |= 1u << bb;
                        }
                        return r;
                    };
                    m0 = word(0); m1 = word(1); m2 = word(2); m3 = word(3); m4 = word(4); m5 = word(5); m6 = word(6); m7 = word(7);
                } else {
                    m0 = S.zmask[0]; m1 = S.zmask[1]; m2 = S.zmask[2]; m3 = S.zmask[3]; m4 = S.zmask[4]; m5 = S.zmask[5]; m6 = S.zmask[6]; m7 = S.zmask[7];
                }
                auto drop = [&](uint32_t x) {                      // clear bit x (no indexed registers: no local memory)
                    const uint32_t q = x >> 5, bit = ~(1u << (x & 31));
                    m0 &= q == 0 ? bit : ~0u; m1 &= q == 1 ? bit : ~0u; m2 &= q == 2 ? bit : ~0u; m3 &= q == 3 ? bit : ~0u;
                    m4 &= q == 4 ? bit : ~0u; m5 &= q == 5 ? bit : ~0u; m6 &= q == 6 ? bit : ~0u; m7 &= q == 7 ? bit : ~0u;
                };
                if (!trig) for (uint32_t e = 0; e < ES; e++) drop(S.esc_sym[e]);
                drop(pr);
                pend = true; pend_slot = atomicAdd(esc_count, 1u); pend_e = ev; pend_info = (c16 & 0xff) | sym << 8;
                pm0 = m0; pm1 = m1; pm2 = m2; pm3 = m3; pm4 = m4; pm5 = m5; pm6 = m6; pm7 = m7;
            }
        }
        const uint32_t applied = first == 0xFFFFFFFFu ? step : first + 1;
        // events that enter their symbol's count: valid non-hits (an escape that rescaled does not); one bump per (warp, symbol)
        const uint32_t b_ap = __ballot_sync(FULLMASK, valid && nonhit && !(esc && trig));
        if (leader && (mE & b_ap)) atomicAdd(&S.cnt[sym], (uint32_t)__popc(mE & b_ap));
        if (tid == applied - 1) {                                                    // totals through the last applied event
            const uint32_t hits_incl = (tid - A) + (hit ? 1 : 0);
            if (first == 0xFFFFFFFFu) { S.s_f256 = S.s_f256 + hits_incl; S.s_f257 = f257 + (esc ? 1 : 0) - (two ? 1 : 0); }
            else S.s_f256 = (S.s_f256 + hits_incl + 1) >> 1;                         // flag 256 -> (f+1)/2 (cr-o2model.c:67)
        }
        // the slots this step has consumed take the events two windows ahead
        if (tid < applied) {
            const uint32_t p = pos + RING + tid;
            if (p < r1) { cr_cp_async4(&S.ringK[slot], K + p); cr_cp_async4(&S.ringV[slot], V + p); }
        }
        cr_cp_async_commit();
        __syncthreads();                                                             // B5
        if (tid == 0) S.s_first = 0xFFFFFFFFu;
        cr_cp_async_wait<1>();                                                       // everything but the group just issued has landed
        if (first != 0xFFFFFFFFu) {                                                  // rescale: halve, count the ones (cr-o2model.c:55-68)
            uint32_t one = 0;
            if (tid < 256) { const uint32_t c = S.cnt[tid] >> 1; S.cnt[tid] = c; one = c == 1; }
            const uint32_t ones = __syncthreads_count(one);                          // B6
            if (tid == 0) S.s_f257 = (1 + ones) & 255;
        } else __syncthreads();                                                      // B6
        pos += applied;
    }
    flush();
    cr_cp_async_wait<0>();
    __syncthreads();
    if (tid < 256) row[tid] = (uint8_t)S.cnt[tid];
    if (tid == 0) { row[256] = (uint8_t)S.s_f256; row[257] = (uint8_t)S.s_f257; }
}

// ------------------------------------------------------------------ o2 pass for HOT contexts, split form (default)
// As for o1 (k_o1_skel / k_o1_eval below): what carries from step to step is the table of counts, flags 256 / 257 and where the table is
// rescaled -- and that needs, per event, only "is it a hit", the count of its own symbol and whether it is an escape or a 1 -> 2
// transition.  Cumulative frequencies, the predicted byte's count, exclusion masks and the triples themselves do not feed the chain.
//   k_o2_hotlist  the contexts with at least O2C_MIN events, as a list (the grids below cover the list, not all 65536 contexts)
//   k_o2_skel     one CTA per hot context walks the steps and records (position, length, flags, table) per step; the record index
//                 comes from a global counter one step ahead.  Out of records (cannot happen within the bound the host allocates, see
//                 LzChain): the context is marked, its row left untouched, and k_o2_hot redoes it
//   k_o2_eval     one CTA per RECORDED STEP: the triples and escape records of k_o2_hot's step, from the recorded table
__global__ void k_o2_hotlist(const uint32_t* __restrict__ bounds, uint32_t* __restrict__ hotlist, O2SplitCtl* __restrict__ ctl, uint8_t* __restrict__ todo) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= 65536) return;
    todo[c] = 0;
    if (bounds[c + 1] - bounds[c] >= O2C_MIN) hotlist[atomicAdd(&ctl->hot_count, 1u)] = c;
}
template <int TH>
__global__ void __launch_bounds__(TH) k_o2_skel(const uint32_t* __restrict__ K, PpmState st, const uint32_t* __restrict__ bounds, const uint32_t* __restrict__ hotlist,
                                                O2SplitCtl* __restrict__ ctl, uint8_t* __restrict__ todo, O2Step* __restrict__ steps, uint8_t* __restrict__ snaps) {
    constexpr int NW = TH / 32, RING = 2 * TH;
    if (blockIdx.x >= ctl->hot_count) return;
    const uint32_t c16 = hotlist[blockIdx.x];
    const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint32_t r0 = bounds[c16], r1 = bounds[c16 + 1];
    __shared__ uint32_t ringK[RING];
    __shared__ __align__(16) uint32_t cnt[256];
    __shared__ uint32_t wtot[3][32];
    __shared__ uint32_t s_f256, s_f257, s_first, s_idx;
    __shared__ uint16_t hexcl[NW][256];
    __shared__ uint8_t hraw[NW][256];
    uint8_t* row = st.o2 + (size_t)c16 * PPM_O2_STRIDE;
    for (uint32_t i = tid; i < 256; i += TH) cnt[i] = row[i];
    for (uint32_t i = tid; i < NW * 64; i += TH) ((uint32_t*)&hraw[0][0])[i] = 0;
    uint32_t next_idx = 0;
    if (tid == 0) { s_f256 = row[256]; s_f257 = row[257]; s_first = 0xFFFFFFFFu; next_idx = atomicAdd(&ctl->rec_count, 1u); }
    for (uint32_t i = tid; i < RING; i += TH) if (r0 + i < r1) cr_cp_async4(&ringK[i], K + r0 + i);
    cr_cp_async_commit();
    cr_cp_async_wait<0>();
    __syncthreads();
    const uint32_t before = (1u << lane) - 1u;
    const uint32_t cap = ctl->rec_cap;
    uint32_t pos = r0;
    bool gave_up = false;
    while (pos < r1) {
        const uint32_t step = r1 - pos < (uint32_t)TH ? r1 - pos : (uint32_t)TH;
        const uint32_t nw = (step + 31) >> 5;
        if (tid == 0) { s_idx = next_idx; next_idx = atomicAdd(&ctl->rec_count, 1u); }       // the next step's record: its round trip hides behind this step
        const bool active = tid < step;
        const uint32_t slot = (pos - r0 + tid) % RING;
        const uint32_t k = active ? ringK[slot] : 0u;
        const uint32_t sym = k >> 24, pr = (k >> 16) & 255;
        const bool hit = active && sym == pr, nonhit = active && sym != pr;
        const uint32_t b_nh = __ballot_sync(FULLMASK, nonhit);
        uint32_t mE = b_nh;
#pragma unroll
        for (int b = 7; b >= 0; b--) {
            const uint32_t Bb = __ballot_sync(FULLMASK, (sym >> b) & 1u);
            mE &= ((sym >> b) & 1u) ? Bb : ~Bb;
        }
        const uint32_t eq = __popc(mE & before);
        const bool leader = nonhit && eq == 0;
        if (leader) hraw[w][sym] = (uint8_t)__popc(mE);
        __syncthreads();                                                             // B1
        const uint32_t idx = s_idx;
        if (idx >= cap) { gave_up = true; break; }                                   // uniform
        if (tid < 64) {                                                              // the table as of the start of this step
            const uint4 a = *(const uint4*)&cnt[tid * 4];
            ((uint32_t*)(snaps + (size_t)idx * 256))[tid] = a.x | a.y << 8 | a.z << 16 | a.w << 24;
        }
        if (tid < 256) { uint32_t run = 0; for (uint32_t q = 0; q < nw; q++) { const uint32_t h = hraw[q][tid]; hexcl[q][tid] = (uint16_t)run; run += h; } }
        __syncthreads();                                                             // B2
        if (leader) hraw[w][sym] = 0;
        const uint32_t fs = nonhit ? cnt[sym] + hexcl[w][sym] + eq : 0u;
        const bool esc = nonhit && fs == 0, two = nonhit && fs == 1;
        const uint32_t b_es = __ballot_sync(FULLMASK, esc), b_tw = __ballot_sync(FULLMASK, two);
        if (lane == 0) { wtot[0][w] = __popc(b_nh); wtot[1][w] = __popc(b_es); wtot[2][w] = __popc(b_tw); }
        __syncthreads();                                                             // B3
        uint32_t p0 = 0, p1 = 0, p2 = 0;
        if (lane < w) { p0 = wtot[0][lane]; p1 = wtot[1][lane]; p2 = wtot[2][lane]; }
        p0 = __reduce_add_sync(FULLMASK, p0); p1 = __reduce_add_sync(FULLMASK, p1); p2 = __reduce_add_sync(FULLMASK, p2);
        const uint32_t A = p0 + __popc(b_nh & before), ES = p1 + __popc(b_es & before), TW = p2 + __popc(b_tw & before);
        const uint32_t f256_0 = s_f256, f257_0 = s_f257;
        const uint32_t f256 = f256_0 + (tid - A), f257 = f257_0 + ES - TW;
        bool trig = false;
        if (hit) trig = f256 + 1 > 250;
        else if (esc) trig = f257 + 1 > 250;
        else if (nonhit) trig = (fs + 1 > 250) || (fs == 1 && f257 == 0);
        const uint32_t b_tr = __ballot_sync(FULLMASK, trig);
        if (lane == 0 && b_tr) atomicMin(&s_first, w * 32 + __ffs(b_tr) - 1);
        __syncthreads();                                                             // B4
        const uint32_t first = s_first;
        const bool valid = active && tid <= first;
        const uint32_t applied = first == 0xFFFFFFFFu ? step : first + 1;
        const uint32_t b_ap = __ballot_sync(FULLMASK, valid && nonhit && !(esc && trig));
        if (leader && (mE & b_ap)) atomicAdd(&cnt[sym], (uint32_t)__popc(mE & b_ap));
        if (tid == applied - 1) {
            const uint32_t hits_incl = (tid - A) + (hit ? 1 : 0);
            if (first == 0xFFFFFFFFu) { s_f256 = f256_0 + hits_incl; s_f257 = f257 + (esc ? 1 : 0) - (two ? 1 : 0); }
            else s_f256 = (f256_0 + hits_incl + 1) >> 1;
        }
        if (tid == 0) { O2Step r; r.c16 = c16; r.pos = pos; r.applied = applied; r.f256 = f256_0; r.f257 = f257_0; r.pad[0] = r.pad[1] = r.pad[2] = 0; steps[idx] = r; }
        if (tid < applied) { const uint32_t p = pos + RING + tid; if (p < r1) cr_cp_async4(&ringK[slot], K + p); }
        cr_cp_async_commit();
        __syncthreads();                                                             // B5
        if (tid == 0) s_first = 0xFFFFFFFFu;
        cr_cp_async_wait<1>();
        if (first != 0xFFFFFFFFu) {
            uint32_t one = 0;
            if (tid < 256) { const uint32_t c = cnt[tid] >> 1; cnt[tid] = c; one = c == 1; }
            const uint32_t ones = __syncthreads_count(one);                          // B6
            if (tid == 0) s_f257 = (1 + ones) & 255;
        } else __syncthreads();                                                      // B6
        pos += applied;
    }
    cr_cp_async_wait<0>();
    __syncthreads();
    if (tid == 0 && next_idx < cap) { O2Step r; memset(&r, 0, sizeof r); steps[next_idx] = r; }   // the record reserved ahead and not used
    if (gave_up) { if (tid == 0) { todo[c16] = 1; ctl->overflowed = 1; } return; }
    if (tid < 256) row[tid] = (uint8_t)cnt[tid];
    if (tid == 0) { row[256] = (uint8_t)s_f256; row[257] = (uint8_t)s_f257; }
}
template <int TH> struct O2EvalSmem {
    static constexpr int NW = TH / 32;
    uint32_t cnt[256], cumt[256], zmask[8];
    uint32_t wtot[3][32];
    uint32_t s_body, pad_[3];
    alignas(16) uint16_t hexcl[NW][256];
    alignas(16) uint16_t below[NW][256];
    uint8_t hraw[NW][256];
    uint16_t ssym[TH];
    uint8_t esc_sym[TH];
};
template <int TH>
__global__ void __launch_bounds__(TH) k_o2_eval(const uint32_t* __restrict__ K, const uint32_t* __restrict__ V, const O2SplitCtl* __restrict__ ctl, const uint8_t* __restrict__ todo,
                                                const O2Step* __restrict__ steps, const uint8_t* __restrict__ snaps,
                                                uint64_t* __restrict__ T1, EscRec* __restrict__ esc_rec, uint32_t* __restrict__ esc_count) {
    extern __shared__ __align__(16) unsigned char o2eval_raw[];
    O2EvalSmem<TH>& S = *reinterpret_cast<O2EvalSmem<TH>*>(o2eval_raw);
    constexpr int NW = TH / 32;
    const uint32_t nrec = ctl->rec_count < ctl->rec_cap ? ctl->rec_count : ctl->rec_cap;
    const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (uint32_t rec_i = blockIdx.x; rec_i < nrec; rec_i += gridDim.x) {
    const O2Step R = steps[rec_i];
    if (R.applied == 0 || todo[R.c16]) continue;                                     // uniform
    __syncthreads();                                                                 // the previous record's tables are no longer read
    const uint32_t c16 = R.c16, step = R.applied;
    for (uint32_t i = tid; i < 256; i += TH) S.cnt[i] = snaps[(size_t)rec_i * 256 + i];
    for (uint32_t i = tid; i < NW * 64; i += TH) ((uint32_t*)&S.hraw[0][0])[i] = 0;
    __syncthreads();
    const uint32_t nw = (step + 31) >> 5;
    const uint32_t before = (1u << lane) - 1u;
    if (w == 0) {
        uint32_t v[8], sum = 0, zb = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { v[k] = S.cnt[lane * 8 + k]; sum += v[k]; zb |= (uint32_t)(v[k] == 0) << k; }
        uint32_t inc = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(FULLMASK, inc, d); if (lane >= d) inc += t; }
        uint32_t run = inc - sum;
#pragma unroll
        for (int k = 0; k < 8; k++) { S.cumt[lane * 8 + k] = run; run += v[k]; }
        if (lane == 31) S.s_body = inc;
        uint32_t z = zb << (8 * (lane & 3));
        z |= __shfl_xor_sync(FULLMASK, z, 1); z |= __shfl_xor_sync(FULLMASK, z, 2);
        if ((lane & 3) == 0) S.zmask[lane >> 2] = z;
    }
    const bool active = tid < step;
    uint32_t k = 0, ev = 0;
    if (active) { k = K[R.pos + tid]; ev = V[R.pos + tid]; }
    const uint32_t sym = k >> 24, pr = (k >> 16) & 255;
    const bool hit = active && sym == pr, nonhit = active && sym != pr;
    const uint32_t b_nh = __ballot_sync(FULLMASK, nonhit);
    uint32_t mL = 0, mE = b_nh, mP = b_nh;
#pragma unroll
    for (int b = 7; b >= 0; b--) {
        const uint32_t Bb = __ballot_sync(FULLMASK, (sym >> b) & 1u);
        if ((sym >> b) & 1u) { mL |= mE & ~Bb; mE &= Bb; } else mE &= ~Bb;
        mP &= ((pr >> b) & 1u) ? Bb : ~Bb;
    }
    const uint32_t lt = __popc(mL & before), eq = __popc(mE & before), peq = __popc(mP & before);
    const bool leader = nonhit && (mE & before) == 0;
    if (leader) S.hraw[w][sym] = (uint8_t)__popc(mE);
    S.ssym[tid] = nonhit ? (uint16_t)sym : (uint16_t)0x100;
    __syncthreads();
    if (tid < 256) { uint32_t run = 0; for (uint32_t q = 0; q < nw; q++) { const uint32_t h = S.hraw[q][tid]; S.hexcl[q][tid] = (uint16_t)run; run += h; } }
    __syncthreads();
    if (w < nw) {
        uint32_t v[8], sum = 0;
        const uint4 hv = *(const uint4*)&S.hexcl[w][lane * 8];
        v[0] = hv.x & 0xffff; v[1] = hv.x >> 16; v[2] = hv.y & 0xffff; v[3] = hv.y >> 16; v[4] = hv.z & 0xffff; v[5] = hv.z >> 16; v[6] = hv.w & 0xffff; v[7] = hv.w >> 16;
#pragma unroll
        for (int q = 0; q < 8; q++) sum += v[q];
        uint32_t inc = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(FULLMASK, inc, d); if (lane >= d) inc += t; }
        uint32_t run = inc - sum, o[8];
#pragma unroll
        for (int q = 0; q < 8; q++) { o[q] = run; run += v[q]; }
        *(uint4*)&S.below[w][lane * 8] = make_uint4(o[0] | o[1] << 16, o[2] | o[3] << 16, o[4] | o[5] << 16, o[6] | o[7] << 16);
    }
    __syncwarp();
    uint32_t fs = 0, pf = 0;
    if (active) { fs = S.cnt[sym] + S.hexcl[w][sym] + eq; pf = S.cnt[pr] + S.hexcl[w][pr] + peq; }
    const bool esc = nonhit && fs == 0, two = nonhit && fs == 1;
    const uint32_t b_es = __ballot_sync(FULLMASK, esc), b_tw = __ballot_sync(FULLMASK, two);
    if (lane == 0) { S.wtot[0][w] = __popc(b_nh); S.wtot[1][w] = __popc(b_es); S.wtot[2][w] = __popc(b_tw); }
    __syncthreads();
    uint32_t p0 = 0, p1 = 0, p2 = 0;
    if (lane < w) { p0 = S.wtot[0][lane]; p1 = S.wtot[1][lane]; p2 = S.wtot[2][lane]; }
    p0 = __reduce_add_sync(FULLMASK, p0); p1 = __reduce_add_sync(FULLMASK, p1); p2 = __reduce_add_sync(FULLMASK, p2);
    const uint32_t A = p0 + __popc(b_nh & before), ES = p1 + __popc(b_es & before), TW = p2 + __popc(b_tw & before);
    const uint32_t f256 = R.f256 + (tid - A), f257 = R.f257 + ES - TW, body = S.s_body + A;
    if (esc) S.esc_sym[ES] = (uint8_t)sym;
    __syncthreads();
    if (!active) continue;
    const uint32_t sum = body + f256 + f257 - pf;
    if (hit) T1[ev] = ppm_pack(body - pf, f256, sum, 0);
    else if (!esc) T1[ev] = ppm_pack(S.cumt[sym] + S.below[w][sym] + lt - (sym >= pr ? pf : 0), fs, sum, 0);
    else {
        T1[ev] = ppm_pack(body + f256 - pf, f257, sum, 1);
        const bool trig = f257 + 1 > 250;
        uint32_t m0, m1, m2, m3, m4, m5, m6, m7;
        if (trig) {
            // the escape update rescales BEFORE the mask is taken (cr-ppm.c:146-151): zero <=> count <= 1 now
            auto word = [&](uint32_t q) {
                uint32_t r = 0;
                for (uint32_t bb = 0; bb < 32; bb++) {
                    const uint32_t x = q * 32 + bb;
                    uint32_t c = S.cnt[x] + S.hexcl[w][x];
                    for (uint32_t j = w * 32; j < tid && c < 2; j++) c += S.ssym[j] == x;
                    if (c < 2) r |= 1u << bb;
                }
                return r;
            };
            m0 = word(0); m1 = word(1); m2 = word(2); m3 = word(3); m4 = word(4); m5 = word(5); m6 = word(6); m7 = word(7);
        } else {
            m0 = S.zmask[0]; m1 = S.zmask[1]; m2 = S.zmask[2]; m3 = S.zmask[3]; m4 = S.zmask[4]; m5 = S.zmask[5]; m6 = S.zmask[6]; m7 = S.zmask[7];
        }
        auto drop = [&](uint32_t x) {
            const uint32_t q = x >> 5, bit = ~(1u << (x & 31));
            m0 &= q == 0 ? bit : ~0u; m1 &= q == 1 ? bit : ~0u; m2 &= q == 2 ? bit : ~0u; m3 &= q == 3 ? bit : ~0u;
            m4 &= q == 4 ? bit : ~0u; m5 &= q == 5 ? bit : ~0u; m6 &= q == 6 ? bit : ~0u; m7 &= q == 7 ? bit : ~0u;
        };
        if (!trig) for (uint32_t e = 0; e < ES; e++) drop(S.esc_sym[e]);
        drop(pr);
        EscRec* rec = esc_rec + atomicAdd(esc_count, 1u);
        rec->e = ev; rec->info = (c16 & 0xff) | sym << 8;
        rec->incl[0] = m0; rec->incl[1] = m1; rec->incl[2] = m2; rec->incl[3] = m3; rec->incl[4] = m4; rec->incl[5] = m5; rec->incl[6] = m6; rec->incl[7] = m7;
    }
    }
}

// ------------------------------------------------------------------ o1 pass, one warp per ctx8
// Escapes arrive sorted by (ctx8, time).  k_o1_gather first makes that order physical, so the pass streams its
// input: 32 records per batch, loaded one batch ahead and staged through shared memory.
__global__ void k_o1_gather(const uint32_t* __restrict__ V, uint32_t n, const EscRec* __restrict__ rec, const uint32_t* __restrict__ ord,
                            uint32_t* __restrict__ info_s, uint32_t* __restrict__ ord_s, uint4* __restrict__ incl_s) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t v = V[i];
    info_s[i] = rec[v].info; ord_s[i] = ord[v];
    const uint32_t* m = rec[v].incl;
    incl_s[2 * (size_t)i] = make_uint4(m[0], m[1], m[2], m[3]);
    incl_s[2 * (size_t)i + 1] = make_uint4(m[4], m[5], m[6], m[7]);
}
// lane l holds o1 counts of symbols 8l..8l+7 in (a0, a1).
__global__ void __launch_bounds__(128) k_o1_pass_warp(const uint64_t* __restrict__ K, uint32_t n, const uint32_t* __restrict__ info_s, const uint32_t* __restrict__ ord_s,
                                                       const uint4* __restrict__ incl_s, PpmState st, uint64_t* __restrict__ T2, uint32_t hot_min) {
    __shared__ uint4 sincl[4][32][2];
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t c8 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c8 >= 256) return;
    auto keyof = [](uint64_t k) { return (uint32_t)(k >> 32) & 0xffu; };
    const uint32_t r0 = lower_bound_key(K, n, c8, keyof);
    if (r0 >= n || keyof(K[r0]) != c8) return;
    const uint32_t r1 = lower_bound_key(K, n, c8 + 1, keyof);
    if (r1 - r0 >= hot_min) return;                       // taken by k_o1_pass_cta
    uint8_t* row = st.o1 + c8 * 256;
    uint2 av = ((const uint2*)row)[lane];
    uint32_t a0 = av.x, a1 = av.y;

    uint32_t n_info = 0, n_ord = 0; uint4 n_i0 = make_uint4(0, 0, 0, 0), n_i1 = n_i0;
    if (r0 + lane < r1) { n_info = info_s[r0 + lane]; n_ord = ord_s[r0 + lane]; n_i0 = incl_s[2 * (size_t)(r0 + lane)]; n_i1 = incl_s[2 * (size_t)(r0 + lane) + 1]; }
    for (uint32_t base = r0; base < r1; base += 32) {
        const uint32_t c_info = n_info, c_ord = n_ord;
        sincl[w][lane][0] = n_i0; sincl[w][lane][1] = n_i1;
        __syncwarp();
        if (base + 32 + lane < r1) {
            const size_t x = base + 32 + lane;
            n_info = info_s[x]; n_ord = ord_s[x]; n_i0 = incl_s[2 * x]; n_i1 = incl_s[2 * x + 1];
        }
        const uint32_t cnt = r1 - base < 32 ? r1 - base : 32;
        uint64_t mine = 0;
        for (uint32_t j = 0; j < cnt; j++) {
            const uint32_t info = __shfl_sync(FULLMASK, c_info, j);
            const uint32_t bits = ((const uint8_t*)&sincl[w][j][0])[lane];
            const uint32_t sym = (info >> 8) & 255, own = sym >> 3, within = sym & 7;
            uint32_t sum = 0, cum = 0;
#pragma unroll
            for (uint32_t k = 0; k < 8; k++) {
                const uint32_t c = ((k < 4 ? a0 : a1) >> (8 * (k & 3))) & 255;
                const uint32_t fr = (bits >> k & 1u) ? c * 8 - 7 : 0u;      // M_freq_o1 (cr-ppm.c:98)
                sum += fr;
                if (lane < own || (lane == own && k < within)) cum += fr;
            }
            // one reduction for both: cum in the high half (sums stay below 2^20)
            const unsigned long long both = ((unsigned long long)cum << 32) | sum;
            const uint32_t lo = __reduce_add_sync(FULLMASK, (uint32_t)both), hi = __reduce_add_sync(FULLMASK, (uint32_t)(both >> 32));
            const uint32_t mysym = ((within < 4 ? a0 : a1) >> (8 * (within & 3))) & 255;
            const uint32_t cs = __shfl_sync(FULLMASK, mysym, own);
            if (lane == j) mine = ppm_pack(hi, cs * 8 - 7, lo, 0);
            // ppm_update_o1 (cr-ppm.c:90-97)
            if (lane == own) { if (within < 4) a0 += 1u << (8 * within); else a1 += 1u << (8 * (within - 4)); }
            if (cs + 1 >= 255) { a0 -= (a0 >> 1) & 0x7f7f7f7fu; a1 -= (a1 >> 1) & 0x7f7f7f7fu; }
        }
        if (lane < cnt) T2[c_ord] = mine;
        __syncwarp();
    }
    ((uint2*)row)[lane] = make_uint2(a0, a1);
}

// ------------------------------------------------------------------ o1 pass for HOT ctx8 rows, one CTA per row
// Same idea as k_o2_pass_cta: between two halvings an o1 row only counts occurrences, so for escape i
//   count_x(i) = count_x(0) + #{j < i with symbol x},  and with its own exclusion mask M_i (from the o2 pass)
//   sum_i = SUM_{x in M_i} (8 count_x(i) - 7),  cum_i = the same restricted to x < s_i      (cr-ppm.c:150-156).
// Each thread evaluates its masked sums against a per-warp table "start-of-step count + occurrences in earlier
// warps" and corrects for the earlier lanes of its own warp.  The first escape whose update halves the row
// (++o1[c] >= 255, cr-ppm.c:91) ends the step.
#define O1C_THREADS 512
#define O1C_WARPS   (O1C_THREADS / 32)
#define O1C_MIN     2048
__global__ void __launch_bounds__(O1C_THREADS) k_o1_pass_cta(const uint64_t* __restrict__ K, uint32_t n, const uint32_t* __restrict__ info_s, const uint32_t* __restrict__ ord_s,
                                                             const uint4* __restrict__ incl_s, PpmState st, uint64_t* __restrict__ T2) {
    const uint32_t c8 = blockIdx.x;
    const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    auto keyof = [](uint64_t k) { return (uint32_t)(k >> 32) & 0xffu; };
    const uint32_t r0 = lower_bound_key(K, n, c8, keyof);
    if (r0 >= n || keyof(K[r0]) != c8) return;
    const uint32_t r1 = lower_bound_key(K, n, c8 + 1, keyof);
    if (r1 - r0 < O1C_MIN) return;
    __shared__ uint32_t cnt[256];
    __shared__ __align__(4) uint16_t hist[O1C_WARPS][256];      // per-warp histogram -> exclusive prefix over warps
    __shared__ uint16_t basew[O1C_WARPS][256];                   // 8 * (cnt + earlier warps) - 7
    __shared__ uint32_t smask[O1C_THREADS][8];
    __shared__ uint32_t s_first;
    uint8_t* row = st.o1 + c8 * 256;
    if (tid < 256) cnt[tid] = row[tid];
    uint32_t pos = r0;
    for (;;) {
        for (uint32_t i = tid; i < O1C_WARPS * 128; i += O1C_THREADS) ((uint32_t*)&hist[0][0])[i] = 0;
        if (tid == 0) s_first = 0xFFFFFFFFu;
        __syncthreads();
        if (pos >= r1) break;
        const uint32_t step = r1 - pos < O1C_THREADS ? r1 - pos : O1C_THREADS;
        const bool active = tid < step;
        uint32_t sym = 0x1FF, ord = 0;
        uint32_t m[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (active) {
            const size_t x = pos + tid;
            sym = (info_s[x] >> 8) & 255; ord = ord_s[x];
            const uint4 a = incl_s[2 * x], b = incl_s[2 * x + 1];
            m[0] = a.x; m[1] = a.y; m[2] = a.z; m[3] = a.w; m[4] = b.x; m[5] = b.y; m[6] = b.z; m[7] = b.w;
            atomicAdd((uint32_t*)&hist[w][0] + (sym >> 1), (sym & 1u) ? 0x10000u : 1u);
        }
#pragma unroll
        for (int q = 0; q < 8; q++) smask[tid][q] = m[q];
        __syncthreads();
        if (tid < 256) {
            uint32_t run = 0;
            for (int q = 0; q < O1C_WARPS; q++) { uint32_t h = hist[q][tid]; hist[q][tid] = (uint16_t)run; basew[q][tid] = (uint16_t)(8 * (cnt[tid] + run) - 7); run += h; }
        }
        __syncthreads();
        // earlier lanes of this warp: same symbol / symbols inside my mask (and below my symbol)
        uint32_t eq = 0, in_all = 0, in_lt = 0;
#pragma unroll 8
        for (uint32_t j = 0; j < 32; j++) {
            const uint32_t sj = __shfl_sync(FULLMASK, sym, j);
            if (j < lane && sj < 256) {
                const uint32_t in = (smask[tid][sj >> 5] >> (sj & 31)) & 1u;
                eq += sj == sym; in_all += in; in_lt += in & (uint32_t)(sj < sym);
            }
        }
        uint32_t sum = 0, cum = 0, count_s = 0;
        if (active) {
#pragma unroll
            for (uint32_t q = 0; q < 8; q++) {
                const uint32_t mq = m[q];
                for (uint32_t b = 0; b < 32; b++) {
                    const uint32_t x = q * 32 + b;
                    const uint32_t v = (mq >> b & 1u) ? (uint32_t)basew[w][x] : 0u;
                    sum += v; cum += x < sym ? v : 0u;
                }
            }
            sum += 8 * in_all; cum += 8 * in_lt;
            count_s = cnt[sym] + hist[w][sym] + eq;
            if (count_s + 1 >= 255) atomicMin(&s_first, tid);
        }
        __syncthreads();
        const uint32_t first = s_first;
        const bool valid = active && tid <= first;
        if (valid) {
            T2[ord] = ppm_pack(cum, 8 * count_s - 7, sum, 0);
            atomicAdd(&cnt[sym], 1u);
        }
        __syncthreads();
        if (first != 0xFFFFFFFFu && tid < 256) cnt[tid] -= cnt[tid] / 2;           // cr-ppm.c:92-94
        pos += first == 0xFFFFFFFFu ? step : first + 1;
        __syncthreads();
    }
    if (tid < 256) row[tid] = (uint8_t)cnt[tid];
}

// ------------------------------------------------------------------ o1 pass for HOT ctx8 rows, split form (default)
// k_o1_pass_cta spends ~1300 instructions per thread and step on the masked sums, and the hottest row (the context "space" of a text
// holds more than half of all escapes) walks its steps one after the other on ONE SM while the others idle (ncu: sm__cycles_active
// avg 66 K vs max 5.6 M, profiles/round2_summary.md section 5).  But the masked sums do not feed the chain: what carries from step to step is only
// the row of counts and where it is halved, and that depends on the escapes' SYMBOLS alone.  So:
//   k_o1_plan   per ctx8: its range of sorted escapes and where its step records start (bound: n/512 + n/127 + 2 steps);
//   k_o1_skel   one CTA per hot row walks the steps -- symbols from a cp.async ring, ranks from ballots, per-warp counts by leader lanes,
//               the first escape whose update halves the row ends the step -- and records (position, length, row of counts) per step;
//   k_o1_eval   one CTA per RECORDED STEP, all rows, all steps at once: the masked sums of k_o1_pass_cta from the recorded row.
#define O1S_TH 512
#define O1S_RING 1024
struct O1Ctx { uint32_t r0, r1, rec_base, nsteps; };
struct O1Step { uint32_t pos, applied; };
CR_HD uint32_t o1_max_steps(uint32_t n) { return n / O1S_TH + n / 127u + 2u; }
__global__ void __launch_bounds__(256) k_o1_plan(const uint64_t* __restrict__ K, uint32_t n, O1Ctx* __restrict__ ctx) {
    __shared__ uint32_t wsum[8];
    const uint32_t c8 = threadIdx.x, lane = c8 & 31, w = c8 >> 5;
    auto keyof = [](uint64_t k) { return (uint32_t)(k >> 32) & 0xffu; };
    const uint32_t r0 = lower_bound_key(K, n, c8, keyof), r1 = lower_bound_key(K, n, c8 + 1, keyof);
    const uint32_t need = r1 - r0 >= O1C_MIN ? o1_max_steps(r1 - r0) : 0u;
    uint32_t incl = need;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(FULLMASK, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    uint32_t off = 0;
    for (uint32_t q = 0; q < w; q++) off += wsum[q];
    O1Ctx c; c.r0 = r0; c.r1 = r1; c.rec_base = off + incl - need; c.nsteps = 0;
    ctx[c8] = c;
}
__global__ void __launch_bounds__(O1S_TH) k_o1_skel(O1Ctx* __restrict__ ctx, const uint32_t* __restrict__ info_s, PpmState st,
                                                     O1Step* __restrict__ steps, uint8_t* __restrict__ snaps) {
    constexpr int NW = O1S_TH / 32;
    const uint32_t c8 = blockIdx.x;
    const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const O1Ctx C = ctx[c8];
    const uint32_t r0 = C.r0, r1 = C.r1;
    if (r1 - r0 < O1C_MIN) return;
    __shared__ uint32_t ring[O1S_RING];
    __shared__ __align__(16) uint32_t cnt[256];
    __shared__ uint16_t hexcl[NW][256];
    __shared__ uint8_t hraw[NW][256];
    __shared__ uint32_t s_first;
    uint8_t* row = st.o1 + c8 * 256;
    if (tid < 256) cnt[tid] = row[tid];
    for (uint32_t i = tid; i < NW * 64; i += O1S_TH) ((uint32_t*)&hraw[0][0])[i] = 0;
    if (tid == 0) s_first = 0xFFFFFFFFu;
    for (uint32_t i = tid; i < O1S_RING; i += O1S_TH) if (r0 + i < r1) cr_cp_async4(&ring[i], info_s + r0 + i);
    cr_cp_async_commit();
    cr_cp_async_wait<0>();
    __syncthreads();
    const uint32_t before = (1u << lane) - 1u;
    uint32_t pos = r0, nrec = 0;
    while (pos < r1) {
        const uint32_t step = r1 - pos < (uint32_t)O1S_TH ? r1 - pos : (uint32_t)O1S_TH;
        const uint32_t nw = (step + 31) >> 5;
        // the row as of the start of this step
        if (tid < 64) {
            const uint4 a = *(const uint4*)&cnt[tid * 4];
            ((uint32_t*)(snaps + (size_t)(C.rec_base + nrec) * 256))[tid] = a.x | a.y << 8 | a.z << 16 | a.w << 24;
        }
        const bool active = tid < step;
        const uint32_t slot = (pos - r0 + tid) % O1S_RING;
        const uint32_t sym = active ? (ring[slot] >> 8) & 255u : 0u;
        uint32_t mE = __ballot_sync(FULLMASK, active);
#pragma unroll
        for (int b = 7; b >= 0; b--) {
            const uint32_t Bb = __ballot_sync(FULLMASK, (sym >> b) & 1u);
            mE &= ((sym >> b) & 1u) ? Bb : ~Bb;
        }
        const uint32_t eq = __popc(mE & before);
        const bool leader = active && eq == 0;
        if (leader) hraw[w][sym] = (uint8_t)__popc(mE);
        __syncthreads();
        if (tid < 256) { uint32_t run = 0; for (uint32_t q = 0; q < nw; q++) { const uint32_t h = hraw[q][tid]; hexcl[q][tid] = (uint16_t)run; run += h; } }
        __syncthreads();
        if (leader) hraw[w][sym] = 0;
        const uint32_t count_s = active ? cnt[sym] + hexcl[w][sym] + eq : 0u;
        const uint32_t b_tr = __ballot_sync(FULLMASK, active && count_s + 1 >= 255);          // ++o1[c] >= 255 halves the row (cr-ppm.c:91)
        if (lane == 0 && b_tr) atomicMin(&s_first, w * 32 + __ffs(b_tr) - 1);
        __syncthreads();
        const uint32_t first = s_first;
        const uint32_t applied = first == 0xFFFFFFFFu ? step : first + 1;
        const uint32_t b_ap = __ballot_sync(FULLMASK, active && tid < applied);
        if (leader && (mE & b_ap)) atomicAdd(&cnt[sym], (uint32_t)__popc(mE & b_ap));
        if (tid == 0) { O1Step r; r.pos = pos; r.applied = applied; steps[C.rec_base + nrec] = r; }
        if (tid < applied) { const uint32_t p = pos + O1S_RING + tid; if (p < r1) cr_cp_async4(&ring[slot], info_s + p); }
        cr_cp_async_commit();
        __syncthreads();
        if (tid == 0) s_first = 0xFFFFFFFFu;
        cr_cp_async_wait<1>();
        if (first != 0xFFFFFFFFu && tid < 256) cnt[tid] -= cnt[tid] / 2;                        // cr-ppm.c:92-94
        __syncthreads();
        pos += applied; nrec++;
    }
    cr_cp_async_wait<0>();
    if (tid < 256) row[tid] = (uint8_t)cnt[tid];
    if (tid == 0) ctx[c8].nsteps = nrec;
}
__global__ void __launch_bounds__(O1C_THREADS) k_o1_eval(const O1Ctx* __restrict__ ctx, const O1Step* __restrict__ steps, const uint8_t* __restrict__ snaps,
                                                         const uint32_t* __restrict__ info_s, const uint32_t* __restrict__ ord_s, const uint4* __restrict__ incl_s,
                                                         uint64_t* __restrict__ T2) {
    static_assert(O1C_THREADS == O1S_TH, "one evaluation CTA per recorded step");
    const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    // which row does this record belong to?  rec_base is ascending over ctx8 (rows that are not hot own no records)
    uint32_t lo = 0, hi = 256;
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (ctx[mid].rec_base <= blockIdx.x) lo = mid; else hi = mid; }
    // several rows may share a base (rows without records): the owner is the last one that starts here and has steps
    const O1Ctx C = ctx[lo];
    if (blockIdx.x - C.rec_base >= C.nsteps) return;
    const O1Step R = steps[blockIdx.x];
    __shared__ uint32_t cnt[256];
    __shared__ __align__(4) uint16_t hist[O1C_WARPS][256];
    __shared__ uint16_t basew[O1C_WARPS][256];
    __shared__ uint32_t smask[O1C_THREADS][8];
    if (tid < 256) cnt[tid] = snaps[(size_t)blockIdx.x * 256 + tid];
    for (uint32_t i = tid; i < O1C_WARPS * 128; i += O1C_THREADS) ((uint32_t*)&hist[0][0])[i] = 0;
    __syncthreads();
    const bool active = tid < R.applied;
    uint32_t sym = 0x1FF, ord = 0;
    uint32_t m[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (active) {
        const size_t x = R.pos + tid;
        sym = (info_s[x] >> 8) & 255; ord = ord_s[x];
        const uint4 a = incl_s[2 * x], b = incl_s[2 * x + 1];
        m[0] = a.x; m[1] = a.y; m[2] = a.z; m[3] = a.w; m[4] = b.x; m[5] = b.y; m[6] = b.z; m[7] = b.w;
        atomicAdd((uint32_t*)&hist[w][0] + (sym >> 1), (sym & 1u) ? 0x10000u : 1u);
    }
#pragma unroll
    for (int q = 0; q < 8; q++) smask[tid][q] = m[q];
    __syncthreads();
    if (tid < 256) {
        uint32_t run = 0;
        for (int q = 0; q < O1C_WARPS; q++) { uint32_t h = hist[q][tid]; hist[q][tid] = (uint16_t)run; basew[q][tid] = (uint16_t)(8 * (cnt[tid] + run) - 7); run += h; }
    }
    __syncthreads();
    uint32_t eq = 0, in_all = 0, in_lt = 0;
#pragma unroll 8
    for (uint32_t j = 0; j < 32; j++) {
        const uint32_t sj = __shfl_sync(FULLMASK, sym, j);
        if (j < lane && sj < 256) {
            const uint32_t in = (smask[tid][sj >> 5] >> (sj & 31)) & 1u;
            eq += sj == sym; in_all += in; in_lt += in & (uint32_t)(sj < sym);
        }
    }
    if (active) {
        uint32_t sum = 0, cum = 0;
#pragma unroll
        for (uint32_t q = 0; q < 8; q++) {
            const uint32_t mq = m[q];
            for (uint32_t b = 0; b < 32; b++) {
                const uint32_t x = q * 32 + b;
                const uint32_t v = (mq >> b & 1u) ? (uint32_t)basew[w][x] : 0u;
                sum += v; cum += x < sym ? v : 0u;
            }
        }
        sum += 8 * in_all; cum += 8 * in_lt;
        const uint32_t count_s = cnt[sym] + hist[w][sym] + eq;
        T2[ord] = ppm_pack(cum, 8 * count_s - 7, sum, 0);
    }
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

// ------------------------------------------------------------------ order-0 side models, epoch-parallel
// One CTA per model (0 = len_model, 1 = idx_model).  Between two halvings ("epoch") the table is
//   snapshot + 4 * (occurrences so far), and the halving instants depend only on how many symbols were coded,
// so every cumulative frequency of an epoch is a snapshot value plus a rank:
//   cum_i = snapcum[s_i] + 4 * #{j < i in epoch : s_j < s_i},  frq_i = snap[s_i] + 4 * #{j < i : s_j == s_i}.
// 1024 symbols are ranked per step: per-warp histograms, a prefix over warps, a prefix over symbols, and a
// 32-way comparison inside each warp.  Serial work left: one step per 1024 symbols plus one halving per epoch.
#define SE_THREADS 1024
__global__ void __launch_bounds__(SE_THREADS) k_side_epochs(const uint8_t* __restrict__ len_sym, const uint32_t* __restrict__ len_pos, uint32_t n_len,
                                                            const uint8_t* __restrict__ idx_sym, const uint32_t* __restrict__ idx_pos, uint32_t n_idx,
                                                            PpmState st, uint64_t* __restrict__ TS) {
    const uint32_t m = blockIdx.x;                       // model
    const uint8_t* sym = m == 0 ? len_sym : idx_sym;
    const uint32_t* pos = m == 0 ? len_pos : idx_pos;
    const uint32_t n = m == 0 ? n_len : n_idx;
    uint16_t* state = st.m0 + m * 256;
    __shared__ uint32_t cnt[256];                         // counts at the start of the current step
    __shared__ uint32_t cumt[256];                        // exclusive prefix of cnt
    __shared__ uint32_t s_total;
    __shared__ __align__(4) uint16_t hist[32][256];       // per-warp histogram of the step -> exclusive prefix over warps
    __shared__ uint16_t below[32][256];                   // per warp: exclusive prefix over symbols of its hist row
    __shared__ uint16_t steptot[256];                     // occurrences of each symbol in the step
    const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid < 256) cnt[tid] = state[tid];
    __syncthreads();
    uint32_t done = 0;
    for (;;) {
        // cumt / total from cnt: warp 0, 8 symbols per lane
        if (w == 0) {
            uint32_t v[8], sum = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) { v[k] = cnt[lane * 8 + k]; sum += v[k]; }
            uint32_t inc = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(FULLMASK, inc, d); if (lane >= d) inc += t; }
            uint32_t run = inc - sum;
#pragma unroll
            for (int k = 0; k < 8; k++) { cumt[lane * 8 + k] = run; run += v[k]; }
            if (lane == 31) s_total = inc;
        }
        for (uint32_t i = tid; i < 32 * 128; i += SE_THREADS) ((uint32_t*)&hist[0][0])[i] = 0;
        __syncthreads();
        if (done >= n) break;
        const uint32_t total = s_total;
        // symbols up to and including the one whose update pushes the total past 32000 (cr-model.c:65)
        const uint32_t to_rescale = (32000 - (total > 32000 ? 32000 : total)) / 4 + 1;
        uint32_t step = n - done; if (step > SE_THREADS) step = SE_THREADS; if (step > to_rescale) step = to_rescale;
        const bool active = tid < step;
        const uint32_t s = active ? sym[done + tid] : 0xFFFFu;
        if (active) atomicAdd((uint32_t*)&hist[w][0] + (s >> 1), (s & 1u) ? 0x10000u : 1u);
        __syncthreads();
        if (tid < 256) {                                   // exclusive prefix over warps, one symbol per thread
            uint32_t run = 0;
            for (int k = 0; k < 32; k++) { uint32_t h = hist[k][tid]; hist[k][tid] = (uint16_t)run; run += h; }
            steptot[tid] = (uint16_t)run;
        }
        __syncthreads();
        {                                                  // exclusive prefix over symbols, one row per warp
            uint32_t v[8], sum = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) { v[k] = hist[w][lane * 8 + k]; sum += v[k]; }
            uint32_t inc = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(FULLMASK, inc, d); if (lane >= d) inc += t; }
            uint32_t run = inc - sum;
#pragma unroll
            for (int k = 0; k < 8; k++) { below[w][lane * 8 + k] = (uint16_t)run; run += v[k]; }
        }
        __syncwarp();
        uint32_t lt = 0, eq = 0;                           // earlier lanes of this warp with a smaller / the same symbol
#pragma unroll 8
        for (uint32_t j = 0; j < 32; j++) { const uint32_t sj = __shfl_sync(FULLMASK, s, j); if (j < lane) { lt += sj < s; eq += sj == s; } }
        if (active) {
            const uint32_t cum = cumt[s] + 4 * ((uint32_t)below[w][s] + lt);
            const uint32_t frq = cnt[s] + 4 * ((uint32_t)hist[w][s] + eq);
            TS[pos[done + tid]] = ppm_pack(cum, frq, total + 4 * tid, 0);
        }
        __syncthreads();
        if (tid < 256) {
            uint32_t c = cnt[tid] + 4u * steptot[tid];
            if (step == to_rescale) c = (c + 1) >> 1;      // halve, rounding up (cr-model.c:66-73)
            cnt[tid] = c;
        }
        done += step;
        __syncthreads();
    }
    if (tid < 256) state[tid] = (uint16_t)cnt[tid];
}

// The same with one CTA per job: any number of models, each with its own increment (comprox's LZ77 front-end uses
// 30 for lengths, 1 for short distances and 1, 4, 16, 64, 256, 1024 for the skewed position models; cr_lz77.cuh).
__global__ void __launch_bounds__(SE_THREADS) k_side_epochs_jobs(SideJobs J, uint64_t* __restrict__ TS) {
    const uint32_t m = blockIdx.x;                       // model
    const uint8_t* sym = J.sym[m];
    const uint32_t* pos = J.pos[m];
    const uint32_t n = J.n[m];
    const uint32_t inc_m = J.inc[m];
    uint16_t* state = J.state[m];
    __shared__ uint32_t cnt[256];                         // counts at the start of the current step
    __shared__ uint32_t cumt[256];                        // exclusive prefix of cnt
    __shared__ uint32_t s_total;
    __shared__ __align__(4) uint16_t hist[32][256];       // per-warp histogram of the step -> exclusive prefix over warps
    __shared__ uint16_t below[32][256];                   // per warp: exclusive prefix over symbols of its hist row
    __shared__ uint16_t steptot[256];                     // occurrences of each symbol in the step
    const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid < 256) cnt[tid] = state[tid];
    __syncthreads();
    uint32_t done = 0;
    for (;;) {
        // cumt / total from cnt: warp 0, 8 symbols per lane
        if (w == 0) {
            uint32_t v[8], sum = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) { v[k] = cnt[lane * 8 + k]; sum += v[k]; }
            uint32_t inc = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(FULLMASK, inc, d); if (lane >= d) inc += t; }
            uint32_t run = inc - sum;
#pragma unroll
            for (int k = 0; k < 8; k++) { cumt[lane * 8 + k] = run; run += v[k]; }
            if (lane == 31) s_total = inc;
        }
        for (uint32_t i = tid; i < 32 * 128; i += SE_THREADS) ((uint32_t*)&hist[0][0])[i] = 0;
        __syncthreads();
        if (done >= n) break;
        const uint32_t total = s_total;
        // symbols up to and including the one whose update pushes the total past 32000 (cr-model.c:65)
        const uint32_t to_rescale = (32000 - (total > 32000 ? 32000 : total)) / inc_m + 1;
        uint32_t step = n - done; if (step > SE_THREADS) step = SE_THREADS; if (step > to_rescale) step = to_rescale;
        const bool active = tid < step;
        const uint32_t s = active ? sym[done + tid] : 0xFFFFu;
        if (active) atomicAdd((uint32_t*)&hist[w][0] + (s >> 1), (s & 1u) ? 0x10000u : 1u);
        __syncthreads();
        if (tid < 256) {                                   // exclusive prefix over warps, one symbol per thread
            uint32_t run = 0;
            for (int k = 0; k < 32; k++) { uint32_t h = hist[k][tid]; hist[k][tid] = (uint16_t)run; run += h; }
            steptot[tid] = (uint16_t)run;
        }
        __syncthreads();
        {                                                  // exclusive prefix over symbols, one row per warp
            uint32_t v[8], sum = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) { v[k] = hist[w][lane * 8 + k]; sum += v[k]; }
            uint32_t inc = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(FULLMASK, inc, d); if (lane >= d) inc += t; }
            uint32_t run = inc - sum;
#pragma unroll
            for (int k = 0; k < 8; k++) { below[w][lane * 8 + k] = (uint16_t)run; run += v[k]; }
        }
        __syncwarp();
        uint32_t lt = 0, eq = 0;                           // earlier lanes of this warp with a smaller / the same symbol
#pragma unroll 8
        for (uint32_t j = 0; j < 32; j++) { const uint32_t sj = __shfl_sync(FULLMASK, s, j); if (j < lane) { lt += sj < s; eq += sj == s; } }
        if (active) {
            const uint32_t cum = cumt[s] + inc_m * ((uint32_t)below[w][s] + lt);
            const uint32_t frq = cnt[s] + inc_m * ((uint32_t)hist[w][s] + eq);
            TS[pos[done + tid]] = ppm_pack(cum, frq, total + inc_m * tid, 0);
        }
        __syncthreads();
        if (tid < 256) {
            uint32_t c = cnt[tid] + inc_m * steptot[tid];
            if (step == to_rescale) c = (c + 1) >> 1;      // halve, rounding up (cr-model.c:66-73)
            cnt[tid] = c;
        }
        done += step;
        __syncthreads();
    }
    if (tid < 256) state[tid] = (uint16_t)cnt[tid];
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
    uint32_t aborted = 0, abort_tri = 0;
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
                if ((t.y & TRI_TOKEND) && c.n >= S.limit) { aborted = 1; abort_tri = (uint32_t)(base + j); break; }
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
        res[s].abort_tri = abort_tri; res[s].pad = 0;
    }
}

// ------------------------------------------------------------------ range coder, split form
// The coder's two recurrences are independent:  range_{n+1} depends only on (range_n, sum_n, frq_n), while `low`
// merely accumulates cum_n * (range_n / sum_n) at a byte offset given by the number of renormalisation shifts so
// far.  So the only truly serial part is the range chain (k_range_chain: ~8 dependent integer ops per symbol);
// the output bytes are the big-number sum  sum_n a_n * 256^-(B_n+4)  which k_low_scatter / k_low_carry evaluate
// for all symbols and all output bytes in parallel (carry resolution by look-right over 0xFF runs).
// chain input: {frq, M_lo, M_hi, 0} with M = floor(2^63 / sum) + 1, for which  (n * M) >> 63 == n / sum  exactly
// for every n < 2^32 and sum < 2^31 (error term n/2^63 < 1/sum).
__global__ void k_chain_inputs(const Tri* __restrict__ dense, uint64_t n, uint4* __restrict__ cin) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Tri t = dense[i];
    const unsigned long long M = (1ull << 63) / t.sum + 1ull;
    cin[i] = make_uint4(t.frq & 0x7FFFFFFFu, (uint32_t)M, (uint32_t)(M >> 32), t.sum);
}

// chain input of VARIANT 7 (double-precision chain, cr_rc.cuh): {inv, frq} as two doubles
__global__ void k_chain_inputs_dp(const Tri* __restrict__ dense, uint64_t n, uint4* __restrict__ cin) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Tri t = dense[i];
    cin[i] = rc_dp_record(t.frq & 0x7FFFFFFFu, t.sum);
}

// VARIANT 1: 32-bit reciprocal + correction, both candidates normalised speculatively (input: Tri)
// VARIANT 2: exact 64-bit reciprocal, FLO-based normalisation                          (input: k_chain_inputs)
// VARIANT 3: exact 64-bit reciprocal, one compare/select for the common 0/1-byte shift, loop for the rest
// VARIANT 4: as 2, but the range is never normalised explicitly: with r = q*frq (un-normalised) and k = leading zero
//            bytes of r, the next quotient is ((r << 8k) * M) >> 63 == (r * M) >> (63 - 8k), so the multiply runs
//            beside the leading-zero count instead of behind it (the count was ~1/4 of the dependent chain).
// VARIANTS 5, 6: one symbol of the chain without a leading-zero count (FLO is a ~25-cycle variable-latency op and sat on the
// chain).  range is the un-normalised product q * frq of the previous symbol; amt = 31 - 8 * (its leading zero bytes) comes from
// three independent compares that run beside the two multiplies.
//   5: amt = (range < 2^16) ? (range < 2^8 ? 7 : 15) : (range < 2^24 ? 23 : 31), then one shift
//   6: only the common case (0 or 1 leading zero bytes: one compare + one select) feeds the shift on the chain; 2 or 3 leading
//      zero bytes (~1 % of the symbols) override the quotient with a predicated second shift.
template <int VARIANT>
CR_D void rc_step_cmp(uint32_t& range, const uint4 t, uint32_t& q, uint32_t& amt) {
    const uint32_t ahi = __umulhi(range, t.y);                                   // t = {frq, M_lo, M_hi, sum}, M = floor(2^63 / sum) + 1
    const unsigned long long S2 = (unsigned long long)range * t.z + ahi;         // (range * M) >> 32
    if (VARIANT == 5) {
        const uint32_t a01 = range < (1u << 24) ? 23u : 31u, a23 = range < (1u << 8) ? 7u : 15u;
        asm("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %1, 65536;\n\tselp.u32 %0, %2, %3, p;\n\t}" : "=r"(amt) : "r"(range), "r"(a23), "r"(a01));
        q = (uint32_t)(S2 >> amt);
    } else {
        const uint32_t lo = (uint32_t)S2, hi = (uint32_t)(S2 >> 32);
        asm("{\n\t.reg .pred p24, p16, p8;\n\t.reg .u32 b;\n\t"
            "setp.lt.u32 p24, %2, 16777216;\n\tsetp.lt.u32 p16, %2, 65536;\n\tsetp.lt.u32 p8, %2, 256;\n\t"
            "selp.u32 %1, 23, 31, p24;\n\tselp.u32 b, 7, 15, p8;\n\t"
            "shf.r.clamp.b32 %0, %3, %4, %1;\n\t"
            "@p16 shf.r.clamp.b32 %0, %3, %4, b;\n\t"
            "@p16 mov.u32 %1, b;\n\t}"
            : "=&r"(q), "=&r"(amt) : "r"(range), "r"(lo), "r"(hi));
    }
    range = q * t.x;                                                             // range *= frq, left un-normalised (cr-rangecoder.c:64)
}

template <int VARIANT>
__global__ void __launch_bounds__(128) k_range_chain(const uint4* __restrict__ in_main, const uint4* __restrict__ in_side, const uint32_t* __restrict__ escord,
                                                      const RcStream* __restrict__ streams, uint32_t nstreams,
                                                      uint32_t* __restrict__ q_main, uint32_t* __restrict__ sh_main, uint32_t* __restrict__ q_side, uint32_t* __restrict__ sh_side) {
    __shared__ uint4 stage[4][RC_BATCH + 8];          // slack: the look-ahead reads past the last symbol need no clamp
    __shared__ uint32_t oq[4][RC_BATCH], os[4][RC_BATCH + 1];   // VARIANT 4: os[j] = top-bit index BEFORE symbol j (stored late, off the chain)
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= nstreams) return;
    const RcStream S = streams[s];
    const uint4* tri = S.is_main ? in_main : in_side;
    uint32_t* qo = S.is_main ? q_main : q_side;
    uint32_t* so = S.is_main ? sh_main : sh_side;
    size_t i0 = S.ev_begin, i1 = S.ev_end;
    if (S.is_main) { i0 += escord[S.ev_begin]; i1 += escord[S.ev_end]; }
    uint32_t range = 0xFFFFFFFFu;
    double R = RC_DP_R0;                                 // VARIANT 7: twice the normalised range as a double (cr_rc.cuh)
    uint32_t msb = 31;                                   // VARIANT 4: index of the top set bit of the un-normalised range
    uint32_t c24 = 24, c7 = 7;
    asm volatile("" : "+r"(c24), "+r"(c7));            // keep the LOP3 operands in registers
    uint4 r0 = make_uint4(1, 1, 1, 1), r1 = r0;
    if (i0 + lane < i1) r0 = tri[i0 + lane];
    if (i0 + 32 + lane < i1) r1 = tri[i0 + 32 + lane];
    for (size_t base = i0; base < i1; base += RC_BATCH) {
        stage[w][lane] = r0; stage[w][lane + 32] = r1;
        __syncwarp();
        const size_t nb = base + RC_BATCH;
        if (nb + lane < i1) r0 = tri[nb + lane];
        if (nb + 32 + lane < i1) r1 = tri[nb + 32 + lane];
        const uint32_t cnt = i1 - base < RC_BATCH ? (uint32_t)(i1 - base) : RC_BATCH;
        if (lane == 0) {
            uint4 t = stage[w][0];
            if (VARIANT == 7) {
                // VARIANT 7: two dependent DFMA + three ALU operations per symbol (rc_dp_step, cr_rc.cuh); q and the top-bit index leave
                // the chain through shared memory like VARIANT 4's (os[j + 1] = top bit after symbol j)
                for (uint32_t j = 0; j < cnt; j++) {
                    const uint4 tn = stage[w][j + 1];
                    uint32_t q, m;
                    rc_dp_step(R, t, q, m);
                    oq[w][j] = q; os[w][j + 1] = m;
                    t = tn;
                }
            }
            for (uint32_t j = 0; j < cnt && VARIANT < 5; j++) {
                const uint4 tn = stage[w][j + 1];
                uint32_t q, sh;
                if (VARIANT == 1) {
                    // Tri: x = cum, y = frq|flag, z = sum, w = floor(2^32/sum): umulhi gives q or q-1 (cr-rangecoder.c:61)
                    const uint32_t q0 = __umulhi(range, t.w);
                    const uint32_t frq = t.y & 0x7FFFFFFFu;
                    const bool up = range - q0 * t.z >= t.z;
                    const uint32_t ra = q0 * frq, rb = ra + frq;                  // range *= frq (:64)
                    const uint32_t sa = __clz(ra) >> 3, sb = __clz(rb) >> 3;      // while (range < 2^24) range <<= 8 (:65-68)
                    const uint32_t na = ra << (8 * sa), nb2 = rb << (8 * sb);
                    q = up ? q0 + 1 : q0; sh = up ? sb : sa; range = up ? nb2 : na;
                } else if (VARIANT == 4) {
                    const uint32_t ahi = __umulhi(range, t.y);                              // these two do not need msb: they run
                    const unsigned long long S2 = (unsigned long long)range * t.z + ahi;      // while the FLO below is in flight
                    uint32_t amt;                                                           // 31 - 8 * leading zero bytes = (msb & 24) | 7
                    asm volatile("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(amt) : "r"(msb), "r"(c24), "r"(c7));
                    q = (uint32_t)(S2 >> amt);
                    os[w][j] = msb;                                                         // top bit before this symbol = after the previous one
                    oq[w][j] = q;
                    range = q * t.x;                                                        // range *= frq, left un-normalised
                    asm volatile("bfind.u32 %0, %1;" : "=r"(msb) : "r"(range));
                    t = tn;
                    continue;
                } else {
                    // x = frq, (y, z) = M: q = (range * M) >> 63, exact
                    const uint32_t ahi = __umulhi(range, t.y);
                    const unsigned long long S2 = (unsigned long long)range * t.z + ahi;
                    q = (uint32_t)(S2 >> 31);
                    uint32_t r = q * t.x;
                    if (VARIANT == 2) { sh = __clz(r) >> 3; range = r << (8 * sh); }
                    else {
                        sh = r < (1u << 24);
                        r = sh ? r << 8 : r;
                        while (r < (1u << 24)) { r <<= 8; sh++; }
                        range = r;
                    }
                }
                oq[w][j] = q; os[w][j] = sh;
                t = tn;
            }
            if (VARIANT == 5 || VARIANT == 6) {
                // a symbol takes less than one LDS latency here: full batches keep 8 symbols in registers and refill each slot
                // 8 symbols ahead of its use
                if (cnt == RC_BATCH) {
                    uint4 p[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) p[u] = stage[w][u];
#pragma unroll 1
                    for (uint32_t j0 = 0; j0 < RC_BATCH; j0 += 8) {
#pragma unroll
                        for (int u = 0; u < 8; u++) {
                            const uint4 t8 = p[u];
                            p[u] = stage[w][j0 + u + 8];
                            uint32_t q, amt;
                            rc_step_cmp<VARIANT>(range, t8, q, amt);
                            os[w][j0 + u] = amt; oq[w][j0 + u] = q;
                        }
                    }
                } else {
                    for (uint32_t j = 0; j < cnt; j++) {
                        uint32_t q, amt;
                        rc_step_cmp<VARIANT>(range, stage[w][j], q, amt);
                        os[w][j] = amt; oq[w][j] = q;
                    }
                }
            }
            if (VARIANT == 4) os[w][cnt] = msb;
            if (VARIANT == 5 || VARIANT == 6) os[w][cnt] = 31u - (__clz(range) & 24u);
        }
        __syncwarp();
        const uint32_t so_off = VARIANT >= 4 ? 1u : 0u;                    // VARIANT 4 keeps the value after symbol j in slot j+1
        if (lane < cnt) { qo[base + lane] = oq[w][lane]; so[base + lane] = os[w][lane + so_off]; }
        if (lane + 32 < cnt) { qo[base + lane + 32] = oq[w][lane + 32]; so[base + lane + 32] = os[w][lane + 32 + so_off]; }
        __syncwarp();
    }
}

// VARIANTS 4, 5 store the top-bit index of the un-normalised range (5: rounded up to 8k+7); shift count = 3 - (msb >> 3)
__global__ void k_msb_to_shifts(uint32_t* __restrict__ sh, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sh[i] = 3u - (sh[i] >> 3);
}

struct LowStream {          // one stream in split form
    uint64_t tri_begin, tri_end;   // dense triple range
    uint64_t dsum_off;             // offset (in bytes == in uint32 digit slots) of the stream's output
    uint32_t length;               // output bytes = shifts + 5
    uint32_t is_main;
};
CR_D uint32_t low_find_stream(const LowStream* __restrict__ ls, uint32_t n, uint64_t i) {
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (ls[mid].tri_begin <= i) lo = mid; else hi = mid; }
    return lo;
}
// digit sums: symbol n adds the 4 bytes of a_n = cum_n * q_n at output positions B_n+1 .. B_n+4 (cr-rangecoder.c:62-63)
__global__ void k_low_scatter(const Tri* __restrict__ dense, const uint32_t* __restrict__ q, const uint32_t* __restrict__ bscan,
                              const LowStream* __restrict__ ls, uint32_t nls, uint64_t ntri, uint32_t* __restrict__ dsum) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ntri) return;
    const LowStream S = ls[low_find_stream(ls, nls, i)];
    if (i < S.tri_begin || i >= S.tri_end) return;               // triple of a stream that runs the serial fallback
    const uint32_t a = dense[i].cum * q[i];
    if (a == 0) return;
    uint32_t* d = dsum + S.dsum_off + (bscan[i] - bscan[S.tri_begin]) + 1;
    if (a >> 24) atomicAdd(d, a >> 24);
    if ((a >> 16) & 255) atomicAdd(d + 1, (a >> 16) & 255);
    if ((a >> 8) & 255) atomicAdd(d + 2, (a >> 8) & 255);
    if (a & 255) atomicAdd(d + 3, a & 255);
}
// value (0..258) of output position p after the digit sums have been spread to single bytes, before carries
CR_D uint32_t low_digit(const uint32_t* __restrict__ d, uint32_t len, uint32_t p) {
    auto D = [&](uint32_t x) { return x < len ? d[x] : 0u; };
    auto E = [&](uint32_t x) { return (D(x) & 255) + ((D(x + 1) >> 8) & 255) + ((D(x + 2) >> 16) & 255) + (D(x + 3) >> 24); };
    return (E(p) & 255) + (E(p + 1) >> 8);
}
__global__ void k_low_carry(const LowStream* __restrict__ ls, const uint32_t* __restrict__ dsum, const RcStream* __restrict__ streams_by_ls,
                            uint8_t* __restrict__ outbuf) {
    const LowStream S = ls[blockIdx.y];
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= S.length) return;
    const uint32_t* d = dsum + S.dsum_off;
    uint32_t v = low_digit(d, S.length, p);
    uint32_t r = p + 1, carry = 0;
    for (; r < S.length; r++) { uint32_t f = low_digit(d, S.length, r); if (f != 255) { carry = f >> 8; break; } }
    outbuf[streams_by_ls[blockIdx.y].out_off + p] = (uint8_t)(v + carry);
}

struct StreamTotals { uint64_t tri_begin, tri_end; uint32_t shifts; uint32_t pad; };
__global__ void k_stream_totals(const RcStream* __restrict__ streams, uint32_t nstreams, const uint32_t* __restrict__ escord,
                                const uint32_t* __restrict__ bscan_main, const uint32_t* __restrict__ bscan_side, StreamTotals* __restrict__ out) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nstreams) return;
    const RcStream S = streams[s];
    uint64_t i0 = S.ev_begin, i1 = S.ev_end;
    if (S.is_main) { i0 += escord[S.ev_begin]; i1 += escord[S.ev_end]; }
    const uint32_t* b = S.is_main ? bscan_main : bscan_side;
    out[s].tri_begin = i0; out[s].tri_end = i1; out[s].shifts = b[i1] - b[i0]; out[s].pad = 0;
}

// ------------------------------------------------------------------ warp-cooperative decoder (one warp per container)
// Same serial chain as k_lzdecode_serial (cr_decode.cuh) -- decoding cannot be replayed per context -- but every
// 256-wide loop of ppm_decode / model_get_decode_symbol runs across the lanes: lane l owns symbols 8l..8l+7 of the
// current o2 row / o1 row / order-0 model, cumulative frequencies come from a warp scan, the symbol search is a
// ballot.  The range decoder state is replicated in all lanes (uniform control flow, no divergence).
// (bounds and error state as in RcDec, cr_decode.cuh: untrusted input never reads behind `end`, divides by zero or spins)
struct WRc {
    uint32_t range, code, err; const uint8_t* in; const uint8_t* end;
    CR_D uint32_t next() { if (in < end) return __ldg(in++); err = DEC_ERR_STREAM; return 0; }
    CR_D void init(const uint8_t* p, const uint8_t* e) { range = 0xFFFFFFFFu; code = 0; err = 0; in = p; end = e; if (p > e) { err = DEC_ERR_STREAM; in = e; } for (int i = 0; i < 5; i++) code = (code << 8) + next(); }
    CR_D uint32_t target(uint32_t sum) {
        if (sum == 0 || range < sum) { err = DEC_ERR_STREAM; range = 1u << 24; return 0; }
        range /= sum; return code / range;
    }
    CR_D void consume(uint32_t cum, uint32_t frq) {
        if (frq == 0) { err = DEC_ERR_STREAM; frq = 1; }
        code -= cum * range; range *= frq;
        if (range == 0) { err = DEC_ERR_STREAM; range = 1u << 24; }
        while (range < (1u << 24)) { code = (code << 8) + next(); range <<= 8; }
    }
};
CR_D uint32_t wscan_incl(uint32_t v, uint32_t lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(FULLMASK, v, d); if (lane >= (uint32_t)d) v += t; }
    return v;
}
CR_D uint32_t byte_of(uint32_t lo, uint32_t hi, uint32_t k) { return ((k < 4 ? lo : hi) >> (8 * (k & 3))) & 255u; }

// ppm_decode (cr-ppm.c:169-235), warp form.  Returns the decoded symbol (uniform).
CR_D uint32_t wdec_ppm(PpmState& st, uint32_t ctx, WRc& rc, uint32_t lane) {
    uint8_t* row = st.o2 + (size_t)(ctx & 0xffff) * PPM_O2_STRIDE;
    const uint32_t slot = ppm_slot(ctx);
    uint2 fv = ((const uint2*)row)[lane];
    uint32_t f0 = fv.x, f1 = fv.y;
    const uint32_t fl = *(const uint16_t*)(row + 256);
    uint32_t f256 = fl & 255, f257 = fl >> 8;
    const uint32_t pred = st.o3_byte[slot];
    uint32_t conf = st.o3_conf[slot];
    const uint32_t pown = pred >> 3;
    const uint32_t pf = __shfl_sync(FULLMASK, byte_of(f0, f1, pred & 7), pown);
    const uint32_t lsum = wsum4(f0) + wsum4(f1);
    uint32_t body = __reduce_add_sync(FULLMASK, lsum);
    const uint32_t tgt = rc.target(body + f256 + f257 - pf);
    const uint32_t ls = lsum - (lane == pown ? pf : 0u);
    const uint32_t incl = wscan_incl(ls, lane);
    const uint32_t hitmask = __ballot_sync(FULLMASK, incl > tgt);
    uint32_t s, acc, frq;
    if (hitmask) {
        const uint32_t L = __ffs(hitmask) - 1;
        acc = __shfl_sync(FULLMASK, incl - ls, L);
        const uint32_t g0 = __shfl_sync(FULLMASK, f0, L), g1 = __shfl_sync(FULLMASK, f1, L);
        uint32_t k = 0; frq = 0;
        for (; k < 8; k++) { const uint32_t sy = L * 8 + k; frq = byte_of(g0, g1, k); const uint32_t wgt = sy == pred ? 0u : frq; if (acc + wgt > tgt) break; acc += wgt; }
        s = L * 8 + k;
    } else {
        acc = body - pf;
        if (tgt < acc + f256) { s = 256; frq = f256; } else { acc += f256; s = 257; frq = f257; }
    }
    rc.consume(acc, frq);
    // ---- o2_model_update(s, +1) and its consequences
    bool rescale = false, wrote_all = false;
    uint32_t sym = s;
    if (s == 256) { f256 = (f256 + 1) & 255; rescale = f256 > 250; sym = pred; }
    else if (s < 256) {
        if (lane == (s >> 3)) { const uint32_t k = s & 7; if (k < 4) f0 += 1u << (8 * k); else f1 += 1u << (8 * (k - 4)); }
        body += 1;
        if (frq + 1 > 250) rescale = true;
        else if (frq + 1 == 2) { f257 = (f257 - 1) & 255; rescale = f257 > 250; }
    } else {
        f257 = (f257 + 1) & 255;
        const bool resc257 = f257 > 250;
        if (resc257) {
            f0 = (f0 >> 1) & 0x7f7f7f7fu; f1 = (f1 >> 1) & 0x7f7f7f7fu;
            const uint32_t ones = (__popc(__vcmpeq4(f0, 0x01010101u)) + __popc(__vcmpeq4(f1, 0x01010101u))) >> 3;
            f257 = (1 + __reduce_add_sync(FULLMASK, ones)) & 255; f256 = (f256 + 1) >> 1;
            wrote_all = true;
        }
        // o1 with exclusion (cr-ppm.c:208-227)
        uint8_t* o1row = st.o1 + (ctx & 0xff) * 256;
        uint2 av = ((const uint2*)o1row)[lane];
        uint32_t a0 = av.x, a1 = av.y;
        const uint32_t z0 = __vcmpeq4(f0, 0u), z1 = __vcmpeq4(f1, 0u);
        uint32_t bits = ((z0 & 1u) | (z0 >> 7 & 2u) | (z0 >> 14 & 4u) | (z0 >> 21 & 8u)) | (((z1 & 1u) | (z1 >> 7 & 2u) | (z1 >> 14 & 4u) | (z1 >> 21 & 8u)) << 4);
        if (lane == pown) bits &= ~(1u << (pred & 7));
        uint32_t ls1 = 0;
#pragma unroll
        for (uint32_t k = 0; k < 8; k++) if (bits >> k & 1u) ls1 += byte_of(a0, a1, k) * 8 - 7;
        const uint32_t sum1 = __reduce_add_sync(FULLMASK, ls1);
        const uint32_t t1 = rc.target(sum1);
        const uint32_t incl1 = wscan_incl(ls1, lane);
        const uint32_t hm = __ballot_sync(FULLMASK, incl1 > t1);
        const uint32_t L = hm ? __ffs(hm) - 1 : 31;
        uint32_t cum1 = __shfl_sync(FULLMASK, incl1 - ls1, L);
        const uint32_t g0 = __shfl_sync(FULLMASK, a0, L), g1 = __shfl_sync(FULLMASK, a1, L), gb = __shfl_sync(FULLMASK, bits, L);
        uint32_t k = 0, fr = 1;
        for (; k < 8; k++) { if (!(gb >> k & 1u)) continue; fr = byte_of(g0, g1, k) * 8 - 7; if (cum1 + fr > t1) break; cum1 += fr; }
        if (k == 8) k = 7;
        const uint32_t d = L * 8 + k;
        rc.consume(cum1, fr);
        const uint32_t cd = (fr + 7) >> 3;                                   // o1[d] before the update
        if (lane == L) { if (k < 4) a0 += 1u << (8 * k); else a1 += 1u << (8 * (k - 4)); }
        if (cd + 1 >= 255) { a0 -= (a0 >> 1) & 0x7f7f7f7fu; a1 -= (a1 >> 1) & 0x7f7f7f7fu; ((uint2*)o1row)[lane] = make_uint2(a0, a1); }
        else if (lane == L) ((uint2*)o1row)[lane] = make_uint2(a0, a1);
        if (!resc257) { if (lane == (d >> 3)) { const uint32_t q = d & 7; if (q < 4) f0 += 1u << (8 * q); else f1 += 1u << (8 * (q - 4)); } }
        sym = d;
    }
    if (rescale) {
        f0 = (f0 >> 1) & 0x7f7f7f7fu; f1 = (f1 >> 1) & 0x7f7f7f7fu;
        const uint32_t ones = (__popc(__vcmpeq4(f0, 0x01010101u)) + __popc(__vcmpeq4(f1, 0x01010101u))) >> 3;
        f257 = (1 + __reduce_add_sync(FULLMASK, ones)) & 255; f256 = (f256 + 1) >> 1;
        wrote_all = true;
    }
    if (wrote_all || lane == (sym >> 3)) ((uint2*)row)[lane] = make_uint2(f0, f1);
    if (lane == 0) *(uint16_t*)(row + 256) = (uint16_t)(f256 | f257 << 8);
    // ---- ppm_update_o3 (cr-ppm.c:69-88)
    if (s == 256) conf += conf < 15;
    else { conf = (conf > 1) + (conf > 2) + (conf > 4) + (conf > 8); if (conf == 0) { if (lane == 0) st.o3_byte[slot] = (uint8_t)sym; conf = 1; } }
    if (lane == 0) st.o3_conf[slot] = (uint8_t)conf;
    __syncwarp();
    return sym;
}
// order-0 model in registers (8 x u16 per lane): M_my_dec_ with increment 4
CR_D uint32_t wdec_m0(uint32_t (&f)[4], uint32_t& total, WRc& rc, uint32_t lane) {
    const uint32_t tgt = rc.target(total);
    const uint32_t ls = side_part(f, 8);
    const uint32_t incl = wscan_incl(ls, lane);
    const uint32_t hm = __ballot_sync(FULLMASK, incl > tgt);
    const uint32_t L = hm ? __ffs(hm) - 1 : 31;
    uint32_t acc = __shfl_sync(FULLMASK, incl - ls, L);
    uint32_t g[4];
#pragma unroll
    for (int q = 0; q < 4; q++) g[q] = __shfl_sync(FULLMASK, f[q], L);
    uint32_t k = 0, fr = 0;
    for (; k < 8; k++) { fr = (g[k >> 1] >> (16 * (k & 1))) & 0xffff; if (acc + fr > tgt) break; acc += fr; }
    if (k == 8) k = 7;
    rc.consume(acc, fr);
    if (lane == L) f[k >> 1] += 4u << (16 * (k & 1));
    total += 4;
    if (total > 32000) { for (int q = 0; q < 4; q++) f[q] = ((f[q] + 0x00010001u) >> 1) & 0x7fff7fffu; total = __reduce_add_sync(FULLMASK, side_part(f, 8)); }
    return L * 8 + k;
}
// order-0 model in memory (8 x u16 per lane, L1 resident): M_my_dec_ with an arbitrary increment (the LZ77 front-end has eight models)
CR_D uint32_t wdec_m0g(uint16_t* __restrict__ fg, WRc& rc, uint32_t inc, uint32_t lane) {
    uint4 v = ((const uint4*)fg)[lane];
    uint32_t f[4] = { v.x, v.y, v.z, v.w };
    const uint32_t ls = side_part(f, 8);
    const uint32_t incl = wscan_incl(ls, lane);
    const uint32_t total = __shfl_sync(FULLMASK, incl, 31);
    const uint32_t tgt = rc.target(total);
    const uint32_t hm = __ballot_sync(FULLMASK, incl > tgt);
    const uint32_t L = hm ? __ffs(hm) - 1 : 31;
    uint32_t acc = __shfl_sync(FULLMASK, incl - ls, L);
    uint32_t g[4];
#pragma unroll
    for (int q = 0; q < 4; q++) g[q] = __shfl_sync(FULLMASK, f[q], L);
    uint32_t k = 0, fr = 0;
    for (; k < 8; k++) { fr = (g[k >> 1] >> (16 * (k & 1))) & 0xffff; if (acc + fr > tgt) break; acc += fr; }
    if (k == 8) k = 7;
    rc.consume(acc, fr);
    if (lane == L) f[k >> 1] += inc << (16 * (k & 1));
    const bool halve = total + inc > 32000;
    if (halve) for (int q = 0; q < 4; q++) f[q] = ((f[q] + 0x00010001u) >> 1) & 0x7fff7fffu;
    if (halve || lane == L) ((uint4*)fg)[lane] = make_uint4(f[0], f[1], f[2], f[3]);
    __syncwarp();
    return L * 8 + k;
}
// copy `len` bytes from out[q..] to out[n..] with the byte-serial semantics of the reference's copy loop
CR_D void wcopy_match(uint8_t* out, uint32_t n, uint32_t q, uint32_t len, uint32_t lane) {
    const uint32_t dist = n - q;
    if (dist >= 32) { for (uint32_t b = 0; b < len; b += 32) { if (b + lane < len) out[n + b + lane] = out[q + b + lane]; __syncwarp(); } }
    else { for (uint32_t b = 0; b < len; b += 32) if (b + lane < len) out[n + b + lane] = out[q + (b + lane) % dist]; __syncwarp(); }
}

__device__ __forceinline__ void lzdecode_warp_body(int variant, const uint8_t* __restrict__ cont, const DecBlock* __restrict__ blocks, uint32_t nb,
                                                   PpmState st, DecTables T, uint32_t* __restrict__ ctx_io, uint8_t* __restrict__ D) {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t ctx = *ctx_io, err = 0;
    uint32_t fa[4], fb[4];
    { const uint4 v = ((const uint4*)(st.m0))[lane]; fa[0] = v.x; fa[1] = v.y; fa[2] = v.z; fa[3] = v.w;
      const uint4 u = ((const uint4*)(st.m0 + 256))[lane]; fb[0] = u.x; fb[1] = u.y; fb[2] = u.z; fb[3] = u.w; }
    uint32_t tota = __reduce_add_sync(FULLMASK, side_part(fa, 8)), totb = __reduce_add_sync(FULLMASK, side_part(fb, 8));
    for (uint32_t b = 0; b < nb && !err; b++) {
        const DecBlock B = blocks[b];
        if (!B.coded) continue;
        const uint8_t* in = cont + B.in_off;
        const uint8_t* in_end = in + B.in_size;
        uint8_t* out = D + B.d_off;
        const uint32_t orig = B.d_size;
        if (orig == 0) continue;
        if (variant == 0) {
            const uint32_t esc = in[2], off_idx = cr_ld32(in + 12);
            const int ctx4 = orig >= 4194304;
            WRc rc, side; rc.init(in + 16, in_end); side.init(in + (off_idx <= B.in_size ? off_idx : B.in_size), in_end);
            for (uint32_t i = lane; i < 256 * 16; i += 32) T.rz_short[i] = 0;
            __syncwarp();
            uint32_t bucket = 0, sbucket = 0, n = 0, hist = 0;                 // hist: last four bytes written
            if (lane == 0) out[0] = in[0];
            hist = in[0]; n = 1;
            __syncwarp();
            // The order-1 lists (16 most recent positions per previous byte) are rings here: lane l keeps the 4-bit heads of
            // bytes 8l..8l+7 in one register, an update is one store instead of a 16-entry shift.  The bucket word the next
            // update needs is loaded by lane 0 as soon as the bucket is known (one symbol ahead of its use).
            uint32_t sheads = 0;
            uint32_t m_pref = lane == 0 ? T.rz_meta[0] : 0u;
            while (n < orig) {
                uint32_t len = 1;
                const uint32_t s = wdec_ppm(st, ctx, rc, lane);
                if (s == esc) {
                    const uint32_t l = wdec_m0(fa, tota, side, lane);
                    if (l == 0) { if (lane == 0) out[n] = (uint8_t)esc; n++; }
                    else {
                        const uint32_t idx = wdec_m0(fb, totb, side, lane);
                        uint32_t q;
                        if (idx < 64) { const uint32_t m = T.rz_meta[bucket]; q = T.rz_items[(size_t)bucket * 64 + (((m & 255) + 64 - idx) & 63)]; }
                        else {
                            const uint32_t hd = (__shfl_sync(FULLMASK, sheads, sbucket >> 3) >> (4 * (sbucket & 7))) & 15u;
                            q = T.rz_short[sbucket * 16 + ((hd + 64 + 16 - idx) & 15)];              // idx - 64 steps back from the newest
                        }
                        len = l;
                        if (q >= n || len > orig - n) { err = DEC_ERR_MATCH; break; }        // a slot no position of this block was stored in
                        wcopy_match(out, n, q, len, lane);
                        n += len;
                    }
                } else { if (lane == 0) out[n] = (uint8_t)s; n++; }
                if ((err = rc.err | side.err) != 0) break;
                __syncwarp();
                for (uint32_t p = n - len; p < n; p++) {                        // matcher_update + ppm_update_context per byte
                    const uint32_t byte = len == 1 ? (s == esc ? esc : s) : out[p];
                    hist = hist << 8 | byte;
                    if (p >= 16) {
                        uint32_t h = (hist & 255) * 1313131u + (hist >> 8 & 255) * 13131u + (hist >> 16 & 255) * 131u;
                        if (ctx4) h += hist >> 24;
                        const uint32_t nbucket = h & (RZ_BUCKETS - 1);
                        if (lane == 0) {
                            uint32_t m = m_pref;
                            if ((m >> 16) != B.epoch) m = B.epoch << 16;
                            const uint32_t head = ((m & 255) + 1) & 63;
                            uint32_t count = ((m >> 8) & 255) + 1; if (count > 64) count = 64;
                            T.rz_items[(size_t)bucket * 64 + head] = p; T.rz_meta[bucket] = B.epoch << 16 | count << 8 | head;
                            m_pref = T.rz_meta[nbucket];                        // behind the store above in program order: never stale
                        }
                        bucket = nbucket;
                        const uint32_t sh4 = 4 * (sbucket & 7);
                        const uint32_t nh = (((__shfl_sync(FULLMASK, sheads, sbucket >> 3) >> sh4) & 15u) + 1u) & 15u;
                        if (lane == (sbucket >> 3)) sheads = (sheads & ~(15u << sh4)) | nh << sh4;
                        if (lane == 0) T.rz_short[sbucket * 16 + nh] = p;
                        sbucket = byte;
                    }
                    ctx = ctx << 8 | byte;
                }
                __syncwarp();
            }
        } else if (variant == 2) {                                             // LZ77, src/roxmain/cr-coder.c:388-526
            const uint32_t mm = in[1], esc = in[2];
            auto at = [&](uint32_t off) { return in + (off <= B.in_size ? off : B.in_size); };
            WRc rc, rs, rp, rl; rc.init(in + 32, in_end); rs.init(at(cr_ld32(in + 20)), in_end); rp.init(at(cr_ld32(in + 24)), in_end); rl.init(at(cr_ld32(in + 28)), in_end);
            // the two ROLZ models cached in registers above alias slots 0 and 1: write them through memory instead
            uint32_t n = 0, last = 0;
            while (n < orig) {
                uint32_t len = 1;
                const uint32_t s = wdec_ppm(st, ctx, rc, lane);
                if (s != esc) { if (lane == 0) out[n] = (uint8_t)s; n++; ctx = ctx << 8 | s; }
                else {
                    const uint32_t l = wdec_m0g(st.m0, rl, 30, lane);
                    if (l == 0) { if (lane == 0) out[n] = (uint8_t)esc; n++; ctx = ctx << 8 | esc; }
                    else {
                        uint32_t dist;
                        if (l < mm) dist = wdec_m0g(st.m0 + 256, rs, 1, lane);
                        else dist = x_decode_distance([&](uint32_t j) { return wdec_m0g(st.m0 + (2 + j) * 256, rp, 1u << (2 * j), lane); });
                        if (dist == 0) dist = last;
                        last = dist; len = l;
                        if (dist == 0 || dist > n || len > orig - n) { err = DEC_ERR_MATCH; break; }
                        wcopy_match(out, n, n - dist, len, lane);
                        n += len;
                        __syncwarp();
                        for (uint32_t i = len < 4 ? len : 4; i; i--) ctx = ctx << 8 | out[n - i];
                    }
                }
                if ((err = rc.err | rs.err | rp.err | rl.err) != 0) break;
                __syncwarp();
            }
        } else {
            const uint32_t esc = in[8];
            if (lane < 9 && lane < orig) out[lane] = in[9 + lane];
            __syncwarp();
            WRc rc; rc.init(in + 20, in_end);
            const unsigned long long tag = (unsigned long long)B.epoch << 32;
            unsigned long long h8 = 0;                                          // last eight bytes written (most recent in the top byte)
            for (int i = 1; i < 9; i++) h8 = h8 >> 8 | (unsigned long long)in[9 + i] << 56;
            uint32_t n = 9;
            while (n < orig) {
                uint32_t len = 1;
                const uint32_t s = wdec_ppm(st, ctx, rc, lane);
                uint32_t lit = s;
                if (s == esc) {
                    ctx = ctx << 8 | esc;
                    len = wdec_ppm(st, ctx, rc, lane);
                    if (len == 0) { len = 1; lit = esc; if (lane == 0) out[n] = (uint8_t)esc; n++; }
                    else {
                        LzpDec m; m.T = T; m.tag = tag;
                        const uint32_t q = m.getpos(out, n);
                        if (q >= n || len > orig - n) { err = DEC_ERR_MATCH; break; }
                        wcopy_match(out, n, q, len, lane);
                        n += len;
                    }
                } else { if (lane == 0) out[n] = (uint8_t)s; n++; }
                if ((err = rc.err) != 0) break;
                __syncwarp();
                for (uint32_t p = n - len; p < n; p++) {
                    const uint32_t byte = len == 1 ? lit : out[p];
                    // hashes of the 8 / 4 / 2 bytes in front of p (src/ropmain/cr-matcher.c:31-33): h8 holds bytes p-8..p-1 little endian
                    if (lane == 0) {
                        const uint32_t x4 = (uint32_t)(h8 >> 32), x2 = (uint32_t)(h8 >> 48);
                        T.lzp8[(uint32_t)((h8 ^ h8 >> 20 ^ h8 >> 40) & 0xffffff)] = tag | p;
                        T.lzp4[(x4 ^ x4 >> 6 ^ x4 >> 12) & 0xfffff] = tag | p;
                        T.lzp2[x2] = tag | p;
                    }
                    h8 = h8 >> 8 | (unsigned long long)byte << 56;
                    ctx = ctx << 8 | byte;
                }
                __syncwarp();
            }
        }
    }
    if (variant != 2) {
        ((uint4*)(st.m0))[lane] = make_uint4(fa[0], fa[1], fa[2], fa[3]);
        ((uint4*)(st.m0 + 256))[lane] = make_uint4(fb[0], fb[1], fb[2], fb[3]);
    }
    if (lane == 0) { ctx_io[0] = ctx; ctx_io[1] = err; }
}
__global__ void __launch_bounds__(32) k_lzdecode_warp(int variant, const uint8_t* __restrict__ cont, const DecBlock* __restrict__ blocks, uint32_t nb,
                                                       PpmState st, DecTables T, uint32_t* __restrict__ ctx_io, uint8_t* __restrict__ D) {
    if (blockIdx.x != 0) return;
    lzdecode_warp_body(variant, cont, blocks, nb, st, T, ctx_io, D);
}
// Many containers in flight: one warp per container, every container with its own model state and
// matcher tables.  The chains are independent, so the grid is simply the list of jobs.
__global__ void __launch_bounds__(32) k_lzdecode_jobs(const DecJob* __restrict__ jobs, uint32_t njobs) {
    if (blockIdx.x >= njobs) return;
    const DecJob j = jobs[blockIdx.x];
    if (j.nb == 0) return;
    lzdecode_warp_body(j.variant, j.cont, j.blocks, j.nb, j.st, j.T, j.ctx_io, j.D);
}
#endif  // !CRGPU_SIM
