// cr_ppm.cuh -- exact replay of the chained PPM / order-0 models, decomposed by context.
//
// The reference codes one event at a time through ppm_encode (src/cr-ppm.c:103-167) with model state that
// is carried across blocks, so a block cannot be coded independently.  The encoder, however,
// knows every (context, symbol) event in advance, and each piece of model state is touched only by the
// events of ONE context:
//     o3 predictor slot  <- events whose 22-bit slot hash is equal        (cr-ppm.c:66-88)
//     o2 model           <- events with equal ctx & 0xFFFF                 (cr-o2model.c)
//     o1 row             <- escape events with equal ctx & 0xFF            (cr-ppm.c:90-98,148-158)
// Grouping the events stably by each of these keys and replaying every group in order reproduces exactly the
// state sequence of the serial coder; the result per event is the (cum, frq, sum) triple(s) the serial coder
// would have handed to range_encoder_encode.  Model state lives in device memory between calls, so a chain
// can be fed window by window (or block by block through the reference-signature shim).
#pragma once
#include "cr_common.cuh"

#define PPM_O3_SLOTS (1u << 22)
#define PPM_O2_STRIDE 264u          // 258 frequencies + padding, 8-byte aligned rows

struct PpmState {
    uint8_t* o3_byte;    // [PPM_O3_SLOTS] predicted byte
    uint8_t* o3_conf;    // [PPM_O3_SLOTS] 4-bit confidence
    uint8_t* o2;         // [65536][PPM_O2_STRIDE]
    uint8_t* o1;         // [256][256]
    uint16_t* m0;        // [2][256] order-0 models: 0 = len_model, 1 = idx_model (src/rolzmain/cr-coder.c:52-56)
};

CR_HD uint32_t ppm_slot(uint32_t ctx) { return (ctx ^ (ctx >> 2)) & 0x3fffff; }           // cr-ppm.c:66

// reset_models(): src/rolzmain/cr-coder.c:78-96, src/cr-ppm.c:34-48, src/cr-o2model.c:31-41
__global__ void k_ppm_reset(PpmState st) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;      // one thread per o2 context
    if (i >= 65536) return;
    uint8_t* f = st.o2 + (size_t)i * PPM_O2_STRIDE;
    for (uint32_t k = 0; k < PPM_O2_STRIDE; k++) f[k] = 0;
    f[256] = 1; f[257] = 1;
    st.o1[i] = 1;
    if (i < 256) {
        st.m0[i] = (i == 0 || i >= 5) ? 1 : 0;                // len_model
        st.m0[256 + i] = i < 80 ? 1 : 0;                      // idx_model
    }
}

// ------------------------------------------------------------------ o3 pass
__global__ void k_o3_keys(const uint32_t* __restrict__ ev_ctx, const uint8_t* __restrict__ ev_sym, uint32_t n,
                          uint32_t* __restrict__ key, uint32_t* __restrict__ val) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    key[e] = (uint32_t)ev_sym[e] << 24 | ppm_slot(ev_ctx[e]);
    val[e] = e;
}

// One thread per slot segment of the slot-sorted event list: replays ppm_update_o3 (cr-ppm.c:69-88) and
// records the byte the predictor held BEFORE each event (predict_ch, cr-ppm.c:114).
struct O3Hot { uint32_t slot, rank, byte, conf; };   // a long slot segment handed over to k_o3_hot (cr_warp.cuh)
#define O3_HANDOVER 2048u
__global__ void k_o3_pass(const uint32_t* __restrict__ K, const uint32_t* __restrict__ V, uint32_t n, PpmState st, uint8_t* __restrict__ pred,
                          O3Hot* __restrict__ hot, uint32_t* __restrict__ hot_count, uint32_t hot_cap) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint32_t slot = K[r] & 0x3fffff;
    if (r > 0 && (K[r - 1] & 0x3fffff) == slot) return;
    uint32_t byte = st.o3_byte[slot], conf = st.o3_conf[slot];
    bool more = true;
    for (uint32_t i = r; more && i < n; i += 8) {
        if (hot && i - r >= O3_HANDOVER) {             // long segment: a warp with staged input continues from here
            uint32_t h = atomicAdd(hot_count, 1u);
            if (h < hot_cap) { hot[h].slot = slot; hot[h].rank = i; hot[h].byte = byte; hot[h].conf = conf; return; }
        }
        uint32_t kk[8], vv[8];                       // 8 independent loads in flight per round trip
#pragma unroll
        for (int u = 0; u < 8; u++) { const uint32_t x = i + u < n ? i + u : n - 1; kk[u] = K[x]; vv[u] = V[x]; }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (!more || i + u >= n || (kk[u] & 0x3fffff) != slot) { more = false; continue; }
            const uint32_t sym = kk[u] >> 24;
            pred[vv[u]] = (uint8_t)byte;
            if (sym == byte) conf += conf < 15;
            else {
                conf = (conf > 1) + (conf > 2) + (conf > 4) + (conf > 8);
                if (conf == 0) { byte = sym; conf = 1; }
            }
        }
    }
    st.o3_byte[slot] = (uint8_t)byte;
    st.o3_conf[slot] = (uint8_t)conf;
}

// ------------------------------------------------------------------ o2 pass
__global__ void k_o2_keys(const uint32_t* __restrict__ ev_ctx, const uint8_t* __restrict__ ev_sym, const uint8_t* __restrict__ pred, uint32_t n,
                          uint32_t* __restrict__ key, uint32_t* __restrict__ val) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    key[e] = (uint32_t)ev_sym[e] << 24 | (uint32_t)pred[e] << 16 | (ev_ctx[e] & 0xffff);
    val[e] = e;
}

// packed triple: cum[0:24) | frq[24:40) | sum[40:63) | escape-follows[63]
CR_HD uint64_t ppm_pack(uint32_t cum, uint32_t frq, uint32_t sum, uint32_t esc) {
    return (uint64_t)cum | (uint64_t)frq << 24 | (uint64_t)sum << 40 | (uint64_t)esc << 63;
}

struct EscRec {          // one o2 escape, handed to the o1 pass
    uint32_t e;          // event index (time order within the chain window)
    uint32_t info;       // ctx8 | sym << 8
    uint32_t incl[8];    // bit i set: symbol i takes part in the o1 sum (o2 frq == 0 after the escape update, and != predicted byte)
};

// o2_model_update (cr-o2model.c:43-72) on a table held as f[258] + 8 group sums + body total.
struct O2Tab {
    uint8_t f[258];
    uint16_t g[8];
    uint32_t body;
};
CR_D void o2_regroup(O2Tab& t) {
    uint32_t b = 0;
    for (int k = 0; k < 8; k++) { uint32_t s = 0; for (int i = 0; i < 32; i++) s += t.f[k * 32 + i]; t.g[k] = (uint16_t)s; b += s; }
    t.body = b;
}
CR_D int o2_bump(O2Tab& t, uint32_t s, int inc) {
    t.f[s] = (uint8_t)(t.f[s] + inc);
    if (s < 256) { t.g[s >> 5] = (uint16_t)(t.g[s >> 5] + inc); t.body += inc; }
    if (t.f[s] > 250) {
        uint32_t ee = 1;
        for (int i = 0; i < 256; i++) { t.f[i] >>= 1; ee += t.f[i] == 1; }
        t.f[256] = (uint8_t)((t.f[256] + 1) / 2);
        t.f[257] = (uint8_t)ee;
        o2_regroup(t);
        return 1;
    }
    return 0;
}
CR_D uint32_t o2_cum_below(const O2Tab& t, uint32_t s) {
    uint32_t c = 0;
    for (uint32_t k = 0; k < (s >> 5); k++) c += t.g[k];
    for (uint32_t i = s & ~31u; i < s; i++) c += t.f[i];
    return c;
}

// One thread per ctx16 segment of the ctx16-sorted event list: replays the o2 part of ppm_encode.
__global__ void k_o2_pass(const uint32_t* __restrict__ K, const uint32_t* __restrict__ V, uint32_t n, PpmState st,
                          uint64_t* __restrict__ T1, EscRec* __restrict__ esc_rec, uint32_t* __restrict__ esc_count) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint32_t c16 = K[r] & 0xffff;
    if (r > 0 && (K[r - 1] & 0xffff) == c16) return;
    uint8_t* row = st.o2 + (size_t)c16 * PPM_O2_STRIDE;
    O2Tab t;
    for (int i = 0; i < 258; i++) t.f[i] = row[i];
    o2_regroup(t);
    for (uint32_t i = r; i < n; i++) {
        const uint32_t k = K[i];
        if ((k & 0xffff) != c16) break;
        const uint32_t sym = k >> 24, pr = (k >> 16) & 255, e = V[i];
        const uint32_t pf = t.f[pr];
        const uint32_t sum = t.body + t.f[256] + t.f[257] - pf;
        if (sym == pr) {                                                       // cr-ppm.c:118-125
            T1[e] = ppm_pack(t.body - pf, t.f[256], sum, 0);
            o2_bump(t, 256, 1);
        } else if (t.f[sym] > 0) {                                             // cr-ppm.c:128-138
            T1[e] = ppm_pack(o2_cum_below(t, sym) - (sym >= pr ? pf : 0), t.f[sym], sum, 0);
            if (!o2_bump(t, sym, 1) && t.f[sym] == 2) o2_bump(t, 257, -1);
        } else {                                                               // cr-ppm.c:140-162
            T1[e] = ppm_pack(t.body + t.f[256] - pf, t.f[257], sum, 1);
            int rescaled = o2_bump(t, 257, 1);
            uint32_t slot = atomicAdd(esc_count, 1u);
            EscRec rec;
            rec.e = e;
            rec.info = (c16 & 0xff) | sym << 8;
            for (int w = 0; w < 8; w++) {
                uint32_t m = 0;
                for (int b = 0; b < 32; b++) m |= (uint32_t)(t.f[w * 32 + b] == 0) << b;
                rec.incl[w] = m;
            }
            rec.incl[pr >> 5] &= ~(1u << (pr & 31));
            esc_rec[slot] = rec;
            if (!rescaled) o2_bump(t, sym, 1);
        }
    }
    for (int i = 0; i < 258; i++) row[i] = t.f[i];
}

// ------------------------------------------------------------------ o1 pass
__global__ void k_o1_keys(const EscRec* __restrict__ rec, uint32_t n, uint64_t* __restrict__ key, uint32_t* __restrict__ val) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    key[i] = (uint64_t)(rec[i].info & 0xff) << 32 | rec[i].e;
    val[i] = i;
}
// also sorted by event index alone, this list gives each escape its ordinal inside the window
__global__ void k_o1_ordinals(const uint32_t* __restrict__ val_by_e, uint32_t n, uint32_t* __restrict__ ord) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ord[val_by_e[i]] = i;
}

// One thread per ctx8 segment of the (ctx8, time)-sorted escapes: o1 coding with exclusion, cr-ppm.c:148-158,90-98.
__global__ void k_o1_pass(const uint64_t* __restrict__ K, const uint32_t* __restrict__ V, uint32_t n, const EscRec* __restrict__ rec,
                          const uint32_t* __restrict__ ord, PpmState st, uint64_t* __restrict__ T2) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint32_t c8 = (uint32_t)(K[r] >> 32) & 0xff;
    if (r > 0 && ((uint32_t)(K[r - 1] >> 32) & 0xff) == c8) return;
    uint8_t* row = st.o1 + c8 * 256;
    uint8_t o1[256];
    for (int i = 0; i < 256; i++) o1[i] = row[i];
    for (uint32_t i = r; i < n; i++) {
        if (((uint32_t)(K[i] >> 32) & 0xff) != c8) break;
        const EscRec& R = rec[V[i]];
        const uint32_t sym = R.info >> 8 & 0xff;
        uint32_t cum = 0, sum = 0;
        for (uint32_t w = 0; w < 8; w++) {
            uint32_t m = R.incl[w];
            while (m) {
                uint32_t b = 31 - __clz(m & (0u - m));
                m &= m - 1;
                uint32_t s = w * 32 + b, fr = (uint32_t)o1[s] * 8 - 7;
                sum += fr;
                if (s < sym) cum += fr;
            }
        }
        T2[ord[V[i]]] = ppm_pack(cum, (uint32_t)o1[sym] * 8 - 7, sum, 0);
        if (++o1[sym] >= 255) for (int j = 0; j < 256; j++) o1[j] -= o1[j] / 2;
    }
    for (int i = 0; i < 256; i++) row[i] = o1[i];
}

// ------------------------------------------------------------------ order-0 side models (ROLZ len / idx)
// model_cum / model_update with increment 4 (src/cr-model.c:55-88, M_my_enc_ src/cr-model.h:58-64).  A single
// context per model, so this is one serial chain over the window's side symbols.
__global__ void k_side_models(const uint16_t* __restrict__ side_sym, uint32_t n, PpmState st, uint64_t* __restrict__ TS) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    uint16_t f[2][256];
    uint32_t g[2][8], total[2];
    for (int m = 0; m < 2; m++) {
        total[m] = 0;
        for (int k = 0; k < 8; k++) { uint32_t s = 0; for (int i = 0; i < 32; i++) { f[m][k * 32 + i] = st.m0[m * 256 + k * 32 + i]; s += f[m][k * 32 + i]; } g[m][k] = s; total[m] += s; }
    }
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t v = side_sym[i], m = v >> 8, s = v & 255;
        uint32_t cum = 0;
        for (uint32_t k = 0; k < (s >> 5); k++) cum += g[m][k];
        for (uint32_t j = s & ~31u; j < s; j++) cum += f[m][j];
        TS[i] = ppm_pack(cum, f[m][s], total[m], 0);
        f[m][s] += 4; g[m][s >> 5] += 4; total[m] += 4;
        if (total[m] > 32000) {
            total[m] = 0;
            for (int k = 0; k < 8; k++) { uint32_t t = 0; for (int j = 0; j < 32; j++) { f[m][k * 32 + j] = (uint16_t)((f[m][k * 32 + j] + 1) / 2); t += f[m][k * 32 + j]; } g[m][k] = t; total[m] += t; }
        }
    }
    for (int m = 0; m < 2; m++) for (int i = 0; i < 256; i++) st.m0[m * 256 + i] = f[m][i];
}
