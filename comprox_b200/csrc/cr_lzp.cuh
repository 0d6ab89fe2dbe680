// cr_lzp.cuh -- LZP match prediction for comprop, all positions in parallel.
//
// Replaces matcher_getpos / matcher_lookup / matcher_update (src/ropmain/cr-matcher.c:59-96) and the look-ahead
// thread (src/ropmain/cr-coder.c:95-118).  The reference keeps three "last occurrence" tables indexed by a
// hash of the 8 / 4 / 2 bytes in front of the position; every position is inserted in order, independent of
// parse decisions, so  table[h] as of time p  ==  the largest q < p with hash(q) == h  (or the table's
// initial value 8 / 4 / 2).  A stable sort by hash puts that q directly in front of p.
#pragma once
#include "cr_common.cuh"
#include "cr_rolz.cuh"   // LzBlock

#define LZP_FIRST   9u      // coding starts at position 9, bytes 0..8 go to the header (src/ropmain/cr-coder.c:143-145)
#define LZP_MINLEN  4u
#define LZP_MAXLEN  255u

CR_HD uint32_t lzp_hash(const uint8_t* x, int kind) {          // src/ropmain/cr-matcher.c:31-33
    if (kind == 2) return x[0] | x[1] << 8;
    if (kind == 1) { uint32_t v = cr_ld32(x); return (v ^ v >> 6 ^ v >> 12) & 0xfffff; }
    uint64_t v = cr_ld32(x) | (uint64_t)cr_ld32(x + 4) << 32;
    return (uint32_t)((v ^ v >> 20 ^ v >> 40) & 0xffffff);
}
CR_HD int lzp_hash_bits(int kind) { return kind == 0 ? 24 : kind == 1 ? 20 : 16; }
CR_HD uint32_t lzp_ctx_bytes(int kind) { return kind == 0 ? 8 : kind == 1 ? 4 : 2; }

__global__ void k_lzp_keys(const uint8_t* __restrict__ D, const LzBlock* __restrict__ blocks, int kind,
                           uint32_t* __restrict__ key, uint32_t* __restrict__ val) {
    const LzBlock B = blocks[blockIdx.y];
    if (B.size < 16) return;
    uint32_t p = LZP_FIRST + blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B.size) return;
    uint32_t e = B.eoff + p - LZP_FIRST;
    key[e] = (blockIdx.y << lzp_hash_bits(kind)) | lzp_hash(D + B.off + p - lzp_ctx_bytes(kind), kind);
    val[e] = p;
}
// cand[e] = table entry position p would read: previous position with the same hash, else the initial value
__global__ void k_lzp_prev(const LzBlock* __restrict__ blocks, int kind, const uint32_t* __restrict__ K, const uint32_t* __restrict__ V, uint32_t n,
                           uint32_t* __restrict__ cand) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint32_t key = K[r], p = V[r];
    const LzBlock B = blocks[key >> lzp_hash_bits(kind)];
    cand[B.eoff + p - LZP_FIRST] = (r > 0 && K[r - 1] == key) ? V[r - 1] : lzp_ctx_bytes(kind);
}
CR_HD bool lzp_same(const uint8_t* a, const uint8_t* b, uint32_t n) { for (uint32_t i = 0; i < n; i++) if (a[i] != b[i]) return false; return true; }

// matcher_lookup: verified hash8 candidate, else verified hash4 candidate, else the (exact) 2-byte context
__global__ void k_lzp_tokens(const uint8_t* __restrict__ D, const LzBlock* __restrict__ blocks, const uint32_t* __restrict__ c8, const uint32_t* __restrict__ c4,
                             const uint32_t* __restrict__ c2, uint8_t* __restrict__ span) {
    const LzBlock B = blocks[blockIdx.y];
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B.size) return;
    uint32_t len = 1;
    if (B.size >= 16 && t >= LZP_FIRST && t + 1024 < B.size) {
        const uint8_t* d = D + B.off;
        uint32_t e = B.eoff + t - LZP_FIRST;
        uint32_t q = c8[e];
        if (!lzp_same(d + q - 8, d + t - 8, 8)) { q = c4[e]; if (!lzp_same(d + q - 4, d + t - 4, 4)) q = c2[e]; }
        uint32_t l = cr_cpl(d + q, d + t, LZP_MAXLEN);
        if (l >= LZP_MINLEN) len = l;
    }
    span[B.off + t] = (uint8_t)len;
}

// ------------------------------------------------------------------ token walk -> PPM events
// coder loop src/ropmain/cr-coder.c:169-207.  A match is coded as two PPM events (esc, then the length with the
// escape shifted into the context); a literal equal to esc is followed by a zero-length event.
// Per chunk we also summarise the bytes the chunk shifts into the PPM context (at most the last four matter),
// so that the context at every chunk entry can be derived without a serial pass.
struct LzpCount {
    typedef struct { uint32_t ev, k, tail; } State;
    const uint8_t* D; const LzBlock* blocks;
    uint32_t *cnt_ev, *tail_k, *tail_bytes;
    CR_D State begin(uint32_t, uint32_t) const { State s = {0, 0, 0}; return s; }
    CR_D void push(State& s, uint32_t byte) const { s.tail = s.tail << 8 | byte; if (s.k < 4) s.k++; }
    CR_D void visit(State& s, uint32_t b, uint32_t t, uint32_t len) const {
        const LzBlock B = blocks[b];
        const uint8_t* d = D + B.off;
        if (len > 1) { s.ev += 2; push(s, B.esc); for (uint32_t i = len > 4 ? len - 4 : 0; i < len; i++) push(s, d[t + i]); }
        else { s.ev += 1 + (d[t] == B.esc); if (d[t] == B.esc) push(s, B.esc); push(s, d[t]); }
    }
    CR_D void end(State& s, uint32_t c, uint32_t) const { cnt_ev[c] = s.ev; tail_k[c] = s.k; tail_bytes[c] = s.tail; }
};

// context on entry to chunk c (c == nchunk: context after the window) by looking back over chunk summaries
__global__ void k_lzp_ctx_in(const uint32_t* __restrict__ tail_k, const uint32_t* __restrict__ tail_bytes, uint32_t nchunk, uint32_t chain_ctx,
                             uint32_t* __restrict__ ctx_in) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > nchunk) return;
    uint32_t have = 0, acc = 0;                    // `have` low bytes of acc are known
    for (uint32_t j = c; j > 0 && have < 4;) {
        j--;
        uint32_t k = tail_k[j], tb = tail_bytes[j];
        for (uint32_t i = 0; i < k && have < 4; i++) { acc |= ((tb >> (8 * i)) & 0xff) << (8 * have); have++; }
    }
    for (uint32_t i = 0; have < 4; i++, have++) acc |= ((chain_ctx >> (8 * i)) & 0xff) << (8 * have);
    ctx_in[c] = acc;
}

struct LzpEmit {
    typedef struct { uint32_t ev, ctx; } State;
    const uint8_t* D; const LzBlock* blocks;
    const uint32_t *scan_ev, *ctx_in;
    uint32_t* ev_ctx; uint8_t* ev_sym; uint8_t* tokend;
    CR_D State begin(uint32_t c, uint32_t) const { State s = { scan_ev[c], ctx_in[c] }; return s; }
    CR_D void put(State& s, uint32_t sym, uint32_t end) const { ev_ctx[s.ev] = s.ctx; ev_sym[s.ev] = (uint8_t)sym; tokend[s.ev] = (uint8_t)end; s.ev++; }
    CR_D void visit(State& s, uint32_t b, uint32_t t, uint32_t len) const {
        const LzBlock B = blocks[b];
        const uint8_t* d = D + B.off;
        if (len > 1) {
            put(s, B.esc, 0); s.ctx = s.ctx << 8 | B.esc;
            put(s, len, 1);
        } else {
            put(s, d[t], d[t] != B.esc);
            if (d[t] == B.esc) { s.ctx = s.ctx << 8 | B.esc; put(s, 0, 1); }
        }
        for (uint32_t i = len > 4 ? len - 4 : 0; i < len; i++) s.ctx = s.ctx << 8 | d[t + i];
    }
    CR_D void end(State&, uint32_t, uint32_t) const {}
};

__global__ void k_lzp_finish_blocks(LzBlock* __restrict__ blocks, uint32_t nb, const uint8_t* __restrict__ esc1) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nb) blocks[b].esc = esc1[b];
}
