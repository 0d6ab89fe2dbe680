// cr_decode.cuh -- decompression (the inverse path): lzdecode, dictionary_decode, inverse filters.
//
// Replaces lzdecode (src/rolzmain/cr-coder.c:287-379, src/ropmain/cr-coder.c:231-292), ppm_decode
// (src/cr-ppm.c:169-235), the decoder half of src/cr-rangecoder.c:81-104, dictionary_decode(_imp)
// (src/cr-diccode.c:223-283,364-425) and filter_inplace(FILTER_DEC).
//
// Decoding one container is ONE serial chain: the context of every symbol depends on the
// bytes decoded before it and the models carry across blocks, so -- unlike compression -- nothing can be replayed
// per context.  k_lzdecode_serial therefore runs one thread per container; throughput comes only from decoding
// many containers at once ("replicas only").  The stages after it are parallel again: one thread per 1 000 000-byte
// sub-chunk for the (backwards) dictionary expansion, byte-parallel inverse filters.
// The matcher tables are not cleared per block: every entry carries the ordinal of the block that wrote it.
#pragma once
#include "cr_common.cuh"
#include "cr_ppm.cuh"
#include "cr_rolz.cuh"
#include "cr_lzp.cuh"

struct DecBlock {
    uint64_t in_off;      // payload (after the 6-byte block header) in the container buffer
    uint32_t in_size;
    uint32_t coded;       // 1 = lzencode'd payload with "compressed" set: run the serial decoder
    uint64_t d_off;       // where the dictionary-coded block goes
    uint32_t d_size;      // its size (original_size of the inner header)
    uint32_t epoch;       // ordinal used to tag matcher table entries (> 0)
};
struct DecTables {        // matcher state of ONE container
    uint32_t* rz_meta;    // [262144]  epoch << 16 | count << 8 | head
    uint32_t* rz_items;   // [262144 * 64]
    uint32_t* rz_short;   // [256 * 16]
    unsigned long long* lzp8;   // [1 << 24]  epoch << 32 | position
    unsigned long long* lzp4;   // [1 << 20]
    unsigned long long* lzp2;   // [1 << 16]
};

struct DecJob {           // everything one decode chain (the consecutive blocks of one container) needs
    int variant; uint32_t nb;
    const uint8_t* cont; const DecBlock* blocks;
    PpmState st; DecTables T;
    uint32_t* ctx_io; uint8_t* D;      // ctx_io[0]: PPM context in / out; ctx_io[1]: error code out (0 = fine)
};

// ------------------------------------------------------------------ range decoder (cr-rangecoder.c:81-104)
// Untrusted input: a stream never reads behind `end` (the end of its payload), a zero divisor / zero range / zero frequency (only a
// damaged stream produces them) sets `err` instead of dividing by zero or spinning in the renormalisation loop.  The decoders test
// `err` once per symbol and give the container up (CRGPU_ERR_CORRUPT); the reference crashes on such input.
enum { DEC_ERR_STREAM = 1, DEC_ERR_MATCH = 2, DEC_ERR_DICT = 3 };
struct RcDec {
    uint32_t range, code, err;
    const uint8_t* in; const uint8_t* end;
    CR_D uint32_t next() { if (in < end) return *in++; err = DEC_ERR_STREAM; return 0; }
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

CR_D int dec_o2_bump(uint8_t* f, uint32_t s, int inc) {       // o2_model_update, cr-o2model.c:43-72
    f[s] = (uint8_t)(f[s] + inc);
    if (f[s] > 250) {
        uint32_t ee = 1;
        for (int i = 0; i < 256; i++) { f[i] >>= 1; ee += f[i] == 1; }
        f[256] = (uint8_t)((f[256] + 1) / 2);
        f[257] = (uint8_t)ee;
        return 1;
    }
    return 0;
}
CR_D void dec_o3_update(PpmState& st, uint32_t slot, int c) {  // ppm_update_o3, cr-ppm.c:69-88
    uint32_t f = st.o3_conf[slot];
    if (c >= 0) { f = (f > 1) + (f > 2) + (f > 4) + (f > 8); if (f == 0) { st.o3_byte[slot] = (uint8_t)c; f = 1; } }
    else f += f < 15;
    st.o3_conf[slot] = (uint8_t)f;
}
// ppm_decode, cr-ppm.c:169-235
CR_D uint32_t dec_ppm(PpmState& st, uint32_t ctx, RcDec& rc) {
    uint8_t* f = st.o2 + (size_t)(ctx & 0xffff) * PPM_O2_STRIDE;
    uint8_t* o1 = st.o1 + (ctx & 0xff) * 256;
    const uint32_t slot = ppm_slot(ctx);
    const uint32_t pred = st.o3_byte[slot];
    uint32_t body = 0;
    for (int i = 0; i < 256; i++) body += f[i];
    const uint32_t pf = f[pred];
    const uint32_t tgt = rc.target(body + f[256] + f[257] - pf);
    uint32_t acc = 0, s = 0;
    for (;; s++) { const uint32_t wgt = s == pred ? 0u : f[s]; if (acc + wgt > tgt || s == 257) break; acc += wgt; }
    rc.consume(acc, f[s]);
    const int rescaled = dec_o2_bump(f, s, 1);
    if (s == 256) { dec_o3_update(st, slot, -1); return pred; }
    if (s < 256) {
        if (!rescaled && f[s] == 2) dec_o2_bump(f, 257, -1);
        dec_o3_update(st, slot, (int)s);
        return s;
    }
    uint32_t sum1 = 0, cum1 = 0, d = 0;
    for (uint32_t i = 0; i < 256; i++) if (f[i] == 0 && i != pred) sum1 += (uint32_t)o1[i] * 8 - 7;
    const uint32_t t1 = rc.target(sum1);
    for (uint32_t i = 0; i < 256; i++)
        if (f[i] == 0 && i != pred) { const uint32_t fr = (uint32_t)o1[i] * 8 - 7; if (cum1 + fr > t1) { d = i; break; } cum1 += fr; }
    rc.consume(cum1, (uint32_t)o1[d] * 8 - 7);
    if (++o1[d] >= 255) for (int j = 0; j < 256; j++) o1[j] -= o1[j] / 2;
    if (!rescaled) dec_o2_bump(f, d, 1);
    dec_o3_update(st, slot, (int)d);
    return d;
}
// M_my_dec_ with increment 4 (cr-model.h:66-74, cr-model.c:55-77,98-115)
CR_D uint32_t dec_m0(uint16_t* f, RcDec& rc, uint32_t inc = 4) {
    uint32_t total = 0;
    for (int i = 0; i < 256; i++) total += f[i];
    const uint32_t tgt = rc.target(total);
    uint32_t acc = 0, s = 0;
    while (s < 255 && acc + f[s] <= tgt) acc += f[s++];
    rc.consume(acc, f[s]);
    f[s] = (uint16_t)(f[s] + inc);
    if (total + inc > 32000) for (int i = 0; i < 256; i++) f[i] = (uint16_t)((f[i] + 1) / 2);
    return s;
}

// ------------------------------------------------------------------ ROLZ tables with block tags
struct RzDec {
    DecTables T; uint32_t epoch, bucket, short_bucket; int ctx4;
    CR_D void begin(const DecTables& t, uint32_t ep, int c4) {
        T = t; epoch = ep; bucket = 0; short_bucket = 0; ctx4 = c4;
        for (int i = 0; i < 256 * 16; i++) T.rz_short[i] = 0;                  // cr-matcher.c:53
    }
    CR_D void insert(const uint8_t* d, uint32_t pos) {                         // matcher_update, cr-matcher.c:65-81
        if (pos < 16) return;
        uint32_t m = T.rz_meta[bucket];
        if ((m >> 16) != epoch) m = epoch << 16;
        const uint32_t head = ((m & 255) + 1) & 63;
        uint32_t count = ((m >> 8) & 255) + 1; if (count > 64) count = 64;
        T.rz_items[(size_t)bucket * 64 + head] = pos;
        T.rz_meta[bucket] = epoch << 16 | count << 8 | head;
        bucket = rz_hash(d + pos, ctx4);
        uint32_t* s = T.rz_short + short_bucket * 16;
        for (int i = 15; i > 0; i--) s[i] = s[i - 1];
        s[0] = pos;
        short_bucket = d[pos];
    }
    CR_D uint32_t getpos(uint32_t idx) const {                                 // matcher_getpos, cr-matcher.c:83-88
        if (idx < 64) { const uint32_t m = T.rz_meta[bucket]; return T.rz_items[(size_t)bucket * 64 + (((m & 255) + 64 - idx) & 63)]; }
        return T.rz_short[short_bucket * 16 + idx - 64];
    }
};
struct LzpDec {
    DecTables T; unsigned long long tag;
    CR_D void begin(const DecTables& t, uint32_t ep) { T = t; tag = (unsigned long long)ep << 32; }
    CR_D uint32_t get(const unsigned long long* tab, uint32_t h, uint32_t dflt) const { const unsigned long long v = tab[h]; return (v >> 32 << 32) == tag ? (uint32_t)v : dflt; }
    CR_D uint32_t getpos(const uint8_t* d, uint32_t pos) const {               // matcher_getpos, ropmain/cr-matcher.c:59-73
        const uint32_t a = get(T.lzp8, lzp_hash(d + pos - 8, 0), 8), b = get(T.lzp4, lzp_hash(d + pos - 4, 1), 4), c = get(T.lzp2, lzp_hash(d + pos - 2, 2), 2);
        if (lzp_same(d + a - 8, d + pos - 8, 8)) return a;
        if (lzp_same(d + b - 4, d + pos - 4, 4)) return b;
        return c;
    }
    CR_D void insert(const uint8_t* d, uint32_t pos) {                         // matcher_update, :91-96
        T.lzp8[lzp_hash(d + pos - 8, 0)] = tag | pos; T.lzp4[lzp_hash(d + pos - 4, 1)] = tag | pos; T.lzp2[lzp_hash(d + pos - 2, 2)] = tag | pos;
    }
};

// distance of a long LZ77 match: lzdecode_pos_thread, src/roxmain/cr-coder.c:350-373.  next(j) decodes one symbol with pos_models[j].
template <class F> CR_D uint32_t x_decode_distance(F next) {
    uint32_t j = 0, v = 0, sym = 0;
    while (j < 2 && (sym = next(j)) >= 128) { v += (sym - 128) * (1u << (7 * j)); j++; }
    if (j < 2) return (v + sym * (1u << (7 * j))) / 8;
    while (j < 5 && (sym = next(j)) >= 64) { v += (sym - 64) * (1u << (6 * j + 2)); j++; }
    return (v + sym * (1u << (6 * j + 2))) / 8;
}

// One thread decodes all lz-coded blocks of one container in order.  ctx_io carries the PPM context in and out.
__global__ void k_lzdecode_serial(int variant, const uint8_t* __restrict__ cont, const DecBlock* __restrict__ blocks, uint32_t nb,
                                  PpmState st, DecTables tabs, uint32_t* __restrict__ ctx_io, uint8_t* __restrict__ D) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    uint32_t ctx = *ctx_io, err = 0;
    for (uint32_t b = 0; b < nb && !err; b++) {
        const DecBlock B = blocks[b];
        if (!B.coded) continue;
        const uint8_t* in = cont + B.in_off;
        const uint8_t* in_end = in + B.in_size;
        uint8_t* out = D + B.d_off;
        const uint32_t orig = B.d_size;
        if (orig == 0) continue;
        if (variant == 0) {                                                    // src/rolzmain/cr-coder.c:287-379
            const uint32_t esc = in[2], off_idx = cr_ld32(in + 12);
            RcDec rc, side; rc.init(in + 16, in_end); side.init(in + (off_idx <= B.in_size ? off_idx : B.in_size), in_end);
            RzDec m; m.begin(tabs, B.epoch, orig >= 4194304);
            uint32_t n = 0;
            out[n++] = in[0];
            while (n < orig) {
                uint32_t len = 1;
                const uint32_t s = dec_ppm(st, ctx, rc);
                if (s == esc) {
                    const uint32_t l = dec_m0(st.m0, side);
                    if (l == 0) out[n++] = (uint8_t)esc;
                    else {
                        const uint32_t idx = dec_m0(st.m0 + 256, side);
                        const uint32_t q = m.getpos(idx);
                        len = l;
                        if (q >= n || len > orig - n) { err = DEC_ERR_MATCH; break; }          // a slot no position of this block was stored in
                        for (uint32_t i = 0; i < len; i++) { out[n] = out[q + i]; n++; }
                    }
                } else out[n++] = (uint8_t)s;
                if ((err = rc.err | side.err) != 0) break;
                for (; len; len--) { const uint32_t p = n - len; m.insert(out, p); ctx = ctx << 8 | out[p]; }
            }
        } else if (variant == 2) {                                             // src/roxmain/cr-coder.c:388-526
            const uint32_t mm = in[1], esc = in[2];
            auto at = [&](uint32_t off) { return in + (off <= B.in_size ? off : B.in_size); };
            RcDec rc, rs, rp, rl; rc.init(in + 32, in_end); rs.init(at(cr_ld32(in + 20)), in_end); rp.init(at(cr_ld32(in + 24)), in_end); rl.init(at(cr_ld32(in + 28)), in_end);
            uint32_t n = 0, last = 0;
            while (n < orig) {
                uint32_t len = 1;
                const uint32_t s = dec_ppm(st, ctx, rc);
                if (s != esc) out[n++] = (uint8_t)s;
                else {
                    const uint32_t l = dec_m0(st.m0, rl, 30);
                    if (l == 0) out[n++] = (uint8_t)esc;
                    else {
                        uint32_t dist;
                        if (l < mm) dist = dec_m0(st.m0 + 256, rs, 1);
                        else dist = x_decode_distance([&](uint32_t j) { return dec_m0(st.m0 + (2 + j) * 256, rp, 1u << (2 * j)); });
                        if (dist == 0) dist = last;
                        last = dist; len = l;
                        if (dist == 0 || dist > n || len > orig - n) { err = DEC_ERR_MATCH; break; }
                        const uint32_t q = n - dist;
                        for (uint32_t i = 0; i < len; i++) { out[n] = out[q + i]; n++; }
                    }
                }
                if ((err = rc.err | rs.err | rp.err | rl.err) != 0) break;
                for (; len; len--) ctx = ctx << 8 | out[n - len];
            }
        } else {                                                               // src/ropmain/cr-coder.c:231-292
            const uint32_t esc = in[8];
            for (uint32_t i = 0; i < 9 && i < orig; i++) out[i] = in[9 + i];
            RcDec rc; rc.init(in + 20, in_end);
            LzpDec m; m.begin(tabs, B.epoch);
            uint32_t n = 9;
            while (n < orig) {
                uint32_t len = 1;
                const uint32_t s = dec_ppm(st, ctx, rc);
                if (s != esc) out[n++] = (uint8_t)s;
                else {
                    ctx = ctx << 8 | esc;
                    len = dec_ppm(st, ctx, rc);
                    if (len == 0) { len = 1; out[n++] = (uint8_t)esc; }
                    else {
                        const uint32_t q = m.getpos(out, n);
                        if (q >= n || len > orig - n) { err = DEC_ERR_MATCH; break; }
                        for (uint32_t i = 0; i < len; i++) { out[n] = out[q + i]; n++; }
                    }
                }
                if ((err = rc.err) != 0) break;
                for (; len; len--) { const uint32_t p = n - len; ctx = ctx << 8 | out[p]; m.insert(out, p); }
            }
        }
    }
    ctx_io[0] = ctx; ctx_io[1] = err;
}

// ------------------------------------------------------------------ dictionary_decode (cr-diccode.c:223-283)
struct DdSub { uint64_t src; uint32_t size; uint32_t block; uint64_t dst; uint32_t orig; uint32_t pad; };   // one sub-chunk
struct DdBlock { uint64_t d_off; uint32_t d_size; uint32_t first_sub; uint64_t raw_off; uint32_t raw_size; uint32_t nsub; };
// Walks the pair framing of every dictionary-coded block (serial over a few hundred headers) and lays out the output.
// totals[2] != 0: the framing of a block is damaged (sizes that leave the block, sub-chunks above 1 000 000 bytes).
__global__ void k_dd_layout(const uint8_t* __restrict__ D, DdBlock* __restrict__ blocks, uint32_t nb, DdSub* __restrict__ subs, uint32_t sub_cap,
                            uint64_t out_base, uint64_t* __restrict__ totals) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    uint64_t raw = out_base; uint32_t ns = 0; uint64_t bad = 0;
    for (uint32_t b = 0; b < nb; b++) {
        DdBlock& B = blocks[b];
        const uint8_t* d = D + B.d_off;
        B.raw_off = raw; B.first_sub = ns; B.nsub = 0;
        if (B.d_size == 0) { B.raw_size = 0; continue; }
        if (d[B.d_size - 1] == 0) { B.raw_size = B.d_size - 1; raw += B.raw_size; continue; }      // stored (:237-241)
        uint32_t size = 0;
        if (B.d_size < 11) { bad = 1; B.raw_size = 0; continue; }
        for (uint32_t pos = 0; pos + 11 < B.d_size;) {
            const uint32_t s1 = cr_ld32(d + pos), s2 = cr_ld32(d + pos + 4);
            // a pair is  u32 len1, u32 len2, codes1 + u32 orig1, codes2 + u32 orig2  and the block ends with esc[10], 1 (cr-diccode.c:173-207)
            if (s1 < 4 || s2 < 4 || (uint64_t)pos + 8 + s1 + s2 + 11 > B.d_size) { bad = 1; break; }
            const uint32_t sz[2] = { s1, s2 }; uint64_t src = B.d_off + pos + 8;
            for (int k = 0; k < 2; k++) {
                const uint32_t orig = cr_ld32(D + src + sz[k] - 4);
                if (orig > 1000000u) bad = 1;                                   // sub-chunks are at most 1 000 000 bytes (cr-diccode.c:174)
                if (ns < sub_cap) { DdSub S; S.src = src; S.size = sz[k]; S.block = b; S.dst = raw + size; S.orig = orig; S.pad = 0; subs[ns] = S; }
                ns++; B.nsub++; size += orig; src += sz[k];
            }
            pos += 8 + s1 + s2;
        }
        B.raw_size = size; raw += size;
    }
    totals[0] = raw - out_base; totals[1] = ns; totals[2] = bad;
}
struct DdDict { const char* words; const uint8_t* lens; int32_t nentries, level1; };   // words: [nentries][24]
CR_HD bool dd_sentence_start(const uint8_t* s, uint32_t i) { return i >= 3 && s[i - 1] == ' ' && (s[i - 2] == '.' || (s[i - 2] == ' ' && s[i - 3] == '.')); }
// dictionary_decode_imp (cr-diccode.c:364-425): each sub-chunk is expanded back to front by one thread
// Untrusted input: the code bytes are read back to front from index size-4; a damaged sub-chunk (a code cut off at the front, a word
// index the dictionary does not have, more bytes than `orig` says) sets *err and stops instead of leaving its buffers.
CR_D void dd_sub_body(const uint8_t* __restrict__ D, const DdBlock* __restrict__ blocks, const DdSub S, const DdDict dic, uint8_t* __restrict__ out, uint32_t* __restrict__ err) {
    const DdBlock B = blocks[S.block];
    const uint8_t* esc = D + B.d_off + B.d_size - 11;
    uint8_t escmap[256];
    for (int i = 0; i < 256; i++) escmap[i] = 0;
    for (int i = 0; i < 10; i++) escmap[esc[i]] = (uint8_t)(i + 1);
    const uint8_t* d = D + S.src;
    uint8_t* o = out + S.dst;
    const int L1 = dic.level1;
    uint32_t src = S.orig, rev = 0xFFFFFFFFu; int dst = (int)S.size - 4;
    while (src > 0) {
        if (dst < 1) { *err = DEC_ERR_DICT; return; }
        const uint32_t ch = d[--dst];
        if (!escmap[ch]) { o[--src] = (uint8_t)ch; continue; }
        if (dst < 1) { *err = DEC_ERR_DICT; return; }
        int id = d[--dst];
        if (id >= L1) {
            if (dst < 1) { *err = DEC_ERR_DICT; return; }
            id = d[--dst] * (256 - L1) + (id - L1);
            if (id == dic.nentries) { o[--src] = (uint8_t)ch; continue; }
        }
        if (id < 0 || id >= dic.nentries) { *err = DEC_ERR_DICT; return; }
        const uint32_t wl = dic.lens[id];
        const char* w = dic.words + (size_t)id * 24;
        if (wl == 0 || wl > 24 || wl > src) { *err = DEC_ERR_DICT; return; }
        src -= wl;
        for (uint32_t i = 0; i < wl; i++) o[src + i] = (uint8_t)w[i];
        const uint32_t e = escmap[ch];
        if (e == 2 || e == 7) o[src + wl - 1] = '.';
        else if (e == 3 || e == 8) o[src + wl - 1] = ',';
        else if (e == 4 || e == 9) o[src + wl - 1] = ';';
        else if (e == 5 || e == 10) o[src + wl - 1] = ':';
        if (e >= 6) o[src] ^= 0x20;
        if (rev != 0xFFFFFFFFu && dd_sentence_start(o, rev)) o[rev] ^= 0x20;
        rev = src;
    }
    if (rev != 0xFFFFFFFFu && dd_sentence_start(o, rev)) o[rev] ^= 0x20;
}
__global__ void k_dd_subs(const uint8_t* __restrict__ D, const DdBlock* __restrict__ blocks, const DdSub* __restrict__ subs, uint32_t nsub, DdDict dic, uint8_t* __restrict__ out, uint32_t* __restrict__ err) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nsub) return;
    dd_sub_body(D, blocks, subs[t], dic, out, err);
}
// the same over the sub-chunks of many containers (crgpu_decompress_batch): first[j] = index of job j's first sub-chunk
struct DdJob { const uint8_t* D; const DdBlock* blocks; const DdSub* subs; DdDict dic; uint8_t* out; uint32_t* err; };
__global__ void k_dd_subs_jobs(const DdJob* __restrict__ jobs, const uint32_t* __restrict__ first, uint32_t njobs, uint32_t nsub) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nsub) return;
    uint32_t lo = 0, hi = njobs;
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (first[mid] <= t) lo = mid; else hi = mid; }
    const DdJob J = jobs[lo];
    dd_sub_body(J.D, J.blocks, J.subs[t - first[lo]], J.dic, J.out, J.err);
}
