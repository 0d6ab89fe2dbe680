// cr_lz77.cuh -- LZ77 match finding and parsing for the `comprox` front-end (src/roxmain), all positions in parallel.
//
// Replaces matcher_init / match / matcher_lookup (src/roxmain/cr-matcher.c:89-340) and the look-ahead thread
// (src/roxmain/cr-coder.c:116-142).
//   * matcher_init threads every position p <= len-256 onto the hash chain of (hash1 % 20, hash2 % bucketsize2): next[p] is
//     the largest q < p with the same pair.  A stable sort of the positions by that pair puts every chain in consecutive
//     ranks, so "walk the chain" is "read the sorted array backwards" (k_x_keys + cr_sort_pairs + k_x_match).
//   * match() and the lazy probes at pos+1..pos+6 are pure functions of the data, so N(p) = the lazy-parse candidate of
//     position p is computed for all p at once.  The short cache (m_short_cache) is written for every position in order,
//     so its content at look-up time is "the most recent q < p with the same 6-byte hash" (k_x_short).
//   * The ONE parse-dependent scalar is m_last_match (the distance of the previous match): it decides whether a repeat
//     match replaces N(p).  The parse is therefore the orbit of (pos, last) -> (pos + len, last'), resolved as a fixed
//     point: guess G[p] = "last on arrival at p", decide all positions, resolve the skip chain (cr_chain.cuh), propagate
//     the true `last` along the resolved path, and repeat until no visited position disagrees with its guess.  A
//     self-consistent path that starts at (0, 0) IS the serial parse (induction over tokens).  A serial kernel is the
//     fallback if the iteration cap is hit.
#pragma once
#include "cr_common.cuh"
#include "cr_rolz.cuh"
#include "cr_chain.cuh"
#include "cr_ppm.cuh"

#define X_MAXLEN    255u
#define X_MIN_NEAR  6u
#define X_NONE      0xFFFFFFFFu
#define X_LOOKAHEAD 1024u     // look-ups only while pos + 1024 < size (src/roxmain/cr-coder.c:125)
#define X_NMODEL    8         // order-0 models: 0 = len, 1 = spos, 2..7 = pos_models[0..5] (src/roxmain/cr-coder.c:55-60)

CR_HD uint32_t x_match_min(uint32_t size) { return 10 + (size > 16777216u); }                  // cr-coder.c:178
CR_HD uint32_t x_bucket2(uint32_t size) { return 20 + size / 25; }                              // cr-matcher.c:91
CR_HD uint32_t x_hashn(const uint8_t* s, uint32_t k) { uint32_t h = 0; for (uint32_t i = 0; i < k; i++) h = (h * 123456791u) ^ s[i]; return h; }   // cr-matcher.c:44-52
CR_HD uint32_t x_entries(uint32_t size) { return size > X_MAXLEN ? size - X_MAXLEN : 0; }      // positions on a chain: p + 255 < size
CR_HD uint32_t x_inc(uint32_t model) { return model == 0 ? 30u : model == 1 ? 1u : 1u << (2 * (model - 2)); }   // 30, 1, M_inc_factor(i)
CR_HD uint32_t x_lcp(const uint8_t* a, const uint8_t* b, uint32_t from, uint32_t cap) { uint32_t n = from; while (n < cap && a[n] == b[n]) n++; return n; }

// PPM context in front of position t (all bytes of all blocks of the chain update it, cr-coder.c:259-261)
CR_HD uint32_t x_ctx_at(const uint8_t* d, uint32_t t, uint32_t cin) {
    if (t >= 4) return (uint32_t)d[t - 4] << 24 | (uint32_t)d[t - 3] << 16 | (uint32_t)d[t - 2] << 8 | d[t - 1];
    uint32_t c = cin;
    for (uint32_t i = 0; i < t; i++) c = c << 8 | d[i];
    return c;
}
__global__ void k_x_finish_blocks(const uint8_t* __restrict__ D, LzBlock* __restrict__ blocks, uint32_t nb, const uint8_t* __restrict__ esc1,
                                  uint32_t ctx_in, uint32_t* __restrict__ ctx_out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    uint32_t c = ctx_in;
    for (uint32_t b = 0; b < nb; b++) {
        blocks[b].esc = esc1[b];
        blocks[b].cin = c;
        c = x_ctx_at(D + blocks[b].off, blocks[b].walk, c);
    }
    *ctx_out = c;
}

// ------------------------------------------------------------------ chains
// key = block << kbits | (hash1 % 20) * bucket2 + hash2 % bucket2; value = position.  Entry index = B.eoff + p.
__global__ void k_x_keys(const uint8_t* __restrict__ D, const LzBlock* __restrict__ blocks, int kbits, uint32_t* __restrict__ key, uint32_t* __restrict__ val) {
    const LzBlock B = blocks[blockIdx.y];
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= x_entries(B.size)) return;
    const uint8_t* s = D + B.off + p;
    const uint32_t b2 = x_bucket2(B.size);
    key[B.eoff + p] = (uint32_t)blockIdx.y << kbits | (((uint32_t)s[0] + s[1]) % 20u * b2 + x_hashn(s, x_match_min(B.size)) % b2);
    val[B.eoff + p] = p;
}
__global__ void k_x_ranks(const uint32_t* __restrict__ K, const uint32_t* __restrict__ V, uint32_t n, const LzBlock* __restrict__ blocks, int kbits, uint32_t* __restrict__ rank_of) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    rank_of[blocks[K[r] >> kbits].eoff + V[r]] = r;
}

struct XRet { uint32_t pos, len; };
// match(), cr-matcher.c:156-194.  r = rank of `pos` in the sorted order; the chain is K/V[r-1], K/V[r-2], ... while the key is equal.
CR_D XRet x_match(const uint8_t* __restrict__ d, const uint32_t* __restrict__ K, const uint32_t* __restrict__ V, uint32_t r, uint32_t pos,
                  uint32_t min, uint32_t limit, uint32_t lazy) {
    XRet ret; ret.pos = 0; ret.len = min - 1;
    const uint32_t key = K[r];
    for (uint32_t i = 0; i < limit && i < r; i++) {
        const uint32_t idx = r - 1 - i;
        if (K[idx] != key) break;
        const uint32_t node = V[idx];
        // the reference extends from ret.len and then memcmp()s the first ret.len bytes: accept iff the common prefix reaches beyond ret.len + price
        const uint32_t dist = pos - node, best = pos - ret.pos;
        const uint32_t price = (uint32_t)(dist / 1048576u > best) + (uint32_t)(dist / 4096u > best) + (uint32_t)(dist / 64u > best);
        const uint32_t need = ret.len + price;                    // new_len must exceed this
        if (need >= X_MAXLEN) continue;
        if (d[node + need] != d[pos + need]) continue;            // new_len > need requires equality at offset `need` (cheap reject)
        const uint32_t pre = x_lcp(d + node, d + pos, 0, need);
        if (pre < need) continue;                                 // prefix differs (memcmp != 0) or the run from ret.len stops early
        ret.pos = node; ret.len = x_lcp(d + node, d + pos, need + 1, X_MAXLEN);
        if ((lazy && lazy < ret.pos) || ret.len == X_MAXLEN) return ret;
    }
    if (ret.len < min) { ret.pos = X_NONE; ret.len = 1; }
    return ret;
}
// N(p): the lazy-parse candidate of matcher_lookup (cr-matcher.c:296-312), before the last-match / short-cache rules.
__global__ void k_x_match(const uint8_t* __restrict__ D, const LzBlock* __restrict__ blocks, const uint32_t* __restrict__ K, const uint32_t* __restrict__ V,
                          const uint32_t* __restrict__ rank_of, uint32_t limit, uint32_t* __restrict__ npos, uint8_t* __restrict__ nlen) {
    const LzBlock B = blocks[blockIdx.y];
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B.size) return;
    XRet ret; ret.pos = X_NONE; ret.len = 1;
    if (p + X_LOOKAHEAD < B.size) {
        const uint8_t* d = D + B.off;
        const uint32_t* ro = rank_of + B.eoff;
        const uint32_t mm = x_match_min(B.size);
        ret = x_match(d, K, V, ro[p], p, mm, limit, 0);
        if (ret.len >= mm) {
            const XRet t2 = x_match(d, K, V, ro[p + 1], p + 1, ret.len + 1, limit / 4, 1);
            if (t2.len > ret.len + (uint32_t)(t2.pos < ret.pos) ||
                x_match(d, K, V, ro[p + 2], p + 2, ret.len + 1, limit / 8, 1).len > 1 ||
                x_match(d, K, V, ro[p + 3], p + 3, ret.len + 2, limit / 8, 1).len > 1 ||
                x_match(d, K, V, ro[p + 4], p + 4, ret.len + 2, limit / 8, 1).len > 1 ||
                x_match(d, K, V, ro[p + 5], p + 5, ret.len + 2, limit / 8, 1).len > 1 ||
                x_match(d, K, V, ro[p + 6], p + 6, ret.len + 3, limit / 8, 1).len > 1) { ret.pos = X_NONE; ret.len = 1; }
        }
    }
    npos[B.off + p] = ret.pos; nlen[B.off + p] = (uint8_t)ret.len;
}
// -f: flexible parsing (cr-matcher.c:247-295).  Needs match(q) for q = p .. p + len; those are the unconditioned
// match() results M0, computed for every position by k_x_match0 first.
__global__ void k_x_match0(const uint8_t* __restrict__ D, const LzBlock* __restrict__ blocks, const uint32_t* __restrict__ K, const uint32_t* __restrict__ V,
                           const uint32_t* __restrict__ rank_of, uint32_t limit, uint32_t* __restrict__ mpos, uint8_t* __restrict__ mlen) {
    const LzBlock B = blocks[blockIdx.y];
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B.size) return;
    XRet ret; ret.pos = X_NONE; ret.len = 1;
    if (p < x_entries(B.size)) ret = x_match(D + B.off, K, V, rank_of[B.eoff + p], p, x_match_min(B.size), limit, 0);
    mpos[B.off + p] = ret.pos; mlen[B.off + p] = (uint8_t)ret.len;
}
CR_HD int x_log2(uint32_t x) { int l = -1; while (x) { l++; x >>= 1; } return l; }             // fast_log2, cr-matcher.c:211-228
__global__ void k_x_flex(const LzBlock* __restrict__ blocks, const uint32_t* __restrict__ mpos, const uint8_t* __restrict__ mlen,
                         uint32_t* __restrict__ npos, uint8_t* __restrict__ nlen) {
    const LzBlock B = blocks[blockIdx.y];
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B.size) return;
    XRet ret; ret.pos = X_NONE; ret.len = 1;
    if (p + X_LOOKAHEAD < B.size) {
        const uint32_t mm = x_match_min(B.size);
        const uint32_t* mp = mpos + B.off; const uint8_t* ml = mlen + B.off;
        ret.pos = mp[p]; ret.len = ml[p];
        if (ret.len >= mm) {
#define XPRICE(at, i, l) ((l) >= mm ? (uint32_t)(((l) - 1) * 3) - (uint32_t)(x_log2((at) - (i)) * 4 / 5) : 9u)
            const uint32_t n = ret.len, p0 = ret.pos;
            uint32_t maxprice = XPRICE(p, p0, n) + XPRICE(p, mp[p + n], (uint32_t)ml[p + n]);
            for (uint32_t i = n - 1; i >= 1; i--) {
                const uint32_t pr = XPRICE(p, p0, i) + XPRICE(p, mp[p + i], (uint32_t)ml[p + i]);
                if (maxprice < pr) { ret.len = i; maxprice = pr; }
            }
#undef XPRICE
            if (ret.len < mm) { ret.pos = X_NONE; ret.len = 1; }
        }
    }
    npos[B.off + p] = ret.pos; nlen[B.off + p] = (uint8_t)ret.len;
}

// ------------------------------------------------------------------ short cache (cr-matcher.c:196-209,318-330)
__global__ void k_x_short_hash(const uint8_t* __restrict__ D, const LzBlock* __restrict__ blocks, uint16_t* __restrict__ h16) {
    const LzBlock B = blocks[blockIdx.y];
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B.size) return;
    h16[B.off + p] = p + X_MIN_NEAR <= B.size ? (uint16_t)(x_hashn(D + B.off + p, X_MIN_NEAR) & 0xFFFF) : 0;
}
// sdist = distance to the cache entry if it lies within 255 positions (else 0), slen = common prefix with it.
// The entry is the most recent earlier position with the same hash; an entry never written still holds position 0.
__global__ void k_x_short(const uint8_t* __restrict__ D, const LzBlock* __restrict__ blocks, const uint16_t* __restrict__ h16,
                          uint8_t* __restrict__ sdist, uint8_t* __restrict__ slen) {
    const LzBlock B = blocks[blockIdx.y];
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B.size) return;
    uint32_t dist = 0, len = 0;
    if (p + X_LOOKAHEAD < B.size && p > 0) {
        const uint16_t* h = h16 + B.off;
        const uint16_t me = h[p];
        const uint32_t lo = p > 255 ? p - 255 : 0;
        uint32_t q = p;
        while (q > lo && h[q - 1] != me) q--;
        if (q > lo) dist = p - (q - 1);
        else if (p < 256) dist = p;                                  // nothing in [0, p): the slot still holds its initial 0
        if (dist) len = x_lcp(D + B.off + p - dist, D + B.off + p, 0, X_MAXLEN);
    }
    sdist[B.off + p] = (uint8_t)dist; slen[B.off + p] = (uint8_t)len;
}

// ------------------------------------------------------------------ the decision at one parse point (cr-matcher.c:314-339)
struct XTables { const uint32_t* npos; const uint8_t* nlen; const uint8_t* sdist; const uint8_t* slen; };
// returns the token length; dist = 0 for a literal.  `last` = m_last_match on arrival.
CR_D uint32_t x_decide(const uint8_t* __restrict__ d, uint32_t size, uint32_t p, uint32_t last, uint32_t np, uint32_t nl, uint32_t sd, uint32_t sl, uint32_t& dist) {
    dist = 0;
    if (p + X_LOOKAHEAD >= size) return 1;
    const uint32_t mm = x_match_min(size);
    uint32_t rpos = np, rlen = nl;
    if (np != X_NONE) {
        uint32_t t1 = 0;
        if (last != 0 && last <= p) t1 = x_lcp(d + p - last, d + p, 0, X_MAXLEN);
        if (rlen < t1 + 3 + (uint32_t)(np + 64 < p) + (uint32_t)(np + 4096 < p) + (uint32_t)(np + 1048576 < p)) { rpos = p - last; rlen = t1; }
    }
    if (rlen < X_MIN_NEAR) { rpos = p - sd; rlen = sd ? sl : 0; if (!sd) rpos = X_NONE - 1024; }     // outside the window: no match
    if (rlen < X_MIN_NEAR || (rlen < mm && rpos + 256 <= p)) return 1;
    dist = p - rpos;
    return rlen;
}
// one decision per position under the guess G[p]; positions without a normal match do not depend on the guess
__global__ void k_x_decide(const uint8_t* __restrict__ D, const LzBlock* __restrict__ blocks, XTables T, const uint32_t* __restrict__ G, int first_pass,
                           uint8_t* __restrict__ span, uint32_t* __restrict__ tdist) {
    const LzBlock B = blocks[blockIdx.y];
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B.size) return;
    const size_t g = B.off + p;
    const uint32_t np = T.npos[g];
    if (!first_pass && np == X_NONE) return;
    uint32_t dist;
    const uint32_t len = x_decide(D + B.off, B.size, p, first_pass ? 0u : G[g], np, T.nlen[g], T.sdist[g], T.slen[g], dist);
    span[g] = (uint8_t)len; tdist[g] = dist;
}

// ------------------------------------------------------------------ propagation of `last` along the resolved path
// Walk A: distance of the last match token of every chunk (0 = the chunk holds no match token)
struct XLastOfChunk {
    typedef struct { uint32_t last; } State;
    const LzBlock* blocks; const uint32_t* tdist; uint32_t* chunk_last;
    CR_D State begin(uint32_t, uint32_t) const { State s = { 0 }; return s; }
    CR_D void visit(State& s, uint32_t b, uint32_t t, uint32_t) const { const uint32_t dd = tdist[blocks[b].off + t]; if (dd) s.last = dd; }
    CR_D void end(State& s, uint32_t c, uint32_t) const { chunk_last[c] = s.last; }
};
// chunk_key[c] = c + 1 if the chunk holds a match token; an inclusive prefix maximum then names the last such chunk
__global__ void k_x_chunk_keys(const uint32_t* __restrict__ chunk_last, uint32_t nchunk, uint32_t* __restrict__ key) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < nchunk) key[c] = chunk_last[c] ? c + 1 : 0;
}
__global__ void k_x_chunk_in(const uint32_t* __restrict__ chunk_last, const uint32_t* __restrict__ pmax, const ChainSeg* __restrict__ segs, uint32_t nseg, uint32_t nchunk,
                             uint32_t* __restrict__ last_in) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunk) return;
    const ChainSeg sg = segs[cr_find_seg(segs, nseg, c)];
    const uint32_t prev = c ? pmax[c - 1] : 0;                       // 1 + index of the last chunk before c with a match
    last_in[c] = (prev && prev - 1 >= sg.chunk0) ? chunk_last[prev - 1] : 0;   // m_last_match = 0 at the start of a block (cr-matcher.c:107)
}
// Walk B: compare the guesses with the truth along the path, and refresh them
struct XCheck {
    typedef struct { uint32_t last, bad; } State;
    const LzBlock* blocks; const uint32_t* tdist; const uint32_t* npos; const uint32_t* last_in; uint32_t* G; uint32_t* mismatches;
    CR_D State begin(uint32_t c, uint32_t) const { State s = { last_in[c], 0 }; return s; }
    CR_D void visit(State& s, uint32_t b, uint32_t t, uint32_t len) const {
        const size_t g = blocks[b].off + t;
        if (npos[g] != X_NONE && G[g] != s.last) s.bad++;            // only positions with a normal match look at `last`
        for (uint32_t i = 0; i < len; i++) G[g + i] = s.last;        // interior positions: best guess for a path that lands there later
        const uint32_t dd = tdist[g];
        if (dd) s.last = dd;
    }
    CR_D void end(State& s, uint32_t, uint32_t) const { if (s.bad) atomicAdd(mismatches, s.bad); }
};
// serial fallback: one thread per block runs the reference loop over the precomputed tables
__global__ void k_x_parse_serial(const uint8_t* __restrict__ D, const LzBlock* __restrict__ blocks, uint32_t nb, XTables T, uint8_t* __restrict__ span, uint32_t* __restrict__ tdist) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const LzBlock B = blocks[b];
    uint32_t last = 0;
    for (uint32_t p = 0; p < B.size;) {
        const size_t g = B.off + p;
        uint32_t dist;
        const uint32_t len = x_decide(D + B.off, B.size, p, last, T.npos[g], T.nlen[g], T.sdist[g], T.slen[g], dist);
        span[g] = (uint8_t)len; tdist[g] = dist;
        if (dist) last = dist;
        p += len;
    }
}

// ------------------------------------------------------------------ the coder's own `last_match` (cr-coder.c:219-243)
// lzencode keeps a second copy of the last distance, and it is NOT the matcher's: after a token coded as "same as last"
// (distance symbol 0) the coder stores 0, so a repeat is never coded twice in a row.  Along the match tokens
//     c' = (dist == c) ? 0 : dist
// and a chunk maps its incoming c to (c == d1 ? a : b) with d1 = its first match distance.  Walk C collects (d1, a, b)
// per chunk, one thread per block composes them in order, and the count / emit walks start from the result.
struct XCoderSum {
    typedef struct { uint32_t d1, a, b; } State;
    const LzBlock* blocks; const uint32_t* tdist; uint32_t* sum3;      // sum3[3 * chunk + {0,1,2}]
    CR_D State begin(uint32_t, uint32_t) const { State s = { 0, 0, 0 }; return s; }
    CR_D void visit(State& s, uint32_t b, uint32_t t, uint32_t) const {
        const uint32_t dd = tdist[blocks[b].off + t];
        if (!dd) return;
        if (!s.d1) { s.d1 = dd; s.a = 0; s.b = dd; }
        else { s.a = dd == s.a ? 0 : dd; s.b = dd == s.b ? 0 : dd; }
    }
    CR_D void end(State& s, uint32_t c, uint32_t) const { sum3[3 * (size_t)c] = s.d1; sum3[3 * (size_t)c + 1] = s.a; sum3[3 * (size_t)c + 2] = s.b; }
};
__global__ void k_x_coder_in(const uint32_t* __restrict__ sum3, const ChainSeg* __restrict__ segs, uint32_t nseg, uint32_t* __restrict__ coder_in) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    const ChainSeg sg = segs[s];
    uint32_t cur = 0;                                                  // last_match = 0 at the start of lzencode (cr-coder.c:156)
    for (uint32_t k = 0; k < sg.nchunk; k++) {
        const size_t c = sg.chunk0 + k;
        coder_in[c] = cur;
        const uint32_t d1 = sum3[3 * c];
        if (d1) cur = cur == d1 ? sum3[3 * c + 1] : sum3[3 * c + 2];
    }
}

// ------------------------------------------------------------------ tokens -> PPM events + order-0 model symbols (cr-coder.c:213-262)
#define X_NCOUNT 11      // per chunk: events, len symbols, spos symbols, pos tokens, pos symbols, pos_models[0..5] symbols
// number of pos_models symbols of a coded distance, and (optionally) the symbols themselves
CR_D uint32_t x_pos_symbols(uint32_t dist, uint8_t* sym, uint8_t* model) {
    uint32_t j = dist * 8, i = 0, n = 0;
    while (j >= 128 && i < 2) { if (sym) { sym[n] = (uint8_t)(j % 128 + 128); model[n] = (uint8_t)i; } n++; i++; j /= 128; }
    if (i >= 2) while (j >= 64 && i < 5) { if (sym) { sym[n] = (uint8_t)(j % 64 + 64); model[n] = (uint8_t)i; } n++; i++; j /= 64; }
    if (sym) { sym[n] = (uint8_t)j; model[n] = (uint8_t)i; }
    return n + 1;
}
struct XCount {
    typedef struct { uint32_t last, c[X_NCOUNT]; } State;
    const uint8_t* D; const LzBlock* blocks; const uint32_t* tdist; const uint32_t* last_in; uint32_t* cnt; uint32_t stride;   // cnt[k * stride + chunk]
    CR_D State begin(uint32_t c, uint32_t) const { State s; s.last = last_in[c]; for (int k = 0; k < X_NCOUNT; k++) s.c[k] = 0; return s; }
    CR_D void visit(State& s, uint32_t b, uint32_t t, uint32_t len) const {
        const LzBlock B = blocks[b];
        const uint32_t dd = tdist[B.off + t];
        s.c[0]++;
        if (dd) {
            s.c[1]++;
            const uint32_t coded = dd == s.last ? 0 : dd;        // s.last = the CODER's last_match here
            if (len < x_match_min(B.size)) s.c[2]++;
            else {
                uint8_t sym[6], model[6];
                const uint32_t n = x_pos_symbols(coded, sym, model);
                s.c[3]++; s.c[4] += n;
                for (uint32_t k = 0; k < n; k++) s.c[5 + model[k]]++;
            }
            s.last = coded;
        } else if (D[B.off + t] == B.esc) s.c[1]++;
    }
    CR_D void end(State& s, uint32_t c, uint32_t) const { for (int k = 0; k < X_NCOUNT; k++) cnt[(size_t)k * stride + c] = s.c[k]; }
};
// side array layout: [all spos symbols][all pos symbols][all len symbols] of the window, each in token order
struct XEmit {
    typedef struct { uint32_t last, ev, nlen, nspos, npossym, nm[6]; } State;
    const uint8_t* D; const LzBlock* blocks; const uint32_t* tdist; const uint32_t* last_in; const uint32_t* scan; uint32_t stride;
    uint32_t base_pos, base_len;                // offsets of the pos / len sections in the side array (spos starts at 0)
    uint32_t* ev_ctx; uint8_t* ev_sym;
    uint8_t* msym[X_NMODEL]; uint32_t* mpos[X_NMODEL];      // per model: k-th use -> symbol, index in the side array
    CR_D State begin(uint32_t c, uint32_t) const {
        State s; s.last = last_in[c]; s.ev = scan[c]; s.nlen = scan[(size_t)stride + c]; s.nspos = scan[(size_t)2 * stride + c]; s.npossym = scan[(size_t)4 * stride + c];
        for (int k = 0; k < 6; k++) s.nm[k] = scan[(size_t)(5 + k) * stride + c];
        return s;
    }
    CR_D void visit(State& s, uint32_t b, uint32_t t, uint32_t len) const {
        const LzBlock B = blocks[b];
        const uint8_t* d = D + B.off;
        const uint32_t dd = tdist[B.off + t];
        ev_ctx[s.ev] = x_ctx_at(d, t, B.cin);
        if (dd) {
            ev_sym[s.ev] = B.esc;
            msym[0][s.nlen] = (uint8_t)len; mpos[0][s.nlen] = base_len + s.nlen; s.nlen++;
            const uint32_t coded = dd == s.last ? 0 : dd;
            if (len < x_match_min(B.size)) { msym[1][s.nspos] = (uint8_t)coded; mpos[1][s.nspos] = s.nspos; s.nspos++; }
            else {
                uint8_t sym[6], model[6];
                const uint32_t n = x_pos_symbols(coded, sym, model);
                for (uint32_t k = 0; k < n; k++) {
                    const uint32_t m = model[k];
                    msym[2 + m][s.nm[m]] = sym[k]; mpos[2 + m][s.nm[m]] = base_pos + s.npossym; s.nm[m]++; s.npossym++;
                }
            }
            s.last = coded;
        } else {
            ev_sym[s.ev] = d[t];
            if (d[t] == B.esc) { msym[0][s.nlen] = 0; mpos[0][s.nlen] = base_len + s.nlen; s.nlen++; }
        }
        s.ev++;
    }
    CR_D void end(State&, uint32_t, uint32_t) const {}
};

// ------------------------------------------------------------------ order-0 models with per-model increments (cr-model.c:55-88)
struct SideJobs { const uint8_t* sym[X_NMODEL]; const uint32_t* pos[X_NMODEL]; uint32_t n[X_NMODEL]; uint32_t inc[X_NMODEL]; uint16_t* state[X_NMODEL]; uint32_t nmodel; };
// scalar form: one thread per model (the models are independent chains over the whole window)
__global__ void k_side_models_jobs(SideJobs J, uint64_t* __restrict__ TS) {
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= J.nmodel) return;
    uint16_t* f = J.state[m];
    uint32_t total = 0;
    for (int i = 0; i < 256; i++) total += f[i];
    const uint32_t inc = J.inc[m];
    for (uint32_t i = 0; i < J.n[m]; i++) {
        const uint32_t s = J.sym[m][i];
        uint32_t cum = 0;
        for (uint32_t j = 0; j < s; j++) cum += f[j];
        TS[J.pos[m][i]] = ppm_pack(cum, f[s], total, 0);
        f[s] = (uint16_t)(f[s] + inc); total += inc;
        if (total > 32000) { total = 0; for (int j = 0; j < 256; j++) { f[j] = (uint16_t)((f[j] + 1) / 2); total += f[j]; } }
    }
}
__global__ void k_x_reset_models(uint16_t* __restrict__ m0) {          // reset_models, cr-coder.c:88-103; slot k*256 = model k
    const uint32_t k = threadIdx.x;
    if (k >= 256) return;
    m0[0 * 256 + k] = (k >= X_MIN_NEAR || k == 0) ? 1 : 0;            // len_model
    m0[1 * 256 + k] = 1;                                               // spos_model
    m0[2 * 256 + k] = (k % 8 == 0) ? 1 : 0;                            // pos_models[0]
    m0[3 * 256 + k] = 1;                                               // pos_models[1]
    for (int i = 2; i < 5; i++) m0[(2 + i) * 256 + k] = k < 128 ? 1 : 0;
    m0[7 * 256 + k] = 1;                                               // pos_models[5]
}
