// cr_rolz.cuh -- ROLZ match finding for comprolz, all positions in parallel.
//
// Replaces matcher_update / match / matcher_lookup (src/rolzmain/cr-matcher.c:65-197) and the look-ahead
// thread (src/rolzmain/cr-coder.c:109-137).  The reference keeps, per 18-bit context hash, a ring of the 64
// most recent positions, and per preceding byte a list of the 16 most recent positions.  Both tables are
// updated for EVERY position in order and never depend on parse decisions, so
//   "the ring of bucket k as of time t"  ==  "the last 64 positions q < t with bucket(q) == k".
// We therefore stable-sort the positions of a block by bucket; the candidates of position p are simply the
// up-to-64 entries in front of it in the sorted order.  The lazy rule needs match(p) evaluated against the
// table as of time p-1 .. p-4 as well; those differ from "as of p" only when one of the 4 previous positions
// shares p's bucket, i.e. by skipping a prefix of the candidate list.
#pragma once
#include "cr_common.cuh"

#define RZ_BUCKET_BITS 18
#define RZ_BUCKETS     262144u
#define RZ_WAYS        64u
#define RZ_SHORT       16u
#define RZ_MINLEN      5u
#define RZ_MAXLEN      255u
#define RZ_TAG_SHIFT   24
#define RZ_KEY_MASK    0x00FFFFFFu
#define RZ_MAX_BLOCKS  64u     // blocks per window: 6 block bits + 18 bucket bits below the tag
#define RZ_LOOKAHEAD   1024u   // look-ups only while pos + 1024 < size (src/rolzmain/cr-coder.c:118)

struct LzBlock {
    uint64_t off;       // offset of the block's byte 0 in the window's concatenated dictionary-coded data
    uint32_t size;      // dictionary-coded size
    uint32_t eoff;      // ROLZ: index of position 16 in the window's entry arrays
    uint32_t cin;       // PPM context on entry to the block (carried across blocks)
    uint8_t  esc;       // rarest byte of the block
    uint8_t  ctx4;      // size >= 4 MiB -> 4-byte context hash (src/rolzmain/cr-coder.c:162)
    uint8_t  pad[2];
    uint32_t walk;      // bytes the coder loop consumed: == size, or the position at which the reference gave up on the block
    uint32_t pad2;      //   ("cannot compress", src/rolzmain/cr-coder.c:231-233): models and context only saw tokens before it
};

CR_HD uint32_t rz_hash(const uint8_t* x, int ctx4) {           // src/rolzmain/cr-matcher.c:38-42
    uint32_t h = x[0] * 1313131u + x[-1] * 13131u + x[-2] * 131u;
    if (ctx4) h += x[-3];
    return h & (RZ_BUCKETS - 1);
}

// sort keys for every position p >= 16 of every block.  bucket(16) = 0 for both tables (m_context and
// m_short_context still hold their initial 0 when position 16 is inserted, cr-matcher.c:52-54,66-79).
__global__ void k_rolz_keys(const uint8_t* __restrict__ D, const LzBlock* __restrict__ blocks,
                            uint32_t* __restrict__ kmain, uint32_t* __restrict__ kshort, uint32_t* __restrict__ val) {
    const LzBlock B = blocks[blockIdx.y];
    uint32_t p = 16 + blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B.size) return;
    const uint8_t* d = D + B.off;
    uint32_t e = B.eoff + p - 16;
    uint32_t bucket = p == 16 ? 0 : rz_hash(d + p - 1, B.ctx4);
    uint32_t sb = p == 16 ? 0 : d[p - 1];
    // bits 24..31 carry the byte AT p (the reference's m_hash, cr-matcher.h:51): candidates whose first byte differs are
    // rejected from the key alone, without touching the data.  Needs block index < 64 (main) -- callers split windows.
    kmain[e] = (uint32_t)d[p] << RZ_TAG_SHIFT | (blockIdx.y << RZ_BUCKET_BITS) | bucket;
    kshort[e] = (uint32_t)d[p] << RZ_TAG_SHIFT | (blockIdx.y << 8) | sb;
    val[e] = p;
}

// Main-table search.  One thread per sorted rank.  Output M[e] = len | idx<<8 (0 = no match >= 5) for the table as of
// time p.  The lazy rule also needs the result as of time p-1..p-4; those differ only when one of the four previous
// positions shares p's bucket, so they are stored (M[v*n + e], v = 1..4) only then and flagged with bit 15 of M[e].
// The common case therefore scatters 2 bytes per position into one block's 32 MB slice, which stays in L2.
#define RZ_VARIANTS 0x8000u
__global__ void k_rolz_match_main(const uint8_t* __restrict__ D, const LzBlock* __restrict__ blocks,
                                  const uint32_t* __restrict__ K, const uint32_t* __restrict__ V, uint32_t n,
                                  uint16_t* __restrict__ M, uint32_t* __restrict__ rank_of) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint32_t key = K[r] & RZ_KEY_MASK, p = V[r];
    const LzBlock B = blocks[key >> RZ_BUCKET_BITS];
    if (rank_of) rank_of[B.eoff + p - 16] = r;               // -f only: k_rolz_flex looks positions up in the sorted order
    if (p + (RZ_LOOKAHEAD - 4) >= B.size) return;            // never looked up (lazy look-ahead reaches pos+4)
    const uint8_t* d = D + B.off;
    const uint8_t* dp = d + p;
    const uint32_t e = B.eoff + p - 16;
    const uint32_t first = K[r] >> RZ_TAG_SHIFT;

    const bool simple = (r == 0) || (K[r - 1] & RZ_KEY_MASK) != key || V[r - 1] + 4 < p;
    if (simple) {
        // no candidate lies in [p-4, p): all five variants see the same list
        uint32_t best = RZ_MINLEN - 1, idx = 0;
        for (uint32_t c = 0; c < RZ_WAYS && c < r && best < RZ_MAXLEN; c++) {
            const uint32_t kc = K[r - 1 - c];
            if ((kc & RZ_KEY_MASK) != key) break;
            if ((kc >> RZ_TAG_SHIFT) != first) continue;
            const uint8_t* dq = d + V[r - 1 - c];
            if (dq[best] != dp[best]) continue;
            uint32_t l = cr_cpl(dp, dq, RZ_MAXLEN);
            if (l > best) { best = l; idx = c; }
        }
        M[e] = best >= RZ_MINLEN ? (uint16_t)(best | idx << 8) : (uint16_t)0;
        return;
    }
    uint32_t best[5], idx[5], skip[5];
#pragma unroll
    for (int v = 0; v < 5; v++) { best[v] = RZ_MINLEN - 1; idx[v] = 0; skip[v] = 0; }
    for (uint32_t c = 0; c < RZ_WAYS + 4 && c < r; c++) {
        const uint32_t kc = K[r - 1 - c];
        if ((kc & RZ_KEY_MASK) != key) break;
        uint32_t q = V[r - 1 - c];
        bool use[5], any = false;
#pragma unroll
        for (int v = 0; v < 5; v++) {
            bool excluded = q + v >= p;                       // inserted at or after time p - v
            if (excluded) skip[v]++;
            use[v] = !excluded && (c - skip[v]) < RZ_WAYS && best[v] < RZ_MAXLEN;
            any |= use[v];
        }
        if (!any) { if (c >= skip[4] + RZ_WAYS) break; continue; }
        if ((kc >> RZ_TAG_SHIFT) != first) continue;
        uint32_t l = cr_cpl(dp, d + q, RZ_MAXLEN);
#pragma unroll
        for (int v = 0; v < 5; v++) if (use[v] && l > best[v]) { best[v] = l; idx[v] = c - skip[v]; }
    }
    M[e] = (uint16_t)((best[0] >= RZ_MINLEN ? (best[0] | idx[0] << 8) : 0u) | RZ_VARIANTS);
#pragma unroll
    for (int v = 1; v < 5; v++) M[(size_t)v * n + e] = best[v] >= RZ_MINLEN ? (uint16_t)(best[v] | idx[v] << 8) : (uint16_t)0;
}

#ifndef CRGPU_SIM
// Main-table search, second form (default on the GPU; same results as k_rolz_match_main, which stays as the definition the CPU
// kernel-logic simulation runs).  What changed, and why (ncu on x86 / BMP data: the first form moves 6x its algorithmic bytes, nearly
// all of them 32-byte sectors fetched for one byte of a candidate that then fails, profiles/round2_summary.md section 6):
//   * a candidate is only worth touching if its first FIVE bytes equal ours (a match shorter than RZ_MINLEN never wins).  Byte 0 is the
//     key's tag; bytes 1..4 are gathered once per sorted entry while the CTA stages its window of the sorted list in shared memory,
//     so the filter runs on shared memory alone and the data is read only for candidates that are real matches;
//   * the window of keys (the CTA's 256 ranks and the 72 in front of them) is staged once instead of being re-read 64 times per rank;
//   * the common prefix is counted four bytes per step (aligned words + funnel shift).
#define RZM_TH   256
#define RZM_BACK 72u        // RZ_WAYS + 4 look-ahead variants, rounded up
CR_D uint32_t rz_ld32(const uint8_t* p) {               // four bytes at any alignment (reads up to 7 bytes past p: buffers carry slack)
    const uintptr_t a = (uintptr_t)p;
    const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
    return __funnelshift_r(w[0], w[1], (uint32_t)(a & 3) * 8);
}
CR_D uint32_t rz_cpl4(const uint8_t* a, const uint8_t* b, uint32_t from, uint32_t cap) {     // common prefix, known equal below `from`
    uint32_t j = from;
    while (j + 4 <= cap) {
        const uint32_t x = rz_ld32(a + j) ^ rz_ld32(b + j);
        if (x) return j + ((__ffs((int)x) - 1) >> 3);
        j += 4;
    }
    while (j < cap && a[j] == b[j]) j++;
    return j;
}
__global__ void __launch_bounds__(RZM_TH) k_rolz_match_main2(const uint8_t* __restrict__ D, const LzBlock* __restrict__ blocks,
                                                              const uint32_t* __restrict__ K, const uint32_t* __restrict__ V, uint32_t n,
                                                              uint16_t* __restrict__ M, uint32_t* __restrict__ rank_of) {
    __shared__ uint32_t sK[RZM_TH + RZM_BACK], sX[RZM_TH + RZM_BACK];
    const uint32_t base = blockIdx.x * RZM_TH;
    for (uint32_t i = threadIdx.x; i < RZM_TH + RZM_BACK; i += RZM_TH) {
        const long long g = (long long)base - RZM_BACK + i;
        uint32_t k = 0, x = 0;
        if (g >= 0 && g < (long long)n) {
            k = K[g];
            const LzBlock B = blocks[(k & RZ_KEY_MASK) >> RZ_BUCKET_BITS];
            const uint32_t q = V[g];
            if (q + 8 < B.size) x = rz_ld32(D + B.off + q + 1);          // a candidate of a looked-up position always has this much behind it
        }
        sK[i] = k; sX[i] = x;
    }
    __syncthreads();
    const uint32_t r = base + threadIdx.x;
    if (r >= n) return;
    const uint32_t me = threadIdx.x + RZM_BACK;                              // my slot; candidate c sits at me - 1 - c
    const uint32_t key = sK[me] & RZ_KEY_MASK, p = V[r];
    const LzBlock B = blocks[key >> RZ_BUCKET_BITS];
    if (rank_of) rank_of[B.eoff + p - 16] = r;
    if (p + (RZ_LOOKAHEAD - 4) >= B.size) return;
    const uint8_t* d = D + B.off;
    const uint8_t* dp = d + p;
    const uint32_t e = B.eoff + p - 16;
    const uint32_t first = sK[me] >> RZ_TAG_SHIFT, myx = sX[me];

    const bool simple = (r == 0) || (sK[me - 1] & RZ_KEY_MASK) != key || V[r - 1] + 4 < p;
    if (simple) {
        uint32_t best = RZ_MINLEN - 1, idx = 0;
        for (uint32_t c = 0; c < RZ_WAYS && c < r && best < RZ_MAXLEN; c++) {
            const uint32_t kc = sK[me - 1 - c];
            if ((kc & RZ_KEY_MASK) != key) break;
            if ((kc >> RZ_TAG_SHIFT) != first || sX[me - 1 - c] != myx) continue;
            const uint8_t* dq = d + V[r - 1 - c];
            if (best >= RZ_MINLEN && dq[best] != dp[best]) continue;
            const uint32_t l = rz_cpl4(dp, dq, RZ_MINLEN, RZ_MAXLEN);
            if (l > best) { best = l; idx = c; }
        }
        M[e] = best >= RZ_MINLEN ? (uint16_t)(best | idx << 8) : (uint16_t)0;
        return;
    }
    uint32_t best[5], idx[5], skip[5];
#pragma unroll
    for (int v = 0; v < 5; v++) { best[v] = RZ_MINLEN - 1; idx[v] = 0; skip[v] = 0; }
    for (uint32_t c = 0; c < RZ_WAYS + 4 && c < r; c++) {
        const uint32_t kc = sK[me - 1 - c];
        if ((kc & RZ_KEY_MASK) != key) break;
        const uint32_t q = V[r - 1 - c];
        bool use[5], any = false;
#pragma unroll
        for (int v = 0; v < 5; v++) {
            const bool excluded = q + v >= p;                       // inserted at or after time p - v
            if (excluded) skip[v]++;
            use[v] = !excluded && (c - skip[v]) < RZ_WAYS && best[v] < RZ_MAXLEN;
            any |= use[v];
        }
        if (!any) { if (c >= skip[4] + RZ_WAYS) break; continue; }
        if ((kc >> RZ_TAG_SHIFT) != first || sX[me - 1 - c] != myx) continue;
        const uint32_t l = rz_cpl4(dp, d + q, RZ_MINLEN, RZ_MAXLEN);
#pragma unroll
        for (int v = 0; v < 5; v++) if (use[v] && l > best[v]) { best[v] = l; idx[v] = c - skip[v]; }
    }
    M[e] = (uint16_t)((best[0] >= RZ_MINLEN ? (best[0] | idx[0] << 8) : 0u) | RZ_VARIANTS);
#pragma unroll
    for (int v = 1; v < 5; v++) M[(size_t)v * n + e] = best[v] >= RZ_MINLEN ? (uint16_t)(best[v] | idx[v] << 8) : (uint16_t)0;
}
#endif

// Order-1 ("short") table, consulted only where the main table found nothing (cr-matcher.c:165-179).
// Empty slots hold position 0 and ARE candidates (the table is zero-initialised, cr-matcher.c:53).
__global__ void k_rolz_match_short(const uint8_t* __restrict__ D, const LzBlock* __restrict__ blocks,
                                   const uint32_t* __restrict__ K, const uint32_t* __restrict__ V, uint32_t n,
                                   const uint16_t* __restrict__ M0, uint16_t* __restrict__ S) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint32_t key = K[r] & RZ_KEY_MASK, p = V[r], first = K[r] >> RZ_TAG_SHIFT;
    const LzBlock B = blocks[key >> 8];
    if (p + RZ_LOOKAHEAD >= B.size) return;
    const uint32_t e = B.eoff + p - 16;
    if ((M0[e] & ~RZ_VARIANTS) != 0) return;
    const uint8_t* d = D + B.off;
    const uint8_t* dp = d + p;
    uint32_t best = RZ_MINLEN - 1, idx = 0, c = 0;
    for (; c < RZ_SHORT && c < r; c++) {
        const uint32_t kc = K[r - 1 - c];
        if ((kc & RZ_KEY_MASK) != key) break;
        if ((kc >> RZ_TAG_SHIFT) != first) continue;               // a match of 5+ bytes needs an equal first byte
        const uint8_t* dq = d + V[r - 1 - c];
        if (dq[best] != dp[best]) continue;
        uint32_t l = cr_cpl(dp, dq, RZ_MAXLEN);
        if (l > best) { best = l; idx = c; }
    }
    if (c < RZ_SHORT) {                                        // remaining slots all hold position 0
        uint32_t l = cr_cpl(dp, d, RZ_MAXLEN);
        if (l > best) { best = l; idx = c; }
    }
    S[e] = best >= RZ_MINLEN ? (uint16_t)(best | (RZ_WAYS + idx) << 8) : (uint16_t)0;
}

CR_HD uint32_t rz_price(uint32_t m) {                          // M_price, cr-matcher.c:146-148
    return m ? ((m & 255) - 1) * 3 * RZ_WAYS - 3 * (m >> 8) : 9 * RZ_WAYS;
}

// Flexible parsing (-f, cr-matcher.c:142-162).  For a position t with a main-table match of length L the
// reference prices match(t+i) for i = 1..L against the table AS OF TIME t and may shorten the match.  In sorted
// order that is: the candidates of p = t+i, minus the leading ones that were inserted at or after t.
__global__ void k_rolz_flex(const uint8_t* __restrict__ D, const LzBlock* __restrict__ blocks, const uint32_t* __restrict__ K, const uint32_t* __restrict__ V,
                            const uint32_t* __restrict__ rank_of, const uint16_t* __restrict__ M0, uint8_t* __restrict__ flexlen) {
    const LzBlock B = blocks[blockIdx.y];
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 16 || t + RZ_LOOKAHEAD >= B.size) return;
    const uint32_t e = B.eoff + t - 16;
    const uint32_t m = M0[e] & ~RZ_VARIANTS;
    if (m == 0) return;
    const uint8_t* d = D + B.off;
    const uint32_t len0 = m & 255, idx0 = m >> 8;
    uint32_t prices[256];
    for (uint32_t i = 1; i <= len0; i++) {
        const uint32_t p = t + i, rp = rank_of[e + i], key = K[rp] & RZ_KEY_MASK;
        const uint8_t* dp = d + p;
        uint32_t best = RZ_MINLEN - 1, idx = 0, taken = 0;
        for (uint32_t c = 0; c < rp && taken < RZ_WAYS && best < RZ_MAXLEN; c++) {
            if ((K[rp - 1 - c] & RZ_KEY_MASK) != key) break;
            const uint32_t q = V[rp - 1 - c];
            if (q >= t) continue;                                  // not yet in the table at time t
            const uint8_t* dq = d + q;
            if (dq[0] == dp[0] && dq[best] == dp[best]) { const uint32_t l = cr_cpl(dp, dq, RZ_MAXLEN); if (l > best) { best = l; idx = taken; } }
            taken++;
        }
        prices[i] = rz_price(best >= RZ_MINLEN ? (best | idx << 8) : 0u);
    }
    uint32_t len = len0, maxprice = rz_price(m) + prices[len0];
    for (uint32_t i = len0 - 1; i >= 1; i--) {
        const uint32_t pi = (i >= RZ_MINLEN ? (i - 1) * 3 * RZ_WAYS - 3 * idx0 : 9 * RZ_WAYS) + prices[i];
        if (pi > maxprice) { len = i; maxprice = pi; }
    }
    flexlen[e] = (uint8_t)len;
}

// Token that the serial parse would emit IF it stood at position t: span[g] = length, tidx[g] = ROLZ index
// (0xFF = literal).  matcher_lookup's selection + lazy rule, cr-matcher.c:139-141,165-195.
__global__ void k_rolz_tokens(const LzBlock* __restrict__ blocks, const uint16_t* __restrict__ M, const uint16_t* __restrict__ S,
                              uint32_t n, uint8_t* __restrict__ span, uint8_t* __restrict__ tidx, const uint8_t* __restrict__ flexlen) {
    const LzBlock B = blocks[blockIdx.y];
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B.size) return;
    uint32_t len = 1, idx = 0xFF;
    if (t >= 16 && t + RZ_LOOKAHEAD < B.size) {
        uint32_t e = B.eoff + t - 16;
        uint32_t m = M[e] & ~RZ_VARIANTS;
        if (flexlen && m != 0) {                              // -f: shortened main match, no lazy rule (cr-matcher.c:143,186)
            const uint32_t fl = flexlen[e];
            if (fl >= RZ_MINLEN) { len = fl; idx = m >> 8; }
            m = 0;
        } else if (m == 0) m = S[e];
        if (m != 0) {
            bool keep = true;
            for (uint32_t i = 1; i < RZ_MINLEN; i++) {
                uint32_t m2 = M[e + i];                      // as of time t = (t+i) - i: stored separately only when it differs
                m2 = (m2 & RZ_VARIANTS) ? M[(size_t)i * n + e + i] : m2;
                if (rz_price(m2) > rz_price(m) + i * RZ_WAYS) { keep = false; break; }
            }
            if (keep) { len = m & 255; idx = m >> 8; }
        }
    }
    span[B.off + t] = (uint8_t)len;
    tidx[B.off + t] = (uint8_t)idx;
}

// ------------------------------------------------------------------ token walk -> PPM events + side symbols
// (coder loop src/rolzmain/cr-coder.c:197-234, minus the entropy coding itself)
struct RolzCount {
    typedef struct { uint32_t ev, match, esclit; } State;
    const uint8_t* D; const LzBlock* blocks; const uint8_t* tidx;
    uint32_t *cnt_ev, *cnt_match, *cnt_esclit;
    CR_D State begin(uint32_t, uint32_t) const { State s = {0, 0, 0}; return s; }
    CR_D void visit(State& s, uint32_t b, uint32_t t, uint32_t) const {
        const LzBlock B = blocks[b];
        s.ev++;
        if (tidx[B.off + t] != 0xFF) s.match++;
        else if (D[B.off + t] == B.esc) s.esclit++;
    }
    CR_D void end(State& s, uint32_t c, uint32_t) const { cnt_ev[c] = s.ev; cnt_match[c] = s.match; cnt_esclit[c] = s.esclit; }
};

// PPM context in front of position t: the last four bytes of the context stream, which is the concatenation
// of bytes 1.. of every block of the chain (byte 0 goes to the header; cr-coder.c:144,227-229).
CR_HD uint32_t rz_ctx_at(const uint8_t* d, uint32_t t, uint32_t cin) {
    if (t >= 5) return (uint32_t)d[t - 4] << 24 | (uint32_t)d[t - 3] << 16 | (uint32_t)d[t - 2] << 8 | d[t - 1];
    uint32_t c = cin;
    for (uint32_t i = 1; i < t; i++) c = c << 8 | d[i];
    return c;
}

struct RolzEmit {
    typedef struct { uint32_t ev, side, nlen, nidx; } State;
    const uint8_t* D; const LzBlock* blocks; const uint8_t* tidx;
    const uint32_t *scan_ev, *scan_match, *scan_esclit;
    uint32_t* ev_ctx; uint8_t* ev_sym; uint16_t* side_sym;     // side symbol: value | model<<8 (0 = len model, 1 = idx model)
    // the same symbols split per model (k-th use of a model -> symbol, position in the side stream), for k_side_epochs
    uint8_t* len_sym; uint32_t* len_pos; uint8_t* idx_sym; uint32_t* idx_pos;
    CR_D State begin(uint32_t c, uint32_t) const { State s = { scan_ev[c], 2 * scan_match[c] + scan_esclit[c], scan_match[c] + scan_esclit[c], scan_match[c] }; return s; }
    CR_D void visit(State& s, uint32_t b, uint32_t t, uint32_t len) const {
        const LzBlock B = blocks[b];
        const uint8_t* d = D + B.off;
        uint32_t idx = tidx[B.off + t];
        ev_ctx[s.ev] = rz_ctx_at(d, t, B.cin);
        if (idx != 0xFF) {
            ev_sym[s.ev] = B.esc;
            len_sym[s.nlen] = (uint8_t)len; len_pos[s.nlen++] = s.side;
            side_sym[s.side++] = (uint16_t)len;
            idx_sym[s.nidx] = (uint8_t)idx; idx_pos[s.nidx++] = s.side;
            side_sym[s.side++] = (uint16_t)(idx | 0x100);
        } else {
            ev_sym[s.ev] = d[t];
            if (d[t] == B.esc) { len_sym[s.nlen] = 0; len_pos[s.nlen++] = s.side; side_sym[s.side++] = 0; }
        }
        s.ev++;
    }
    CR_D void end(State&, uint32_t, uint32_t) const {}
};
