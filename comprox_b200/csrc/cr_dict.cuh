// cr_dict.cuh -- static word dictionary on the GPU: word statistics (dicpick) and substitution (diccode).
//
// dicpick  replaces the tokenizer + hash map of src/cr-dicpick.c:164-216,95-146.  A word start is a local
//          property of two neighbouring bytes, so every byte position is examined in parallel; accepted words
//          are counted in a device-wide open-addressing table keyed by a 64-bit hash.  The host receives only
//          the (first position, count) pairs with count > 5 and performs the tiny ordering step
//          (src/cr-dicpick.c:218-257) there.  Counts are order independent as long as fewer than 325001 distinct
//          words exist; beyond that the reference prunes in arrival order and we refuse loudly.
// diccode  replaces dictionary_encode_imp (src/cr-diccode.c:285-362): the trie walk is evaluated at every word
//          start in parallel, the serial "i = j" skip is resolved with cr_chain.cuh, output offsets come from a
//          scan of the per-position code sizes.
#pragma once
#include "cr_common.cuh"
#include "cr_hostdict.h"

// ------------------------------------------------------------------ dicpick
#define DP_CHUNK      200000u            // fread granularity of the reference tokenizer (cr-dicpick.c:162)
#define DP_SLOTS_LOG2 21
#define DP_SLOTS      (1u << DP_SLOTS_LOG2)
#define DP_MAXWORDS   325001u            // HASHMAP_MAXSIZE (cr-dicpick.c:34)
#define DP_MINLEN     2u
#define DP_MAXLEN     20u

struct DpTable {
    unsigned long long* key;   // [DP_SLOTS] 0 = empty
    uint32_t* count;           // [DP_SLOTS]
    uint32_t* first;           // [DP_SLOTS] smallest position at which the word starts
    uint32_t* stats;           // [0] distinct words, [1] error flags, [2] number of selected entries, [8] set by k_dp_verify: two spellings share a hash
};

CR_HD uint32_t dp_byte(const uint8_t* in, uint64_t chunk0, uint32_t c, uint32_t flen) {
    return c == flen - 1 ? 0u : in[chunk0 + c];                 // the last byte of every chunk reads as 0 (:192)
}
// If a word that the reference would count starts at file position x, returns its length (2..20), else 0.
CR_HD uint32_t dp_word_at(const uint8_t* in, uint64_t n, uint64_t x, unsigned long long* hash_out) {
    const uint64_t chunk0 = x / DP_CHUNK * DP_CHUNK;
    const uint32_t c = (uint32_t)(x - chunk0);
    const uint32_t flen = (uint32_t)(n - chunk0 < DP_CHUNK ? n - chunk0 : DP_CHUNK);
    if (c == 0) return 0;
    uint32_t b = dp_byte(in, chunk0, c, flen);
    if (!cr_is_alpha(b) || cr_is_alpha(in[x - 1])) return 0;
    unsigned long long h = 1469598103934665603ull;
    h = (h ^ (b | 32)) * 1099511628211ull;
    uint32_t y = c + 1;
    while (y < flen) {
        uint32_t v = dp_byte(in, chunk0, y, flen);
        if (!cr_is_lower(v)) break;
        if (y - c >= DP_MAXLEN) return 0;                          // longer than 20: rejected whatever follows
        h = (h ^ v) * 1099511628211ull;
        y++;
    }
    const uint32_t len = y - c;
    if (len < DP_MINLEN || y >= flen) return 0;
    const uint32_t s = dp_byte(in, chunk0, y, flen);
    if (!(s == ' ' || s == ',' || s == '.' || s == ':' || s == ';')) return 0;
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    if (h == 0) h = 1;
    *hash_out = h;
    return len;
}

// positions [x0, x1) of the file; the table persists across launches so a file can be fed in windows
__global__ void k_dp_count(const uint8_t* __restrict__ in, uint64_t n, uint64_t x0, uint64_t x1, DpTable T) {
    uint64_t x = x0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= x1) return;
    unsigned long long h;
    if (!dp_word_at(in, n, x, &h)) return;
    uint32_t slot = (uint32_t)h & (DP_SLOTS - 1);
    for (uint32_t probe = 0; probe < DP_SLOTS; probe++) {
        unsigned long long prev = atomicCAS(&T.key[slot], 0ull, h);
        if (prev == 0ull) {
            if (atomicAdd(&T.stats[0], 1u) + 1 >= DP_MAXWORDS) atomicOr(&T.stats[1], 1u);
            prev = h;
        }
        if (prev == h) { atomicAdd(&T.count[slot], 1u); atomicMin(&T.first[slot], (uint32_t)x); return; }
        slot = (slot + 1) & (DP_SLOTS - 1);
    }
    atomicOr(&T.stats[1], 1u);
}
#ifndef CRGPU_SIM
// The same count with the hot words taken off the device-wide table: a CTA tokenizes a tile of DPS_TILE bytes, 256 positions per round;
// the lanes of a warp that found the same word in a round are found with __match_any_sync and entered once (multiplicity = popcount,
// first position = the lowest lane's); the tile's words are counted in a shared-memory open-addressing table and only the table's
// distinct entries go to the global table -- one CAS / add / min per (tile, distinct word) instead of one per occurrence ("the" alone
// is half a million serialised L2 atomics on a 100 MiB text).  A word that finds no slot within DPS_PROBES probes goes to the global
// table directly.  Counts and first positions are sums and minima: the result does not depend on the order.
// MEASURED (B200, text-100M, ncu): 2.26 ms against 1.4 ms for k_dp_count -- the L2 atomics of the plain kernel are not its limit, and this
// one issues 61 % of its slots on 64 serial rounds per thread with match / shared-memory atomics (1.08 M bank conflicts per 16 MiB).
// It is therefore NOT the default (crgpu_set_option "dp_tiles" = 1 selects it; same dictionary, tests/test_gpu_compress.py).
#define DPS_TILE   16384u
#define DPS_SLOTS  2048u
#define DPS_PROBES 12u
CR_D void dp_global_add(DpTable T, unsigned long long h, uint32_t count, uint32_t first) {
    uint32_t slot = (uint32_t)h & (DP_SLOTS - 1);
    for (uint32_t probe = 0; probe < DP_SLOTS; probe++) {
        unsigned long long prev = atomicCAS(&T.key[slot], 0ull, h);
        if (prev == 0ull) {
            if (atomicAdd(&T.stats[0], 1u) + 1 >= DP_MAXWORDS) atomicOr(&T.stats[1], 1u);
            prev = h;
        }
        if (prev == h) { atomicAdd(&T.count[slot], count); atomicMin(&T.first[slot], first); return; }
        slot = (slot + 1) & (DP_SLOTS - 1);
    }
    atomicOr(&T.stats[1], 1u);
}
__global__ void __launch_bounds__(256) k_dp_count_tiles(const uint8_t* __restrict__ in, uint64_t n, uint64_t x0, uint64_t x1, DpTable T) {
    __shared__ unsigned long long skey[DPS_SLOTS];
    __shared__ uint32_t scount[DPS_SLOTS], sfirst[DPS_SLOTS];
    for (uint32_t i = threadIdx.x; i < DPS_SLOTS; i += 256) { skey[i] = 0ull; scount[i] = 0; sfirst[i] = 0xFFFFFFFFu; }
    __syncthreads();
    const uint64_t tbase = x0 + (uint64_t)blockIdx.x * DPS_TILE;
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t j = 0; j < DPS_TILE / 256; j++) {
        const uint64_t x = tbase + j * 256 + threadIdx.x;
        unsigned long long h = 0ull;
        const bool word = x < x1 && dp_word_at(in, n, x, &h) != 0;
        // lanes of this warp with the same word (lanes without a word carry distinct dummies: their own lane number, never a hash...
        // a real hash below 32 would only merge the group with a lane that is masked out below)
        const unsigned long long tag = word ? h : (unsigned long long)lane;
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, tag) & __ballot_sync(0xFFFFFFFFu, word);
        if (word && lane == (uint32_t)(__ffs(peers) - 1)) {
            const uint32_t mult = __popc(peers);
            uint32_t slot = (uint32_t)(h >> 21) & (DPS_SLOTS - 1), probe = 0;
            for (; probe < DPS_PROBES; probe++) {
                unsigned long long prev = atomicCAS(&skey[slot], 0ull, h);
                if (prev == 0ull || prev == h) { atomicAdd(&scount[slot], mult); atomicMin(&sfirst[slot], (uint32_t)x); break; }
                slot = (slot + 1) & (DPS_SLOTS - 1);
            }
            if (probe == DPS_PROBES) dp_global_add(T, h, mult, (uint32_t)x);
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < DPS_SLOTS; i += 256) if (skey[i] != 0ull) dp_global_add(T, skey[i], scount[i], sfirst[i]);
}
#endif

// every occurrence must spell the same word as the table entry's first occurrence (hash collisions are loud)
__global__ void k_dp_verify(const uint8_t* __restrict__ in, uint64_t n, uint64_t x0, uint64_t x1, DpTable T) {
    uint64_t x = x0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= x1) return;
    unsigned long long h;
    uint32_t len = dp_word_at(in, n, x, &h);
    if (!len) return;
    uint32_t slot = (uint32_t)h & (DP_SLOTS - 1);
    for (uint32_t probe = 0; probe < DP_SLOTS && T.key[slot] != h; probe++) slot = (slot + 1) & (DP_SLOTS - 1);
    const uint8_t* a = in + x; const uint8_t* b = in + T.first[slot];
    bool same = true;
    for (uint32_t i = 0; i < len; i++) same &= (a[i] | 32) == (b[i] | 32);
    same &= !cr_is_lower(b[len]);
    if (!same) atomicOr(&T.stats[8], 1u);
}
// ---- exact handling of vocabulary overflow (cr-dicpick.c:115-144).  When the 325001st distinct word
// arrives the reference drops every word whose count is <= (smallest count) + 5 and carries on, so the result
// depends on WHEN each word first appeared.  That moment is found without a serial pass: list the first position
// of every word that is not yet in the table, take the k-th smallest (k = free places left), count everything up
// to and including that position, prune, continue behind it.  One "epoch" per prune.
// k_dp_first: like k_dp_count but only records first positions; new keys enter with count 0 ("candidates").
__global__ void k_dp_first(const uint8_t* __restrict__ in, uint64_t n, uint64_t x0, uint64_t x1, DpTable T) {
    uint64_t x = x0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= x1) return;
    unsigned long long h;
    if (!dp_word_at(in, n, x, &h)) return;
    uint32_t slot = (uint32_t)h & (DP_SLOTS - 1);
    for (uint32_t probe = 0; probe < DP_SLOTS; probe++) {
        unsigned long long prev = atomicCAS(&T.key[slot], 0ull, h);
        if (prev == 0ull) { atomicAdd(&T.stats[0], 1u); prev = h; }
        if (prev == h) { atomicMin(&T.first[slot], (uint32_t)x); return; }
        slot = (slot + 1) & (DP_SLOTS - 1);
    }
    atomicOr(&T.stats[1], 4u);                                   // table full
}
// first positions (>= x0) of all candidates (count == 0)
__global__ void k_dp_list_new(DpTable T, uint32_t x0, uint32_t* __restrict__ list, uint32_t cap, uint32_t* __restrict__ n_new) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= DP_SLOTS || T.key[s] == 0ull || T.count[s] != 0 || T.first[s] < x0) return;
    uint32_t i = atomicAdd(n_new, 1u);
    if (i < cap) list[i] = T.first[s];
}
__global__ void k_dp_min_count(DpTable T, uint32_t* __restrict__ mn) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= DP_SLOTS || T.key[s] == 0ull || T.count[s] == 0) return;
    atomicMin(mn, T.count[s]);
}
// survivors (count > threshold) move to a fresh table; candidates and pruned words vanish
__global__ void k_dp_rebuild(DpTable A, DpTable B, uint32_t threshold) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= DP_SLOTS || A.key[s] == 0ull || A.count[s] <= threshold) return;
    const unsigned long long h = A.key[s];
    uint32_t slot = (uint32_t)h & (DP_SLOTS - 1);
    for (;;) {
        unsigned long long prev = atomicCAS(&B.key[slot], 0ull, h);
        if (prev == 0ull) { B.count[slot] = A.count[s]; B.first[slot] = A.first[s]; atomicAdd(&B.stats[0], 1u); return; }
        slot = (slot + 1) & (DP_SLOTS - 1);
    }
}

// a selected word leaves the device complete: first position, count, and its letters (first letter lower-cased, at most 20, zero padded)
// as three big-endian 64-bit keys -- what HdWord::set (cr_hostdict.h) used to fetch from the host copy of the input, one cache miss per word
struct DpEntry { uint32_t first, count; unsigned long long k[3]; uint32_t len, pad; };
__global__ void k_dp_collect(DpTable T, const uint8_t* __restrict__ in, uint64_t n, DpEntry* __restrict__ out, uint32_t cap) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= DP_SLOTS || T.key[s] == 0ull || T.count[s] <= 5) return;      // WORD_MIN_FREQ (:221)
    uint32_t i = atomicAdd(&T.stats[2], 1u);
    if (i >= cap) return;
    const uint64_t first = T.first[s];
    uint8_t w[24];
    for (int q = 0; q < 24; q++) w[q] = 0;
    w[0] = (uint8_t)(in[first] | 32);
    uint32_t len = 1;
    for (uint64_t x = first + 1; x < n && in[x] >= 'a' && in[x] <= 'z' && len < 20; x++) w[len++] = in[x];      // copyword, cr-dicpick.c:88-94
    DpEntry e; e.first = (uint32_t)first; e.count = T.count[s]; e.len = len; e.pad = 0;
    for (int q = 0; q < 3; q++) { unsigned long long v = 0; for (int b = 0; b < 8; b++) v = v << 8 | w[q * 8 + b]; e.k[q] = v; }
    out[i] = e;
}

// ------------------------------------------------------------------ diccode
#define DC_SUB 1000000u                  // sub-chunk size (cr-diccode.c:177-180)

struct DcTrie { const HdEdge* edge; const int32_t* id; uint32_t mask; int32_t nentries; int32_t level1; };
// child of `node` along byte `ch` (0 = none); same table as HdTrie (cr_hostdict.h)
CR_D uint32_t dc_child(const DcTrie& T, uint32_t node, uint32_t ch) {
    const uint32_t key = ((node << 7) | ch) + 1;
    for (uint32_t h = (key * 2654435761u) >> 8;; h++) { const HdEdge e = T.edge[h & T.mask]; if (e.key == key) return e.val; if (e.key == 0) return 0; }
}
struct DcSub {                           // one sub-chunk = one chain segment
    uint64_t off;                        // offset of the sub-chunk in the raw window
    uint32_t size;
    uint32_t block;                      // owning block (for the escape set)
    uint64_t out;                        // where its codes start in the dictionary-coded window (set after sizing)
};
#define DC_HIT 0x80000000u

CR_HD bool dc_sentence_start(const uint8_t* s, uint32_t i) {     // M_check_reverse_case (cr-diccode.c:313)
    return i >= 3 && s[i - 1] == ' ' && (s[i - 2] == '.' || (s[i - 2] == ' ' && s[i - 3] == '.'));
}

// span[g] and hit[g] for every raw position of the window (grid.y = sub-chunk)
__global__ void k_dc_spans(const uint8_t* __restrict__ raw, const DcSub* __restrict__ subs, DcTrie T, uint8_t* __restrict__ span, uint32_t* __restrict__ hit) {
    const DcSub S = subs[blockIdx.y];
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.size) return;
    const uint8_t* d = raw + S.off;
    uint32_t sp = 1;
    if (i > 0 && i + 40 < S.size && cr_is_alpha(d[i]) && !cr_is_alpha(d[i - 1])) {
        uint32_t j = i, node = 0;
        while (d[j] < 128 && (node = dc_child(T, node, d[j])) != 0 && T.id[node] == -1) j++;
        if (d[j] < 128 && node != 0) {
            uint32_t rev = (uint32_t)cr_is_upper(d[i]) ^ (uint32_t)dc_sentence_start(d, i);
            uint32_t tail = d[j] == ':' ? 4 : d[j] == ';' ? 3 : d[j] == ',' ? 2 : d[j] == '.' ? 1 : 0;
            sp = j - i + 1;
            hit[S.off + i] = DC_HIT | (uint32_t)T.id[node] | (rev * 5 + tail) << 24;
        }
    }
    span[S.off + i] = (uint8_t)sp;
}

#ifndef CRGPU_SIM
// The same in three steps (default on the GPU).  k_dc_spans gives every byte a thread, and a warp then walks the trie for as long as its
// longest word while five of its 32 lanes have a word at all (ncu, round 1: 68-78 % of the issue slots, 3 % of the HBM peak).  Here the
// word starts are listed first -- counted per CTA (__syncthreads_count), offsets from the device scan, positions written in order with
// a ballot / popcount prefix inside the CTA -- and the trie is walked with one LISTED word per lane.
#define DCS_TH 256
CR_D bool dc_is_start(const uint8_t* d, uint32_t i, uint32_t size) { return i > 0 && i + 40 < size && cr_is_alpha(d[i]) && !cr_is_alpha(d[i - 1]); }
__global__ void __launch_bounds__(DCS_TH) k_dc_count_starts(const uint8_t* __restrict__ raw, const DcSub* __restrict__ subs, uint32_t* __restrict__ cta_count) {
    const DcSub S = subs[blockIdx.y];
    const uint32_t i = blockIdx.x * DCS_TH + threadIdx.x;
    const int c = __syncthreads_count(i < S.size && dc_is_start(raw + S.off, i, S.size));
    if (threadIdx.x == 0) cta_count[blockIdx.y * gridDim.x + blockIdx.x] = (uint32_t)c;
}
__global__ void __launch_bounds__(DCS_TH) k_dc_list_starts(const uint8_t* __restrict__ raw, const DcSub* __restrict__ subs, const uint32_t* __restrict__ cta_off,
                                                           uint8_t* __restrict__ span, uint2* __restrict__ list) {
    __shared__ uint32_t wsum[DCS_TH / 32];
    const DcSub S = subs[blockIdx.y];
    const uint32_t i = blockIdx.x * DCS_TH + threadIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const bool in = i < S.size;
    const bool ws = in && dc_is_start(raw + S.off, i, S.size);
    if (in) span[S.off + i] = 1;                                       // k_dc_walk raises the spans of the words it finds
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, ws);
    if (lane == 0) wsum[w] = __popc(m);
    __syncthreads();
    uint32_t off = cta_off[blockIdx.y * gridDim.x + blockIdx.x];
    for (uint32_t q = 0; q < w; q++) off += wsum[q];
    if (ws) list[off + __popc(m & ((1u << lane) - 1u))] = make_uint2((uint32_t)(S.off + i), blockIdx.y);
}
__global__ void __launch_bounds__(DCS_TH) k_dc_walk(const uint8_t* __restrict__ raw, const DcSub* __restrict__ subs, DcTrie T, const uint2* __restrict__ list,
                                                    const uint32_t* __restrict__ total, uint8_t* __restrict__ span, uint32_t* __restrict__ hit) {
    const uint32_t t = blockIdx.x * DCS_TH + threadIdx.x;
    if (t >= *total) return;
    const uint2 e = list[t];
    const DcSub S = subs[e.y];
    const uint8_t* d = raw + S.off;
    const uint32_t i = e.x - (uint32_t)S.off;
    uint32_t j = i, node = 0;
    while (d[j] < 128 && (node = dc_child(T, node, d[j])) != 0 && T.id[node] == -1) j++;
    if (d[j] < 128 && node != 0) {
        const uint32_t rev = (uint32_t)cr_is_upper(d[i]) ^ (uint32_t)dc_sentence_start(d, i);
        const uint32_t tail = d[j] == ':' ? 4 : d[j] == ';' ? 3 : d[j] == ',' ? 2 : d[j] == '.' ? 1 : 0;
        span[e.x] = (uint8_t)(j - i + 1);
        hit[e.x] = DC_HIT | (uint32_t)T.id[node] | (rev * 5 + tail) << 24;
    }
}
#endif

struct DcCount {
    typedef uint32_t State;
    const uint8_t* raw; const DcSub* subs; const uint8_t* span; const uint32_t* hit; const uint32_t* escmask;   // escmask[block*8 + w]
    int32_t level1; uint32_t* cnt;
    CR_D State begin(uint32_t, uint32_t) const { return 0; }
    CR_D void visit(State& n, uint32_t s, uint32_t i, uint32_t sp) const {
        const DcSub S = subs[s];
        if (sp > 1) { uint32_t id = hit[S.off + i] & 0xFFFFFF; n += (int32_t)id < level1 ? 2 : 3; }
        else { uint32_t b = raw[S.off + i]; n += (escmask[S.block * 8 + (b >> 5)] >> (b & 31) & 1) ? 3 : 1; }
    }
    CR_D void end(State& n, uint32_t c, uint32_t) const { cnt[c] = n; }
};
struct DcEmit {
    typedef uint64_t State;
    const uint8_t* raw; const DcSub* subs; const uint8_t* span; const uint32_t* hit; const uint32_t* escmask; const uint8_t* esc10;
    int32_t level1, nentries; const uint32_t* scan; const uint32_t* sub_chunk0; uint8_t* out;
    CR_D State begin(uint32_t c, uint32_t s) const { return subs[s].out == ~0ull ? ~0ull : subs[s].out + (scan[c] - scan[sub_chunk0[s]]); }
    CR_D void visit(State& o, uint32_t s, uint32_t i, uint32_t sp) const {
        if (o == ~0ull) return;                                   // block is stored raw: nothing to emit
        const DcSub S = subs[s];
        const uint32_t wide = 256 - level1;
        if (sp > 1) {                                             // cr-diccode.c:327-336
            uint32_t h = hit[S.off + i], id = h & 0xFFFFFF;
            if ((int32_t)id < level1) out[o++] = (uint8_t)id;
            else { out[o++] = (uint8_t)(id / wide); out[o++] = (uint8_t)(id % wide + level1); }
            out[o++] = esc10[S.block * 10 + ((h >> 24) & 0x7F)];
        } else {                                                  // cr-diccode.c:338-346
            uint32_t b = raw[S.off + i];
            if (escmask[S.block * 8 + (b >> 5)] >> (b & 31) & 1) { out[o++] = (uint8_t)(nentries / wide); out[o++] = (uint8_t)(nentries % wide + level1); }
            out[o++] = (uint8_t)b;
        }
    }
    CR_D void end(State&, uint32_t, uint32_t) const {}
};
__global__ void k_dc_escmask(const uint8_t* __restrict__ esc10, uint32_t nblocks, uint32_t* __restrict__ mask) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    uint32_t m[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < 10; k++) { uint32_t e = esc10[b * 10 + k]; m[e >> 5] |= 1u << (e & 31); }
    for (int w = 0; w < 8; w++) mask[b * 8 + w] = m[w];
}
