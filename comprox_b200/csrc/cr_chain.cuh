// cr_chain.cuh -- exact parallel resolution of a serial "skip" parse.
//
// Three loops of the reference are serial only because each step decides where the next one starts:
//   * the ROLZ / LZP parse           pos += match_len          (src/rolzmain/cr-coder.c:115-129, src/ropmain/cr-coder.c:101-111)
//   * dictionary substitution        i = j on a word hit       (src/cr-diccode.c:300-347)
// In all of them the step taken AT a position is a pure function of the data, so we first compute
// span[p] for every position in parallel and then find the orbit of the start position:
//   A  per 1024-position chunk, a backward DP gives, for each of the <=255 possible entry offsets, the offset at
//      which the parse enters the next chunk                           (parallel over chunks)
//   B  one thread per segment composes those tables chunk by chunk      (serial, n/1024 dependent loads)
//   C  every chunk re-walks from its true entry and visits exactly the positions the serial loop visits.
// Spans are <= 255 everywhere in comprox, which bounds the entry offsets.
#pragma once
#include "cr_common.cuh"

#define CR_CHUNK 1024u

struct ChainSeg {
    uint64_t off;      // offset of the segment's position 0 in the span array
    uint32_t len;      // positions in the segment
    uint32_t start;    // first position the serial loop visits (< 256)
    uint32_t chunk0;   // index of the segment's first chunk in the flattened chunk list
    uint32_t nchunk;
};

CR_D uint32_t cr_find_seg(const ChainSeg* __restrict__ segs, uint32_t nseg, uint32_t chunk) {
    uint32_t lo = 0, hi = nseg;
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (segs[mid].chunk0 <= chunk) lo = mid; else hi = mid; }
    return lo;
}

// A: exit table per chunk.  xt[chunk*256 + o] = offset (relative to the next chunk) reached when entering at o.
__global__ void k_chain_exits(const uint8_t* __restrict__ span, const ChainSeg* __restrict__ segs, uint32_t nseg, uint32_t nchunk, uint8_t* __restrict__ xt) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunk) return;
    uint32_t s = cr_find_seg(segs, nseg, c);
    ChainSeg sg = segs[s];
    uint32_t base = (c - sg.chunk0) * CR_CHUNK;
    uint32_t clen = sg.len - base < CR_CHUNK ? sg.len - base : CR_CHUNK;
    const uint8_t* sp = span + sg.off + base;
    uint8_t x[CR_CHUNK];
    for (int p = (int)clen - 1; p >= 0; p--) {
        uint32_t l = sp[p]; if (l == 0) l = 1;
        uint32_t np = (uint32_t)p + l;
        x[p] = (uint8_t)(np >= CR_CHUNK ? np - CR_CHUNK : (np >= clen ? 0 : x[np]));
    }
    uint32_t m = clen < 256 ? clen : 256;
    for (uint32_t o = 0; o < m; o++) xt[(size_t)c * 256 + o] = x[o];
}

// B: entry offset of every chunk.
__global__ void k_chain_entries(const ChainSeg* __restrict__ segs, uint32_t nseg, const uint8_t* __restrict__ xt, uint8_t* __restrict__ entry) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    ChainSeg sg = segs[s];
    uint32_t e = sg.start;
    for (uint32_t k = 0; k < sg.nchunk; k++) {
        entry[sg.chunk0 + k] = (uint8_t)e;
        e = xt[(size_t)(sg.chunk0 + k) * 256 + e];
    }
}

// C: walk every chunk from its entry.  F must provide:
//   typedef State;  State begin(uint32_t chunk, uint32_t seg) const;
//   void visit(State&, uint32_t seg, uint32_t pos /*within segment*/, uint32_t span) const;
//   void end(State&, uint32_t chunk, uint32_t seg) const;
template <class F>
__global__ void k_chain_walk(const uint8_t* __restrict__ span, const ChainSeg* __restrict__ segs, uint32_t nseg, uint32_t nchunk, const uint8_t* __restrict__ entry, F f) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunk) return;
    uint32_t s = cr_find_seg(segs, nseg, c);
    ChainSeg sg = segs[s];
    uint32_t base = (c - sg.chunk0) * CR_CHUNK;
    uint32_t clen = sg.len - base < CR_CHUNK ? sg.len - base : CR_CHUNK;
    const uint8_t* sp = span + sg.off + base;
    typename F::State st = f.begin(c, s);
    for (uint32_t p = entry[c]; p < clen;) {
        uint32_t l = sp[p]; if (l == 0) l = 1;
        f.visit(st, s, base + p, l);
        p += l;
    }
    f.end(st, c, s);
}

// Host helper: lay out the flattened chunk list.  Returns total chunks.
static inline uint32_t cr_chain_layout(ChainSeg* segs, uint32_t nseg) {
    uint32_t c = 0;
    for (uint32_t s = 0; s < nseg; s++) {
        segs[s].chunk0 = c;
        segs[s].nchunk = (segs[s].len + CR_CHUNK - 1) / CR_CHUNK;
        c += segs[s].nchunk;
    }
    return c;
}
