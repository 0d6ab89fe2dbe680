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
#include <vector>
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

// B: entry offset of every chunk.  Exit tables compose (entering chunk c at o leaves chunk c+1 at xt[c+1][xt[c][o]]),
// so the serial walk over n/1024 tables is cut into groups of CR_GROUP chunks:
//   B1  every group composes its tables into one 256-entry map         (parallel over groups x entries)
//   B2  one thread per segment walks the group maps                      (n / (1024 * CR_GROUP) dependent loads)
//   B3  every group re-walks its own tables from its true entry          (parallel over groups)
#define CR_GROUP 64u
__global__ void k_chain_group_maps(const ChainSeg* __restrict__ segs, uint32_t nseg, const uint8_t* __restrict__ xt, const uint32_t* __restrict__ group0,
                                   uint32_t ngroup, uint8_t* __restrict__ gmap) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;          // (group, entry)
    const uint32_t g = t >> 8, o = t & 255;
    if (g >= ngroup) return;
    // segment of this group: last segment whose first group is <= g
    uint32_t lo = 0, hi = nseg;
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (group0[mid] <= g) lo = mid; else hi = mid; }
    const ChainSeg sg = segs[lo];
    const uint32_t c0 = sg.chunk0 + (g - group0[lo]) * CR_GROUP;
    const uint32_t c1 = c0 + CR_GROUP < sg.chunk0 + sg.nchunk ? c0 + CR_GROUP : sg.chunk0 + sg.nchunk;
    uint32_t e = o;
    for (uint32_t c = c0; c < c1; c++) e = xt[(size_t)c * 256 + e];
    gmap[(size_t)g * 256 + o] = (uint8_t)e;
}
__global__ void k_chain_group_entries(const ChainSeg* __restrict__ segs, uint32_t nseg, const uint32_t* __restrict__ group0, const uint8_t* __restrict__ gmap,
                                      uint8_t* __restrict__ gentry) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    const ChainSeg sg = segs[s];
    const uint32_t ng = (sg.nchunk + CR_GROUP - 1) / CR_GROUP;
    uint32_t e = sg.start;
    for (uint32_t k = 0; k < ng; k++) { gentry[group0[s] + k] = (uint8_t)e; e = gmap[(size_t)(group0[s] + k) * 256 + e]; }
}
__global__ void k_chain_entries_grouped(const ChainSeg* __restrict__ segs, uint32_t nseg, const uint8_t* __restrict__ xt, const uint32_t* __restrict__ group0,
                                        uint32_t ngroup, const uint8_t* __restrict__ gentry, uint8_t* __restrict__ entry) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngroup) return;
    uint32_t lo = 0, hi = nseg;
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (group0[mid] <= g) lo = mid; else hi = mid; }
    const ChainSeg sg = segs[lo];
    const uint32_t c0 = sg.chunk0 + (g - group0[lo]) * CR_GROUP;
    const uint32_t c1 = c0 + CR_GROUP < sg.chunk0 + sg.nchunk ? c0 + CR_GROUP : sg.chunk0 + sg.nchunk;
    uint32_t e = gentry[g];
    for (uint32_t c = c0; c < c1; c++) { entry[c] = (uint8_t)e; e = xt[(size_t)c * 256 + e]; }
}
// plain serial form (kept for tiny inputs and as the definition)
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

// Host-side driver of step B.  `work` must hold (ngroup * 257 + nseg * 4 + 64) bytes.
struct ChainEntryPlan { std::vector<uint32_t> group0; uint32_t ngroup; };
static inline ChainEntryPlan cr_chain_groups(const ChainSeg* segs, uint32_t nseg) {
    ChainEntryPlan p; p.group0.resize(nseg + 1); uint32_t g = 0;
    for (uint32_t s = 0; s < nseg; s++) { p.group0[s] = g; g += (segs[s].nchunk + CR_GROUP - 1) / CR_GROUP; }
    p.group0[nseg] = g; p.ngroup = g;
    return p;
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

// Runs step B (entry offsets of all chunks) on `stream`.  d_work: device scratch of at least cr_chain_entries_scratch() bytes.
static inline size_t cr_chain_entries_scratch(uint32_t nseg, uint32_t nchunk) { return (size_t)(nchunk / CR_GROUP + nseg + 2) * 257 + (size_t)(nseg + 1) * 4 + 512; }
template <class Chain>
static int cr_chain_run_entries(Chain& C, DevBuf& d_work, const std::vector<ChainSeg>& segs, const ChainSeg* d_segs, uint32_t nchunk, const uint8_t* d_xt, uint8_t* d_entry) {
    const uint32_t nseg = (uint32_t)segs.size();
    cudaStream_t stream = C.stream;
    if (nchunk < 4 * CR_GROUP) {        // tiny: the serial form is one launch
        CR_LAUNCH(k_chain_entries, dim3(cr_div_up(nseg, 32)), dim3(32), stream, d_segs, nseg, d_xt, d_entry);
        return CRGPU_OK;
    }
    ChainEntryPlan plan = cr_chain_groups(segs.data(), nseg);
    CR_TRY(d_work.reserve(cr_chain_entries_scratch(nseg, nchunk)));
    uint32_t* d_group0 = d_work.as<uint32_t>();
    uint8_t* d_gmap = (uint8_t*)d_work.p + (((size_t)(nseg + 1) * 4 + 255) & ~(size_t)255);
    uint8_t* d_gentry = d_gmap + (size_t)plan.ngroup * 256;
    CR_CUDA(cudaMemcpyAsync(d_group0, plan.group0.data(), (size_t)(nseg + 1) * 4, cudaMemcpyHostToDevice, stream));
    CR_LAUNCH(k_chain_group_maps, dim3(cr_div_up((size_t)plan.ngroup * 256, 256)), dim3(256), stream, d_segs, nseg, d_xt, d_group0, plan.ngroup, d_gmap);
    CR_LAUNCH(k_chain_group_entries, dim3(cr_div_up(nseg, 32)), dim3(32), stream, d_segs, nseg, d_group0, d_gmap, d_gentry);
    CR_LAUNCH(k_chain_entries_grouped, dim3(cr_div_up(plan.ngroup, 64)), dim3(64), stream, d_segs, nseg, d_xt, d_group0, plan.ngroup, d_gentry, d_entry);
    return CRGPU_OK;
}
