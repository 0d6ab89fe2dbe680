// cr_prims.cuh -- device-wide building blocks: byte histograms and escape selection (sort / scan live in cr_sort.cuh).
#pragma once
#include "cr_common.cuh"
#include "cr_sort.cuh"

// ------------------------------------------------------------------ segmented byte histogram
// One 256-bin histogram per segment (a raw block or a dictionary-coded block).
// Replaces the counting loops src/cr-diccode.c:161-163 and src/rolzmain/cr-coder.c:171-173.
// Roofline: HBM, 1 byte read per input byte.  Grid = (tiles, segments); each CTA keeps one sub-histogram per
// warp in shared memory (bank-conflict-free for distinct bytes, ATOMS on equal bytes) and flushes 256 REDs.
#define CR_HIST_TILE 65536u

__global__ void k_hist256(const uint8_t* __restrict__ data, const uint64_t* __restrict__ seg_off, const uint32_t* __restrict__ seg_len, uint32_t* __restrict__ hist) {
    const uint32_t s = blockIdx.y;
    const uint32_t len = seg_len[s];
    const uint32_t t0 = blockIdx.x * CR_HIST_TILE;
    if (t0 >= len) return;
    const uint32_t t1 = (len - t0 < CR_HIST_TILE) ? len : t0 + CR_HIST_TILE;
    const uint8_t* d = data + seg_off[s];
#ifdef CRGPU_SIM
    if (threadIdx.x == 0) for (uint32_t i = t0; i < t1; i++) hist[s * 256 + d[i]]++;
#else
    __shared__ uint32_t sh[8][256];
    const uint32_t w = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < 8 * 256; i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();
    // 16-byte vector loads over the aligned middle, byte loads at the ragged ends
    const uint8_t* p = d + t0;
    uint32_t n = t1 - t0;
    uint32_t head = (uint32_t)((16 - ((uintptr_t)p & 15)) & 15);
    if (head > n) head = n;
    for (uint32_t i = threadIdx.x; i < head; i += blockDim.x) atomicAdd(&sh[w][p[i]], 1u);
    const uint4* v = (const uint4*)(p + head);
    uint32_t nv = (n - head) / 16;
    for (uint32_t i = threadIdx.x; i < nv; i += blockDim.x) {
        uint4 q = v[i];
        uint32_t ws[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
        for (int k = 0; k < 4; k++) {
            atomicAdd(&sh[w][ws[k] & 255], 1u); atomicAdd(&sh[w][(ws[k] >> 8) & 255], 1u);
            atomicAdd(&sh[w][(ws[k] >> 16) & 255], 1u); atomicAdd(&sh[w][ws[k] >> 24], 1u);
        }
    }
    for (uint32_t i = head + nv * 16 + threadIdx.x; i < n; i += blockDim.x) atomicAdd(&sh[w][p[i]], 1u);
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < 256; b += blockDim.x) {
        uint32_t c = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) c += sh[k][b];
        if (c) atomicAdd(&hist[s * 256 + b], c);
    }
#endif
}

// Escape selection from the histograms, one thread per segment.
//   esc10: the 10 rarest byte values, ties to the lowest value, picked one at a time (src/cr-diccode.c:164-171)
//   esc1 : the rarest byte value, ties to the lowest (src/rolzmain/cr-coder.c:174-178)
__global__ void k_pick_escapes(const uint32_t* __restrict__ hist, uint32_t nseg, uint8_t* __restrict__ esc10, uint8_t* __restrict__ esc1) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    const uint32_t* h = hist + s * 256;
    uint32_t taken[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < 10; k++) {
        uint32_t best = 0, bc = (taken[0] & 1u) ? 0xFFFFFFFFu : h[0];
        for (uint32_t j = 1; j < 256; j++) {
            uint32_t c = ((taken[j >> 5] >> (j & 31)) & 1u) ? 0xFFFFFFFFu : h[j];
            if (c < bc) { bc = c; best = j; }
        }
        taken[best >> 5] |= 1u << (best & 31);
        if (esc10) esc10[s * 10 + k] = (uint8_t)best;
        if (k == 0 && esc1) esc1[s] = (uint8_t)best;
        if (!esc10) break;
    }
}
