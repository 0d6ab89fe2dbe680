// cr_sort.cuh -- hand-written device-wide primitives: exclusive scan and stable LSD radix sort of (key, value) pairs.
//
// Sorting positions by context bucket IS the construction of the ROLZ / LZP "per-context offset tables", and sorting
// events by model context is what turns the serial coder into per-context replays, so these sit on the hot path.
//   scan : three-level (1024-element blocks -> block sums -> recurse), in place capable (out may alias in).
//   sort : 8 bits per pass.  Per pass: k_rs_hist (per-tile digit histograms, digit-major), scan, k_rs_scatter.
//          Stability inside a tile: a tile is 8 warps x 512 consecutive elements; per-warp digit counts are prefixed over
//          warps, and inside a warp elements are ranked 32 at a time with __match_any_sync + per-digit running counters.
// The CPU kernel-logic simulation (CRGPU_SIM) replaces both by std:: algorithms (they are cooperative kernels).
#pragma once
#include "cr_common.cuh"
#ifdef CRGPU_SIM
#include <algorithm>
#include <vector>
#endif

struct Prims {
    cudaStream_t stream = 0;
    DevBuf temp, temp2, pk, pv;      // digit counts, scan scratch, ping-pong keys / values
    void release() { temp.release(); temp2.release(); pk.release(); pv.release(); }
};

#ifndef CRGPU_SIM
// ------------------------------------------------------------------ scan
#define SC_BLOCK 1024u
// each CTA scans SC_BLOCK elements (one per thread) and writes its total
__global__ void __launch_bounds__(SC_BLOCK) k_scan_block(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t n, uint32_t* __restrict__ sums) {
    __shared__ uint32_t wsum[32];
    const size_t i = (size_t)blockIdx.x * SC_BLOCK + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t v = i < n ? in[i] : 0u;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= (uint32_t)d) inc += t; }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t s = wsum[lane], si = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, si, d); if (lane >= (uint32_t)d) si += t; }
        wsum[lane] = si - s;
        if (lane == 31 && sums) sums[blockIdx.x] = si;
    }
    __syncthreads();
    if (i < n) out[i] = inc - v + wsum[w];
}
__global__ void k_scan_add(uint32_t* __restrict__ out, size_t n, const uint32_t* __restrict__ offs) {
    const size_t i = (size_t)blockIdx.x * SC_BLOCK + threadIdx.x;
    if (i < n) out[i] += offs[blockIdx.x];
}
// exclusive prefix sum of n uint32 (mod 2^32); scratch must hold 2 * (n / SC_BLOCK + 2) uint32 (+ the same recursively / 1024)
static int cr_scan_rec(cudaStream_t stream, const uint32_t* in, uint32_t* out, size_t n, uint32_t* scratch) {
    const unsigned nblk = cr_div_up(n, SC_BLOCK);
    if (nblk <= 1) { CR_LAUNCH(k_scan_block, dim3(1), dim3(SC_BLOCK), stream, in, out, n, (uint32_t*)nullptr); return CRGPU_OK; }
    uint32_t* sums = scratch;
    CR_LAUNCH(k_scan_block, dim3(nblk), dim3(SC_BLOCK), stream, in, out, n, sums);
    CR_TRY(cr_scan_rec(stream, sums, sums, nblk, scratch + ((nblk + 31) & ~31u)));
    CR_LAUNCH(k_scan_add, dim3(nblk), dim3(SC_BLOCK), stream, out, n, sums);
    return CRGPU_OK;
}
#endif

#ifndef CRGPU_SIM
// inclusive prefix MAXIMUM, same three-level shape
__global__ void __launch_bounds__(SC_BLOCK) k_maxscan_block(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t n, uint32_t* __restrict__ tops) {
    __shared__ uint32_t wmax[32];
    const size_t i = (size_t)blockIdx.x * SC_BLOCK + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t v = i < n ? in[i] : 0u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, d); if (lane >= (uint32_t)d && t > v) v = t; }
    if (lane == 31) wmax[w] = v;
    __syncthreads();
    if (w == 0) {
        uint32_t s = wmax[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, s, d); if (lane >= (uint32_t)d && t > s) s = t; }
        wmax[lane] = s;
        if (lane == 31 && tops) tops[blockIdx.x] = s;
    }
    __syncthreads();
    if (w > 0 && wmax[w - 1] > v) v = wmax[w - 1];
    if (i < n) out[i] = v;
}
__global__ void k_maxscan_add(uint32_t* __restrict__ out, size_t n, const uint32_t* __restrict__ tops) {
    const size_t i = (size_t)blockIdx.x * SC_BLOCK + threadIdx.x;
    if (i < n && blockIdx.x > 0) { const uint32_t t = tops[blockIdx.x - 1]; if (t > out[i]) out[i] = t; }
}
static int cr_maxscan_rec(cudaStream_t stream, const uint32_t* in, uint32_t* out, size_t n, uint32_t* scratch) {
    const unsigned nblk = cr_div_up(n, SC_BLOCK);
    if (nblk <= 1) { CR_LAUNCH(k_maxscan_block, dim3(1), dim3(SC_BLOCK), stream, in, out, n, (uint32_t*)nullptr); return CRGPU_OK; }
    uint32_t* tops = scratch;
    CR_LAUNCH(k_maxscan_block, dim3(nblk), dim3(SC_BLOCK), stream, in, out, n, tops);
    CR_TRY(cr_maxscan_rec(stream, tops, tops, nblk, scratch + ((nblk + 31) & ~31u)));
    CR_LAUNCH(k_maxscan_add, dim3(nblk), dim3(SC_BLOCK), stream, out, n, tops);
    return CRGPU_OK;
}
#endif
static int cr_inclusive_max(Prims& P, const uint32_t* in, uint32_t* out, size_t n) {
    if (n == 0) return CRGPU_OK;
#ifdef CRGPU_SIM
    uint32_t acc = 0;
    for (size_t i = 0; i < n; i++) { if (in[i] > acc) acc = in[i]; out[i] = acc; }
    (void)P;
    return CRGPU_OK;
#else
    CR_TRY(P.temp2.reserve((n / SC_BLOCK + 64) * 8 + 4096));
    return cr_maxscan_rec(P.stream, in, out, n, P.temp2.as<uint32_t>());
#endif
}

static int cr_exclusive_sum(Prims& P, const uint32_t* in, uint32_t* out, size_t n) {
    if (n == 0) return CRGPU_OK;
#ifdef CRGPU_SIM
    uint32_t acc = 0;
    for (size_t i = 0; i < n; i++) { uint32_t v = in[i]; out[i] = acc; acc += v; }
    (void)P;
    return CRGPU_OK;
#else
    CR_TRY(P.temp2.reserve((n / SC_BLOCK + 64) * 8 + 4096));
    return cr_scan_rec(P.stream, in, out, n, P.temp2.as<uint32_t>());
#endif
}

#ifndef CRGPU_SIM
// ------------------------------------------------------------------ radix sort
#define RS_THREADS 256u
#define RS_ITEMS   16u
#define RS_TILE    (RS_THREADS * RS_ITEMS)      // 4096 elements: 8 warps x 512 consecutive elements
#define RS_WARPS   (RS_THREADS / 32)

template <class K>
__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const K* __restrict__ keys, size_t n, int shift, uint32_t mask, uint32_t ntiles, uint32_t* __restrict__ counts) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (uint32_t j = 0; j < RS_ITEMS; j++) {
        const size_t i = base + j * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & mask], 1u);
    }
    __syncthreads();
    counts[(size_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];            // digit-major: a scan gives global bases
}

template <class K>
__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const K* __restrict__ kin, const uint32_t* __restrict__ vin, K* __restrict__ kout, uint32_t* __restrict__ vout,
                                                           size_t n, int shift, uint32_t mask, uint32_t ntiles, const uint32_t* __restrict__ bases) {
    __shared__ uint32_t wcnt[RS_WARPS][256];      // per-warp digit counts -> exclusive prefix over warps + global base
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wcnt[0][0])[i] = 0;
    __syncthreads();
    // warp w owns elements [w*512, (w+1)*512) of the tile, round r = 32 consecutive elements
    const size_t wbase = (size_t)blockIdx.x * RS_TILE + (size_t)w * (RS_ITEMS * 32);
    K key[RS_ITEMS]; uint32_t val[RS_ITEMS], dig[RS_ITEMS];
#pragma unroll
    for (uint32_t r = 0; r < RS_ITEMS; r++) {
        const size_t i = wbase + r * 32 + lane;
        const bool ok = i < n;
        key[r] = ok ? kin[i] : (K)0; val[r] = ok ? vin[i] : 0u;
        dig[r] = ok ? ((uint32_t)(key[r] >> shift) & mask) : 0xFFFFFFFFu;
        // count: one lane per distinct digit of the round adds the multiplicity
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, dig[r]);
        if (ok && lane == (uint32_t)(__ffs(peers) - 1)) wcnt[w][dig[r]] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // exclusive prefix over warps per digit, plus the global base of (digit, tile)
    {
        const uint32_t d = threadIdx.x;
        uint32_t run = bases[(size_t)d * ntiles + blockIdx.x];
#pragma unroll
        for (uint32_t q = 0; q < RS_WARPS; q++) { const uint32_t c = wcnt[q][d]; wcnt[q][d] = run; run += c; }
    }
    __syncthreads();
    // rank and scatter, in element order
#pragma unroll
    for (uint32_t r = 0; r < RS_ITEMS; r++) {
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, dig[r]);
        const bool ok = dig[r] != 0xFFFFFFFFu;
        uint32_t basepos = 0;
        const uint32_t leader = __ffs(peers) - 1;
        if (ok && lane == leader) { basepos = wcnt[w][dig[r]]; wcnt[w][dig[r]] = basepos + __popc(peers); }
        basepos = __shfl_sync(0xFFFFFFFFu, basepos, leader);
        if (ok) {
            const uint32_t dst = basepos + __popc(peers & ((1u << lane) - 1u));
            kout[dst] = key[r]; vout[dst] = val[r];
        }
        __syncwarp();
    }
}

// 32-bit keys: the same ranking, but the tile is first put in digit order in shared memory and leaves in runs -- consecutive threads
// write consecutive addresses of one digit's range (a run is ~16 elements with 256 digits and 4096 elements per tile), where
// k_rs_scatter writes 32 scattered 4-byte words per warp and round (one 32-byte sector each).
__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter_staged(const uint32_t* __restrict__ kin, const uint32_t* __restrict__ vin, uint32_t* __restrict__ kout, uint32_t* __restrict__ vout,
                                                                  size_t n, int shift, uint32_t mask, uint32_t ntiles, const uint32_t* __restrict__ bases) {
    __shared__ uint32_t wcnt[RS_WARPS][256];
    __shared__ uint32_t sk[RS_TILE], sv[RS_TILE];
    __shared__ uint32_t goff[256], wsum[RS_WARPS];
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wcnt[0][0])[i] = 0;
    __syncthreads();
    const size_t tbase = (size_t)blockIdx.x * RS_TILE;
    const size_t wbase = tbase + (size_t)w * (RS_ITEMS * 32);
    uint32_t key[RS_ITEMS], val[RS_ITEMS], dig[RS_ITEMS];
#pragma unroll
    for (uint32_t r = 0; r < RS_ITEMS; r++) {
        const size_t i = wbase + r * 32 + lane;
        const bool ok = i < n;
        key[r] = ok ? kin[i] : 0u; val[r] = ok ? vin[i] : 0u;
        dig[r] = ok ? ((key[r] >> shift) & mask) : 0xFFFFFFFFu;
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, dig[r]);
        if (ok && lane == (uint32_t)(__ffs(peers) - 1)) wcnt[w][dig[r]] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    {   // per digit: prefix over warps; over digits: where the digit starts inside the tile
        const uint32_t d = threadIdx.x;
        uint32_t cq[RS_WARPS], tot = 0;
#pragma unroll
        for (uint32_t q = 0; q < RS_WARPS; q++) { cq[q] = wcnt[q][d]; tot += cq[q]; }
        uint32_t inc = tot;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, s); if (lane >= (uint32_t)s) inc += t; }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        uint32_t off = 0;
#pragma unroll
        for (uint32_t q = 0; q < RS_WARPS; q++) if (q < w) off += wsum[q];
        uint32_t run = off + inc - tot;                              // first local position of digit d
        goff[d] = bases[(size_t)d * ntiles + blockIdx.x] - run;
#pragma unroll
        for (uint32_t q = 0; q < RS_WARPS; q++) { wcnt[q][d] = run; run += cq[q]; }
    }
    __syncthreads();
#pragma unroll
    for (uint32_t r = 0; r < RS_ITEMS; r++) {
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, dig[r]);
        const bool ok = dig[r] != 0xFFFFFFFFu;
        uint32_t basepos = 0;
        const uint32_t leader = __ffs(peers) - 1;
        if (ok && lane == leader) { basepos = wcnt[w][dig[r]]; wcnt[w][dig[r]] = basepos + __popc(peers); }
        basepos = __shfl_sync(0xFFFFFFFFu, basepos, leader);
        if (ok) { const uint32_t loc = basepos + __popc(peers & ((1u << lane) - 1u)); sk[loc] = key[r]; sv[loc] = val[r]; }
        __syncwarp();
    }
    __syncthreads();
    const uint32_t tn = n - tbase < RS_TILE ? (uint32_t)(n - tbase) : RS_TILE;
#pragma unroll
    for (uint32_t j = 0; j < RS_ITEMS; j++) {
        const uint32_t i = j * RS_THREADS + threadIdx.x;
        if (i < tn) { const uint32_t k = sk[i]; const uint32_t dst = goff[(k >> shift) & mask] + i; kout[dst] = k; vout[dst] = sv[i]; }
    }
}
template <class K> struct RsScatter {
    static void launch(const K* ks, const uint32_t* vs, K* kd, uint32_t* vd, size_t n, int shift, uint32_t mask, uint32_t ntiles, const uint32_t* counts, cudaStream_t st) {
        k_rs_scatter<K><<<dim3(ntiles), dim3(RS_THREADS), 0, st>>>(ks, vs, kd, vd, n, shift, mask, ntiles, counts);
    }
};
template <> struct RsScatter<uint32_t> {
    static void launch(const uint32_t* ks, const uint32_t* vs, uint32_t* kd, uint32_t* vd, size_t n, int shift, uint32_t mask, uint32_t ntiles, const uint32_t* counts, cudaStream_t st) {
        k_rs_scatter_staged<<<dim3(ntiles), dim3(RS_THREADS), 0, st>>>(ks, vs, kd, vd, n, shift, mask, ntiles, counts);
    }
};
#endif

// Stable sort of (key, value) pairs on key bits [begin_bit, end_bit).  Full keys travel with the pairs; the inputs are
// preserved (passes ping-pong between the output and a private buffer, arranged so that the last pass lands in the output).
template <class K>
static int cr_sort_pairs(Prims& P, const K* kin, K* kout, const uint32_t* vin, uint32_t* vout, size_t n, int begin_bit, int end_bit) {
    if (n == 0) return CRGPU_OK;
#ifdef CRGPU_SIM
    std::vector<uint32_t> order(n);
    for (size_t i = 0; i < n; i++) order[i] = (uint32_t)i;
    const K mask = (end_bit - begin_bit >= (int)(8 * sizeof(K))) ? ~(K)0 : ((((K)1) << (end_bit - begin_bit)) - 1);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        return ((kin[a] >> begin_bit) & mask) < ((kin[b] >> begin_bit) & mask);
    });
    for (size_t i = 0; i < n; i++) { kout[i] = kin[order[i]]; vout[i] = vin[order[i]]; }
    (void)P;
    return CRGPU_OK;
#else
    if (n >= ((size_t)1 << 32)) return CRGPU_ERR_ARG;
    const uint32_t ntiles = cr_div_up(n, RS_TILE);
    const size_t ncounts = (size_t)256 * ntiles;
    CR_TRY(P.temp.reserve(ncounts * 4 + 64));
    uint32_t* counts = P.temp.as<uint32_t>();
    const int bits = end_bit - begin_bit;
    const int passes = bits <= 0 ? 0 : (bits + 7) / 8;
    if (passes == 0) {
        CR_CUDA(cudaMemcpyAsync(kout, kin, n * sizeof(K), cudaMemcpyDeviceToDevice, P.stream));
        CR_CUDA(cudaMemcpyAsync(vout, vin, n * 4, cudaMemcpyDeviceToDevice, P.stream));
        return CRGPU_OK;
    }
    K* kalt = nullptr; uint32_t* valt = nullptr;
    if (passes > 1) {
        CR_TRY(P.pk.reserve(n * sizeof(K) + 16)); CR_TRY(P.pv.reserve(n * 4 + 16));
        kalt = P.pk.as<K>(); valt = P.pv.as<uint32_t>();
    }
    const K* ks = kin; const uint32_t* vs = vin;
    for (int p = 0; p < passes; p++) {
        const int shift = begin_bit + 8 * p;
        const int w = bits - 8 * p < 8 ? bits - 8 * p : 8;
        const uint32_t mask = (1u << w) - 1u;
        CR_LAUNCH(k_rs_hist<K>, dim3(ntiles), dim3(RS_THREADS), P.stream, ks, n, shift, mask, ntiles, counts);
        CR_TRY(cr_exclusive_sum(P, counts, counts, ncounts));
        const bool to_out = ((passes - 1 - p) & 1) == 0;          // the last pass writes the output
        K* kd = to_out ? kout : kalt; uint32_t* vd = to_out ? vout : valt;
        __atomic_fetch_add(&g_cr_launches, 1ull, __ATOMIC_RELAXED);
        RsScatter<K>::launch(ks, vs, kd, vd, n, shift, mask, ntiles, counts, P.stream);
        CR_CUDA(cudaGetLastError());
        ks = kd; vs = vd;
    }
    return CRGPU_OK;
#endif
}
