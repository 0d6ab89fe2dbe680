// cr_filter.cuh -- executable / bitmap pre-filters (the -F switch).
//
// Replaces filter_inplace (src/cr-filter.c:33-73) and its three sub-filters
//   pe_i386_transform  src/filter_x86_pe.c:128-159      elf_i386_transform src/filter_x86_elf.c:131-156
//   bmp_transform      src/filter_bmp.c:149-204         i386_e8e9          src/filter_x86opcode.h:38-62
// Split of work:
//   GPU  k_filter_scan      every byte position is tested for the 2/4-byte magics that can start an image
//   host FilterHost::walk   the reference's tiny sticky state machine runs over those few candidates and the
//                           image headers (a few dozen bytes per image) and emits a list of transform ops
//   GPU  e8e9 ops           the E8/E9 skip automaton is resolved exactly with cr_chain.cuh (a hit skips its
//                           4 operand bytes), then every live CALL/JMP operand is rewritten in parallel
//   GPU  bmp ops            colour + left + up delta in closed form per byte from the untouched tile
// State that the reference keeps in function-local statics (lastproc, flag/curr/imsz, BMP geometry) lives in
// FilterHost and carries across blocks and windows exactly like the statics do.
#pragma once
#include <algorithm>
#include <vector>
#include "cr_common.cuh"
#include "cr_chain.cuh"
#include "cr_rc.cuh"      // CopyDesc / k_copy_segments

// ------------------------------------------------------------------ candidate scan
// One thread tests the 16 positions of one aligned 16-byte chunk: a SWAR zero-byte test rejects chunks that hold none of the
// three first bytes ('M', 'B', 0x7F); only the others look at their positions one by one.  d must be 16-byte aligned and
// readable for 4 bytes past the chunk that holds position n - 1 (the window buffers carry 128 bytes of slack).
CR_D uint32_t fs_has(uint32_t w, uint32_t b) { const uint32_t x = w ^ (b * 0x01010101u); return (x - 0x01010101u) & ~x & 0x80808080u; }
// Positions skip .. n - 1 of d are tested and reported relative to skip (skip < 16 aligns a window that does not start on a 16-byte boundary).
__global__ void k_filter_scan(const uint8_t* __restrict__ d, uint32_t skip, uint64_t n, uint32_t* __restrict__ list, uint32_t cap, uint32_t* __restrict__ count) {
    const uint64_t p0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (p0 + 1 >= n) return;
    const uint4 v = *(const uint4*)(d + p0);
    const uint32_t w[4] = { v.x, v.y, v.z, v.w };
    uint32_t any = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) any |= fs_has(w[k], 'M') | fs_has(w[k], 'B') | fs_has(w[k], 0x7F);
    if (!any) return;
    const uint32_t w5[5] = { v.x, v.y, v.z, v.w, *(const uint32_t*)(d + p0 + 16) };
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t m = fs_has(w[k], 'M') | fs_has(w[k], 'B') | fs_has(w[k], 0x7F);          // bit 7 of every candidate byte (may over-report, never under)
        while (m) {
            const uint32_t b = (uint32_t)(__ffs((int)m) - 1) >> 3;                        // byte within the word
            m &= m - 1;
            const uint32_t q = (uint32_t)((((unsigned long long)w5[k + 1] << 32) | w5[k]) >> (8 * b));   // the four bytes at the position
            const uint64_t p = p0 + 4 * k + b;
            const uint32_t a = q & 255u, c = (q >> 8) & 255u;
            const bool hit = p >= skip && p + 1 < n && ((a == 'M' && c == 'Z') || (a == 'B' && c == 'M') || (p + 3 < n && q == 0x464C457Fu));
            if (hit) { uint32_t i = atomicAdd(count, 1u); if (i < cap) list[i] = (uint32_t)(p - skip); }
        }
    }
}

// ------------------------------------------------------------------ E8/E9
struct E8Op {
    uint64_t off;        // window offset of the region's byte 0
    uint32_t limit;      // region length the reference passes to i386_e8e9
    uint32_t valid;      // bytes of the region that really exist in the block (the ELF detection call runs 52 bytes past it)
    int32_t  ncur, nend;
    uint64_t span_off;   // offset of the region in the op-local span array
};
CR_D uint32_t e8_byte(const uint8_t* d, const E8Op& o, uint32_t i) { return i < o.valid ? d[o.off + i] : 0u; }

__global__ void k_e8e9_spans(const uint8_t* __restrict__ d, const E8Op* __restrict__ ops, uint8_t* __restrict__ span) {
    const E8Op o = ops[blockIdx.y];
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= o.limit) return;
    span[o.span_off + i] = (i < o.limit - 8 && (e8_byte(d, o, i) & 254) == 0xe8) ? 5 : 1;      // src/filter_x86opcode.h:41-42,58
}
struct E8Apply {
    typedef int State;
    uint8_t* d; const E8Op* ops; const uint8_t* span; int en_de;
    CR_D State begin(uint32_t, uint32_t) const { return 0; }
    CR_D void visit(State&, uint32_t s, uint32_t i, uint32_t sp) const {
        if (sp != 5) return;
        const E8Op o = ops[s];
        const uint32_t at = i + 1;                               // index of the operand (the reference's i after i++)
        int32_t op = (int32_t)(e8_byte(d, o, at) | e8_byte(d, o, at + 1) << 8 | e8_byte(d, o, at + 2) << 16 | e8_byte(d, o, at + 3) << 24);
        const int32_t pos = o.ncur + (int32_t)at;
        if (en_de == 0) {
            if (op >= -pos && op < o.nend - pos) op = (int32_t)((uint32_t)op + (uint32_t)pos);      // :46-47
            else if (op > 0 && op < o.nend) op = (int32_t)((uint32_t)op - (uint32_t)o.nend);        // :48-49
            else return;
        } else {
            if (op < 0) { if ((int32_t)((uint32_t)op + (uint32_t)pos) >= 0) op = (int32_t)((uint32_t)op + (uint32_t)o.nend); else return; }   // :51-53
            else if (op < o.nend) op = (int32_t)((uint32_t)op - (uint32_t)pos);                      // :54-55
            else return;
        }
        for (uint32_t k = 0; k < 4; k++) if (at + k < o.valid) d[o.off + at + k] = (uint8_t)((uint32_t)op >> (8 * k));
    }
    CR_D void end(State&, uint32_t, uint32_t) const {}
};

// ------------------------------------------------------------------ BMP
struct BmpOp {
    uint64_t off;        // window offset of the tile
    uint64_t src_off;    // offset of the tile's untouched copy
    uint32_t rows, width, row_size, bytes;   // bytes per pixel (3 or 4)
};
// Forward transform, 16 output bytes per thread (one aligned 16-byte chunk of the window).  The untouched copy of the tile sits
// in `snap` at an address congruent to the window address mod 16 (k_copy_chunks), so the chunk, its two neighbours and the
// 28-byte window one row up are aligned vector / word loads.  Per byte (row y, byte xb of the row, channel ch = xb % BPP):
//   c(y, xb)  = b - g for ch 0 and 2, b otherwise                                  colour       src/filter_bmp.c:63-73
//   h(y, xb)  = c(y, xb) - c(y, xb - BPP)          for xb >= BPP                   left delta   :75-88
//   out       = h(y, xb) - h(y - 1, xb)            for y > 0                       up delta     :89-102
// Bytes of the row padding keep their value.  Chunks that straddle the ends of the tile write their own bytes one by one
// (the neighbouring bytes may belong to another tile whose thread writes them).
__global__ void k_copy_chunks(const CopyDesc* __restrict__ descs, const uint8_t* __restrict__ src, uint8_t* __restrict__ dst) {
    const CopyDesc c = descs[blockIdx.y];                                     // c.dst == c.src (mod 16)
    const uint64_t a0 = c.src & ~15ull, n = ((c.src + c.len + 15) & ~15ull) - a0 >> 4;
    const uint4* s = (const uint4*)(src + a0);
    uint4* t = (uint4*)(dst + (c.dst - (c.src - a0)));
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) t[i] = s[i];
}
CR_D uint32_t cr_funnel_r(uint32_t lo, uint32_t hi, uint32_t bits) { return (uint32_t)((((unsigned long long)hi << 32) | lo) >> bits); }
// byte-wise a - b (mod 256) on four packed bytes
CR_D uint32_t bmp_sub4(uint32_t a, uint32_t b) { return ((a | 0x80808080u) - (b & 0x7f7f7f7fu)) ^ ((a ^ ~b) & 0x80808080u); }
// 0xFF in every byte b of a word whose channel (phase + b) mod BPP equals `want`
template <int BPP> CR_D uint32_t bmp_chmask(uint32_t phase, uint32_t want) {
    uint32_t m = 0;
#pragma unroll
    for (uint32_t b = 0; b < 4; b++) if ((phase + b) % BPP == want) m |= 255u << (8 * b);
    return m;
}
template <int BPP>
CR_D void bmp_chunk(const uint8_t* __restrict__ snap, uint8_t* __restrict__ d, const BmpOp& o, const uint64_t chunk) {
    const int64_t total = (int64_t)o.rows * o.row_size;
    const uint64_t A = (o.off & ~15ull) + 16 * chunk;                         // window address of the chunk
    const int64_t idx0 = (int64_t)A - (int64_t)o.off;                         // tile index of its byte 0 (negative only in the first chunk)
    if (idx0 >= total) return;
    const uint8_t* sp = snap + (int64_t)o.src_off + idx0;                     // 16-byte aligned by construction
    uint32_t cur[12], up[7];                                                  // tile bytes idx0 - 16 .. idx0 + 31;  idx0 - row_size - 8 .. + 19
    { const uint4 a = *(const uint4*)(sp - 16), b = *(const uint4*)sp, c = *(const uint4*)(sp + 16);
      cur[0] = a.x; cur[1] = a.y; cur[2] = a.z; cur[3] = a.w; cur[4] = b.x; cur[5] = b.y; cur[6] = b.z; cur[7] = b.w; cur[8] = c.x; cur[9] = c.y; cur[10] = c.z; cur[11] = c.w; }
    const bool has_up = idx0 + 15 >= (int64_t)o.row_size;
#pragma unroll
    for (int k = 0; k < 7; k++) up[k] = has_up ? ((const uint32_t*)(sp - o.row_size - 8))[k] : 0u;
    auto CB = [&](int r) -> uint32_t { return (cur[(r + 16) >> 2] >> (((r + 16) & 3) * 8)) & 255u; };
    auto UB = [&](int r) -> uint32_t { return (up[(r + 8) >> 2] >> (((r + 8) & 3) * 8)) & 255u; };
    const uint32_t wb = o.width * BPP;
    const int64_t first = idx0 < 0 ? 0 : idx0;
    uint32_t y = (uint32_t)(first / o.row_size), xb = (uint32_t)(first % o.row_size), ch = xb % BPP;
    uint32_t out[4] = { cur[4], cur[5], cur[6], cur[7] };
    const bool interior = idx0 >= 0 && idx0 + 16 <= total;
    if (interior && xb >= 8 && xb + 16 <= wb) {
        // All 16 bytes are pixel bytes of one row with their left neighbours in the same row: four bytes per operation.
        // Word i covers bytes 4i .. 4i+3 of the chunk; the channel of byte b of word i is (ch + 4i + b) mod BPP.
        uint32_t C[5], UC[5];                                                   // colour-decorrelated words -1 .. 3, this row and the row above
#pragma unroll
        for (int i = -1; i < 4; i++) {
            const uint32_t ph = (ch + (uint32_t)(4 * i + 4 * BPP)) % BPP;
            const uint32_t m0 = bmp_chmask<BPP>(ph, 0), m2 = bmp_chmask<BPP>(ph, 2);
            const uint32_t o = cur[i + 4], op = cr_funnel_r(cur[i + 4], cur[i + 5], 8), om = cr_funnel_r(cur[i + 3], cur[i + 4], 24);
            C[i + 1] = bmp_sub4(o, (op & m0) | (om & m2));
            const uint32_t u = up[i + 2], upl = cr_funnel_r(up[i + 2], up[i + 3], 8), umi = cr_funnel_r(up[i + 1], up[i + 2], 24);
            UC[i + 1] = bmp_sub4(u, (upl & m0) | (umi & m2));
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t cl = BPP == 4 ? C[i] : cr_funnel_r(C[i], C[i + 1], 8);        // the same channel one pixel to the left
            const uint32_t ul = BPP == 4 ? UC[i] : cr_funnel_r(UC[i], UC[i + 1], 8);
            const uint32_t h = bmp_sub4(C[i + 1], cl);
            out[i] = y > 0 ? bmp_sub4(h, bmp_sub4(UC[i + 1], ul)) : h;
        }
        *(uint4*)(d + A) = make_uint4(out[0], out[1], out[2], out[3]);
        return;
    }
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int64_t idx = idx0 + k;
        if (idx < 0 || idx >= total) continue;
        if (xb < wb) {
            auto C = [&](int r) -> uint32_t { return ch == 0 ? CB(r) - CB(r + 1) : ch == 2 ? CB(r) - CB(r - 1) : CB(r); };
            auto U = [&](int r) -> uint32_t { return ch == 0 ? UB(r) - UB(r + 1) : ch == 2 ? UB(r) - UB(r - 1) : UB(r); };
            uint32_t v = C(k);
            if (xb >= BPP) v -= C(k - BPP);
            if (y > 0) { v -= U(k); if (xb >= BPP) v += U(k - BPP); }
            v &= 255u;
            if (interior) out[k >> 2] = (out[k >> 2] & ~(255u << ((k & 3) * 8))) | v << ((k & 3) * 8);
            else d[A + k] = (uint8_t)v;
        }
        if (++xb == o.row_size) { xb = 0; ch = 0; y++; } else if (++ch == BPP) ch = 0;
    }
    if (interior) *(uint4*)(d + A) = make_uint4(out[0], out[1], out[2], out[3]);
}
__global__ void k_bmp_rows(const uint8_t* __restrict__ snap, uint8_t* __restrict__ d, const BmpOp* __restrict__ ops) {
    const BmpOp o = ops[blockIdx.y];
    const uint64_t nchunk = ((o.off & 15) + (uint64_t)o.rows * o.row_size + 15) >> 4;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nchunk; c += (uint64_t)gridDim.x * blockDim.x) {
        if (o.bytes == 3) bmp_chunk<3>(snap, d, o, c); else bmp_chunk<4>(snap, d, o, c);
    }
}

// inverse (src/filter_bmp.c:104-145): horizontal prefix sums, vertical prefix sums, colour re-correlation, in place
__global__ void k_bmp_dec_rows(uint8_t* __restrict__ d, const BmpOp* __restrict__ ops) {
    const BmpOp o = ops[blockIdx.y];
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;             // (row, channel)
    if (t >= o.rows * o.bytes) return;
    uint8_t* p = d + o.off + (uint64_t)(t / o.bytes) * o.row_size + t % o.bytes;
    uint32_t acc = p[0];
    for (uint32_t x = 1; x < o.width; x++) { acc += p[(size_t)x * o.bytes]; p[(size_t)x * o.bytes] = (uint8_t)acc; }
}
__global__ void k_bmp_dec_cols(uint8_t* __restrict__ d, const BmpOp* __restrict__ ops) {
    const BmpOp o = ops[blockIdx.y];
    const uint32_t xb = blockIdx.x * blockDim.x + threadIdx.x;
    if (xb >= o.width * o.bytes) return;
    uint8_t* p = d + o.off + xb;
    uint32_t acc = p[0];
    for (uint32_t y = 1; y < o.rows; y++) { acc += p[(uint64_t)y * o.row_size]; p[(uint64_t)y * o.row_size] = (uint8_t)acc; }
}
__global__ void k_bmp_dec_colour(uint8_t* __restrict__ d, const BmpOp* __restrict__ ops) {
    const BmpOp o = ops[blockIdx.y];
    const uint64_t total = (uint64_t)o.rows * o.width;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
        uint8_t* p = d + o.off + (t / o.width) * o.row_size + (t % o.width) * o.bytes;
        p[0] = (uint8_t)(p[0] + p[1]); p[2] = (uint8_t)(p[2] + p[1]);
    }
}

// ------------------------------------------------------------------ host state machine
struct FilterHost {
    int lastproc = 0;                                   // 0 none, 1 pe, 2 elf, 3 bmp (src/cr-filter.c:41)
    struct { int flag; uint32_t curr, imsz; } pe = {0, 0, 0}, elf = {0, 0, 0};
    struct { int flag, curr, size, row_size, bpp, width, height, skip_size; } bmp = {0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<E8Op> e8ops; std::vector<BmpOp> bmpops;
    DevBuf b_list, b_count, b_e8, b_bmp, b_span, b_segs, b_xt, b_entry, b_tmp, b_copy;

    void reset() { lastproc = 0; pe = {0, 0, 0}; elf = {0, 0, 0}; bmp = {0, 0, 0, 0, 0, 0, 0, 0}; }
    void release() { DevBuf* all[] = { &b_list, &b_count, &b_e8, &b_bmp, &b_span, &b_segs, &b_xt, &b_entry, &b_tmp, &b_copy }; for (DevBuf* b : all) b->release(); }

    // view of one block for header parsing: bytes past the block read as 0 ("filter envelope")
    struct View {
        const uint8_t* p; uint32_t len;
        uint32_t u8(uint32_t i) const { return i < len ? p[i] : 0; }
        uint32_t u16(uint32_t i) const { return u8(i) | u8(i + 1) << 8; }
        uint32_t u32(uint32_t i) const { return u16(i) | u16(i + 2) << 16; }
    };
    static uint32_t umin(uint32_t a, uint32_t b) { return a < b ? a : b; }

    void push_e8(uint64_t off, uint32_t limit, uint32_t valid, uint32_t ncur, uint32_t nend) {
        if (limit < 8) return;                          // the reference's loop bound wraps here (undefined); we transform nothing
        E8Op o; o.off = off; o.limit = limit; o.valid = umin(valid, limit); o.ncur = (int32_t)ncur; o.nend = (int32_t)nend; o.span_off = 0;
        e8ops.push_back(o);
    }
    uint32_t run_elf(const View& v, uint64_t woff) {                                              // src/filter_x86_elf.c:131-156
        uint32_t size = umin(elf.imsz - elf.curr, v.len), start = 0;
        if (!elf.flag) {
            if (v.len < 52 || v.u32(0) != 0x464C457Fu || v.u16(18) != 3) return 0;
            uint32_t shoff = v.u32(32), est = shoff - 52;
            if (shoff < 52 || est >= (1u << 30)) return 0;
            elf.imsz = est - 52; start = 52; size = umin(elf.imsz, v.len);
        }
        push_e8(woff + start, size, v.len > start ? v.len - start : 0, elf.curr, elf.imsz);
        elf.curr += size; elf.flag = elf.curr < elf.imsz;
        return size;
    }
    uint32_t run_pe(const View& v, uint64_t woff) {                                               // src/filter_x86_pe.c:75-159
        uint32_t size = umin(pe.imsz - pe.curr, v.len), ret = size, start = 0;
        if (!pe.flag) {
            pe.curr = 0;
            if (v.len < 0x3C + 4 || v.u16(0) != 0x5A4D) return 0;
            uint32_t hdr = v.u32(0x3C);
            if (hdr >= v.len || v.u32(hdr) != 0x00004550u || hdr == 0) return 0;
            if (hdr + v.len < 24) return 0;
            uint32_t machine = v.u16(hdr + 4), nsec = v.u16(hdr + 6), optsz = v.u16(hdr + 20), chars = v.u16(hdr + 22);
            if (machine != 0x14c && (chars & 2)) return 0;
            uint32_t sec_off = 24 + optsz, size_hdr = sec_off + nsec * 40, est = size_hdr;
            if (hdr + v.len < size_hdr) return 0;
            for (uint32_t i = 0; i < nsec; i++) est += v.u32(hdr + sec_off + i * 40 + 16);
            if (est > (1u << 28)) return 0;
            start = size_hdr; pe.imsz = est - size_hdr;
            size = umin(pe.imsz, v.len - size_hdr);
            ret = size + size_hdr;
        }
        push_e8(woff + start, size, v.len > start ? v.len - start : 0, pe.curr, pe.imsz);
        pe.curr += size; pe.flag = pe.curr < pe.imsz;
        return ret;
    }
    uint32_t run_bmp(const View& v, uint64_t woff) {                                              // src/filter_bmp.c:149-204
        if (!bmp.flag) {
            if (v.len < 54 || v.u16(0) != 0x4d42 || v.u16(26) != 1 || v.u32(30) != 0) return 0;
            uint32_t fsize = v.u32(2), off = v.u32(10), isz = v.u32(34), bpp = v.u16(28);
            if ((isz != 0 && off + isz != fsize) || (bpp != 24 && bpp != 32)) return 0;
            int w = (int32_t)v.u32(18), h = (int32_t)v.u32(22);
            bmp.width = w < 0 ? -w : w; bmp.height = h < 0 ? -h : h;
            bmp.row_size = ((int)bpp * bmp.width + 31) / 32 * 4; bmp.bpp = (int)bpp;
            if (bmp.width < 4 || bmp.height < 4 || bmp.width >= (1 << 20) || bmp.height >= (1 << 20)) return 0;
            bmp.curr = (int)off; bmp.size = bmp.height * bmp.row_size; bmp.skip_size = 0; bmp.flag = 1;
            return off;
        }
        if (bmp.skip_size > 0) { uint32_t t = umin((uint32_t)bmp.skip_size, v.len); bmp.curr += (int)t; bmp.skip_size -= (int)t; return t; }
        uint32_t avail = umin(v.len, (uint32_t)(bmp.size - bmp.curr));
        uint32_t rows = avail / (uint32_t)bmp.row_size, t = rows * (uint32_t)bmp.row_size;
        if (rows) { BmpOp o; o.off = woff; o.src_off = 0; o.rows = rows; o.width = (uint32_t)bmp.width; o.row_size = (uint32_t)bmp.row_size; o.bytes = (uint32_t)bmp.bpp / 8; bmpops.push_back(o); }
        bmp.curr += (int)t;
        if (bmp.curr < bmp.size) bmp.skip_size = (int)umin((uint32_t)bmp.row_size, (uint32_t)(bmp.size - bmp.curr)); else bmp.flag = 0;
        return t;
    }
    uint32_t run(int which, const View& v, uint64_t woff) { return which == 1 ? run_pe(v, woff) : which == 2 ? run_elf(v, woff) : run_bmp(v, woff); }

    // filter_inplace over one block (src/cr-filter.c:50-71); cand = sorted candidate offsets (window relative)
    int walk_block(const uint8_t* h_block, uint64_t boff, uint32_t len, const std::vector<uint32_t>& cand) {
        int filt = 0;
        size_t ci = std::lower_bound(cand.begin(), cand.end(), (uint32_t)boff) - cand.begin();
        for (uint32_t pos = 0; pos < len;) {
            View v = { h_block + pos, len - pos };
            if (lastproc) {
                uint32_t n = run(lastproc, v, boff + pos);
                if ((int)n == 0) lastproc = 0; else { filt = 1; pos += n; continue; }
            }
            // A sub-filter whose sticky flag is set (an image in progress) needs no magic to fire: after bmp_transform has
            // returned 0 for a broken row, the loop over all sub-filters calls it again at the SAME position and it skips
            // the row (src/filter_bmp.c:188-203, src/cr-filter.c:60-68).  Only with all flags clear may we jump to the next magic.
            if (!(pe.flag || elf.flag || bmp.flag)) {
                while (ci < cand.size() && cand[ci] < boff + pos) ci++;
                if (ci >= cand.size() || cand[ci] >= boff + len) break;
                pos = (uint32_t)(cand[ci] - boff);
                v.p = h_block + pos; v.len = len - pos;
            }
            bool fired = false;
            for (int k = 1; k <= 3 && !fired; k++) {
                uint32_t n = run(k, v, boff + pos);
                if ((int)n > 0) { filt = 1; lastproc = k; pos += n; fired = true; }
            }
            if (!fired) pos++;
        }
        return filt;
    }

    template <class Chain>
    int run_window(Chain& C, const uint8_t* h_win, uint8_t* d_win, uint64_t nwin_total, const std::vector<uint64_t>& roff, const std::vector<uint32_t>& rsize,
                   std::vector<uint8_t>& flags, int& filt_flag, int en_de = 0) {
        cudaStream_t stream = C.stream;
        uint64_t wlen = 0;
        for (size_t b = 0; b < roff.size(); b++) if (roff[b] + rsize[b] > wlen) wlen = roff[b] + rsize[b];
        (void)nwin_total;
        const uint32_t mis = (uint32_t)((uintptr_t)d_win & 15);          // the vector kernels address the window from the 16-byte boundary below it
        // ---- candidates
        std::vector<uint32_t> cand;
        if (wlen >= 2) {
            uint32_t cap = (uint32_t)(wlen / 64 + 4096);
            for (;;) {
                CR_TRY(b_list.reserve((size_t)cap * 4)); CR_TRY(b_count.reserve(16));
                CR_CUDA(cudaMemsetAsync(b_count.p, 0, 4, stream));
                CR_LAUNCH(k_filter_scan, dim3(cr_div_up(cr_div_up(wlen + mis, 16), 256)), dim3(256), stream, d_win - mis, mis, wlen + mis, b_list.as<uint32_t>(), cap, b_count.as<uint32_t>());
                std::vector<uint32_t> cnt;
                CR_TRY(C.download(cnt, b_count.p, 1));
                if (cnt[0] <= cap) { CR_TRY(C.download(cand, b_list.p, cnt[0])); break; }
                cap = cnt[0] + 16;
            }
            std::sort(cand.begin(), cand.end());
        }
        // ---- the reference's state machine, block by block
        e8ops.clear(); bmpops.clear();
        for (size_t b = 0; b < roff.size(); b++) { filt_flag = walk_block(h_win + roff[b], roff[b], rsize[b], cand); flags[b] = (uint8_t)filt_flag; }
        // ---- E8/E9 regions
        if (!e8ops.empty()) {
            std::vector<ChainSeg> segs(e8ops.size());
            uint64_t so = 0; uint32_t maxlimit = 0;
            for (size_t i = 0; i < e8ops.size(); i++) {
                e8ops[i].span_off = so;
                segs[i].off = so; segs[i].len = e8ops[i].limit; segs[i].start = 0;
                so += e8ops[i].limit; if (e8ops[i].limit > maxlimit) maxlimit = e8ops[i].limit;
            }
            const uint32_t nseg = (uint32_t)segs.size(), nchunk = cr_chain_layout(segs.data(), nseg);
            CR_TRY(C.upload(b_e8, e8ops)); CR_TRY(C.upload(b_segs, segs));
            CR_TRY(b_span.reserve(so + 16)); CR_TRY(b_xt.reserve((size_t)nchunk * 256 + 16)); CR_TRY(b_entry.reserve(nchunk + 16));
            CR_LAUNCH(k_e8e9_spans, dim3(cr_div_up(maxlimit, 256), nseg), dim3(256), stream, d_win, b_e8.as<E8Op>(), b_span.as<uint8_t>());
            CR_LAUNCH(k_chain_exits, dim3(cr_div_up(nchunk, 64)), dim3(64), stream, b_span.as<uint8_t>(), b_segs.as<ChainSeg>(), nseg, nchunk, b_xt.as<uint8_t>());
            CR_TRY(cr_chain_run_entries(C, C.b_chainwork, segs, b_segs.as<ChainSeg>(), nchunk, b_xt.as<uint8_t>(), b_entry.as<uint8_t>()));
            E8Apply f = { d_win, b_e8.as<E8Op>(), b_span.as<uint8_t>(), en_de };
            CR_LAUNCH(k_chain_walk<E8Apply>, dim3(cr_div_up(nchunk, 64)), dim3(64), stream, b_span.as<uint8_t>(), b_segs.as<ChainSeg>(), nseg, nchunk, b_entry.as<uint8_t>(), f);
        }
        // ---- BMP tiles: snapshot the tiles, then transform from the snapshot
        if (!bmpops.empty() && en_de) {
            uint32_t maxrc = 0, maxwb = 0;
            for (auto& o : bmpops) { if (o.rows * o.bytes > maxrc) maxrc = o.rows * o.bytes; if (o.width * o.bytes > maxwb) maxwb = o.width * o.bytes; }
            CR_TRY(C.upload(b_bmp, bmpops));
            CR_LAUNCH(k_bmp_dec_rows, dim3(cr_div_up(maxrc, 128), (unsigned)bmpops.size()), dim3(128), stream, d_win, b_bmp.as<BmpOp>());
            CR_LAUNCH(k_bmp_dec_cols, dim3(cr_div_up(maxwb, 128), (unsigned)bmpops.size()), dim3(128), stream, d_win, b_bmp.as<BmpOp>());
            CR_LAUNCH(k_bmp_dec_colour, dim3(296, (unsigned)bmpops.size()), dim3(256), stream, d_win, b_bmp.as<BmpOp>());
        } else if (!bmpops.empty()) {
            // snapshots: 32 bytes of slack on both sides (the kernel reads the neighbouring chunks and 8 bytes in front of the
            // row above), each placed congruent to its window address mod 16
            std::vector<CopyDesc> copies(bmpops.size());
            uint64_t to = 0, maxchunk = 0;
            for (size_t i = 0; i < bmpops.size(); i++) {
                bmpops[i].off += mis;                                                       // relative to d_win - mis from here on
                const uint64_t bytes = (uint64_t)bmpops[i].rows * bmpops[i].row_size, mo = bmpops[i].off & 15;
                bmpops[i].src_off = to + 32 + mo;
                CopyDesc c = { bmpops[i].off, bmpops[i].src_off, (uint32_t)bytes, 0 }; copies[i] = c;
                to += 32 + ((mo + bytes + 15) & ~15ull) + 32;
                maxchunk = std::max<uint64_t>(maxchunk, (mo + bytes + 15) >> 4);
            }
            CR_TRY(b_tmp.reserve(to + 16)); CR_TRY(C.upload(b_copy, copies)); CR_TRY(C.upload(b_bmp, bmpops));
            const unsigned gx = (unsigned)std::min<uint64_t>(cr_div_up(maxchunk, 256), 148 * 8);
            CR_LAUNCH(k_copy_chunks, dim3(gx, (unsigned)copies.size()), dim3(256), stream, b_copy.as<CopyDesc>(), d_win - mis, b_tmp.as<uint8_t>());
            CR_LAUNCH(k_bmp_rows, dim3(gx, (unsigned)bmpops.size()), dim3(256), stream, b_tmp.as<uint8_t>(), d_win - mis, b_bmp.as<BmpOp>());
        }
        return CRGPU_OK;
    }
};
