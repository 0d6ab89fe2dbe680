// cr_lzchain.cuh -- host orchestration of one window of an lzencode chain (ROLZ or LZP + PPM + range coder).
//
// A "chain" is a run of blocks that share adaptive model state: the dictionary payload is a
// chain of one block; all data blocks of a container form the second chain.  A window is a group of
// consecutive blocks of a chain that is resident in HBM at once.  Everything model-independent (histograms,
// match finding, parse resolution) is parallel over all positions of the window; the model passes are
// parallel over contexts; only the range coders are serial, one per (block, stream).
#pragma once
#include <vector>
#include <algorithm>
#include "cr_common.cuh"
#include "cr_prims.cuh"
#include "cr_chain.cuh"
#include "cr_rolz.cuh"
#include "cr_lzp.cuh"
#include "cr_lz77.cuh"
#include "cr_ppm.cuh"
#include "cr_rc.cuh"
#include "cr_warp.cuh"
#include "cr_rcpar.cuh"

enum { CR_ROLZ = 0, CR_LZP = 1, CR_LZ77 = 2 };

struct BlockIO {
    uint64_t off;        // in: offset of the dictionary-coded block in the window buffer
    uint32_t size;       // in
    uint8_t  filt, prec; // in: bytes 4,5 of the container block header (prefix_mode 1)
    uint64_t out_off;    // out: offset of prefix+payload in the output buffer
    uint32_t out_size;   // out: payload size (without prefix)
    uint32_t raw;        // out: 1 = stored ("cannot compress")
    uint32_t cut;        // in: 0, or the position at which the reference's coder loop gives up on this block (encode_blocks sets it
                         //     for the re-run after a mid-chain "cannot compress"): only tokens before it reach models and context
};

struct LzChain {
    cudaStream_t stream = 0;
    Prims prims;
    int variant = CR_ROLZ;
#ifndef CRGPU_SIM
    cudaStream_t side_stream = 0;       // the order-0 side-model chain runs beside the PPM passes
    cudaEvent_t ev_side_go = 0, ev_side_done = 0;
#endif
    // ---- persistent model state (the reference's file-scope `m`, src/rolzmain/cr-coder.c:52-56)
    DevBuf s_o3b, s_o3c, s_o2, s_o1, s_m0;
    PpmState st;
    uint32_t chain_ctx = 0;
    bool inited = false;
    // ---- window work buffers (kept between calls)
    DevBuf b_blocks, b_segoff, b_seglen, b_hist, b_esc1, b_first, b_ctxout;
    DevBuf b_k0, b_k1, b_v0, b_v1, b_ks0, b_M, b_S, b_span, b_tidx;
    DevBuf b_segs, b_xt, b_entry, b_cnt, b_scan, b_chainwork;
    DevBuf b_evctx, b_evsym, b_tokend, b_pred, b_T1, b_T2, b_TS, b_side;
    DevBuf b_escrec, b_esccount, b_k64a, b_k64b, b_ord, b_flag, b_escord;
    DevBuf b_lensym, b_lenpos, b_idxsym, b_idxpos;
    DevBuf b_hits;
    DevBuf b_o1ctx, b_o1steps, b_o1snaps, b_o2ctl, b_o2hot, b_o2todo, b_o2steps, b_o2snaps;
    DevBuf b_o1info, b_o1ord, b_o1incl, b_bounds, b_o3hot, b_cinm, b_cins, b_segstart, b_segkey, b_rank, b_flexlen;
    DevBuf b_qm, b_shm, b_bm, b_qs, b_shs, b_bs, b_stot, b_dsum, b_lsm, b_lss, b_rsm, b_rss, b_fb;
    DevBuf b_dense, b_denseside, b_streams, b_rcres, b_rcout, b_copy, b_hdr;
    // LZ77 (cr_lz77.cuh): per-position candidates, guesses and token distances; per-chunk `last` propagation; per-model symbol lists
    DevBuf x_npos, x_nlen, x_sdist, x_slen, x_G, x_tdist, x_h16, x_rank, x_mpos0, x_mlen0, x_clast, x_ckey, x_cmax, x_lastin, x_mism, x_msym, x_mpos;
    uint32_t match_limit = 40;     // -m (src/roxmain/cr-matcher.c:38)
    uint32_t x_max_iter = 48;      // fixed-point passes before the serial parse takes over (0 = always serial; tests)
    uint32_t last_x_iters = 0;
    // ---- sizes of the last window (for the debug/trace fetch used by the tests)
    uint32_t last_nev = 0, last_nside = 0, last_nesc = 0, last_nent = 0;
    size_t last_dtotal = 0;
    StageTimer timer;
    bool flexible = false;         // -f flexible parsing (ROLZ)
    int rc_variant = 8;            // range-chain formulation: 8 = cut into jobs that run side by side (cr_rcpar.cuh); 1..7 = one serial walk per stream (cr_warp.cuh: k_range_chain<1..7>)
    bool hot_contexts = true;      // hot o2 contexts run the rank-based CTA kernel (k_o2_pass_cta)
    int rolz_match_variant = 2;    // 2 = k_rolz_match_main2 (five-byte filter in shared memory), 1 = k_rolz_match_main
    int o1_hot_variant = 2;        // 2 = k_o1_skel + k_o1_eval (the chain of steps carries only the counts), 1 = k_o1_pass_cta
    int o2_hot_variant = 3;        // 3 = k_o2_skel + k_o2_eval (the chain of steps carries only counts and flags), 2 = k_o2_hot (event ring, ballot ranks), 1 = k_o2_pass_cta
    bool o2_attr_done = false;
    uint32_t o2_width = 0;         // events per step of the split o2 pass: 0 = by the window's hit rate (256 / 1024), or 256 / 512 / 1024
    uint32_t o2_rec_cap_test = 0;  // tests only: capacity of the step-record table (0 = the bound)
    bool scalar_models = false;   // GPU A/B switch: run the scalar (simulation-checked) model/coder kernels
    bool exact_aborts = true;      // replay a mid-chain "cannot compress" exactly (encode_blocks); false = CRGPU_ERR_MIDCHAIN_ABORT
    DevBuf b_abort;
#ifndef CRGPU_SIM
    RcPar rcpar;                   // rc_variant 8: the range chain cut into jobs (cr_rcpar.cuh)
#endif

    // the switches a second chain of the same handle must share (Compressor's dictionary-payload chain)
    void copy_options(const LzChain& o) {
        match_limit = o.match_limit; x_max_iter = o.x_max_iter; flexible = o.flexible; rc_variant = o.rc_variant; hot_contexts = o.hot_contexts;
        rolz_match_variant = o.rolz_match_variant; o1_hot_variant = o.o1_hot_variant; o2_hot_variant = o.o2_hot_variant; o2_rec_cap_test = o.o2_rec_cap_test; o2_width = o.o2_width;
        scalar_models = o.scalar_models; exact_aborts = o.exact_aborts;
#ifndef CRGPU_SIM
        rcpar.job_symbols = o.rcpar.job_symbols; rcpar.late_cfg = o.rcpar.late_cfg; rcpar.serial_only = o.rcpar.serial_only; rcpar.crowded = o.rcpar.crowded;
#endif
    }
    int init(int variant_, cudaStream_t s) {
        variant = variant_; stream = s; prims.stream = s;
        CR_TRY(s_o3b.reserve(PPM_O3_SLOTS)); CR_TRY(s_o3c.reserve(PPM_O3_SLOTS));
        CR_TRY(s_o2.reserve((size_t)65536 * PPM_O2_STRIDE)); CR_TRY(s_o1.reserve(65536)); CR_TRY(s_m0.reserve(X_NMODEL * 256 * 2));
        st.o3_byte = s_o3b.as<uint8_t>(); st.o3_conf = s_o3c.as<uint8_t>(); st.o2 = s_o2.as<uint8_t>();
        st.o1 = s_o1.as<uint8_t>(); st.m0 = s_m0.as<uint16_t>();
#ifndef CRGPU_SIM
        if (!side_stream) {
            CR_CUDA(cudaStreamCreateWithFlags(&side_stream, cudaStreamNonBlocking));
            CR_CUDA(cudaEventCreateWithFlags(&ev_side_go, cudaEventDisableTiming));
            CR_CUDA(cudaEventCreateWithFlags(&ev_side_done, cudaEventDisableTiming));
        }
#endif
        inited = true;
        return reset_models();
    }
    void release() {
        DevBuf* all[] = { &s_o3b, &s_o3c, &s_o2, &s_o1, &s_m0, &b_blocks, &b_segoff, &b_seglen, &b_hist, &b_esc1, &b_first, &b_ctxout,
            &b_k0, &b_k1, &b_v0, &b_v1, &b_ks0, &b_M, &b_S, &b_span, &b_tidx, &b_segs, &b_xt, &b_entry, &b_cnt, &b_scan, &b_chainwork,
            &b_evctx, &b_evsym, &b_tokend, &b_pred, &b_T1, &b_T2, &b_TS, &b_side, &b_escrec, &b_esccount, &b_k64a, &b_k64b, &b_ord,
            &b_flag, &b_escord, &b_lensym, &b_lenpos, &b_idxsym, &b_idxpos, &b_hits, &b_o1ctx, &b_o1steps, &b_o1snaps, &b_o2ctl, &b_o2hot, &b_o2todo, &b_o2steps, &b_o2snaps, &b_o1info, &b_o1ord, &b_o1incl, &b_bounds, &b_o3hot, &b_cinm, &b_cins, &b_segstart, &b_segkey, &b_rank, &b_flexlen, &b_qm, &b_shm, &b_bm, &b_qs, &b_shs, &b_bs, &b_stot, &b_dsum, &b_lsm, &b_lss, &b_rsm, &b_rss, &b_fb, &b_dense, &b_denseside, &b_streams, &b_rcres, &b_rcout, &b_copy, &b_hdr,
            &x_npos, &x_nlen, &x_sdist, &x_slen, &x_G, &x_tdist, &x_h16, &x_rank, &x_mpos0, &x_mlen0, &x_clast, &x_ckey, &x_cmax, &x_lastin, &x_mism, &x_msym, &x_mpos,
            &snap_o3b, &snap_o3c, &snap_o2, &snap_o1, &snap_m0, &b_abort };
        for (DevBuf* b : all) b->release();
        prims.release();
#ifndef CRGPU_SIM
        rcpar.release();
        if (side_stream) { cudaStreamDestroy(side_stream); cudaEventDestroy(ev_side_go); cudaEventDestroy(ev_side_done); side_stream = 0; }
#endif
        inited = false;
    }
    // reset_models(): src/rolzmain/cr-coder.c:78-96 / src/ropmain/cr-coder.c:73-83
    int reset_models() {
        CR_CUDA(cudaMemsetAsync(st.o3_byte, 0, PPM_O3_SLOTS, stream));
        CR_CUDA(cudaMemsetAsync(st.o3_conf, 0, PPM_O3_SLOTS, stream));
        CR_LAUNCH(k_ppm_reset, dim3(65536 / 256), dim3(256), stream, st);
        if (variant == CR_LZ77) CR_LAUNCH(k_x_reset_models, dim3(1), dim3(256), stream, st.m0);
        chain_ctx = 0;
        return CRGPU_OK;
    }

    template <class T> int upload(DevBuf& b, const std::vector<T>& v) {
        CR_TRY(b.reserve(v.size() * sizeof(T) + 16));
        CR_CUDA(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, stream));
        return CRGPU_OK;
    }
    template <class T> int download(std::vector<T>& v, const void* src, size_t n) {
        v.resize(n);
        CR_CUDA(cudaMemcpyAsync(v.data(), src, n * sizeof(T), cudaMemcpyDeviceToHost, stream));
        CR_CUDA(cudaStreamSynchronize(stream));
        return CRGPU_OK;
    }

    // Encode blk[0..nb) (consecutive blocks of this chain, dictionary-coded bytes at dD + off) into `out` at
    // out_base.  prefix_mode 0: [u32 payload_len] (dictionary payload, src/main.c:169);
    // 1: [u32 payload_len, u8 filt, u8 prec] (data block header, src/main.c:199-204); 2: none.
    // chain_ends: no further block of this chain follows (an aborted last block is then harmless).
    int encode_window(const uint8_t* dD, std::vector<BlockIO>& blk, int prefix_mode, bool chain_ends, DevBuf& out, size_t out_base, size_t& out_total);

    // ---- exact handling of a mid-chain "cannot compress" (SURVEY.md F11).  The reference's coder loop gives up on a block as soon as
    // the main stream is as large as the block (src/rolzmain/cr-coder.c:231-233, src/ropmain/cr-coder.c:204, src/roxmain/cr-coder.c:273)
    // and stores it raw, but the models, the side models and the PPM context keep what the tokens BEFORE that point did to them, and the
    // following blocks are coded on top of that.  encode_window() reports such a block (CR_RETRY_CUT + retry_block/retry_cut);
    // encode_blocks() restores the model snapshot taken before the window, re-runs the window up to and including that block with the
    // block's token walk cut at that position (BlockIO::cut), and carries on behind it.
    enum { CR_RETRY_CUT = 1 };
    DevBuf snap_o3b, snap_o3c, snap_o2, snap_o1, snap_m0;
    uint32_t snap_ctx = 0, retry_block = 0, retry_cut = 0;
    uint32_t cut_blocks = 0;         // statistics: blocks re-run with a cut since the handle was created
    int copy_state(bool save) {
        DevBuf* live[] = { &s_o3b, &s_o3c, &s_o2, &s_o1, &s_m0 }; DevBuf* snap[] = { &snap_o3b, &snap_o3c, &snap_o2, &snap_o1, &snap_m0 };
        const size_t bytes[] = { (size_t)PPM_O3_SLOTS, (size_t)PPM_O3_SLOTS, (size_t)65536 * PPM_O2_STRIDE, 65536, (size_t)X_NMODEL * 256 * 2 };
        for (int k = 0; k < 5; k++) {
            CR_TRY(snap[k]->reserve(bytes[k]));
            CR_CUDA(cudaMemcpyAsync(save ? snap[k]->p : live[k]->p, save ? live[k]->p : snap[k]->p, bytes[k], cudaMemcpyDeviceToDevice, stream));
        }
        if (save) snap_ctx = chain_ctx; else chain_ctx = snap_ctx;
        return CRGPU_OK;
    }
    int encode_blocks(const uint8_t* dD, std::vector<BlockIO>& blk, int prefix_mode, bool chain_ends, DevBuf& out, size_t out_base, size_t& out_total);
};

// Abort path only: the token whose last event is `event` ends at position t + len of block `block`.  Events per token as in the
// Count functors: ROLZ and LZ77 one, LZP two for a match and one (two if the literal equals the escape byte) for a literal.
struct AbortPos {
    typedef struct { uint32_t ev; } State;
    const uint8_t* D; const LzBlock* blocks; const uint32_t* scan_ev; int variant; uint32_t block, event; uint32_t* out;
    CR_D State begin(uint32_t c, uint32_t) const { State s = { scan_ev[c] }; return s; }
    CR_D void visit(State& s, uint32_t b, uint32_t t, uint32_t len) const {
        uint32_t n = 1;
        if (variant == CR_LZP) { const LzBlock B = blocks[b]; n = len > 1 ? 2 : 1 + (D[B.off + t] == B.esc); }
        if (b == block && event >= s.ev && event < s.ev + n) *out = t + len;
        s.ev += n;
    }
    CR_D void end(State&, uint32_t, uint32_t) const {}
};
// the event a dense main-stream triple belongs to: the largest e in [ev_begin, ev_end) with e + escord[e] <= dense_idx
__global__ void k_abort_event(const uint32_t* __restrict__ escord, uint32_t ev_begin, uint32_t ev_end, uint32_t dense_idx, uint32_t* __restrict__ out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    uint32_t lo = ev_begin, hi = ev_end;
    while (hi - lo > 1) { const uint32_t mid = lo + (hi - lo) / 2; if (mid + escord[mid] <= dense_idx) lo = mid; else hi = mid; }
    out[0] = lo; out[1] = 0;
}

__global__ void k_first_bytes(const uint8_t* __restrict__ D, const LzBlock* __restrict__ blocks, uint32_t nb, uint8_t* __restrict__ first) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    for (uint32_t i = 0; i < 16; i++) first[b * 16 + i] = i < blocks[b].size ? D[blocks[b].off + i] : 0;
}

// escape byte + PPM context carried into each block; serial over the blocks of the window (nb steps)
__global__ void k_rolz_finish_blocks(const uint8_t* __restrict__ D, LzBlock* __restrict__ blocks, uint32_t nb, const uint8_t* __restrict__ esc1,
                                     uint32_t ctx_in, uint32_t* __restrict__ ctx_out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    uint32_t c = ctx_in;
    for (uint32_t b = 0; b < nb; b++) {
        blocks[b].esc = esc1[b];
        blocks[b].cin = c;
        c = rz_ctx_at(D + blocks[b].off, blocks[b].walk, c);
    }
    *ctx_out = c;
}

static inline int cr_bits_for(uint32_t n) { int b = 0; while ((1u << b) < n) b++; return b; }

inline int LzChain::encode_window(const uint8_t* dD, std::vector<BlockIO>& blk, int prefix_mode, bool chain_ends, DevBuf& out, size_t out_base, size_t& out_total) {
    const uint32_t nb = (uint32_t)blk.size();
    if (nb == 0) { out_total = 0; return CRGPU_OK; }
    if (variant == CR_ROLZ && nb > RZ_MAX_BLOCKS) return CRGPU_ERR_ARG;            // callers split (sort key layout, cr_rolz.cuh)
    const uint32_t hdr_size = variant == CR_ROLZ ? 16 : variant == CR_LZP ? 20 : 32;
    const uint32_t prefix = prefix_mode == 0 ? 4 : prefix_mode == 1 ? 6 : 0;

    timer.begin(stream);
    // ---- block table
    std::vector<LzBlock> hb(nb);
    std::vector<uint64_t> segoff(nb);
    std::vector<uint32_t> seglen(nb);
    uint32_t nent = 0, maxsize = 0;
    size_t dtotal = 0;
    for (uint32_t b = 0; b < nb; b++) {
        memset(&hb[b], 0, sizeof(LzBlock));
        hb[b].off = blk[b].off; hb[b].size = blk[b].size; hb[b].eoff = nent;
        hb[b].walk = blk[b].cut ? blk[b].cut : blk[b].size;
        hb[b].ctx4 = blk[b].size >= 4194304;
        uint32_t first_entry = variant == CR_ROLZ ? 16 : LZP_FIRST;
        if (variant == CR_LZP && blk[b].size < 16) first_entry = blk[b].size;      // stored without coding (src/ropmain/cr-coder.c:140)
        if (variant == CR_LZ77) nent += x_entries(blk[b].size);
        else nent += blk[b].size > first_entry ? blk[b].size - first_entry : 0;
        segoff[b] = blk[b].off; seglen[b] = blk[b].size;
        if (blk[b].size > maxsize) maxsize = blk[b].size;
        if (blk[b].off + blk[b].size > dtotal) dtotal = blk[b].off + blk[b].size;
    }
    last_nent = nent; last_dtotal = dtotal;
    CR_TRY(upload(b_blocks, hb)); CR_TRY(upload(b_segoff, segoff)); CR_TRY(upload(b_seglen, seglen));
    LzBlock* d_blocks = b_blocks.as<LzBlock>();

    // ---- escape byte per block (rarest byte), first bytes, carried contexts
    CR_TRY(b_hist.reserve((size_t)nb * 256 * 4)); CR_TRY(b_esc1.reserve(nb)); CR_TRY(b_first.reserve((size_t)nb * 16)); CR_TRY(b_ctxout.reserve(16));
    CR_CUDA(cudaMemsetAsync(b_hist.p, 0, (size_t)nb * 256 * 4, stream));
    if (maxsize) CR_LAUNCH(k_hist256, dim3(cr_div_up(maxsize, CR_HIST_TILE), nb), dim3(256), stream, dD, b_segoff.as<uint64_t>(), b_seglen.as<uint32_t>(), b_hist.as<uint32_t>());
    CR_LAUNCH(k_pick_escapes, dim3(cr_div_up(nb, 64)), dim3(64), stream, b_hist.as<uint32_t>(), nb, (uint8_t*)nullptr, b_esc1.as<uint8_t>());
    CR_LAUNCH(k_first_bytes, dim3(cr_div_up(nb, 64)), dim3(64), stream, dD, d_blocks, nb, b_first.as<uint8_t>());

    timer.mark("hist_esc");
    // ---- per-position tokens
    CR_TRY(b_span.reserve(dtotal + 16)); CR_TRY(b_tidx.reserve(dtotal + 16));
    const int bbits = cr_bits_for(nb);
    if (variant == CR_ROLZ) {
        CR_LAUNCH(k_rolz_finish_blocks, dim3(1), dim3(1), stream, dD, d_blocks, nb, b_esc1.as<uint8_t>(), chain_ctx, b_ctxout.as<uint32_t>());
        CR_TRY(b_k0.reserve((size_t)nent * 4 + 16)); CR_TRY(b_k1.reserve((size_t)nent * 4 + 16)); CR_TRY(b_ks0.reserve((size_t)nent * 4 + 16));
        CR_TRY(b_v0.reserve((size_t)nent * 4 + 16)); CR_TRY(b_v1.reserve((size_t)nent * 4 + 16));
        CR_TRY(b_M.reserve((size_t)nent * 2 * 5 + 16)); CR_TRY(b_S.reserve((size_t)nent * 2 + 16));
        if (nent) {
            CR_LAUNCH(k_rolz_keys, dim3(cr_div_up(maxsize, 256), nb), dim3(256), stream, dD, d_blocks, b_k0.as<uint32_t>(), b_ks0.as<uint32_t>(), b_v0.as<uint32_t>());
            CR_TRY(cr_sort_pairs<uint32_t>(prims, b_k0.as<uint32_t>(), b_k1.as<uint32_t>(), b_v0.as<uint32_t>(), b_v1.as<uint32_t>(), nent, 0, RZ_BUCKET_BITS + bbits));
            if (flexible) { CR_TRY(b_rank.reserve((size_t)nent * 4 + 16)); CR_TRY(b_flexlen.reserve(nent + 16)); }
#ifndef CRGPU_SIM
            if (rolz_match_variant == 2)
                CR_LAUNCH(k_rolz_match_main2, dim3(cr_div_up(nent, RZM_TH)), dim3(RZM_TH), stream, dD, d_blocks, b_k1.as<uint32_t>(), b_v1.as<uint32_t>(), nent, b_M.as<uint16_t>(),
                          flexible ? b_rank.as<uint32_t>() : (uint32_t*)nullptr);
            else
#endif
            CR_LAUNCH(k_rolz_match_main, dim3(cr_div_up(nent, 128)), dim3(128), stream, dD, d_blocks, b_k1.as<uint32_t>(), b_v1.as<uint32_t>(), nent, b_M.as<uint16_t>(),
                      flexible ? b_rank.as<uint32_t>() : (uint32_t*)nullptr);
            if (flexible) CR_LAUNCH(k_rolz_flex, dim3(cr_div_up(maxsize, 128), nb), dim3(128), stream, dD, d_blocks, b_k1.as<uint32_t>(), b_v1.as<uint32_t>(), b_rank.as<uint32_t>(), b_M.as<uint16_t>(), b_flexlen.as<uint8_t>());
            CR_TRY(cr_sort_pairs<uint32_t>(prims, b_ks0.as<uint32_t>(), b_k1.as<uint32_t>(), b_v0.as<uint32_t>(), b_v1.as<uint32_t>(), nent, 0, 8 + bbits));
            CR_LAUNCH(k_rolz_match_short, dim3(cr_div_up(nent, 128)), dim3(128), stream, dD, d_blocks, b_k1.as<uint32_t>(), b_v1.as<uint32_t>(), nent, b_M.as<uint16_t>(), b_S.as<uint16_t>());
        }
        if (maxsize) CR_LAUNCH(k_rolz_tokens, dim3(cr_div_up(maxsize, 256), nb), dim3(256), stream, d_blocks, b_M.as<uint16_t>(), b_S.as<uint16_t>(), nent, b_span.as<uint8_t>(), b_tidx.as<uint8_t>(), flexible ? b_flexlen.as<uint8_t>() : (const uint8_t*)nullptr);
    } else if (variant == CR_LZ77) {
        CR_LAUNCH(k_x_finish_blocks, dim3(1), dim3(1), stream, dD, d_blocks, nb, b_esc1.as<uint8_t>(), chain_ctx, b_ctxout.as<uint32_t>());
        CR_TRY(x_npos.reserve(dtotal * 4 + 64)); CR_TRY(x_nlen.reserve(dtotal + 64)); CR_TRY(x_sdist.reserve(dtotal + 64)); CR_TRY(x_slen.reserve(dtotal + 64));
        CR_TRY(x_h16.reserve(dtotal * 2 + 64)); CR_TRY(x_G.reserve(dtotal * 4 + 64)); CR_TRY(x_tdist.reserve(dtotal * 4 + 64));
        if (maxsize) {
            const dim3 gp(cr_div_up(maxsize, 256), nb), tp(256);
            uint32_t maxkey = 0;
            for (uint32_t b = 0; b < nb; b++) { const uint32_t k = 20u * x_bucket2(blk[b].size); if (k > maxkey) maxkey = k; }
            const int kbits = cr_bits_for(maxkey);
            if (kbits + bbits > 32) return CRGPU_ERR_UNSUPPORTED;
            CR_TRY(b_k0.reserve((size_t)nent * 4 + 16)); CR_TRY(b_k1.reserve((size_t)nent * 4 + 16));
            CR_TRY(b_v0.reserve((size_t)nent * 4 + 16)); CR_TRY(b_v1.reserve((size_t)nent * 4 + 16)); CR_TRY(x_rank.reserve((size_t)nent * 4 + 16));
            if (nent) {
                CR_LAUNCH(k_x_keys, gp, tp, stream, dD, d_blocks, kbits, b_k0.as<uint32_t>(), b_v0.as<uint32_t>());
                CR_TRY(cr_sort_pairs<uint32_t>(prims, b_k0.as<uint32_t>(), b_k1.as<uint32_t>(), b_v0.as<uint32_t>(), b_v1.as<uint32_t>(), nent, 0, kbits + bbits));
                CR_LAUNCH(k_x_ranks, dim3(cr_div_up(nent, 256)), dim3(256), stream, b_k1.as<uint32_t>(), b_v1.as<uint32_t>(), nent, d_blocks, kbits, x_rank.as<uint32_t>());
            }
            if (flexible) {
                CR_TRY(x_mpos0.reserve(dtotal * 4 + 64)); CR_TRY(x_mlen0.reserve(dtotal + 64));
                CR_LAUNCH(k_x_match0, gp, tp, stream, dD, d_blocks, b_k1.as<uint32_t>(), b_v1.as<uint32_t>(), x_rank.as<uint32_t>(), match_limit, x_mpos0.as<uint32_t>(), x_mlen0.as<uint8_t>());
                CR_LAUNCH(k_x_flex, gp, tp, stream, d_blocks, x_mpos0.as<uint32_t>(), x_mlen0.as<uint8_t>(), x_npos.as<uint32_t>(), x_nlen.as<uint8_t>());
            } else {
                CR_LAUNCH(k_x_match, gp, tp, stream, dD, d_blocks, b_k1.as<uint32_t>(), b_v1.as<uint32_t>(), x_rank.as<uint32_t>(), match_limit, x_npos.as<uint32_t>(), x_nlen.as<uint8_t>());
            }
            CR_LAUNCH(k_x_short_hash, gp, tp, stream, dD, d_blocks, x_h16.as<uint16_t>());
            CR_LAUNCH(k_x_short, gp, tp, stream, dD, d_blocks, x_h16.as<uint16_t>(), x_sdist.as<uint8_t>(), x_slen.as<uint8_t>());
            CR_CUDA(cudaMemsetAsync(x_G.p, 0, dtotal * 4, stream));
        }
    } else {
        CR_LAUNCH(k_lzp_finish_blocks, dim3(cr_div_up(nb, 64)), dim3(64), stream, d_blocks, nb, b_esc1.as<uint8_t>());
        CR_TRY(b_k0.reserve((size_t)nent * 4 + 16)); CR_TRY(b_k1.reserve((size_t)nent * 4 + 16));
        CR_TRY(b_v0.reserve((size_t)nent * 4 + 16)); CR_TRY(b_v1.reserve((size_t)nent * 4 + 16));
        CR_TRY(b_M.reserve((size_t)nent * 12 + 16));
        uint32_t* cand = b_M.as<uint32_t>();
        for (int kind = 0; kind < 3 && nent; kind++) {
            CR_LAUNCH(k_lzp_keys, dim3(cr_div_up(maxsize, 256), nb), dim3(256), stream, dD, d_blocks, kind, b_k0.as<uint32_t>(), b_v0.as<uint32_t>());
            CR_TRY(cr_sort_pairs<uint32_t>(prims, b_k0.as<uint32_t>(), b_k1.as<uint32_t>(), b_v0.as<uint32_t>(), b_v1.as<uint32_t>(), nent, 0, lzp_hash_bits(kind) + bbits));
            CR_LAUNCH(k_lzp_prev, dim3(cr_div_up(nent, 256)), dim3(256), stream, d_blocks, kind, b_k1.as<uint32_t>(), b_v1.as<uint32_t>(), nent, cand + (size_t)kind * nent);
        }
        if (maxsize) CR_LAUNCH(k_lzp_tokens, dim3(cr_div_up(maxsize, 256), nb), dim3(256), stream, dD, d_blocks, cand, cand + nent, cand + 2 * (size_t)nent, b_span.as<uint8_t>());
    }

    timer.mark("match");
    // ---- resolve the parse (cr_chain.cuh)
    std::vector<ChainSeg> segs(nb);
    for (uint32_t b = 0; b < nb; b++) {
        segs[b].off = blk[b].off; segs[b].len = hb[b].walk;
        segs[b].start = variant == CR_ROLZ ? 1 : variant == CR_LZ77 ? 0 : (blk[b].size < 16 ? blk[b].size : LZP_FIRST);
    }
    for (uint32_t b = 0; b < nb; b++) if (segs[b].start > 255) { segs[b].len = 0; segs[b].start = 0; }   // tiny stored LZP block: nothing to walk
    const uint32_t nchunk = cr_chain_layout(segs.data(), nb);
    CR_TRY(upload(b_segs, segs));
    CR_TRY(b_xt.reserve((size_t)nchunk * 256 + 16)); CR_TRY(b_entry.reserve(nchunk + 16));
    const uint32_t ncount = variant == CR_LZ77 ? X_NCOUNT : 3;
    CR_TRY(b_cnt.reserve((size_t)(nchunk + 1) * 4 * ncount + 16)); CR_TRY(b_scan.reserve((size_t)(nchunk + 1) * 4 * ncount + 16));
    const ChainSeg* d_segs = b_segs.as<ChainSeg>();
    uint32_t* cnt = b_cnt.as<uint32_t>(); uint32_t* scan = b_scan.as<uint32_t>();
    CR_CUDA(cudaMemsetAsync(cnt, 0, (size_t)(nchunk + 1) * 4 * ncount, stream));
    if (nchunk && variant == CR_LZ77) {
        // fixed point over m_last_match (cr_lz77.cuh header)
        const dim3 gp(cr_div_up(maxsize, 256), nb), tp(256), gc(cr_div_up(nchunk, 64)), tc(64);
        const XTables XT = { x_npos.as<uint32_t>(), x_nlen.as<uint8_t>(), x_sdist.as<uint8_t>(), x_slen.as<uint8_t>() };
        CR_TRY(x_clast.reserve((size_t)nchunk * 4 + 16)); CR_TRY(x_ckey.reserve((size_t)nchunk * 12 + 16)); CR_TRY(x_cmax.reserve((size_t)nchunk * 4 + 16));
        CR_TRY(x_lastin.reserve((size_t)nchunk * 4 + 16)); CR_TRY(x_mism.reserve(16));
        auto propagate = [&]() -> int {
            CR_LAUNCH(k_chain_exits, gc, tc, stream, b_span.as<uint8_t>(), d_segs, nb, nchunk, b_xt.as<uint8_t>());
            CR_TRY(cr_chain_run_entries(*this, b_chainwork, segs, d_segs, nchunk, b_xt.as<uint8_t>(), b_entry.as<uint8_t>()));
            XLastOfChunk fa = { d_blocks, x_tdist.as<uint32_t>(), x_clast.as<uint32_t>() };
            CR_LAUNCH(k_chain_walk<XLastOfChunk>, gc, tc, stream, b_span.as<uint8_t>(), d_segs, nb, nchunk, b_entry.as<uint8_t>(), fa);
            CR_LAUNCH(k_x_chunk_keys, dim3(cr_div_up(nchunk, 256)), dim3(256), stream, x_clast.as<uint32_t>(), nchunk, x_ckey.as<uint32_t>());
            CR_TRY(cr_inclusive_max(prims, x_ckey.as<uint32_t>(), x_cmax.as<uint32_t>(), nchunk));
            CR_LAUNCH(k_x_chunk_in, dim3(cr_div_up(nchunk, 256)), dim3(256), stream, x_clast.as<uint32_t>(), x_cmax.as<uint32_t>(), d_segs, nb, nchunk, x_lastin.as<uint32_t>());
            return CRGPU_OK;
        };
        bool converged = false;
        uint32_t it = 0;
        for (; it < x_max_iter; it++) {
            CR_LAUNCH(k_x_decide, gp, tp, stream, dD, d_blocks, XT, x_G.as<uint32_t>(), it == 0 ? 1 : 0, b_span.as<uint8_t>(), x_tdist.as<uint32_t>());
            CR_TRY(propagate());
            CR_CUDA(cudaMemsetAsync(x_mism.p, 0, 4, stream));
            XCheck fb = { d_blocks, x_tdist.as<uint32_t>(), x_npos.as<uint32_t>(), x_lastin.as<uint32_t>(), x_G.as<uint32_t>(), x_mism.as<uint32_t>() };
            CR_LAUNCH(k_chain_walk<XCheck>, gc, tc, stream, b_span.as<uint8_t>(), d_segs, nb, nchunk, b_entry.as<uint8_t>(), fb);
            std::vector<uint32_t> hm;
            CR_TRY(download(hm, x_mism.p, 1));
            if (hm[0] == 0) { converged = true; it++; break; }
        }
        last_x_iters = it;
        timer.count("#x_iters", it);
        if (!converged) {
            CR_LAUNCH(k_x_parse_serial, dim3(cr_div_up(nb, 32)), dim3(32), stream, dD, d_blocks, nb, XT, b_span.as<uint8_t>(), x_tdist.as<uint32_t>());
            CR_TRY(propagate());
        }
        // the coder's own last_match on entry to every chunk (reuses the guess / key buffers of the iteration)
        CR_TRY(x_ckey.reserve((size_t)nchunk * 12 + 16));
        XCoderSum fs = { d_blocks, x_tdist.as<uint32_t>(), x_ckey.as<uint32_t>() };
        CR_LAUNCH(k_chain_walk<XCoderSum>, gc, tc, stream, b_span.as<uint8_t>(), d_segs, nb, nchunk, b_entry.as<uint8_t>(), fs);
        CR_LAUNCH(k_x_coder_in, dim3(cr_div_up(nb, 32)), dim3(32), stream, x_ckey.as<uint32_t>(), d_segs, nb, x_lastin.as<uint32_t>());
        XCount fc = { dD, d_blocks, x_tdist.as<uint32_t>(), x_lastin.as<uint32_t>(), cnt, nchunk + 1 };
        CR_LAUNCH(k_chain_walk<XCount>, gc, tc, stream, b_span.as<uint8_t>(), d_segs, nb, nchunk, b_entry.as<uint8_t>(), fc);
    } else if (nchunk) {
        CR_LAUNCH(k_chain_exits, dim3(cr_div_up(nchunk, 64)), dim3(64), stream, b_span.as<uint8_t>(), d_segs, nb, nchunk, b_xt.as<uint8_t>());
        CR_TRY(cr_chain_run_entries(*this, b_chainwork, segs, d_segs, nchunk, b_xt.as<uint8_t>(), b_entry.as<uint8_t>()));
        if (variant == CR_ROLZ) {
            RolzCount f = { dD, d_blocks, b_tidx.as<uint8_t>(), cnt, cnt + (nchunk + 1), cnt + 2 * (nchunk + 1) };
            CR_LAUNCH(k_chain_walk<RolzCount>, dim3(cr_div_up(nchunk, 64)), dim3(64), stream, b_span.as<uint8_t>(), d_segs, nb, nchunk, b_entry.as<uint8_t>(), f);
        } else {
            LzpCount f = { dD, d_blocks, cnt, cnt + (nchunk + 1), cnt + 2 * (nchunk + 1) };
            CR_LAUNCH(k_chain_walk<LzpCount>, dim3(cr_div_up(nchunk, 64)), dim3(64), stream, b_span.as<uint8_t>(), d_segs, nb, nchunk, b_entry.as<uint8_t>(), f);
        }
    }
    timer.mark("chain_count");
    for (uint32_t k = 0; k < ncount; k++) CR_TRY(cr_exclusive_sum(prims, cnt + (size_t)k * (nchunk + 1), scan + (size_t)k * (nchunk + 1), nchunk + 1));
    std::vector<uint32_t> hscan; std::vector<uint8_t> hfirst, hesc;
    CR_TRY(download(hscan, scan, (size_t)(nchunk + 1) * ncount));
    CR_TRY(download(hfirst, b_first.p, (size_t)nb * 16));
    CR_TRY(download(hesc, b_esc1.p, nb));
    const uint32_t* sc_ev = hscan.data(); const uint32_t* sc_a = sc_ev + (nchunk + 1); const uint32_t* sc_b = sc_a + (nchunk + 1);
    // ROLZ: a = match tokens, b = escape literals -> side symbols 2a+b.  LZP: a = tokens closing... unused.
    const uint32_t nev = sc_ev[nchunk];
    // LZ77: counters 1 = len symbols, 2 = spos symbols, 3 = pos tokens, 4 = pos symbols, 5.. = symbols per pos model
    auto xsc = [&](uint32_t k, uint32_t c) { return hscan[(size_t)k * (nchunk + 1) + c]; };
    const uint32_t x_nlen_sym = variant == CR_LZ77 ? xsc(1, nchunk) : 0, x_nspos = variant == CR_LZ77 ? xsc(2, nchunk) : 0, x_npossym = variant == CR_LZ77 ? xsc(4, nchunk) : 0;
    const uint32_t nside = variant == CR_ROLZ ? 2 * sc_a[nchunk] + sc_b[nchunk] : variant == CR_LZ77 ? x_nlen_sym + x_nspos + x_npossym : 0;
    last_nev = nev; last_nside = nside;

    // ---- events
    CR_TRY(b_evctx.reserve((size_t)nev * 4 + 16)); CR_TRY(b_evsym.reserve(nev + 16)); CR_TRY(b_tokend.reserve(nev + 16)); CR_TRY(b_pred.reserve(nev + 16));
    CR_TRY(b_T1.reserve((size_t)nev * 8 + 16)); CR_TRY(b_side.reserve((size_t)nside * 2 + 16)); CR_TRY(b_TS.reserve((size_t)nside * 8 + 16));
    const uint32_t n_idxsym = variant == CR_ROLZ ? sc_a[nchunk] : 0, n_lensym = variant == CR_ROLZ ? sc_a[nchunk] + sc_b[nchunk] : 0;
    SideJobs xjobs; memset(&xjobs, 0, sizeof xjobs);
    CR_TRY(b_lensym.reserve(n_lensym + 16)); CR_TRY(b_lenpos.reserve((size_t)n_lensym * 4 + 16));
    CR_TRY(b_idxsym.reserve(n_idxsym + 16)); CR_TRY(b_idxpos.reserve((size_t)n_idxsym * 4 + 16));
    if (nchunk) {
        if (variant == CR_ROLZ) {
            RolzEmit f = { dD, d_blocks, b_tidx.as<uint8_t>(), scan, scan + (nchunk + 1), scan + 2 * (nchunk + 1), b_evctx.as<uint32_t>(), b_evsym.as<uint8_t>(), b_side.as<uint16_t>(),
                           b_lensym.as<uint8_t>(), b_lenpos.as<uint32_t>(), b_idxsym.as<uint8_t>(), b_idxpos.as<uint32_t>() };
            CR_LAUNCH(k_chain_walk<RolzEmit>, dim3(cr_div_up(nchunk, 64)), dim3(64), stream, b_span.as<uint8_t>(), d_segs, nb, nchunk, b_entry.as<uint8_t>(), f);
        } else if (variant == CR_LZ77) {
            // per-model symbol lists: [len | spos | pos0..pos5], each (symbol u8, index in the side array u32)
            uint32_t mcount[X_NMODEL] = { x_nlen_sym, x_nspos, 0, 0, 0, 0, 0, 0 };
            for (int k = 0; k < 6; k++) mcount[2 + k] = xsc(5 + k, nchunk);
            size_t mtotal = 0; size_t moff[X_NMODEL];
            for (int k = 0; k < X_NMODEL; k++) { moff[k] = mtotal; mtotal += mcount[k]; }
            CR_TRY(x_msym.reserve(mtotal + 64)); CR_TRY(x_mpos.reserve(mtotal * 4 + 64));
            XEmit f;
            f.D = dD; f.blocks = d_blocks; f.tdist = x_tdist.as<uint32_t>(); f.last_in = x_lastin.as<uint32_t>(); f.scan = scan; f.stride = nchunk + 1;
            f.base_pos = x_nspos; f.base_len = x_nspos + x_npossym;
            f.ev_ctx = b_evctx.as<uint32_t>(); f.ev_sym = b_evsym.as<uint8_t>();
            for (int k = 0; k < X_NMODEL; k++) {
                f.msym[k] = x_msym.as<uint8_t>() + moff[k]; f.mpos[k] = x_mpos.as<uint32_t>() + moff[k];
                xjobs.sym[k] = f.msym[k]; xjobs.pos[k] = f.mpos[k]; xjobs.n[k] = mcount[k]; xjobs.inc[k] = x_inc(k); xjobs.state[k] = st.m0 + k * 256;
            }
            xjobs.nmodel = X_NMODEL;
            CR_LAUNCH(k_chain_walk<XEmit>, dim3(cr_div_up(nchunk, 64)), dim3(64), stream, b_span.as<uint8_t>(), d_segs, nb, nchunk, b_entry.as<uint8_t>(), f);
        } else {
            CR_TRY(b_ord.reserve((size_t)(nchunk + 1) * 4 + 16));
            CR_LAUNCH(k_lzp_ctx_in, dim3(cr_div_up(nchunk + 1, 128)), dim3(128), stream, cnt + (nchunk + 1), cnt + 2 * (nchunk + 1), nchunk, chain_ctx, b_ord.as<uint32_t>());
            CR_CUDA(cudaMemcpyAsync(b_ctxout.p, b_ord.as<uint32_t>() + nchunk, 4, cudaMemcpyDeviceToDevice, stream));
            LzpEmit f = { dD, d_blocks, scan, b_ord.as<uint32_t>(), b_evctx.as<uint32_t>(), b_evsym.as<uint8_t>(), b_tokend.as<uint8_t>() };
            CR_LAUNCH(k_chain_walk<LzpEmit>, dim3(cr_div_up(nchunk, 64)), dim3(64), stream, b_span.as<uint8_t>(), d_segs, nb, nchunk, b_entry.as<uint8_t>(), f);
        }
    } else if (variant == CR_LZP) {
        std::vector<uint32_t> keep(1, chain_ctx);
        CR_TRY(upload(b_ctxout, keep));
    }

    timer.mark("events");
    // ---- order-0 side models (len / idx) are independent of the PPM passes: run them on their own stream
    bool side_async = false;
#ifndef CRGPU_SIM
    if (nside && !scalar_models) {
        CR_TRY(b_denseside.reserve((size_t)nside * sizeof(Tri) + 16));
        CR_CUDA(cudaEventRecord(ev_side_go, stream));
        CR_CUDA(cudaStreamWaitEvent(side_stream, ev_side_go, 0));
        if (variant == CR_LZ77) CR_LAUNCH(k_side_epochs_jobs, dim3(X_NMODEL), dim3(SE_THREADS), side_stream, xjobs, b_TS.as<uint64_t>());
        else CR_LAUNCH(k_side_epochs, dim3(2), dim3(SE_THREADS), side_stream, b_lensym.as<uint8_t>(), b_lenpos.as<uint32_t>(), n_lensym, b_idxsym.as<uint8_t>(), b_idxpos.as<uint32_t>(), n_idxsym, st, b_TS.as<uint64_t>());
        CR_LAUNCH(k_expand_side, dim3(cr_div_up(nside, 256)), dim3(256), side_stream, b_TS.as<uint64_t>(), nside, b_denseside.as<Tri>());
        CR_CUDA(cudaEventRecord(ev_side_done, side_stream));
        side_async = true;
    }
#endif
    // ---- model passes
    uint32_t nesc = 0;
    CR_TRY(b_esccount.reserve(16));
    CR_CUDA(cudaMemsetAsync(b_esccount.p, 0, 4, stream));
    if (nev) {
        CR_TRY(b_k0.reserve((size_t)nev * 4 + 16)); CR_TRY(b_k1.reserve((size_t)nev * 4 + 16));
        CR_TRY(b_v0.reserve((size_t)nev * 4 + 16)); CR_TRY(b_v1.reserve((size_t)nev * 4 + 16));
        CR_TRY(b_escrec.reserve((size_t)nev * sizeof(EscRec) + 16));
        const dim3 ge(cr_div_up(nev, 256)), te(256);
        CR_LAUNCH(k_o3_keys, ge, te, stream, b_evctx.as<uint32_t>(), b_evsym.as<uint8_t>(), nev, b_k0.as<uint32_t>(), b_v0.as<uint32_t>());
        CR_TRY(cr_sort_pairs<uint32_t>(prims, b_k0.as<uint32_t>(), b_k1.as<uint32_t>(), b_v0.as<uint32_t>(), b_v1.as<uint32_t>(), nev, 0, 22));
#ifndef CRGPU_SIM
        if (!scalar_models && hot_contexts) {
            const uint32_t hot_cap = nev / O3_HANDOVER + 16;
            CR_TRY(b_o3hot.reserve((size_t)hot_cap * sizeof(O3Hot) + 16));
            CR_CUDA(cudaMemsetAsync(b_esccount.as<uint32_t>() + 1, 0, 8, stream));      // hot-segment count and the queue head of k_o3_hot_spec
            // list the slot segments, longest first
            CR_TRY(b_flag.reserve((size_t)(nev + 1) * 4 + 16)); CR_TRY(b_escord.reserve((size_t)(nev + 1) * 4 + 16));
            CR_LAUNCH(k_o3_headflags, dim3(cr_div_up(nev + 1, 256)), dim3(256), stream, b_k1.as<uint32_t>(), nev, b_flag.as<uint32_t>());
            CR_TRY(cr_exclusive_sum(prims, b_flag.as<uint32_t>(), b_escord.as<uint32_t>(), nev + 1));
            std::vector<uint32_t> hseg;
            CR_TRY(download(hseg, b_escord.as<uint32_t>() + nev, 1));
            const uint32_t nseg = hseg[0];
            CR_TRY(b_segstart.reserve((size_t)(nseg + 1) * 4 + 16)); CR_TRY(b_segkey.reserve((size_t)nseg * 16 + 64));
            uint32_t* sk0 = b_segkey.as<uint32_t>(); uint32_t* sk1 = sk0 + nseg; uint32_t* sv0 = sk1 + nseg; uint32_t* sv1 = sv0 + nseg;
            CR_LAUNCH(k_o3_segstarts, dim3(cr_div_up(nev + 1, 256)), dim3(256), stream, b_flag.as<uint32_t>(), b_escord.as<uint32_t>(), nev, b_segstart.as<uint32_t>());
            CR_LAUNCH(k_o3_seglen_keys, dim3(cr_div_up(nseg, 256)), dim3(256), stream, b_segstart.as<uint32_t>(), nseg, sk0, sv0);
            CR_TRY(cr_sort_pairs<uint32_t>(prims, sk0, sk1, sv0, sv1, nseg, 0, 32));
            CR_LAUNCH(k_o3_pass_sorted, dim3(cr_div_up(nseg, 128)), dim3(128), stream, b_k1.as<uint32_t>(), b_v1.as<uint32_t>(), nev, b_segstart.as<uint32_t>(), sv1, nseg, st,
                      b_pred.as<uint8_t>(), b_o3hot.as<O3Hot>(), b_esccount.as<uint32_t>() + 1, hot_cap);
            CR_LAUNCH(k_o3_hot_spec, dim3(296), dim3(O3S_THREADS), stream, b_k1.as<uint32_t>(), b_v1.as<uint32_t>(), nev, b_segstart.as<uint32_t>(), st, b_pred.as<uint8_t>(), b_o3hot.as<O3Hot>(), b_esccount.as<uint32_t>() + 1, hot_cap);
        } else
#endif
        CR_LAUNCH(k_o3_pass, dim3(cr_div_up(nev, 128)), dim3(128), stream, b_k1.as<uint32_t>(), b_v1.as<uint32_t>(), nev, st, b_pred.as<uint8_t>(), (O3Hot*)nullptr, (uint32_t*)nullptr, 0u);
    timer.mark("o3");
        CR_LAUNCH(k_o2_keys, ge, te, stream, b_evctx.as<uint32_t>(), b_evsym.as<uint8_t>(), b_pred.as<uint8_t>(), nev, b_k0.as<uint32_t>(), b_v0.as<uint32_t>());
        CR_TRY(cr_sort_pairs<uint32_t>(prims, b_k0.as<uint32_t>(), b_k1.as<uint32_t>(), b_v0.as<uint32_t>(), b_v1.as<uint32_t>(), nev, 0, 16));
#ifndef CRGPU_SIM
        if (!scalar_models) {
            const uint32_t hot_min = hot_contexts ? O2C_MIN : 0xFFFFFFFFu;
            CR_TRY(b_bounds.reserve(65537 * 4 + 16));
            CR_LAUNCH(k_o2_bounds, dim3(cr_div_up(65537, 256)), dim3(256), stream, b_k1.as<uint32_t>(), nev, b_bounds.as<uint32_t>());
            if (hot_contexts) {
                CR_TRY(b_hits.reserve(64));
                CR_CUDA(cudaMemsetAsync(b_hits.p, 0, 4, stream));
                CR_LAUNCH(k_o2_hits, dim3(cr_div_up(nev, 256)), dim3(256), stream, b_k1.as<uint32_t>(), nev, b_hits.as<uint32_t>());
                std::vector<uint32_t> hh;
                CR_TRY(download(hh, b_hits.p, 1));
                timer.count("#o3_hits", hh[0]);
                const bool narrow = (double)hh[0] > 0.25 * (double)nev;
                if (o2_hot_variant >= 2) {
                    if (!o2_attr_done) {
                        CR_CUDA(cudaFuncSetAttribute(k_o2_hot<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(O2HotSmem<256>)));
                        CR_CUDA(cudaFuncSetAttribute(k_o2_hot<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(O2HotSmem<1024>)));
                        CR_CUDA(cudaFuncSetAttribute(k_o2_hot<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(O2HotSmem<512>)));
                        CR_CUDA(cudaFuncSetAttribute(k_o2_eval<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(O2EvalSmem<512>)));
                        CR_CUDA(cudaFuncSetAttribute(k_o2_eval<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(O2EvalSmem<256>)));
                        CR_CUDA(cudaFuncSetAttribute(k_o2_eval<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(O2EvalSmem<1024>)));
                        o2_attr_done = true;
                    }
                    const uint32_t* nolist = nullptr; const O2SplitCtl* noctl = nullptr; const uint8_t* notodo = nullptr;
#define CR_O2_LAUNCH(KERNEL, GRID, TH, SMEM, ...) do { __atomic_fetch_add(&g_cr_launches, 1ull, __ATOMIC_RELAXED); \
                        KERNEL<TH><<<dim3(GRID), dim3(TH), (SMEM), stream>>>(__VA_ARGS__); CR_CUDA(cudaGetLastError()); } while (0)
                    if (o2_hot_variant == 2) {
                        if (narrow) CR_O2_LAUNCH(k_o2_hot, 65536, 256, sizeof(O2HotSmem<256>), b_k1.as<uint32_t>(), b_v1.as<uint32_t>(), nev, st, b_T1.as<uint64_t>(), b_escrec.as<EscRec>(), b_esccount.as<uint32_t>(), b_bounds.as<uint32_t>(), nolist, noctl, notodo);
                        else CR_O2_LAUNCH(k_o2_hot, 65536, 1024, sizeof(O2HotSmem<1024>), b_k1.as<uint32_t>(), b_v1.as<uint32_t>(), nev, st, b_T1.as<uint64_t>(), b_escrec.as<EscRec>(), b_esccount.as<uint32_t>(), b_bounds.as<uint32_t>(), nolist, noctl, notodo);
                    } else {
                        // split form: the chain of steps (k_o2_skel) records (position, length, flags, table) per step, k_o2_eval turns every record
                        // into triples and escape records.  Records: one per TH events plus one per rescale; a rescale needs 125 hits or 125
                        // occurrences of one symbol since the last one (short cascades aside), so n / 120 bounds them with room to spare; a context
                        // that still runs out is redone by k_o2_hot (nothing of it has been written by then)
                        const uint32_t th = o2_width ? o2_width : narrow ? 256u : 1024u;
                        const uint32_t hot_bound = nev / O2C_MIN + 1 < 65536u ? nev / O2C_MIN + 1 : 65536u;
                        const uint32_t cap = o2_rec_cap_test ? o2_rec_cap_test : nev / th + nev / 120u + 2u * hot_bound + 256u;
                        CR_TRY(b_o2hot.reserve(65536 * 4)); CR_TRY(b_o2todo.reserve(65536)); CR_TRY(b_o2steps.reserve((size_t)cap * sizeof(O2Step))); CR_TRY(b_o2snaps.reserve((size_t)cap * 256));
                        std::vector<uint32_t> ctl0 = { 0u, 0u, cap, 0u };
                        CR_TRY(upload(b_o2ctl, ctl0));
                        CR_LAUNCH(k_o2_hotlist, dim3(256), dim3(256), stream, b_bounds.as<uint32_t>(), b_o2hot.as<uint32_t>(), b_o2ctl.as<O2SplitCtl>(), b_o2todo.as<uint8_t>());
#define CR_O2_SPLIT(TH) do { \
                            CR_O2_LAUNCH(k_o2_skel, hot_bound, TH, 0, b_k1.as<uint32_t>(), st, b_bounds.as<uint32_t>(), b_o2hot.as<uint32_t>(), b_o2ctl.as<O2SplitCtl>(), b_o2todo.as<uint8_t>(), b_o2steps.as<O2Step>(), b_o2snaps.as<uint8_t>()); \
                            CR_O2_LAUNCH(k_o2_eval, 148u * (2048u / TH), TH, sizeof(O2EvalSmem<TH>), b_k1.as<uint32_t>(), b_v1.as<uint32_t>(), b_o2ctl.as<O2SplitCtl>(), b_o2todo.as<uint8_t>(), b_o2steps.as<O2Step>(), b_o2snaps.as<uint8_t>(), b_T1.as<uint64_t>(), b_escrec.as<EscRec>(), b_esccount.as<uint32_t>()); \
                            CR_O2_LAUNCH(k_o2_hot, hot_bound, TH, sizeof(O2HotSmem<TH>), b_k1.as<uint32_t>(), b_v1.as<uint32_t>(), nev, st, b_T1.as<uint64_t>(), b_escrec.as<EscRec>(), b_esccount.as<uint32_t>(), b_bounds.as<uint32_t>(), b_o2hot.as<uint32_t>(), b_o2ctl.as<O2SplitCtl>(), b_o2todo.as<uint8_t>()); } while (0)
                        if (th == 256) CR_O2_SPLIT(256); else if (th == 512) CR_O2_SPLIT(512); else CR_O2_SPLIT(1024);
#undef CR_O2_SPLIT
                    }
#undef CR_O2_LAUNCH
                } else if (narrow)
                    CR_LAUNCH(k_o2_pass_cta<256>, dim3(65536), dim3(256), stream, b_k1.as<uint32_t>(), b_v1.as<uint32_t>(), nev, st, b_T1.as<uint64_t>(), b_escrec.as<EscRec>(), b_esccount.as<uint32_t>(), b_bounds.as<uint32_t>());
                else
                    CR_LAUNCH(k_o2_pass_cta<1024>, dim3(65536), dim3(1024), stream, b_k1.as<uint32_t>(), b_v1.as<uint32_t>(), nev, st, b_T1.as<uint64_t>(), b_escrec.as<EscRec>(), b_esccount.as<uint32_t>(), b_bounds.as<uint32_t>());
            }
            CR_LAUNCH(k_o2_pass_warp, dim3(65536 * 32 / 128), dim3(128), stream, b_k1.as<uint32_t>(), b_v1.as<uint32_t>(), nev, st, b_T1.as<uint64_t>(), b_escrec.as<EscRec>(), b_esccount.as<uint32_t>(), hot_min, b_bounds.as<uint32_t>());
        }
        else
#endif
        CR_LAUNCH(k_o2_pass, dim3(cr_div_up(nev, 128)), dim3(128), stream, b_k1.as<uint32_t>(), b_v1.as<uint32_t>(), nev, st, b_T1.as<uint64_t>(), b_escrec.as<EscRec>(), b_esccount.as<uint32_t>());
    timer.mark("o2");
        std::vector<uint32_t> hc;
        CR_TRY(download(hc, b_esccount.p, 1));
        nesc = hc[0];
        if (nesc) {
            CR_TRY(b_k64a.reserve((size_t)nesc * 8 + 16)); CR_TRY(b_k64b.reserve((size_t)nesc * 8 + 16)); CR_TRY(b_ord.reserve((size_t)nesc * 4 + 16));
            CR_TRY(b_T2.reserve((size_t)nesc * 8 + 16));
            const dim3 gx(cr_div_up(nesc, 256));
            CR_LAUNCH(k_o1_keys, gx, te, stream, b_escrec.as<EscRec>(), nesc, b_k64a.as<uint64_t>(), b_v0.as<uint32_t>());
            CR_TRY(cr_sort_pairs<uint64_t>(prims, b_k64a.as<uint64_t>(), b_k64b.as<uint64_t>(), b_v0.as<uint32_t>(), b_v1.as<uint32_t>(), nesc, 0, 32));
            CR_LAUNCH(k_o1_ordinals, gx, te, stream, b_v1.as<uint32_t>(), nesc, b_ord.as<uint32_t>());
            CR_TRY(cr_sort_pairs<uint64_t>(prims, b_k64a.as<uint64_t>(), b_k64b.as<uint64_t>(), b_v0.as<uint32_t>(), b_v1.as<uint32_t>(), nesc, 0, 40));
#ifndef CRGPU_SIM
            if (!scalar_models) {
                CR_TRY(b_o1info.reserve((size_t)nesc * 4 + 16)); CR_TRY(b_o1ord.reserve((size_t)nesc * 4 + 16)); CR_TRY(b_o1incl.reserve((size_t)nesc * 32 + 32));
                CR_LAUNCH(k_o1_gather, gx, te, stream, b_v1.as<uint32_t>(), nesc, b_escrec.as<EscRec>(), b_ord.as<uint32_t>(), b_o1info.as<uint32_t>(), b_o1ord.as<uint32_t>(), b_o1incl.as<uint4>());
                if (hot_contexts && o1_hot_variant == 2) {
                    const uint32_t max_rec = nesc / O1S_TH + nesc / 127u + 2u * 256u;
                    CR_TRY(b_o1ctx.reserve(256 * sizeof(O1Ctx))); CR_TRY(b_o1steps.reserve((size_t)max_rec * sizeof(O1Step))); CR_TRY(b_o1snaps.reserve((size_t)max_rec * 256));
                    CR_LAUNCH(k_o1_plan, dim3(1), dim3(256), stream, b_k64b.as<uint64_t>(), nesc, b_o1ctx.as<O1Ctx>());
                    CR_LAUNCH(k_o1_skel, dim3(256), dim3(O1S_TH), stream, b_o1ctx.as<O1Ctx>(), b_o1info.as<uint32_t>(), st, b_o1steps.as<O1Step>(), b_o1snaps.as<uint8_t>());
                    CR_LAUNCH(k_o1_eval, dim3(max_rec), dim3(O1C_THREADS), stream, b_o1ctx.as<O1Ctx>(), b_o1steps.as<O1Step>(), b_o1snaps.as<uint8_t>(),
                              b_o1info.as<uint32_t>(), b_o1ord.as<uint32_t>(), b_o1incl.as<uint4>(), b_T2.as<uint64_t>());
                } else
                if (hot_contexts) CR_LAUNCH(k_o1_pass_cta, dim3(256), dim3(O1C_THREADS), stream, b_k64b.as<uint64_t>(), nesc, b_o1info.as<uint32_t>(), b_o1ord.as<uint32_t>(), b_o1incl.as<uint4>(), st, b_T2.as<uint64_t>());
                CR_LAUNCH(k_o1_pass_warp, dim3(256 * 32 / 128), dim3(128), stream, b_k64b.as<uint64_t>(), nesc, b_o1info.as<uint32_t>(), b_o1ord.as<uint32_t>(), b_o1incl.as<uint4>(), st, b_T2.as<uint64_t>(), hot_contexts ? O1C_MIN : 0xFFFFFFFFu);
            } else
#endif
            CR_LAUNCH(k_o1_pass, dim3(cr_div_up(nesc, 64)), dim3(64), stream, b_k64b.as<uint64_t>(), b_v1.as<uint32_t>(), nesc, b_escrec.as<EscRec>(), b_ord.as<uint32_t>(), st, b_T2.as<uint64_t>());
        }
    }
    timer.mark("o1");
    last_nesc = nesc;
    timer.count("#events", nev); timer.count("#triples", (double)nev + nesc + nside); timer.count("#escapes", nesc); timer.count("#side_symbols", nside);
    if (nside && !side_async && variant == CR_LZ77) {
        CR_LAUNCH(k_side_models_jobs, dim3(1), dim3(X_NMODEL), stream, xjobs, b_TS.as<uint64_t>());
    } else if (nside && !side_async) {
#ifndef CRGPU_SIM
        if (!scalar_models) CR_LAUNCH(k_side_epochs, dim3(2), dim3(SE_THREADS), stream, b_lensym.as<uint8_t>(), b_lenpos.as<uint32_t>(), n_lensym, b_idxsym.as<uint8_t>(), b_idxpos.as<uint32_t>(), n_idxsym, st, b_TS.as<uint64_t>());
        else
#endif
        CR_LAUNCH(k_side_models, dim3(1), dim3(1), stream, b_side.as<uint16_t>(), nside, st, b_TS.as<uint64_t>());
    }

    timer.mark("side_models");
    // ---- dense triple streams
    CR_TRY(b_flag.reserve((size_t)(nev + 1) * 4 + 16)); CR_TRY(b_escord.reserve((size_t)(nev + 1) * 4 + 16));
    CR_TRY(b_dense.reserve(((size_t)nev + nesc) * sizeof(Tri) + 16)); CR_TRY(b_denseside.reserve((size_t)nside * sizeof(Tri) + 16));
    CR_LAUNCH(k_esc_flags, dim3(cr_div_up(nev + 1, 256)), dim3(256), stream, b_T1.as<uint64_t>(), nev, b_flag.as<uint32_t>());
    CR_TRY(cr_exclusive_sum(prims, b_flag.as<uint32_t>(), b_escord.as<uint32_t>(), nev + 1));
    if (nev) CR_LAUNCH(k_expand_main, dim3(cr_div_up(nev, 256)), dim3(256), stream, b_T1.as<uint64_t>(), b_T2.as<uint64_t>(), b_escord.as<uint32_t>(),
                       variant == CR_LZP ? b_tokend.as<uint8_t>() : (const uint8_t*)nullptr, nev, b_dense.as<Tri>());
    if (nside && !side_async) CR_LAUNCH(k_expand_side, dim3(cr_div_up(nside, 256)), dim3(256), stream, b_TS.as<uint64_t>(), nside, b_denseside.as<Tri>());
#ifndef CRGPU_SIM
    if (side_async) CR_CUDA(cudaStreamWaitEvent(stream, ev_side_done, 0));
#endif

    timer.mark("expand");
    // ---- range coding: one serial coder per (block, stream)
    const uint32_t spb = variant == CR_ROLZ ? 2 : variant == CR_LZ77 ? 4 : 1;
    std::vector<uint32_t> x_hdr_counts((size_t)nb * 3, 0);          // LZ77: num_spos, num_pos (tokens), num_len per block
    std::vector<RcStream> streams((size_t)nb * spb);
    std::vector<uint32_t> num_idx(nb, 0);
    size_t rc_total = 0;
    for (uint32_t b = 0; b < nb; b++) {
        const uint32_t c0 = segs[b].chunk0, c1 = c0 + segs[b].nchunk;
        RcStream& m = streams[(size_t)b * spb];
        memset(&m, 0, sizeof m);
        m.ev_begin = sc_ev[c0]; m.ev_end = sc_ev[c1]; m.is_main = 1;
        m.limit = blk[b].size > hdr_size ? blk[b].size - hdr_size : 0;
        m.out_off = rc_total; m.out_cap = blk[b].size + 16; rc_total += m.out_cap;
        if (variant == CR_ROLZ) {
            RcStream& s = streams[(size_t)b * spb + 1];
            memset(&s, 0, sizeof s);
            s.ev_begin = 2 * sc_a[c0] + sc_b[c0]; s.ev_end = 2 * sc_a[c1] + sc_b[c1]; s.is_main = 0; s.limit = 0xFFFFFFFFu;
            s.out_off = rc_total; s.out_cap = 3 * (s.ev_end - s.ev_begin) + 16; rc_total += s.out_cap;
            num_idx[b] = (sc_a[c1] - sc_a[c0]) + (sc_b[c1] - sc_b[c0]);
        }
        if (variant == CR_LZ77) {                                    // side streams in payload order: spos, pos, len (cr-coder.c:283-294)
            const uint32_t lo[3] = { xsc(2, c0), x_nspos + xsc(4, c0), x_nspos + x_npossym + xsc(1, c0) };
            const uint32_t hi[3] = { xsc(2, c1), x_nspos + xsc(4, c1), x_nspos + x_npossym + xsc(1, c1) };
            for (int k = 0; k < 3; k++) {
                RcStream& s = streams[(size_t)b * spb + 1 + k];
                memset(&s, 0, sizeof s);
                s.ev_begin = lo[k]; s.ev_end = hi[k]; s.is_main = 0; s.limit = 0xFFFFFFFFu;
                s.out_off = rc_total; s.out_cap = 3 * (hi[k] - lo[k]) + 16; rc_total += s.out_cap;
            }
            x_hdr_counts[(size_t)b * 3] = xsc(2, c1) - xsc(2, c0); x_hdr_counts[(size_t)b * 3 + 1] = xsc(3, c1) - xsc(3, c0); x_hdr_counts[(size_t)b * 3 + 2] = xsc(1, c1) - xsc(1, c0);
        }
    }
    // (A block re-run with a cut is stored raw whatever its streams say, so its serial chain is wasted work.  Emptying its streams
    // here (ev_end = ev_begin) is exact on the scalar path and k_low_scatter skips triples outside every stream's range, but that
    // variant has not run on a GPU yet: left for the next measurement session, it only matters for runs of incompressible blocks.)
    CR_TRY(upload(b_streams, streams));
    CR_TRY(b_rcres.reserve(streams.size() * sizeof(RcResult) + 16)); CR_TRY(b_rcout.reserve(rc_total + 16));
    std::vector<RcResult> res;
    bool rc_done = false;
#ifndef CRGPU_SIM
    if (!scalar_models) {
        // split form (cr_warp.cuh): serial range chain, then the output bytes as a parallel big-number sum
        const size_t ntm = (size_t)nev + nesc, nts = nside;
        const uint32_t nstr = (uint32_t)streams.size();
        CR_TRY(b_qm.reserve((ntm + 1) * 4 + 16)); CR_TRY(b_shm.reserve((ntm + 1) * 4 + 16)); CR_TRY(b_bm.reserve((ntm + 1) * 4 + 16));
        CR_TRY(b_qs.reserve((nts + 1) * 4 + 16)); CR_TRY(b_shs.reserve((nts + 1) * 4 + 16)); CR_TRY(b_bs.reserve((nts + 1) * 4 + 16));
        CR_TRY(b_stot.reserve((size_t)nstr * sizeof(StreamTotals) + 16));
        CR_CUDA(cudaMemsetAsync(b_shm.as<uint32_t>() + ntm, 0, 4, stream));
        CR_CUDA(cudaMemsetAsync(b_shs.as<uint32_t>() + nts, 0, 4, stream));
        const dim3 gchain(cr_div_up((size_t)nstr * 32, 128));
        if (rc_variant == 1) {
            CR_LAUNCH(k_range_chain<1>, gchain, dim3(128), stream, (const uint4*)b_dense.p, (const uint4*)b_denseside.p, b_escord.as<uint32_t>(),
                      b_streams.as<RcStream>(), nstr, b_qm.as<uint32_t>(), b_shm.as<uint32_t>(), b_qs.as<uint32_t>(), b_shs.as<uint32_t>());
        } else {
            CR_TRY(b_cinm.reserve((ntm + 1) * 16 + 16)); CR_TRY(b_cins.reserve((nts + 1) * 16 + 16));
            if (rc_variant >= 7) {
                if (ntm) CR_LAUNCH(k_chain_inputs_dp, dim3(cr_div_up(ntm, 256)), dim3(256), stream, b_dense.as<Tri>(), (uint64_t)ntm, b_cinm.as<uint4>());
                if (nts) CR_LAUNCH(k_chain_inputs_dp, dim3(cr_div_up(nts, 256)), dim3(256), stream, b_denseside.as<Tri>(), (uint64_t)nts, b_cins.as<uint4>());
            } else {
            if (ntm) CR_LAUNCH(k_chain_inputs, dim3(cr_div_up(ntm, 256)), dim3(256), stream, b_dense.as<Tri>(), (uint64_t)ntm, b_cinm.as<uint4>());
            if (nts) CR_LAUNCH(k_chain_inputs, dim3(cr_div_up(nts, 256)), dim3(256), stream, b_denseside.as<Tri>(), (uint64_t)nts, b_cins.as<uint4>());
            }
            timer.mark("expand");
            if (rc_variant == 8) {
                CR_TRY(rcpar.run(stream, b_streams.as<RcStream>(), nstr, b_escord.as<uint32_t>(), ntm, nts, b_dense.as<Tri>(), b_denseside.as<Tri>(),
                                 b_cinm.as<uint4>(), b_cins.as<uint4>(), b_qm.as<uint32_t>(), b_shm.as<uint32_t>(), b_qs.as<uint32_t>(), b_shs.as<uint32_t>(), timer.enabled));
                if (timer.enabled) {
                    timer.count("#rcp_state_steps", (double)rcpar.last.state_steps); timer.count("#rcp_live_jobs", rcpar.last.live_jobs);
                    timer.count("#rcp_merged_jobs", rcpar.last.merged_jobs); timer.count("#rcp_seed_retries", rcpar.last.seed_retries);
                    timer.count("#rcp_demoted_jobs", rcpar.last.demoted_jobs); timer.count("#rcp_flagged_streams", rcpar.last.flagged_streams);
                    timer.count("#rcp_max_exit_set", rcpar.last.max_e);
                    timer.count("#rcp_steps_seed", (double)rcpar.last.phase_steps[0]); timer.count("#rcp_steps_mid", (double)rcpar.last.phase_steps[1]);
                    timer.count("#rcp_steps_late", (double)rcpar.last.phase_steps[2]); timer.count("#rcp_steps_follow", (double)rcpar.last.phase_steps[3]);
                    timer.count("#rcp_decided_serial", rcpar.last.decided_serial); timer.count("#rcp_pilot_jobs", rcpar.last.pilot_jobs);
                    timer.count("#rcp_est_steps", (double)rcpar.last.est_steps); timer.count("#rcp_serial_equiv", (double)rcpar.last.serial_equiv);
                }
            }
            else if (rc_variant == 7) CR_LAUNCH(k_range_chain<7>, gchain, dim3(128), stream, b_cinm.as<uint4>(), b_cins.as<uint4>(), b_escord.as<uint32_t>(),
                      b_streams.as<RcStream>(), nstr, b_qm.as<uint32_t>(), b_shm.as<uint32_t>(), b_qs.as<uint32_t>(), b_shs.as<uint32_t>());
            else if (rc_variant == 2) CR_LAUNCH(k_range_chain<2>, gchain, dim3(128), stream, b_cinm.as<uint4>(), b_cins.as<uint4>(), b_escord.as<uint32_t>(),
                      b_streams.as<RcStream>(), nstr, b_qm.as<uint32_t>(), b_shm.as<uint32_t>(), b_qs.as<uint32_t>(), b_shs.as<uint32_t>());
            else if (rc_variant == 6) CR_LAUNCH(k_range_chain<6>, gchain, dim3(128), stream, b_cinm.as<uint4>(), b_cins.as<uint4>(), b_escord.as<uint32_t>(),
                      b_streams.as<RcStream>(), nstr, b_qm.as<uint32_t>(), b_shm.as<uint32_t>(), b_qs.as<uint32_t>(), b_shs.as<uint32_t>());
            else if (rc_variant == 5) CR_LAUNCH(k_range_chain<5>, gchain, dim3(128), stream, b_cinm.as<uint4>(), b_cins.as<uint4>(), b_escord.as<uint32_t>(),
                      b_streams.as<RcStream>(), nstr, b_qm.as<uint32_t>(), b_shm.as<uint32_t>(), b_qs.as<uint32_t>(), b_shs.as<uint32_t>());
            else if (rc_variant == 4) CR_LAUNCH(k_range_chain<4>, gchain, dim3(128), stream, b_cinm.as<uint4>(), b_cins.as<uint4>(), b_escord.as<uint32_t>(),
                      b_streams.as<RcStream>(), nstr, b_qm.as<uint32_t>(), b_shm.as<uint32_t>(), b_qs.as<uint32_t>(), b_shs.as<uint32_t>());
            else CR_LAUNCH(k_range_chain<3>, gchain, dim3(128), stream, b_cinm.as<uint4>(), b_cins.as<uint4>(), b_escord.as<uint32_t>(),
                      b_streams.as<RcStream>(), nstr, b_qm.as<uint32_t>(), b_shm.as<uint32_t>(), b_qs.as<uint32_t>(), b_shs.as<uint32_t>());
        }
        timer.mark("range_chain");
        if (rc_variant >= 4) {
            if (ntm) CR_LAUNCH(k_msb_to_shifts, dim3(cr_div_up(ntm, 256)), dim3(256), stream, b_shm.as<uint32_t>(), (uint64_t)ntm);
            if (nts) CR_LAUNCH(k_msb_to_shifts, dim3(cr_div_up(nts, 256)), dim3(256), stream, b_shs.as<uint32_t>(), (uint64_t)nts);
        }
        CR_TRY(cr_exclusive_sum(prims, b_shm.as<uint32_t>(), b_bm.as<uint32_t>(), ntm + 1));
        CR_TRY(cr_exclusive_sum(prims, b_shs.as<uint32_t>(), b_bs.as<uint32_t>(), nts + 1));
        CR_LAUNCH(k_stream_totals, dim3(cr_div_up(nstr, 64)), dim3(64), stream, b_streams.as<RcStream>(), nstr, b_escord.as<uint32_t>(), b_bm.as<uint32_t>(), b_bs.as<uint32_t>(), b_stot.as<StreamTotals>());
        std::vector<StreamTotals> tot;
        CR_TRY(download(tot, b_stot.p, nstr));
        res.assign(nstr, RcResult());
        std::vector<LowStream> lsm, lss; std::vector<RcStream> rsm, rss, fallback; std::vector<uint32_t> fb_index;
        size_t dsum_total = 0; uint32_t maxlen = 0;
        for (uint32_t i = 0; i < nstr; i++) {
            const bool par = !streams[i].is_main || tot[i].shifts < streams[i].limit;      // bytes out <= shifts: no "cannot compress" possible
            if (!par) { fallback.push_back(streams[i]); fb_index.push_back(i); continue; }
            LowStream L; L.tri_begin = tot[i].tri_begin; L.tri_end = tot[i].tri_end; L.dsum_off = dsum_total; L.length = tot[i].shifts + 5; L.is_main = streams[i].is_main;
            dsum_total += L.length; if (L.length > maxlen) maxlen = L.length;
            if (L.length > streams[i].out_cap) return CRGPU_ERR_ARG;
            if (streams[i].is_main) { lsm.push_back(L); rsm.push_back(streams[i]); } else { lss.push_back(L); rss.push_back(streams[i]); }
            res[i].nbytes = L.length; res[i].aborted = 0;
        }
        // k_low_scatter finds the stream of a triple by binary search over tri_begin: keep the side list ordered (LZ77 lays its
        // three side streams per block out by kind, not by block)
        if (lss.size() > 1) {
            std::vector<size_t> order(lss.size());
            for (size_t k = 0; k < order.size(); k++) order[k] = k;
            std::sort(order.begin(), order.end(), [&](size_t a, size_t b) {      // empty streams first among equal starts: the search takes the last one
                return lss[a].tri_begin != lss[b].tri_begin ? lss[a].tri_begin < lss[b].tri_begin : lss[a].tri_end < lss[b].tri_end; });
            std::vector<LowStream> l2(lss.size()); std::vector<RcStream> r2(rss.size());
            for (size_t k = 0; k < order.size(); k++) { l2[k] = lss[order[k]]; r2[k] = rss[order[k]]; }
            lss.swap(l2); rss.swap(r2);
        }
        CR_TRY(b_dsum.reserve(dsum_total * 4 + 16));
        CR_CUDA(cudaMemsetAsync(b_dsum.p, 0, dsum_total * 4, stream));
        if (!lsm.empty()) {
            CR_TRY(upload(b_lsm, lsm)); CR_TRY(upload(b_rsm, rsm));
            if (ntm) CR_LAUNCH(k_low_scatter, dim3(cr_div_up(ntm, 256)), dim3(256), stream, b_dense.as<Tri>(), b_qm.as<uint32_t>(), b_bm.as<uint32_t>(), b_lsm.as<LowStream>(), (uint32_t)lsm.size(), (uint64_t)ntm, b_dsum.as<uint32_t>());
        }
        if (!lss.empty()) {
            CR_TRY(upload(b_lss, lss)); CR_TRY(upload(b_rss, rss));
            if (nts) CR_LAUNCH(k_low_scatter, dim3(cr_div_up(nts, 256)), dim3(256), stream, b_denseside.as<Tri>(), b_qs.as<uint32_t>(), b_bs.as<uint32_t>(), b_lss.as<LowStream>(), (uint32_t)lss.size(), (uint64_t)nts, b_dsum.as<uint32_t>());
        }
        if (!lsm.empty()) CR_LAUNCH(k_low_carry, dim3(cr_div_up(maxlen, 256), (unsigned)lsm.size()), dim3(256), stream, b_lsm.as<LowStream>(), b_dsum.as<uint32_t>(), b_rsm.as<RcStream>(), b_rcout.as<uint8_t>());
        if (!lss.empty()) CR_LAUNCH(k_low_carry, dim3(cr_div_up(maxlen, 256), (unsigned)lss.size()), dim3(256), stream, b_lss.as<LowStream>(), b_dsum.as<uint32_t>(), b_rss.as<RcStream>(), b_rcout.as<uint8_t>());
        if (!fallback.empty()) {                       // barely compressible streams: exact serial coder decides "cannot compress"
            CR_TRY(upload(b_fb, fallback));
            CR_LAUNCH(k_range_encode_warp, dim3(cr_div_up(fallback.size() * 32, 128)), dim3(128), stream, b_dense.as<Tri>(), b_denseside.as<Tri>(), b_escord.as<uint32_t>(),
                      b_fb.as<RcStream>(), (uint32_t)fallback.size(), b_rcout.as<uint8_t>(), b_rcres.as<RcResult>());
            std::vector<RcResult> fr;
            CR_TRY(download(fr, b_rcres.p, fallback.size()));
            for (size_t k = 0; k < fallback.size(); k++) res[fb_index[k]] = fr[k];
        }
        rc_done = true;
    }
#endif
    if (!rc_done) {
        CR_LAUNCH(k_range_encode, dim3(cr_div_up(streams.size(), 32)), dim3(32), stream, b_dense.as<Tri>(), b_denseside.as<Tri>(), b_escord.as<uint32_t>(),
                  b_streams.as<RcStream>(), (uint32_t)streams.size(), b_rcout.as<uint8_t>(), b_rcres.as<RcResult>());
    }
    timer.mark("range_coder");
    std::vector<uint32_t> hctx;
    if (!rc_done) CR_TRY(download(res, b_rcres.p, streams.size()));
    CR_TRY(download(hctx, b_ctxout.p, 1)); chain_ctx = hctx[0];

    // ---- payload layout + headers (src/rolzmain/cr-coder.c:241-262, src/ropmain/cr-coder.c:212-228)
    std::vector<CopyDesc> copies; std::vector<HeaderDesc> hdrs(nb);
    size_t pos = out_base;
    // a block the coder gave up on while more blocks of the chain follow: report where, for the exact re-run (encode_blocks)
    for (uint32_t b = 0; b < nb; b++) {
        const RcStream& m = streams[(size_t)b * spb]; const RcResult& rm = res[(size_t)b * spb];
        if (!rm.aborted || blk[b].cut || (chain_ends && b == nb - 1)) continue;
        if (!exact_aborts) {
            fprintf(stderr, "crgpu: block %u of %u hit 'cannot compress' in the middle of a model chain (SURVEY.md F11)\n", b, nb);
            return CRGPU_ERR_MIDCHAIN_ABORT;
        }
        std::vector<uint32_t> got;
        CR_TRY(b_abort.reserve(16));
        CR_LAUNCH(k_abort_event, dim3(1), dim3(1), stream, b_escord.as<uint32_t>(), m.ev_begin, m.ev_end, rm.abort_tri, b_abort.as<uint32_t>());
        CR_TRY(download(got, b_abort.p, 1));
        AbortPos f = { dD, d_blocks, scan, variant, b, got[0], b_abort.as<uint32_t>() + 1 };
        CR_LAUNCH(k_chain_walk<AbortPos>, dim3(cr_div_up(nchunk, 64)), dim3(64), stream, b_span.as<uint8_t>(), d_segs, nb, nchunk, b_entry.as<uint8_t>(), f);
        CR_TRY(download(got, b_abort.p, 2));
        if (got[1] == 0 || got[1] > blk[b].size) { fprintf(stderr, "crgpu: internal error: no abort position for block %u (event %u)\n", b, got[0]); return CRGPU_ERR_MIDCHAIN_ABORT; }
        retry_block = b; retry_cut = got[1];
        timer.finish();
        return CR_RETRY_CUT;
    }
    for (uint32_t b = 0; b < nb; b++) {
        const RcStream& m = streams[(size_t)b * spb]; const RcResult& rm = res[(size_t)b * spb];
        const bool stored = rm.aborted || blk[b].cut || (variant == CR_LZP && blk[b].size < 16);
        HeaderDesc& h = hdrs[b];
        memset(&h, 0, sizeof h);
        uint8_t* hp = h.bytes + prefix;
        uint32_t payload;
        blk[b].raw = stored; blk[b].out_off = pos;
        if (stored) {
            payload = hdr_size + blk[b].size;
            CopyDesc c = { blk[b].off, pos + prefix + hdr_size, blk[b].size, 1 }; copies.push_back(c);
        } else if (variant == CR_ROLZ) {
            const RcStream& s = streams[(size_t)b * spb + 1]; const RcResult& rs = res[(size_t)b * spb + 1];
            payload = 16 + rm.nbytes + rs.nbytes;
            hp[0] = hfirst[b * 16]; hp[1] = 1; hp[2] = hesc[b];
            uint32_t v[3] = { blk[b].size, num_idx[b], 16 + rm.nbytes };
            memcpy(hp + 4, v, 12);
            CopyDesc c0 = { m.out_off, pos + prefix + 16, rm.nbytes, 0 }; copies.push_back(c0);
            CopyDesc c1 = { s.out_off, pos + prefix + 16 + rm.nbytes, rs.nbytes, 0 }; copies.push_back(c1);
        } else if (variant == CR_LZ77) {                             // block_header, src/roxmain/cr-coder.c:68-80,276-294
            const uint32_t nby[4] = { rm.nbytes, res[(size_t)b * spb + 1].nbytes, res[(size_t)b * spb + 2].nbytes, res[(size_t)b * spb + 3].nbytes };
            payload = 32 + nby[0] + nby[1] + nby[2] + nby[3];
            hp[0] = 1; hp[1] = (uint8_t)x_match_min(blk[b].size); hp[2] = hesc[b];
            uint32_t v[7] = { blk[b].size, x_hdr_counts[(size_t)b * 3], x_hdr_counts[(size_t)b * 3 + 1], x_hdr_counts[(size_t)b * 3 + 2],
                              32 + nby[0], 32 + nby[0] + nby[1], 32 + nby[0] + nby[1] + nby[2] };
            memcpy(hp + 4, v, 28);
            size_t at = pos + prefix + 32;
            for (int k = 0; k < 4; k++) { CopyDesc c = { streams[(size_t)b * spb + k].out_off, at, nby[k], 0 }; copies.push_back(c); at += nby[k]; }
        } else {
            payload = 20 + rm.nbytes;
            hp[0] = 1; memcpy(hp + 4, &blk[b].size, 4); hp[8] = hesc[b]; memcpy(hp + 9, &hfirst[b * 16], 9);
            CopyDesc c0 = { m.out_off, pos + prefix + 20, rm.nbytes, 0 }; copies.push_back(c0);
        }
        if (prefix) memcpy(h.bytes, &payload, 4);
        if (prefix == 6) { h.bytes[4] = blk[b].filt; h.bytes[5] = blk[b].prec; }
        h.dst = pos; h.len = prefix + hdr_size;
        blk[b].out_size = payload;
        pos += prefix + payload;
    }
    out_total = pos - out_base;
    if (out.cap < pos) {   // grow, preserving what earlier windows wrote
        DevBuf nb2; CR_TRY(nb2.reserve(pos + pos / 4));
        if (out.p && out_base) CR_CUDA(cudaMemcpyAsync(nb2.p, out.p, out_base, cudaMemcpyDeviceToDevice, stream));
        CR_CUDA(cudaStreamSynchronize(stream));
        out.release(); out = nb2;
    }
    CR_TRY(upload(b_copy, copies)); CR_TRY(upload(b_hdr, hdrs));
    CR_LAUNCH(k_write_headers, dim3(cr_div_up(nb, 64)), dim3(64), stream, b_hdr.as<HeaderDesc>(), nb, out.as<uint8_t>());
    if (!copies.empty()) CR_LAUNCH(k_copy_segments, dim3(64, (unsigned)copies.size()), dim3(256), stream, b_copy.as<CopyDesc>(), b_rcout.as<uint8_t>(), dD, out.as<uint8_t>());
    timer.mark("assemble");
    CR_CUDA(cudaStreamSynchronize(stream));
    timer.finish();
    return CRGPU_OK;
}

// encode_window with the exact replay of mid-chain "cannot compress" blocks (see the comment in LzChain).  Same contract as
// encode_window; the blocks may end up being encoded in several consecutive sub-windows, which changes no output byte.
inline int LzChain::encode_blocks(const uint8_t* dD, std::vector<BlockIO>& blk, int prefix_mode, bool chain_ends, DevBuf& out, size_t out_base, size_t& out_total) {
    const size_t nb = blk.size();
    out_total = 0;
    if (nb == 0) return CRGPU_OK;
    if (!exact_aborts || (nb == 1 && chain_ends)) return encode_window(dD, blk, prefix_mode, chain_ends, out, out_base, out_total);
    size_t done = 0, take = nb;
    while (done < nb) {
        if (take > nb - done) take = nb - done;
        const bool tail = done + take == nb;
        std::vector<BlockIO> sub(blk.begin() + done, blk.begin() + done + take);
        size_t wrote = 0;
        CR_TRY(copy_state(true));
        int rc = encode_window(dD, sub, prefix_mode, chain_ends && tail, out, out_base + out_total, wrote);
        if (rc == CRGPU_OK) {
            std::copy(sub.begin(), sub.end(), blk.begin() + done);
            out_total += wrote; done += take;
            take = take * 2;                                  // grow again after a window without incident
            continue;
        }
        if (rc != CR_RETRY_CUT) return rc;
        // blocks [done, done + retry_block) are fine, block done + retry_block is stored and leaves the models where its cut says
        const uint32_t rb = retry_block;
        CR_TRY(copy_state(false));
        std::vector<BlockIO> head(blk.begin() + done, blk.begin() + done + rb + 1);
        head[rb].cut = retry_cut;
        rc = encode_window(dD, head, prefix_mode, false, out, out_base + out_total, wrote);
        if (rc != CRGPU_OK) return rc == CR_RETRY_CUT ? CRGPU_ERR_MIDCHAIN_ABORT : rc;
        std::copy(head.begin(), head.end(), blk.begin() + done);
        out_total += wrote; done += rb + 1; cut_blocks++;
        take = 1;                                             // incompressible blocks come in runs: probe one block at a time
    }
    return CRGPU_OK;
}
