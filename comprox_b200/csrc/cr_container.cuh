// cr_container.cuh -- whole-container compression: what cr_main does between fopen and fclose
// (src/main.c:137-218), restructured so that the GPU sees the whole input at once.
//
//   magic | u32 dict_len | lzencode(dic_lcp_encode(dicpick(file)))          <- chain 0, models reset after
//         | { u32 len, u8 filt, u8 prec, payload }*                           <- chain 1, models carried (F2)
// The serial per-block loop of the reference becomes: one dicpick pass over the file, then per WINDOW of
// blocks: filters -> dictionary substitution -> lzencode chain (cr_lzchain.cuh).  The container bytes are
// assembled on the device and leave with a single copy.
#pragma once
#include <thread>
#include "cr_lzchain.cuh"
#include "cr_dict.cuh"
#include "cr_filter.cuh"
#include "cr_hostdict.h"
#include <chrono>

struct CrConfig {
    uint32_t block_size;     // bytes, reference default 16 MiB (src/main.c:62)
    int filt;                // -F
    int prec;                // -p
    int flexible;            // -f (ROLZ, LZ77)
    uint64_t window_bytes;   // raw bytes per window (0 = default)
};

struct Compressor {
    LzChain* chain = nullptr;
    cudaStream_t stream = 0;
    DevBuf d_raw, d_D, d_out, d_dic;
    DevBuf t_key, t_count, t_first, t_stats, t_entries;          // dicpick table
    DevBuf d_trie_edge, d_trie_id;
    DevBuf b_dcoff, b_dclist;
    int dc_listed = 1;                      // 1: word starts are listed and the trie is walked one listed word per lane (k_dc_walk); 0: k_dc_spans
    DevBuf b_subs, b_hist, b_esc10, b_escmask, b_span, b_hit, b_segs, b_xt, b_entry, b_cnt, b_scan, b_chunk0, b_hdr, b_copy, b_segoff, b_seglen;
    HdTrie trie;
    FilterHost filt;
    const uint8_t* staged_ptr = nullptr; uint64_t staged_n = 0;
    // The dictionary payload is a model chain of its own: reset_models() stands before and after its lzencode (src/main.c:166-167).  It
    // therefore runs on a second LzChain (own models, own buffers, own stream, driven by a helper thread) beside the data blocks: its
    // range walk -- fresh models, small sums, nothing to cut: ~200 000 symbols one after the other, 2.6 ms -- no longer stands in front
    // of them.  dict_mode 0 turns this off (the payload is then coded on the main chain first, as before).
    // the input arrives in pieces on a copy stream; the word count of a piece starts when the piece (and the one behind it: a word may
    // straddle the border) has landed, so the H2D copy hides behind k_dp_count instead of standing in front of it
#ifndef CRGPU_SIM
    cudaStream_t copy_stream = 0;
    std::vector<cudaEvent_t> copy_ev;
    cudaEvent_t ev_counted = 0;
    // lowest priority: the check kernel that runs on it (k_dp_verify) must not hold up the kernels the host is waiting for
    int make_copy_stream() {
        int lo = 0, hi = 0;
        CR_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CR_CUDA(cudaStreamCreateWithPriority(&copy_stream, cudaStreamNonBlocking, lo));
        return CRGPU_OK;
    }
#endif
    uint64_t copy_chunk = 0;                // 0: the input is resident as a whole (no events to wait for)
    LzChain* dict_chain = nullptr;
    cudaStream_t dict_stream = 0;
    DevBuf d_dictout;
    int dict_mode = 1;
    bool crowded = false;                   // set by crgpu_compress_batch when three or more handles share the device: no extra streams
                                            // (every stream beyond the 32 hardware queues aliases behind some other handle's serial walk)
    int dp_tiles = 0;                       // 1: k_dp_count_tiles (shared-memory table per tile) instead of k_dp_count; measured slower, see cr_dict.cuh

    void release() {
        DevBuf* all[] = { &d_raw, &d_D, &d_out, &d_dic, &t_key, &t_count, &t_first, &t_stats, &t_entries, &d_trie_edge, &d_trie_id, &b_subs, &b_hist,
                          &b_esc10, &b_escmask, &b_span, &b_hit, &b_segs, &b_xt, &b_entry, &b_cnt, &b_scan, &b_chunk0, &b_hdr, &b_copy, &b_segoff, &b_seglen };
        for (DevBuf* b : all) b->release();
        filt.release();
        d_dictout.release(); b_dcoff.release(); b_dclist.release();
        if (dict_chain) { dict_chain->release(); delete dict_chain; dict_chain = nullptr; }
#ifndef CRGPU_SIM
        if (dict_stream) { cudaStreamDestroy(dict_stream); dict_stream = 0; }
        if (copy_stream) { cudaStreamDestroy(copy_stream); copy_stream = 0; }
        for (cudaEvent_t e : copy_ev) cudaEventDestroy(e);
        copy_ev.clear();
        if (ev_counted) { cudaEventDestroy(ev_counted); ev_counted = 0; }
#endif
    }
    template <class T> int upload(DevBuf& b, const std::vector<T>& v) { return chain->upload(b, v); }
    template <class T> int download(std::vector<T>& v, const void* src, size_t n) { return chain->download(v, src, n); }

    int stage(const uint8_t* in, uint64_t n) {
        CR_TRY(d_raw.reserve(n + 256));
        if (n) CR_CUDA(cudaMemcpyAsync(d_raw.p, in, n, cudaMemcpyHostToDevice, stream));
        CR_CUDA(cudaMemsetAsync(d_raw.as<uint8_t>() + n, 0, 128, stream));
        CR_CUDA(cudaStreamSynchronize(stream));
        staged_ptr = nullptr;                  // only crgpu_stage_input arms the "already resident" fast path of compress()
        return CRGPU_OK;
    }
    int dicpick(const uint8_t* h_in, const uint8_t* d_in, uint64_t n, std::string& text);
    int dicpick_epochs(const uint8_t* d_in, uint64_t n);
    int load_dictionary(const std::string& text);
    int dict_encode_window(const uint8_t* d_rawwin, const std::vector<uint64_t>& roff, const std::vector<uint32_t>& rsize, std::vector<BlockIO>& blk, size_t& dtotal);
    int compress(const CrConfig& cfg, const uint8_t* in, uint64_t n, uint8_t* out, uint64_t out_cap, uint64_t* out_n);
};

// ------------------------------------------------------------------ dicpick (src/cr-dicpick.c:164-259)
inline int Compressor::dicpick(const uint8_t* h_in, const uint8_t* d_in, uint64_t n, std::string& text) {
    CR_TRY(t_key.reserve((size_t)DP_SLOTS * 8)); CR_TRY(t_count.reserve((size_t)DP_SLOTS * 4)); CR_TRY(t_first.reserve((size_t)DP_SLOTS * 4));
    CR_TRY(t_stats.reserve(64)); CR_TRY(t_entries.reserve((size_t)DP_MAXWORDS * sizeof(DpEntry)));
    CR_CUDA(cudaMemsetAsync(t_key.p, 0, (size_t)DP_SLOTS * 8, stream));
    CR_CUDA(cudaMemsetAsync(t_count.p, 0, (size_t)DP_SLOTS * 4, stream));
    CR_CUDA(cudaMemsetAsync(t_first.p, 0xFF, (size_t)DP_SLOTS * 4, stream));
    CR_CUDA(cudaMemsetAsync(t_stats.p, 0, 64, stream));
    DpTable T = { t_key.as<unsigned long long>(), t_count.as<uint32_t>(), t_first.as<uint32_t>(), t_stats.as<uint32_t>() };
    const uint64_t step = 1ull << 30;
#ifndef CRGPU_SIM
    if (copy_chunk) {
        const uint64_t nchunks = (n + copy_chunk - 1) / copy_chunk;
        for (uint64_t c = 0; c < nchunks; c++) {
            CR_CUDA(cudaStreamWaitEvent(stream, copy_ev[c + 1 < nchunks ? c + 1 : c], 0));
            const uint64_t x0 = c * copy_chunk, x1 = x0 + copy_chunk < n ? x0 + copy_chunk : n;
            if (dp_tiles) CR_LAUNCH(k_dp_count_tiles, dim3(cr_div_up(x1 - x0, DPS_TILE)), dim3(256), stream, d_in, n, x0, x1, T);
            else CR_LAUNCH(k_dp_count, dim3(cr_div_up(x1 - x0, 256)), dim3(256), stream, d_in, n, x0, x1, T);
        }
        copy_chunk = 0;
    } else
#endif
    for (uint64_t x0 = 0; x0 < n; x0 += step) {
        uint64_t x1 = x0 + step < n ? x0 + step : n;
#ifndef CRGPU_SIM
        if (dp_tiles) CR_LAUNCH(k_dp_count_tiles, dim3(cr_div_up(x1 - x0, DPS_TILE)), dim3(256), stream, d_in, n, x0, x1, T);
        else
#endif
        CR_LAUNCH(k_dp_count, dim3(cr_div_up(x1 - x0, 256)), dim3(256), stream, d_in, n, x0, x1, T);
    }
    // every occurrence must spell the word its table entry stands for (a 64-bit hash collision is an error, not a wrong dictionary).
    // Nothing downstream waits for the answer, so the check runs on a second stream beside the collect kernel and the host's
    // ordering of the words; its flag is read at the end of this function.
    cudaStream_t vstream = stream;
#ifndef CRGPU_SIM
    if (!crowded) {
    if (!copy_stream) CR_TRY(make_copy_stream());
    if (!ev_counted) CR_CUDA(cudaEventCreateWithFlags(&ev_counted, cudaEventDisableTiming));
    CR_CUDA(cudaEventRecord(ev_counted, stream));
    CR_CUDA(cudaStreamWaitEvent(copy_stream, ev_counted, 0));
    vstream = copy_stream;
    }
#endif
    for (uint64_t x0 = 0; x0 < n; x0 += step) {
        uint64_t x1 = x0 + step < n ? x0 + step : n;
        CR_LAUNCH(k_dp_verify, dim3(cr_div_up(x1 - x0, 256)), dim3(256), vstream, d_in, n, x0, x1, T);
    }
    auto collided = [&](bool* yes) -> int {
        uint32_t flag = 0;
        CR_CUDA(cudaMemcpyAsync(&flag, t_stats.as<uint32_t>() + 8, 4, cudaMemcpyDeviceToHost, vstream));
        CR_CUDA(cudaStreamSynchronize(vstream));
        *yes = flag != 0;
        return CRGPU_OK;
    };
    CR_LAUNCH(k_dp_collect, dim3(DP_SLOTS / 256), dim3(256), stream, T, d_in, n, t_entries.as<DpEntry>(), DP_MAXWORDS);
    chain->timer.mark("dp_kernels");
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!getenv("CRGPU_TIMING")) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "crgpu timing: %-16s %.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    };
    std::vector<uint32_t> stats;
    CR_TRY(download(stats, t_stats.p, 4));
    if (stats[1] & 1u) {
        // more than 325000 distinct words: replay the reference's prune epochs exactly (cr_dict.cuh)
        bool bad = false;
        CR_TRY(collided(&bad));                                          // (the epochs rebuild the tables the check reads)
        if (bad) return CRGPU_ERR_HASH_COLLISION;
        CR_TRY(dicpick_epochs(d_in, n));
        CR_TRY(download(stats, t_stats.p, 4));
        if (stats[1] & 4u) return CRGPU_ERR_VOCAB_OVERFLOW;          // more distinct words in one window than the table holds
    }
    lap("dp sync+stats");
    std::vector<DpEntry> ent;
    CR_TRY(download(ent, t_entries.p, stats[2]));
    lap("dp entries d2h");
    std::vector<HdWord> words(ent.size());
    for (size_t i = 0; i < ent.size(); i++) { words[i].k[0] = ent[i].k[0]; words[i].k[1] = ent[i].k[1]; words[i].k[2] = ent[i].k[2]; words[i].count = ent[i].count; words[i].len = ent[i].len; }
    (void)h_in;
    lap("dp words");
    if (getenv("CRGPU_TIMING")) fprintf(stderr, "crgpu timing: %zu words with count > 5\n", words.size());
    text = hd_dictionary_text(words);
    lap("dp text");
    bool bad = false;
    CR_TRY(collided(&bad));
    lap("dp verify wait");
    chain->timer.mark("dp_host");
    return bad ? CRGPU_ERR_HASH_COLLISION : CRGPU_OK;
}

// Exact vocabulary-overflow path.  Leaves the final table in t_key/t_count/t_first and the selected
// entries in t_entries / stats[2], like the fast path.
inline int Compressor::dicpick_epochs(const uint8_t* d_in, uint64_t n) {
    DevBuf k2, c2, f2, list, small;
    CR_TRY(k2.reserve((size_t)DP_SLOTS * 8)); CR_TRY(c2.reserve((size_t)DP_SLOTS * 4)); CR_TRY(f2.reserve((size_t)DP_SLOTS * 4));
    CR_TRY(list.reserve((size_t)DP_SLOTS * 4)); CR_TRY(small.reserve(64));
    auto clear = [&](DevBuf& k, DevBuf& c, DevBuf& f) -> int {
        CR_CUDA(cudaMemsetAsync(k.p, 0, (size_t)DP_SLOTS * 8, stream)); CR_CUDA(cudaMemsetAsync(c.p, 0, (size_t)DP_SLOTS * 4, stream));
        CR_CUDA(cudaMemsetAsync(f.p, 0xFF, (size_t)DP_SLOTS * 4, stream));
        return CRGPU_OK;
    };
    DevBuf *ka = &t_key, *ca = &t_count, *fa = &t_first, *kb = &k2, *cb = &c2, *fb = &f2;
    CR_TRY(clear(*ka, *ca, *fa));
    CR_CUDA(cudaMemsetAsync(t_stats.p, 0, 64, stream));
    uint32_t alive = 0;
    const uint64_t window = 64ull << 20;
    int rc = CRGPU_OK;
    for (uint64_t pos = 0; pos < n && rc == CRGPU_OK;) {
        DpTable T = { ka->as<unsigned long long>(), ca->as<uint32_t>(), fa->as<uint32_t>(), t_stats.as<uint32_t>() };
        const uint64_t wend = pos + window < n ? pos + window : n;
        CR_LAUNCH(k_dp_first, dim3(cr_div_up(wend - pos, 256)), dim3(256), stream, d_in, n, pos, wend, T);
        CR_CUDA(cudaMemsetAsync(small.p, 0, 8, stream));
        CR_LAUNCH(k_dp_list_new, dim3(DP_SLOTS / 256), dim3(256), stream, T, (uint32_t)pos, list.as<uint32_t>(), DP_SLOTS, small.as<uint32_t>());
        std::vector<uint32_t> cnt, st;
        CR_TRY(download(cnt, small.p, 1));
        CR_TRY(download(st, t_stats.p, 4));
        if (st[1] & 4u) break;                                           // table full: reported by the caller
        const uint32_t n_new = cnt[0], need = DP_MAXWORDS - alive;
        if (n_new < need) {                                              // no prune inside this window
            CR_LAUNCH(k_dp_count, dim3(cr_div_up(wend - pos, 256)), dim3(256), stream, d_in, n, pos, wend, T);
            alive += n_new; pos = wend;
            continue;
        }
        std::vector<uint32_t> firsts;
        CR_TRY(download(firsts, list.p, n_new));
        std::nth_element(firsts.begin(), firsts.begin() + (need - 1), firsts.end());
        const uint64_t X = firsts[need - 1];                             // the word starting here is the 325001st
        CR_LAUNCH(k_dp_count, dim3(cr_div_up(X + 1 - pos, 256)), dim3(256), stream, d_in, n, pos, X + 1, T);
        CR_CUDA(cudaMemsetAsync(small.p, 0xFF, 4, stream));
        CR_LAUNCH(k_dp_min_count, dim3(DP_SLOTS / 256), dim3(256), stream, T, small.as<uint32_t>());
        std::vector<uint32_t> mn;
        CR_TRY(download(mn, small.p, 1));
        CR_TRY(clear(*kb, *cb, *fb));
        CR_CUDA(cudaMemsetAsync(t_stats.p, 0, 4, stream));
        DpTable B = { kb->as<unsigned long long>(), cb->as<uint32_t>(), fb->as<uint32_t>(), t_stats.as<uint32_t>() };
        CR_LAUNCH(k_dp_rebuild, dim3(DP_SLOTS / 256), dim3(256), stream, T, B, mn[0] + 5);
        CR_TRY(download(st, t_stats.p, 1));
        alive = st[0];
        std::swap(ka, kb); std::swap(ca, cb); std::swap(fa, fb);
        pos = X + 1;
    }
    // the final table must live in t_key/t_count/t_first
    if (ka != &t_key) {
        CR_CUDA(cudaMemcpyAsync(t_key.p, ka->p, (size_t)DP_SLOTS * 8, cudaMemcpyDeviceToDevice, stream));
        CR_CUDA(cudaMemcpyAsync(t_count.p, ca->p, (size_t)DP_SLOTS * 4, cudaMemcpyDeviceToDevice, stream));
        CR_CUDA(cudaMemcpyAsync(t_first.p, fa->p, (size_t)DP_SLOTS * 4, cudaMemcpyDeviceToDevice, stream));
    }
    DpTable T = { t_key.as<unsigned long long>(), t_count.as<uint32_t>(), t_first.as<uint32_t>(), t_stats.as<uint32_t>() };
    CR_CUDA(cudaMemsetAsync(t_stats.as<uint32_t>() + 2, 0, 4, stream));
    CR_LAUNCH(k_dp_collect, dim3(DP_SLOTS / 256), dim3(256), stream, T, d_in, n, t_entries.as<DpEntry>(), DP_MAXWORDS);
    CR_CUDA(cudaStreamSynchronize(stream));
    k2.release(); c2.release(); f2.release(); list.release(); small.release();
    return rc;
}

inline int Compressor::load_dictionary(const std::string& text) {
    auto t0 = std::chrono::steady_clock::now();
    trie.load(text.c_str());
    if (getenv("CRGPU_TIMING")) fprintf(stderr, "crgpu timing: trie.load        %.3f ms (%zu nodes, table %u)\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), trie.id.size(), trie.mask + 1);
    CR_TRY(upload(d_trie_edge, trie.edge)); CR_TRY(upload(d_trie_id, trie.id));
    return CRGPU_OK;
}

// ------------------------------------------------------------------ dictionary_encode for a window of blocks
// (src/cr-diccode.c:142-221).  Produces the dictionary-coded blocks back to back (16-byte aligned) in d_D.
inline int Compressor::dict_encode_window(const uint8_t* d_rawwin, const std::vector<uint64_t>& roff, const std::vector<uint32_t>& rsize,
                                          std::vector<BlockIO>& blk, size_t& dtotal) {
    const uint32_t nb = (uint32_t)rsize.size();
    Prims& prims = chain->prims;
    // ---- sub-chunks: pairs of up to 1 000 000 bytes (:173-180); the second of a pair may be empty
    std::vector<DcSub> subs;
    std::vector<uint32_t> first_sub(nb + 1);
    uint32_t maxsize = 0;
    for (uint32_t b = 0; b < nb; b++) {
        first_sub[b] = (uint32_t)subs.size();
        for (uint32_t pos = 0; pos < rsize[b];) {
            for (int k = 0; k < 2; k++) {
                uint32_t s = rsize[b] - pos < DC_SUB ? rsize[b] - pos : DC_SUB;
                DcSub S; S.off = roff[b] + pos; S.size = s; S.block = b; S.out = 0;
                subs.push_back(S);
                pos += s;
            }
        }
        if (rsize[b] > maxsize) maxsize = rsize[b];
    }
    first_sub[nb] = (uint32_t)subs.size();
    const uint32_t nsub = (uint32_t)subs.size();
    size_t rawtotal = 0;
    for (uint32_t b = 0; b < nb; b++) if (roff[b] + rsize[b] > rawtotal) rawtotal = roff[b] + rsize[b];

    // ---- the 10 rarest bytes of every raw block (:161-171)
    CR_TRY(upload(b_segoff, roff)); CR_TRY(upload(b_seglen, rsize));
    CR_TRY(b_hist.reserve((size_t)nb * 1024)); CR_TRY(b_esc10.reserve((size_t)nb * 10 + 16)); CR_TRY(b_escmask.reserve((size_t)nb * 32));
    CR_CUDA(cudaMemsetAsync(b_hist.p, 0, (size_t)nb * 1024, stream));
    if (maxsize) CR_LAUNCH(k_hist256, dim3(cr_div_up(maxsize, CR_HIST_TILE), nb), dim3(256), stream, d_rawwin, b_segoff.as<uint64_t>(), b_seglen.as<uint32_t>(), b_hist.as<uint32_t>());
    CR_LAUNCH(k_pick_escapes, dim3(cr_div_up(nb, 64)), dim3(64), stream, b_hist.as<uint32_t>(), nb, b_esc10.as<uint8_t>(), (uint8_t*)nullptr);
    CR_LAUNCH(k_dc_escmask, dim3(cr_div_up(nb, 64)), dim3(64), stream, b_esc10.as<uint8_t>(), nb, b_escmask.as<uint32_t>());

    // ---- per-position trie walk, parse resolution, code sizes
    std::vector<ChainSeg> segs(nsub);
    for (uint32_t s = 0; s < nsub; s++) { segs[s].off = subs[s].off; segs[s].len = subs[s].size; segs[s].start = 0; }
    const uint32_t nchunk = cr_chain_layout(segs.data(), nsub);
    std::vector<uint32_t> chunk0(nsub + 1);
    for (uint32_t s = 0; s < nsub; s++) chunk0[s] = segs[s].chunk0;
    chunk0[nsub] = nchunk;
    std::vector<uint32_t> hscan(nchunk + 1, 0);
    DcTrie T = { d_trie_edge.as<HdEdge>(), d_trie_id.as<int32_t>(), trie.mask, trie.nentries, trie.level1() };
    if (nsub) {
        CR_TRY(upload(b_subs, subs)); CR_TRY(upload(b_segs, segs)); CR_TRY(upload(b_chunk0, chunk0));
        CR_TRY(b_span.reserve(rawtotal + 16)); CR_TRY(b_hit.reserve(rawtotal * 4 + 16));
        CR_TRY(b_xt.reserve((size_t)nchunk * 256 + 16)); CR_TRY(b_entry.reserve(nchunk + 16));
        CR_TRY(b_cnt.reserve((size_t)(nchunk + 1) * 4)); CR_TRY(b_scan.reserve((size_t)(nchunk + 1) * 4));
#ifndef CRGPU_SIM
        if (dc_listed) {
            const uint32_t gx = cr_div_up(DC_SUB, DCS_TH);
            const size_t nct = (size_t)gx * nsub;
            CR_TRY(b_dcoff.reserve((nct + 1) * 4 + 16)); CR_TRY(b_dclist.reserve(rawtotal / 2 * sizeof(uint2) + 64));       // a word start needs a non-letter in front of it
            CR_CUDA(cudaMemsetAsync(b_dcoff.as<uint32_t>() + nct, 0, 4, stream));
            CR_LAUNCH(k_dc_count_starts, dim3(gx, nsub), dim3(DCS_TH), stream, d_rawwin, b_subs.as<DcSub>(), b_dcoff.as<uint32_t>());
            CR_TRY(cr_exclusive_sum(prims, b_dcoff.as<uint32_t>(), b_dcoff.as<uint32_t>(), nct + 1));                     // [nct] = number of word starts
            CR_LAUNCH(k_dc_list_starts, dim3(gx, nsub), dim3(DCS_TH), stream, d_rawwin, b_subs.as<DcSub>(), b_dcoff.as<uint32_t>(), b_span.as<uint8_t>(), b_dclist.as<uint2>());
            CR_LAUNCH(k_dc_walk, dim3(cr_div_up(rawtotal / 2 + 1, DCS_TH)), dim3(DCS_TH), stream, d_rawwin, b_subs.as<DcSub>(), T, b_dclist.as<uint2>(), b_dcoff.as<uint32_t>() + nct,
                      b_span.as<uint8_t>(), b_hit.as<uint32_t>());
        } else
#endif
        CR_LAUNCH(k_dc_spans, dim3(cr_div_up(DC_SUB, 256), nsub), dim3(256), stream, d_rawwin, b_subs.as<DcSub>(), T, b_span.as<uint8_t>(), b_hit.as<uint32_t>());
        CR_CUDA(cudaMemsetAsync(b_cnt.p, 0, (size_t)(nchunk + 1) * 4, stream));
        if (nchunk) {
            CR_LAUNCH(k_chain_exits, dim3(cr_div_up(nchunk, 64)), dim3(64), stream, b_span.as<uint8_t>(), b_segs.as<ChainSeg>(), nsub, nchunk, b_xt.as<uint8_t>());
            CR_TRY(cr_chain_run_entries(*chain, chain->b_chainwork, segs, b_segs.as<ChainSeg>(), nchunk, b_xt.as<uint8_t>(), b_entry.as<uint8_t>()));
            DcCount f = { d_rawwin, b_subs.as<DcSub>(), b_span.as<uint8_t>(), b_hit.as<uint32_t>(), b_escmask.as<uint32_t>(), T.level1, b_cnt.as<uint32_t>() };
            CR_LAUNCH(k_chain_walk<DcCount>, dim3(cr_div_up(nchunk, 64)), dim3(64), stream, b_span.as<uint8_t>(), b_segs.as<ChainSeg>(), nsub, nchunk, b_entry.as<uint8_t>(), f);
        }
        CR_TRY(cr_exclusive_sum(prims, b_cnt.as<uint32_t>(), b_scan.as<uint32_t>(), nchunk + 1));
        CR_TRY(download(hscan, b_scan.p, nchunk + 1));
    }
    std::vector<uint8_t> hesc;
    CR_TRY(download(hesc, b_esc10.p, (size_t)nb * 10));

    // ---- layout: coded if smaller than raw, else raw + 0 (:208-217)
    std::vector<HeaderDesc> hdrs; std::vector<CopyDesc> copies;
    blk.assign(nb, BlockIO());
    size_t pos = 0;
    auto put_bytes = [&](uint64_t dst, const void* p, uint32_t len) { HeaderDesc h; memset(&h, 0, sizeof h); h.dst = dst; h.len = len; memcpy(h.bytes, p, len); hdrs.push_back(h); };
    for (uint32_t b = 0; b < nb; b++) {
        memset(&blk[b], 0, sizeof(BlockIO));
        uint64_t coded = 11;
        for (uint32_t s = first_sub[b]; s < first_sub[b + 1]; s += 2) coded += 8 + (hscan[chunk0[s + 1]] - hscan[chunk0[s]]) + 4 + (hscan[chunk0[s + 2]] - hscan[chunk0[s + 1]]) + 4;
        blk[b].off = pos;
        if (coded >= rsize[b]) {
            blk[b].size = rsize[b] + 1;
            if (rsize[b]) { CopyDesc c = { roff[b], pos, rsize[b], 0 }; copies.push_back(c); }
            uint8_t z = 0; put_bytes(pos + rsize[b], &z, 1);
            for (uint32_t s = first_sub[b]; s < first_sub[b + 1]; s++) subs[s].out = ~0ull;   // nothing to emit
        } else {
            blk[b].size = (uint32_t)coded;
            uint64_t p = pos;
            for (uint32_t s = first_sub[b]; s < first_sub[b + 1]; s += 2) {
                uint32_t l1 = hscan[chunk0[s + 1]] - hscan[chunk0[s]] + 4, l2 = hscan[chunk0[s + 2]] - hscan[chunk0[s + 1]] + 4;
                uint32_t pair[2] = { l1, l2 };
                put_bytes(p, pair, 8); p += 8;
                subs[s].out = p; p += l1 - 4; put_bytes(p, &subs[s].size, 4); p += 4;
                subs[s + 1].out = p; p += l2 - 4; put_bytes(p, &subs[s + 1].size, 4); p += 4;
            }
            uint8_t tail[11]; memcpy(tail, &hesc[(size_t)b * 10], 10); tail[10] = 1;
            put_bytes(p, tail, 11);
        }
        pos += ((size_t)blk[b].size + 15) & ~(size_t)15;
    }
    dtotal = pos;
    CR_TRY(d_D.reserve(dtotal + 64));
    // emit codes of the coded blocks (sub-chunks of raw blocks carry out = ~0 and emit nothing)
    if (nsub && nchunk) {
        CR_TRY(upload(b_subs, subs));
        DcEmit f = { d_rawwin, b_subs.as<DcSub>(), b_span.as<uint8_t>(), b_hit.as<uint32_t>(), b_escmask.as<uint32_t>(), b_esc10.as<uint8_t>(),
                     T.level1, T.nentries, b_scan.as<uint32_t>(), b_chunk0.as<uint32_t>(), d_D.as<uint8_t>() };
        CR_LAUNCH(k_chain_walk<DcEmit>, dim3(cr_div_up(nchunk, 64)), dim3(64), stream, b_span.as<uint8_t>(), b_segs.as<ChainSeg>(), nsub, nchunk, b_entry.as<uint8_t>(), f);
    }
    if (!hdrs.empty()) { CR_TRY(upload(b_hdr, hdrs)); CR_LAUNCH(k_write_headers, dim3(cr_div_up(hdrs.size(), 64)), dim3(64), stream, b_hdr.as<HeaderDesc>(), (uint32_t)hdrs.size(), d_D.as<uint8_t>()); }
    if (!copies.empty()) { CR_TRY(upload(b_copy, copies)); CR_LAUNCH(k_copy_segments, dim3(64, (unsigned)copies.size()), dim3(256), stream, b_copy.as<CopyDesc>(), d_rawwin, d_rawwin, d_D.as<uint8_t>()); }
    return CRGPU_OK;
}

static inline const char* cr_magic(int variant) {          // src/rolzmain/main.c:35, src/ropmain/main.c:35, src/roxmain/main.c:35
    return variant == CR_ROLZ ? "\x1f\x9d\x01\x01::0.11.0-comprolz" : variant == CR_LZP ? "\x1f\x9d\x01\x01::0.11.0-comprop" : "\x1f\x9d\x01\x01::0.11.0-comprox";
}
static inline uint64_t cr_compress_bound(uint64_t n, uint32_t block_size) {
    uint64_t nblocks = n / (block_size ? block_size : 1) + 2;
    return 64 + (1u << 20) + n + nblocks * 64;
}

inline int Compressor::compress(const CrConfig& cfg, const uint8_t* in, uint64_t n, uint8_t* out, uint64_t out_cap, uint64_t* out_n) {
    if (cfg.block_size == 0 || (n && !in) || !out || !out_n) return CRGPU_ERR_ARG;
    if (cfg.flexible && chain->variant == CR_LZP) return CRGPU_ERR_ARG;      // comprop has no -f (src/ropmain/main.c)
    if (n > (1ull << 32)) {                                                  // word positions of the dicpick table are 32 bit (the 4 GiB config fits)
        fprintf(stderr, "crgpu: inputs above 4 GiB are not supported in one container (%llu bytes); split into shards\n", (unsigned long long)n);
        return CRGPU_ERR_UNSUPPORTED;
    }
    stream = chain->stream;
    StageTimer& tm = chain->timer;
    const size_t mlen = strlen(cr_magic(chain->variant));
    if (out_cap < cr_compress_bound(n, cfg.block_size)) return CRGPU_ERR_ARG;
    memcpy(out, cr_magic(chain->variant), mlen);

    // ---- whole input to HBM
    const bool staged = staged_ptr == in && staged_n == n && d_raw.p != nullptr && !cfg.filt;   // filters modify d_raw in place
    staged_ptr = nullptr;
    copy_chunk = 0;
    if (!staged) {
        CR_TRY(d_raw.reserve(n + 256));
        CR_CUDA(cudaMemsetAsync(d_raw.as<uint8_t>() + n, 0, 128, stream));
#ifndef CRGPU_SIM
        const uint64_t CH = 16ull << 20;
        if (n > 2 * CH && !crowded) {
            if (!copy_stream) CR_TRY(make_copy_stream());
            const uint64_t nchunks = (n + CH - 1) / CH;
            while (copy_ev.size() < nchunks + 1) { cudaEvent_t e; CR_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); copy_ev.push_back(e); }
            CR_CUDA(cudaEventRecord(copy_ev[nchunks], stream));                 // the buffer exists (stream-ordered allocation) and nothing reads it any more
            CR_CUDA(cudaStreamWaitEvent(copy_stream, copy_ev[nchunks], 0));
            for (uint64_t c = 0; c < nchunks; c++) {
                const uint64_t x0 = c * CH, len = x0 + CH < n ? CH : n - x0;
                CR_CUDA(cudaMemcpyAsync(d_raw.as<uint8_t>() + x0, in + x0, len, cudaMemcpyHostToDevice, copy_stream));
                CR_CUDA(cudaEventRecord(copy_ev[c], copy_stream));
            }
            copy_chunk = CH;
        } else
#endif
        if (n) CR_CUDA(cudaMemcpyAsync(d_raw.p, in, n, cudaMemcpyHostToDevice, stream));
    }

    // ---- static dictionary: build, load, emit as its own model chain (src/main.c:156-172)
    std::string text;
    tm.begin(stream);
    CR_TRY(dicpick(in, d_raw.as<uint8_t>(), n, text));
    if (tm.enabled) { CR_CUDA(cudaStreamSynchronize(stream)); tm.finish(); }
    auto th0 = std::chrono::steady_clock::now();
    auto hlap = [&](const char* what) {
        if (!getenv("CRGPU_TIMING")) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "crgpu timing: %-16s %.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - th0).count());
        th0 = t1;
    };
    CR_TRY(load_dictionary(text));
    hlap("trie load+upload");
    std::vector<uint8_t> lcp;                  // front-coded dictionary text (made by whoever encodes it: the helper thread, or below)
    chain->flexible = cfg.flexible != 0;       // flexible_parsing is a process-wide switch: it also applies to the dictionary payload
    std::vector<BlockIO> dblk(1);
    memset(&dblk[0], 0, sizeof(BlockIO));
    size_t out_pos = 0, wrote = 0, dict_wrote = 0;
    bool dict_async = false;
#ifndef CRGPU_SIM
    std::thread dict_thread;
    int dict_rc = CRGPU_OK;
    if (dict_mode == 1 && !crowded) {
        int dev = 0;
        CR_CUDA(cudaGetDevice(&dev));
        if (!dict_chain) {
            CR_CUDA(cudaStreamCreateWithFlags(&dict_stream, cudaStreamNonBlocking));
            const cudaStream_t keep = g_cr_alloc_stream;
            g_cr_alloc_stream = dict_stream;
            dict_chain = new LzChain();
            const int rc = dict_chain->init(chain->variant, dict_stream);
            g_cr_alloc_stream = keep;
            if (rc != CRGPU_OK) return rc;
        }
        dict_chain->copy_options(*chain);
        dict_async = true;
        dict_thread = std::thread([&, dev]() {
            dict_rc = [&]() -> int {
                CR_CUDA(cudaSetDevice(dev));
                g_cr_alloc_stream = dict_stream; g_cr_alloc_async = true;
                lcp = hd_lcp_encode(text); dblk[0].size = (uint32_t)lcp.size();
                CR_TRY(dict_chain->upload(d_dic, lcp));
                CR_TRY(dict_chain->reset_models());
                size_t w = 0;
                CR_TRY(dict_chain->encode_blocks(d_dic.as<uint8_t>(), dblk, 0, true, d_dictout, 0, w));
                CR_CUDA(cudaStreamSynchronize(dict_stream));
                dict_wrote = w;
                return CRGPU_OK;
            }();
        });
    }
    struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{dict_thread};      // every return path waits for the helper
#endif
    if (!dict_async) {
        lcp = hd_lcp_encode(text); dblk[0].size = (uint32_t)lcp.size();
        CR_TRY(upload(d_dic, lcp));
        CR_TRY(chain->reset_models());
        CR_TRY(chain->encode_blocks(d_dic.as<uint8_t>(), dblk, 0, true, d_out, out_pos, wrote));
        out_pos += wrote;
    }
    CR_TRY(chain->reset_models());
    hlap("dict chain start");

    // ---- data blocks, window by window.  A trailing empty block appears when n % block_size == 0 (F8).
    const uint64_t nblocks = n / cfg.block_size + 1;
    uint64_t wbytes = cfg.window_bytes ? cfg.window_bytes : (512ull << 20);
    uint64_t per_window = wbytes / cfg.block_size; if (per_window == 0) per_window = 1;
    if (chain->variant == CR_ROLZ && per_window > RZ_MAX_BLOCKS) per_window = RZ_MAX_BLOCKS;
    filt.reset();
    int filt_flag = 0;
    for (uint64_t b0 = 0; b0 < nblocks; b0 += per_window) {
        const uint64_t b1 = b0 + per_window < nblocks ? b0 + per_window : nblocks;
        std::vector<uint64_t> roff; std::vector<uint32_t> rsize;
        const uint64_t wbase = b0 * cfg.block_size;              // offsets below are relative to the window
        uint8_t* d_win = d_raw.as<uint8_t>() + wbase;
        for (uint64_t b = b0; b < b1; b++) {
            uint64_t off = b * cfg.block_size;
            roff.push_back(off - wbase);
            rsize.push_back((uint32_t)(n - off < cfg.block_size ? n - off : cfg.block_size));
        }
        std::vector<uint8_t> filt_flags(rsize.size(), 0);
        tm.begin(stream);
        if (cfg.filt) CR_TRY(filt.run_window(*chain, in + wbase, d_win, n - wbase, roff, rsize, filt_flags, filt_flag));
        tm.mark("filters");
        if (tm.enabled) { CR_CUDA(cudaStreamSynchronize(stream)); tm.finish(); }
        std::vector<BlockIO> blk; size_t dtotal = 0;
        tm.begin(stream);
        CR_TRY(dict_encode_window(d_win, roff, rsize, blk, dtotal));
        tm.mark("diccode");
        if (tm.enabled) { CR_CUDA(cudaStreamSynchronize(stream)); tm.finish(); }
        for (size_t i = 0; i < blk.size(); i++) { blk[i].filt = filt_flags[i]; blk[i].prec = (uint8_t)cfg.prec; }
        const bool last = b1 == nblocks;
        if (!cfg.prec) {
            CR_TRY(chain->encode_blocks(d_D.as<uint8_t>(), blk, 1, last, d_out, out_pos, wrote));
        } else {
            // -p: the dictionary-coded block is the payload (src/main.c:191-195)
            std::vector<HeaderDesc> hdrs; std::vector<CopyDesc> copies; size_t p = out_pos;
            for (size_t i = 0; i < blk.size(); i++) {
                HeaderDesc h; memset(&h, 0, sizeof h); h.dst = p; h.len = 6; memcpy(h.bytes, &blk[i].size, 4); h.bytes[4] = blk[i].filt; h.bytes[5] = 1; hdrs.push_back(h);
                CopyDesc c = { blk[i].off, p + 6, blk[i].size, 0 }; copies.push_back(c);
                p += 6 + blk[i].size;
            }
            wrote = p - out_pos;
            if (d_out.cap < p) {
                DevBuf nb2; CR_TRY(nb2.reserve(p + p / 4));
                if (d_out.p && out_pos) CR_CUDA(cudaMemcpyAsync(nb2.p, d_out.p, out_pos, cudaMemcpyDeviceToDevice, stream));
                CR_CUDA(cudaStreamSynchronize(stream));
                d_out.release(); d_out = nb2;
            }
            CR_TRY(upload(b_hdr, hdrs)); CR_TRY(upload(b_copy, copies));
            CR_LAUNCH(k_write_headers, dim3(cr_div_up(hdrs.size(), 64)), dim3(64), stream, b_hdr.as<HeaderDesc>(), (uint32_t)hdrs.size(), d_out.as<uint8_t>());
            CR_LAUNCH(k_copy_segments, dim3(64, (unsigned)copies.size()), dim3(256), stream, b_copy.as<CopyDesc>(), d_D.as<uint8_t>(), d_D.as<uint8_t>(), d_out.as<uint8_t>());
        }
        out_pos += wrote;
    }
    hlap("data windows");
#ifndef CRGPU_SIM
    if (dict_async) {
        dict_thread.join();
        if (dict_rc != CRGPU_OK) return dict_rc;
        if (mlen + dict_wrote + out_pos > out_cap) return CRGPU_ERR_ARG;
        CR_CUDA(cudaMemcpyAsync(out + mlen, d_dictout.p, dict_wrote, cudaMemcpyDeviceToHost, stream));     // (the helper has synchronised its stream)
    }
#endif
    if (mlen + dict_wrote + out_pos > out_cap) return CRGPU_ERR_ARG;
    CR_CUDA(cudaMemcpyAsync(out + mlen + dict_wrote, d_out.p, out_pos, cudaMemcpyDeviceToHost, stream));
    CR_CUDA(cudaStreamSynchronize(stream));
    hlap("d2h");
    *out_n = mlen + dict_wrote + out_pos;
    return CRGPU_OK;
}
