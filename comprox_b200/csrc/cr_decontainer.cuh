// cr_decontainer.cuh -- whole-container decompression: cr_main's decode branch (src/main.c:220-302).
#pragma once
#include "cr_container.cuh"
#include "cr_decode.cuh"

struct Decompressor {
    LzChain* chain = nullptr;
    cudaStream_t stream = 0;
    DevBuf d_cont, d_D, d_out, d_blocks, d_ctx, d_ddblocks, d_subs, d_totals, d_words, d_lens, d_copy, d_jobs, d_first, d_err;
    DevBuf t_meta, t_items, t_short, t_l8, t_l4, t_l2;
    FilterHost filt;
    uint32_t epoch = 0;

    void release() {
        DevBuf* all[] = { &d_cont, &d_D, &d_out, &d_blocks, &d_ctx, &d_ddblocks, &d_subs, &d_totals, &d_words, &d_lens, &d_copy, &d_jobs, &d_first, &d_err, &t_meta, &t_items, &t_short, &t_l8, &t_l4, &t_l2 };
        for (DevBuf* b : all) b->release();
        filt.release();
    }
    int tables(DecTables& T) {
        if (chain->variant == CR_LZ77) { memset(&T, 0, sizeof T); return CRGPU_OK; }     // LZ77 copies by distance: no matcher state
        if (chain->variant == CR_ROLZ) {
            const bool fresh = t_meta.p == nullptr;
            CR_TRY(t_meta.reserve((size_t)RZ_BUCKETS * 4)); CR_TRY(t_items.reserve((size_t)RZ_BUCKETS * 64 * 4)); CR_TRY(t_short.reserve(256 * 16 * 4));
            if (fresh) { CR_CUDA(cudaMemsetAsync(t_meta.p, 0, (size_t)RZ_BUCKETS * 4, stream)); epoch = 0; }
        } else {
            const bool fresh = t_l8.p == nullptr;
            CR_TRY(t_l8.reserve((size_t)8 << 24)); CR_TRY(t_l4.reserve((size_t)8 << 20)); CR_TRY(t_l2.reserve((size_t)8 << 16));
            if (fresh) {
                CR_CUDA(cudaMemsetAsync(t_l8.p, 0, (size_t)8 << 24, stream)); CR_CUDA(cudaMemsetAsync(t_l4.p, 0, (size_t)8 << 20, stream));
                CR_CUDA(cudaMemsetAsync(t_l2.p, 0, (size_t)8 << 16, stream)); epoch = 0;
            }
        }
        T.rz_meta = t_meta.as<uint32_t>(); T.rz_items = t_items.as<uint32_t>(); T.rz_short = t_short.as<uint32_t>();
        T.lzp8 = t_l8.as<unsigned long long>(); T.lzp4 = t_l4.as<unsigned long long>(); T.lzp2 = t_l2.as<unsigned long long>();
        return CRGPU_OK;
    }
    // lzdecode of `blk` (consecutive blocks of one model chain); fills D at blk[i].d_off.
    // lz_prepare() uploads the descriptors and fills `job`; the caller launches the chain (alone or in a batch); lz_finish() reads the context back.
    DecJob job;
    int lz_prepare(std::vector<DecBlock>& blk) {
        DecTables T; CR_TRY(tables(T));
        std::vector<CopyDesc> copies;
        for (auto& b : blk) {
            if (epoch >= 65000 && t_meta.p) { CR_CUDA(cudaMemsetAsync(t_meta.p, 0, (size_t)RZ_BUCKETS * 4, stream)); epoch = 0; }   // ROLZ tags are 16 bit
            b.epoch = ++epoch;
            if (!b.coded && b.d_size) { CopyDesc c = { b.in_off + (b.in_size - b.d_size), b.d_off, b.d_size, 0 }; copies.push_back(c); }
        }
        CR_TRY(chain->upload(d_blocks, blk));
        std::vector<uint32_t> c(2, 0); c[0] = chain->chain_ctx;            // [context in/out, error code out]
        CR_TRY(chain->upload(d_ctx, c));
        if (!copies.empty()) {
            CR_TRY(chain->upload(d_copy, copies));
            CR_LAUNCH(k_copy_segments, dim3(64, (unsigned)copies.size()), dim3(256), stream, d_copy.as<CopyDesc>(), d_cont.as<uint8_t>(), d_cont.as<uint8_t>(), d_D.as<uint8_t>());
        }
        job.variant = chain->variant; job.nb = (uint32_t)blk.size();
        job.cont = d_cont.as<uint8_t>(); job.blocks = d_blocks.as<DecBlock>();
        job.st = chain->st; job.T = T; job.ctx_io = d_ctx.as<uint32_t>(); job.D = d_D.as<uint8_t>();
        return CRGPU_OK;
    }
    int lz_launch() {
        if (job.nb == 0) return CRGPU_OK;
#ifndef CRGPU_SIM
        if (!chain->scalar_models) CR_LAUNCH(k_lzdecode_warp, dim3(1), dim3(32), stream, job.variant, job.cont, job.blocks, job.nb, job.st, job.T, job.ctx_io, job.D);
        else
#endif
        CR_LAUNCH(k_lzdecode_serial, dim3(1), dim3(1), stream, job.variant, job.cont, job.blocks, job.nb, job.st, job.T, job.ctx_io, job.D);
        return CRGPU_OK;
    }
    int lz_finish() {
        std::vector<uint32_t> c;
        CR_TRY(chain->download(c, d_ctx.p, 2));
        chain->chain_ctx = c[0];
        return c[1] ? CRGPU_ERR_CORRUPT : CRGPU_OK;       // a damaged stream / match (cr_decode.cuh): nothing outside the block was touched
    }
    // one container in three phases; between the phases the decode chain `job` must have run (lz_launch() or a batched launch)
    const uint8_t* c_in = nullptr; uint64_t c_n = 0; uint8_t* c_out = nullptr; uint64_t c_out_cap = 0; uint64_t* c_out_n = nullptr;
    uint64_t c_p = 0;
    std::vector<DecBlock> c_dblk, c_blk; std::vector<uint8_t> c_filt; DdDict c_dic;
    int describe(uint64_t off, uint32_t size, int prec, uint64_t d_off, DecBlock& b);
    int begin(const uint8_t* in, uint64_t n, uint8_t* out, uint64_t out_cap, uint64_t* out_n);   // -> job = dictionary payload
    int middle();                                                                                  // -> job = data blocks (nb may be 0)
    int finish_layout(bool may_resume = false);                                                                         // -> ddjob = the sub-chunks to expand
    int layout(bool may_resume = false);
    int load_words(const char* text);
    int lzdecode_block(const uint8_t* in, uint32_t n, uint8_t* out, uint64_t out_cap, uint32_t* out_n);          // stage shims (crgpu_lzdecode,
    int dict_decode_block(const uint8_t* in, uint32_t n, uint8_t* out, uint64_t out_cap, uint32_t* out_n);       //  crgpu_dictionary_decode)
    int dd_launch();
    int finish_output();                                                                           // inverse filters, copy out
    DdJob ddjob; uint32_t dd_nsub = 0; std::vector<DdBlock> c_ddb; uint64_t c_raw_total = 0;
    const uint8_t* resume_in = nullptr; uint64_t resume_n = 0, resume_tag = 0;     // container whose lz stage is done but whose output did not fit
    static uint64_t tag_of(const uint8_t* p, uint64_t n) {         // guards the resume against a different container at the same address
        uint64_t h = 1469598103934665603ull ^ n;
        for (uint64_t i = 0; i < n && i < 256; i++) h = (h ^ p[i]) * 1099511628211ull;
        for (uint64_t i = n > 256 ? n - 256 : 0; i < n; i++) h = (h ^ p[i]) * 1099511628211ull;
        return h;
    }
    int decompress(const uint8_t* in, uint64_t n, uint8_t* out, uint64_t out_cap, uint64_t* out_n) {
        if (resume_in && resume_in == in && resume_n == n && out_n && resume_tag == tag_of(in, n)) {
            // second call for the same container after "out_cap too small": d_D still holds the decoded blocks, the models are past them
            resume_in = nullptr;
            c_out = out; c_out_cap = out_cap; c_out_n = out_n;
            CR_TRY(layout(true)); CR_TRY(dd_launch());
            return finish_output();
        }
        resume_in = nullptr;
        CR_TRY(begin(in, n, out, out_cap, out_n)); CR_TRY(lz_launch());
        CR_TRY(middle()); CR_TRY(lz_launch());
        CR_TRY(finish_layout(true)); CR_TRY(dd_launch());
        return finish_output();
    }
};

// payload -> DecBlock: sizes come from the inner headers (src/rolzmain/cr-coder.c:63-71, src/ropmain/cr-coder.c:60-66)
inline int Decompressor::describe(uint64_t off, uint32_t size, int prec, uint64_t d_off, DecBlock& b) {
    const int variant = chain->variant;
    const uint32_t hdr = variant == CR_ROLZ ? 16 : variant == CR_LZP ? 20 : 32;
    memset(&b, 0, sizeof b);
    b.in_off = off; b.in_size = size; b.d_off = d_off;
    if (prec) { b.coded = 0; b.d_size = size; return CRGPU_OK; }
    if (size < hdr) return CRGPU_ERR_CORRUPT;
    const int compressed = variant == CR_ROLZ ? c_in[off + 1] : c_in[off];
    if (!compressed) { b.coded = 0; b.d_size = size - hdr; return CRGPU_OK; }
    b.coded = 1; memcpy(&b.d_size, c_in + off + 4, 4);
    // stream offsets of the inner header must lie inside the payload (the reference trusts them: src/rolzmain/cr-coder.c:300-303)
    auto ld = [&](uint32_t at) { uint32_t v; memcpy(&v, c_in + off + at, 4); return v; };
    if (variant == CR_ROLZ) { const uint32_t o1 = ld(12); if (o1 < 16 || o1 > size) return CRGPU_ERR_CORRUPT; }
    if (variant == CR_LZ77) { const uint32_t o1 = ld(20), o2 = ld(24), o3 = ld(28); if (o1 < 32 || o1 > o2 || o2 > o3 || o3 > size) return CRGPU_ERR_CORRUPT; }
    if (variant == CR_LZP && b.d_size < 9 && b.d_size != 0) return CRGPU_ERR_CORRUPT;          // a coded LZP block carries its first nine bytes in the header
    return CRGPU_OK;
}

inline int Decompressor::begin(const uint8_t* in, uint64_t n, uint8_t* out, uint64_t out_cap, uint64_t* out_n) {
    stream = chain->stream;
    resume_in = nullptr;
    const char* magic = cr_magic(chain->variant);
    const size_t mlen = strlen(magic);
    if (!in || !out_n || n < mlen + 4 || memcmp(in, magic, mlen) != 0) return CRGPU_ERR_ARG;     // check_magic, src/main.c:72-79
    c_in = in; c_n = n; c_out = out; c_out_cap = out_cap; c_out_n = out_n;
    CR_TRY(d_cont.reserve(n + 64));
    CR_CUDA(cudaMemcpyAsync(d_cont.p, in, n, cudaMemcpyHostToDevice, stream));
    // ---- static dictionary (src/main.c:244-259)
    uint64_t p = mlen;
    uint32_t dict_len; memcpy(&dict_len, in + p, 4); p += 4;
    if (p + dict_len > n) return CRGPU_ERR_CORRUPT;
    c_dblk.assign(1, DecBlock());
    CR_TRY(describe(p, dict_len, 0, 0, c_dblk[0]));
    c_p = p + dict_len;
    if (c_dblk[0].d_size > (8u << 20)) return CRGPU_ERR_CORRUPT;              // the front-coded dictionary text is a few hundred KB at most
    // the data blocks are parsed here too, so that every device buffer of the call is sized before any chain runs
    c_blk.clear(); c_filt.clear();
    uint64_t dtotal = 0;
    p = c_p;
    while (p + 6 <= n) {
        uint32_t size; memcpy(&size, in + p, 4); const int f = in[p + 4], prec = in[p + 5];
        p += 6;
        if (p + size > n) return CRGPU_ERR_CORRUPT;
        DecBlock b; CR_TRY(describe(p, size, prec, dtotal, b));
        c_blk.push_back(b); c_filt.push_back((uint8_t)f);
        dtotal += ((uint64_t)b.d_size + 15) & ~15ull;
        p += size;
    }
    CR_TRY(d_D.reserve((dtotal > c_dblk[0].d_size ? dtotal : (uint64_t)c_dblk[0].d_size) + 64));
    CR_TRY(chain->reset_models());
    return lz_prepare(c_dblk);
}

// dictionary_load(dicstr, 0) (src/cr-diccode.c:76-104): the word table the expansion kernel indexes
inline int Decompressor::load_words(const char* text) {
    const std::vector<std::string> entries = hd_entries(text);
    if (entries.size() > 25000 + 2) return CRGPU_ERR_CORRUPT;                  // dic[25000][22] in the reference (src/cr-diccode.c:33-35)
    std::vector<char> words(entries.size() * 24 + 24, 0); std::vector<uint8_t> lens(entries.size() + 1, 0);
    for (size_t i = 0; i < entries.size(); i++) {
        if (entries[i].size() > 24) return CRGPU_ERR_CORRUPT;                  // the expansion kernel's word slots are 24 bytes
        lens[i] = (uint8_t)entries[i].size(); memcpy(&words[i * 24], entries[i].data(), entries[i].size());
    }
    CR_TRY(chain->upload(d_words, words)); CR_TRY(chain->upload(d_lens, lens));
    c_dic = DdDict{ d_words.as<char>(), d_lens.as<uint8_t>(), (int32_t)entries.size(), HD_LEVEL1((int)entries.size()) };
    return CRGPU_OK;
}

inline int Decompressor::middle() {
    CR_TRY(lz_finish());
    std::vector<uint8_t> lcp;
    CR_TRY(chain->download(lcp, d_D.p, c_dblk[0].d_size));
    CR_TRY(chain->reset_models());
    const std::string text = hd_lcp_decode(lcp.data(), lcp.size());
    CR_TRY(load_words(text.c_str()));
    // ---- data blocks (src/main.c:263-292)
    job.nb = 0;
    if (!c_blk.empty()) CR_TRY(lz_prepare(c_blk));
    return CRGPU_OK;
}

inline int Decompressor::finish_layout(bool may_resume) {
    if (!c_blk.empty()) CR_TRY(lz_finish());
    return layout(may_resume);
}

// dictionary_decode of the blocks described by c_blk (their dictionary-coded bytes sit in d_D): pair framing -> sub-chunks
inline int Decompressor::layout(bool may_resume) {
    std::vector<DecBlock>& blk = c_blk;
    // ---- dictionary_decode
    const DdDict dic = c_dic;
    const uint64_t out_cap = c_out_cap;
    const uint32_t nb = (uint32_t)blk.size();
    std::vector<DdBlock> ddb(nb);
    uint64_t sub_cap = 2 * nb + 16;
    for (uint32_t b = 0; b < nb; b++) { memset(&ddb[b], 0, sizeof(DdBlock)); ddb[b].d_off = blk[b].d_off; ddb[b].d_size = blk[b].d_size; sub_cap += blk[b].d_size / 8 + 2; }
    if (sub_cap > (1u << 24)) sub_cap = 1u << 24;
    CR_TRY(chain->upload(d_ddblocks, ddb));
    CR_TRY(d_subs.reserve(sub_cap * sizeof(DdSub))); CR_TRY(d_totals.reserve(64));
    CR_LAUNCH(k_dd_layout, dim3(1), dim3(1), stream, d_D.as<uint8_t>(), d_ddblocks.as<DdBlock>(), nb, d_subs.as<DdSub>(), (uint32_t)sub_cap, (uint64_t)0, d_totals.as<uint64_t>());
    std::vector<uint64_t> totals;
    CR_TRY(chain->download(totals, d_totals.p, 3));
    const uint64_t raw_total = totals[0]; const uint32_t nsub = (uint32_t)totals[1];
    if (totals[2] || nsub > sub_cap) return CRGPU_ERR_CORRUPT;
    if (raw_total > out_cap) {
        // the decoded size is only known here, after the (serial, slow) lzdecode chain: report it and keep the dictionary-coded
        // blocks, so that the caller's second call with a large enough buffer resumes at this point (decompress())
        if (c_out_n) *c_out_n = raw_total;
        // (armed only on the whole-container route: the stage calls and the batch route reuse d_D / c_blk for other data)
        if (may_resume) { resume_in = c_in; resume_n = c_n; resume_tag = tag_of(c_in, c_n); }
        return CRGPU_ERR_ARG;
    }
    CR_TRY(chain->download(ddb, d_ddblocks.p, nb));
    CR_TRY(d_out.reserve(raw_total + 256));
    CR_CUDA(cudaMemsetAsync(d_out.p, 0, raw_total + 128, stream));      // (bytes a damaged sub-chunk never reaches stay defined)
    std::vector<CopyDesc> copies;
    for (uint32_t b = 0; b < nb; b++) if (ddb[b].nsub == 0 && ddb[b].raw_size) { CopyDesc c = { ddb[b].d_off, ddb[b].raw_off, ddb[b].raw_size, 0 }; copies.push_back(c); }
    if (!copies.empty()) {
        CR_TRY(chain->upload(d_copy, copies));
        CR_LAUNCH(k_copy_segments, dim3(64, (unsigned)copies.size()), dim3(256), stream, d_copy.as<CopyDesc>(), d_D.as<uint8_t>(), d_D.as<uint8_t>(), d_out.as<uint8_t>());
    }
    c_ddb = ddb; c_raw_total = raw_total;
    CR_TRY(d_err.reserve(16));
    CR_CUDA(cudaMemsetAsync(d_err.p, 0, 4, stream));
    ddjob = DdJob{ d_D.as<uint8_t>(), d_ddblocks.as<DdBlock>(), d_subs.as<DdSub>(), dic, d_out.as<uint8_t>(), d_err.as<uint32_t>() };
    dd_nsub = nsub;
    return CRGPU_OK;
}

inline int Decompressor::dd_launch() {
    if (dd_nsub) CR_LAUNCH(k_dd_subs, dim3(cr_div_up(dd_nsub, 32)), dim3(32), stream, ddjob.D, ddjob.blocks, ddjob.subs, dd_nsub, ddjob.dic, ddjob.out, ddjob.err);
    return CRGPU_OK;
}

inline int Decompressor::finish_output() {
    const std::vector<DdBlock>& ddb = c_ddb;
    const uint64_t raw_total = c_raw_total;
    const std::vector<uint8_t>& filt_flags = c_filt;
    const uint32_t nb = (uint32_t)c_blk.size();
    uint8_t* out = c_out;
    if (dd_nsub) {                                                     // a sub-chunk that left its buffers stopped and said so
        std::vector<uint32_t> e;
        CR_TRY(chain->download(e, d_err.p, 1));
        if (e[0]) return CRGPU_ERR_CORRUPT;
    }

    // ---- inverse filters (src/main.c:284-286): the state machine needs the decoded headers on the host
    bool any_filt = false;
    for (uint8_t f : filt_flags) any_filt |= f != 0;
    if (any_filt) {
        CR_CUDA(cudaMemcpyAsync(out, d_out.p, raw_total, cudaMemcpyDeviceToHost, stream));
        CR_CUDA(cudaStreamSynchronize(stream));
        filt.reset();
        std::vector<uint64_t> roff; std::vector<uint32_t> rsize;
        for (uint32_t b = 0; b < nb; b++) if (filt_flags[b]) { roff.push_back(ddb[b].raw_off); rsize.push_back(ddb[b].raw_size); }
        std::vector<uint8_t> ff(roff.size(), 0); int dummy = 0;
        CR_TRY(filt.run_window(*chain, out, d_out.as<uint8_t>(), raw_total, roff, rsize, ff, dummy, /*decode=*/1));
    }
    CR_CUDA(cudaMemcpyAsync(out, d_out.p, raw_total, cudaMemcpyDeviceToHost, stream));
    CR_CUDA(cudaStreamSynchronize(stream));
    *c_out_n = raw_total;
    return CRGPU_OK;
}

// ------------------------------------------------------------------ stage-level forms (one block per call, like the reference)
// lzdecode(ib, ob) (src/rolzmain/cr-coder.c:287-379, src/ropmain/cr-coder.c:231-292, src/roxmain/cr-coder.c:321-526): models and
// the PPM context carry over from the previous call of this handle until reset_models().
inline int Decompressor::lzdecode_block(const uint8_t* in, uint32_t n, uint8_t* out, uint64_t out_cap, uint32_t* out_n) {
    stream = chain->stream;
    resume_in = nullptr;
    c_in = in; c_n = n;
    CR_TRY(d_cont.reserve((size_t)n + 64));
    if (n) CR_CUDA(cudaMemcpyAsync(d_cont.p, in, n, cudaMemcpyHostToDevice, stream));
    std::vector<DecBlock> blk(1);
    CR_TRY(describe(0, n, 0, 0, blk[0]));
    if (blk[0].d_size > out_cap) return CRGPU_ERR_ARG;
    CR_TRY(d_D.reserve((size_t)blk[0].d_size + 64));
    CR_TRY(lz_prepare(blk)); CR_TRY(lz_launch()); CR_TRY(lz_finish());
    if (blk[0].d_size) CR_CUDA(cudaMemcpyAsync(out, d_D.p, blk[0].d_size, cudaMemcpyDeviceToHost, stream));
    CR_CUDA(cudaStreamSynchronize(stream));
    *out_n = blk[0].d_size;
    return CRGPU_OK;
}

// dictionary_decode(ib, ob, NULL) (src/cr-diccode.c:223-283) of one block; needs load_words() first.
inline int Decompressor::dict_decode_block(const uint8_t* in, uint32_t n, uint8_t* out, uint64_t out_cap, uint32_t* out_n) {
    stream = chain->stream;
    resume_in = nullptr; c_in = nullptr; c_n = 0;
    if (n == 0 || !d_words.p) return CRGPU_ERR_ARG;          // the reference reads ib->m_data[size - 1]: an empty block is not a valid input
    CR_TRY(d_D.reserve((size_t)n + 64));
    CR_CUDA(cudaMemcpyAsync(d_D.p, in, n, cudaMemcpyHostToDevice, stream));
    c_blk.assign(1, DecBlock()); memset(&c_blk[0], 0, sizeof(DecBlock)); c_blk[0].d_size = n;
    c_filt.assign(1, 0);
    uint64_t raw_n = 0;
    c_out = out; c_out_cap = out_cap; c_out_n = &raw_n;
    CR_TRY(layout()); CR_TRY(dd_launch()); CR_TRY(finish_output());
    *out_n = (uint32_t)raw_n;
    return CRGPU_OK;
}
