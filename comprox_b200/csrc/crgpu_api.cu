// crgpu_api.cu -- the C ABI declared in include/crgpu.h.
#include "../../include/crgpu.h"
#include "cr_lzchain.cuh"
#include "cr_container.cuh"
#include "cr_decontainer.cuh"
#include <string>
#include <atomic>
#include <thread>

#ifdef CRGPU_SIM
thread_local crsim_dim3 threadIdx, blockIdx, blockDim, gridDim;
#define CR_SET_DEVICE(h) do { } while (0)
#else
unsigned long long g_cr_launches = 0;
// Handles that share a GPU work on private streams, and a chain's serial range walk is one kernel that runs for tens to hundreds of
// milliseconds.  With the default of 8 hardware work queues, streams alias onto the same queue and wait behind each other's long kernels
// (measured: 24 shards, 8 handles: 3.8 s with 8 queues, 2.0 s with 32; profiles/round2_summary.md section 3).  The variable only counts before the
// CUDA context exists, so it is set when the library is loaded -- unless the caller has chosen a value.
__attribute__((constructor)) static void cr_more_work_queues() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }
thread_local cudaStream_t g_cr_alloc_stream = 0;
thread_local bool g_cr_alloc_async = false;
// every entry point that touches the device: select the handle's GPU and bind DevBuf growth (cr_common.cuh) to the handle's stream
#define CR_SET_DEVICE(h) do { CR_CUDA(cudaSetDevice((h)->device)); g_cr_alloc_stream = (h)->stream; g_cr_alloc_async = true; } while (0)
#endif

struct crgpu_handle {
    int device = 0;
    int variant = CRGPU_ROLZ;
    cudaStream_t stream = 0;
    LzChain chain;
    DevBuf d_in, d_out;
    Compressor comp;
    Decompressor decomp;
    FilterHost stage_filt;          // continuation state of crgpu_filter_inplace (the reference's function-local statics)
    bool owns_stream = false;
    uint32_t last_cut_blocks = 0;   // mid-chain blocks stored raw by the most recent crgpu_compress (crgpu_get_stat)
};

extern "C" const char* crgpu_strerror(int code) {
    switch (code) {
        case CRGPU_OK: return "ok";
        case CRGPU_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU path)";
        case CRGPU_ERR_CUDA: return "CUDA runtime error";
        case CRGPU_ERR_ARG: return "bad argument";
        case CRGPU_ERR_VOCAB_OVERFLOW: return "more than 325000 distinct words (reference prune is order dependent)";
        case CRGPU_ERR_HASH_COLLISION: return "word hash collision";
        case CRGPU_ERR_MIDCHAIN_ABORT: return "a non-final block could not be compressed (reference desyncs here too)";
        case CRGPU_ERR_UNSUPPORTED: return "unsupported option";
        case CRGPU_ERR_OOM: return "out of device memory";
        case CRGPU_ERR_CORRUPT: return "corrupt container";
    }
    return "unknown error";
}

extern "C" int crgpu_create(crgpu_handle** out, int variant, int device, void* stream) {
    if (!out || (variant != CRGPU_ROLZ && variant != CRGPU_LZP && variant != CRGPU_LZ77)) return CRGPU_ERR_ARG;
#ifndef CRGPU_SIM
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        fprintf(stderr, "crgpu: no usable CUDA device %d (found %d); there is no CPU fallback\n", device, ndev);
        return CRGPU_ERR_NO_DEVICE;
    }
    CR_CUDA(cudaSetDevice(device));
#endif
    crgpu_handle* h = new crgpu_handle();
    h->device = device; h->variant = variant; h->stream = (cudaStream_t)stream;
#ifndef CRGPU_SIM
    if (stream == CRGPU_OWN_STREAM) {            // private stream: several handles can then work side by side on one GPU
        cudaStream_t own;
        if (cudaStreamCreateWithFlags(&own, cudaStreamNonBlocking) != cudaSuccess) { delete h; return CRGPU_ERR_CUDA; }
        h->stream = own; h->owns_stream = true;
    }
#else
    if (stream == CRGPU_OWN_STREAM) h->stream = 0;
#endif
#ifndef CRGPU_SIM
    {   // keep freed blocks in the stream-ordered pool instead of returning them to the driver at every synchronisation
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) { uint64_t keep = ~0ull; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep); }
        g_cr_alloc_stream = h->stream; g_cr_alloc_async = true;
    }
#endif
    int rc = h->chain.init(variant, h->stream);
    if (rc != CRGPU_OK) { delete h; return rc; }
    *out = h;
    return CRGPU_OK;
}

extern "C" void crgpu_destroy(crgpu_handle* h) {
    if (!h) return;
#ifndef CRGPU_SIM
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
#endif
    h->chain.release(); h->d_in.release(); h->d_out.release(); h->comp.release(); h->decomp.release(); h->stage_filt.release();
#ifndef CRGPU_SIM
    if (h->owns_stream) cudaStreamDestroy(h->stream);
#endif
    delete h;
}

extern "C" int crgpu_reset_models(crgpu_handle* h) {
    if (!h) return CRGPU_ERR_ARG;
    CR_SET_DEVICE(h);
    return h->chain.reset_models();
}

extern "C" int crgpu_lzencode(crgpu_handle* h, const uint8_t* in, const uint32_t* sizes, uint32_t nblocks, int chain_ends,
                              uint8_t* out, uint64_t out_cap, uint32_t* out_sizes) {
    if (!h || !sizes || !out_sizes || (nblocks && !in)) return CRGPU_ERR_ARG;
#ifndef CRGPU_SIM
    CR_SET_DEVICE(h);
#endif
    // every payload is at most header + block bytes (stored form): demand that much up front, so that a too small `out` is
    // reported BEFORE any block has advanced the models of the chain
    { uint64_t need = 0; for (uint32_t b = 0; b < nblocks; b++) need += (uint64_t)sizes[b] + 32; if (nblocks == 0) need = 32; if (out_cap < need) return CRGPU_ERR_ARG; }
    // windows of at most RZ_MAX_BLOCKS consecutive blocks; model state carries from window to window
    size_t src = 0, written = 0;
    for (uint32_t b0 = 0; b0 < nblocks || (nblocks == 0 && b0 == 0); b0 += RZ_MAX_BLOCKS) {
        const uint32_t b1 = b0 + RZ_MAX_BLOCKS < nblocks ? b0 + RZ_MAX_BLOCKS : nblocks;
        std::vector<BlockIO> blk(b1 - b0);
        size_t total = 0;
        for (uint32_t b = b0; b < b1; b++) {
            memset(&blk[b - b0], 0, sizeof(BlockIO));
            blk[b - b0].off = total; blk[b - b0].size = sizes[b];
            total += ((size_t)sizes[b] + 15) & ~(size_t)15;      // keep blocks 16-byte aligned in the window
        }
        CR_TRY(h->d_in.reserve(total + 64));
        for (uint32_t b = b0; b < b1; b++) {
            CR_CUDA(cudaMemcpyAsync(h->d_in.as<uint8_t>() + blk[b - b0].off, in + src, sizes[b], cudaMemcpyHostToDevice, h->stream));
            src += sizes[b];
        }
        size_t out_total = 0;
        CR_TRY(h->chain.encode_blocks(h->d_in.as<uint8_t>(), blk, 2, chain_ends != 0 && b1 == nblocks, h->d_out, 0, out_total));
        if (written + out_total > out_cap) return CRGPU_ERR_ARG;
        CR_CUDA(cudaMemcpyAsync(out + written, h->d_out.p, out_total, cudaMemcpyDeviceToHost, h->stream));
        CR_CUDA(cudaStreamSynchronize(h->stream));
        written += out_total;
        for (uint32_t b = b0; b < b1; b++) out_sizes[b] = blk[b - b0].out_size;
        if (nblocks == 0) break;
    }
    return CRGPU_OK;
}

extern "C" int64_t crgpu_debug_fetch(crgpu_handle* h, const char* what, void* dst, uint64_t cap) {
    if (!h || !what) return CRGPU_ERR_ARG;
    LzChain& c = h->chain;
    std::string w(what);
    const void* src = nullptr; uint64_t bytes = 0;
    if (w == "span") { src = c.b_span.p; bytes = c.last_dtotal; }
    else if (w == "tidx") { src = c.b_tidx.p; bytes = c.last_dtotal; }
    else if (w == "ev_ctx") { src = c.b_evctx.p; bytes = (uint64_t)c.last_nev * 4; }
    else if (w == "ev_sym") { src = c.b_evsym.p; bytes = c.last_nev; }
    else if (w == "pred") { src = c.b_pred.p; bytes = c.last_nev; }
    else if (w == "dense") { src = c.b_dense.p; bytes = ((uint64_t)c.last_nev + c.last_nesc) * sizeof(Tri); }
    else if (w == "dense_side") { src = c.b_denseside.p; bytes = (uint64_t)c.last_nside * sizeof(Tri); }
    else return CRGPU_ERR_ARG;
    uint64_t n = bytes < cap ? bytes : cap;
    if (n && dst) {
        if (cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) return CRGPU_ERR_CUDA;
        if (cudaStreamSynchronize(h->stream) != cudaSuccess) return CRGPU_ERR_CUDA;
    }
    return (int64_t)bytes;
}

extern "C" int crgpu_debug_sort(crgpu_handle* h, const void* keys, const uint32_t* vals, uint64_t n, int key_bytes, int begin_bit, int end_bit,
                                void* keys_out, uint32_t* vals_out) {
    if (!h || !vals || !vals_out || (key_bytes != 0 && key_bytes != 4 && key_bytes != 8) || (key_bytes && (!keys || !keys_out))) return CRGPU_ERR_ARG;
    if (n == 0) return CRGPU_OK;
    LzChain& c = h->chain;
    DevBuf ki, ko, vi, vo;
    int rc = CRGPU_OK;
    auto run = [&]() -> int {
        CR_TRY(vi.reserve(n * 4 + 16)); CR_TRY(vo.reserve(n * 4 + 16));
        CR_CUDA(cudaMemcpyAsync(vi.p, vals, n * 4, cudaMemcpyHostToDevice, h->stream));
        if (key_bytes == 0) {
            CR_TRY(cr_exclusive_sum(c.prims, vi.as<uint32_t>(), vo.as<uint32_t>(), n));
        } else {
            CR_TRY(ki.reserve(n * key_bytes + 16)); CR_TRY(ko.reserve(n * key_bytes + 16));
            CR_CUDA(cudaMemcpyAsync(ki.p, keys, n * key_bytes, cudaMemcpyHostToDevice, h->stream));
            if (key_bytes == 4) CR_TRY(cr_sort_pairs<uint32_t>(c.prims, ki.as<uint32_t>(), ko.as<uint32_t>(), vi.as<uint32_t>(), vo.as<uint32_t>(), n, begin_bit, end_bit));
            else CR_TRY(cr_sort_pairs<uint64_t>(c.prims, ki.as<uint64_t>(), ko.as<uint64_t>(), vi.as<uint32_t>(), vo.as<uint32_t>(), n, begin_bit, end_bit));
            CR_CUDA(cudaMemcpyAsync(keys_out, ko.p, n * key_bytes, cudaMemcpyDeviceToHost, h->stream));
        }
        CR_CUDA(cudaMemcpyAsync(vals_out, vo.p, n * 4, cudaMemcpyDeviceToHost, h->stream));
        CR_CUDA(cudaStreamSynchronize(h->stream));
        return CRGPU_OK;
    };
    rc = run();
    ki.release(); ko.release(); vi.release(); vo.release();
    return rc;
}

extern "C" int crgpu_profile(crgpu_handle* h, int enable) {
    if (!h) return CRGPU_ERR_ARG;
    h->chain.timer.enabled = enable != 0;
    h->chain.timer.result.clear();
    return CRGPU_OK;
}

extern "C" int crgpu_profile_report(crgpu_handle* h, char* buf, uint64_t cap) {
    if (!h || !buf || cap == 0) return CRGPU_ERR_ARG;
    std::string s;
    char line[128];
    for (auto& r : h->chain.timer.result) { snprintf(line, sizeof line, "%s %.3f\n", r.first.c_str(), r.second); s += line; }
    size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
    memcpy(buf, s.data(), n); buf[n] = 0;
    return (int)n;
}

extern "C" int crgpu_set_option(crgpu_handle* h, const char* name, int64_t value) {
    if (!h || !name) return CRGPU_ERR_ARG;
    std::string n(name);
    if (n == "scalar_models") { h->chain.scalar_models = value != 0; return CRGPU_OK; }
    if (n == "dc_listed") { h->comp.dc_listed = value != 0; return CRGPU_OK; }
    if (n == "dp_tiles") { h->comp.dp_tiles = value != 0; return CRGPU_OK; }
    if (n == "dict_mode") { if (value < 0 || value > 1) return CRGPU_ERR_ARG; h->comp.dict_mode = (int)value; return CRGPU_OK; }
    if (n == "rolz_match_variant") { if (value < 1 || value > 2) return CRGPU_ERR_ARG; h->chain.rolz_match_variant = (int)value; return CRGPU_OK; }
    if (n == "o1_hot_variant") { if (value < 1 || value > 2) return CRGPU_ERR_ARG; h->chain.o1_hot_variant = (int)value; return CRGPU_OK; }
    if (n == "o2_hot_variant") { if (value < 1 || value > 3) return CRGPU_ERR_ARG; h->chain.o2_hot_variant = (int)value; return CRGPU_OK; }
    if (n == "o2_width") { if (value != 0 && value != 256 && value != 512 && value != 1024) return CRGPU_ERR_ARG; h->chain.o2_width = (uint32_t)value; return CRGPU_OK; }
    if (n == "o2_rec_cap_test") { h->chain.o2_rec_cap_test = (uint32_t)value; return CRGPU_OK; }     // tests: force the out-of-records path
    if (n == "rc_variant") { if (value < 1 || value > 8) return CRGPU_ERR_ARG; h->chain.rc_variant = (int)value; return CRGPU_OK; }
#ifndef CRGPU_SIM
    if (n == "rc_job_symbols") { if (value != 0 && (value < 4096 || value > (1 << 24))) return CRGPU_ERR_ARG; h->chain.rcpar.job_symbols = (uint32_t)value; return CRGPU_OK; }
    if (n == "rc_late_cfg") { h->chain.rcpar.late_cfg = (int)value; return CRGPU_OK; }
    if (n == "rc_serial") { if (value < -1 || value > 1) return CRGPU_ERR_ARG; h->chain.rcpar.serial_only = (int)value; return CRGPU_OK; }
#else
    if (n == "rc_job_symbols" || n == "rc_late_cfg" || n == "rc_serial") return CRGPU_OK;
#endif
    if (n == "hot_contexts") { h->chain.hot_contexts = value != 0; return CRGPU_OK; }
    if (n == "match_limit") { if (value < 1 || value > 1000000) return CRGPU_ERR_ARG; h->chain.match_limit = (uint32_t)value; return CRGPU_OK; }   // comprox -m
    if (n == "exact_aborts") { h->chain.exact_aborts = value != 0; return CRGPU_OK; }             // 0: mid-chain "cannot compress" -> CRGPU_ERR_MIDCHAIN_ABORT
    if (n == "flexible") { h->chain.flexible = value != 0; return CRGPU_OK; }                    // -f: the reference's global flexible_parsing
    if (n == "lz77_max_iter") { if (value < 0) return CRGPU_ERR_ARG; h->chain.x_max_iter = (uint32_t)value; return CRGPU_OK; }
    return CRGPU_ERR_ARG;
}

extern "C" uint64_t crgpu_compress_bound(uint64_t n, uint32_t block_size) { return cr_compress_bound(n, block_size); }

extern "C" int crgpu_compress(crgpu_handle* h, const crgpu_config* cfg, const uint8_t* in, uint64_t n, uint8_t* out, uint64_t out_cap, uint64_t* out_n) {
    if (!h || !cfg) return CRGPU_ERR_ARG;
#ifndef CRGPU_SIM
    CR_SET_DEVICE(h);
#endif
    CrConfig c; c.block_size = cfg->block_size; c.filt = cfg->filt; c.prec = cfg->prec; c.flexible = cfg->flexible; c.window_bytes = cfg->window_bytes;
    h->comp.chain = &h->chain;
    const uint32_t cuts0 = h->chain.cut_blocks;
    const int rc = h->comp.compress(c, in, n, out, out_cap, out_n);
    h->last_cut_blocks = h->chain.cut_blocks - cuts0;
    return rc;
}

extern "C" int64_t crgpu_get_stat(crgpu_handle* h, const char* name) {
    if (!h || !name) return CRGPU_ERR_ARG;
    const std::string n(name);
    if (n == "cut_blocks") return h->chain.cut_blocks;
    if (n == "last_cut_blocks") return h->last_cut_blocks;
    return CRGPU_ERR_ARG;
}

extern "C" uint64_t crgpu_launch_count(void) {
#ifdef CRGPU_SIM
    return 0;
#else
    return g_cr_launches;
#endif
}

// Stages the input in HBM ahead of crgpu_compress (for timing the device-resident path): a following
// crgpu_compress call with the same `in` pointer and length skips its host-to-device copy.
extern "C" int crgpu_stage_input(crgpu_handle* h, const uint8_t* in, uint64_t n) {
    if (!h || (n && !in)) return CRGPU_ERR_ARG;
#ifndef CRGPU_SIM
    CR_SET_DEVICE(h);
#endif
    h->comp.chain = &h->chain; h->comp.stream = h->chain.stream;
    CR_TRY(h->comp.stage(in, n));
    h->comp.staged_ptr = in; h->comp.staged_n = n;       // one-shot: the next crgpu_compress consumes it, every other staging call clears it
    return CRGPU_OK;
}

extern "C" int crgpu_decompress(crgpu_handle* h, const uint8_t* in, uint64_t n, uint8_t* out, uint64_t out_cap, uint64_t* out_n) {
    if (!h) return CRGPU_ERR_ARG;
#ifndef CRGPU_SIM
    CR_SET_DEVICE(h);
#endif
    h->decomp.chain = &h->chain;
    return h->decomp.decompress(in, n, out, out_cap, out_n);
}

extern "C" int crgpu_decompress_batch(crgpu_handle* const* hs, uint32_t count, const uint8_t* const* ins, const uint64_t* in_lens,
                                      uint8_t* const* outs, const uint64_t* out_caps, uint64_t* out_lens) {
    if (!hs || !ins || !in_lens || !outs || !out_caps || !out_lens) return CRGPU_ERR_ARG;
    if (count == 0) return CRGPU_OK;
    for (uint32_t i = 0; i < count; i++) {
        if (!hs[i] || hs[i]->device != hs[0]->device || hs[i]->variant != hs[0]->variant || hs[i]->stream != hs[0]->stream) return CRGPU_ERR_ARG;
        for (uint32_t k = 0; k < i; k++) if (hs[k] == hs[i]) return CRGPU_ERR_ARG;
        hs[i]->decomp.chain = &hs[i]->chain;
    }
    // one container after the other (the simulation, and the scalar kernels on the GPU); a damaged container fails alone
    auto one_by_one = [&]() -> int {
        int first_err = CRGPU_OK;
        for (uint32_t i = 0; i < count; i++) {
            out_lens[i] = 0;
            const int rc = hs[i]->decomp.decompress(ins[i], in_lens[i], outs[i], out_caps[i], &out_lens[i]);
            if (rc != CRGPU_OK && rc != CRGPU_ERR_CORRUPT && rc != CRGPU_ERR_ARG) return rc;
            if (rc != CRGPU_OK) { out_lens[i] = 0; if (first_err == CRGPU_OK) first_err = rc; }
        }
        return first_err;
    };
#ifdef CRGPU_SIM
    return one_by_one();
#else
    CR_SET_DEVICE(hs[0]);
    if (hs[0]->chain.scalar_models) return one_by_one();
    cudaStream_t stream = hs[0]->stream;
    DevBuf& d_jobs = hs[0]->decomp.d_jobs;
    std::vector<DecJob> jobs(count);
    // a damaged container is given up on its own: the others are decoded, its out_lens[i] is 0 and the call returns the first error
    std::vector<int> st(count, CRGPU_OK);
    auto launch = [&]() -> int {
        for (uint32_t i = 0; i < count; i++) { jobs[i] = hs[i]->decomp.job; if (st[i] != CRGPU_OK) jobs[i].nb = 0; }
        CR_TRY(hs[0]->chain.upload(d_jobs, jobs));
        CR_LAUNCH(k_lzdecode_jobs, dim3(count), dim3(32), stream, d_jobs.as<DecJob>(), count);
        return CRGPU_OK;
    };
    auto hard = [](int rc) { return rc != CRGPU_OK && rc != CRGPU_ERR_CORRUPT && rc != CRGPU_ERR_ARG; };     // CUDA / memory errors end the call
    for (uint32_t i = 0; i < count; i++) { out_lens[i] = 0; st[i] = hs[i]->decomp.begin(ins[i], in_lens[i], outs[i], out_caps[i], &out_lens[i]); if (hard(st[i])) return st[i]; }
    CR_TRY(launch());
    for (uint32_t i = 0; i < count; i++) if (st[i] == CRGPU_OK) { st[i] = hs[i]->decomp.middle(); if (hard(st[i])) return st[i]; }
    CR_TRY(launch());
    // dictionary_decode: the sub-chunks of all containers in one launch (each is a serial back-to-front expansion)
    std::vector<DdJob> ddjobs(count); std::vector<uint32_t> first(count + 1, 0);
    for (uint32_t i = 0; i < count; i++) {
        if (st[i] == CRGPU_OK) { st[i] = hs[i]->decomp.finish_layout(); if (hard(st[i])) return st[i]; }
        const bool ok = st[i] == CRGPU_OK;
        if (ok) ddjobs[i] = hs[i]->decomp.ddjob; else memset(&ddjobs[i], 0, sizeof(DdJob));
        first[i + 1] = first[i] + (ok ? hs[i]->decomp.dd_nsub : 0);
    }
    if (first[count]) {
        DevBuf& d_first = hs[0]->decomp.d_first;
        CR_TRY(hs[0]->chain.upload(d_jobs, ddjobs)); CR_TRY(hs[0]->chain.upload(d_first, first));
        CR_LAUNCH(k_dd_subs_jobs, dim3(cr_div_up(first[count], 32)), dim3(32), stream, d_jobs.as<DdJob>(), d_first.as<uint32_t>(), count, first[count]);
    }
    for (uint32_t i = 0; i < count; i++) if (st[i] == CRGPU_OK) { st[i] = hs[i]->decomp.finish_output(); if (hard(st[i])) return st[i]; }
    for (uint32_t i = 0; i < count; i++) if (st[i] != CRGPU_OK) { out_lens[i] = 0; return st[i]; }
    return CRGPU_OK;
#endif
}

// dicpick(fp, dic_block): the dictionary text (NUL terminated) the reference builds from the whole input.
extern "C" int crgpu_dicpick(crgpu_handle* h, const uint8_t* in, uint64_t n, uint8_t* out, uint64_t out_cap, uint64_t* out_n) {
    if (!h || (n && !in) || !out || !out_n) return CRGPU_ERR_ARG;
    if (n > (1ull << 32)) return CRGPU_ERR_UNSUPPORTED;                        // word positions of the table are 32 bit
#ifndef CRGPU_SIM
    CR_SET_DEVICE(h);
#endif
    h->comp.chain = &h->chain; h->comp.stream = h->chain.stream;
    CR_TRY(h->comp.stage(in, n));
    std::string text;
    CR_TRY(h->comp.dicpick(in, h->comp.d_raw.as<uint8_t>(), n, text));
    if (text.size() > out_cap) return CRGPU_ERR_ARG;
    memcpy(out, text.data(), text.size());
    *out_n = text.size();
    return CRGPU_OK;
}

// ------------------------------------------------------------------ stage-level entry points: one block per call, with the
// argument meaning of the reference's cr-* functions.  host/cr_shim.c wraps them into the reference's
// own signatures so that the UNMODIFIED src/main.c links against this library.

// filter_inplace(buf, len, en_de) -- src/cr-filter.c:33-73
extern "C" int crgpu_filter_inplace(crgpu_handle* h, uint8_t* buf, uint32_t len, int en_de) {
    if (!h || (len && !buf) || (en_de != 0 && en_de != 1)) return CRGPU_ERR_ARG;
    if (len == 0) return 0;
    CR_SET_DEVICE(h);
    cudaStream_t stream = h->chain.stream;
    CR_TRY(h->d_in.reserve((size_t)len + 256));
    CR_CUDA(cudaMemcpyAsync(h->d_in.p, buf, len, cudaMemcpyHostToDevice, stream));
    CR_CUDA(cudaMemsetAsync(h->d_in.as<uint8_t>() + len, 0, 128, stream));
    std::vector<uint64_t> roff(1, 0); std::vector<uint32_t> rsize(1, len); std::vector<uint8_t> flags(1, 0);
    int fired = 0;
    CR_TRY(h->stage_filt.run_window(h->chain, buf, h->d_in.as<uint8_t>(), len, roff, rsize, flags, fired, en_de));
    if (flags[0]) CR_CUDA(cudaMemcpyAsync(buf, h->d_in.p, len, cudaMemcpyDeviceToHost, stream));
    CR_CUDA(cudaStreamSynchronize(stream));
    return flags[0] ? 1 : 0;
}

// dic_lcp_encode / dic_lcp_decode -- src/cr-dicpick.c:261-346.  Host only: the dictionary text is at most a few hundred KB.
extern "C" int crgpu_dic_lcp_encode(const uint8_t* text, uint64_t n, uint8_t* out, uint64_t out_cap, uint64_t* out_n) {
    if (!text || !out || !out_n || n == 0 || text[n - 1] != 0) return CRGPU_ERR_ARG;        // NUL-terminated, as dicpick leaves it
    // dicpick's text is a list of '\n'-terminated lines (at least the two fixed ones, src/cr-dicpick.c:238-240) with no NUL inside;
    // the front coder scans for '\n' without a bound, so anything else is rejected here instead of being read past its end
    if (n < 2 || text[n - 2] != '\n' || memchr(text, 0, n - 1) != nullptr) return CRGPU_ERR_ARG;
    const std::vector<uint8_t> enc = hd_lcp_encode(std::string((const char*)text, n - 1));
    if (enc.size() > out_cap) return CRGPU_ERR_ARG;
    memcpy(out, enc.data(), enc.size());
    *out_n = enc.size();
    return CRGPU_OK;
}
extern "C" int crgpu_dic_lcp_decode(const uint8_t* in, uint64_t n, uint8_t* out, uint64_t out_cap, uint64_t* out_n) {
    if ((n && !in) || !out || !out_n) return CRGPU_ERR_ARG;
    const std::string text = hd_lcp_decode(in, n);
    if (text.size() + 1 > out_cap) return CRGPU_ERR_ARG;
    memcpy(out, text.c_str(), text.size() + 1);
    *out_n = text.size() + 1;
    return CRGPU_OK;
}

// dictionary_load(dicstr, init_trie) -- src/cr-diccode.c:76-120.  Returns the number of dictionary entries (>= 0) or an error.
extern "C" int crgpu_dictionary_load(crgpu_handle* h, const char* dicstr, int init_trie) {
    if (!h || !dicstr) return CRGPU_ERR_ARG;
    CR_SET_DEVICE(h);
    h->comp.chain = &h->chain; h->comp.stream = h->chain.stream;
    h->decomp.chain = &h->chain; h->decomp.stream = h->chain.stream;
    if (init_trie) CR_TRY(h->comp.load_dictionary(std::string(dicstr)));
    CR_TRY(h->decomp.load_words(dicstr));
    return h->decomp.c_dic.nentries;
}

// dictionary_encode(ib, ob) -- src/cr-diccode.c:142-221
extern "C" int crgpu_dictionary_encode(crgpu_handle* h, const uint8_t* in, uint32_t n, uint8_t* out, uint64_t out_cap, uint32_t* out_n) {
    if (!h || (n && !in) || !out || !out_n) return CRGPU_ERR_ARG;
    if (!h->comp.d_trie_edge.p) return CRGPU_ERR_ARG;                                         // crgpu_dictionary_load(.., 1) first
    CR_SET_DEVICE(h);
    h->comp.chain = &h->chain; h->comp.stream = h->chain.stream;
    CR_TRY(h->comp.stage(in, n));
    std::vector<uint64_t> roff(1, 0); std::vector<uint32_t> rsize(1, n); std::vector<BlockIO> blk; size_t dtotal = 0;
    CR_TRY(h->comp.dict_encode_window(h->comp.d_raw.as<uint8_t>(), roff, rsize, blk, dtotal));
    if (blk[0].size > out_cap) return CRGPU_ERR_ARG;
    CR_CUDA(cudaMemcpyAsync(out, h->comp.d_D.as<uint8_t>() + blk[0].off, blk[0].size, cudaMemcpyDeviceToHost, h->chain.stream));
    CR_CUDA(cudaStreamSynchronize(h->chain.stream));
    *out_n = blk[0].size;
    return CRGPU_OK;
}

// dictionary_decode(ib, ob, NULL) -- src/cr-diccode.c:223-283
extern "C" int crgpu_dictionary_decode(crgpu_handle* h, const uint8_t* in, uint32_t n, uint8_t* out, uint64_t out_cap, uint32_t* out_n) {
    if (!h || !in || !out_n || (out_cap && !out)) return CRGPU_ERR_ARG;
    CR_SET_DEVICE(h);
    h->decomp.chain = &h->chain;
    return h->decomp.dict_decode_block(in, n, out, out_cap, out_n);
}

// lzdecode(ib, ob, print) -- src/main.c:59
extern "C" int crgpu_lzdecode(crgpu_handle* h, const uint8_t* in, uint32_t n, uint8_t* out, uint64_t out_cap, uint32_t* out_n) {
    if (!h || !in || !out_n || (out_cap && !out)) return CRGPU_ERR_ARG;
    CR_SET_DEVICE(h);
    h->decomp.chain = &h->chain;
    return h->decomp.lzdecode_block(in, n, out, out_cap, out_n);
}

// size of the block an lzencode payload decodes to (the inner headers, src/rolzmain/cr-coder.c:63-71 etc.), so callers can size `out`
extern "C" int64_t crgpu_lzdecode_size(int variant, const uint8_t* in, uint32_t n) {
    const uint32_t hdr = variant == CRGPU_ROLZ ? 16 : variant == CRGPU_LZP ? 20 : variant == CRGPU_LZ77 ? 32 : 0;
    if (!hdr || !in || n < hdr) return CRGPU_ERR_ARG;
    const int compressed = variant == CRGPU_ROLZ ? in[1] : in[0];
    if (!compressed) return (int64_t)n - hdr;
    uint32_t sz; memcpy(&sz, in + 4, 4);
    return sz;
}

// ------------------------------------------------------------------ shard mode
// Independent containers compressed side by side: one host thread per handle inside the call, containers dealt to whichever handle is
// idle.  The handles may sit on different devices (then this is the multi-GPU form: no data-path collective, every container goes
// straight from its GPU to its host buffer) or share one device on private streams (the serial range chains of different containers
// then overlap on different SMs).  Output bytes do not depend on which handle compressed a container.
extern "C" int crgpu_compress_batch(crgpu_handle* const* hs, uint32_t nhandles, const crgpu_config* cfg, uint32_t count,
                                    const uint8_t* const* ins, const uint64_t* in_lens, uint8_t* const* outs, const uint64_t* out_caps, uint64_t* out_lens) {
    if (!hs || nhandles == 0 || !cfg || (count && (!ins || !in_lens || !outs || !out_caps || !out_lens))) return CRGPU_ERR_ARG;
    for (uint32_t i = 0; i < nhandles; i++) {
        if (!hs[i] || hs[i]->variant != hs[0]->variant) return CRGPU_ERR_ARG;
        for (uint32_t k = 0; k < i; k++) {
            if (hs[k] == hs[i]) return CRGPU_ERR_ARG;
#ifndef CRGPU_SIM
            if (hs[k]->device == hs[i]->device && hs[k]->stream == hs[i]->stream) return CRGPU_ERR_ARG;     // would serialise: use CRGPU_OWN_STREAM
#endif
        }
    }
#ifndef CRGPU_SIM
    // three or more handles on one device: the serial range walks of different containers overlap each other and everything else, while
    // the tracking kernels of the cut chain (cr_rcpar.cuh) fill the device and do not -- unless the caller chose with "rc_serial"
    for (uint32_t i = 0; i < nhandles; i++) {
        uint32_t same = 0;
        for (uint32_t k = 0; k < nhandles; k++) same += hs[k]->device == hs[i]->device;
        hs[i]->chain.rcpar.crowded = same >= 3 && count >= 3;
        hs[i]->comp.crowded = hs[i]->chain.rcpar.crowded;
    }
#endif
    std::atomic<uint32_t> next(0);
    std::atomic<int> first_error(CRGPU_OK);
    auto worker = [&](uint32_t j) {
        for (;;) {
            const uint32_t i = next.fetch_add(1);
            if (i >= count || first_error.load() != CRGPU_OK) return;
            const int rc = crgpu_compress(hs[j], cfg, ins[i], in_lens[i], outs[i], out_caps[i], &out_lens[i]);
            if (rc != CRGPU_OK) { int expect = CRGPU_OK; first_error.compare_exchange_strong(expect, rc); return; }
        }
    };
    const uint32_t nthreads = nhandles < count ? nhandles : count;
    std::vector<std::thread> pool;
    for (uint32_t j = 1; j < nthreads; j++) pool.emplace_back(worker, j);
    if (nthreads) worker(0);
    for (auto& t : pool) t.join();
#ifndef CRGPU_SIM
    for (uint32_t i = 0; i < nhandles; i++) { hs[i]->chain.rcpar.crowded = false; hs[i]->comp.crowded = false; }
#endif
    return first_error.load();
}

// Test aid for k_range_chain<7>: walks the double-precision form of the range recurrence (rc_dp_record / rc_dp_step, cr_rc.cuh -- the
// very functions the kernel calls) over n symbols ON THE HOST, from the coder's initial range.  q_out[i] = range / sum, shift_out[i] =
// renormalisation bytes after symbol i.  Lets the CPU tests compare the formulation with the integer recurrence without a GPU.
extern "C" int crgpu_debug_rc_dp(const uint32_t* frq, const uint32_t* sum, uint64_t n, uint32_t* q_out, uint32_t* shift_out) {
    if (!frq || !sum || !q_out || !shift_out) return CRGPU_ERR_ARG;
    double R = RC_DP_R0;
    for (uint64_t i = 0; i < n; i++) {
        if (sum[i] == 0 || frq[i] == 0 || frq[i] > sum[i] || sum[i] >= (1u << 23)) return CRGPU_ERR_ARG;
        uint32_t q, msb;
        rc_dp_step(R, rc_dp_record(frq[i], sum[i]), q, msb);
        q_out[i] = q; shift_out[i] = 3u - (msb >> 3);
    }
    return CRGPU_OK;
}

// Test aid for the parallel range chain (cr_rcpar.cuh, rc_variant 8): runs it ON THE GPU over caller-supplied (frq, sum) symbols cut
// into `nstreams` consecutive streams of lens[] symbols, each starting from the coder's initial range.  q_out[i] = range / sum[i],
// shift_out[i] = renormalisation bytes after symbol i; stats_out (8 x u64): state steps, live jobs, merged jobs, seed retries, failed
// seeds, streams re-done serially, largest exit set, 0.
extern "C" int crgpu_debug_rc_parallel(crgpu_handle* h, const uint32_t* frq, const uint32_t* sum, uint64_t n, const uint64_t* lens, uint32_t nstreams,
                                       uint32_t job_symbols, uint32_t* q_out, uint32_t* shift_out, uint64_t* stats_out) {
#ifdef CRGPU_SIM
    (void)h; (void)frq; (void)sum; (void)n; (void)lens; (void)nstreams; (void)job_symbols; (void)q_out; (void)shift_out; (void)stats_out;
    return CRGPU_ERR_UNSUPPORTED;
#else
    if (!h || !frq || !sum || !lens || !q_out || !shift_out || n == 0 || n >= (1ull << 32) || nstreams == 0) return CRGPU_ERR_ARG;
    CR_SET_DEVICE(h);
    LzChain& c = h->chain;
    cudaStream_t stream = c.stream;
    std::vector<Tri> tri(n);
    for (uint64_t i = 0; i < n; i++) {
        if (sum[i] == 0 || frq[i] == 0 || frq[i] > sum[i] || sum[i] >= (1u << 23)) return CRGPU_ERR_ARG;
        tri[i].cum = 0; tri[i].frq = frq[i]; tri[i].sum = sum[i]; tri[i].magic = rc_magic(sum[i]);
    }
    std::vector<RcStream> streams(nstreams);
    uint64_t at = 0;
    for (uint32_t s = 0; s < nstreams; s++) {
        memset(&streams[s], 0, sizeof(RcStream));
        streams[s].ev_begin = (uint32_t)at; at += lens[s]; streams[s].ev_end = (uint32_t)at; streams[s].is_main = 0; streams[s].limit = 0xFFFFFFFFu;
    }
    if (at != n) return CRGPU_ERR_ARG;
    DevBuf d_tri, d_cin, d_q, d_sh, d_str;
    int rc = CRGPU_OK;
    auto run = [&]() -> int {
        CR_TRY(d_tri.reserve(n * sizeof(Tri) + 64)); CR_TRY(d_cin.reserve((n + 1) * 16 + 64)); CR_TRY(d_q.reserve(n * 4 + 64)); CR_TRY(d_sh.reserve(n * 4 + 64));
        CR_TRY(d_str.reserve(nstreams * sizeof(RcStream) + 64));
        CR_CUDA(cudaMemcpyAsync(d_tri.p, tri.data(), n * sizeof(Tri), cudaMemcpyHostToDevice, stream));
        CR_CUDA(cudaMemcpyAsync(d_str.p, streams.data(), nstreams * sizeof(RcStream), cudaMemcpyHostToDevice, stream));
        CR_LAUNCH(k_chain_inputs_dp, dim3(cr_div_up(n, 256)), dim3(256), stream, d_tri.as<Tri>(), (uint64_t)n, d_cin.as<uint4>());
        const uint32_t keep = c.rcpar.job_symbols;
        if (job_symbols) c.rcpar.job_symbols = job_symbols;
        const int r = c.rcpar.run(stream, d_str.as<RcStream>(), nstreams, d_q.as<uint32_t>() /* unused: no main streams */, 0, n, d_tri.as<Tri>(), d_tri.as<Tri>(),
                                  d_cin.as<uint4>(), d_cin.as<uint4>(), d_q.as<uint32_t>(), d_sh.as<uint32_t>(), d_q.as<uint32_t>(), d_sh.as<uint32_t>(), true);
        c.rcpar.job_symbols = keep;
        CR_TRY(r);
        CR_LAUNCH(k_msb_to_shifts, dim3(cr_div_up(n, 256)), dim3(256), stream, d_sh.as<uint32_t>(), (uint64_t)n);
        CR_CUDA(cudaMemcpyAsync(q_out, d_q.p, n * 4, cudaMemcpyDeviceToHost, stream));
        CR_CUDA(cudaMemcpyAsync(shift_out, d_sh.p, n * 4, cudaMemcpyDeviceToHost, stream));
        CR_CUDA(cudaStreamSynchronize(stream));
        return CRGPU_OK;
    };
    rc = run();
    d_tri.release(); d_cin.release(); d_q.release(); d_sh.release(); d_str.release();
    if (rc == CRGPU_OK && stats_out) {
        const RcpStats& st = c.rcpar.last;
        stats_out[0] = st.state_steps; stats_out[1] = st.live_jobs; stats_out[2] = st.merged_jobs; stats_out[3] = st.seed_retries;
        stats_out[4] = st.demoted_jobs; stats_out[5] = st.flagged_streams; stats_out[6] = st.max_e; stats_out[7] = st.decided_serial;
    }
    return rc;
#endif
}
