/*
 * crgpu.h -- C ABI of libcrgpu.so, the B200-native (sm_100a CUDA) replacement for comprox's
 * block-compression hot path.  Plain C types only; every buffer is caller-owned HOST memory unless a
 * function says otherwise.  All functions return CRGPU_OK (0) or a negative CRGPU_ERR_* code; there is no
 * CPU fallback -- without a CUDA device crgpu_create() fails with CRGPU_ERR_NO_DEVICE.
 *
 * Each entry point names the reference interface (file:line under /root/reference) it replaces.  A handle
 * owns the state the reference keeps in file-scope statics (adaptive models, dictionary trie, filter
 * continuation state), so -- unlike the reference -- several handles can coexist in one process.
 * Like the reference, a single handle is not thread-safe.
 */
#ifndef CRGPU_H
#define CRGPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRGPU_OK                    0
#define CRGPU_ERR_NO_DEVICE        -1
#define CRGPU_ERR_CUDA             -2
#define CRGPU_ERR_ARG              -3
#define CRGPU_ERR_VOCAB_OVERFLOW   -4   /* > 325000 distinct words: reference prune is order dependent (src/cr-dicpick.c:115-144) */
#define CRGPU_ERR_HASH_COLLISION   -5
#define CRGPU_ERR_MIDCHAIN_ABORT   -6   /* a non-final block hit "cannot compress" (src/rolzmain/cr-coder.c:231-233) while the exact
                                           replay of that case is switched off (crgpu_set_option "exact_aborts" = 0) */
#define CRGPU_ERR_UNSUPPORTED      -7
#define CRGPU_ERR_OOM              -8
#define CRGPU_ERR_CORRUPT          -9   /* decoder: the container is damaged (a stream, match or dictionary reference leaves its buffer) */

#define CRGPU_ROLZ 0   /* comprolz: src/rolzmain */
#define CRGPU_LZP  1   /* comprop : src/ropmain  */
#define CRGPU_LZ77 2   /* comprox : src/roxmain  */

typedef struct crgpu_handle crgpu_handle;

/* Creates a handle on CUDA device `device`.  `stream` is a cudaStream_t passed as void* (NULL = the legacy
 * default stream, CRGPU_OWN_STREAM = a private stream so that several handles overlap on one GPU); all work of the
 * handle is enqueued on it.  `variant` selects the lzencode implementation
 * exactly as linking src/rolzmain or src/ropmain does in the reference (Makefile:12-27). */
#define CRGPU_OWN_STREAM ((void*)1)   /* pass as `stream`: the handle creates (and owns) a private non-blocking stream */
int  crgpu_create(crgpu_handle** out, int variant, int device, void* stream);
void crgpu_destroy(crgpu_handle* h);
const char* crgpu_strerror(int code);

/* reset_models()  -- src/main.c:57, src/rolzmain/cr-coder.c:78-96, src/ropmain/cr-coder.c:73-83 */
int crgpu_reset_models(crgpu_handle* h);

/* lzencode(ib, ob, print)  -- src/main.c:58, src/rolzmain/cr-coder.c:139-264, src/ropmain/cr-coder.c:119-229.
 * Batch form: encodes `nblocks` CONSECUTIVE blocks of the current model chain in one call (the reference
 * calls lzencode once per block; models carry over from block to block and from call to call).
 * in        : the blocks' bytes back to back (what dictionary_encode produced)
 * sizes     : nblocks sizes
 * out       : receives the payloads back to back; out_sizes[i] = size of payload i
 * chain_ends: nonzero if reset_models() follows before any further block (as after the dictionary payload,
 *             src/main.c:164-165, or at end of file).
 * A block the reference's coder loop gives up on ("cannot compress") is stored raw and leaves the models exactly where the
 * reference's aborted loop leaves them, so the following blocks stay byte-identical.
 * out_cap must be >= sum(sizes[i] + 32) (a stored block is its bytes behind the inner header); a smaller `out` is refused with
 * CRGPU_ERR_ARG before any block has touched the models, so the call can be repeated. */
int crgpu_lzencode(crgpu_handle* h, const uint8_t* in, const uint32_t* sizes, uint32_t nblocks, int chain_ends,
                   uint8_t* out, uint64_t out_cap, uint32_t* out_sizes);

/* Whole-container compression: everything cr_main does between write_magic and the last fwrite
 * (src/main.c:153-206): dicpick over the file, dictionary payload, then per block filter_inplace (-F),
 * dictionary_encode and lzencode (unless -p), framed with the 6-byte block headers.  `out` receives the bytes the
 * reference CLI would have written to its output file for the same input and switches.
 * out_cap must be >= crgpu_compress_bound(n, block_size).  n <= 4 GiB per container (CRGPU_ERR_UNSUPPORTED above that:
 * split the input into shards, crgpu_compress_batch). */
typedef struct crgpu_config {
    uint32_t block_size;    /* -b, in BYTES (reference default 16 MiB, src/main.c:62) */
    int32_t  filt;          /* -F  cr_filt_enable */
    int32_t  prec;          /* -p  cr_prec_enable */
    int32_t  flexible;      /* -f  flexible parsing (comprolz only, src/rolzmain/cr-matcher.c:143-162) */
    uint64_t window_bytes;  /* raw bytes resident per window, 0 = default (512 MiB) */
} crgpu_config;
uint64_t crgpu_compress_bound(uint64_t n, uint32_t block_size);
int crgpu_compress(crgpu_handle* h, const crgpu_config* cfg, const uint8_t* in, uint64_t n, uint8_t* out, uint64_t out_cap, uint64_t* out_n);

/* Shard mode: `count` independent containers compressed side by side by `nhandles` handles, one host thread
 * per handle inside the call, containers dealt to whichever handle is idle; what `xargs -P` over the reference CLI does with host
 * cores.  All handles must be of one variant; they may sit on different devices (the multi-GPU form: no data-path collective, every
 * container goes straight from its GPU to outs[i]) or share a device on private streams (CRGPU_OWN_STREAM), where the serial range
 * chains of different containers overlap.  With three or more handles on one device (and as many containers) the range chains are
 * walked serially instead of cut into jobs -- one warp each, they overlap each other and all other work, where the tracking kernels
 * of the cut chain fill the device; "rc_serial" (crgpu_set_option) overrides the choice.  Handles that share a device need their own
 * hardware work queues: the library sets CUDA_DEVICE_MAX_CONNECTIONS=32 when it is loaded unless the variable is already set; a host
 * program that creates its CUDA context BEFORE loading the library must export it itself.  outs[i] / out_lens[i] receive container
 * i, exactly the bytes crgpu_compress gives for ins[i]; out_caps[i] >= crgpu_compress_bound(in_lens[i], cfg->block_size).  Returns
 * the first error. */
int crgpu_compress_batch(crgpu_handle* const* hs, uint32_t nhandles, const crgpu_config* cfg, uint32_t count,
                         const uint8_t* const* ins, const uint64_t* in_lens, uint8_t* const* outs, const uint64_t* out_caps, uint64_t* out_lens);

/* dicpick(fp, dic_block)  -- src/cr-dicpick.h:40, src/cr-dicpick.c:164-259: word statistics of the whole input ->
 * dictionary text ("\x20\x20\n" "http://www.\n" word\n ... NUL).  Stage-level entry point (crgpu_compress calls the
 * same code); exact also beyond 325000 distinct words, where the reference prunes in arrival order. */
int crgpu_dicpick(crgpu_handle* h, const uint8_t* in, uint64_t n, uint8_t* out, uint64_t out_cap, uint64_t* out_n);

/* Whole-container decompression: cr_main's decode branch (src/main.c:220-302): check_magic, dictionary payload,
 * then per block lzdecode (unless the block was written with -p), dictionary_decode and filter_inplace(FILTER_DEC).
 * One container is one serial model chain, so this call runs the entropy stage on a single GPU thread: it is
 * provided for completeness and for decoding MANY containers side by side (one handle each); a single container
 * decodes faster on the host.  Returns CRGPU_ERR_ARG if the magic does not match the handle's variant or out_cap
 * is too small.  The container does not store its decoded size; when out_cap is too small *out_n receives the size needed
 * (it stays 0 for the other errors -- set it to 0 before the call), and calling again with the SAME `in` pointer and length and
 * a large enough buffer resumes behind the entropy stage instead of decoding twice. */
int crgpu_decompress(crgpu_handle* h, const uint8_t* in, uint64_t n, uint8_t* out, uint64_t out_cap, uint64_t* out_n);

/* Many containers in flight.  `count` containers are decoded side by side, one warp each, by
 * ONE launch per phase (dictionary payloads, then data blocks); hs[i] supplies the model state, matcher tables and
 * buffers of container i, so the number of handles is the number of containers in flight.  All handles must have been
 * created for the same variant and device and on the SAME stream (not CRGPU_OWN_STREAM).  out_lens[i] receives the
 * decoded size.  Returns the first error (outputs are then undefined). */
int crgpu_decompress_batch(crgpu_handle* const* hs, uint32_t count, const uint8_t* const* ins, const uint64_t* in_lens,
                           uint8_t* const* outs, const uint64_t* out_caps, uint64_t* out_lens);

/* ------------------------------------------------------------------------------------------------------------------
 * Stage-level entry points: ONE block per call, argument meaning of the reference's cr-* functions.
 * The handle carries what the reference keeps in statics between calls (filter continuation state, dictionary trie and
 * word table, adaptive models + PPM context).  comprox_b200/host/cr_shim.c wraps them into the reference's own
 * signatures (data_block_t, void returns), so that the UNMODIFIED src/main.c links against this library.  The
 * whole-container calls above are the fast path: per-block calls are exact but expose one block of parallelism.
 * ------------------------------------------------------------------------------------------------------------------ */

/* filter_inplace(buf, len, en_de)  -- src/cr-filter.h:38, src/cr-filter.c:33-73.  In place on HOST memory;
 * en_de: 0 = FILTER_ENC, 1 = FILTER_DEC.  Returns 1 if a filter fired (-> block header m_filt), 0 if not, < 0 on error.
 * An image may straddle calls, exactly as it straddles blocks in the reference. */
int crgpu_filter_inplace(crgpu_handle* h, uint8_t* buf, uint32_t len, int en_de);

/* dic_lcp_encode / dic_lcp_decode  -- src/cr-dicpick.h:41-42, src/cr-dicpick.c:261-346.  Front coding of the dictionary
 * text; host-side (the text is a few hundred KB at most), no handle needed.  `text` includes its final NUL, as dicpick
 * leaves it; the decoder returns the text with the NUL. */
int crgpu_dic_lcp_encode(const uint8_t* text, uint64_t n, uint8_t* out, uint64_t out_cap, uint64_t* out_n);
int crgpu_dic_lcp_decode(const uint8_t* in, uint64_t n, uint8_t* out, uint64_t out_cap, uint64_t* out_n);

/* dictionary_load(dicstr, init_trie)  -- src/cr-diccode.h:44, src/cr-diccode.c:76-120.  init_trie = 1 before encoding
 * (builds and uploads the trie), 0 before decoding.  Returns the number of dictionary words, or < 0. */
int crgpu_dictionary_load(crgpu_handle* h, const char* dicstr, int init_trie);

/* dictionary_encode(ib, ob)  -- src/cr-diccode.h:46, src/cr-diccode.c:142-221,285-362: escape selection, 1 000 000-byte
 * sub-chunk pairs, framing, "raw + 0" when not smaller.  out_cap >= n + 1. */
int crgpu_dictionary_encode(crgpu_handle* h, const uint8_t* in, uint32_t n, uint8_t* out, uint64_t out_cap, uint32_t* out_n);

/* dictionary_decode(ib, ob, NULL)  -- src/cr-diccode.h:48, src/cr-diccode.c:223-283,364-425.  Returns CRGPU_ERR_ARG if
 * out_cap is too small (the decoded size is only known after the framing has been walked). */
int crgpu_dictionary_decode(crgpu_handle* h, const uint8_t* in, uint32_t n, uint8_t* out, uint64_t out_cap, uint32_t* out_n);

/* lzdecode(ib, ob, print)  -- src/main.c:59, src/rolzmain/cr-coder.c:287-379, src/ropmain/cr-coder.c:231-292,
 * src/roxmain/cr-coder.c:321-526.  One payload per call; models carry over from call to call until
 * crgpu_reset_models().  crgpu_lzdecode_size() reads the decoded size from the payload's inner header. */
int crgpu_lzdecode(crgpu_handle* h, const uint8_t* in, uint32_t n, uint8_t* out, uint64_t out_cap, uint32_t* out_n);
int64_t crgpu_lzdecode_size(int variant, const uint8_t* in, uint32_t n);

/* Copies `in` to HBM ahead of time.  A following crgpu_compress(h, cfg, in, n, ...) with the same pointer and
 * length (and no -F, which rewrites the staged bytes in place) then skips its host-to-device copy; used by
 * bench.py to time the device-resident path separately from the end-to-end path. */
int crgpu_stage_input(crgpu_handle* h, const uint8_t* in, uint64_t n);

/* Counters of a handle.  "cut_blocks": blocks of the data chain that hit "cannot compress" in mid-chain since the handle was created
 * and were stored raw with the exact replay.  Such a container is byte-identical to the reference CLI's, but NEITHER
 * the reference decoder NOR crgpu_decompress can read the blocks behind the stored one (the decoder skips a stored block without
 * touching its models; the reference crashes, this library returns CRGPU_ERR_CORRUPT).  crgpu_compress still returns CRGPU_OK for
 * parity with the reference; callers that need a decodable archive check "last_cut_blocks" (the most recent crgpu_compress call)
 * or switch the replay off ("exact_aborts" = 0 -> CRGPU_ERR_MIDCHAIN_ABORT).  Returns < 0 for an unknown name. */
int64_t crgpu_get_stat(crgpu_handle* h, const char* name);

/* Number of CUDA kernels this library has launched in this process (all handles). */
uint64_t crgpu_launch_count(void);

/* Test / profiling aid: copies an intermediate array of the most recent lzencode window to the host.
 * what: "span" (u8 per position), "tidx" (u8 per position), "ev_ctx" (u32 per event), "ev_sym" (u8 per event),
 *       "pred" (u8 per event), "dense" (4 x u32 per main-stream triple), "dense_side" (4 x u32 per side triple).
 * Returns the number of BYTES available (copies min(cap, available)), or a negative error. */
int64_t crgpu_debug_fetch(crgpu_handle* h, const char* what, void* dst, uint64_t cap);

/* Test aid for the device-wide primitives (cr_sort.cuh).  key_bytes = 4 or 8: stable radix sort of n (key, value)
 * pairs on key bits [begin_bit, end_bit); key_bytes = 0: exclusive prefix sum of the n uint32 in `vals` (keys unused).
 * Host arrays in, host arrays out (keys_out may be NULL for the scan). */
int crgpu_debug_sort(crgpu_handle* h, const void* keys, const uint32_t* vals, uint64_t n, int key_bytes, int begin_bit, int end_bit,
                     void* keys_out, uint32_t* vals_out);

/* Test aid: the double-precision form of the range recurrence (k_range_chain<7>, opt-in through crgpu_set_option "rc_variant" = 7)
 * walked on the HOST with the same step function the kernel uses; q_out[i] = range / sum[i], shift_out[i] = renormalisation bytes. */
int crgpu_debug_rc_dp(const uint32_t* frq, const uint32_t* sum, uint64_t n, uint32_t* q_out, uint32_t* shift_out);

/* Test aid for the parallel range chain (cr_rcpar.cuh, "rc_variant" = 8): runs it on the GPU over caller-supplied (frq, sum) symbols cut
 * into `nstreams` consecutive streams of lens[] symbols, each from the coder's initial range.  q_out / shift_out as above;
 * stats_out (8 x u64, may be NULL): state steps, live jobs, merged jobs, seed retries, failed seeds, streams re-done serially, largest
 * exit set, 0.  job_symbols = 0 keeps the handle's setting. */
int crgpu_debug_rc_parallel(crgpu_handle* h, const uint32_t* frq, const uint32_t* sum, uint64_t n, const uint64_t* lens, uint32_t nstreams,
                            uint32_t job_symbols, uint32_t* q_out, uint32_t* shift_out, uint64_t* stats_out);

/* Tuning / test switches.  "scalar_models" = 1 runs the scalar model and coder kernels (the ones the CPU
 * kernel-logic simulation checks) instead of the warp-cooperative ones; results are identical.
 * "exact_aborts" = 0 turns the exact replay of mid-chain "cannot compress" blocks off (CRGPU_ERR_MIDCHAIN_ABORT instead).
 * "flexible" = the reference's global flexible_parsing (-f) for crgpu_lzencode (crgpu_compress takes it from its
 * config); "match_limit" = comprox -m.
 * Formulation switches (every setting gives the same bytes; the defaults are the measured best, profiles/round2_summary.md):
 * "rc_variant" 1..8 (8 = range chain cut into jobs), "rc_serial" -1 / 0 / 1 (automatic / always cut / one serial job per stream),
 * "rc_job_symbols", "rc_late_cfg"; "o2_hot_variant" 1..3 and "o1_hot_variant" 1..2 (hot-context passes: 3 / 2 = chain of steps +
 * parallel evaluation), "o2_width" 0 / 256 / 512 / 1024 (events per o2 step, 0 = by the window's hit rate); "rolz_match_variant"
 * 1..2; "dict_mode" 0 / 1 (1 = the dictionary payload is coded by a second model chain beside the data blocks); "dc_listed" 0 / 1
 * (1 = dictionary substitution walks the trie one listed word start per lane); "dp_tiles" 0 / 1 (1 = word count through a
 * shared-memory table per tile: measured slower, off). */
int crgpu_set_option(crgpu_handle* h, const char* name, int64_t value);

/* Stage timing (CUDA events on the handle's stream, accumulated over calls until reset).
 * crgpu_profile(h, 1) enables and clears, crgpu_profile(h, 0) disables.  crgpu_profile_report writes
 * "stage_name milliseconds\n" lines into buf and returns the number of bytes written. */
int crgpu_profile(crgpu_handle* h, int enable);
int crgpu_profile_report(crgpu_handle* h, char* buf, uint64_t cap);

#ifdef __cplusplus
}
#endif
#endif
