/*
 * cr_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement ("oracle") of the comprox block-compression hot path, written from scratch as a
 * re-entrant library (the reference keeps everything in file-scope statics).  It is the checker used by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.  Nothing under comprox_b200/ may include,
 * link or call it: the product path is CUDA only.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py and tests/golden/ check this restatement
 *   (a) against the digests in tests/golden/kat.json (containers produced by the UNMODIFIED reference
 *       compiled by oracle/Makefile into oracle/_ref/), and
 *   (b) differentially against oracle/_ref/{comprolz,comprop} whenever those binaries are present.
 *
 * Every function cites the reference file:line (under /root/reference) whose behaviour it restates.
 */
#ifndef CR_ORACLE_H
#define CR_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { CRO_ROLZ = 0, CRO_LZP = 1, CRO_LZ77 = 2 };   /* comprolz, comprop, comprox */

typedef struct cro_buf { uint8_t* data; size_t size, cap; } cro_buf;
void cro_buf_free(cro_buf* b);

/* ---- trace records (optional; enabled with cro_trace_enable) ---- */
typedef struct cro_token  { uint32_t pos, len, idx; } cro_token;      /* idx = 0xFFFFFFFF for literals      */
typedef struct cro_event  { uint32_t ctx; uint32_t sym; } cro_event;  /* one ppm_encode call                */
typedef struct cro_triple { uint32_t cum, frq, sum, stream; } cro_triple; /* stream 0 = main, 1 = idx/len  */

typedef struct cro_ctx cro_ctx;

cro_ctx* cro_new(int variant);
void     cro_free(cro_ctx* c);
void     cro_reset_models(cro_ctx* c);               /* src/<variant>/cr-coder.c reset_models()             */
void     cro_set_flexible(cro_ctx* c, int on);       /* src/rolzmain/cr-matcher.c:31                         */
void     cro_set_match_limit(cro_ctx* c, uint32_t limit); /* -m, src/roxmain/cr-matcher.c:38 (LZ77 only)       */
void     cro_trace_enable(cro_ctx* c, int on);
void     cro_trace_clear(cro_ctx* c);
size_t   cro_trace_tokens(cro_ctx* c, const cro_token** out);
size_t   cro_trace_events(cro_ctx* c, const cro_event** out);
size_t   cro_trace_triples(cro_ctx* c, const cro_triple** out);

/* ---- stage API (mirrors the reference's cr-* C API, re-entrant) ---- */
int  cro_filter_inplace(cro_ctx* c, uint8_t* buf, uint32_t len, int en_de);            /* src/cr-filter.c:33      */
void cro_dicpick(const uint8_t* data, size_t n, cro_buf* dic_text);                     /* src/cr-dicpick.c:164    */
void cro_dic_lcp_encode(cro_buf* dic);                                                  /* src/cr-dicpick.c:261    */
void cro_dic_lcp_decode(cro_buf* dic);                                                  /* src/cr-dicpick.c:307    */
int  cro_dictionary_load(cro_ctx* c, const char* dicstr, int init_trie);                /* src/cr-diccode.c:76     */
void cro_dictionary_encode(cro_ctx* c, const uint8_t* in, uint32_t n, cro_buf* out);    /* src/cr-diccode.c:142    */
void cro_dictionary_decode(cro_ctx* c, const uint8_t* in, uint32_t n, cro_buf* out);    /* src/cr-diccode.c:223    */
void cro_lzencode(cro_ctx* c, const uint8_t* in, uint32_t n, cro_buf* out);             /* src/*main/cr-coder.c    */
void cro_lzdecode(cro_ctx* c, const uint8_t* in, uint32_t n, cro_buf* out);

/* ROLZ parse only (no entropy coding): token list of one block. src/rolzmain/cr-coder.c:109-137 */
size_t cro_rolz_parse(const uint8_t* data, uint32_t n, int flexible, cro_token** out);
/* LZP parse only. src/ropmain/cr-coder.c:95-118 */
size_t cro_lzp_parse(const uint8_t* data, uint32_t n, cro_token** out);

/* LZ77 parse only: idx = match position (0xFFFFFFFF for literals). src/roxmain/cr-coder.c:116-142 */
size_t cro_lz77_parse(const uint8_t* data, uint32_t n, int flexible, uint32_t match_limit, cro_token** out);

/* ---- container API: what cr_main does between fopen and fclose. src/main.c:137-218 / 220-302 ---- */
typedef struct cro_config {
    int      variant;       /* CRO_ROLZ (comprolz) or CRO_LZP (comprop) */
    uint32_t block_size;    /* bytes; reference default 16 MiB (src/main.c:62) */
    int      filt;          /* -F */
    int      prec;          /* -p */
    int      flexible;      /* -f (ROLZ, LZ77) */
    uint32_t match_limit;   /* -m (LZ77 only; 0 = the default 40) */
} cro_config;
int cro_compress(const cro_config* cfg, const uint8_t* in, size_t n, cro_buf* out);
int cro_decompress(int variant, const uint8_t* in, size_t n, cro_buf* out);

#ifdef __cplusplus
}
#endif
#endif
