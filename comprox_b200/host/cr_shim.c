/*
 * cr_shim.c -- the reference's cr-* C API, symbol for symbol, on top of libcrgpu.so.
 *
 * Compile with -DCR_VARIANT=0|1|2 (comprolz | comprop | comprox) and link it in place of the reference's
 * cr-datablock.c, cr-filter.c, filter_*.c, cr-dicpick.c, cr-diccode.c, cr-ppm.c, cr-o2model.c, cr-model.c,
 * cr-rangecoder.c and <variant>/cr-coder.c + cr-matcher.c: the UNMODIFIED src/main.c and src/<variant>/main.c then run
 * their serial block loop (src/main.c:174-206, 263-292) over the CUDA path, one stage call per block.  This is the
 * literal drop-in; it is exact because the handle below lives as long as the reference's file-scope statics do, but it
 * exposes one block of parallelism at a time -- the fast route is the whole-container call (host/cr_main.c).
 *
 *   symbol here                      replaces (file:line under /root/reference)
 *   data_block_reserve/resize/add/destroy   src/cr-datablock.c:31-57       (host memory only)
 *   filter_inplace                   src/cr-filter.c:33-73
 *   dicpick                          src/cr-dicpick.c:164-259
 *   dic_lcp_encode / dic_lcp_decode  src/cr-dicpick.c:261-346
 *   dictionary_load                  src/cr-diccode.c:76-120
 *   dictionary_encode / _decode      src/cr-diccode.c:142-221 / 223-283
 *   reset_models, lzencode, lzdecode src/<variant>/cr-coder.c
 *   flexible_parsing, match_limit    src/rolzmain/cr-matcher.c:31, src/roxmain/cr-matcher.c:32,37
 *
 * Error behaviour: the reference API is void; here every failure of the CUDA path prints the crgpu error and abort()s.
 * There is no CPU fallback.  libcrgpu.so is found through $CRGPU_LIB or next to the executable
 * (../comprox_b200/libcrgpu.so), like host/cr_main.c.
 */
#include <dlfcn.h>
#include <libgen.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "../../include/crgpu.h"

#ifndef CR_VARIANT
#define CR_VARIANT 0
#endif

/* layout of src/cr-datablock.h:35-39 */
typedef struct data_block_t { uint8_t* m_data; uint32_t m_size; uint32_t m_capacity; } data_block_t;

#if CR_VARIANT != 1
int flexible_parsing = 0;
#endif
#if CR_VARIANT == 2
uint32_t match_limit = 40;
#endif

/* ------------------------------------------------------------------ growable host buffer */
static void die(const char* what, int rc);

void data_block_reserve(data_block_t* b, uint32_t size) {
    if (size <= b->m_capacity && b->m_data) return;
    uint64_t cap = (uint64_t)size + size / 4 + 64;
    if (cap > 0xFFFFFFFFull) cap = 0xFFFFFFFFull;
    uint8_t* p = realloc(b->m_data, cap);
    if (!p) { fprintf(stderr, "cr_shim: out of host memory (%llu bytes)\n", (unsigned long long)cap); abort(); }
    b->m_data = p; b->m_capacity = (uint32_t)cap;
}
void data_block_resize(data_block_t* b, uint32_t size) { data_block_reserve(b, size); b->m_size = size; }
void data_block_add(data_block_t* b, uint8_t byte) { data_block_resize(b, b->m_size + 1); b->m_data[b->m_size - 1] = byte; }
void data_block_destroy(data_block_t* b) { free(b->m_data); b->m_data = NULL; b->m_size = b->m_capacity = 0; }

/* ------------------------------------------------------------------ the library and the one handle of this process */
static struct {
    void* lib;
    int (*create)(crgpu_handle**, int, int, void*);
    const char* (*strerror)(int);
    int (*reset_models)(crgpu_handle*);
    int (*lzencode)(crgpu_handle*, const uint8_t*, const uint32_t*, uint32_t, int, uint8_t*, uint64_t, uint32_t*);
    int (*lzdecode)(crgpu_handle*, const uint8_t*, uint32_t, uint8_t*, uint64_t, uint32_t*);
    int64_t (*lzdecode_size)(int, const uint8_t*, uint32_t);
    int (*dicpick)(crgpu_handle*, const uint8_t*, uint64_t, uint8_t*, uint64_t, uint64_t*);
    int (*lcp_encode)(const uint8_t*, uint64_t, uint8_t*, uint64_t, uint64_t*);
    int (*lcp_decode)(const uint8_t*, uint64_t, uint8_t*, uint64_t, uint64_t*);
    int (*dictionary_load)(crgpu_handle*, const char*, int);
    int (*dictionary_encode)(crgpu_handle*, const uint8_t*, uint32_t, uint8_t*, uint64_t, uint32_t*);
    int (*dictionary_decode)(crgpu_handle*, const uint8_t*, uint32_t, uint8_t*, uint64_t, uint32_t*);
    int (*filter_inplace)(crgpu_handle*, uint8_t*, uint32_t, int);
    int (*set_option)(crgpu_handle*, const char*, int64_t);
    crgpu_handle* h;
} G;

static void die(const char* what, int rc) {
    fprintf(stderr, "cr_shim: %s failed: %s (no CPU fallback)\n", what, G.strerror ? G.strerror(rc) : "library not loaded");
    abort();
}

static void* sym(const char* name) {
    void* p = dlsym(G.lib, name);
    if (!p) { fprintf(stderr, "cr_shim: %s is missing from libcrgpu.so\n", name); abort(); }
    return p;
}

static crgpu_handle* handle(void) {
    if (G.h) return G.h;
    char path[4096];
    const char* env = getenv("CRGPU_LIB");
    if (env) snprintf(path, sizeof path, "%s", env);
    else {
        char exe[4096]; ssize_t k = readlink("/proc/self/exe", exe, sizeof exe - 1); exe[k > 0 ? k : 0] = 0;
        snprintf(path, sizeof path, "%s/../comprox_b200/libcrgpu.so", dirname(exe));
    }
    G.lib = dlopen(path, RTLD_NOW);
    if (!G.lib) { fprintf(stderr, "cr_shim: cannot load %s: %s (no CPU fallback)\n", path, dlerror()); abort(); }
    G.create = sym("crgpu_create"); G.strerror = sym("crgpu_strerror"); G.reset_models = sym("crgpu_reset_models");
    G.lzencode = sym("crgpu_lzencode"); G.lzdecode = sym("crgpu_lzdecode"); G.lzdecode_size = sym("crgpu_lzdecode_size");
    G.dicpick = sym("crgpu_dicpick"); G.lcp_encode = sym("crgpu_dic_lcp_encode"); G.lcp_decode = sym("crgpu_dic_lcp_decode");
    G.dictionary_load = sym("crgpu_dictionary_load"); G.dictionary_encode = sym("crgpu_dictionary_encode");
    G.dictionary_decode = sym("crgpu_dictionary_decode"); G.filter_inplace = sym("crgpu_filter_inplace"); G.set_option = sym("crgpu_set_option");
    int rc = G.create(&G.h, CR_VARIANT, getenv("CRGPU_DEVICE") ? atoi(getenv("CRGPU_DEVICE")) : 0, NULL);
    if (rc) die("crgpu_create", rc);
    return G.h;
}

/* ------------------------------------------------------------------ src/cr-filter.h:38 */
int filter_inplace(unsigned char* buf, uint32_t len, int en_de) {
    int rc = (handle(), G.filter_inplace(G.h, buf, len, en_de));
    if (rc < 0) die("filter_inplace", rc);
    return rc;
}

/* ------------------------------------------------------------------ src/cr-dicpick.h:40-42 */
void dicpick(FILE* fp, data_block_t* dic_block) {
    handle();
    size_t cap = 1 << 20, n = 0, r;
    uint8_t* in = malloc(cap);
    if (!in) abort();
    while ((r = fread(in + n, 1, cap - n, fp)) > 0) {
        n += r;
        if (n == cap) { cap *= 2; in = realloc(in, cap); if (!in) abort(); }
    }
    const uint64_t tcap = 25000 * 24 + 64;
    uint8_t* text = malloc(tcap);
    uint64_t tn = 0;
    int rc = G.dicpick(G.h, in, n, text, tcap, &tn);
    if (rc) die("dicpick", rc);
    const uint32_t at = dic_block->m_size;                     /* the reference appends (data_block_add) */
    data_block_resize(dic_block, at + (uint32_t)tn);
    memcpy(dic_block->m_data + at, text, tn);
    free(text); free(in);
}

static void lcp(data_block_t* b, int decode) {
    handle();
    const uint64_t cap = decode ? 25000 * 24 + 64 : (uint64_t)b->m_size + 64;
    uint8_t* out = malloc(cap);
    uint64_t n = 0;
    int rc = decode ? G.lcp_decode(b->m_data, b->m_size, out, cap, &n) : G.lcp_encode(b->m_data, b->m_size, out, cap, &n);
    if (rc) die(decode ? "dic_lcp_decode" : "dic_lcp_encode", rc);
    data_block_resize(b, (uint32_t)n);
    memcpy(b->m_data, out, n);
    free(out);
}
void dic_lcp_encode(data_block_t* dic_block) { lcp(dic_block, 0); }
void dic_lcp_decode(data_block_t* dic_block) { lcp(dic_block, 1); }

/* ------------------------------------------------------------------ src/cr-diccode.h:44-48 */
int dictionary_load(const char* dicstr, int init_trie) {
    int rc = (handle(), G.dictionary_load(G.h, dicstr, init_trie));
    if (rc < 0) die("dictionary_load", rc);
    return rc;
}

void dictionary_encode(data_block_t* ib, data_block_t* ob) {
    handle();
    uint32_t n = 0;
    fprintf(stderr, "%s\n", "-> running static dictionary encoding...");
    data_block_resize(ob, ib->m_size + 1);
    int rc = G.dictionary_encode(G.h, ib->m_data, ib->m_size, ob->m_data, ob->m_capacity, &n);
    if (rc) die("dictionary_encode", rc);
    ob->m_size = n;
}

void dictionary_decode(data_block_t* ib, data_block_t* ob, FILE* fpout_sync) {
    handle();
    fprintf(stderr, "%s\n", "-> running static dictionary decoding...");
    if (ib->m_size == 0) die("dictionary_decode (empty block)", CRGPU_ERR_ARG);
    const int coded = ib->m_data[ib->m_size - 1] != 0;
    uint64_t cap = coded ? (uint64_t)ib->m_size * 4 + (1u << 20) : ib->m_size;
    uint32_t n = 0;
    int rc;
    for (;;) {                                                  /* the decoded size is not stored: grow until it fits */
        if (cap > 0xFFFFFF00ull) cap = 0xFFFFFF00ull;
        data_block_resize(ob, (uint32_t)cap);
        rc = G.dictionary_decode(G.h, ib->m_data, ib->m_size, ob->m_data, ob->m_capacity, &n);
        if (rc != CRGPU_ERR_ARG || cap >= 0xFFFFFF00ull) break;
        cap *= 2;
    }
    if (rc) die("dictionary_decode", rc);
    ob->m_size = n;
    /* src/cr-diccode.c:275-278: a coded block is flushed to the output file pair by pair and `ob` is left empty
       (which is why the reference cannot inverse-filter such a block); a stored block stays in `ob`. */
    if (coded && fpout_sync) { fwrite(ob->m_data, 1, ob->m_size, fpout_sync); data_block_resize(ob, 0); }
}

/* ------------------------------------------------------------------ src/main.c:57-59 */
void reset_models(void) {
    int rc = (handle(), G.reset_models(G.h));
    if (rc) die("reset_models", rc);
}

void lzencode(data_block_t* ib, data_block_t* ob, int print_information) {
    handle();
    (void)print_information;
    int rc;
#if CR_VARIANT != 1
    if ((rc = G.set_option(G.h, "flexible", flexible_parsing)) != 0) die("set_option(flexible)", rc);
#endif
#if CR_VARIANT == 2
    if ((rc = G.set_option(G.h, "match_limit", match_limit)) != 0) die("set_option(match_limit)", rc);
#endif
    uint32_t n = 0;
    data_block_resize(ob, ib->m_size + 64);                     /* payload <= inner header + input */
    /* chain_ends = 0: more blocks of this model chain may follow, so a block that hits "cannot compress"
       (src/rolzmain/cr-coder.c:231-233) must leave the models exactly where the reference's aborted loop leaves them;
       the library replays that case exactly (LzChain::encode_blocks). */
    rc = G.lzencode(G.h, ib->m_data, &ib->m_size, 1, /*chain_ends=*/0, ob->m_data, ob->m_capacity, &n);
    if (rc) die("lzencode", rc);
    ob->m_size = n;
}

void lzdecode(data_block_t* ib, data_block_t* ob, int print_information) {
    handle();
    (void)print_information;
    int64_t size = G.lzdecode_size(CR_VARIANT, ib->m_data, ib->m_size);
    if (size < 0) die("lzdecode (payload shorter than its header)", (int)size);
    uint32_t n = 0;
    data_block_resize(ob, (uint32_t)size);
    int rc = G.lzdecode(G.h, ib->m_data, ib->m_size, ob->m_data, ob->m_capacity, &n);
    if (rc) die("lzdecode", rc);
    ob->m_size = n;
}
