/*
 * cr_oracle.c -- TEST INFRASTRUCTURE ONLY (see cr_oracle.h).
 *
 * A from-scratch, re-entrant CPU restatement of the comprox hot path:
 *   filters -> static word dictionary -> ROLZ / LZP parse -> PPM + order-0 models -> range coder,
 * and the inverse.  It restates *behaviour* (including the quirks listed in SURVEY.md App. B); the data
 * structures are our own (flat model tables, explicit state structs instead of function-local statics).
 * Parity: pinned against the unmodified reference build in oracle/_ref (tests/test_oracle_vs_ref.py) and the
 * committed digests tests/golden/kat.json.
 */
#include "cr_oracle.h"
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ small helpers */
static int is_lower(int c) { return c >= 'a' && c <= 'z'; }
static int is_upper(int c) { return c >= 'A' && c <= 'Z'; }
static int is_alpha(int c) { return is_lower(c) || is_upper(c); }
static int to_lower(int c) { return is_upper(c) ? c + 32 : c; }
static uint32_t umin(uint32_t a, uint32_t b) { return a < b ? a : b; }
static uint32_t rd32(const uint8_t* p) { return p[0] | p[1] << 8 | p[2] << 16 | (uint32_t)p[3] << 24; }
static uint32_t rd16(const uint8_t* p) { return p[0] | p[1] << 8; }
static void wr32(uint8_t* p, uint32_t v) { p[0] = v; p[1] = v >> 8; p[2] = v >> 16; p[3] = v >> 24; }

static void buf_reserve(cro_buf* b, size_t n) {
    if (n > b->cap) {
        size_t c = b->cap ? b->cap : 256;
        while (c < n) c += c / 2 + 64;
        b->data = (uint8_t*)realloc(b->data, c);
        b->cap = c;
    }
}
static void buf_push(cro_buf* b, uint8_t v) { buf_reserve(b, b->size + 1); b->data[b->size++] = v; }
static void buf_append(cro_buf* b, const void* p, size_t n) {
    buf_reserve(b, b->size + n);
    if (n) memcpy(b->data + b->size, p, n);
    b->size += n;
}
static void buf_push32(cro_buf* b, uint32_t v) { uint8_t t[4]; wr32(t, v); buf_append(b, t, 4); }
void cro_buf_free(cro_buf* b) { free(b->data); b->data = NULL; b->size = b->cap = 0; }

/* ------------------------------------------------------------------ trace */
typedef struct { void* p; size_t n, cap; } vec_t;
static void* vec_grow(vec_t* v, size_t elem) {
    if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 1024; v->p = realloc(v->p, v->cap * elem); }
    return (char*)v->p + (v->n++) * elem;
}

/* ------------------------------------------------------------------ range coder (src/cr-rangecoder.c) */
typedef struct { uint32_t low, range, follow, carry, cache; } rc_t;

static void rc_enc_init(rc_t* r) { r->low = 0; r->range = 0xFFFFFFFFu; r->follow = 0; r->cache = 0; r->carry = 0; }

/* one byte leaves the 32-bit window; src/cr-rangecoder.c:44-58 */
static void rc_shift_out(rc_t* r, cro_buf* o) {
    if (r->low < 0xFF000000u || r->carry) {
        buf_push(o, (uint8_t)(r->cache + r->carry));
        for (; r->follow; r->follow--) buf_push(o, (uint8_t)(r->carry - 1));
        r->cache = r->low >> 24;
        r->carry = 0;
    } else {
        r->follow++;
    }
    r->low <<= 8;
}
/* src/cr-rangecoder.c:60-70 */
static void rc_encode(rc_t* r, uint32_t cum, uint32_t frq, uint32_t sum, cro_buf* o) {
    uint32_t q = r->range / sum, add = cum * q, nl = r->low + add;
    r->carry += nl < r->low;
    r->low = nl;
    r->range = q * frq;
    while (r->range < (1u << 24)) { r->range <<= 8; rc_shift_out(r, o); }
}
static void rc_flush(rc_t* r, cro_buf* o) { for (int i = 0; i < 5; i++) rc_shift_out(r, o); } /* :72-79 */

/* decoder: `cache` holds the code window; src/cr-rangecoder.c:81-104 */
static void rc_dec_init(rc_t* r, const uint8_t** in) {
    rc_enc_init(r);
    for (int i = 0; i < 5; i++) r->cache = (r->cache << 8) + *(*in)++;
}
static uint32_t rc_dec_cum(rc_t* r, uint32_t sum) { r->range /= sum; return r->cache / r->range; }
static void rc_dec_consume(rc_t* r, uint32_t cum, uint32_t frq, const uint8_t** in) {
    r->cache -= cum * r->range;
    r->range *= frq;
    while (r->range < (1u << 24)) { r->cache = (r->cache << 8) + *(*in)++; r->range <<= 8; }
}

/* ------------------------------------------------------------------ order-0 adaptive model (src/cr-model.c) */
typedef struct { uint16_t frq[256]; uint32_t total; } m0_t;

static void m0_recount(m0_t* m) { m->total = 0; for (int i = 0; i < 256; i++) m->total += m->frq[i]; }
static uint32_t m0_cum(const m0_t* m, int s) { uint32_t c = 0; for (int i = 0; i < s; i++) c += m->frq[i]; return c; }
/* src/cr-model.c:55-77: add, then halve (rounding up) once the total passes 32000 */
static void m0_update(m0_t* m, int s, int inc) {
    m->frq[s] += inc; m->total += inc;
    if (m->total > 32000) { for (int i = 0; i < 256; i++) m->frq[i] = (m->frq[i] + 1) / 2; m0_recount(m); }
}

/* ------------------------------------------------------------------ PPM (src/cr-ppm.c, src/cr-o2model.c) */
#define O3_SLOTS (1u << 22)
typedef struct {
    uint8_t  (*o1)[256];        /* [256][256]                                                    */
    uint8_t  (*o2)[258];        /* [65536][258]; 256 = o3-hit flag, 257 = escape                 */
    uint8_t* o3_byte;           /* predicted byte per slot  (reference packs 2 slots in 3 bytes)  */
    uint8_t* o3_conf;           /* 4-bit confidence per slot                                      */
    uint32_t ctx;
} ppm_t;

static void ppm_alloc(ppm_t* p) {
    p->o1 = malloc(256 * 256); p->o2 = malloc(65536 * 258);
    p->o3_byte = malloc(O3_SLOTS); p->o3_conf = malloc(O3_SLOTS);
}
static void ppm_release(ppm_t* p) { free(p->o1); free(p->o2); free(p->o3_byte); free(p->o3_conf); }
/* src/cr-ppm.c:34-48 + lazy o2 init src/cr-o2model.c:31-41 (done eagerly here: same observable state) */
static void ppm_reset(ppm_t* p) {
    memset(p->o1, 1, 256 * 256);
    memset(p->o2, 0, 65536 * 258);
    for (int i = 0; i < 65536; i++) p->o2[i][256] = p->o2[i][257] = 1;
    memset(p->o3_byte, 0, O3_SLOTS); memset(p->o3_conf, 0, O3_SLOTS);
    p->ctx = 0;
}
static uint32_t o3_slot(uint32_t ctx) { return (ctx ^ (ctx >> 2)) & 0x3fffff; }      /* src/cr-ppm.c:66 */

/* src/cr-ppm.c:69-88 */
static void o3_update(ppm_t* p, int c) {
    uint32_t s = o3_slot(p->ctx);
    int f = p->o3_conf[s];
    if (c >= 0) {
        f = (f > 1) + (f > 2) + (f > 4) + (f > 8);
        if (f == 0) { p->o3_byte[s] = (uint8_t)c; f = 1; }
    } else {
        f += f < 15;
    }
    p->o3_conf[s] = f;
}
static uint32_t o2_body(const uint8_t* f) { uint32_t t = 0; for (int i = 0; i < 256; i++) t += f[i]; return t; }
static uint32_t o2_cum(const uint8_t* f, int s) { uint32_t c = 0; for (int i = 0; i < s && i < 256; i++) c += f[i]; if (s > 256) c += f[256]; return c; }
/* src/cr-o2model.c:43-72; returns 1 when the table was rescaled. frq is uint8 and wraps like the reference. */
static int o2_update(uint8_t* f, int s, int inc) {
    f[s] = (uint8_t)(f[s] + inc);
    if (f[s] > 250) {
        int ee = 1;
        for (int i = 0; i < 256; i++) { f[i] >>= 1; ee += f[i] == 1; }
        f[256] = (f[256] + 1) / 2;
        f[257] = (uint8_t)ee;
        return 1;
    }
    return 0;
}
static void o1_update(uint8_t* o1, int c) {                                           /* src/cr-ppm.c:90-97 */
    if (++o1[c] >= 255) for (int i = 0; i < 256; i++) o1[i] -= o1[i] / 2;
}
#define O1F(v) ((int)(v) * 8 - 7)

/* forward decl of trace sink */
struct cro_ctx;
static void trace_event(struct cro_ctx* c, uint32_t ctx, uint32_t sym);
static void trace_triple(struct cro_ctx* c, uint32_t cum, uint32_t frq, uint32_t sum, uint32_t stream);
static void emit(struct cro_ctx* c, rc_t* r, uint32_t cum, uint32_t frq, uint32_t sum, uint32_t stream, cro_buf* o) {
    trace_triple(c, cum, frq, sum, stream);
    rc_encode(r, cum, frq, sum, o);
}

/* src/cr-ppm.c:103-167 */
static void ppm_encode(struct cro_ctx* c, ppm_t* p, rc_t* r, int ch, cro_buf* o) {
    uint8_t* f = p->o2[p->ctx & 0xffff];
    uint8_t* o1 = p->o1[p->ctx & 0xff];
    int pred = p->o3_byte[o3_slot(p->ctx)];
    uint32_t pf = f[pred], body = o2_body(f), sum = body + f[256] + f[257] - pf;
    trace_event(c, p->ctx, (uint32_t)ch);
    if (ch == pred) {
        emit(c, r, body - pf, f[256], sum, 0, o);
        o2_update(f, 256, 1);
        o3_update(p, -1);
        return;
    }
    if (f[ch] > 0) {
        emit(c, r, o2_cum(f, ch) - (ch >= pred ? pf : 0), f[ch], sum, 0, o);
        if (!o2_update(f, ch, 1) && f[ch] == 2) o2_update(f, 257, -1);
    } else {
        emit(c, r, body + f[256] - pf, f[257], sum, 0, o);
        int rescaled = o2_update(f, 257, 1);
        if (o1[ch] > 0) {
            uint32_t cum1 = 0, sum1 = 0;
            for (int i = 0; i < 256; i++)
                if (f[i] == 0 && i != pred) { if (i < ch) cum1 += O1F(o1[i]); sum1 += O1F(o1[i]); }
            emit(c, r, cum1, O1F(o1[ch]), sum1, 0, o);
            o1_update(o1, ch);
        }
        if (!rescaled) o2_update(f, ch, 1);
    }
    o3_update(p, ch);
}

/* src/cr-ppm.c:169-235 */
static int ppm_decode(ppm_t* p, rc_t* r, const uint8_t** in) {
    uint8_t* f = p->o2[p->ctx & 0xffff];
    uint8_t* o1 = p->o1[p->ctx & 0xff];
    int pred = p->o3_byte[o3_slot(p->ctx)];
    uint32_t pf = f[pred], body = o2_body(f);
    uint32_t target = rc_dec_cum(r, body + f[256] + f[257] - pf);
    /* symbol search with `pred` excluded (src/cr-o2model.c:93-113) */
    uint32_t acc = 0; int s = 0;
    for (;; s++) { uint32_t w = (s == pred) ? 0 : f[s]; if (acc + w > target) break; acc += w; }
    rc_dec_consume(r, acc, f[s], in);
    int rescaled = o2_update(f, s, 1);
    if (s == 256) { o3_update(p, -1); return pred; }
    if (s < 256) {
        if (!rescaled && f[s] == 2) o2_update(f, 257, -1);
        o3_update(p, s);
        return s;
    }
    uint32_t sum1 = 0, cum1 = 0;
    for (int i = 0; i < 256; i++) if (f[i] == 0 && i != pred) sum1 += O1F(o1[i]);
    target = rc_dec_cum(r, sum1);
    int d = 257;
    for (int i = 0; i < 256; i++)
        if (f[i] == 0 && i != pred) { if (cum1 + O1F(o1[i]) > target) { d = i; break; } cum1 += O1F(o1[i]); }
    rc_dec_consume(r, cum1, O1F(o1[d]), in);
    o1_update(o1, d);
    if (!rescaled) o2_update(f, d, 1);
    o3_update(p, d);
    return d;
}

/* ------------------------------------------------------------------ filter state (function-local statics in the ref) */
typedef struct { int flag; uint32_t curr, imsz; } x86_state;
typedef struct { int flag, curr, size, row_size, bpp, width, height, skip_size; } bmp_state;

/* ------------------------------------------------------------------ dictionary (src/cr-diccode.c) */
#define DIC_MAXWORDS 25000
#define DIC_WORDBUF  22
#define L1_WORDS(n)  ((65535 - (n)) / 255 - 1)
typedef struct {
    char     (*word)[DIC_WORDBUF];
    int      nentries;          /* the reference's dic_len */
    int32_t* next;              /* nnode x 128 */
    int32_t* id;                /* nnode */
    uint32_t nnode, nword, cap;
} dict_t;

/* ------------------------------------------------------------------ context */
struct cro_ctx {
    int variant, flexible, tracing;
    ppm_t ppm;
    m0_t idx_model, len_model;
    m0_t x_len, x_spos, x_pos[6];   /* LZ77 (comprox): src/roxmain/cr-coder.c:55-60 */
    uint32_t match_limit;           /* -m, src/roxmain/cr-matcher.c:38 */
    dict_t dic;
    int last_filter;            /* 0 none, 1 pe, 2 elf, 3 bmp  (src/cr-filter.c:41) */
    x86_state pe, elf;
    bmp_state bmp;
    vec_t tokens, events, triples;
};

static void trace_event(cro_ctx* c, uint32_t ctx, uint32_t sym) {
    if (c && c->tracing) { cro_event* e = vec_grow(&c->events, sizeof *e); e->ctx = ctx; e->sym = sym; }
}
static void trace_triple(cro_ctx* c, uint32_t cum, uint32_t frq, uint32_t sum, uint32_t stream) {
    if (c && c->tracing) { cro_triple* t = vec_grow(&c->triples, sizeof *t); t->cum = cum; t->frq = frq; t->sum = sum; t->stream = stream; }
}
static void trace_token(cro_ctx* c, uint32_t pos, uint32_t len, uint32_t idx) {
    if (c && c->tracing) { cro_token* t = vec_grow(&c->tokens, sizeof *t); t->pos = pos; t->len = len; t->idx = idx; }
}

cro_ctx* cro_new(int variant) {
    cro_ctx* c = calloc(1, sizeof *c);
    c->variant = variant;
    c->match_limit = 40;
    ppm_alloc(&c->ppm);
    c->dic.word = calloc(DIC_MAXWORDS, DIC_WORDBUF);
    cro_reset_models(c);
    return c;
}
void cro_free(cro_ctx* c) {
    if (!c) return;
    ppm_release(&c->ppm);
    free(c->dic.word); free(c->dic.next); free(c->dic.id);
    free(c->tokens.p); free(c->events.p); free(c->triples.p);
    free(c);
}
/* src/rolzmain/cr-coder.c:78-96, src/ropmain/cr-coder.c:73-83 */
void cro_reset_models(cro_ctx* c) {
    ppm_reset(&c->ppm);
    for (int i = 0; i < 256; i++) {
        c->idx_model.frq[i] = i < 64 + 16;
        c->len_model.frq[i] = (i == 0 || i >= 5);
    }
    m0_recount(&c->idx_model); m0_recount(&c->len_model);
    /* src/roxmain/cr-coder.c:88-103 */
    for (int i = 0; i < 5; i++) {
        for (int k = 0; k < 256; k++) c->x_pos[i].frq[k] = (i == 0 && k % 8 == 0) || (i > 0 && (i < 2 || k < 128));
        m0_recount(&c->x_pos[i]);
    }
    for (int k = 0; k < 256; k++) { c->x_pos[5].frq[k] = 1; c->x_spos.frq[k] = 1; c->x_len.frq[k] = (k >= 6) || (k == 0); }
    m0_recount(&c->x_pos[5]); m0_recount(&c->x_spos); m0_recount(&c->x_len);
}
void cro_set_match_limit(cro_ctx* c, uint32_t limit) { c->match_limit = limit; }
void cro_set_flexible(cro_ctx* c, int on) { c->flexible = on; }
void cro_trace_enable(cro_ctx* c, int on) { c->tracing = on; }
void cro_trace_clear(cro_ctx* c) { c->tokens.n = c->events.n = c->triples.n = 0; }
size_t cro_trace_tokens(cro_ctx* c, const cro_token** o) { *o = c->tokens.p; return c->tokens.n; }
size_t cro_trace_events(cro_ctx* c, const cro_event** o) { *o = c->events.p; return c->events.n; }
size_t cro_trace_triples(cro_ctx* c, const cro_triple** o) { *o = c->triples.p; return c->triples.n; }

/* ================================================================== ROLZ (src/rolzmain) */
#define RZ_BUCKETS 262144u
#define RZ_WAYS    64u
#define RZ_SHORT   16u
#define RZ_MINLEN  5u
#define RZ_MAXLEN  255u
#define RZ_NONE    0xFFFFFFFFu

typedef struct {
    uint32_t* ring;             /* RZ_BUCKETS x RZ_WAYS, RZ_NONE = empty */
    uint8_t*  head;             /* RZ_BUCKETS */
    uint32_t  short_ring[256][RZ_SHORT];
    uint32_t  bucket;           /* m_context */
    uint8_t   short_bucket;
    int       ctx4;
} rz_t;

static uint32_t rz_hash(const uint8_t* x, int ctx4) {                                /* cr-matcher.c:38-42 */
    uint32_t h = x[0] * 1313131u + x[-1] * 13131u + x[-2] * 131u;
    if (ctx4) h += x[-3];
    return h % RZ_BUCKETS;
}
static void rz_init(rz_t* m, int ctx4) {                                             /* cr-matcher.c:44-58 */
    m->ring = malloc((size_t)RZ_BUCKETS * RZ_WAYS * 4);
    m->head = calloc(RZ_BUCKETS, 1);
    memset(m->ring, 0xFF, (size_t)RZ_BUCKETS * RZ_WAYS * 4);
    memset(m->short_ring, 0, sizeof m->short_ring);
    m->bucket = 0; m->short_bucket = 0; m->ctx4 = ctx4;
}
static void rz_free(rz_t* m) { free(m->ring); free(m->head); }
/* n-th most recent entry of a bucket */
static uint32_t rz_item(const rz_t* m, uint32_t b, uint32_t n) { return m->ring[(size_t)b * RZ_WAYS + ((m->head[b] + RZ_WAYS - n) % RZ_WAYS)]; }
static void rz_insert(rz_t* m, const uint8_t* d, uint32_t pos) {                      /* cr-matcher.c:65-81 */
    if (pos < 16) return;
    m->head[m->bucket] = (m->head[m->bucket] + 1) % RZ_WAYS;
    m->ring[(size_t)m->bucket * RZ_WAYS + m->head[m->bucket]] = pos;
    m->bucket = rz_hash(d + pos, m->ctx4);
    uint32_t* s = m->short_ring[m->short_bucket];
    memmove(s + 1, s, (RZ_SHORT - 1) * 4);
    s[0] = pos;
    m->short_bucket = d[pos];
}
static uint32_t rz_getpos(const rz_t* m, uint32_t idx) {                             /* cr-matcher.c:83-88 */
    return idx < RZ_WAYS ? rz_item(m, m->bucket, idx) : m->short_ring[m->short_bucket][idx - RZ_WAYS];
}
static uint32_t common_prefix(const uint8_t* a, const uint8_t* b, uint32_t cap) { uint32_t j = 0; while (j < cap && a[j] == b[j]) j++; return j; }

typedef struct { uint32_t idx, len; } rz_match;
/* cr-matcher.c:90-120: most-recent-first, strictly longer wins, stop at 255 */
static rz_match rz_best(const rz_t* m, const uint8_t* d, uint32_t pos, uint32_t bucket) {
    rz_match r = { RZ_NONE, RZ_MINLEN - 1 };
    for (uint32_t i = 0; i < RZ_WAYS && r.len < RZ_MAXLEN; i++) {
        uint32_t q = rz_item(m, bucket, i);
        if (q == RZ_NONE) break;
        if (d[q] != d[pos]) continue;
        uint32_t l = common_prefix(d + pos, d + q, RZ_MAXLEN);
        if (l > r.len) { r.idx = i; r.len = l; }
    }
    if (r.len < RZ_MINLEN) { r.idx = RZ_NONE; r.len = 1; }
    return r;
}
static uint32_t rz_price(rz_match r) {                                               /* cr-matcher.c:146-148 */
    return r.len >= RZ_MINLEN ? (r.len - 1) * 3 * RZ_WAYS - 3 * r.idx : 9 * RZ_WAYS;
}
/* cr-matcher.c:122-197 */
static rz_match rz_lookup(const rz_t* m, const uint8_t* d, uint32_t pos, int flexible) {
    rz_match r = { RZ_NONE, 1 };
    if (pos < 16) return r;
    r = rz_best(m, d, pos, m->bucket);
    int find_short = r.len < RZ_MINLEN;
    if (flexible && !find_short) {
        uint32_t prices[260], best;
        for (uint32_t i = 1; i <= r.len; i++) prices[i] = rz_price(rz_best(m, d, pos + i, rz_hash(d + pos + i - 1, m->ctx4)));
        best = rz_price(r) + prices[r.len];
        for (uint32_t i = r.len - 1; i >= 1; i--) {
            rz_match t = { r.idx, i };
            if (rz_price(t) + prices[i] > best) { r.len = i; best = rz_price(t) + prices[i]; }
        }
    }
    if (find_short) {
        r.len = RZ_MINLEN - 1; r.idx = RZ_NONE;
        for (uint32_t i = 0; i < RZ_SHORT; i++) {
            uint32_t l = common_prefix(d + pos, d + m->short_ring[m->short_bucket][i], RZ_MAXLEN);
            if (l > r.len) { r.idx = RZ_WAYS + i; r.len = l; }
        }
    }
    if (r.len < RZ_MINLEN) { r.idx = RZ_NONE; r.len = 1; }
    if ((!flexible || find_short) && r.len > 1) {
        for (uint32_t i = 1; i < RZ_MINLEN; i++) {
            rz_match t = rz_best(m, d, pos + i, rz_hash(d + pos + i - 1, m->ctx4));
            if (rz_price(t) > rz_price(r) + i * RZ_WAYS) { r.idx = RZ_NONE; r.len = 1; break; }
        }
    }
    return r;
}

/* token list of one block: the look-ahead thread of src/rolzmain/cr-coder.c:109-137 run to completion */
size_t cro_rolz_parse(const uint8_t* d, uint32_t n, int flexible, cro_token** out) {
    vec_t v = {0};
    rz_t m; rz_init(&m, n >= 4194304);
    for (uint32_t pos = 1; pos < n;) {
        rz_match r = { RZ_NONE, 1 };
        if (pos + 1024 < n) r = rz_lookup(&m, d, pos, flexible);
        for (uint32_t i = 0; i < r.len; i++) rz_insert(&m, d, pos + i);
        cro_token* t = vec_grow(&v, sizeof *t); t->pos = pos; t->len = r.len; t->idx = r.idx;
        pos += r.len;
    }
    rz_free(&m);
    *out = v.p;
    return v.n;
}

static uint8_t rarest_byte(const uint8_t* d, uint32_t n) {                           /* cr-coder.c:171-178 */
    uint32_t cnt[256] = {0}; int e = 0;
    for (uint32_t i = 0; i < n; i++) cnt[d[i]]++;
    for (int i = 1; i < 256; i++) if (cnt[e] > cnt[i]) e = i;
    return (uint8_t)e;
}
static void m0_code(cro_ctx* c, rc_t* r, m0_t* m, int s, cro_buf* o) {               /* M_my_enc_, cr-model.h:58-64 */
    emit(c, r, m0_cum(m, s), m->frq[s], m->total, 1, o);
    m0_update(m, s, 4);
}
static void store_raw(const uint8_t* d, uint32_t n, uint32_t hdr, cro_buf* out) {
    out->size = 0; buf_reserve(out, hdr + n);
    memset(out->data, 0, hdr); out->size = hdr; buf_append(out, d, n);
}

/* src/rolzmain/cr-coder.c:139-264 */
static void rolz_encode(cro_ctx* c, const uint8_t* d, uint32_t n, cro_buf* out) {
    cro_token* tok; size_t nt = cro_rolz_parse(d, n, c->flexible, &tok);
    cro_buf side = {0};
    rc_t main_rc, side_rc; rc_enc_init(&main_rc); rc_enc_init(&side_rc);
    uint8_t esc = rarest_byte(d, n);
    uint32_t nidx = 0; int aborted = 0;
    out->size = 0; buf_reserve(out, 16); memset(out->data, 0, 16); out->size = 16;
    for (size_t k = 0; k < nt; k++) {
        uint32_t pos = tok[k].pos;
        trace_token(c, pos, tok[k].len, tok[k].idx);
        if (tok[k].idx != RZ_NONE) {
            ppm_encode(c, &c->ppm, &main_rc, esc, out);
            m0_code(c, &side_rc, &c->len_model, tok[k].len, &side);
            m0_code(c, &side_rc, &c->idx_model, tok[k].idx, &side);
            nidx++;
        } else {
            ppm_encode(c, &c->ppm, &main_rc, d[pos], out);
            if (d[pos] == esc) { m0_code(c, &side_rc, &c->len_model, 0, &side); nidx++; }
        }
        for (uint32_t i = 0; i < tok[k].len; i++) c->ppm.ctx = c->ppm.ctx << 8 | d[pos + i];
        if (out->size >= n) { aborted = 1; break; }                                  /* :231-233 */
    }
    free(tok);
    if (aborted) { store_raw(d, n, 16, out); cro_buf_free(&side); return; }
    rc_flush(&main_rc, out); rc_flush(&side_rc, &side);
    out->data[0] = n ? d[0] : 0; out->data[1] = 1; out->data[2] = esc; out->data[3] = 0;
    wr32(out->data + 4, n); wr32(out->data + 8, nidx); wr32(out->data + 12, (uint32_t)out->size);
    buf_append(out, side.data, side.size);
    cro_buf_free(&side);
}

static int m0_decode(rc_t* r, m0_t* m, const uint8_t** in) {                         /* M_my_dec_, cr-model.h:66-74 */
    uint32_t target = rc_dec_cum(r, m->total), acc = 0; int s = 0;
    while (acc + m->frq[s] <= target) acc += m->frq[s++];
    rc_dec_consume(r, acc, m->frq[s], in);
    m0_update(m, s, 4);
    return s;
}
/* src/rolzmain/cr-coder.c:287-379 (the side-stream thread only pre-decodes; order of model use is unchanged) */
static void rolz_decode(cro_ctx* c, const uint8_t* in, uint32_t n, cro_buf* out) {
    if (!in[1]) { buf_append(out, in + 16, n - 16); return; }
    uint8_t esc = in[2]; uint32_t orig = rd32(in + 4), off_idx = rd32(in + 12);
    const uint8_t* pm = in + 16; const uint8_t* ps = in + off_idx;
    rc_t main_rc, side_rc; rc_dec_init(&main_rc, &pm); rc_dec_init(&side_rc, &ps);
    size_t base = out->size;
    buf_reserve(out, base + orig + 300); buf_push(out, in[0]);
    rz_t m; rz_init(&m, orig >= 4194304);
#define OUTD (out->data + base)
    while (out->size - base < orig) {
        uint32_t len = 1;
        int s = ppm_decode(&c->ppm, &main_rc, &pm);
        if (s == esc) {
            int l = m0_decode(&side_rc, &c->len_model, &ps);
            if (l == 0) buf_push(out, esc);
            else {
                int idx = m0_decode(&side_rc, &c->idx_model, &ps);
                uint32_t q = rz_getpos(&m, idx);
                len = l;
                buf_reserve(out, out->size + len);
                for (uint32_t i = 0; i < len; i++) { uint8_t b = OUTD[q + i]; out->data[out->size++] = b; }
            }
        } else buf_push(out, (uint8_t)s);
        for (; len; len--) {
            uint32_t p = (uint32_t)(out->size - base) - len;
            rz_insert(&m, OUTD, p);
            c->ppm.ctx = c->ppm.ctx << 8 | OUTD[p];
        }
    }
#undef OUTD
    rz_free(&m);
}

/* ================================================================== LZP (src/ropmain) */
typedef struct { uint32_t *t8, *t4, *t2; } lzp_t;
static uint32_t lzp_h2(const uint8_t* x) { return rd16(x); }
static uint32_t lzp_h4(const uint8_t* x) { uint32_t v = rd32(x); return (v ^ v >> 6 ^ v >> 12) & 0xfffff; }
static uint32_t lzp_h8(const uint8_t* x) { uint64_t v = rd32(x) | (uint64_t)rd32(x + 4) << 32; return (uint32_t)((v ^ v >> 20 ^ v >> 40) & 0xffffff); }
static void lzp_init(lzp_t* m) {                                                     /* cr-matcher.c:35-50 */
    m->t8 = malloc(4u << 24); m->t4 = malloc(4u << 20); m->t2 = malloc(4u << 16);
    for (uint32_t i = 0; i < 1u << 24; i++) m->t8[i] = 8;
    for (uint32_t i = 0; i < 1u << 20; i++) m->t4[i] = 4;
    for (uint32_t i = 0; i < 1u << 16; i++) m->t2[i] = 2;
}
static void lzp_free(lzp_t* m) { free(m->t8); free(m->t4); free(m->t2); }
static uint32_t lzp_getpos(const lzp_t* m, const uint8_t* d, uint32_t pos) {         /* cr-matcher.c:59-73 */
    uint32_t a = m->t8[lzp_h8(d + pos - 8)], b = m->t4[lzp_h4(d + pos - 4)], c2 = m->t2[lzp_h2(d + pos - 2)];
    if (memcmp(d + a - 8, d + pos - 8, 8) == 0) return a;
    if (memcmp(d + b - 4, d + pos - 4, 4) == 0) return b;
    return c2;
}
static uint32_t lzp_lookup(const lzp_t* m, const uint8_t* d, uint32_t pos) {         /* cr-matcher.c:75-89 */
    uint32_t q = lzp_getpos(m, d, pos), l = 0;
    if (q != 0) l = common_prefix(d + q, d + pos, 255);
    return l < 4 ? 1 : l;
}
static void lzp_insert(lzp_t* m, const uint8_t* d, uint32_t pos) {                   /* cr-matcher.c:91-96 */
    m->t8[lzp_h8(d + pos - 8)] = pos; m->t4[lzp_h4(d + pos - 4)] = pos; m->t2[lzp_h2(d + pos - 2)] = pos;
}
size_t cro_lzp_parse(const uint8_t* d, uint32_t n, cro_token** out) {                /* cr-coder.c:95-118 */
    vec_t v = {0};
    if (n >= 16) {
        lzp_t m; lzp_init(&m);
        for (uint32_t pos = 9; pos < n;) {
            uint32_t len = 1;
            if (pos + 1024 < n) { len = lzp_lookup(&m, d, pos); for (uint32_t i = 0; i < len; i++) lzp_insert(&m, d, pos + i); }
            cro_token* t = vec_grow(&v, sizeof *t); t->pos = pos; t->len = len; t->idx = len > 1 ? 0 : RZ_NONE;
            pos += len;
        }
        lzp_free(&m);
    }
    *out = v.p;
    return v.n;
}
/* src/ropmain/cr-coder.c:119-229 */
static void lzp_encode(cro_ctx* c, const uint8_t* d, uint32_t n, cro_buf* out) {
    if (n < 16) { store_raw(d, n, 20, out); return; }
    cro_token* tok; size_t nt = cro_lzp_parse(d, n, &tok);
    uint8_t esc = rarest_byte(d, n);
    rc_t r; rc_enc_init(&r);
    int aborted = 0;
    out->size = 0; buf_reserve(out, 20); memset(out->data, 0, 20); out->size = 20;
    for (size_t k = 0; k < nt; k++) {
        uint32_t pos = tok[k].pos, len = tok[k].len;
        trace_token(c, pos, len, tok[k].idx);
        if (len > 1) {
            ppm_encode(c, &c->ppm, &r, esc, out);
            c->ppm.ctx = c->ppm.ctx << 8 | esc;
            ppm_encode(c, &c->ppm, &r, (int)len, out);
        } else {
            ppm_encode(c, &c->ppm, &r, d[pos], out);
            if (d[pos] == esc) { c->ppm.ctx = c->ppm.ctx << 8 | esc; ppm_encode(c, &c->ppm, &r, 0, out); }
        }
        for (uint32_t i = 0; i < len; i++) c->ppm.ctx = c->ppm.ctx << 8 | d[pos + i];
        if (out->size >= n) { aborted = 1; break; }
    }
    free(tok);
    if (aborted) { store_raw(d, n, 20, out); return; }
    rc_flush(&r, out);
    out->data[0] = 1; wr32(out->data + 4, n); out->data[8] = esc; memcpy(out->data + 9, d, 9);
}
/* src/ropmain/cr-coder.c:231-292 */
static void lzp_decode(cro_ctx* c, const uint8_t* in, uint32_t n, cro_buf* out) {
    if (!in[0]) { buf_append(out, in + 20, n - 20); return; }
    uint32_t orig = rd32(in + 4); uint8_t esc = in[8];
    size_t base = out->size;
    buf_reserve(out, base + orig + 300); buf_append(out, in + 9, 9);
    const uint8_t* p = in + 20;
    rc_t r; rc_dec_init(&r, &p);
    lzp_t m; lzp_init(&m);
#define OUTD (out->data + base)
    while (out->size - base < orig) {
        uint32_t len = 1;
        int s = ppm_decode(&c->ppm, &r, &p);
        if (s != esc) buf_push(out, (uint8_t)s);
        else {
            c->ppm.ctx = c->ppm.ctx << 8 | esc;
            len = ppm_decode(&c->ppm, &r, &p);
            if (len == 0) { len = 1; buf_push(out, esc); }
            else {
                uint32_t q = lzp_getpos(&m, OUTD, (uint32_t)(out->size - base));
                buf_reserve(out, out->size + len);
                for (uint32_t i = 0; i < len; i++) { uint8_t b = OUTD[q + i]; out->data[out->size++] = b; }
            }
        }
        for (; len; len--) {
            uint32_t q = (uint32_t)(out->size - base) - len;
            c->ppm.ctx = c->ppm.ctx << 8 | OUTD[q];
            lzp_insert(&m, OUTD, q);
        }
    }
#undef OUTD
    lzp_free(&m);
}

/* ================================================================== LZ77 (src/roxmain, the `comprox` binary) */
#define X_MAXLEN   255u
#define X_MIN_NEAR 6u
#define X_NONE     0xFFFFFFFFu
typedef struct {
    uint32_t* next;             /* previous position with the same (hash1 % 20, hash2 % bucketsize2), X_NONE = end */
    uint32_t* short_cache;      /* [65536] most recent position per 6-byte hash */
    uint32_t  last_match;       /* distance of the most recent match (the one parse-dependent scalar) */
    uint32_t  match_min, limit;
} x_t;
typedef struct { uint32_t pos, len; } x_ret;

static uint32_t x_hashn(const uint8_t* s, uint32_t k) { uint32_t h = 0; for (uint32_t i = 0; i < k; i++) h = (h * 123456791u) ^ s[i]; return h; }   /* cr-matcher.c:44-52,196-204 */
/* matcher_init, cr-matcher.c:89-147.  The reference threads positions through two bucket passes; the result is:
 * next[p] = the largest q < p with the same (hash1 % 20, hash2 % bucketsize2), for p <= len - 256. */
static void x_init(x_t* m, const uint8_t* d, uint32_t len, uint32_t match_min, uint32_t limit) {
    const uint32_t b2 = 20 + len / 25;
    m->next = malloc(((size_t)len + 1) * 4); m->short_cache = calloc(65536, 4);
    m->last_match = 0; m->match_min = match_min; m->limit = limit;
    memset(m->next, 0xFF, ((size_t)len + 1) * 4);
    if (len < X_MAXLEN + 1) return;
    uint32_t* head = malloc((size_t)20 * b2 * 4);
    memset(head, 0xFF, (size_t)20 * b2 * 4);
    for (uint32_t p = 0; p + X_MAXLEN < len; p++) {
        size_t k = (size_t)((d[p] + d[p + 1]) % 20) * b2 + x_hashn(d + p, match_min) % b2;
        m->next[p] = head[k]; head[k] = p;
    }
    free(head);
}
static void x_free(x_t* m) { free(m->next); free(m->short_cache); }
/* match(), cr-matcher.c:156-194 */
static x_ret x_match(const x_t* m, const uint8_t* d, uint32_t pos, uint32_t min, uint32_t limit, uint32_t lazy) {
    x_ret r = { 0, min - 1 };
    uint32_t node = m->next[pos];
    for (uint32_t i = 0; i < limit && node != X_NONE; i++) {
        uint32_t nl = r.len;
        while (nl < X_MAXLEN && d[node + nl] == d[pos + nl]) nl++;
        uint32_t price = 0, dist = pos - node, best = pos - r.pos;
        price += dist / 1048576 > best; price += dist / 4096 > best; price += dist / 64 > best;
        if (nl > r.len + price && memcmp(d + pos, d + node, r.len) == 0) {
            r.pos = node; r.len = nl;
            if ((lazy && lazy < r.pos) || r.len == X_MAXLEN) return r;
        }
        node = m->next[node];
    }
    if (r.len < min) { r.pos = X_NONE; r.len = 1; }
    return r;
}
static int x_log2(uint32_t x) { int l = -1; while (x) { l++; x >>= 1; } return l; }   /* fast_log2, cr-matcher.c:211-228: floor(log2 x), 0 -> -1 */
/* matcher_lookup, cr-matcher.c:230-340 */
static x_ret x_lookup(x_t* m, const uint8_t* d, uint32_t pos, int flexible) {
    x_ret t1 = { pos - m->last_match, 0 }, ret;
    const uint32_t mm = m->match_min;
    if (t1.pos < pos) while (t1.len < X_MAXLEN && d[pos + t1.len] == d[t1.pos + t1.len]) t1.len++;
    if (flexible) {
        /* the reference caches match() results between calls (m_ret_cache); match() is pure, so we recompute */
        x_ret r0 = x_match(m, d, pos, mm, m->limit, 0);
        ret = r0;
        if (r0.len >= mm) {
#define XP(i, l) ((l) >= mm ? (int)(((l) - 1) * 3) - (x_log2(pos - (i)) * 4 / 5) : 9)
            uint32_t n = r0.len;
            x_ret* rs = malloc((n + 1) * sizeof *rs);
            rs[0] = r0;
            for (uint32_t i = 1; i <= n; i++) rs[i] = x_match(m, d, pos + i, mm, m->limit, 0);
            uint32_t maxprice = (uint32_t)(XP(rs[0].pos, rs[0].len) + XP(rs[n].pos, rs[n].len));
            for (uint32_t i = n - 1; i >= 1; i--) {
                uint32_t pr = (uint32_t)(XP(ret.pos, i) + XP(rs[i].pos, rs[i].len));
                if (maxprice < pr) { ret.len = i; maxprice = pr; }
            }
            free(rs);
#undef XP
            if (ret.len < mm) { ret.pos = X_NONE; ret.len = 1; }
        }
    } else {
        ret = x_match(m, d, pos, mm, m->limit, 0);
        if (ret.len >= mm) {
            x_ret t2 = x_match(m, d, pos + 1, ret.len + 1, m->limit / 4, 1);
            if (t2.len > ret.len + (t2.pos < ret.pos) ||
                x_match(m, d, pos + 2, ret.len + 1, m->limit / 8, 1).len > 1 ||
                x_match(m, d, pos + 3, ret.len + 2, m->limit / 8, 1).len > 1 ||
                x_match(m, d, pos + 4, ret.len + 2, m->limit / 8, 1).len > 1 ||
                x_match(m, d, pos + 5, ret.len + 2, m->limit / 8, 1).len > 1 ||
                x_match(m, d, pos + 6, ret.len + 3, m->limit / 8, 1).len > 1) { ret.pos = X_NONE; ret.len = 1; }
        }
    }
    if (ret.pos != X_NONE && ret.len < t1.len + 3 + (ret.pos + 64 < pos) + (ret.pos + 4096 < pos) + (ret.pos + 1048576 < pos)) ret = t1;
    if (ret.len < X_MIN_NEAR) {
        ret.pos = m->short_cache[x_hashn(d + pos, X_MIN_NEAR) % 65536]; ret.len = 0;
        if (ret.pos < pos && ret.pos + 256 > pos) { uint32_t i = 0; while (i < X_MAXLEN && d[ret.pos + i] == d[pos + i]) i++; ret.len = i; }
    }
    if (ret.len < X_MIN_NEAR || (ret.len < mm && ret.pos + 256 <= pos)) { ret.pos = X_NONE; ret.len = 1; }
    else m->last_match = pos - ret.pos;
    return ret;
}
/* lzmatch_thread, cr-coder.c:116-142: the token list (the helper thread only runs ahead; the order of lookups is the serial one) */
size_t cro_lz77_parse(const uint8_t* d, uint32_t n, int flexible, uint32_t match_limit, cro_token** out) {
    vec_t v = {0};
    x_t m; x_init(&m, d, n, 10 + (n > 16777216), match_limit);
    for (uint32_t pos = 0; pos < n;) {
        x_ret r = { X_NONE, 1 };
        if (pos + 1024 < n) {
            r = x_lookup(&m, d, pos, flexible);
            for (uint32_t i = 0; i < r.len; i++) m.short_cache[x_hashn(d + pos + i, X_MIN_NEAR) % 65536] = pos + i;
        }
        cro_token* t = vec_grow(&v, sizeof *t); t->pos = pos; t->len = r.len; t->idx = r.pos;
        pos += r.len;
    }
    x_free(&m);
    *out = v.p;
    return v.n;
}
static void x_code(cro_ctx* c, rc_t* r, m0_t* m, int s, int inc, uint32_t stream, cro_buf* o) {   /* M_my_enc_, cr-model.h:58-64 */
    emit(c, r, m0_cum(m, s), m->frq[s], m->total, stream, o);
    m0_update(m, s, inc);
}
#define X_INC(i) (1 << (i) << (i))
/* lzencode, src/roxmain/cr-coder.c:144-318.  Streams: 0 = ppm, 1 = spos, 2 = pos, 3 = len. */
static void lz77_encode(cro_ctx* c, const uint8_t* d, uint32_t n, cro_buf* out) {
    cro_token* tok; size_t nt = cro_lz77_parse(d, n, c->flexible, c->match_limit, &tok);
    cro_buf spos = {0}, posb = {0}, lenb = {0};
    rc_t rc, rc_spos, rc_pos, rc_len; rc_enc_init(&rc); rc_enc_init(&rc_spos); rc_enc_init(&rc_pos); rc_enc_init(&rc_len);
    const uint8_t esc = rarest_byte(d, n);
    const uint32_t mm = 10 + (n > 16777216);
    uint32_t n_spos = 0, n_pos = 0, n_len = 0, last = 0; int aborted = 0;
    out->size = 0; buf_reserve(out, 32); memset(out->data, 0, 32); out->size = 32;
    for (size_t k = 0; k < nt; k++) {
        const uint32_t pos = tok[k].pos, len = tok[k].len; uint32_t mp = tok[k].idx;
        trace_token(c, pos, len, mp);
        if (mp != X_NONE) {
            ppm_encode(c, &c->ppm, &rc, esc, out);
            if (pos - mp == last) mp = pos;
            x_code(c, &rc_len, &c->x_len, (int)len, 30, 3, &lenb); n_len++;
            if (len < mm) { x_code(c, &rc_spos, &c->x_spos, (int)(pos - mp), 1, 1, &spos); n_spos++; }
            else {
                uint32_t j = (pos - mp) * 8; int i = 0;
                while (j >= 128 && i < 2) { x_code(c, &rc_pos, &c->x_pos[i], (int)(j % 128 + 128), X_INC(i), 2, &posb); i++; j /= 128; }
                if (i >= 2) while (j >= 64 && i < 5) { x_code(c, &rc_pos, &c->x_pos[i], (int)(j % 64 + 64), X_INC(i), 2, &posb); i++; j /= 64; }
                x_code(c, &rc_pos, &c->x_pos[i], (int)j, X_INC(i), 2, &posb); n_pos++;
            }
            last = pos - mp;
        } else {
            ppm_encode(c, &c->ppm, &rc, d[pos], out);
            if (d[pos] == esc) { x_code(c, &rc_len, &c->x_len, 0, 30, 3, &lenb); n_len++; }
        }
        for (uint32_t i = 0; i < len; i++) c->ppm.ctx = c->ppm.ctx << 8 | d[pos + i];
        if (out->size >= n) { aborted = 1; break; }
    }
    free(tok);
    if (aborted) { store_raw(d, n, 32, out); cro_buf_free(&spos); cro_buf_free(&posb); cro_buf_free(&lenb); return; }
    rc_flush(&rc, out); rc_flush(&rc_spos, &spos); rc_flush(&rc_pos, &posb); rc_flush(&rc_len, &lenb);
    /* block_header, cr-coder.c:68-80: u8 compressed, u8 match_min, u8 esc, pad, then 7 x u32 */
    out->data[0] = 1; out->data[1] = (uint8_t)mm; out->data[2] = esc; out->data[3] = 0;
    wr32(out->data + 4, n); wr32(out->data + 8, n_spos); wr32(out->data + 12, n_pos); wr32(out->data + 16, n_len);
    wr32(out->data + 20, (uint32_t)out->size); wr32(out->data + 24, (uint32_t)(out->size + spos.size)); wr32(out->data + 28, (uint32_t)(out->size + spos.size + posb.size));
    buf_append(out, spos.data, spos.size); buf_append(out, posb.data, posb.size); buf_append(out, lenb.data, lenb.size);
    cro_buf_free(&spos); cro_buf_free(&posb); cro_buf_free(&lenb);
}
static int x_decode_sym(rc_t* r, m0_t* m, int inc, const uint8_t** in) {               /* M_my_dec_, cr-model.h:66-74 */
    uint32_t target = rc_dec_cum(r, m->total), acc = 0; int s = 0;
    while (acc + m->frq[s] <= target) acc += m->frq[s++];
    rc_dec_consume(r, acc, m->frq[s], in);
    m0_update(m, s, inc);
    return s;
}
/* lzdecode, src/roxmain/cr-coder.c:388-526 (the queue threads only pre-decode; per-stream symbol order is unchanged) */
static void lz77_decode(cro_ctx* c, const uint8_t* in, uint32_t n, cro_buf* out) {
    if (!in[0]) { buf_append(out, in + 32, n - 32); return; }
    const uint32_t mm = in[1], esc = in[2], orig = rd32(in + 4);
    const uint8_t *pm = in + 32, *ps = in + rd32(in + 20), *pp = in + rd32(in + 24), *pl = in + rd32(in + 28);
    rc_t rc, rc_spos, rc_pos, rc_len; rc_dec_init(&rc, &pm); rc_dec_init(&rc_spos, &ps); rc_dec_init(&rc_pos, &pp); rc_dec_init(&rc_len, &pl);
    size_t base = out->size; uint32_t last = 0;
    buf_reserve(out, base + orig + 300);
    while (out->size - base < orig) {
        uint32_t len = 1, from = 0;
        int s = ppm_decode(&c->ppm, &rc, &pm);
        if ((uint32_t)s == esc) {
            uint32_t l = (uint32_t)x_decode_sym(&rc_len, &c->x_len, 30, &pl);
            if (l == 0) { s = (int)esc; }
            else {
                uint32_t dist;
                if (l < mm) dist = (uint32_t)x_decode_sym(&rc_spos, &c->x_spos, 1, &ps);
                else {
                    uint32_t j = 0, v = 0, sym = 0;
                    while (j < 2 && (sym = (uint32_t)x_decode_sym(&rc_pos, &c->x_pos[j], X_INC(j), &pp)) >= 128) { v += (sym - 128) * (1u << (7 * j)); j++; }
                    if (j < 2) dist = (v + sym * (1u << (7 * j))) / 8;
                    else {
                        while (j < 5 && (sym = (uint32_t)x_decode_sym(&rc_pos, &c->x_pos[j], X_INC(j), &pp)) >= 64) { v += (sym - 64) * (1u << (6 * j + 2)); j++; }
                        dist = (v + sym * (1u << (6 * j + 2))) / 8;
                    }
                }
                len = l;
                if (len > 1) { if (dist == 0) dist = last; from = (uint32_t)(out->size - base) - dist; last = dist; }
            }
        }
        buf_reserve(out, out->size + len);
        if (len > 1) for (uint32_t i = 0; i < len; i++) { uint8_t b = out->data[base + from + i]; out->data[out->size++] = b; }
        else out->data[out->size++] = (uint8_t)s;
        for (uint32_t i = 0; i < len; i++) c->ppm.ctx = c->ppm.ctx << 8 | out->data[out->size - len + i];
    }
}

void cro_lzencode(cro_ctx* c, const uint8_t* in, uint32_t n, cro_buf* out) {
    if (c->variant == CRO_ROLZ) rolz_encode(c, in, n, out); else if (c->variant == CRO_LZP) lzp_encode(c, in, n, out); else lz77_encode(c, in, n, out);
}
void cro_lzdecode(cro_ctx* c, const uint8_t* in, uint32_t n, cro_buf* out) {
    if (c->variant == CRO_ROLZ) rolz_decode(c, in, n, out); else if (c->variant == CRO_LZP) lzp_decode(c, in, n, out); else lz77_decode(c, in, n, out);
}

/* ================================================================== dicpick (src/cr-dicpick.c) */
#define WP_SLOTS   (1u << 21)
#define WP_MAXSIZE (25000 * 13 + 1)
typedef struct { char w[21]; int32_t count; } wp_entry;   /* count 0 = empty */

static uint32_t wp_hash(const char* w) { uint32_t h = 2166136261u; for (; *w; w++) h = (h ^ (uint8_t)*w) * 16777619u; return h; }
static wp_entry* wp_find(wp_entry* t, const char* w) {
    uint32_t i = wp_hash(w) & (WP_SLOTS - 1);
    while (t[i].count && strcmp(t[i].w, w)) i = (i + 1) & (WP_SLOTS - 1);
    return &t[i];
}
/* src/cr-dicpick.c:95-146. The reference's probe layout is irrelevant to the result; the prune is not:
 * when the 325001st distinct word arrives, every word with count <= min+5 is dropped (:115-144). */
static void wp_add(wp_entry** tp, uint32_t* size, const char* w) {
    wp_entry* e = wp_find(*tp, w);
    if (e->count) { e->count++; return; }
    strcpy(e->w, w); e->count = 1;
    if (++*size == WP_MAXSIZE) {
        wp_entry* old = *tp; wp_entry* nt = calloc(WP_SLOTS, sizeof *nt);
        int32_t mn = INT32_MAX;
        for (uint32_t i = 0; i < WP_SLOTS; i++) if (old[i].count && old[i].count < mn) mn = old[i].count;
        *size = 0;
        for (uint32_t i = 0; i < WP_SLOTS; i++) if (old[i].count > mn + 5) { *wp_find(nt, old[i].w) = old[i]; ++*size; }
        free(old); *tp = nt;
    }
}
static int cmp_word(const void* a, const void* b) { return strcmp(((const wp_entry*)a)->w, ((const wp_entry*)b)->w); }
static int cmp_count_desc(const void* a, const void* b) {                            /* :60-67 */
    const wp_entry *x = a, *y = b;
    if (x->count != y->count) return y->count - x->count;
    return strcmp(y->w, x->w);
}
void cro_dicpick(const uint8_t* data, size_t n, cro_buf* dic) {
    wp_entry* t = calloc(WP_SLOTS, sizeof *t); uint32_t size = 0;
    uint8_t* chunk = malloc(200000);
    for (size_t off = 0; off < n; off += 200000) {                                   /* :191-215 */
        int flen = (int)(n - off < 200000 ? n - off : 200000);
        memcpy(chunk, data + off, flen);
        chunk[flen - 1] = 0;
        for (int x = 1; x < flen; x++) {
            if (!is_alpha(chunk[x]) || is_alpha(chunk[x - 1])) continue;
            int y = x + 1;
            while (y < flen && is_lower(chunk[y])) y++;
            uint8_t s = chunk[y];
            if (y - x >= 2 && y - x <= 20 && (s == ' ' || s == ',' || s == '.' || s == ':' || s == ';')) {
                char w[21]; for (int i = x; i < y; i++) w[i - x] = (char)to_lower(chunk[i]); w[y - x] = 0;
                wp_add(&t, &size, w);
            }
            x = y;
        }
    }
    free(chunk);
    wp_entry* sel = malloc(sizeof *sel * (size + 1)); int y = 0;                     /* :218-236 */
    for (uint32_t i = 0; i < WP_SLOTS; i++) if (t[i].count > 5) sel[y++] = t[i];
    free(t);
    qsort(sel, y, sizeof *sel, cmp_count_desc);
    if (y > DIC_MAXWORDS - 2) y = DIC_MAXWORDS - 2;
    if (y > L1_WORDS(y) - 2) { int x = L1_WORDS(y) - 2; qsort(sel + x, y - x, sizeof *sel, cmp_word); }
    dic->size = 0;                                                                   /* :238-257 */
    buf_append(dic, "\x20\x20\n", 3); buf_append(dic, "http://www.\n", 12);
    for (int x = 0; x < y; x++)
        if (x < L1_WORDS(y) || strlen(sel[x].w) >= 3) { buf_append(dic, sel[x].w, strlen(sel[x].w)); buf_push(dic, '\n'); }
    buf_push(dic, 0);
    free(sel);
}
/* src/cr-dicpick.c:261-305: front coding against the previous line */
void cro_dic_lcp_encode(cro_buf* dic) {
    cro_buf o = {0}; const uint8_t* d = dic->data; size_t prev = 0, cur = 0;
    while (d[cur] != '\n') buf_push(&o, d[cur++]);
    cur++; buf_push(&o, '\n');
    while (d[cur] != 0) {
        int lcp = 0; while (d[prev + lcp] == d[cur + lcp]) lcp++;
        buf_push(&o, (uint8_t)lcp);
        prev = cur; cur += lcp;
        while (d[cur] != '\n') buf_push(&o, d[cur++]);
        cur++; buf_push(&o, '\n');
    }
    buf_push(&o, 255);
    cro_buf_free(dic); *dic = o;
}
/* src/cr-dicpick.c:307-346 */
void cro_dic_lcp_decode(cro_buf* dic) {
    cro_buf o = {0}; const uint8_t* d = dic->data; size_t wi = 0, wo = 0;
    while (d[wi] != '\n') buf_push(&o, d[wi++]);
    wi++; buf_push(&o, '\n');
    while (d[wi] != 255) {
        int lcp = d[wi++];
        while (lcp-- > 0) buf_push(&o, o.data[wo++]);
        while (d[wi] != '\n') buf_push(&o, d[wi++]);
        wi++; buf_push(&o, '\n');
        while (o.data[wo] != '\n') wo++;
        wo++;
    }
    buf_push(&o, 0);
    cro_buf_free(dic); *dic = o;
}

/* ================================================================== diccode (src/cr-diccode.c) */
static uint32_t trie_new_node(dict_t* t) {
    if (t->nnode >= t->cap) {
        t->cap = t->cap ? t->cap * 2 : 4096;
        t->next = realloc(t->next, (size_t)t->cap * 128 * 4);
        t->id = realloc(t->id, (size_t)t->cap * 4);
    }
    memset(t->next + (size_t)t->nnode * 128, 0, 128 * 4);
    t->id[t->nnode] = 0;
    return t->nnode++;
}
static void trie_add(dict_t* t, const char* w) {                                      /* :47-70 */
    uint32_t node = 0;
    for (; *w; w++) {
        uint8_t ch = (uint8_t)*w;
        if (t->next[(size_t)node * 128 + ch] == 0) {
            uint32_t nn = trie_new_node(t);
            t->id[node] = -1;
            t->next[(size_t)node * 128 + ch] = (int32_t)nn;
        }
        node = (uint32_t)t->next[(size_t)node * 128 + ch];
    }
    t->id[node] = (int32_t)t->nword++;
}
int cro_dictionary_load(cro_ctx* c, const char* s, int init_trie) {                   /* :76-120 */
    dict_t* t = &c->dic; int p = 0;
    memset(t->word, 0, DIC_MAXWORDS * DIC_WORDBUF); t->nentries = 0;
    for (; *s; s++) {
        char* w = t->word[t->nentries];
        if (*s == '\n') {
            if (p > 0 && is_alpha((uint8_t)w[p - 1])) { w[p++] = ' '; w[p++] = 0; }
            p = 0; t->nentries++;
        } else w[p++] = *s;
    }
    t->nnode = 0; t->nword = 0;
    if (init_trie) {
        trie_new_node(t);
        for (int i = 0; i < t->nentries; i++) trie_add(t, t->word[i]);
        for (int i = 'A'; i < 'Z'; i++) t->next[i] = t->next[to_lower(i)];            /* sic: 'Z' is left out, :107 */
        for (uint32_t i = 0; i < t->nnode; i++) {
            int32_t* nx = t->next + (size_t)i * 128;
            if (nx[' '] > 0) { if (!nx['.']) nx['.'] = nx[' ']; if (!nx[',']) nx[','] = nx[' ']; if (!nx[':']) nx[':'] = nx[' ']; if (!nx[';']) nx[';'] = nx[' ']; }
        }
    }
    return (int)t->nword;
}
static int sentence_start(const uint8_t* s, int i) {                                  /* M_check_reverse_case, :313 */
    return i >= 3 && s[i - 1] == ' ' && (s[i - 2] == '.' || (s[i - 2] == ' ' && s[i - 3] == '.'));
}
static void put_literal(const dict_t* t, const uint8_t* escmap, uint8_t b, cro_buf* o) {
    int L1 = L1_WORDS(t->nentries);
    if (escmap[b]) { buf_push(o, (uint8_t)(t->nentries / (256 - L1))); buf_push(o, (uint8_t)(t->nentries % (256 - L1) + L1)); }
    buf_push(o, b);
}
/* src/cr-diccode.c:285-362 */
static void dic_encode_sub(const dict_t* t, const uint8_t* d, uint32_t size, const uint8_t esc[10], cro_buf* o) {
    uint8_t escmap[256] = {0};
    int L1 = L1_WORDS(t->nentries);
    uint32_t i = 0;
    for (int k = 0; k < 10; k++) escmap[esc[k]] = (uint8_t)(k + 1);
    for (; i + 40 < size; i++) {
        uint32_t j = i, node = 0;
        if (i > 0 && is_alpha(d[i]) && !is_alpha(d[i - 1])) {
            while (d[j] < 128 && (node = (uint32_t)t->next[(size_t)node * 128 + d[j]]) != 0 && t->id[node] == -1) j++;
        }
        if (d[j] < 128 && node != 0) {
            int rev = is_upper(d[i]) ^ sentence_start(d, (int)i);
            int tail = d[j] == ':' ? 4 : d[j] == ';' ? 3 : d[j] == ',' ? 2 : d[j] == '.' ? 1 : 0;
            int id = t->id[node];
            if (id < L1) buf_push(o, (uint8_t)id);
            else { buf_push(o, (uint8_t)(id / (256 - L1))); buf_push(o, (uint8_t)(id % (256 - L1) + L1)); }
            buf_push(o, esc[rev * 5 + tail]);
            i = j;
        } else put_literal(t, escmap, d[i], o);
    }
    for (; i < size; i++) put_literal(t, escmap, d[i], o);
    buf_push32(o, size);
}
/* src/cr-diccode.c:142-221 */
void cro_dictionary_encode(cro_ctx* c, const uint8_t* in, uint32_t n, cro_buf* out) {
    uint32_t cnt[256] = {0}; uint8_t esc[10] = {0};
    for (uint32_t i = 0; i < n; i++) cnt[in[i]]++;
    for (int k = 0; k < 10; k++) {
        for (int j = 0; j < 256; j++) if (cnt[j] < cnt[esc[k]]) esc[k] = (uint8_t)j;
        cnt[esc[k]] = 0xFFFFFFFFu;
    }
    out->size = 0;
    for (uint32_t pos = 0; pos < n;) {
        cro_buf a = {0}, b = {0};
        uint32_t s1 = umin(1000000, n - pos); pos += s1;
        uint32_t s2 = umin(1000000, n - pos); pos += s2;
        dic_encode_sub(&c->dic, in + pos - s2 - s1, s1, esc, &a);
        dic_encode_sub(&c->dic, in + pos - s2, s2, esc, &b);
        buf_push32(out, (uint32_t)a.size); buf_push32(out, (uint32_t)b.size);
        buf_append(out, a.data, a.size); buf_append(out, b.data, b.size);
        cro_buf_free(&a); cro_buf_free(&b);
    }
    buf_append(out, esc, 10); buf_push(out, 1);
    if (out->size >= n) { out->size = 0; buf_append(out, in, n); buf_push(out, 0); }
}
/* src/cr-diccode.c:364-425: each sub-chunk is decoded back to front */
static void dic_decode_sub(const dict_t* t, const uint8_t* d, uint32_t size, const uint8_t esc[10], cro_buf* out) {
    uint8_t escmap[256] = {0};
    int L1 = L1_WORDS(t->nentries);
    for (int k = 0; k < 10; k++) escmap[esc[k]] = (uint8_t)(k + 1);
    uint32_t src = rd32(d + size - 4); int dst = (int)size - 4;
    size_t base = out->size; buf_reserve(out, base + src); out->size = base + src;
    uint8_t* o = out->data + base;
    uint32_t rev_pos = 0xFFFFFFFFu;
    while (src > 0) {
        int ch = d[--dst];
        if (!escmap[ch]) { o[--src] = (uint8_t)ch; continue; }
        int id = d[--dst];
        if (id >= L1) {
            id = d[--dst] * (256 - L1) + (id - L1);
            if (id == t->nentries) { o[--src] = (uint8_t)ch; continue; }
        }
        uint32_t wl = (uint32_t)strlen(t->word[id]);
        src -= wl; memcpy(o + src, t->word[id], wl);
        switch (escmap[ch]) {
            case 2: case 7:  o[src + wl - 1] = '.'; break;
            case 3: case 8:  o[src + wl - 1] = ','; break;
            case 4: case 9:  o[src + wl - 1] = ';'; break;
            case 5: case 10: o[src + wl - 1] = ':'; break;
        }
        if (escmap[ch] >= 6) o[src] ^= 0x20;
        if (rev_pos != 0xFFFFFFFFu && sentence_start(o, (int)rev_pos)) o[rev_pos] ^= 0x20;
        rev_pos = src;
    }
    if (rev_pos != 0xFFFFFFFFu && sentence_start(o, (int)rev_pos)) o[rev_pos] ^= 0x20;
}
/* src/cr-diccode.c:223-283 (without the fpout_sync side channel; see F4 in SURVEY.md) */
void cro_dictionary_decode(cro_ctx* c, const uint8_t* in, uint32_t n, cro_buf* out) {
    if (in[n - 1] == 0) { buf_append(out, in, n - 1); return; }
    const uint8_t* esc = in + n - 11;
    for (uint32_t pos = 0; pos + 11 < n;) {
        uint32_t s1 = rd32(in + pos), s2 = rd32(in + pos + 4);
        pos += 8 + s1 + s2;
        dic_decode_sub(&c->dic, in + pos - s2 - s1, s1, esc, out);
        dic_decode_sub(&c->dic, in + pos - s2, s2, esc, out);
    }
}

/* ================================================================== filters (src/cr-filter.c, filter_*.c) */
/* src/filter_x86opcode.h:38-62.  limit < 8 makes the reference loop bound wrap (undefined behaviour there);
 * we transform nothing in that case. */
static void e8e9(uint8_t* buf, uint32_t limit, int en_de, int32_t ncur, int32_t nend) {
    if (limit < 8) return;
    for (int32_t i = 0; (uint32_t)i < limit - 8;) {
        if ((buf[i++] & 254) != 0xe8) continue;
        int32_t op = (int32_t)rd32(buf + i), at = ncur + i;
        if (en_de == 0) {
            if (op >= -at && op < nend - at) op = (int32_t)((uint32_t)op + (uint32_t)at);
            else if (op > 0 && op < nend) op = (int32_t)((uint32_t)op - (uint32_t)nend);
        } else {
            if (op < 0) { if ((int32_t)((uint32_t)op + (uint32_t)at) >= 0) op = (int32_t)((uint32_t)op + (uint32_t)nend); }
            else if (op < nend) op = (int32_t)((uint32_t)op - (uint32_t)at);
        }
        wr32(buf + i, (uint32_t)op);
        i += 4;
    }
}
/* src/filter_x86_elf.c:106-156.  NOTE the detection call transforms [52, 52+min(imsz,len)) -- up to 52 bytes
 * past `len` -- and returns only `size`; callers must give 64 bytes of zeroed slack after the block. */
static uint32_t elf_filter(x86_state* s, uint8_t* buf, uint32_t len, int en_de) {
    uint32_t size = umin(s->imsz - s->curr, len);
    uint8_t* start = buf;
    if (!s->flag) {
        if (len < 52 || rd32(buf) != 0x464C457Fu || rd16(buf + 18) != 3) return 0;
        uint32_t shoff = rd32(buf + 32), est = shoff - 52;
        if (shoff < 52 || est >= (1u << 30)) return 0;
        s->imsz = est - 52;
        start = buf + 52;
        size = umin(s->imsz, len);
    }
    e8e9(start, size, en_de, (int32_t)s->curr, (int32_t)s->imsz);
    s->curr += size;
    s->flag = s->curr < s->imsz;
    return size;
}
/* src/filter_x86_pe.c:75-159.  The reference parses the COFF header and the section table wherever e_lfanew points inside the
 * block, without checking that they END inside it (its test `hdr_off + len < size` is always false for real sizes), so a
 * header near the end of a block makes it read heap memory behind the block -- undefined behaviour, and up to 2.6 MB of it.
 * This restatement defines those bytes as ZERO (zb16/zb32), which is what the CUDA path does (FilterHost::View) and what the
 * reference sees when the block sits in a freshly mapped buffer; and where the section table itself ends behind the block
 * (size_hdr > len: the reference's `len - size_hdr` wraps and it rewrites memory behind the block) it keeps the state
 * arithmetic and transforms nothing. */
static uint32_t zb16(const uint8_t* b, uint32_t len, uint64_t o) { return (o < len ? b[o] : 0) | (o + 1 < len ? b[o + 1] : 0) << 8; }
static uint32_t zb32(const uint8_t* b, uint32_t len, uint64_t o) { return zb16(b, len, o) | zb16(b, len, o + 2) << 16; }
static uint32_t pe_filter(x86_state* s, uint8_t* buf, uint32_t len, int en_de) {
    uint32_t size = umin(s->imsz - s->curr, len), ret = size;
    uint8_t* start = buf;
    int in_bounds = 1;
    if (!s->flag) {
        s->curr = 0;
        if (len < 0x3C + 4 || rd16(buf) != 0x5A4D) return 0;
        uint32_t hdr = rd32(buf + 0x3C);
        if (hdr >= len || zb32(buf, len, hdr) != 0x00004550u || hdr == 0) return 0;
        if (hdr + len < 24) return 0;
        uint32_t machine = zb16(buf, len, (uint64_t)hdr + 4), nsec = zb16(buf, len, (uint64_t)hdr + 6), optsz = zb16(buf, len, (uint64_t)hdr + 20),
                 chars = zb16(buf, len, (uint64_t)hdr + 22);
        if (machine != 0x14c && (chars & 2)) return 0;
        uint32_t sec_off = 24 + optsz, size_hdr = sec_off + nsec * 40, est = size_hdr;
        if (hdr + len < size_hdr) return 0;
        for (uint32_t i = 0; i < nsec; i++) est += zb32(buf, len, (uint64_t)hdr + sec_off + (uint64_t)i * 40 + 16);
        if (est > (1u << 28)) return 0;
        start = buf + size_hdr;
        s->imsz = est - size_hdr;
        size = umin(s->imsz, len - size_hdr);
        ret = size + size_hdr;
        in_bounds = size_hdr <= len;
    }
    if (in_bounds) e8e9(start, size, en_de, (int32_t)s->curr, (int32_t)s->imsz);
    s->curr += size;
    s->flag = s->curr < s->imsz;
    return ret;
}
/* src/filter_bmp.c:57-147: colour decorrelation, then left delta, then up delta over `rows` whole rows */
static uint32_t bmp_rows(uint8_t* b, uint32_t len, int width, int row_size, int bpp, int en_de) {
    int rows = (int)(len / (uint32_t)row_size), bytes = bpp / 8;
    if (en_de == 0) {
        for (int y = 0; y < rows; y++) for (int x = 0; x < width; x++) { uint8_t* p = b + (size_t)y * row_size + x * bytes; p[0] -= p[1]; p[2] -= p[1]; }
        for (int y = 0; y < rows; y++) for (int x = width - 1; x > 0; x--) { uint8_t* p = b + (size_t)y * row_size + x * bytes; for (int k = 0; k < bytes; k++) p[k] -= p[k - bytes]; }
        for (int y = rows - 1; y > 0; y--) for (int x = 0; x < width * bytes; x++) b[(size_t)y * row_size + x] -= b[(size_t)(y - 1) * row_size + x];
    } else {
        for (int y = 0; y < rows; y++) for (int x = 1; x < width; x++) { uint8_t* p = b + (size_t)y * row_size + x * bytes; for (int k = 0; k < bytes; k++) p[k] += p[k - bytes]; }
        for (int y = 1; y < rows; y++) for (int x = 0; x < width * bytes; x++) b[(size_t)y * row_size + x] += b[(size_t)(y - 1) * row_size + x];
        for (int y = 0; y < rows; y++) for (int x = 0; x < width; x++) { uint8_t* p = b + (size_t)y * row_size + x * bytes; p[0] += p[1]; p[2] += p[1]; }
    }
    return (uint32_t)row_size * (uint32_t)rows;
}
/* src/filter_bmp.c:149-204 */
static uint32_t bmp_filter(bmp_state* s, uint8_t* buf, uint32_t len, int en_de) {
    if (!s->flag) {
        if (len < 54 || rd16(buf) != 0x4d42 || rd16(buf + 26) != 1 || rd32(buf + 30) != 0) return 0;
        uint32_t fsize = rd32(buf + 2), off = rd32(buf + 10), isz = rd32(buf + 34), bpp = rd16(buf + 28);
        if ((isz != 0 && off + isz != fsize) || (bpp != 24 && bpp != 32)) return 0;
        int w = (int32_t)rd32(buf + 18), h = (int32_t)rd32(buf + 22);
        s->width = w < 0 ? -w : w; s->height = h < 0 ? -h : h;
        s->row_size = (int)((int)bpp * s->width + 31) / 32 * 4;
        s->bpp = (int)bpp;
        if (s->width < 4 || s->height < 4 || s->width >= (1 << 20) || s->height >= (1 << 20)) return 0;
        s->curr = (int)off; s->size = s->height * s->row_size; s->skip_size = 0; s->flag = 1;
        return off;
    }
    if (s->skip_size > 0) {
        uint32_t t = umin((uint32_t)s->skip_size, len);
        s->curr += (int)t; s->skip_size -= (int)t;
        return t;
    }
    uint32_t t = bmp_rows(buf, umin(len, (uint32_t)(s->size - s->curr)), s->width, s->row_size, s->bpp, en_de);
    s->curr += (int)t;
    if (s->curr < s->size) s->skip_size = (int)umin((uint32_t)s->row_size, (uint32_t)(s->size - s->curr));
    else s->flag = 0;
    return t;
}
static uint32_t run_filter(cro_ctx* c, int which, uint8_t* buf, uint32_t len, int en_de) {
    return which == 1 ? pe_filter(&c->pe, buf, len, en_de) : which == 2 ? elf_filter(&c->elf, buf, len, en_de) : bmp_filter(&c->bmp, buf, len, en_de);
}
/* src/cr-filter.c:33-73 */
int cro_filter_inplace(cro_ctx* c, uint8_t* buf, uint32_t len, int en_de) {
    int filt = 0;
    for (int pos = 0; pos < (int)len; pos++) {
        if (c->last_filter) {
            int n = (int)run_filter(c, c->last_filter, buf + pos, len - pos, en_de);
            if (n == 0) c->last_filter = 0; else { filt = 1; pos += n - 1; continue; }
        }
        for (int k = 1; k <= 3; k++) {
            int n = (int)run_filter(c, k, buf + pos, len - pos, en_de);
            if (n > 0) { filt = 1; c->last_filter = k; pos += n - 1; break; }
        }
    }
    return filt;
}

/* ================================================================== container (src/main.c) */
static const char* magic_of(int variant) {
    return variant == CRO_ROLZ ? "\x1f\x9d\x01\x01::0.11.0-comprolz" : variant == CRO_LZP ? "\x1f\x9d\x01\x01::0.11.0-comprop" : "\x1f\x9d\x01\x01::0.11.0-comprox";
}

/* src/main.c:137-218 */
int cro_compress(const cro_config* cfg, const uint8_t* in, size_t n, cro_buf* out) {
    cro_ctx* c = cro_new(cfg->variant);
    c->flexible = cfg->flexible;
    if (cfg->match_limit) c->match_limit = cfg->match_limit;
    cro_buf dic = {0}, tmp = {0}, pay = {0};
    int filt = 0;
    out->size = 0;
    buf_append(out, magic_of(cfg->variant), strlen(magic_of(cfg->variant)));
    cro_dicpick(in, n, &dic);
    cro_dictionary_load(c, (const char*)dic.data, 1);
    cro_dic_lcp_encode(&dic);
    cro_lzencode(c, dic.data, (uint32_t)dic.size, &pay);
    cro_reset_models(c);
    buf_push32(out, (uint32_t)pay.size); buf_append(out, pay.data, pay.size);
    uint8_t* blk = malloc((size_t)cfg->block_size + 64);
    for (size_t off = 0;; ) {
        uint32_t take = (uint32_t)(n - off < cfg->block_size ? n - off : cfg->block_size);
        memcpy(blk, in + off, take); memset(blk + take, 0, 64);
        off += take;
        if (cfg->filt) filt = cro_filter_inplace(c, blk, take, 0);
        cro_dictionary_encode(c, blk, take, &tmp);
        const cro_buf* fin = &tmp;
        if (!cfg->prec) { cro_lzencode(c, tmp.data, (uint32_t)tmp.size, &pay); fin = &pay; }
        if (fin->size > 0) {
            buf_push32(out, (uint32_t)fin->size); buf_push(out, (uint8_t)filt); buf_push(out, (uint8_t)cfg->prec);
            buf_append(out, fin->data, fin->size);
        }
        if (take < cfg->block_size) break;                                           /* feof only after a short read (F8) */
    }
    free(blk); cro_buf_free(&dic); cro_buf_free(&tmp); cro_buf_free(&pay);
    cro_free(c);
    return 0;
}
/* src/main.c:220-302.  Deviation (documented): the reference streams dictionary_decode output straight to the
 * file before the inverse filter runs (bug F4); we filter the decoded block, i.e. what the format intends. */
int cro_decompress(int variant, const uint8_t* in, size_t n, cro_buf* out) {
    size_t ml = strlen(magic_of(variant)), p = ml;
    if (n < ml + 4 || memcmp(in, magic_of(variant), ml)) return -1;
    cro_ctx* c = cro_new(variant);
    cro_buf dic = {0}, a = {0}, b = {0};
    uint32_t dl = rd32(in + p); p += 4;
    cro_lzdecode(c, in + p, dl, &dic); p += dl;
    cro_reset_models(c);
    cro_dic_lcp_decode(&dic);
    cro_dictionary_load(c, (const char*)dic.data, 0);
    out->size = 0;
    while (p + 6 <= n) {
        uint32_t sz = rd32(in + p); int filt = in[p + 4], prec = in[p + 5]; p += 6;
        a.size = b.size = 0;
        if (!prec) cro_lzdecode(c, in + p, sz, &a); else buf_append(&a, in + p, sz);
        p += sz;
        cro_dictionary_decode(c, a.data, (uint32_t)a.size, &b);
        if (filt) { buf_reserve(&b, b.size + 64); memset(b.data + b.size, 0, 64); cro_filter_inplace(c, b.data, (uint32_t)b.size, 1); }
        buf_append(out, b.data, b.size);
    }
    cro_buf_free(&dic); cro_buf_free(&a); cro_buf_free(&b);
    cro_free(c);
    return 0;
}
