// cr_hostdict.h -- the small serial parts of the static-dictionary stage that stay on the host.
//
// After the GPU has counted the words (cr_dict.cuh), at most a few hundred thousand (first position, count)
// pairs come back.  Ordering them, writing the dictionary text, front-coding it and building the lookup trie
// are O(#words) control logic (< 25 000 words kept), exactly the steps of
//   src/cr-dicpick.c:218-257 (selection + two-level ordering), :261-305 (dic_lcp_encode),
//   src/cr-diccode.c:76-120 (dictionary_load).
// No block data is processed here.
#pragma once
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>

#define HD_TOTAL_WORDS 25000
#define HD_LEVEL1(n)   ((65535 - (n)) / 255 - 1)

// a counted word: up to 20 lower-case letters, zero padded, kept as three big-endian 64-bit keys so that
// comparing keys compares the strings (strcmp order) without touching memory twice
struct HdWord {
    uint64_t k[3]; uint32_t count; uint32_t len;
    void set(const uint8_t* in, uint64_t n, uint64_t first) {       // copyword, cr-dicpick.c:88-94
        char w[24] = {0};
        w[0] = (char)(in[first] | 32); len = 1;
        for (uint64_t i = first + 1; i < n && in[i] >= 'a' && in[i] <= 'z' && len < 20; i++) w[len++] = (char)in[i];
        for (int q = 0; q < 3; q++) { uint64_t v = 0; for (int b = 0; b < 8; b++) v = v << 8 | (uint8_t)w[q * 8 + b]; k[q] = v; }
    }
    void append_to(std::string& t) const { for (uint32_t i = 0; i < len; i++) t.push_back((char)(k[i >> 3] >> (56 - 8 * (i & 7)))); }
    bool word_less(const HdWord& o) const { return k[0] != o.k[0] ? k[0] < o.k[0] : k[1] != o.k[1] ? k[1] < o.k[1] : k[2] < o.k[2]; }
};

// src/cr-dicpick.c:218-257.  `words` = all words with count > 5.  Returns the dictionary text incl. the final NUL.
static inline std::string hd_dictionary_text(std::vector<HdWord>& words) {
    std::sort(words.begin(), words.end(), [](const HdWord& a, const HdWord& b) {
        if (a.count != b.count) return a.count > b.count;          // count descending
        return b.word_less(a);                                     // ties: word descending (:60-67)
    });
    int y = (int)words.size();
    if (y > HD_TOTAL_WORDS - 2) y = HD_TOTAL_WORDS - 2;
    if (y > HD_LEVEL1(y) - 2) {
        int x = HD_LEVEL1(y) - 2;
        std::sort(words.begin() + x, words.begin() + y, [](const HdWord& a, const HdWord& b) { return a.word_less(b); });
    }
    std::string t("\x20\x20\n" "http://www.\n");
    t.reserve((size_t)y * 10 + 64);
    for (int x = 0; x < y; x++)
        if (x < HD_LEVEL1(y) || words[x].len >= 3) { words[x].append_to(t); t.push_back('\n'); }
    t.push_back('\0');
    return t;
}

// src/cr-dicpick.c:261-305: first line verbatim, then <lcp with previous line><suffix>\n ..., terminator 0xFF
static inline std::vector<uint8_t> hd_lcp_encode(const std::string& t) {
    std::vector<uint8_t> o;
    const uint8_t* d = (const uint8_t*)t.data();
    size_t prev = 0, cur = 0;
    while (d[cur] != '\n') o.push_back(d[cur++]);
    cur++; o.push_back('\n');
    while (d[cur] != 0) {
        int lcp = 0;
        while (d[prev + lcp] == d[cur + lcp]) lcp++;
        o.push_back((uint8_t)lcp);
        prev = cur; cur += lcp;
        while (d[cur] != '\n') o.push_back(d[cur++]);
        cur++; o.push_back('\n');
    }
    o.push_back(255);
    return o;
}

// src/cr-dicpick.c:307-346: inverse of hd_lcp_encode; returns the dictionary text without the final NUL
static inline std::string hd_lcp_decode(const uint8_t* d, size_t n) {
    std::string o;
    size_t wi = 0, wo = 0;
    while (wi < n && d[wi] != '\n') o.push_back((char)d[wi++]);
    wi++; o.push_back('\n');
    while (wi < n && d[wi] != 255) {
        int lcp = d[wi++];
        while (lcp-- > 0 && wo < o.size()) o.push_back(o[wo++]);          // (wo < size always holds for a well-formed text; damaged input must not read past the end)
        while (wi < n && d[wi] != '\n') o.push_back((char)d[wi++]);
        wi++; o.push_back('\n');
        while (wo < o.size() && o[wo] != '\n') wo++;
        wo++;
    }
    return o;
}
// the dictionary entries as dictionary_load stores them (src/cr-diccode.c:82-93): one per line, a blank appended
// to entries that end in a letter
static inline std::vector<std::string> hd_entries(const char* text) {
    std::vector<std::string> entries;
    std::string cur;
    for (const char* s = text; *s; s++) {
        if (*s == '\n') {
            if (!cur.empty() && (unsigned)(((unsigned char)cur.back() | 32) - 'a') < 26u) cur.push_back(' ');
            entries.push_back(cur); cur.clear();
        } else cur.push_back(*s);
    }
    return entries;
}

// src/cr-diccode.c:47-120: entries get a trailing blank when they end in a letter; 128-ary trie with the
// root's upper-case links ('A'..'Y', sic) and the ". , : ;" aliases of every blank edge.
// The reference stores 128 child slots per node (516 B/node, tens of MB).  Only the edges that exist are kept
// here, in an open-addressing table keyed by (parent node, byte): a few MB that stay resident in L2 and are
// cheap to upload.  Lookup semantics are identical: a missing edge is child 0 (= the root, "no match").
struct HdTrie {
    std::vector<uint32_t> ekey, eval;   // edge table: key = ((parent << 7) | byte) + 1, 0 = empty; value = child node
    std::vector<int32_t> id;            // per node: word index, -1 = inner node, 0 for a fresh node (as the reference)
    uint32_t mask = 0;
    int nentries = 0, nword = 0;
    int level1() const { return HD_LEVEL1(nentries); }
    static uint32_t hash(uint32_t key) { return key * 2654435761u; }
    uint32_t child(uint32_t node, uint32_t ch) const {
        const uint32_t key = ((node << 7) | ch) + 1;
        for (uint32_t h = hash(key) >> 8;; h++) { const uint32_t k = ekey[h & mask]; if (k == key) return eval[h & mask]; if (k == 0) return 0; }
    }
    void link(uint32_t node, uint32_t ch, uint32_t to) {
        const uint32_t key = ((node << 7) | ch) + 1;
        uint32_t h = hash(key) >> 8;
        while (ekey[h & mask] != 0 && ekey[h & mask] != key) h++;
        ekey[h & mask] = key; eval[h & mask] = to;
    }
    void add(const std::string& w) {                     // dictionary_add_word, :47-70
        uint32_t node = 0;
        for (unsigned char ch : w) {
            uint32_t nx = child(node, ch);
            if (nx == 0) { nx = (uint32_t)id.size(); id.push_back(0); id[node] = -1; link(node, ch, nx); }
            node = nx;
        }
        id[node] = nword++;
    }
    // `text` = dictionary text up to (not including) the NUL
    void load(const char* text) {
        id.clear(); nentries = 0; nword = 0;
        std::vector<std::string> entries;
        std::string cur;
        size_t chars = 0;
        for (const char* s = text; *s; s++) {
            if (*s == '\n') {
                if (!cur.empty() && (unsigned)(((unsigned char)cur.back() | 32) - 'a') < 26u) cur.push_back(' ');
                chars += cur.size(); entries.push_back(cur); cur.clear();
            } else cur.push_back(*s);
        }
        nentries = (int)entries.size();
        uint32_t cap = 1024;
        while (cap < 2 * (chars + 5 * entries.size() + 64)) cap <<= 1;      // edges <= chars, + 4 aliases per word, + root links
        mask = cap - 1;
        ekey.assign(cap, 0); eval.assign(cap, 0);
        id.push_back(0);                                                     // root
        for (auto& e : entries) add(e);
        for (int c = 'A'; c < 'Z'; c++) { uint32_t lo = child(0, (uint32_t)c + 32); if (lo || child(0, (uint32_t)c)) link(0, (uint32_t)c, lo); }   // :107-109
        const uint32_t nnode = (uint32_t)id.size();
        for (uint32_t i = 0; i < nnode; i++) {                               // :110-117
            const uint32_t sp = child(i, ' ');
            if (sp > 0) for (char a : { '.', ',', ':', ';' }) if (!child(i, (uint32_t)a)) link(i, (uint32_t)a, sp);
        }
    }
};
