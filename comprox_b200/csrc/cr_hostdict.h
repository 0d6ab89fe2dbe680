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
// std::sort on (64-bit proxy key, index) pairs, then the full comparator inside runs of equal proxies.  proxy(a) < proxy(b) must imply
// less(a, b).  (Two sorts of 20 000 32-byte records with three-key comparisons were 2.6 ms of host time with the GPU idle.)
template <class KeyFn, class Less>
static inline void hd_proxy_sort(std::vector<HdWord>& w, size_t lo, size_t hi, KeyFn key, Less less) {
    const size_t n = hi - lo;
    if (n < 2) return;
    std::vector<std::pair<uint64_t, uint32_t>> p(n);
    for (size_t i = 0; i < n; i++) p[i] = std::make_pair(key(w[lo + i]), (uint32_t)i);
    std::sort(p.begin(), p.end());
    for (size_t a = 0; a < n;) {
        size_t b = a + 1;
        while (b < n && p[b].first == p[a].first) b++;
        if (b - a > 1) std::sort(p.begin() + a, p.begin() + b, [&](const std::pair<uint64_t, uint32_t>& x, const std::pair<uint64_t, uint32_t>& y) { return less(w[lo + x.second], w[lo + y.second]); });
        a = b;
    }
    std::vector<HdWord> t(n);
    for (size_t i = 0; i < n; i++) t[i] = w[lo + p[i].second];
    for (size_t i = 0; i < n; i++) w[lo + i] = t[i];
}
static inline std::string hd_dictionary_text(std::vector<HdWord>& words) {
    auto by_count = [](const HdWord& a, const HdWord& b) {
        if (a.count != b.count) return a.count > b.count;          // count descending
        return b.word_less(a);                                     // ties: word descending (:60-67)
    };
    // The reference sorts everything by count, keeps the first y and then re-sorts all but the first x = level1 - 2 of those by word.
    // Words are distinct, so the order by count is total and only two things of it are used: WHICH words are the first y, and the first
    // x in order.  Select instead of sorting (nth_element), sort the x, sort the rest by word: one sort of y records instead of two.
    int y = (int)words.size();
    if (y > HD_TOTAL_WORDS - 2) {
        y = HD_TOTAL_WORDS - 2;
        std::nth_element(words.begin(), words.begin() + y, words.end(), by_count);
    }
    if (y > HD_LEVEL1(y) - 2) {
        const int x = HD_LEVEL1(y) - 2;
        if (x > 0) { std::nth_element(words.begin(), words.begin() + x, words.begin() + y, by_count); std::sort(words.begin(), words.begin() + x, by_count); }
        hd_proxy_sort(words, (size_t)(x > 0 ? x : 0), (size_t)y, [](const HdWord& a) { return a.k[0]; }, [](const HdWord& a, const HdWord& b) { return a.word_less(b); });
    } else std::sort(words.begin(), words.begin() + y, by_count);
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
struct alignas(8) HdEdge { uint32_t key, val; };   // key = ((parent << 7) | byte) + 1, 0 = empty; val = child node: one 8-byte load per probe
struct HdTrie {
    std::vector<HdEdge> edge;           // the edge table
    std::vector<int32_t> id;            // per node: word index, -1 = inner node, 0 for a fresh node (as the reference)
    uint32_t mask = 0;
    int nentries = 0, nword = 0;
    int level1() const { return HD_LEVEL1(nentries); }
    static uint32_t hash(uint32_t key) { return key * 2654435761u; }
    uint32_t child(uint32_t node, uint32_t ch) const {
        const uint32_t key = ((node << 7) | ch) + 1;
        for (uint32_t h = hash(key) >> 8;; h++) { const HdEdge e = edge[h & mask]; if (e.key == key) return e.val; if (e.key == 0) return 0; }
    }
    void link(uint32_t node, uint32_t ch, uint32_t to) {
        const uint32_t key = ((node << 7) | ch) + 1;
        uint32_t h = hash(key) >> 8;
        while (edge[h & mask].key != 0 && edge[h & mask].key != key) h++;
        edge[h & mask].key = key; edge[h & mask].val = to;
    }
    void link_if_absent(uint32_t node, uint32_t ch, uint32_t to) {       // one probe sequence for "look up, add if missing"
        const uint32_t key = ((node << 7) | ch) + 1;
        for (uint32_t h = hash(key) >> 8;; h++) {
            HdEdge& e = edge[h & mask];
            if (e.key == key) return;
            if (e.key == 0) { e.key = key; e.val = to; return; }
        }
    }
    // child(node, ch), made if it does not exist yet (one probe sequence instead of two); *made tells
    uint32_t child_or_new(uint32_t node, uint32_t ch, bool* made) {
        const uint32_t key = ((node << 7) | ch) + 1;
        uint32_t h = hash(key) >> 8;
        for (;; h++) {
            HdEdge& e = edge[h & mask];
            if (e.key == key) { *made = false; return e.val; }
            if (e.key == 0) { e.key = key; e.val = (uint32_t)id.size(); *made = true; return e.val; }
        }
    }
    std::vector<uint32_t> space_parents;                 // nodes that own a blank edge, in the order the edges appeared
    void add(const char* w, size_t len, bool blank) {    // dictionary_add_word, :47-70 (`blank`: the entry ends in a letter and gets a ' ')
        uint32_t node = 0;
        for (size_t i = 0; i < len + (blank ? 1 : 0); i++) {
            const unsigned char ch = i < len ? (unsigned char)w[i] : (unsigned char)' ';
            bool made;
            const uint32_t nx = child_or_new(node, ch, &made);
            if (made) { id.push_back(0); id[node] = -1; if (ch == ' ') space_parents.push_back(node); }
            node = nx;
        }
        id[node] = nword++;
    }
    // `text` = dictionary text up to (not including) the NUL
    void load(const char* text) {
        id.clear(); space_parents.clear(); nentries = 0; nword = 0;
        size_t chars = 0, lines = 0;
        for (const char* s = text; *s; s++) { if (*s == '\n') lines++; else chars++; }
        nentries = (int)lines;
        // edges <= characters + one blank per entry + 4 aliases per entry + the root's upper-case links; load factor <= 2/3
        const size_t bound = chars + 5 * lines + 64;
        uint32_t cap = 1024;
        while ((size_t)cap * 2 < bound * 3) cap <<= 1;
        mask = cap - 1;
        edge.assign(cap, HdEdge{0, 0});
        id.reserve(bound);
        id.push_back(0);                                                     // root
        for (const char* s = text; *s;) {                                    // one entry per line (:82-93)
            const char* e = s;
            while (*e != '\n') e++;
            const size_t len = (size_t)(e - s);
            add(s, len, len > 0 && (unsigned)(((unsigned char)e[-1] | 32) - 'a') < 26u);
            s = e + 1;
        }
        for (int c = 'A'; c < 'Z'; c++) { uint32_t lo = child(0, (uint32_t)c + 32); if (lo || child(0, (uint32_t)c)) link(0, (uint32_t)c, lo); }   // :107-109
        // :110-117: every node with a blank edge gets ". , : ;" aliases of it (where those edges do not exist).  The reference walks
        // all nodes; only the owners of a blank edge do anything, and they were noted when the edge was made
        for (uint32_t i : space_parents) {
            const uint32_t sp = child(i, ' ');
            for (char a : { '.', ',', ':', ';' }) link_if_absent(i, (uint32_t)a, sp);
        }
    }
};
