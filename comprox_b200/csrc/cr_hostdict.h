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

struct HdWord { std::string w; uint32_t count; };

// word text of a counted entry: the letters starting at `first`, lower-cased (copyword, cr-dicpick.c:88-94)
static inline std::string hd_word_at(const uint8_t* in, uint64_t n, uint64_t first) {
    std::string w;
    w.push_back((char)(in[first] | 32));
    for (uint64_t i = first + 1; i < n && in[i] >= 'a' && in[i] <= 'z' && w.size() < 20; i++) w.push_back((char)in[i]);
    return w;
}

// src/cr-dicpick.c:218-257.  `words` = all words with count > 5.  Returns the dictionary text incl. the final NUL.
static inline std::string hd_dictionary_text(std::vector<HdWord>& words) {
    std::sort(words.begin(), words.end(), [](const HdWord& a, const HdWord& b) {
        if (a.count != b.count) return a.count > b.count;          // count descending
        return a.w > b.w;                                          // ties: word descending (:60-67)
    });
    int y = (int)words.size();
    if (y > HD_TOTAL_WORDS - 2) y = HD_TOTAL_WORDS - 2;
    if (y > HD_LEVEL1(y) - 2) {
        int x = HD_LEVEL1(y) - 2;
        std::sort(words.begin() + x, words.begin() + y, [](const HdWord& a, const HdWord& b) { return a.w < b.w; });
    }
    std::string t("\x20\x20\n" "http://www.\n");
    for (int x = 0; x < y; x++)
        if (x < HD_LEVEL1(y) || words[x].w.size() >= 3) { t += words[x].w; t.push_back('\n'); }
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

// src/cr-diccode.c:47-120: entries get a trailing blank when they end in a letter; 128-ary trie with the
// root's upper-case links ('A'..'Y', sic) and the ". , : ;" aliases of every blank edge.
struct HdTrie {
    std::vector<int32_t> next;   // nnode x 128
    std::vector<int32_t> id;     // nnode; -1 = inner node
    int nentries = 0, nword = 0;
    int level1() const { return HD_LEVEL1(nentries); }
    uint32_t new_node() { next.resize(next.size() + 128, 0); id.push_back(0); return (uint32_t)id.size() - 1; }
    void add(const std::string& w) {
        uint32_t node = 0;
        for (unsigned char ch : w) {
            if (next[(size_t)node * 128 + ch] == 0) {
                uint32_t nn = new_node();
                id[node] = -1;
                next[(size_t)node * 128 + ch] = (int32_t)nn;
            }
            node = (uint32_t)next[(size_t)node * 128 + ch];
        }
        id[node] = nword++;
    }
    // `text` = dictionary text up to (not including) the NUL
    void load(const char* text) {
        next.clear(); id.clear(); nentries = 0; nword = 0;
        std::vector<std::string> entries;
        std::string cur;
        for (const char* s = text; *s; s++) {
            if (*s == '\n') {
                if (!cur.empty() && (unsigned)(((unsigned char)cur.back() | 32) - 'a') < 26u) cur.push_back(' ');
                entries.push_back(cur); cur.clear();
            } else cur.push_back(*s);
        }
        nentries = (int)entries.size();
        new_node();
        for (auto& e : entries) add(e);
        for (int c = 'A'; c < 'Z'; c++) next[c] = next[c + 32];
        for (size_t i = 0; i < id.size(); i++) {
            int32_t* nx = &next[i * 128];
            if (nx[' '] > 0) { for (char a : { '.', ',', ':', ';' }) if (!nx[(int)a]) nx[(int)a] = nx[' ']; }
        }
    }
};
