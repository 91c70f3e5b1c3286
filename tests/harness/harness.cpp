// TEST CODE — compiles the product's algorithmic core (csrc/tok_core.cuh) and host table builders
// (csrc/tables.cpp) with g++ so that tests/test_core_host.py can check the matcher / merge / trie
// logic against the oracle on a machine without a GPU.  Never linked into libb200tok.so.
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../openvino_tokenizers_b200/csrc/tables.hpp"

using namespace b200tok;
#define HZ extern "C" __attribute__((visibility("default")))

HZ int hz_char_class(uint32_t cp) {
    const auto& t = host_class_tables();
    if (cp >= 0x110000) return 0;
    return t.stage2[(uint32_t)t.stage1[cp >> 8] * 256u + (cp & 255u)];
}

// Split one element with the sequential chain.  Returns piece count or a negative error.
HZ int64_t hz_split(const char* pattern, int64_t plen, const char* behaviour, int invert, int max_splits,
                    const uint8_t* chars, int64_t n, int32_t* out_b, int32_t* out_e, int64_t cap) {
    b200tok_regexsplit_desc d{pattern, plen, behaviour, invert, max_splits, 0};
    HostSplit hs;
    std::string err;
    int rc = parse_split(d, hs, err);
    if (rc) return rc;
    ScanCtx c{chars, (int)n, host_class_tables().view(), hs.spec.pat, hs.spec.class_mask};
    const SplitSpec spec = hs.host_spec();          // (PAT_VM: the compiled program, host pointers)
    SplitEmitter em;
    em.reset(hs.mode, hs.invert, hs.max_splits, (int)n);
    int64_t k = 0;
    split_element_scan(c, spec, hs.repeat, 0, (int)n, em, [&](int b, int e) {
        if (k < cap) { out_b[k] = b; out_e[k] = e; }
        ++k;
    });
    return k;
}

HZ int hz_contiguous_batch(const int32_t* rb, const int32_t* re, const int32_t* eb, const int32_t* ee, int64_t B, int64_t E, int64_t N) {
    return contiguous_batch(rb, re, eb, ee, B, E, N) ? 1 : 0;
}

struct HzBpe { HostBpe t; };
HZ void* hz_bpe_create(const b200tok_bpe_desc* d) {
    auto h = std::make_unique<HzBpe>();
    std::string err;
    if (build_bpe(*d, h->t, err)) return nullptr;
    return h.release();
}
HZ void hz_bpe_destroy(void* h) { delete (HzBpe*)h; }
HZ int64_t hz_bpe_info(void* h, int what) {
    auto* b = (HzBpe*)h;
    switch (what) {
    case 0: return b->t.n_duplicate_products;
    case 1: return b->t.bytes_only;
    case 2: return (int64_t)b->t.trie.n_nodes();
    case 3: return b->t.unk_id;
    default: return -1;
    }
}
// mode 0: in-place serial loop, 1: heap loop, 2: packed-key serial loop.  The end_suffix is appended like the reference does.
HZ int64_t hz_bpe_piece(void* h, const uint8_t* bytes, int64_t n, int mode, int32_t* out) {
    auto* b = (HzBpe*)h;
    std::string s((const char*)bytes, (size_t)n);
    s += b->t.end_suffix;
    const BpeTables T = b->t.view();
    const int L = (int)s.size();
    std::vector<int32_t> ids(2 * L + 2), rank(L + 2), newid(L + 2), prev(2 * L + 2), next(2 * L + 2);
    std::vector<uint16_t> birth(L + 2);
    std::vector<HeapEntry> heap(3 * L + 3);
    int m = bpe_symbolize(T, (const uint8_t*)s.data(), 0, L, ids.data());
    int cnt;
    if (mode == 0) {
        cnt = bpe_merge_serial(T.merges, ids.data(), rank.data(), newid.data(), birth.data(), m);
        std::memcpy(out, ids.data(), (size_t)cnt * 4);
    } else if (mode == 2) {
        if (m > kPackedMaxSymbols) return -1;
        std::vector<uint32_t> key(L + 2);
        cnt = bpe_merge_packed(T.merges, ids.data(), key.data(), m);
        std::memcpy(out, ids.data(), (size_t)cnt * 4);
    } else {
        cnt = bpe_merge_heap(T.merges, m, ids.data(), prev.data(), next.data(), heap.data(), out);
    }
    return cnt;
}

// The packed-key loop with its tie report (MergeTable::tie_check): returns the count, *tie = 1 if a merge found its own product on both sides.
HZ int64_t hz_bpe_piece_tie(void* h, const uint8_t* bytes, int64_t n, int32_t* out, int32_t* tie) {
    auto* b = (HzBpe*)h;
    std::string s((const char*)bytes, (size_t)n);
    s += b->t.end_suffix;
    const BpeTables T = b->t.view();
    const int L = (int)s.size();
    std::vector<int32_t> ids(2 * L + 2);
    std::vector<uint32_t> key(L + 2);
    const int m = bpe_symbolize(T, (const uint8_t*)s.data(), 0, L, ids.data());
    if (m > kPackedMaxSymbols) return -1;
    bool t = false;
    const int cnt = bpe_merge_packed(T.merges, ids.data(), key.data(), m, T.merges.tie_check ? &t : nullptr);
    std::memcpy(out, ids.data(), (size_t)cnt * 4);
    *tie = t ? 1 : 0;
    return cnt;
}

struct HzWp { HostWordpiece t; };
HZ void* hz_wp_create(const b200tok_wordpiece_desc* d) {
    auto h = std::make_unique<HzWp>();
    std::string err;
    if (build_wordpiece(*d, h->t, err)) return nullptr;
    return h.release();
}
HZ void hz_wp_destroy(void* h) { delete (HzWp*)h; }
HZ int64_t hz_wp_word(void* h, const uint8_t* bytes, int64_t n, int32_t unk, int32_t* out) {
    return wordpiece_word(((HzWp*)h)->t.view(), bytes, 0, (int)n, unk, out);
}

// Closed-form GPT-2 piece-start predicate (tok_core.cuh gpt2_piece_starts_at) over one element; the class array is
// built exactly like the kernel's pass A.  Writes piece begins; returns the count.
HZ int64_t hz_gpt2_closed_form(const uint8_t* s, int64_t n, int single_digits, int32_t* out_begins) {
    const ClassTables T = host_class_tables().view();
    std::vector<uint8_t> k((size_t)n + 8, 0);
    const int end = (int)n;
    for (int w = 0; w < end; ++w) {
        const uint8_t b = s[w];
        uint8_t c;
        if (b < 0x80) c = T.ascii[b];
        else if (is_cont_byte(b) && w > 0) {
            int j = w - 1;
            while (j >= 0 && j > w - 4 && is_cont_byte(s[j])) --j;
            c = C_CONT;
            if (j >= 0 && j > w - 4 && s[j] >= 0xC0) c |= char_class(s, j, end, T);
        } else c = char_class(s, w, end, T);
        k[w] = c;
    }
    int64_t cnt = 0;
    for (int i = 0; i < end; ++i)
        if (gpt2_piece_starts_at(s, k.data(), i, 0, end, single_digits != 0)) out_begins[cnt++] = i;
    return cnt;
}

// The shuffle-driven neighbour form (gpt2_start_nb), evaluated sequentially over an all-ASCII element.
HZ int64_t hz_gpt2_neighbour_form(const uint8_t* s, int64_t n, int single_digits, int32_t* out_begins) {
    const ClassTables T = host_class_tables().view();
    auto word = [&](int64_t i) -> uint32_t {
        if (i == -1) return G_BOS;
        if (i < 0 || i >= n) return 0;
        return gpt2_class_word(s[i], T.ascii[s[i] & 0x7F]);
    };
    int64_t cnt = 0;
    for (int64_t i = 0; i < n; ++i) {
        const bool apos = (word(i - 1) | word(i - 2) | word(i - 3)) & G_AP;
        if (gpt2_start_nb(word(i), word(i - 1), word(i - 2), word(i - 3), word(i - 4), word(i + 1), single_digits != 0, apos))
            out_begins[cnt++] = (int32_t)i;
    }
    return cnt;
}

// The word (bit-mask) form of the predicate (tok_core.cuh g2_starts), evaluated over one element the way the window
// kernel does: class bytes as in pass A, masks per 32 positions, carries from the neighbouring words.
HZ int64_t hz_gpt2_word_form(const uint8_t* s, int64_t n, int single_digits, int32_t* out_begins) {
    const ClassTables T = host_class_tables().view();
    const int end = (int)n;
    std::vector<uint8_t> k((size_t)n + 8, 0);
    for (int w = 0; w < end; ++w) {
        const uint8_t b = s[w];
        uint8_t c;
        if (b < 0x80) c = T.ascii[b];
        else if (is_cont_byte(b) && w > 0) {
            int j = w - 1;
            while (j >= 0 && j > w - 4 && is_cont_byte(s[j])) --j;
            c = C_CONT;
            if (j >= 0 && j > w - 4 && s[j] >= 0xC0) c |= char_class(s, j, end, T);
        } else c = char_class(s, w, end, T);
        k[w] = c;
    }
    const int nw = (end + 31) / 32;
    std::vector<G2Word> W((size_t)nw + 2, G2Word{0, 0, 0, 0, 0, 0, 0, 0, 0});   // W[i + 1] = word i; W[0] and W[nw + 1] stay empty
    for (int w = 0; w < end; ++w) {
        G2Word& g = W[(size_t)(w >> 5) + 1];
        const uint32_t bit = 1u << (w & 31);
        g.X |= bit;
        if (k[w] & C_L) g.L |= bit;
        if (k[w] & C_N) g.N |= bit;
        if (k[w] & C_S) g.S |= bit;
        if (k[w] & C_CONT) g.CONT |= bit;
        if (s[w] == 0x20) g.SP |= bit;
        if (s[w] == '\'') {
            const int cl = gpt2_contraction_len(s, w, end);
            if (cl == 2) g.A2 |= bit;
            if (cl == 3) g.A3 |= bit;
        }
        if ((k[w] & C_S) && s[w] >= 0x80 && !(k[w] & C_CONT)) {
            int j = w + 1;
            while (j < end && (k[j] & C_CONT)) ++j;
            if (j < end && !(k[j] & C_S)) g.MB |= bit;
        }
    }
    int64_t cnt = 0;
    uint32_t pc2 = 0, pc3 = 0;
    for (int i = 0; i < nw; ++i) {
        const G2Word& g = W[(size_t)i + 1];
        const G2Word& p = W[(size_t)i];
        const G2Word& nx = W[(size_t)i + 2];
        const uint32_t bos = i == 0 ? 1u : 0u;
        uint32_t c2, c3;
        g2_contractions(g, g2_ok1(p), bos, c2, c3);
        const uint32_t st = g2_starts(g, p, c2, c3, pc2, pc3, nx.X & ~nx.S, bos, single_digits != 0);
        for (int b = 0; b < 32; ++b) if ((st >> b) & 1u) out_begins[cnt++] = i * 32 + b;
        pc2 = c2; pc3 = c3;
    }
    return cnt;
}

// The word (bit-mask) form of the Llama-3 pattern (tok_core.cuh l3_starts) over one element: class bytes as in pass A, masks per
// 32 positions, the two cross-word fills and the number phase resolved sequentially (the kernel resolves them with ballots).
HZ int64_t hz_llama3_word_form(const uint8_t* s, int64_t n, int32_t* out_begins) {
    const ClassTables T = host_class_tables().view();
    const int end = (int)n;
    std::vector<uint8_t> k((size_t)n + 8, 0);
    for (int w = 0; w < end; ++w) {
        const uint8_t b = s[w];
        uint8_t c;
        if (b < 0x80) c = T.ascii[b];
        else if (is_cont_byte(b) && w > 0) {
            int j = w - 1;
            while (j >= 0 && j > w - 4 && is_cont_byte(s[j])) --j;
            c = C_CONT;
            if (j >= 0 && j > w - 4 && s[j] >= 0xC0) c |= char_class(s, j, end, T);
        } else c = char_class(s, w, end, T);
        k[w] = c;
    }
    auto is_other = [&](int i) { return !(k[i] & (C_L | C_N | C_S)); };
    const int nw = (end + 31) / 32;
    const L3Word Z{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<L3Word> W((size_t)nw + 2, Z);      // W[i + 1] = word i
    for (int w = 0; w < end; ++w) {
        L3Word& g = W[(size_t)(w >> 5) + 1];
        const uint32_t bit = 1u << (w & 31);
        g.X |= bit;
        if (k[w] & C_L) g.L |= bit;
        if (k[w] & C_N) { g.N |= bit; if (s[w] >= 0x80) return -2; }      // non-ASCII digit: the kernel hands such rows to the generic path
        if (k[w] & C_S) g.S |= bit;
        if ((k[w] & C_NL) && !(k[w] & C_CONT)) g.NL |= bit;
        if (k[w] & C_CONT) g.CONT |= bit;
        if (s[w] == 0x20) g.SP |= bit;
        if (s[w] == '\'') {
            const int cl = llama3_contraction_len(s, w, end);
            if (cl == 2) g.A2 |= bit;
            if (cl == 3) g.A3 |= bit;
        }
        if ((k[w] & C_S) && s[w] >= 0x80 && !(k[w] & C_CONT)) {
            int j = w + 1;
            while (j < end && (k[j] & C_CONT)) ++j;
            if (j < end && !(k[j] & C_S)) g.MB |= bit;
        }
        if (w > 0 && !(k[w] & C_CONT) && (k[w - 1] & C_CONT)) {          // previous char is multi-byte: is it an other char at which a match starts?
            int j = w - 1;
            while (j > 0 && (k[j] & C_CONT)) --j;
            if (is_other(j) && (j == 0 || (!is_other(j - 1) && s[j - 1] != 0x20))) g.PG |= bit;
        }
    }
    std::vector<uint32_t> mso((size_t)nw + 2, 0), d2((size_t)nw + 2, 0), d3((size_t)nw + 2, 0), lead((size_t)nw + 2, 0), tail((size_t)nw + 2, 0), nst((size_t)nw + 2, 0);
    for (int i = 1; i <= nw; ++i) {
        mso[(size_t)i] = l3_mso(W[(size_t)i], W[(size_t)i - 1]);
        d2[(size_t)i] = W[(size_t)i].A2 & mso[(size_t)i];
        d3[(size_t)i] = W[(size_t)i].A3 & mso[(size_t)i];
    }
    uint32_t carry = 0;                                                     // upward fill of newline runs that follow an other char
    for (int i = 1; i <= nw; ++i) {
        const L3Word& g = W[(size_t)i];
        uint32_t seeds = l3_lead_seeds(g, W[(size_t)i - 1]);
        if (carry && (g.NL & 1u)) seeds |= 1u;
        lead[(size_t)i] = l3_fill_up(g.NL, seeds);
        carry = lead[(size_t)i] >> 31;
    }
    carry = 0;                                                              // downward fill of the tail of every whitespace run
    for (int i = nw; i >= 1; --i) {
        const L3Word& g = W[(size_t)i];
        const uint32_t Tb = g.S & ~g.NL;
        uint32_t seeds = l3_tail_seeds(g, W[(size_t)i + 1].S);
        if (carry && (Tb >> 31)) seeds |= 0x80000000u;
        tail[(size_t)i] = l3_fill_down(Tb, seeds);
        carry = tail[(size_t)i] & 1u;
    }
    for (int i = 1; i <= nw; ++i) {                                         // digits of the run before the word's first bit
        int phase = 0;
        for (int p = (i - 1) * 32 - 1; p >= 0 && (k[p] & C_N); --p) ++phase;
        nst[(size_t)i] = l3_number_starts(W[(size_t)i].N, W[(size_t)i - 1].N >> 31, phase, i == 1);
    }
    int64_t cnt = 0;
    for (int i = 1; i <= nw; ++i) {
        const uint32_t st = l3_starts(W[(size_t)i], W[(size_t)i - 1], mso[(size_t)i], mso[(size_t)i - 1], d2[(size_t)i], d3[(size_t)i], d2[(size_t)i - 1],
                                      d3[(size_t)i - 1], lead[(size_t)i], lead[(size_t)i - 1], tail[(size_t)i], nst[(size_t)i],
                                      W[(size_t)i + 1].X & ~W[(size_t)i + 1].S, i == 1 ? 1u : 0u);
        for (int b = 0; b < 32; ++b) if ((st >> b) & 1u) out_begins[cnt++] = (i - 1) * 32 + b;
    }
    return cnt;
}

// SpecialTokensSplit on the host: the pattern parser (tables.cpp parse_special) + the per-position matcher and scan of
// tok_core.cuh over one element.  Returns the piece count (< 0: the parser's error code).
HZ int64_t hz_special_split(const char* pattern, int64_t plen, const uint8_t* chars, int64_t n, int32_t* out_b, int32_t* out_e,
                            uint8_t* out_skip, int64_t cap) {
    HostSpecial hs;
    std::string err;
    const int rc = parse_special(pattern, plen, hs, err);
    if (rc) return rc;
    SpecialTables st{};
    st.n_groups = (int32_t)hs.groups.size();
    for (int g = 0; g < st.n_groups; ++g) {
        st.trie[g] = hs.groups[(size_t)g].trie.view();
        st.strip_left[g] = hs.groups[(size_t)g].strip_left;
        st.strip_right[g] = hs.groups[(size_t)g].strip_right;
    }
    for (int k = 0; k < 8; ++k) st.first[k] = hs.first[(size_t)k];
    st.ws_token = hs.ws_token;
    int64_t k = 0;
    special_split_element(st, host_class_tables().view(), chars, 0, (int)n, [&](int b, int e, int skip) {
        if (k < cap) { out_b[k] = b; out_e[k] = e; out_skip[k] = (uint8_t)skip; }
        ++k;
    });
    return k;
}


// Normalisers on the host: the product's pattern / blob parsers (tables.cpp) + the sequential scan of tok_core.cuh over
// each element.  kind 0: RegexNormalization (a = search, b = replace, flag = global_replace); kind 1: CharsMapNormalization
// (a = blob).  Returns the bytes produced (< 0: the parser's error code).
HZ int64_t hz_normalize(int kind, const uint8_t* a, int64_t alen, const uint8_t* b, int64_t blen, int flag, const int32_t* begins,
                        const int32_t* ends, const uint8_t* chars, const uint8_t* skips, int64_t n, int32_t* ob, int32_t* oe, uint8_t* oc,
                        int64_t cap) {
    HostNorm hn;
    std::string err;
    uint16_t b2c[256];
    uint8_t pair_map[256] = {};
    if (kind >= 2 && kind <= 4) {      // BytesToChars / UTF8Validate(flag = replace mode) / CharsToBytes elements run on the same scan (api.cu builds these rules inline)
        hn.rule = NormRule{};
        hn.rule.kind = kind == 2 ? NORM_B2C : kind == 3 ? NORM_UTF8 : NORM_C2B;
        hn.rule.literal_cp = -1;
        hn.rule.global = kind == 3 ? (flag != 0) : 1;
    } else {
        const int rc = kind == 0 ? parse_regex_norm((const char*)a, alen, (const char*)b, blen, flag, hn, err) : parse_charsmap(a, alen, 0, 0, 0, hn, err);
        if (rc) return rc;
    }
    NormRule R = hn.rule;
    R.cls = host_norm_class_tables().view();
    R.units = hn.units.data(); R.n_units = (uint32_t)hn.units.size();
    R.normalized = hn.normalized.data(); R.n_normalized = (uint32_t)hn.normalized.size();
    R.atab = hn.atab.data();
    if (kind == 2 || kind == 4) gpt2_build_byte_codepoints(b2c);
    if (kind == 2) { R.normalized = reinterpret_cast<const uint8_t*>(b2c); R.n_normalized = 512; }
    if (kind == 4) {                   // flag = size of the chars buffer (the follower of a trailing lead byte may lie beyond the element)
        for (int x = 0; x < 256; ++x) if (b2c[x] >= 0x80) pair_map[((0xC0 | (b2c[x] >> 6)) - 194) * 64 + ((0x80 | (b2c[x] & 0x3F)) - 128)] = (uint8_t)x;
        R.normalized = pair_map; R.n_normalized = 256; R.n_units = (uint32_t)flag;
    }
    int64_t cur = 0;
    for (int64_t i = 0; i < n; ++i) {
        ob[i] = (int32_t)cur;
        int l;
        if (skips && skips[i]) { l = ends[i] - begins[i]; if (cur + l <= cap) std::memcpy(oc + cur, chars + begins[i], (size_t)l); }
        else { l = norm_string(R, chars, begins[i], ends[i], nullptr); if (cur + l <= cap) norm_string(R, chars, begins[i], ends[i], oc + cur); }
        cur += l;
        oe[i] = (int32_t)cur;
    }
    return cur;
}


// Consistency of the per-ASCII-byte shortcut tables of a charsmap with the general step: for every ASCII byte a whose
// flags allow the shortcut, the step at [a, next] must consume 1 byte and emit amap[a].  Returns the number of violations.
HZ int64_t hz_charsmap_ascii_table_check(const uint8_t* blob, int64_t len) {
    HostNorm hn;
    std::string err;
    if (parse_charsmap(blob, len, 0, 0, 0, hn, err)) return -1;
    NormRule R = hn.rule;
    R.units = hn.units.data(); R.n_units = (uint32_t)hn.units.size();
    R.normalized = hn.normalized.data(); R.n_normalized = (uint32_t)hn.normalized.size();
    int64_t bad = 0;
    const uint8_t tails[][3] = {{0xCC, 0x81, 0}, {0xCC, 0x8A, 0}, {0xE3, 0x82, 0x99}, {0xD6, 0xBC, 0}};
    for (int a = 0; a < 128; ++a) {
        const uint8_t aflag = hn.atab[128 + (size_t)a], amap = hn.atab[(size_t)a];
        if (aflag & (NA_COMPLEX | NA_ASCII_KIDS)) continue;
        for (int nx = 0; nx < 128 + 4; ++nx) {
            uint8_t s[4] = {(uint8_t)a, 0, 0, 0};
            int n = 2;
            if (nx < 128) s[1] = (uint8_t)nx;
            else { if (aflag & NA_OTHER_KIDS) continue; std::memcpy(s + 1, tails[nx - 128], 3); n = tails[nx - 128][2] ? 4 : 3; }
            const NormStep st = norm_eval(R, s, 0, 0, n, false);
            uint8_t out[64] = {0};
            if (st.olen <= 64) norm_emit(R, st, s, 0, out);
            if (st.consumed != 1 || st.olen != 1 || out[0] != amap) ++bad;
        }
    }
    return bad;
}


// Composed byte table of a chain of normalisers (tables.cpp compose_norm_chain).  hz_chain_reset(); hz_chain_add(...) per
// op (same arguments as hz_normalize); hz_chain_table(T) -> 1 composable, 0 not, < 0 parser error of the last add.
namespace { std::vector<std::unique_ptr<HostNorm>> g_chain; int g_chain_err = 0; }
HZ void hz_chain_reset() { g_chain.clear(); g_chain_err = 0; }
HZ int hz_chain_add(int kind, const uint8_t* a, int64_t alen, const uint8_t* b, int64_t blen, int flag) {
    auto hn = std::make_unique<HostNorm>();
    std::string err;
    const int rc = kind == 0 ? parse_regex_norm((const char*)a, alen, (const char*)b, blen, flag, *hn, err) : parse_charsmap(a, alen, 0, 0, 0, *hn, err);
    if (rc) { g_chain_err = rc; return rc; }
    g_chain.push_back(std::move(hn));
    return 0;
}
HZ int hz_chain_table(uint8_t* T) {
    if (g_chain_err) return g_chain_err;
    std::vector<const HostNorm*> ops;
    for (auto& h : g_chain) ops.push_back(h.get());
    return compose_norm_chain(ops.data(), (int)ops.size(), T) ? 1 : 0;
}
