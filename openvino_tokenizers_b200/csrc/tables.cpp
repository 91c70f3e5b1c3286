// tables.cpp — see tables.hpp.
#include "tables.hpp"

#include <algorithm>
#include <cctype>
#include <cstring>
#include <map>
#include <unordered_map>

namespace b200tok {

// ------------------------------------------------------------------------------------------
// Unicode class tables (two-stage), from the PCRE2-derived run list.
// ------------------------------------------------------------------------------------------
namespace {
struct Run { uint32_t cp; uint8_t flags; };
const Run kRuns[] = {
#include "unicode_ranges.inc"
};

bool in_bert_punct_or_cjk(uint32_t cp, uint8_t f) {
    if ((cp >= 0x21 && cp <= 0x2F) || (cp >= 0x3A && cp <= 0x40) || (cp >= 0x5B && cp <= 0x60) || (cp >= 0x7B && cp <= 0x7E))
        return true;
    if (f & C_P) return true;
    static const uint32_t cjk[][2] = {{0x4E00, 0x9FFF}, {0x3400, 0x4DBF}, {0x20000, 0x2A6DF}, {0x2A700, 0x2B73F},
                                      {0x2B740, 0x2B81F}, {0x2B820, 0x2CEAF}, {0xF900, 0xFAFF}, {0x2F800, 0x2FA1F}};
    for (auto& r : cjk)
        if (cp >= r[0] && cp <= r[1]) return true;
    return false;
}
void two_stage(const std::vector<uint8_t>& flat, HostClassTables& t);
}  // namespace

const HostClassTables& host_class_tables() {
    static const HostClassTables tables = [] {
        HostClassTables t;
        std::vector<uint8_t> flat(0x110000);
        const size_t n = sizeof(kRuns) / sizeof(kRuns[0]);
        for (size_t i = 0; i < n; ++i) {
            const uint32_t a = kRuns[i].cp, b = (i + 1 < n) ? kRuns[i + 1].cp : 0x110000u;
            for (uint32_t cp = a; cp < b; ++cp) {
                uint8_t f = kRuns[i].flags;
                if (in_bert_punct_or_cjk(cp, f)) f |= C_BP;
                if (cp == '\r' || cp == '\n') f |= C_NL;
                flat[cp] = f;
            }
        }
        two_stage(flat, t);
        return t;
    }();
    return tables;
}

namespace {
void two_stage(const std::vector<uint8_t>& flat, HostClassTables& t) {
    t.ascii.assign(flat.begin(), flat.begin() + 128);
    t.stage1.resize(0x1100);
    std::map<std::vector<uint8_t>, uint16_t> seen;
    for (uint32_t blk = 0; blk < 0x1100; ++blk) {
        std::vector<uint8_t> v(flat.begin() + blk * 256, flat.begin() + blk * 256 + 256);
        auto it = seen.find(v);
        if (it == seen.end()) {
            const uint16_t idx = (uint16_t)seen.size();
            it = seen.emplace(v, idx).first;
            t.stage2.insert(t.stage2.end(), v.begin(), v.end());
        }
        t.stage1[blk] = it->second;
    }
}
const Run kNormRuns[] = {
#include "unicode_norm_ranges.inc"
};
}  // namespace

const HostClassTables& host_norm_class_tables() {
    static const HostClassTables tables = [] {
        HostClassTables t;
        std::vector<uint8_t> flat(0x110000);
        const size_t n = sizeof(kNormRuns) / sizeof(kNormRuns[0]);
        for (size_t i = 0; i < n; ++i) {
            const uint32_t a = kNormRuns[i].cp, b = (i + 1 < n) ? kNormRuns[i + 1].cp : 0x110000u;
            for (uint32_t cp = a; cp < b; ++cp) flat[cp] = kNormRuns[i].flags;
        }
        two_stage(flat, t);
        return t;
    }();
    return tables;
}

// ------------------------------------------------------------------------------------------
// Normalisers: pattern recognition (RegexNormalization) and blob parsing (CharsMapNormalization).
// ------------------------------------------------------------------------------------------
namespace {
struct KnownNormPattern { const char* pattern; uint8_t mask; uint8_t any; int32_t literal; uint8_t negate; uint8_t anchored; int char_group; int n_groups; };
// Search patterns the converter emits that match exactly one character (python/openvino_tokenizers/tokenizer_pipeline.py:230-278
// of the reference), after the legacy rewrites of src/regex_normalization.cpp:33-37.
const KnownNormPattern kNormPatterns[] = {
    {R"(([\x00-\x08\x0B\x0C\x0E-\x1F\x7F-\x9F\p{Cf}]))", NC_DEL, 0, -1, 0, 0, 1, 1},   // del_control_chars_regex
    {R"(\s)", NC_S, 0, -1, 0, 0, 0, 0},                                                   // replace_whitespace_regex
    {R"(([\p{Han}]))", NC_HAN, 0, -1, 0, 0, 1, 1},                                         // handle_chinese_chars_regex
    {R"(\p{Mn})", NC_MN, 0, -1, 0, 0, 0, 0},                                              // strip_accents_regex
    {R"(^(\S))", NC_S, 0, -1, 1, 1, 1, 1},                                                // add_prefix_whitespace_regex
    {R"(^([^ ]))", 0, 0, ' ', 1, 1, 1, 1},                                                // add_prefix_whitespace_to_not_whitespace_regex
    {R"((?:^)([\s\S]))", 0, 1, -1, 0, 1, 1, 1},                                           // prepend_regex
    {R"((^)([\s\S]))", 0, 1, -1, 0, 1, 2, 2},                                             // legacy (^)(.) / (^)(.+) after the rewrite
};
}  // namespace

int parse_regex_norm(const char* search, int64_t slen, const char* replace, int64_t rlen, int global_replace, HostNorm& out, std::string& err) {
    if (!search || slen < 0 || rlen < 0 || (rlen > 0 && !replace)) { err = "RegexNormalization: null pattern"; return B200TOK_E_INVALID; }
    std::string pat(search, (size_t)slen);
    if (pat == R"((^)(.))" || pat == R"((^)(.+))") pat = R"((^)([\s\S]))";       // src/regex_normalization.cpp:35-36
    NormRule& R = out.rule;
    R = NormRule{};
    R.kind = NORM_CLASS;
    R.literal_cp = -1;
    R.global = global_replace != 0;
    int char_group = 0, n_groups = 0;
    bool known = false;
    for (const auto& k : kNormPatterns) {
        if (pat != k.pattern) continue;
        R.mask = k.mask; R.any = k.any; R.literal_cp = k.literal; R.negate = k.negate; R.anchored = k.anchored;
        char_group = k.char_group; n_groups = k.n_groups;
        known = true;
        break;
    }
    if (!known) {
        // a single literal character that is not a regex metacharacter (e.g. " " -> metaspace, replace_spaces_metaspace)
        uint32_t cp = 0;
        const int l = pat.empty() ? 0 : utf8_strict_len((const uint8_t*)pat.data(), 0, (int)pat.size(), cp);
        static const std::string meta = R"(\^$.|?*+()[]{})";
        if (l > 0 && (size_t)l == pat.size() && !(cp < 0x80 && meta.find((char)cp) != std::string::npos)) { R.literal_cp = (int32_t)cp; known = true; }
    }
    if (!known) {
        err = "RegexNormalization: search pattern `" + pat + "` is not one of the single-character patterns this build runs on the GPU";
        return B200TOK_E_UNSUPPORTED;
    }
    // replacement: \N was already rewritten to $N by the reference (src/regex_normalization.cpp:19-31); do the same here
    std::string rep(replace ? replace : "", (size_t)rlen);
    for (char d = '1'; d <= '9'; ++d) {
        const std::string from = std::string("\\") + d, to = std::string("$") + d;
        size_t pos = 0;
        while ((pos = rep.find(from, pos)) != std::string::npos) { rep.replace(pos, from.size(), to); pos += to.size(); }
    }
    std::string pre, post;
    bool seen_char = false;
    for (size_t i = 0; i < rep.size();) {
        if (rep[i] != '$') { (seen_char ? post : pre).push_back(rep[i++]); continue; }
        if (i + 1 < rep.size() && rep[i + 1] == '$') { (seen_char ? post : pre).push_back('$'); i += 2; continue; }
        size_t j = i + 1;
        const bool brace = j < rep.size() && rep[j] == '{';
        if (brace) ++j;
        size_t d0 = j;
        long g = 0;
        while (j < rep.size() && rep[j] >= '0' && rep[j] <= '9') { g = g * 10 + (rep[j] - '0'); if (g > 1000) break; ++j; }
        if (j == d0 || (brace && (j >= rep.size() || rep[j] != '}'))) { err = "RegexNormalization: unsupported replacement syntax `" + rep + "`"; return B200TOK_E_UNSUPPORTED; }
        if (brace) ++j;
        if (g > n_groups) { err = "RegexNormalization: replacement refers to group " + std::to_string(g) + " which the pattern does not have"; return B200TOK_E_UNSUPPORTED; }
        if (g == 0 || g == char_group) {
            if (seen_char) { err = "RegexNormalization: the replacement may refer to the matched character once"; return B200TOK_E_UNSUPPORTED; }
            seen_char = true;
        }   // any other group of these patterns is the empty (^) group
        i = j;
    }
    if (pre.size() > sizeof(R.pre) || post.size() > sizeof(R.post)) { err = "RegexNormalization: replacement literal longer than 16 bytes"; return B200TOK_E_UNSUPPORTED; }
    R.keep = seen_char;
    out.atab.assign(256, 0);
    for (int a = 0; a < 128; ++a) { out.atab[(size_t)a] = (uint8_t)a; out.atab[128 + (size_t)a] = host_norm_class_tables().ascii[(size_t)a]; }
    R.pre_len = (uint8_t)pre.size(); R.post_len = (uint8_t)post.size();
    std::memcpy(R.pre, pre.data(), pre.size());
    std::memcpy(R.post, post.data(), post.size());
    return B200TOK_OK;
}

// What every ASCII byte becomes after all ops of a chain: a byte, NT_DEL (dropped) or NT_GENERAL (some op does more than
// map / drop it).  false: the chain has an op whose effect depends on the position in the string (anchored / first match
// only), so there is no composed table.  Strings made only of bytes with a simple fate take kernels_norm.cuh:compose_kernel.
bool compose_norm_chain(const HostNorm* const* ops, int n_ops, uint8_t* T) {
    for (int a = 0; a < 128; ++a) T[a] = (uint8_t)a;
    for (int k = 0; k < n_ops; ++k) {
        const HostNorm& h = *ops[k];
        const NormRule& R = h.rule;
        if (R.kind == NORM_CLASS && (R.anchored || !R.global)) return false;
        uint8_t A[128];
        for (int a = 0; a < 128; ++a) {
            if (R.kind == NORM_CHARSMAP) { A[a] = (h.atab[128 + (size_t)a] & (NA_COMPLEX | NA_ASCII_KIDS)) ? (uint8_t)NT_GENERAL : h.atab[(size_t)a]; continue; }
            const bool in_class = R.any || (R.literal_cp >= 0 ? a == R.literal_cp : (h.atab[128 + (size_t)a] & R.mask) != 0);
            if (in_class == (R.negate != 0)) { A[a] = (uint8_t)a; continue; }
            const int l = R.pre_len + (R.keep ? 1 : 0) + R.post_len;
            const uint8_t r = R.pre_len ? R.pre[0] : R.keep ? (uint8_t)a : R.post_len ? R.post[0] : 0;
            A[a] = l == 0 ? (uint8_t)NT_DEL : (l == 1 && r < 0x80) ? r : (uint8_t)NT_GENERAL;
        }
        for (int a = 0; a < 128; ++a) if (T[a] < 0x80) T[a] = A[T[a]];
    }
    return true;
}

int parse_charsmap(const uint8_t* blob, int64_t len, int add_dummy_prefix, int remove_extra_whitespaces, int escape_whitespaces, HostNorm& out, std::string& err) {
    if (len < 0 || (len > 0 && !blob)) { err = "CharsMapNormalization: null charsmap"; return B200TOK_E_INVALID; }
    if (add_dummy_prefix || remove_extra_whitespaces || escape_whitespaces) {
        err = "CharsMapNormalization: add_dummy_prefix / remove_extra_whitespaces / escape_whitespaces are not built for the GPU path (the converter's NormalizeUnicode / CaseFold steps leave them off)";
        return B200TOK_E_UNSUPPORTED;
    }
    NormRule& R = out.rule;
    R = NormRule{};
    R.kind = NORM_CHARSMAP;
    R.literal_cp = -1;
    out.units.clear(); out.normalized.clear();
    out.atab.assign(256, 0);
    for (int a = 0; a < 128; ++a) out.atab[(size_t)a] = (uint8_t)a;
    if (len == 0) return B200TOK_OK;                    // empty charsmap = identity (with U+FFFD for malformed bytes)
    uint32_t tsz = 0;
    if (len < 4) { err = "CharsMapNormalization: charsmap blob is truncated"; return B200TOK_E_INVALID; }
    std::memcpy(&tsz, blob, 4);
    if ((int64_t)tsz + 4 > len || (tsz & 3u) || tsz < 4) { err = "CharsMapNormalization: charsmap blob has a bad trie size"; return B200TOK_E_INVALID; }
    out.units.resize(tsz / 4);
    std::memcpy(out.units.data(), blob + 4, tsz);
    out.normalized.assign(blob + 4 + tsz, blob + len);
    // ASCII shortcuts: what the rule starting with byte a does, and whether longer rules start with it
    const std::vector<uint32_t>& U = out.units;
    auto offset = [](uint32_t u) { return (u >> 10) << ((u & (1u << 9)) >> 6); };
    auto label = [](uint32_t u) { return u & ((1u << 31) | 0xFFu); };
    const uint32_t root = offset(U[0]);
    for (uint32_t a = 0; a < 128; ++a) {
        uint32_t pos = root ^ a;
        if (pos >= U.size() || label(U[pos]) != a) continue;            // no rule starts with a
        const uint32_t u = U[pos];
        pos ^= offset(u);
        uint8_t fl = 0;
        if ((u >> 8) & 1u) {
            const uint32_t v = pos < U.size() ? (U[pos] & 0x7FFFFFFFu) : 0xFFFFFFFFu;
            if (v + 1 < out.normalized.size() && out.normalized[v] && out.normalized[v] < 0x80 && out.normalized[v + 1] == 0) out.atab[a] = out.normalized[v];
            else fl |= NA_COMPLEX;
        }
        for (uint32_t c = 1; c < 256; ++c) {
            const uint32_t q = pos ^ c;
            if (q < U.size() && label(U[q]) == c) fl |= c < 0x80 ? NA_ASCII_KIDS : NA_OTHER_KIDS;
        }
        out.atab[128 + a] = fl;
    }
    return B200TOK_OK;
}


// ------------------------------------------------------------------------------------------
// Flattened trie.
// ------------------------------------------------------------------------------------------
void HostTrie::build(const std::vector<std::pair<std::string, int32_t>>& entries) {
    struct Node { std::map<uint8_t, int32_t> kids; int32_t value = -1; };
    std::vector<Node> nodes(1);
    for (const auto& [key, id] : entries) {
        if (key.empty()) continue;   // the root's value is never consulted (utils.cpp:524-535)
        int32_t cur = 0;
        for (unsigned char c : key) {
            auto it = nodes[cur].kids.find(c);
            if (it == nodes[cur].kids.end()) {
                const int32_t nn = (int32_t)nodes.size();
                nodes[cur].kids.emplace(c, nn);
                nodes.emplace_back();
                cur = nn;
            } else {
                cur = it->second;
            }
        }
        nodes[cur].value = id;
    }
    const size_t N = nodes.size();
    first.assign(N + 1, 0);
    value.assign(N, -1);
    edge_byte.clear();
    edge_child.clear();
    root_child.assign(256, -1);
    for (size_t i = 0; i < N; ++i) {
        first[i] = (int32_t)edge_byte.size();
        value[i] = nodes[i].value;
        for (const auto& [b, ch] : nodes[i].kids) {
            edge_byte.push_back(b);
            edge_child.push_back(ch);
            if (i == 0) root_child[b] = ch;
        }
    }
    first[N] = (int32_t)edge_byte.size();
    if (edge_byte.empty()) { edge_byte.push_back(0); edge_child.push_back(-1); }  // keep device buffers non-empty
}

void HostTrie::build_rank() {
    const size_t N = value.size();
    std::vector<int32_t> order, newid(N, -1);       // breadth-first order over the CSR trie; node 0 = root
    order.reserve(N);
    order.push_back(0);
    newid[0] = 0;
    for (size_t q = 0; q < order.size(); ++q) {
        const int32_t n = order[q];
        for (int32_t e = first[(size_t)n]; e < first[(size_t)n + 1]; ++e) {
            const int32_t ch = edge_child[(size_t)e];
            if (ch < 0) continue;
            newid[(size_t)ch] = (int32_t)order.size();
            order.push_back(ch);
        }
    }
    rank_nodes.assign(order.size(), RankNode{});
    for (size_t q = 0; q < order.size(); ++q) {
        const int32_t n = order[q];
        RankNode& r = rank_nodes[q];
        r.value = value[(size_t)n];
        r.base = -1;
        for (int32_t e = first[(size_t)n]; e < first[(size_t)n + 1]; ++e) {       // edges are sorted by byte
            const int32_t ch = edge_child[(size_t)e];
            if (ch < 0) continue;
            if (r.base < 0) r.base = newid[(size_t)ch];
            r.bits[edge_byte[(size_t)e] >> 5] |= 1u << (edge_byte[(size_t)e] & 31);
        }
        int c = 0;
        for (int w = 0; w < 8; ++w) { r.cum[w] = (uint8_t)c; c += __builtin_popcount(r.bits[w]); }   // (at most 255 children before the last word)
        if (r.base < 0) r.base = 0;
    }
    rank_root.assign(256, -1);
    for (int b = 0; b < 256; ++b) if (root_child[(size_t)b] >= 0) rank_root[(size_t)b] = newid[(size_t)root_child[(size_t)b]];
    // one-byte values and the two-byte jump table (tok_core.cuh RankJump): the first two steps of every walk, precomputed
    rank_val1.assign(256, -1);
    rank_jump.assign(128 * 128, RankJump{-1, 0u});
    for (int b0 = 0; b0 < 256; ++b0) {
        const int32_t n1 = rank_root[(size_t)b0];
        if (n1 < 0) continue;
        const int32_t v1 = rank_nodes[(size_t)n1].value;
        rank_val1[(size_t)b0] = v1;
        if (b0 >= 128) continue;
        for (int b1 = 0; b1 < 128; ++b1) {
            RankJump& j = rank_jump[(size_t)(b0 * 128 + b1)];
            const int32_t n2 = rank_child(rank_nodes[(size_t)n1], (uint32_t)b1);
            int32_t found = v1;
            uint32_t len = v1 != -1 ? 1u : 0u;
            if (n2 >= 0 && rank_nodes[(size_t)n2].value != -1) { found = rank_nodes[(size_t)n2].value; len = 2u; }
            j.node2 = n2;
            j.info = kJumpHas1 | (len << 24) | (uint32_t)((found + 1) & 0xFFFFFF);
        }
    }
    if (value.size() >= (1u << 24) - 2) rank_jump.clear();      // (values must fit 24 bits; never the case for a tokenizer vocabulary)
}

#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("avx2")))
#endif
static bool contiguous_batch_impl(const int32_t* __restrict rb, const int32_t* __restrict re, const int32_t* __restrict eb,
                                  const int32_t* __restrict ee, int64_t B, int64_t E, int32_t n) {
    int acc = 0;
    for (int64_t r = 0; r + 1 < B; ++r) acc |= (rb[r + 1] != re[r]) | (re[r] < rb[r]);
    for (int64_t p = 0; p < E; ++p) acc |= (eb[p] < 0) | (ee[p] < eb[p]) | (ee[p] > n);
    for (int64_t p = 0; p + 1 < E; ++p) acc |= (eb[p + 1] < ee[p]);
    return acc == 0;
}
bool contiguous_batch(const int32_t* rb, const int32_t* re, const int32_t* eb, const int32_t* ee, int64_t B, int64_t E, int64_t N) {
    if (B <= 0 || N >= (1ll << 31) || rb[0] != 0 || re[B - 1] != E || re[B - 1] < rb[B - 1]) return false;
#if defined(__x86_64__) && defined(__GNUC__)
    if (!__builtin_cpu_supports("avx2")) {           // same predicate, baseline instruction set
        for (int64_t r = 0; r + 1 < B; ++r) if (rb[r + 1] != re[r] || re[r] < rb[r]) return false;
        for (int64_t p = 0; p < E; ++p) if (eb[p] < 0 || ee[p] < eb[p] || ee[p] > N || (p + 1 < E && eb[p + 1] < ee[p])) return false;
        return true;
    }
#endif
    return contiguous_batch_impl(rb, re, eb, ee, B, E, (int32_t)N);
}

static std::string str_at(const b200tok_strings& s, int64_t i) {
    return std::string((const char*)s.chars + s.begins[i], (const char*)s.chars + s.ends[i]);
}

static bool check_strings(const b200tok_strings& s, const char* what, std::string& err, bool allow_empty = true) {
    if (s.n < 0 || (s.n > 0 && (!s.begins || !s.ends)) || (!allow_empty && s.n == 0)) { err = std::string("bad string tensor: ") + what; return false; }
    for (int64_t i = 0; i < s.n; ++i)
        if (s.begins[i] < 0 || s.ends[i] < s.begins[i] || s.ends[i] > s.n_chars) { err = std::string("string offsets out of range in ") + what; return false; }
    return true;
}

// ------------------------------------------------------------------------------------------
// BPE tables.
// ------------------------------------------------------------------------------------------
int build_bpe(const b200tok_bpe_desc& d, HostBpe& out, std::string& err) {
    if (!check_strings(d.vocab, "vocab", err) || !check_strings(d.merges_left, "merges", err)) return B200TOK_E_INVALID;
    const bool pairs = d.merges_right.begins != nullptr;
    if (pairs && (!check_strings(d.merges_right, "merges(right)", err) || d.merges_right.n != d.merges_left.n)) {
        if (err.empty()) err = "left/right merge tensors differ in length";
        return B200TOK_E_INVALID;
    }
    if (d.added_tokens.n > 0 && (!check_strings(d.added_tokens, "added tokens", err) || !d.added_ids)) return B200TOK_E_INVALID;

    // bpe_tokenizer.cpp:62-66 (std::map::insert: first occurrence of a key wins)
    std::map<std::string, int32_t> added;
    for (int64_t i = 0; i < d.added_tokens.n; ++i) added.insert({str_at(d.added_tokens, i), d.added_ids[i]});
    // :77-82 (insert_or_assign: last id wins)
    std::unordered_map<std::string, int32_t> vocab;
    vocab.reserve((size_t)(d.vocab.n + d.added_tokens.n) * 2);
    for (int64_t i = 0; i < d.vocab.n; ++i) vocab[str_at(d.vocab, i)] = (int32_t)i;
    // :110-114 (insert: an existing key keeps its id)
    for (const auto& kv : added) vocab.insert({kv.first, kv.second});

    const std::string unk(d.unk_token ? d.unk_token : "", d.unk_token ? (size_t)d.unk_token_len : 0);
    out.unk_id = -1;
    if (auto it = vocab.find(unk); it != vocab.end()) out.unk_id = it->second;   // :353-355
    out.end_suffix.assign(d.end_suffix ? d.end_suffix : "", d.end_suffix ? (size_t)d.end_suffix_len : 0);

    const int64_t M = d.merges_left.n;
    out.n_merges = M;
    if (M >= (1ll << 20)) { err = "BPE: more than 2^20 merges is not supported by the packed (rank,birth) key"; return B200TOK_E_UNSUPPORTED; }
    // :361-373 — key (left_id,right_id); a later duplicate key overwrites (bpe_tokenizer.hpp:61-65)
    std::unordered_map<uint64_t, std::pair<int32_t, int32_t>> merges;
    merges.reserve((size_t)M * 2);
    std::vector<std::string> products;
    products.reserve((size_t)M);
    std::unordered_map<int32_t, int> product_count;
    for (int64_t i = 0; i < M; ++i) {
        std::string l, r;
        if (pairs) { l = str_at(d.merges_left, i); r = str_at(d.merges_right, i); }
        else {   // :91-96 split on the first ' '
            const std::string m = str_at(d.merges_left, i);
            const size_t sp = m.find(' ');
            l = m.substr(0, sp);
            r = sp == std::string::npos ? m : m.substr(sp + 1);
        }
        auto li = vocab.find(l), ri = vocab.find(r);
        std::string joined = l + r;
        auto ji = vocab.find(joined);
        if (li == vocab.end() || ri == vocab.end() || ji == vocab.end()) {   // vocab.at() throws in the reference
            err = "merge #" + std::to_string(i) + " refers to a token that is not in the vocab";
            return B200TOK_E_VOCAB;
        }
        merges[((uint64_t)(uint32_t)li->second << 32) | (uint32_t)ri->second] = {(int32_t)i, ji->second};
        if (++product_count[ji->second] == 2) ++out.n_duplicate_products;
        products.push_back(std::move(joined));
    }
    for (const auto& p : products) vocab.erase(p);   // :375-377

    // trie over what is left (:382-386) and the per-byte shortcuts
    std::vector<std::pair<std::string, int32_t>> entries;
    entries.reserve(vocab.size());
    for (const auto& kv : vocab) entries.emplace_back(kv.first, kv.second);
    out.trie.build(entries);
    out.byte_sym.assign(256, -1);
    out.byte_miss.assign(256, -1);
    out.bytes_only = true;
    for (int c = 0; c < 256; ++c) {
        const int32_t node = out.trie.root_child[c];
        if (node >= 0) {
            const bool has_kids = out.trie.first[node + 1] > out.trie.first[node];
            if (has_kids) { out.byte_sym[c] = kSymWalk; out.bytes_only = false; }
            else out.byte_sym[c] = out.trie.value[node];   // a leaf always carries a value
        }
        int32_t miss = -1;
        if (d.byte_fallback) {   // :242-248, looked up in the post-erase vocab
            char name[8];
            std::snprintf(name, sizeof(name), "<0x%02X>", (unsigned)c);
            if (auto it = vocab.find(name); it != vocab.end()) miss = it->second;
        }
        if (miss == -1 && out.unk_id != -1) miss = out.unk_id;   // :250-254 (fuse_unk never suppresses: it compares with -1)
        out.byte_miss[c] = miss;
    }

    // device hash table: power-of-two capacity, load factor <= 0.25 (an unsuccessful linear probe then ends after ~1.4 slots)
    size_t cap = 16;
    while (cap < merges.size() * 4 + 2) cap <<= 1;
    out.slots.assign(cap, MergeSlot{kEmptyKey, kEmptyKey, kNoRank, -1});
    out.mask = (uint32_t)(cap - 1);
    out.rank_newid.assign((size_t)std::max<int64_t>(M, 1), -1);
    for (const auto& kv : merges) {
        const uint32_t l = (uint32_t)(kv.first >> 32), r = (uint32_t)kv.first;
        out.rank_newid[(size_t)kv.second.first] = kv.second.second;
        uint32_t h = merge_hash(l, r) & out.mask;
        while (out.slots[h].left != kEmptyKey) h = (h + 1) & out.mask;
        out.slots[h] = MergeSlot{l, r, kv.second.first, kv.second.second};
    }
    out.max_id = 0;
    for (const auto& kv : vocab) out.max_id = std::max<int64_t>(out.max_id, kv.second);
    for (int32_t v : out.rank_newid) out.max_id = std::max<int64_t>(out.max_id, v);
    out.newid_base = M > 0 ? out.rank_newid[0] : -1;
    for (int64_t r = 0; r < M && out.newid_base >= 0; ++r)
        if (out.rank_newid[(size_t)r] != out.newid_base + (int32_t)r) out.newid_base = -1;
    // direct table for the initial pairs of one-byte symbols: sym1(b) is what the position-parallel symbolisation
    // assigns to byte b when no longer token starts there (the one-byte token, else the byte-fallback / unk id)
    out.pair_rank.assign(65536, kNoKey);
    int32_t sym1[256];
    for (int c = 0; c < 256; ++c) {
        const int32_t node = out.trie.root_child[c];
        int32_t id = (node >= 0) ? out.trie.value[node] : -1;
        if (id < 0) id = out.byte_miss[c];
        sym1[c] = id;
    }
    for (int b0 = 0; b0 < 256; ++b0)
        for (int b1 = 0; b1 < 256; ++b1) {
            if (sym1[b0] < 0 || sym1[b1] < 0) continue;
            auto it = merges.find(((uint64_t)(uint32_t)sym1[b0] << 32) | (uint32_t)sym1[b1]);
            if (it != merges.end()) out.pair_rank[(size_t)(b0 << 8 | b1)] = (uint32_t)it->second.first;
        }
    // [0, 512): mergeable ASCII byte pairs; [512, 512 + 2048): for every first byte, the second bytes that continue a
    // token of the symbolisation trie (bit b1 of words [512 + 8 * b0, +8)) — lets the window kernel skip trie walks
    // [2560, 2560 + 8192): ranks of the ASCII pairs as u16 (index b0 << 7 | b1; 0xFFFF = none, 0xFFFE = rank too large for
    // 16 bits, look in pair_rank) — 32 KB, stays L1-resident in the window kernel.  The table is allocated for a first byte
    // up to 255 (another 32 KB of 0xFFFF that is never touched in normal operation): the window kernel indexes it with the byte
    // BEFORE the window's first position too, whose rank is never used (a window starts at a piece start) but which may be any
    // byte of the preceding text, or shared-memory garbage at the start of an element — found by compute-sanitizer.
    out.pair_bits.assign(512 + 2048 + 16384, 0u);
    {
        uint16_t* r16 = reinterpret_cast<uint16_t*>(out.pair_bits.data() + 2560);
        std::fill(r16 + 16384, r16 + 32768, (uint16_t)0xFFFF);
        for (int b0 = 0; b0 < 128; ++b0)
            for (int b1 = 0; b1 < 128; ++b1) {
                const uint32_t r = out.pair_rank[(size_t)(b0 << 8 | b1)];
                r16[b0 << 7 | b1] = r == kNoKey ? (uint16_t)0xFFFF : r >= 0xFFFEu ? (uint16_t)0xFFFE : (uint16_t)r;
            }
    }
    for (int b0 = 0; b0 < 256; ++b0) {
        const int32_t node = out.trie.root_child[b0];
        if (node < 0) continue;
        for (int32_t e = out.trie.first[node]; e < out.trie.first[node + 1]; ++e) {
            const int b1 = out.trie.edge_byte[e];
            out.pair_bits[(size_t)(512 + 8 * b0 + (b1 >> 5))] |= 1u << (b1 & 31);
        }
    }
    for (int b0 = 0; b0 < 128; ++b0)
        for (int b1 = 0; b1 < 128; ++b1)
            if (out.pair_rank[(size_t)(b0 << 8 | b1)] != kNoKey) out.pair_bits[(size_t)(b0 << 2 | b1 >> 5)] |= 1u << (b1 & 31);
    return B200TOK_OK;
}

// ------------------------------------------------------------------------------------------
// WordPiece tries (wordpiece_tokenizer.cpp:61-71).
// ------------------------------------------------------------------------------------------
int build_wordpiece(const b200tok_wordpiece_desc& d, HostWordpiece& out, std::string& err) {
    if (!check_strings(d.vocab, "vocab", err)) return B200TOK_E_INVALID;
    const std::string ind(d.suffix_indicator ? d.suffix_indicator : "", d.suffix_indicator ? (size_t)d.suffix_indicator_len : 0);
    std::vector<std::pair<std::string, int32_t>> r, s;
    for (int64_t i = 0; i < d.vocab.n; ++i) {
        std::string w = str_at(d.vocab, i);
        if (w.compare(0, ind.size(), ind) == 0 && w.size() >= ind.size()) s.emplace_back(w.substr(ind.size()), (int32_t)i);
        else r.emplace_back(std::move(w), (int32_t)i);
    }
    out.root.build(r);
    out.sub.build(s);
    out.root.build_rank();
    out.sub.build_rank();
    out.max_bytes = d.max_bytes_per_word;
    return B200TOK_OK;
}

// ------------------------------------------------------------------------------------------
// VocabEncoder table (vocab_encoder.cpp:63-78: insert => first duplicate key wins).
// ------------------------------------------------------------------------------------------
int build_vocabenc(const b200tok_vocabenc_desc& d, HostVocabEnc& out, std::string& err) {
    if (!check_strings(d.keys, "vocab keys", err) || (d.keys.n > 0 && !d.values)) { if (err.empty()) err = "missing values"; return B200TOK_E_INVALID; }
    size_t cap = 16;
    while (cap < (size_t)d.keys.n * 2 + 2) cap <<= 1;
    out.slots.assign(cap, VocabEncSlot{0, 0, -1, 0});
    out.mask = (uint32_t)(cap - 1);
    out.key_bytes.assign(d.keys.chars, d.keys.chars + d.keys.n_chars);
    if (out.key_bytes.empty()) out.key_bytes.push_back(0);
    out.max_len = 0;
    for (int64_t i = 0; i < d.keys.n; ++i) {
        const int32_t b = d.keys.begins[i], len = d.keys.ends[i] - b;
        const uint64_t h = fnv1a64(d.keys.chars + b, len);
        uint32_t k = (uint32_t)h & out.mask;
        bool dup = false;
        while (out.slots[k].len >= 0) {
            const auto& sl = out.slots[k];
            if (sl.hash == h && sl.len == len && std::memcmp(d.keys.chars + sl.begin, d.keys.chars + b, (size_t)len) == 0) { dup = true; break; }
            k = (k + 1) & out.mask;
        }
        if (dup) continue;
        const int64_t v = d.values_are_i64 ? ((const int64_t*)d.values)[i] : (int64_t)((const int32_t*)d.values)[i];
        out.slots[k] = VocabEncSlot{h, b, len, v};
        out.max_len = std::max(out.max_len, len);
    }
    return B200TOK_OK;
}

// ------------------------------------------------------------------------------------------
// Split pattern recognition.
// ------------------------------------------------------------------------------------------
static const char* kGpt2 = R"('s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+)";
static const char* kGpt2Digits = R"('s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+|\p{N}| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+)";
static const char* kLlama3 = R"((?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+)";
static const char* kBertPunct =
    R"([!-/]|[:-@]|[\[-`]|[{-~]|[\p{P}]|[\x{4E00}-\x{9FFF}]|[\x{3400}-\x{4DBF}]|[\x{20000}-\x{2A6DF}]|[\x{2A700}-\x{2B73F}]|[\x{2B740}-\x{2B81F}]|[\x{2B820}-\x{2CEAF}]|[\x{F900}-\x{FAFF}]|[\x{2F800}-\x{2FA1F}])";

int parse_split(const b200tok_regexsplit_desc& d, HostSplit& out, std::string& err) {
    if (!d.pattern || d.pattern_len < 0 || !d.behaviour) { err = "RegexSplit: missing pattern or behaviour"; return B200TOK_E_INVALID; }
    const std::string pat(d.pattern, (size_t)d.pattern_len), beh(d.behaviour);
    // modes: src/regex_split.cpp:16-22
    if (beh == "remove") out.mode = MODE_REMOVED;
    else if (beh == "isolate" || beh == "contiguous") out.mode = MODE_ISOLATED;
    else if (beh == "mergedwithprevious") out.mode = MODE_MERGED_PREV;
    else if (beh == "mergedwithnext") out.mode = MODE_MERGED_NEXT;
    else { err = "RegexSplit doesn't support unknown split mode: " + beh; return B200TOK_E_INVALID; }
    if (!(d.max_splits == -1 || d.max_splits > 0)) {   // src/regex_split.cpp:114-117
        err = "RegexSplit max_splits attribute must be greater then `0` or equal to `-1`, got " + std::to_string(d.max_splits);
        return B200TOK_E_INVALID;
    }
    out.invert = d.invert != 0;
    out.max_splits = d.max_splits;
    out.pattern = pat;
    out.repeat = beh == "contiguous" && !pat.empty() && pat.back() != '+';   // :33-37
    out.spec = SplitSpec{};
    auto cls = [&](uint8_t mask) { out.spec.pat = PAT_CLASS_CHAR; out.spec.class_mask = mask; return B200TOK_OK; };
    if (pat == kGpt2) { out.spec.pat = PAT_GPT2; return B200TOK_OK; }
    if (pat == kGpt2Digits) { out.spec.pat = PAT_GPT2_DIGITS; return B200TOK_OK; }
    if (pat == kLlama3) { out.spec.pat = PAT_LLAMA3; return B200TOK_OK; }
    if (pat == R"(\s+)") { out.spec.pat = PAT_WS; return B200TOK_OK; }
    if (pat == kBertPunct) { out.spec.pat = PAT_BERT_PUNCT; return B200TOK_OK; }
    if (pat == R"(\w+|[^\w\s]+)") { out.spec.pat = PAT_WORD_OR_PUNCT; return B200TOK_OK; }
    if (pat == ".") { out.spec.pat = PAT_ANYCHAR; return B200TOK_OK; }
    if (pat == R"(\p{N})" || pat == R"(\p{Nd}|\p{Nl}|\p{No})") return cls(C_N);
    if (pat == R"(\p{P})" || pat == R"([\p{P}])") return cls(C_P);
    if (pat == R"(\p{Nd}|\p{Nl}|\p{No}|\p{P})" || pat == R"(\p{P}|\p{Nd}|\p{Nl}|\p{No})") return cls(C_N | C_P);
    if (pat == R"(\p{L})") return cls(C_L);
    // a literal: no regex metacharacters at all (metaspace "▁", " ", ...)
    if (!pat.empty() && pat.size() <= sizeof(out.spec.lit) && pat.find_first_of("\\^$.|?*+()[]{}") == std::string::npos) {
        out.spec.pat = PAT_LITERAL;
        out.spec.lit_len = (uint8_t)pat.size();
        std::memcpy(out.spec.lit, pat.data(), pat.size());
        return B200TOK_OK;
    }
    // anything else: compile it for the regex machine (regex_vm.cuh); patterns outside its syntax are refused, never approximated
    const int rc = compile_regex(pat, out.vm, err);
    if (rc == B200TOK_OK) out.spec.pat = PAT_VM;
    return rc;
}

// ------------------------------------------------------------------------------------------
// SpecialTokensSplit pattern.
// ------------------------------------------------------------------------------------------
int parse_special(const char* pattern, int64_t len, HostSpecial& out, std::string& err) {
    if (!pattern || len < 0) { err = "SpecialTokensSplit: missing pattern"; return B200TOK_E_INVALID; }
    const std::string pat(pattern, (size_t)len);
    out.pattern = pat;
    out.groups.clear();
    auto unsupported = [&](const char* why) {
        err = std::string("SpecialTokensSplit: pattern is not an alternation of (?:\\s*)?(literal|...)(?:\\s*)? groups (") + why + "): " + pat;
        return B200TOK_E_UNSUPPORTED;
    };
    static const std::string kWs = "(?:\\s*)";
    size_t i = 0;
    while (i < pat.size()) {
        HostSpecialGroup g;
        if (pat.compare(i, kWs.size(), kWs) == 0) { g.strip_left = true; i += kWs.size(); }
        if (i >= pat.size() || pat[i] != '(' || (i + 1 < pat.size() && pat[i + 1] == '?')) return unsupported("expected a capture group");
        ++i;
        std::string tok;
        bool closed = false;
        while (i < pat.size()) {
            const char c = pat[i];
            if (c == '\\') {
                if (i + 1 >= pat.size()) return unsupported("dangling backslash");
                const unsigned char n = (unsigned char)pat[i + 1];
                if (std::isalnum(n)) return unsupported("escape sequence with a meaning");   // \s, \d, \x.. are not literals
                tok.push_back((char)n);
                i += 2;
            } else if (c == '|') { g.tokens.push_back(tok); tok.clear(); ++i; }
            else if (c == ')') { g.tokens.push_back(tok); closed = true; ++i; break; }
            else if (std::strchr("^$.?*+([{}]", c)) return unsupported("unescaped metacharacter inside a group");
            else { tok.push_back(c); ++i; }
        }
        if (!closed) return unsupported("unterminated group");
        if (pat.compare(i, kWs.size(), kWs) == 0) { g.strip_right = true; i += kWs.size(); }
        if (i < pat.size()) {
            if (pat[i] != '|') return unsupported("expected | between groups");
            ++i;
            if (i >= pat.size()) return unsupported("empty alternative");
        }
        // neighbouring groups with the same flags behave like one group with the alternatives concatenated
        if (!out.groups.empty() && out.groups.back().strip_left == g.strip_left && out.groups.back().strip_right == g.strip_right)
            out.groups.back().tokens.insert(out.groups.back().tokens.end(), g.tokens.begin(), g.tokens.end());
        else out.groups.push_back(std::move(g));
    }
    if (out.groups.empty()) return unsupported("empty pattern");
    if ((int)out.groups.size() > kMaxSpecialGroups) return unsupported("too many groups");
    out.first.assign(8, 0u);
    out.ws_token = false;
    const HostClassTables& ct = host_class_tables();
    bool any_strip_left = false;
    for (auto& g : out.groups) {
        std::vector<std::pair<std::string, int32_t>> entries;
        for (size_t k = g.tokens.size(); k-- > 0;) {      // reversed: an earlier duplicate overwrites a later one
            if (g.tokens[k].empty()) return unsupported("empty alternative");   // would match the empty string everywhere
            entries.emplace_back(g.tokens[k], (int32_t)k);
        }
        g.trie.build(entries);
        for (const auto& t : g.tokens) {
            const unsigned char b = (unsigned char)t[0];
            out.first[b >> 5] |= 1u << (b & 31);
            if (g.strip_left && (b >= 0x80 || (ct.ascii[b] & C_S))) out.ws_token = true;
        }
        any_strip_left |= g.strip_left;
    }
    if (any_strip_left)      // whitespace can start a match: ASCII \s bytes and every multi-byte lead (classified exactly later)
        for (int b = 0; b < 256; ++b)
            if (b >= 0xC2 || (b < 0x80 && (ct.ascii[b] & C_S))) out.first[b >> 5] |= 1u << (b & 31);
    return B200TOK_OK;
}

}  // namespace b200tok
