// tok_core.cuh — the algorithmic core of the hot path, written once as __host__ __device__ inline
// functions over plain pointers.  The CUDA kernels (kernels.cu) call these on shared / global
// memory; tests/harness compiles the very same header with g++ so that the matching, merge and
// trie logic can be checked against the oracle in the CPU-only test tier (the harness is test
// code: the product library has no CPU execution path).
//
// Reference semantics restated here (paths relative to /root/reference):
//   * regex matching for the tokenizer patterns  — PCRE2 leftmost / ordered-alternation /
//     backtracking semantics as driven by src/regex_split.cpp:287-309 and src/utils.cpp:396-420
//   * BPE merge order (rank, push-sequence)       — src/bpe_tokenizer.cpp:166-172,269-323
//   * trie longest match                          — src/utils.cpp:517-538
//   * WordPiece word loop                         — src/wordpiece_tokenizer.cpp:96-130
#pragma once
#include <stdint.h>

#include "regex_vm.cuh"

#if defined(__CUDACC__)
#define B2_HD __host__ __device__ __forceinline__
#else
#define B2_HD inline
#endif

namespace b200tok {

// ------------------------------------------------------------------------------------------
// Character classes (PCRE2_UTF|PCRE2_UCP semantics; tables generated from PCRE2 itself).
// ------------------------------------------------------------------------------------------
enum : uint8_t {
    C_L = 1,      // \p{L}
    C_N = 2,      // \p{N}
    C_S = 4,      // \s
    C_P = 8,      // \p{P}
    C_W = 16,     // \w
    C_BP = 32,    // BERT "punctuation or CJK" class (tokenizer_pipeline.py:403-431)
    C_NL = 64,    // \r or \n
    C_CONT = 128  // UTF-8 continuation byte (class bits copied from the owning character)
};

struct ClassTables {
    const uint8_t* ascii;     // [128]
    const uint16_t* stage1;   // [0x1100]  cp >> 8 -> block
    const uint8_t* stage2;    // [n_blocks * 256]
};

B2_HD bool is_cont_byte(uint8_t b) { return (b & 0xC0) == 0x80; }

// Class of the character whose first byte is s[i]; `end` bounds the character's bytes.
// Malformed sequences (out of contract for the reference: PCRE2 JIT skips validation,
// SURVEY App. B item 6) are classified as "other" deterministically.
B2_HD uint8_t char_class(const uint8_t* s, int i, int end, const ClassTables& t) {
    const uint8_t b0 = s[i];
    if (b0 < 0x80) return t.ascii[b0];
    const int need = b0 >= 0xF0 ? 3 : b0 >= 0xE0 ? 2 : b0 >= 0xC0 ? 1 : -1;
    if (need < 0 || b0 >= 0xF8 || i + need >= end) return 0;
    uint32_t cp = need == 1 ? (b0 & 0x1Fu) : need == 2 ? (b0 & 0x0Fu) : (b0 & 0x07u);
    for (int k = 1; k <= need; ++k) {
        const uint8_t b = s[i + k];
        if (!is_cont_byte(b)) return 0;
        cp = (cp << 6) | (b & 0x3Fu);
    }
    if (cp >= 0x110000u) return 0;
    return t.stage2[(uint32_t)t.stage1[cp >> 8] * 256u + (cp & 255u)];
}

// ------------------------------------------------------------------------------------------
// Split patterns the GPU splitter implements.  Each is a direct transcription of the regex's
// alternatives in order (PCRE2 tries alternatives left to right and takes the first that
// matches, with greedy quantifiers and backtracking).
// ------------------------------------------------------------------------------------------
enum PatternId : int {
    PAT_NONE = 0,
    PAT_GPT2 = 1,          // 's|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+
    PAT_GPT2_DIGITS = 2,   // same with \p{N} in place of " ?\p{N}+"
    PAT_LLAMA3 = 3,        // (?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+
    PAT_WS = 4,            // \s+
    PAT_BERT_PUNCT = 5,    // [!-/]|[:-@]|[\[-`]|[{-~]|[\p{P}]|<CJK ranges>
    PAT_WORD_OR_PUNCT = 6, // \w+|[^\w\s]+
    PAT_LITERAL = 7,       // one literal string (metaspace etc.)
    PAT_ANYCHAR = 8,       // .
    PAT_CLASS_CHAR = 9,    // one character of a class: \p{N} == \p{Nd}|\p{Nl}|\p{No}, or \p{P}
    PAT_BERT_FUSED = 10,   // internal: \s+ (remove) followed by PAT_BERT_PUNCT (isolate), one pass
    PAT_VM = 11            // any other pattern inside the supported syntax: compiled to the regex machine of regex_vm.cuh
};

// Partition of characters into "kinds" such that every quantified class of the pattern is a
// maximal run of one kind.
B2_HD int kind_of(uint8_t cls, int pat, uint8_t class_mask) {
    switch (pat) {
    case PAT_GPT2: case PAT_GPT2_DIGITS: case PAT_LLAMA3:
        return (cls & C_L) ? 0 : (cls & C_N) ? 1 : (cls & C_S) ? 2 : 3;
    case PAT_WS: return (cls & C_S) ? 2 : 3;
    case PAT_BERT_PUNCT: return (cls & C_BP) ? 1 : 3;
    case PAT_BERT_FUSED: return (cls & C_S) ? 2 : (cls & C_BP) ? 1 : 3;
    case PAT_WORD_OR_PUNCT: return (cls & C_W) ? 0 : (cls & C_S) ? 2 : 3;
    case PAT_CLASS_CHAR: return (cls & class_mask) ? 1 : 3;
    default: return 3;
    }
}

struct SplitSpec {
    int pat;
    uint8_t class_mask;       // PAT_CLASS_CHAR
    uint8_t lit_len;          // PAT_LITERAL
    uint8_t lit[22];
    VmProgram vm;             // PAT_VM (pointers into host memory in the host harness, device memory in the kernels)
};

struct Match {
    int len;    // 0 => the pattern does not match at p
    int peek;   // 1 + highest byte index examined while deciding
    int drop;   // PAT_BERT_FUSED only: 1 => this match is a removed delimiter (whitespace)
};

// A context gives the matcher random access to one element [lo, end) of the chars buffer.
//   byte(i), cls(i): defined for lo <= i < lim()
//   run_end(i): end (exclusive, <= known()) of the maximal same-kind run starting at char i
//   last_nl(i): index of the last \r|\n inside the \s-run that contains i, at or after i; -1 if none
//   known(): everything below this index is exact; lim(): bytes below this index may be read
// ScanCtx computes everything by scanning (used for pieces that outgrow a shared-memory window,
// and by the host harness); the kernels' window context reads precomputed arrays.
struct ScanCtx {
    const uint8_t* s;
    int end;
    ClassTables t;
    int pat;
    uint8_t class_mask;
    B2_HD int known() const { return end; }
    B2_HD int lim() const { return end; }
    B2_HD uint8_t byte(int i) const { return s[i]; }
    B2_HD uint8_t cls(int i) const { return (i > 0 && is_cont_byte(s[i])) ? (uint8_t)C_CONT : char_class(s, i, end, t); }
    B2_HD int next(int i) const {
        ++i;
        while (i < end && is_cont_byte(s[i])) ++i;
        return i;
    }
    B2_HD int run_end(int i) const {
        const int k = kind_of(cls(i), pat, class_mask);
        int j = next(i);
        while (j < end && kind_of(cls(j), pat, class_mask) == k) j = next(j);
        return j;
    }
    B2_HD int last_nl(int i) const {
        int r = -1;
        for (int j = i; j < end; j = next(j)) {
            const uint8_t c = cls(j);
            if (!(c & C_S)) break;
            if (c & C_NL) r = j;
        }
        return r;
    }
};

template <class C>
B2_HD bool ctx_has(const C& c, int i, int& peek) {
    if (i + 1 > peek) peek = i + 1;
    return i < c.lim();
}

template <class C>
B2_HD int prev_char_start(const C& c, int e, int lo) {
    int j = e - 1;
    while (j > lo && (c.cls(j) & C_CONT)) --j;
    return j;
}

// \s+(?!\S) | \s+   at a whitespace character p  (shared by GPT-2 and Llama-3 patterns)
template <class C>
B2_HD Match match_ws_tail(const C& c, int p, int end, int peek) {
    const int e = c.run_end(p);
    if (e + 1 > peek) peek = e + 1;
    if (e >= end) return Match{e - p, peek, 0};          // (?!\S) holds at end of subject
    const int last = prev_char_start(c, e, p);
    if (last > p) return Match{last - p, peek, 0};       // give back one character
    return Match{e - p, peek, 0};                        // single whitespace char: plain \s+
}

template <class C>
B2_HD Match match_gpt2(const C& c, int p, int end, bool single_digits) {
    int peek = p + 1;
    const uint8_t b0 = c.byte(p);
    const uint8_t k0 = c.cls(p);
    if (b0 == '\'') {  // 's|'t|'re|'ve|'m|'ll|'d
        if (ctx_has(c, p + 1, peek) && p + 1 < end) {
            const uint8_t b1 = c.byte(p + 1);
            if (b1 == 's' || b1 == 't' || b1 == 'm' || b1 == 'd') return Match{2, peek, 0};
            if ((b1 == 'r' || b1 == 'v' || b1 == 'l') && ctx_has(c, p + 2, peek) && p + 2 < end) {
                const uint8_t b2 = c.byte(p + 2);
                if ((b1 == 'l') ? (b2 == 'l') : (b2 == 'e')) return Match{3, peek, 0};
            }
        }
    }
    int q = p;
    uint8_t kq = k0;
    if (b0 == ' ' && ctx_has(c, p + 1, peek) && p + 1 < end) {  // optional leading U+0020
        const uint8_t k1 = c.cls(p + 1);
        if (!(k1 & C_S) && !(single_digits && (k1 & C_N))) { q = p + 1; kq = k1; }
    }
    if (!(kq & C_S)) {
        if (single_digits && (kq & C_N)) {  // \p{N}
            const int e = c.next(q);
            return Match{e - p, peek > e ? peek : e, 0};
        }
        const int e = c.run_end(q);   //  ?\p{L}+ |  ?\p{N}+ |  ?[^\s\p{L}\p{N}]+
        if (e + 1 > peek) peek = e + 1;
        return Match{e - p, peek, 0};
    }
    return match_ws_tail(c, p, end, peek);
}

B2_HD bool ci_eq(uint8_t b, char lower) { return (b | 0x20) == (uint8_t)lower && ((b | 0x20) >= 'a'); }

template <class C>
B2_HD Match match_llama3(const C& c, int p, int end) {
    int peek = p + 1;
    const uint8_t b0 = c.byte(p);
    const uint8_t k0 = c.cls(p);
    if (b0 == '\'' && ctx_has(c, p + 1, peek) && p + 1 < end) {  // (?i:'s|'t|'re|'ve|'m|'ll|'d)
        const uint8_t b1 = c.byte(p + 1);
        if (ci_eq(b1, 's') || ci_eq(b1, 't') || ci_eq(b1, 'm') || ci_eq(b1, 'd')) return Match{2, peek, 0};
        if (ctx_has(c, p + 2, peek) && p + 2 < end) {
            const uint8_t b2 = c.byte(p + 2);
            if (b1 == 0xC5 && b2 == 0xBF) return Match{3, peek, 0};  // U+017F folds to 's' under PCRE2 caseless UTF
            if ((ci_eq(b1, 'r') || ci_eq(b1, 'v')) && ci_eq(b2, 'e')) return Match{3, peek, 0};
            if (ci_eq(b1, 'l') && ci_eq(b2, 'l')) return Match{3, peek, 0};
        }
    }
    // [^\r\n\p{L}\p{N}]?\p{L}+
    if (k0 & C_L) {
        const int e = c.run_end(p);
        if (e + 1 > peek) peek = e + 1;
        return Match{e - p, peek, 0};
    }
    if (!(k0 & (C_N | C_NL))) {
        const int q = c.next(p);
        if (ctx_has(c, q, peek) && q < end && (c.cls(q) & C_L)) {
            const int e = c.run_end(q);
            if (e + 1 > peek) peek = e + 1;
            return Match{e - p, peek, 0};
        }
    }
    // \p{N}{1,3}
    if (k0 & C_N) {
        const int re = c.run_end(p);
        if (re + 1 > peek) peek = re + 1;
        int e = c.next(p), n = 1;
        while (n < 3 && e < re) { e = c.next(e); ++n; }
        return Match{e - p, peek, 0};
    }
    //  ?[^\s\p{L}\p{N}]+[\r\n]*
    {
        int q = p;
        uint8_t kq = k0;
        if (b0 == ' ' && ctx_has(c, p + 1, peek) && p + 1 < end) { q = p + 1; kq = c.cls(q); }
        if (!(kq & (C_L | C_N | C_S))) {
            int e = c.run_end(q);
            while (ctx_has(c, e, peek) && e < end && (c.byte(e) == '\r' || c.byte(e) == '\n')) ++e;
            return Match{e - p, peek, 0};
        }
    }
    // here k0 is whitespace:  \s*[\r\n]+ | \s+(?!\S) | \s+
    {
        const int e = c.run_end(p);
        if (e + 1 > peek) peek = e + 1;
        const int j = c.last_nl(p);
        if (j >= 0) return Match{j + 1 - p, peek, 0};
    }
    return match_ws_tail(c, p, end, peek);
}

// Returns the match of `spec` anchored at character position p of the element ending at `end`.
template <class C>
B2_HD Match match_at(const C& c, const SplitSpec& spec, int p, int end) {
    switch (spec.pat) {
    case PAT_GPT2: return match_gpt2(c, p, end, false);
    case PAT_GPT2_DIGITS: return match_gpt2(c, p, end, true);
    case PAT_LLAMA3: return match_llama3(c, p, end);
    case PAT_WS: {
        if (!(c.cls(p) & C_S)) return Match{0, p + 1, 0};
        const int e = c.run_end(p);
        return Match{e - p, e + 1, 0};
    }
    case PAT_BERT_PUNCT: case PAT_CLASS_CHAR: {
        const uint8_t m = spec.pat == PAT_BERT_PUNCT ? (uint8_t)C_BP : spec.class_mask;
        if (!(c.cls(p) & m)) return Match{0, p + 1, 0};
        const int e = c.next(p);
        return Match{e - p, e, 0};
    }
    case PAT_BERT_FUSED: {
        const uint8_t k = c.cls(p);
        if (k & C_S) { const int e = c.run_end(p); return Match{e - p, e + 1, 1}; }
        if (k & C_BP) { const int e = c.next(p); return Match{e - p, e, 0}; }
        return Match{0, p + 1, 0};
    }
    case PAT_WORD_OR_PUNCT: {
        if (c.cls(p) & C_S) return Match{0, p + 1, 0};
        const int e = c.run_end(p);
        return Match{e - p, e + 1, 0};
    }
    case PAT_ANYCHAR: {
        if (c.byte(p) == '\n') return Match{0, p + 1, 0};
        const int e = c.next(p);
        return Match{e - p, e, 0};
    }
    case PAT_LITERAL: {
        int peek = p + 1;
        for (int k = 0; k < spec.lit_len; ++k) {
            if (!ctx_has(c, p + k, peek) || p + k >= end || c.byte(p + k) != spec.lit[k]) return Match{0, peek, 0};
        }
        return Match{(int)spec.lit_len, peek, 0};
    }
    case PAT_VM: {
        int len, peek;
        vm_match(c, spec.vm, p, end, len, peek);
        return Match{len, peek, 0};
    }
    default: return Match{0, p + 1, 0};
    }
}

// ------------------------------------------------------------------------------------------
// Closed form of the GPT-2 byte-level pattern in "isolate" mode.  The pattern tiles the subject, and
// whether a piece starts at a character depends only on a few neighbouring characters, so the
// sequential match loop collapses to a per-position predicate (checked against PCRE2 exhaustively
// in tests/test_core_host.py).  Inputs: b[] bytes, k[] per-byte classes where continuation bytes
// carry their owner's class bits plus C_CONT; [eb, ee) = the element.  Derivation (alternatives in
// order 's|'t|'re|'ve|'m|'ll|'d,  ?\p{L}+,  ?\p{N}+ (or \p{N}),  ?[^\s\p{L}\p{N}]+, \s+(?!\S), \s+):
//   * a whitespace run followed by a non-space gives up its last character (\s+(?!\S) backtracks);
//     if that character is U+0020 it becomes the optional prefix of the following run;
//   * runs of one kind (letter / number / other) are consumed whole, except that a contraction
//     ("'" reached as the start of a match, i.e. not preceded by an "other" character or by U+0020)
//     takes its one or two letters out of the following letter run.
// ------------------------------------------------------------------------------------------
// class accessors: per-byte class array, or (ASCII-only subjects) the 128-entry table applied to the byte itself
struct ClsArray { const uint8_t* k; B2_HD uint8_t operator()(int i) const { return k[i]; } };
struct ClsAsciiLut { const uint8_t* b; const uint8_t* lut; B2_HD uint8_t operator()(int i) const { return lut[b[i]]; } };

template <class KF>
B2_HD bool gpt2_apostrophe_t(const uint8_t* b, const KF& K, int q, int eb) {
    if (q == eb) return true;
    const uint8_t p = K(q - 1);
    if (p & (C_L | C_N)) return true;
    if (p & C_S) return b[q - 1] != 0x20;
    return false;
}
B2_HD int gpt2_contraction_len(const uint8_t* b, int q, int ee) {
    if (q + 1 >= ee) return 0;
    const uint8_t b1 = b[q + 1];
    if (b1 == 's' || b1 == 't' || b1 == 'm' || b1 == 'd') return 2;
    if (q + 2 < ee) {
        const uint8_t b2 = b[q + 2];
        if ((b1 == 'r' || b1 == 'v') && b2 == 'e') return 3;
        if (b1 == 'l' && b2 == 'l') return 3;
    }
    return 0;
}
template <class KF>
B2_HD bool gpt2_piece_starts_t(const uint8_t* b, const KF& K, int i, int eb, int ee, bool single_digits) {
    const uint8_t c = K(i);
    if (c & C_CONT) return false;
    if (i == eb) return true;
    const uint8_t p = K(i - 1);
    if (c & C_S) {
        if (!(p & C_S)) return true;
        int j = i + 1;
        while (j < ee && (K(j) & C_CONT)) ++j;
        return j < ee && !(K(j) & C_S);
    }
    if (single_digits && (c & C_N)) return true;
    if (b[i - 1] == 0x20) return false;
    if (c & C_N) return !(p & C_N);
    if (!(c & C_L)) return (p & (C_L | C_N | C_S)) != 0;          // "other": starts unless it continues an other-run
    bool inside = false, ends_here = false;
    if (b[i - 1] == '\'' && gpt2_apostrophe_t(b, K, i - 1, eb) && gpt2_contraction_len(b, i - 1, ee) >= 2) inside = true;
    if (i - 2 >= eb && b[i - 2] == '\'' && gpt2_apostrophe_t(b, K, i - 2, eb)) {
        const int cl = gpt2_contraction_len(b, i - 2, ee);
        if (cl == 3) inside = true;
        else if (cl == 2) ends_here = true;
    }
    if (i - 3 >= eb && b[i - 3] == '\'' && gpt2_apostrophe_t(b, K, i - 3, eb) && gpt2_contraction_len(b, i - 3, ee) == 3) ends_here = true;
    return !inside && (ends_here || !(p & C_L));
}
B2_HD bool gpt2_piece_starts_at(const uint8_t* b, const uint8_t* k, int i, int eb, int ee, bool single_digits) {
    return gpt2_piece_starts_t(b, ClsArray{k}, i, eb, ee, single_digits);
}

// Branch-free form of the same predicate for all-ASCII subjects, driven by the class words of the five neighbours
// (the window kernel obtains them with warp shuffles).  Class word bits: G_L / G_N / G_S as above, G_SP = U+0020,
// G_AP = apostrophe, G_ES = one of s t m d, G_ERV = r or v, G_EE = e, G_EL = l, G_X = "this position exists",
// G_BOS = the virtual position just before the element start.  A missing neighbour has word 0.
enum : uint32_t { G_L = 1, G_N = 2, G_S = 4, G_SP = 8, G_AP = 16, G_ES = 32, G_ERV = 64, G_EE = 128, G_EL = 256, G_BOS = 512, G_X = 1024 };
B2_HD uint32_t gpt2_class_word(uint8_t byte, uint8_t cls) {
    uint32_t w = G_X | (cls & (C_L | C_N | C_S));
    if (byte == 0x20) w |= G_SP;
    if (byte == '\'') w |= G_AP;
    if (byte == 's' || byte == 't' || byte == 'm' || byte == 'd') w |= G_ES;
    if (byte == 'r' || byte == 'v') w |= G_ERV;
    if (byte == 'e') w |= G_EE;
    if (byte == 'l') w |= G_EL;
    return w;
}
B2_HD bool gpt2_nb_okprev(uint32_t x) { return (x & (G_L | G_N | G_BOS)) || ((x & G_S) && !(x & G_SP)); }
B2_HD bool gpt2_nb_len3(uint32_t a, uint32_t b) { return !(a & G_ES) && (((a & G_ERV) && (b & G_EE)) || ((a & G_EL) && (b & G_EL))); }
// `apos_near` = an apostrophe exists among p1..p3 (lets the caller skip fetching p2..p4 when false).
B2_HD bool gpt2_start_nb(uint32_t c, uint32_t p1, uint32_t p2, uint32_t p3, uint32_t p4, uint32_t n1, bool digits, bool apos_near) {
    const bool prev_sp = (p1 & G_SP) != 0;
    const bool s_start = !(p1 & G_S) || ((n1 & G_X) && !(n1 & G_S));
    const bool n_start = digits || (!prev_sp && !(p1 & G_N));
    const bool o_start = !prev_sp && (p1 & (G_L | G_N | G_S));
    bool inside = false, ends = false;
    if (apos_near) {
        const bool d1 = (p1 & G_AP) && gpt2_nb_okprev(p2);
        const bool d2 = (p2 & G_AP) && gpt2_nb_okprev(p3);
        const bool d3 = (p3 & G_AP) && gpt2_nb_okprev(p4);
        inside = (d1 && ((c & G_ES) || gpt2_nb_len3(c, n1))) || (d2 && gpt2_nb_len3(p1, c));
        ends = (d2 && (p1 & G_ES)) || (d3 && gpt2_nb_len3(p2, p1));
    }
    const bool l_start = !prev_sp && !inside && (ends || !(p1 & G_L));
    const bool r = (c & G_S) ? s_start : (c & G_N) ? n_start : (c & G_L) ? l_start : o_start;
    return r || (p1 & G_BOS);
}

// Length of the case-insensitive contraction (?i:'s|'t|'re|'ve|'m|'ll|'d) at the apostrophe s[q] (0 = none); U+017F folds to 's'.
B2_HD int llama3_contraction_len(const uint8_t* b, int q, int ee) {
    if (q + 1 >= ee) return 0;
    const uint8_t b1 = b[q + 1];
    if (ci_eq(b1, 's') || ci_eq(b1, 't') || ci_eq(b1, 'm') || ci_eq(b1, 'd')) return 2;
    if (q + 2 < ee) {
        const uint8_t b2 = b[q + 2];
        if (b1 == 0xC5 && b2 == 0xBF) return 3;
        if ((ci_eq(b1, 'r') || ci_eq(b1, 'v')) && ci_eq(b2, 'e')) return 3;
        if (ci_eq(b1, 'l') && ci_eq(b2, 'l')) return 3;
    }
    return 0;
}
// Word form of the same predicate: 32 byte positions at a time as bit masks (bit i = position base + i), so one thread
// evaluates 32 positions with ~40 logic operations.  Per-word class masks: L / N / S (continuation bytes carry their
// owner's bits), SP = byte 0x20, A2 / A3 = an apostrophe followed by a 2- / 3-byte contraction ('s 't 'm 'd / 're 've 'll),
// CONT = UTF-8 continuation byte, MB = a multi-byte whitespace character whose next character exists and is not
// whitespace, X = the position exists (inside the element and loaded).  Masks of positions that do not exist are 0.
struct G2Word { uint32_t L, N, S, SP, A2, A3, CONT, MB, X; };
B2_HD uint32_t g2_fsl(uint32_t lo, uint32_t hi, int k) { return k == 0 ? hi : (hi << k) | (lo >> (32 - k)); }   // bits of (hi:lo) moved up by k
B2_HD uint32_t g2_fsr(uint32_t lo, uint32_t hi, int k) { return k == 0 ? lo : (lo >> k) | (hi << (32 - k)); }   // moved down by k
B2_HD uint32_t g2_ok1(const G2Word& w) { return w.L | w.N | (w.S & ~w.SP); }
// contraction starts of this word: apostrophes reached as the start of a match (gpt2_apostrophe_t).  p_ok1 = g2_ok1 of
// the previous word, bos = bit of the element's first position if it lies in this word.
B2_HD void g2_contractions(const G2Word& w, uint32_t p_ok1, uint32_t bos, uint32_t& c2, uint32_t& c3) {
    const uint32_t okp = g2_fsl(p_ok1, g2_ok1(w), 1) | bos;
    c2 = w.A2 & okp;
    c3 = w.A3 & okp;
}
// piece starts of this word.  p = previous word's masks, (pc2, pc3) / (c2, c3) = g2_contractions of the previous / this
// word, n_ns = X & ~S of the next word.
B2_HD uint32_t g2_starts(const G2Word& w, const G2Word& p, uint32_t c2, uint32_t c3, uint32_t pc2, uint32_t pc3, uint32_t n_ns,
                         uint32_t bos, bool single_digits) {
    const uint32_t p1L = g2_fsl(p.L, w.L, 1), p1N = g2_fsl(p.N, w.N, 1), p1S = g2_fsl(p.S, w.S, 1), p1SP = g2_fsl(p.SP, w.SP, 1);
    const uint32_t n1ns = g2_fsr(w.X & ~w.S, n_ns, 1);
    const uint32_t s_start = w.S & (~p1S | n1ns | w.MB);
    const uint32_t n_start = single_digits ? w.N : (w.N & ~p1SP & ~p1N);
    const uint32_t o_start = w.X & ~(w.L | w.N | w.S) & ~p1SP & (p1L | p1N | p1S);
    const uint32_t c23 = c2 | c3, pc23 = pc2 | pc3;
    const uint32_t inside = g2_fsl(pc23, c23, 1) | g2_fsl(pc3, c3, 2);
    const uint32_t ends = g2_fsl(pc2, c2, 2) | g2_fsl(pc3, c3, 3);
    const uint32_t l_start = w.L & ~p1SP & ~inside & (ends | ~p1L);
    return (s_start | n_start | o_start | l_start | bos) & w.X & ~w.CONT;
}

// ------------------------------------------------------------------------------------------
// Word form of the Llama-3 / cl100k pattern in "isolate" mode
//   (?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+
// The pattern tiles the subject; which characters start a piece (derived from PCRE2's leftmost / first-alternative /
// backtracking semantics, checked against PCRE2 in tests/test_core_host.py):
//   other (not space / letter / number) char: starts a piece iff it is the first of its run and the byte before is not U+0020
//          ("a match starts here", MSO); a following newline run belongs to the same piece ([\r\n]*)
//   letter: starts iff the previous char is a newline, a number, the element start, or an "other" char at which NO match starts;
//          a letter run is glued to one preceding non-newline space or MSO char ([^\r\n\p{L}\p{N}]?\p{L}+), except that a
//          contraction ('s 't 'm 'd 're 've 'll, any case) at an MSO apostrophe takes its letters first and the next letter starts
//   number (ASCII digits): every third digit from the start of its run
//   whitespace run [a, b): a' = a, or a + its leading newlines if an "other" char precedes (they went to that piece); a' starts;
//          the char after the last newline starts (\s*[\r\n]+ ends there); the last char starts if it is not a newline and a
//          non-space follows (\s+(?!\S) gives it back; it then glues to a following letter, or, if U+0020, to a following other run)
// Masks per 32 byte positions as for the GPT-2 form, plus NL = \r or \n and PG = "the previous char is a MULTI-byte other char
// at which a match starts" (for one-byte chars this is a shift of MSO).
// ------------------------------------------------------------------------------------------
struct L3Word { uint32_t L, N, S, NL, SP, A2, A3, CONT, MB, PG, X; };
B2_HD uint32_t l3_mod3(int r) { return r == 0 ? 0x49249249u : r == 1 ? 0x92492492u : 0x24924924u; }     // bit positions = r (mod 3)
// fill every run of `run` upwards from its seed bit (at most one seed per run), inside one word
B2_HD uint32_t l3_fill_up(uint32_t run, uint32_t seeds) { return ((run + seeds) ^ run) & run; }
B2_HD uint32_t l3_brev(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
#endif
}
B2_HD uint32_t l3_fill_down(uint32_t run, uint32_t seeds) { return l3_brev(l3_fill_up(l3_brev(run), l3_brev(seeds))); }
B2_HD uint32_t l3_other(const L3Word& w) { return w.X & ~(w.L | w.N | w.S); }
// "a match starts at this other char": first of its run, byte before is not U+0020.  p = previous word.
B2_HD uint32_t l3_mso(const L3Word& w, const L3Word& p) {
    return l3_other(w) & ~w.CONT & ~g2_fsl(l3_other(p), l3_other(w), 1) & ~g2_fsl(p.SP, w.SP, 1);
}
// seeds of the two cross-word fills: newline runs right after an other char (upwards), tails of whitespace runs (downwards)
B2_HD uint32_t l3_lead_seeds(const L3Word& w, const L3Word& p) { return w.NL & g2_fsl(l3_other(p), l3_other(w), 1); }
B2_HD uint32_t l3_tail_seeds(const L3Word& w, uint32_t n_s /* S of the next word */) { return (w.S & ~w.NL) & ~g2_fsr(w.S, n_s, 1); }
// number starts of one word.  phase_in = digits of the run that continues into bit 0, before it (only used if it does).
B2_HD uint32_t l3_number_starts(uint32_t N, uint32_t pN_bit31, int phase_in, bool force0) {
    const uint32_t cont0 = (N & 1u) && pN_bit31 && !force0 ? 1u : 0u;     // bit 0 continues a run from the previous word
    const uint32_t rs = N & ~((N << 1) | (pN_bit31 && !force0 ? 1u : 0u));
    const int r0 = (3 - phase_in % 3) % 3;
    uint32_t st = 0;
    for (int r = 0; r < 3; ++r) {
        uint32_t seeds = rs & l3_mod3(r);
        if (cont0 && r == r0) seeds |= 1u;
        st |= l3_fill_up(N, seeds) & l3_mod3(r);
    }
    return st;
}
// piece starts of one word.  lead / tail = the filled masks of this word, p_lead = of the previous word; (d2, d3) / (pd2, pd3)
// = contraction starts (A2 / A3 & MSO) of this / the previous word; nst = number starts; n_ns = X & ~S of the next word.
B2_HD uint32_t l3_starts(const L3Word& w, const L3Word& p, uint32_t mso, uint32_t p_mso, uint32_t d2, uint32_t d3, uint32_t pd2, uint32_t pd3,
                         uint32_t lead, uint32_t p_lead, uint32_t tail, uint32_t nst, uint32_t n_ns, uint32_t bos) {
    const uint32_t O = l3_other(w), pO = l3_other(p);
    const uint32_t p1L = g2_fsl(p.L, w.L, 1), p1S = g2_fsl(p.S, w.S, 1), p1NL = g2_fsl(p.NL, w.NL, 1), p1O = g2_fsl(pO, O, 1);
    const uint32_t p1T = p1S & ~p1NL;
    const uint32_t p1MSO = g2_fsl(p_mso, mso, 1) | w.PG;                 // (a one-byte MSO char right before, or a multi-byte one)
    const uint32_t ends = g2_fsl(pd2, d2, 2) | g2_fsl(pd3, d3, 3);
    const uint32_t l_start = w.L & ((~p1L & ~p1T & ~(p1O & p1MSO)) | ends);
    const uint32_t T = w.S & ~w.NL;
    const uint32_t n1ns = g2_fsr(w.X & ~w.S, n_ns, 1);
    const uint32_t s_start = (w.S & ~p1S & ~(p1O & w.NL)) | (T & g2_fsl(p_lead, lead, 1)) | (tail & p1NL) | (T & n1ns) | w.MB;
    return (l_start | nst | mso | s_start | bos) & w.X & ~w.CONT;
}

// (p)+ for the "contiguous" rewrite (src/regex_split.cpp:33-37): greedy repetition of the pattern.
template <class C>
B2_HD Match match_rep(const C& c, const SplitSpec& spec, bool repeat, int p, int end) {
    Match m = match_at(c, spec, p, end);
    if (!repeat || m.len == 0) return m;
    int q = p + m.len;
    while (q < end && q < c.known()) {
        const Match n = match_at(c, spec, q, end);
        if (n.peek > m.peek) m.peek = n.peek;
        if (n.len == 0) break;
        q += n.len;
    }
    if (q >= c.known() && q < end && q + 1 > m.peek) m.peek = q + 1;
    m.len = q - p;
    return m;
}

// The behaviour logic of RegexSplit's add_split lambda (src/regex_split.cpp:243-284), kept in the
// reference's element-relative coordinates including its quirks: `last_begin` is never reset,
// an unset last_begin (-1) clamps to 0, and `max_splits` only stretches the end of one piece.
enum : int { SPLIT_REMOVED = 0, SPLIT_ISOLATED = 1, SPLIT_MERGED_PREV = 2, SPLIT_MERGED_NEXT = 3 };
struct SplitEmitter {
    int mode;
    bool invert;
    int max_splits;
    int len;            // element length in bytes
    int64_t last_begin; // -1 = unset (size_t(-1) in the reference)
    uint32_t n_splits;
    B2_HD void reset(int mode_, bool invert_, int max_splits_, int len_) {
        mode = mode_; invert = invert_; max_splits = max_splits_; len = len_; last_begin = -1; n_splits = 0;
    }
    // Returns true if a piece [ob, oe) (relative to the element) is produced.
    B2_HD bool add(int begin, int end, bool inv, int& ob, int& oe) {
        switch (mode) {
        case SPLIT_REMOVED: if (inv) return false; break;
        case SPLIT_ISOLATED: break;
        case SPLIT_MERGED_PREV:
            if (!inv && end != len) { last_begin = begin; return false; }
            else if (inv) begin = (int)last_begin;
            break;
        case SPLIT_MERGED_NEXT:
            if (!inv) { if (last_begin != -1) begin = (int)last_begin; }
            else { last_begin = begin; return false; }
            break;
        }
        if (begin < 0) begin = 0;
        if (end > len) end = len;
        if (n_splits == (uint32_t)max_splits) end = len;
        ob = begin; oe = end;
        ++n_splits;
        return true;
    }
    // after the match loop (src/regex_split.cpp:302-309); `start` = end of the last match
    B2_HD bool finish(int start, int& ob, int& oe) {
        if (start < len) return add(start, len, invert, ob, oe);
        if (mode == SPLIT_MERGED_NEXT && last_begin != (int64_t)len) return add((int)last_begin, len, invert, ob, oe);
        return false;
    }
};

// Sequential split of one element [lo, end) (reference loop src/regex_split.cpp:287-309).
// sink(b, e) receives absolute piece offsets.  Used for elements that outgrow the window path
// and by the host harness; the window kernels run the same chain in parallel.
template <class C, class Sink>
B2_HD void split_element_scan(const C& c, const SplitSpec& spec, bool repeat, int lo, int end,
                              SplitEmitter& em, Sink&& sink) {
    em.len = end - lo;
    int start = lo, p = lo, ob, oe;
    while (p < end) {
        const Match m = match_rep(c, spec, repeat, p, end);
        if (m.len > 0) {
            if (p != start && em.add(start - lo, p - lo, em.invert, ob, oe)) sink(lo + ob, lo + oe);
            if (em.add(p - lo, p + m.len - lo, !em.invert, ob, oe)) sink(lo + ob, lo + oe);
            p += m.len;
            start = p;
        } else {
            p = c.next(p);
        }
    }
    if (em.finish(start - lo, ob, oe)) sink(lo + ob, lo + oe);
}

// ------------------------------------------------------------------------------------------
// Merge-rank table: open addressing over (left,right) -> (rank,new_id), 16-byte slots so one
// probe is one vector load.  (The reference's MergesMap, src/bpe_tokenizer.hpp:40-115, plays the
// same role on the CPU; hash function and load factor here are our own.)
// ------------------------------------------------------------------------------------------
struct MergeSlot { uint32_t left, right; int32_t rank, new_id; };
constexpr uint32_t kEmptyKey = 0xFFFFFFFFu;
constexpr int32_t kNoRank = 0x7FFFFFFF;

B2_HD uint32_t merge_hash(uint32_t l, uint32_t r) {
    uint32_t h = l * 0x9E3779B1u ^ (r + 0x7F4A7C15u) * 0x85EBCA77u;
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    h ^= h >> 13;
    return h;
}

struct MergeTable {
    const MergeSlot* slots;
    uint32_t mask;
    const int32_t* rank_newid;   // [n_merges] token produced by the merge of that rank
    // != 0: some token is the product of more than one merge, so two queue entries can tie on (rank, seq) — when a merge finds its
    // own product on BOTH sides.  The mask-based and packed merge loops report that (they break such a tie left pair first); the row
    // is then redone by bpe_merge_heap, which pops ties exactly like the reference's std::priority_queue.
    int32_t tie_check;
};

#if defined(__CUDA_ARCH__)
B2_HD MergeSlot load_slot(const MergeSlot* p) {
    const int4 v = __ldg(reinterpret_cast<const int4*>(p));
    return MergeSlot{(uint32_t)v.x, (uint32_t)v.y, v.z, v.w};
}
#else
B2_HD MergeSlot load_slot(const MergeSlot* p) { return *p; }
#endif

B2_HD bool merge_find(const MergeTable& t, int32_t l, int32_t r, int32_t& rank, int32_t& new_id) {
    uint32_t h = merge_hash((uint32_t)l, (uint32_t)r) & t.mask;
    for (;;) {
        const MergeSlot s = load_slot(t.slots + h);
        if (s.left == (uint32_t)l && s.right == (uint32_t)r) { rank = s.rank; new_id = s.new_id; return true; }
        if (s.left == kEmptyKey) { rank = kNoRank; new_id = -1; return false; }
        h = (h + 1) & t.mask;
    }
}

// ------------------------------------------------------------------------------------------
// Flattened byte trie (longest match).  Node n owns edges [first[n], first[n+1]) sorted by byte.
// ------------------------------------------------------------------------------------------
struct FlatTrie {
    const int32_t* first;       // [n_nodes + 1]
    const int32_t* value;       // [n_nodes]   (-1 = no token ends here)
    const uint8_t* edge_byte;   // [n_edges]
    const int32_t* edge_child;  // [n_edges]
    const int32_t* root_child;  // [256] child of the root per byte, -1 if none
};

B2_HD int32_t trie_child(const FlatTrie& t, int32_t node, uint8_t ch) {
    int32_t lo = t.first[node], hi = t.first[node + 1];
    while (lo < hi) {
        const int32_t mid = (lo + hi) >> 1;
        const uint8_t b = t.edge_byte[mid];
        if (b == ch) return t.edge_child[mid];
        if (b < ch) lo = mid + 1; else hi = mid;
    }
    return -1;
}

// Longest match starting at s[idx] (idx < end).  On success returns the token id and advances
// idx to the end of the match; otherwise returns -1 and leaves idx unchanged (utils.cpp:517-538).
B2_HD int32_t trie_longest(const FlatTrie& t, const uint8_t* s, int& idx, int end) {
    int32_t node = t.root_child[s[idx]];
    int32_t found = -1;
    int best = idx, i = idx;
    while (node >= 0) {
        ++i;
        const int32_t v = t.value[node];
        if (v != -1) { found = v; best = i; }
        if (i >= end) break;
        node = trie_child(t, node, s[i]);
    }
    idx = best;
    return found;
}

// ------------------------------------------------------------------------------------------
// BPE.  Per-byte symbolisation tables are precomputed on the host from the trie over
// "vocab minus merge products" (src/bpe_tokenizer.cpp:375-386):
//   byte_sym[c]  >= 0  : c starts no longer token; the single-byte token id
//                == -2 : must walk the trie (a longer token starts with c)
//                == -1 : no token starts with c (use byte_miss)
//   byte_miss[c] >= 0  : <0xNN> byte-fallback id or the unk id (src/bpe_tokenizer.cpp:242-254)
//                == -1 : the byte is dropped
// ------------------------------------------------------------------------------------------
struct BpeTables {
    const int32_t* byte_sym;   // [256]
    const int32_t* byte_miss;  // [256]
    const uint32_t* pair_rank; // [65536] rank of the merge of the one-byte symbols of (b0,b1), index b0<<8|b1; kNoKey if none
    FlatTrie trie;
    MergeTable merges;
    const uint32_t* pair_bits; // [512] bit (b0 << 7 | b1) set iff pair_rank[b0 << 8 | b1] != kNoKey, for b0, b1 < 128;
                               // then [2048]: bit (b0 << 8 | b1) set iff the trie has a token prefix b0 b1; then the u16 ASCII rank table
    int32_t newid_base;        // >= 0: the token produced by the merge of rank r is newid_base + r (vocab in merge order); -1: use rank_newid
};
constexpr int32_t kSymWalk = -2;

// Symbolise s[b..e) (+ optional suffix bytes) into ids[]; returns the number of symbols.
B2_HD int bpe_symbolize(const BpeTables& T, const uint8_t* s, int b, int e, int32_t* ids) {
    int n = 0;
    for (int i = b; i < e;) {
        const uint8_t c = s[i];
        int32_t id = T.byte_sym[c];
        if (id == kSymWalk) {
            int j = i;
            id = trie_longest(T.trie, s, j, e);
            if (id >= 0) { ids[n++] = id; i = j; continue; }
            id = -1;
        }
        if (id < 0) id = T.byte_miss[c];
        if (id >= 0) ids[n++] = id;
        ++i;
    }
    return n;
}

// In-place merge loop for a piece of n symbols held in ids[0..n).  rank/newid/birth are scratch
// arrays of at least n-1 entries.  Order: smallest (rank, birth) first, where birth is the
// reference's push-sequence number: pair k of the initial sequence gets k, and the (up to two)
// pairs created by the m-th merge get n-1+m  (src/bpe_tokenizer.cpp:269-285,314-322).
// Returns the number of tokens left in ids[].
template <class BirthT>
B2_HD int bpe_merge_serial(const MergeTable& M, int32_t* ids, int32_t* rank, int32_t* newid, BirthT* birth, int n) {
    if (n < 2) return n;
    bool any = false;
    for (int k = 0; k + 1 < n; ++k) {
        int32_t r, v;
        any |= merge_find(M, ids[k], ids[k + 1], r, v);
        rank[k] = r; newid[k] = v; birth[k] = (BirthT)k;
    }
    if (!any) return n;
    int seq = n - 1;
    while (n >= 2) {
        int32_t best = kNoRank;
        int bk = -1;
        uint32_t bb = 0xFFFFFFFFu;
        for (int k = 0; k + 1 < n; ++k) {
            const int32_t r = rank[k];
            if (r < best || (r == best && r != kNoRank && (uint32_t)birth[k] < bb)) { best = r; bk = k; bb = (uint32_t)birth[k]; }
        }
        if (bk < 0) break;
        ids[bk] = newid[bk];
        for (int k = bk + 1; k + 1 < n; ++k) ids[k] = ids[k + 1];
        for (int k = bk + 1; k + 2 < n; ++k) { rank[k] = rank[k + 1]; newid[k] = newid[k + 1]; birth[k] = birth[k + 1]; }
        --n;
        ++seq;
        if (bk > 0) {
            int32_t r, v;
            merge_find(M, ids[bk - 1], ids[bk], r, v);
            rank[bk - 1] = r; newid[bk - 1] = v; birth[bk - 1] = (BirthT)seq;
        }
        if (bk + 1 < n) {
            int32_t r, v;
            merge_find(M, ids[bk], ids[bk + 1], r, v);
            rank[bk] = r; newid[bk] = v; birth[bk] = (BirthT)seq;
        }
    }
    return n;
}

// Same loop with (rank, birth) packed into one 32-bit key = rank << 12 | birth, so the argmin is a
// single unsigned compare and the state fits in shared memory.  Valid for n <= 2048 symbols
// (birth <= 2n-2 < 4096) and ranks < 2^20; the table builder enforces the latter.
constexpr uint32_t kNoKey = 0xFFFFFFFFu;
constexpr int kPackedBirthBits = 12;
constexpr int kPackedMaxSymbols = 2048;
template <class IdT>
B2_HD int bpe_merge_packed(const MergeTable& M, IdT* ids, uint32_t* key, int n, bool* tie = nullptr) {
    if (n < 2) return n;
    bool any = false;
    for (int k = 0; k + 1 < n; ++k) {
        int32_t r, v;
        const bool f = merge_find(M, (int32_t)ids[k], (int32_t)ids[k + 1], r, v);
        any |= f;
        key[k] = f ? (((uint32_t)r << kPackedBirthBits) | (uint32_t)k) : kNoKey;
    }
    if (!any) return n;
    int seq = n - 1;
    while (n >= 2) {
        uint32_t best = kNoKey;
        int bk = -1;
        for (int k = 0; k + 1 < n; ++k) {
            const uint32_t q = key[k];
            if (q < best) { best = q; bk = k; }
        }
        if (bk < 0) break;
        ids[bk] = (IdT)M.rank_newid[best >> kPackedBirthBits];
        for (int k = bk + 1; k + 1 < n; ++k) ids[k] = ids[k + 1];
        for (int k = bk + 1; k + 2 < n; ++k) key[k] = key[k + 1];
        --n;
        ++seq;
        bool fl = false, fr = false;
        if (bk > 0) {
            int32_t r, v;
            fl = merge_find(M, (int32_t)ids[bk - 1], (int32_t)ids[bk], r, v);
            key[bk - 1] = fl ? (((uint32_t)r << kPackedBirthBits) | (uint32_t)seq) : kNoKey;
        }
        if (bk + 1 < n) {
            int32_t r, v;
            fr = merge_find(M, (int32_t)ids[bk], (int32_t)ids[bk + 1], r, v);
            key[bk] = fr ? (((uint32_t)r << kPackedBirthBits) | (uint32_t)seq) : kNoKey;
        }
        if (tie && fl && fr && ids[bk - 1] == ids[bk] && ids[bk + 1] == ids[bk]) *tie = true;      // both new pairs are (x, x): equal (rank, seq)
    }
    return n;
}

// Heap form of the same loop: O(n log n), state in caller-provided scratch.
//   sym_id/prev/next : [2n]   (dead symbols have id == -1 after being merged)
//   heap             : [3n] entries (at most n-1 initial pushes + 2 per merge)
// This is the EXACT form of the reference's loop (src/bpe_tokenizer.cpp:262-327), heap included: entries with equal (rank, seq) —
// reachable only when two merges produce the same token, SURVEY App. B item 1 — are popped in the order libstdc++'s
// std::priority_queue would pop them, because heap_push / heap_pop below are std::push_heap / std::pop_heap operation for operation
// (bits/stl_heap.h __push_heap / __adjust_heap) under the reference's CompareRank (:166-172).  Vocabularies with such merges run every
// row that meets the tie through this function (MergeTable::tie_check); checked against std::priority_queue itself in the CPU tier.
// Returns the token count; tokens are written to out[0..count).
struct HeapEntry { int32_t rank, birth, a, b; };
B2_HD bool heap_comp(const HeapEntry& x, const HeapEntry& y) {      // CompareRank: "x sits below y"
    return x.rank != y.rank ? x.rank > y.rank : x.birth > y.birth;
}
B2_HD void heap_sift_up(HeapEntry* h, int hole, int top, const HeapEntry& v) {      // std::__push_heap
    int parent = (hole - 1) / 2;
    while (hole > top && heap_comp(h[parent], v)) {
        h[hole] = h[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    h[hole] = v;
}
B2_HD void heap_push(HeapEntry* h, int& n, HeapEntry e) {      // c.push_back(e); std::push_heap(c.begin(), c.end(), comp)
    heap_sift_up(h, n, 0, e);
    ++n;
}
B2_HD HeapEntry heap_pop(HeapEntry* h, int& n) {               // top(); std::pop_heap(c.begin(), c.end(), comp); c.pop_back()
    const HeapEntry top = h[0];
    if (n > 1) {
        const HeapEntry v = h[n - 1];
        const int len = n - 1;                                  // std::__adjust_heap(first, 0, len, v)
        int hole = 0, child = 0;
        while (child < (len - 1) / 2) {
            child = 2 * (child + 1);
            if (heap_comp(h[child], h[child - 1])) --child;
            h[hole] = h[child];
            hole = child;
        }
        if ((len & 1) == 0 && child == (len - 2) / 2) {
            child = 2 * (child + 1);
            h[hole] = h[child - 1];
            hole = child - 1;
        }
        heap_sift_up(h, hole, 0, v);
    }
    --n;
    return top;
}
B2_HD int bpe_merge_heap(const MergeTable& M, int n, int32_t* sym_id, int32_t* sym_prev, int32_t* sym_next,
                         HeapEntry* heap, int32_t* out) {
    if (n == 0) return 0;
    int hn = 0, total = n, seq = 0;
    for (int k = 0; k < n; ++k) { sym_prev[k] = k - 1; sym_next[k] = (k + 1 < n) ? k + 1 : -1; }
    for (int k = 0; k + 1 < n; ++k) {
        int32_t r, v;
        if (merge_find(M, sym_id[k], sym_id[k + 1], r, v)) heap_push(heap, hn, HeapEntry{r, seq, k, k + 1});
        ++seq;
    }
    int head = 0, live = n;
    while (hn > 0 && live >= 2) {
        const HeapEntry e = heap_pop(heap, hn);
        if (sym_id[e.a] < 0 || sym_id[e.b] < 0 || sym_next[e.a] != e.b) continue;
        int32_t r, v;
        merge_find(M, sym_id[e.a], sym_id[e.b], r, v);
        const int32_t pv = sym_prev[e.a], nx = sym_next[e.b], m = total++;
        sym_id[m] = v; sym_prev[m] = pv; sym_next[m] = nx;
        sym_id[e.a] = -1; sym_id[e.b] = -1;
        if (pv != -1) sym_next[pv] = m; else head = m;
        if (nx != -1) sym_prev[nx] = m;
        --live;
        ++seq;
        int32_t r1, v1;
        if (pv != -1 && merge_find(M, sym_id[pv], v, r1, v1)) heap_push(heap, hn, HeapEntry{r1, seq, pv, m});
        int32_t r2, v2;
        if (nx != -1 && merge_find(M, v, sym_id[nx], r2, v2)) heap_push(heap, hn, HeapEntry{r2, seq, m, nx});
    }
    int cnt = 0;
    for (int k = head; k != -1; k = sym_next[k]) out[cnt++] = sym_id[k];
    return cnt;
}

// ------------------------------------------------------------------------------------------
// WordPiece: one word s[b..e) -> ids; returns count (>= 1).  src/wordpiece_tokenizer.cpp:96-130.
// A zero-length word yields [unk] (the reference reads out of bounds there).
// ------------------------------------------------------------------------------------------
// Rank-indexed byte trie for the WordPiece walks: a node carries a 256-bit child bitmap with per-word prefix counts and the
// index of its first child (children are numbered consecutively, in byte order), so one step is one round of independent
// loads — bitmap word, prefix count, base, value — instead of a binary search over an edge list.
struct RankNode { uint32_t bits[8]; int32_t base; int32_t value; uint8_t cum[8]; };      // 48 bytes
// Two-byte jump table of a RankTrie (ASCII pairs only, 128 x 128 entries): what the first TWO steps of a walk starting with the
// bytes (b0, b1) find — node2 = the node after both bytes (-1 if none), and the longest valued node on the way (0, 1 or 2 bytes
// long) — so a walk starts with ONE table read instead of the root lookup and two dependent node loads, and a word of up to two
// bytes needs no node at all.  val1[b] = value of the root's child for byte b (one-byte words).
struct RankJump { int32_t node2; uint32_t info; };      // info: bit 31 root child exists | bits 24-25 length of the best value | bits 0-23 value + 1
constexpr uint32_t kJumpHas1 = 0x80000000u;
struct RankTrie {
    const RankNode* nodes;
    const int32_t* root_child;   // [256] child of the root per byte, -1 if none
    const RankJump* jump2;       // [128 * 128] or nullptr
    const int32_t* val1;         // [256] or nullptr
};
B2_HD int32_t rank_popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
B2_HD int32_t rank_child(const RankNode& nd, uint32_t ch) {
    const uint32_t w = ch >> 5, bit = ch & 31u;
    const uint32_t bw = nd.bits[w];
    return ((bw >> bit) & 1u) ? nd.base + (int32_t)nd.cum[w] + rank_popc(bw & ((1u << bit) - 1u)) : -1;
}
// Longest match starting at s[idx] (same contract as trie_longest).
B2_HD int32_t rank_trie_longest(const RankTrie& t, const uint8_t* s, int& idx, int end) {
    int32_t node, found = -1;
    int best = idx, i = idx;
    bool jumped = false;
    if (t.jump2 && end - idx >= 2 && ((s[idx] | s[idx + 1]) & 0x80) == 0) {
        const RankJump j = t.jump2[((uint32_t)s[idx] << 7) | s[idx + 1]];
        if (!(j.info & kJumpHas1)) return -1;
        found = (int32_t)(j.info & 0xFFFFFFu) - 1;
        best = idx + (int)((j.info >> 24) & 3u);
        i = idx + 2;
        node = i < end ? j.node2 : -1;            // (the value of node2 is already in `found`)
        jumped = true;
    } else if (t.val1 && end - idx == 1) {
        found = t.val1[s[idx]];
        if (found >= 0) ++idx;
        return found;
    } else node = t.root_child[s[idx]];
    while (node >= 0) {
        const RankNode& nd = t.nodes[node];
        if (!jumped) {
            ++i;
            const int32_t v = nd.value;
            if (v != -1) { found = v; best = i; }
        }
        jumped = false;
        if (i >= end) break;
        node = rank_child(nd, s[i]);
    }
    idx = best;
    return found;
}

struct WordpieceTables {
    RankTrie root;
    RankTrie sub;
    int32_t max_bytes;
};

B2_HD int wordpiece_word(const WordpieceTables& T, const uint8_t* s, int b, int e, int32_t unk, int32_t* out) {
    if (e - b > T.max_bytes || e <= b) { out[0] = unk; return 1; }
    int idx = b;
    int32_t id = rank_trie_longest(T.root, s, idx, e);
    if (id < 0) { out[0] = unk; return 1; }
    int n = 0;
    out[n++] = id;
    while (idx < e) {
        id = rank_trie_longest(T.sub, s, idx, e);
        if (id < 0) { out[0] = unk; return 1; }
        out[n++] = id;
    }
    return n;
}

// ------------------------------------------------------------------------------------------
// SpecialTokensSplit (src/special_tokens_split.cpp:61-162): the match that starts at a given byte, for a pattern that is an
// alternation of groups  (?:\s*)?(tok|tok|...)(?:\s*)?  of literal tokens (PCRE2: leftmost, first alternative, greedy \s* with
// backtracking).  Used by kernels_special.cuh per position and by the host harness for the CPU-tier check against PCRE2.
// ------------------------------------------------------------------------------------------
constexpr int kSpecialGroups = 8;
struct SpecialTables {
    FlatTrie trie[kSpecialGroups];     // value = position of the token inside its group (smaller = earlier alternative)
    uint8_t strip_left[kSpecialGroups], strip_right[kSpecialGroups];
    int32_t n_groups;
    uint32_t first[8];                 // bytes at which a match can start
    int32_t ws_token;                  // a token of a strip_left group starts with whitespace: full backtracking needed
};

// Earliest alternative among the group's tokens matching at chars[q..ee); they all lie on one trie path.
B2_HD bool special_token_at(const FlatTrie& t, const uint8_t* chars, int q, int ee, int& tok_end) {
    int32_t node = t.root_child[chars[q]];
    int32_t best = 0x7FFFFFFF;
    int i = q;
    while (node >= 0) {
        ++i;
        const int32_t v = t.value[node];
        if (v != -1 && v < best) { best = v; tok_end = i; }
        if (i >= ee) break;
        node = trie_child(t, node, chars[i]);
    }
    return best != 0x7FFFFFFF;
}

// Length in bytes of the whitespace character starting at chars[i] (0 if it is not one).
B2_HD int special_ws_len(const uint8_t* chars, int i, int ee, const ClassTables& T) {
    const uint8_t b = chars[i];
    if (b < 0x80) return (T.ascii[b] & C_S) ? 1 : 0;
    if (b < 0xC2 || !(char_class(chars, i, ee, T) & C_S)) return 0;
    return b >= 0xF0 ? 4 : b >= 0xE0 ? 3 : 2;
}

// The match starting exactly at pos, or m1 = 0.  [g0, g1) = the token (capture group), [pos, m1) = the full match.
B2_HD void special_match_at(const SpecialTables& ST, const ClassTables& T, const uint8_t* chars, int pos, int ee,
                                                 int& m1, int& g0, int& g1) {
    m1 = 0;
    int ws_end = -1;     // end of the whitespace run starting at pos (computed on first use)
    for (int g = 0; g < ST.n_groups; ++g) {
        int q = pos, te = 0;
        bool hit = false;
        if (ST.strip_left[g]) {
            if (ws_end < 0) { ws_end = pos; int l; while (ws_end < ee && (l = special_ws_len(chars, ws_end, ee, T)) > 0) ws_end += l; }
            // greedy \s*, then give back one character at a time
            q = ws_end;
            for (;;) {
                if (q < ee && ((ST.first[chars[q] >> 5] >> (chars[q] & 31)) & 1u) && special_token_at(ST.trie[g], chars, q, ee, te)) { hit = true; break; }
                if (q <= pos || !ST.ws_token) break;     // no token starts with whitespace: only the end of the run can match
                --q;
                while (q > pos && is_cont_byte(chars[q])) --q;
            }
        } else {
            hit = special_token_at(ST.trie[g], chars, pos, ee, te);
        }
        if (!hit) continue;
        g0 = q; g1 = te; m1 = te;
        if (ST.strip_right[g]) { int l; while (m1 < ee && (l = special_ws_len(chars, m1, ee, T)) > 0) m1 += l; }
        return;
    }
}


// Sequential scan of one element with the matcher above (the kernel resolves the same scan per 32-position chunk).
template <class Sink>
B2_HD void special_split_element(const SpecialTables& ST, const ClassTables& T, const uint8_t* chars, int eb, int ee, Sink&& sink) {
    int cur = eb;
    for (int pos = eb; pos < ee; ++pos) {
        if (pos < cur || is_cont_byte(chars[pos]) || !((ST.first[chars[pos] >> 5] >> (chars[pos] & 31)) & 1u)) continue;
        int m1 = 0, g0 = 0, g1 = 0;
        special_match_at(ST, T, chars, pos, ee, m1, g0, g1);
        if (m1 <= pos) continue;
        if (cur < pos) sink(cur, pos, 0);
        sink(g0, g1, 1);
        cur = m1;
    }
    if (cur < ee) sink(cur, ee, 0);
}

// ------------------------------------------------------------------------------------------
// Normalisers: RegexNormalization for single-character patterns and CharsMapNormalization.
// Both are "at an active byte position: consume c bytes, emit o bytes" machines; norm_eval() is that step, shared by
// the host scan below (tests) and the kernel (kernels_norm.cuh), which evaluates 32 positions at once and resolves which
// of them the scan really visits.
// ------------------------------------------------------------------------------------------
// GPT-2 byte -> code point map (public algorithm of the GPT-2 encoder: printable Latin-1 bytes map to themselves, the other
// 68 bytes to U+0100 + n in byte order); the table BytesToChars indexes (reference src/bytes_to_chars.cpp:11-268).
inline void gpt2_build_byte_codepoints(uint16_t* cp) {
    int n = 0;
    for (int b = 0; b < 256; ++b) {
        const bool keep = (b >= '!' && b <= '~') || (b >= 0xA1 && b <= 0xAC) || (b >= 0xAE && b <= 0xFF);
        cp[b] = keep ? (uint16_t)b : (uint16_t)(256 + n++);
    }
}

enum : uint8_t { NC_DEL = 1, NC_S = 2, NC_HAN = 4, NC_MN = 8 };   // flags of the second class table (unicode_norm_ranges.inc)
enum : int { NORM_CLASS = 0, NORM_CHARSMAP = 1, NORM_B2C = 2, NORM_UTF8 = 3, NORM_C2B = 4 };   // + BytesToChars, UTF8Validate, CharsToBytes: the same kind of scan

struct NormRule {
    int32_t kind;
    // NORM_CLASS: the search pattern matches exactly one character of a class, optionally only at the start of the string
    // (reference patterns: python/openvino_tokenizers/tokenizer_pipeline.py:230-278); the replacement is pre + [the character] + post.
    ClassTables cls;
    int32_t literal_cp;      // >= 0: the class is this single code point
    uint8_t mask;            // else: characters whose NC_* flags intersect mask ...
    uint8_t any;             // ... or every character ([\s\S])
    uint8_t negate;          // the class is negated (\S, [^ ])
    uint8_t anchored;        // ^: only the first character of the string
    uint8_t global;          // PCRE2_SUBSTITUTE_GLOBAL, else only the first match
    uint8_t keep;            // the replacement refers to the matched character
    uint8_t pre_len, post_len;
    uint8_t pre[16], post[16];
    // NORM_CHARSMAP: sentencepiece precompiled charsmap = Darts-clone double array + '\0'-separated replacements
    const uint32_t* units;
    uint32_t n_units;
    const uint8_t* normalized;
    uint32_t n_normalized;
    // per-ASCII-byte shortcuts (a chunk of 32 ASCII bytes skips the general step): NORM_CLASS: aflag = the byte's NC_* flags;
    // NORM_CHARSMAP: amap = the byte it maps to when the rule for it is "one ASCII byte -> one ASCII byte" (or none), aflag =
    // NA_* reasons why the byte needs the trie walk
    // (divergent indices: the table lives in global memory, not in the parameter block)
    const uint8_t* atab;     // [0,128) amap, [128,256) aflag
};
enum : uint8_t { NT_DEL = 0xFE, NT_GENERAL = 0xFF };                     // fates of an ASCII byte in a composed chain table (tables.cpp compose_norm_chain)
enum : uint8_t { NA_COMPLEX = 1, NA_ASCII_KIDS = 2, NA_OTHER_KIDS = 4 };   // replacement not 1 ASCII byte / longer rules continue with an ASCII / non-ASCII byte

struct NormStep {
    int32_t consumed;    // input bytes
    int32_t olen;        // output bytes
    int32_t src;         // >= 0: copy from normalized + src; -1: the rule's pre/char/post; -2: input bytes verbatim; -3: olen / 3 x U+FFFD;
                         // -4: the byte's BytesToChars character (1 or 2 bytes); -5: the byte of a CharsToBytes pair
    uint8_t matched;     // NORM_CLASS: the character matched the class
};

// Well-formed UTF-8 character at s[i] (strict: no overlongs / surrogates, like sentencepiece's DecodeUTF8)?  Returns its
// length, or 0 if malformed.
B2_HD int utf8_strict_len(const uint8_t* s, int i, int end, uint32_t& cp) {
    const uint32_t b0 = s[i];
    if (b0 < 0x80) { cp = b0; return 1; }
    const int left = end - i;
    if ((b0 & 0xE0) == 0xC0) {
        if (left < 2 || !is_cont_byte(s[i + 1])) return 0;
        cp = ((b0 & 0x1Fu) << 6) | (s[i + 1] & 0x3Fu);
        return cp >= 0x80 ? 2 : 0;
    }
    if ((b0 & 0xF0) == 0xE0) {
        if (left < 3 || !is_cont_byte(s[i + 1]) || !is_cont_byte(s[i + 2])) return 0;
        cp = ((b0 & 0x0Fu) << 12) | ((s[i + 1] & 0x3Fu) << 6) | (s[i + 2] & 0x3Fu);
        return (cp >= 0x800 && (cp < 0xD800 || cp >= 0xE000)) ? 3 : 0;
    }
    if ((b0 & 0xF8) == 0xF0) {
        if (left < 4 || !is_cont_byte(s[i + 1]) || !is_cont_byte(s[i + 2]) || !is_cont_byte(s[i + 3])) return 0;
        cp = ((b0 & 0x07u) << 18) | ((s[i + 1] & 0x3Fu) << 12) | ((s[i + 2] & 0x3Fu) << 6) | (s[i + 3] & 0x3Fu);
        return (cp >= 0x10000 && cp <= 0x10FFFF) ? 4 : 0;
    }
    return 0;
}

// Darts::DoubleArray::commonPrefixSearch + the "longest rule" loop of sentencepiece's Normalizer::NormalizePrefix
// (at most 32 prefixes are reported; the last reported one is the longest).
B2_HD int charsmap_longest(const NormRule& R, const uint8_t* s, int i, int end, uint32_t& value) {
    if (!R.n_units) return 0;
    uint32_t pos = 0, u = R.units[0];
    pos ^= (u >> 10) << ((u & (1u << 9)) >> 6);
    int mlen = 0, found = 0;
    for (int k = i; k < end; ++k) {
        const uint32_t c = s[k];
        pos ^= c;
        if (pos >= R.n_units) break;
        u = R.units[pos];
        if ((u & ((1u << 31) | 0xFFu)) != c) break;
        pos ^= (u >> 10) << ((u & (1u << 9)) >> 6);
        if ((u >> 8) & 1u) {
            if (found < 32 && pos < R.n_units) { mlen = k + 1 - i; value = R.units[pos] & 0x7FFFFFFFu; }
            ++found;
        }
    }
    return mlen;
}

// One step of the scan at byte i of the string [b, e).  `first_done`: a non-global rule already replaced its match.
B2_HD NormStep norm_eval(const NormRule& R, const uint8_t* s, int b, int i, int e, bool first_done) {
    NormStep st;
    st.matched = 0;
    uint32_t cp = 0;
    if (R.kind == NORM_CHARSMAP) {
        uint32_t value = 0;
        const int m = charsmap_longest(R, s, i, e, value);
        if (m > 0) {
            int l = 0;
            if (value < R.n_normalized) while (value + (uint32_t)l < R.n_normalized && R.normalized[value + l]) ++l;
            st.consumed = m; st.olen = l; st.src = (int32_t)value;
            return st;
        }
        const int l = utf8_strict_len(s, i, e, cp);
        if (l == 0) { st.consumed = 1; st.olen = 3; st.src = -3; }
        else { st.consumed = l; st.olen = l; st.src = -2; }
        return st;
    }
    if (R.kind == NORM_B2C) {          // reference src/bytes_to_chars.cpp:284-339: every byte becomes one character of the GPT-2 byte map
        const uint16_t c = reinterpret_cast<const uint16_t*>(R.normalized)[s[i]];
        st.consumed = 1; st.olen = c >= 0x80 ? 2 : 1; st.src = -4;
        return st;
    }
    if (R.kind == NORM_C2B) {          // reference src/chars_to_bytes.cpp:52-60: a byte >= 128 takes the following byte with it (even past the element's end)
        st.olen = 1;
        if (s[i] < 128) { st.consumed = 1; st.src = -2; } else { st.consumed = 2; st.src = -5; }
        return st;
    }
    if (R.kind == NORM_UTF8) {         // reference src/utf8_validate.cpp:18-137 as "at a start byte: consume c, emit o"; R.global = replace mode
        const uint32_t c = s[i];
        const int bad = R.global ? 3 : 0;
        st.src = -3;
        if (c < 128) { st.consumed = 1; st.olen = 1; st.src = -2; return st; }
        const int num = (c >> 5) == 0b110 ? 2 : (c >> 4) == 0b1110 ? 3 : (c >> 3) == 0b11110 ? 4 : 0;
        if (!num) { st.consumed = 1; st.olen = bad; return st; }               // a continuation or 11111xxx byte at a start position (:80-87)
        uint32_t v = num == 2 ? (c & 0b11111u) << 6 : num == 3 ? (c & 0b1111u) << 12 : (c & 0b111u) << 18;
        for (int j = 1; j < num; ++j) {
            if (i + j >= e || (s[i + j] >> 6) != 0b10) { st.consumed = j; st.olen = bad; return st; }   // broken (:96-105; the byte starts anew) or unfinished (:131-134)
            v |= (uint32_t)(s[i + j] & 0b111111u) << (6 * (num - 1 - j));
        }
        st.consumed = num;
        const uint32_t starts = num == 2 ? 0x80u : num == 3 ? 0x800u : 0x10000u;
        if (v < starts) st.olen = bad * num;                                      // overlong: one replacement per byte (:112-122)
        else { st.olen = num; st.src = -2; }
        return st;
    }
    // malformed bytes never match and pass through one at a time (PCRE2 runs with NO_UTF_CHECK: out of contract)
    const int l = utf8_strict_len(s, i, e, cp);
    st.consumed = l ? l : 1;
    st.olen = st.consumed;
    st.src = -2;
    if (!l || (R.anchored && i != b) || (!R.global && first_done)) return st;
    bool in_class;
    if (R.any) in_class = true;
    else if (R.literal_cp >= 0) in_class = cp == (uint32_t)R.literal_cp;
    else {
        const uint8_t f = cp < 0x80 ? R.cls.ascii[cp] : R.cls.stage2[(uint32_t)R.cls.stage1[cp >> 8] * 256u + (cp & 255u)];
        in_class = (f & R.mask) != 0;
    }
    if (in_class == (R.negate != 0)) return st;
    st.matched = 1;
    st.src = -1;
    st.olen = (int32_t)R.pre_len + (R.keep ? l : 0) + (int32_t)R.post_len;
    return st;
}

// Writes the step's output bytes.
B2_HD void norm_emit(const NormRule& R, const NormStep& st, const uint8_t* s, int i, uint8_t* out) {
    if (st.src >= 0) { for (int k = 0; k < st.olen; ++k) out[k] = R.normalized[st.src + k]; }
    else if (st.src == -2) { for (int k = 0; k < st.olen; ++k) out[k] = s[i + k]; }
    else if (st.src == -3) { for (int k = 0; k < st.olen; k += 3) { out[k] = 0xEF; out[k + 1] = 0xBF; out[k + 2] = 0xBD; } }
    else if (st.src == -5) {            // R.normalized = pair map [4 * 64] (src/chars_to_bytes.cpp:20-29), R.n_units = size of the chars buffer
        const int fi = (int)s[i] - 194, si = ((uint32_t)(i + 1) < R.n_units ? (int)s[i + 1] : 128) - 128;
        out[0] = (fi >= 0 && fi < 4 && si >= 0 && si < 64) ? R.normalized[fi * 64 + si] : 0;      // outside the map: 0 (the reference reads out of bounds)
    }
    else if (st.src == -4) {
        const uint16_t c = reinterpret_cast<const uint16_t*>(R.normalized)[s[i]];
        if (c < 0x80) out[0] = (uint8_t)c;
        else { out[0] = (uint8_t)(0xC0 | (c >> 6)); out[1] = (uint8_t)(0x80 | (c & 0x3F)); }
    }
    else {
        int o = 0;
        for (int k = 0; k < R.pre_len; ++k) out[o++] = R.pre[k];
        if (R.keep) for (int k = 0; k < st.consumed; ++k) out[o++] = s[i + k];
        for (int k = 0; k < R.post_len; ++k) out[o++] = R.post[k];
    }
}

// Sequential scan of one string (host tests; the kernel does the same 32 positions at a time).  Returns the output length;
// writes when out != nullptr.
B2_HD int norm_string(const NormRule& R, const uint8_t* s, int b, int e, uint8_t* out) {
    int o = 0;
    bool done = false;
    for (int i = b; i < e;) {
        const NormStep st = norm_eval(R, s, b, i, e, done);
        if (out) norm_emit(R, st, s, i, out + o);
        done = done || st.matched;
        o += st.olen;
        i += st.consumed;
    }
    return o;
}

}  // namespace b200tok
