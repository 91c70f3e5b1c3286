// tables.hpp — host-side construction of the read-only device tables (pure C++17, no CUDA).
// Corresponds to the reference's lazy first-evaluate initialisation:
//   BPETokenizer::evaluate call_once block   src/bpe_tokenizer.cpp:50-120
//   BPETokenizerImpl ctor                     src/bpe_tokenizer.cpp:341-388
//   WordpieceTokenizer trie construction      src/wordpiece_tokenizer.cpp:50-73
//   VocabEncoder map construction             src/vocab_encoder.cpp:63-78
//   RegexSplit::compile_pattern_if_necessary  src/regex_split.cpp:26-39
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#include "../../include/b200tok.h"
#include "tok_core.cuh"

namespace b200tok {

struct HostClassTables {
    std::vector<uint8_t> ascii;
    std::vector<uint16_t> stage1;
    std::vector<uint8_t> stage2;
    ClassTables view() const { return ClassTables{ascii.data(), stage1.data(), stage2.data()}; }
};
const HostClassTables& host_class_tables();
const HostClassTables& host_norm_class_tables();   // NC_* flags (normaliser patterns)

struct HostTrie {
    std::vector<int32_t> first, value, edge_child, root_child;
    std::vector<uint8_t> edge_byte;
    // entries: (key bytes, id); later duplicates of a key overwrite earlier ones; empty keys ignored
    void build(const std::vector<std::pair<std::string, int32_t>>& entries);
    FlatTrie view() const {
        return FlatTrie{first.data(), value.data(), edge_byte.data(), edge_child.data(), root_child.data()};
    }
    size_t n_nodes() const { return value.size(); }
    // the same trie in rank-indexed form (breadth-first numbering: a node's children are consecutive)
    std::vector<RankNode> rank_nodes;
    std::vector<int32_t> rank_root, rank_val1;
    std::vector<RankJump> rank_jump;
    void build_rank();
    RankTrie rank_view() const { return RankTrie{rank_nodes.data(), rank_root.data(), rank_jump.empty() ? nullptr : rank_jump.data(), rank_val1.data()}; }
};

struct HostBpe {
    std::vector<int32_t> byte_sym, byte_miss;
    HostTrie trie;
    std::vector<MergeSlot> slots;
    std::vector<int32_t> rank_newid;
    std::vector<uint32_t> pair_rank;
    std::vector<uint32_t> pair_bits;
    int32_t newid_base = -1;
    int64_t max_id = 0;                 // largest token id any symbol can take (vocab + added tokens)
    uint32_t mask = 0;
    std::string end_suffix;
    int32_t unk_id = -1;
    int64_t n_merges = 0;
    int64_t n_duplicate_products = 0;   // merges whose product token another merge also produces (tie hazard)
    bool bytes_only = false;            // every byte symbolises without a trie walk
    BpeTables view() const {
        return BpeTables{byte_sym.data(), byte_miss.data(), pair_rank.data(), trie.view(), MergeTable{slots.data(), mask, rank_newid.data(), n_duplicate_products > 0 ? 1 : 0}, pair_bits.data(), newid_base};
    }
};
// returns B200TOK_OK or an error code (message in err)
int build_bpe(const b200tok_bpe_desc& d, HostBpe& out, std::string& err);

struct HostWordpiece {
    HostTrie root, sub;
    int32_t max_bytes = 100;
    WordpieceTables view() const { return WordpieceTables{root.rank_view(), sub.rank_view(), max_bytes}; }
};
int build_wordpiece(const b200tok_wordpiece_desc& d, HostWordpiece& out, std::string& err);

// String -> value open-addressing table for VocabEncoder (FNV-1a 64; full key compare on hit).
struct VocabEncSlot { uint64_t hash; int32_t begin, len; int64_t value; };   // len < 0 => empty
struct HostVocabEnc {
    std::vector<VocabEncSlot> slots;
    std::vector<uint8_t> key_bytes;
    uint32_t mask = 0;
    int32_t max_len = 0;
};
int build_vocabenc(const b200tok_vocabenc_desc& d, HostVocabEnc& out, std::string& err);

// Host-buffer fast path qualification: rows contiguous (rb[0] == 0, rb[r + 1] == re[r], re[B - 1] == E), elements inside [0, N],
// increasing and non-overlapping — what StringTensorUnpack / RegexSplit produce.  Branch-free so that the compiler vectorises it
// (the early-exit form cost 0.19 ms per 65 536-row call, 7 % of the whole host-to-host C1 call).
bool contiguous_batch(const int32_t* rb, const int32_t* re, const int32_t* eb, const int32_t* ee, int64_t B, int64_t E, int64_t N);

// General split patterns: program of the regex machine (regex_vm.cuh), compiled by regex_compile.cpp.
struct HostVm {
    std::vector<VmInst> code;
    std::vector<VmSet> sets;
    std::vector<uint32_t> ranges;
};
struct HostGcTables { std::vector<uint16_t> stage1; std::vector<uint8_t> stage2; };     // Unicode general category per code point
const HostGcTables& host_gc_tables();
int compile_regex(const std::string& pattern, HostVm& out, std::string& err);

enum SplitMode : int { MODE_REMOVED = 0, MODE_ISOLATED = 1, MODE_MERGED_PREV = 2, MODE_MERGED_NEXT = 3 };
struct HostSplit {
    SplitSpec spec{};
    HostVm vm;             // spec.pat == PAT_VM
    // spec with vm pointing at this object's host vectors (host harness / tests)
    SplitSpec host_spec() const {
        SplitSpec s = spec;
        if (s.pat == PAT_VM) s.vm = VmProgram{vm.code.data(), vm.sets.data(), vm.ranges.data(), host_gc_tables().stage1.data(), host_gc_tables().stage2.data(), (int32_t)vm.code.size()};
        return s;
    }
    int mode = MODE_REMOVED;
    bool invert = false;
    int max_splits = -1;
    bool repeat = false;   // "contiguous" rewrite (p)+ applied (src/regex_split.cpp:33-37)
    std::string pattern;
};
int parse_split(const b200tok_regexsplit_desc& d, HostSplit& out, std::string& err);

// SpecialTokensSplit pattern (src/special_tokens_split.cpp + the converter's tokenizer_pipeline.py:138-158): an alternation
// of groups  (?:\s*)?(tok|tok|...)(?:\s*)?  of quote_meta-escaped literals.  value of a trie entry = position of the token
// inside its group (PCRE2 takes the first alternative that matches).
constexpr int kMaxSpecialGroups = 8;
struct HostSpecialGroup { bool strip_left = false, strip_right = false; std::vector<std::string> tokens; HostTrie trie; };
struct HostSpecial {
    std::vector<HostSpecialGroup> groups;
    std::vector<uint32_t> first;     // [8]: bytes at which a match can start
    bool ws_token = false;           // some token of a strip_left group starts with a whitespace byte (forces full backtracking)
    std::string pattern;
};
int parse_special(const char* pattern, int64_t len, HostSpecial& out, std::string& err);

// Normalisers.  rule.cls / units / normalized pointers are filled in by the user of the tables (host or device views).
struct HostNorm {
    NormRule rule{};
    std::vector<uint32_t> units;
    std::vector<uint8_t> normalized;
    std::vector<uint8_t> atab;       // [0,128) amap, [128,256) aflag (NormRule::atab)
};
int parse_regex_norm(const char* search, int64_t slen, const char* replace, int64_t rlen, int global_replace, HostNorm& out, std::string& err);
bool compose_norm_chain(const HostNorm* const* ops, int n_ops, uint8_t* T /* [128] */);
int parse_charsmap(const uint8_t* blob, int64_t len, int add_dummy_prefix, int remove_extra_whitespaces, int escape_whitespaces, HostNorm& out, std::string& err);

inline uint64_t fnv1a64(const uint8_t* p, int64_t n) {
    uint64_t h = 1469598103934665603ull;
    for (int64_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

}  // namespace b200tok
