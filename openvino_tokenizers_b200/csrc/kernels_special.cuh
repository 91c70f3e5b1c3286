// kernels_special.cuh — SpecialTokensSplit (reference src/special_tokens_split.cpp:61-162, src/utils.cpp:423-461).
//
// The op's pattern is an alternation of groups  (?:\s*)?(tok|tok|...)(?:\s*)?  of literal special tokens, matched by
// PCRE2 with leftmost / first-alternative semantics.  Here: one warp per row; every lane evaluates "the match that would
// start at my byte" (a first-byte filter, then per group: optional greedy whitespace with backtracking, a walk of the
// group's token trie taking the earliest alternative, optional trailing whitespace); the reference's sequential
// scan (:115-145) is then resolved per 32-position chunk from the ballot of match starts.  Pieces go to the row-local
// worst-case slots shared with the RegexSplit kernel (tmp_a / tmp_b / tmp_c) and are compacted by the same pass.
#pragma once
#include "kernels.cuh"

namespace b200tok {

constexpr int kSpecialGroups = 8;
struct SpecialTables {
    FlatTrie trie[kSpecialGroups];     // value = position of the token inside its group (smaller = earlier alternative)
    uint8_t strip_left[kSpecialGroups], strip_right[kSpecialGroups];
    int32_t n_groups;
    uint32_t first[8];                 // bytes at which a match can start
    int32_t ws_token;                  // a token of a strip_left group starts with whitespace: full backtracking needed
};

// Earliest alternative among the group's tokens matching at chars[q..ee); they all lie on one trie path.
__device__ __forceinline__ bool special_token_at(const FlatTrie& t, const uint8_t* chars, int q, int ee, int& tok_end) {
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
__device__ __forceinline__ int special_ws_len(const uint8_t* chars, int i, int ee, const ClassTables& T) {
    const uint8_t b = chars[i];
    if (b < 0x80) return (T.ascii[b] & C_S) ? 1 : 0;
    if (b < 0xC2 || !(char_class(chars, i, ee, T) & C_S)) return 0;
    return b >= 0xF0 ? 4 : b >= 0xE0 ? 3 : 2;
}

// The match starting exactly at pos, or m1 = 0.  [g0, g1) = the token (capture group), [pos, m1) = the full match.
__device__ __forceinline__ void special_match_at(const SpecialTables& ST, const ClassTables& T, const uint8_t* chars, int pos, int ee,
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

__global__ void __launch_bounds__(256) special_split_kernel(const __grid_constant__ RowParams P, const __grid_constant__ SpecialTables ST) {
    const int lane = threadIdx.x & 31;
    const ClassTables T = P.cls;
    for (;;) {
        int row = 0;
        if (lane == 0) row = atomicAdd(&P.status[ST_TICKET], 1);
        row = __shfl_sync(FULL, row, 0);
        if (row >= P.n_rows) break;
        const int p0 = P.rb[row], p1 = P.re[row];
        int64_t base;
        if (P.direct_base) {
            base = p1 > p0 ? (int64_t)(P.begins[p0] - P.direct_byte0) + (int64_t)(p0 - P.direct_elem0) * P.direct_extra : 0;
            if (lane == 0) const_cast<int32_t*>(P.row_base)[row] = (int32_t)base;
        } else base = P.row_base[row];
        int emitted = 0;
        auto emit = [&](int b, int e, int skip) {     // lane 0 writes; every lane counts
            if (lane == 0) {
                const int64_t o = base + emitted;
                if (o < P.tmp_cap) { P.tmp_a[o] = b; P.tmp_b[o] = e; P.tmp_c[o] = (uint8_t)skip; }
                else atomicOr(&P.status[ST_ERROR], ERR_TMP_OVERFLOW);
            }
            ++emitted;
        };
        for (int p = p0; p < p1; ++p) {
            const int eb = P.begins[p], ee = P.ends[p];
            if (P.skips && P.skips[p]) { emit(eb, ee, 1); continue; }     // :109-112
            int cur = eb;
            for (int c0 = eb; c0 < ee; c0 += 32) {
                const int pos = c0 + lane;
                int m1 = 0, g0 = 0, g1 = 0;
                if (pos < ee && pos >= cur) {      // (a match can only start at or after the end of the previous one)
                    const uint8_t b = P.chars[pos];
                    if (!is_cont_byte(b) && ((ST.first[b >> 5] >> (b & 31)) & 1u)) special_match_at(ST, T, P.chars, pos, ee, m1, g0, g1);
                }
                uint32_t mask = __ballot_sync(FULL, m1 > pos);      // empty matches end the scan (:118); tokens are non-empty
                while (mask) {
                    const int l = __ffs(mask) - 1;
                    const int mm1 = __shfl_sync(FULL, m1, l), mg0 = __shfl_sync(FULL, g0, l), mg1 = __shfl_sync(FULL, g1, l);
                    const int m0 = c0 + l;
                    if (m0 >= cur) {
                        if (cur < m0) emit(cur, m0, 0);
                        emit(mg0, mg1, 1);
                        cur = mm1;
                    }
                    mask &= mask - 1;
                }
            }
            if (cur < ee) emit(cur, ee, 0);
        }
        if (lane == 0) { P.row_ext[row] = emitted; P.row_cnt[row] = emitted; P.row_flag[row] = 0; }
    }
}

}  // namespace b200tok
