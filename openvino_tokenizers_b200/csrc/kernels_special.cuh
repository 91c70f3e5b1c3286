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
