// kernels_tail.cuh — the post-tokenizer tail: Truncate, CombineSegments, RaggedToDense and their fusion.
//   Truncate          reference src/truncate.cpp:37-147
//   CombineSegments   reference src/combine_segments.cpp:36-134   (i32 elements)
//   RaggedToDense     reference src/ragged_to_dense.cpp:70-174    (i32 elements, no trailing dense dimensions)
// All three are offset arithmetic plus coalesced copies: HBM-bound, no shared memory needed.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200tok {

enum : int { TRUNC_RIGHT = 0, TRUNC_LEFT = 1 };
enum : int { TRUNC_ONLY_FIRST = 0, TRUNC_ONLY_SECOND = 1, TRUNC_LONGEST_FIRST = 2 };

// One thread per row; edits begins / ends in place like the reference's aliased outputs.
__global__ void truncate_kernel(int num_inputs, int32_t* b0, int32_t* e0, int32_t* b1, int32_t* e1, int64_t n, int32_t max_length,
                                int side, int mode) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (num_inputs == 1) {   // :57-68
        const int32_t len = e0[i] - b0[i];
        const int32_t t = len < max_length ? len : max_length;
        if (side == TRUNC_RIGHT) e0[i] = b0[i] + t;
        else b0[i] = e0[i] - t;
        return;
    }
    const int32_t fb = b0[i], fe = e0[i], sb = b1[i], se = e1[i];
    const int32_t first = fe - fb, second = se - sb;
    if (first + second <= max_length) return;   // :83
    const int32_t first_rem = (max_length % 2) * (first >= second), second_rem = (max_length % 2) * (first < second);
    const int32_t half = max_length / 2, half_up = max_length / 2 + max_length % 2;
    int32_t nf = first, ns = second;   // new lengths
    if (mode == TRUNC_ONLY_FIRST) { if (first > max_length) nf = max_length; }
    else if (mode == TRUNC_ONLY_SECOND) { if (second > max_length) ns = max_length; }
    else if (first >= half_up && second <= half) nf = max_length - second;
    else if (first < half_up && second > half) ns = max_length - first;
    else { nf = half + first_rem; ns = half + second_rem; }
    if (side == TRUNC_RIGHT) { e0[i] = fb + nf; e1[i] = sb + ns; }
    else { b0[i] = fe - nf; b1[i] = se - ns; }
}

constexpr int kMaxSegments = 16;
struct SegmentList {
    const int32_t* begins[kMaxSegments];
    const int32_t* ends[kMaxSegments];
    const int32_t* elems[kMaxSegments];
    int32_t broadcast[kMaxSegments];   // 1 => the segment has one row used for every output row (:102-104)
    int32_t ids[kMaxSegments];
    int32_t num;
};

// Row lengths of the combined tensor.
__global__ void combine_len_kernel(const SegmentList S, int64_t rows, int32_t* len) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    int32_t t = 0;
    for (int j = 0; j < S.num; ++j) {
        const int64_t r = S.broadcast[j] ? 0 : i;
        const int32_t l = S.ends[j][r] - S.begins[j][r];
        t += l > 0 ? l : 0;
    }
    len[i] = t;
}
// One warp per row: copy every segment's slice, label it with the segment's id (:98-121).
__global__ void combine_copy_kernel(const SegmentList S, int64_t rows, const int32_t* out_begins, const int32_t* len, int32_t* out_ends,
                                    int32_t* out_elems, int32_t* out_ids, int64_t* total) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < rows; i += nwarps) {
        int64_t off = out_begins[i];
        if (lane == 0) {
            out_ends[i] = (int32_t)(off + len[i]);
            if (i == rows - 1) *total = off + len[i];
        }
        for (int j = 0; j < S.num; ++j) {
            const int64_t r = S.broadcast[j] ? 0 : i;
            const int32_t b = S.begins[j][r], l = S.ends[j][r] - b;
            const int32_t* src = S.elems[j] + b;
            const int32_t id = S.ids[j];
            for (int t = lane; t < l; t += 32) { out_elems[off + t] = src[t]; out_ids[off + t] = id; }
            off += l > 0 ? l : 0;
        }
    }
}

// One thread per output element (:128-165).  Reads that the reference would do beyond the elems buffer (pad_max_length
// with a short row) yield the default value.
__global__ void ragged_to_dense_kernel(const int32_t* begins, const int32_t* ends, int64_t n, const int32_t* elems, int64_t n_elems,
                                       int32_t target_dim, int32_t default_value, int pad_right, int pad_max_length, int32_t* out, uint8_t* mask) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * (int64_t)target_dim) return;
    const int64_t i = idx / target_dim;
    const int32_t t = (int32_t)(idx - i * target_dim);
    const int32_t b = begins[i];
    const int64_t data_len = (int64_t)ends[i] - b;
    const int64_t target_len = pad_max_length ? target_dim : (data_len < target_dim ? (data_len > 0 ? data_len : 0) : target_dim);
    const int64_t k = pad_right ? t : t - (target_dim - target_len);     // index inside the row, valid if 0 <= k < target_len
    const bool data = k >= 0 && k < target_len;
    int32_t v = default_value;
    if (data) { const int64_t src = (int64_t)b + k; v = (src >= 0 && src < n_elems) ? elems[src] : default_value; }
    out[idx] = v;
    if (mask) mask[idx] = data ? 1 : 0;
}

// Truncate (single input) -> CombineSegments(prefix constants, tokens, suffix constants) -> RaggedToDense in one pass:
// one thread per element of the dense [rows, target_dim] result, reading the ragged ids exactly once.
struct PostParams {
    int32_t max_length; int32_t trunc_left;
    int32_t prefix[8]; int32_t n_prefix;
    int32_t suffix[8]; int32_t n_suffix;
    int32_t target_dim; int32_t pad_value; int32_t pad_right;
};
__global__ void post_dense_kernel(const PostParams Q, const int32_t* begins, const int32_t* ends, int64_t rows, const int32_t* ids, int64_t n_ids,
                                  int32_t* out, uint8_t* mask) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * (int64_t)Q.target_dim) return;
    const int64_t i = idx / Q.target_dim;
    const int32_t t = (int32_t)(idx - i * Q.target_dim);
    int32_t b = begins[i], e = ends[i];
    {   // Truncate, one input
        const int32_t len = e - b, tl = len < Q.max_length ? len : Q.max_length;
        if (Q.trunc_left) b = e - tl; else e = b + tl;
    }
    const int32_t tok_len = e - b > 0 ? e - b : 0;
    const int64_t data_len = (int64_t)Q.n_prefix + tok_len + Q.n_suffix;                 // CombineSegments row
    const int64_t target_len = data_len < Q.target_dim ? data_len : Q.target_dim;         // RaggedToDense
    const int64_t k = Q.pad_right ? t : t - (Q.target_dim - target_len);
    const bool data = k >= 0 && k < target_len;
    int32_t v = Q.pad_value;
    if (data) {
        if (k < Q.n_prefix) v = Q.prefix[k];
        else if (k < Q.n_prefix + tok_len) { const int64_t src = (int64_t)b + (k - Q.n_prefix); v = (src >= 0 && src < n_ids) ? ids[src] : Q.pad_value; }
        else v = Q.suffix[k - Q.n_prefix - tok_len];
    }
    out[idx] = v;
    if (mask) mask[idx] = data ? 1 : 0;
}

}  // namespace b200tok
