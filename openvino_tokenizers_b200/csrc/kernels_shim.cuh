// kernels_shim.cuh — byte-level shims and detokenizer tail:
//   BytesToChars   reference src/bytes_to_chars.cpp:284-339   (GPT-2 byte -> printable-char map, 1 or 2 UTF-8 bytes per byte)
//   CharsToBytes   reference src/chars_to_bytes.cpp:31-68     (its inverse; fuses the ragged dimension: one string per row)
//   FuzeRagged     reference src/fuze.cpp:20-40
//   UTF8Validate   reference src/utf8_validate.cpp:18-137     (replace / drop malformed sequences)
// All are per-string state machines with data-dependent output sizes: lengths pass, cub scan of the lengths, write pass.
// BytesToChars, UTF8Validate and CharsToBytes are scans of the form "at a start byte: consume c bytes, emit o bytes" and run on the
// warp-per-string kernel of the normalisers (kernels_norm.cuh: 32 positions at a time, ASCII chunks table-driven);
// so do the elements of CharsToBytes, whose rows (it fuses the ragged dimension) are then stitched by c2b_rows_kernel;
// FuzeRagged is offset arithmetic.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tok_core.cuh"      // gpt2_build_byte_codepoints

namespace b200tok {

// GPT-2 byte <-> unicode map (public algorithm of the GPT-2 encoder: printable Latin-1 bytes map to themselves, the other 68
// bytes to U+0100 + n in byte order).  code point of byte b:
__host__ __device__ inline uint32_t gpt2_byte_codepoint(uint32_t b, const uint16_t* shifted /* [256] or null */) {
    return shifted ? shifted[b] : b;
}
struct ByteCharTables {
    const uint16_t* cp;        // [256] code point of every byte (all < 0x800: one or two UTF-8 bytes)
    const uint8_t* pair_map;   // [4 * 64] byte of the 2-byte sequence (first - 194, second - 128); src/chars_to_bytes.cpp:20-29
};

// ---- BytesToChars runs on the warp-per-string scan of kernels_norm.cuh (rule NORM_B2C, tok_core.cuh norm_eval) ----

// ---- CharsToBytes: the elements run on the scan kernel (rule NORM_C2B); rows take the offsets of their first / last element ----
__global__ void c2b_rows_kernel(const int32_t* rb, const int32_t* re, int64_t rows, const int32_t* off, const int32_t* len, int64_t n,
                                int32_t* out_begins, int32_t* out_ends) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const int32_t a = rb[r], b = re[r];
    const int32_t start = a < n ? off[a] : (n > 0 ? off[n - 1] + len[n - 1] : 0);
    out_begins[r] = start;
    out_ends[r] = b > a ? off[b - 1] + len[b - 1] : start;
}

// ---- FuzeRagged ----
__global__ void fuze_ragged_kernel(const int32_t* rb, const int32_t* re, int64_t rows, const int32_t* begins, const int32_t* ends,
                                   int32_t* out_begins, int32_t* out_ends) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    out_begins[r] = begins[rb[r]];
    out_ends[r] = ends[re[r] > rb[r] ? re[r] - 1 : re[r]];      // src/fuze.cpp:36-37 (an empty row reads element re[r])
}

// ---- UTF8Validate runs on the same scan (rule NORM_UTF8): the reference's byte automaton restated per start byte ----

// Device-side form of the host check "rows cover the elements contiguously and in order" (BytesToChars / CharsToBytes take their
// row extents from an element-order scan, which is the reference's row-order walk only under that condition): flag != 0 = violated.
__global__ void rows_partition_check_kernel(const int32_t* rb, const int32_t* re, int64_t n_rows, int64_t n_elems, int32_t* flag) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const int32_t b = rb[r], e = re[r];
    const int32_t prev_end = r == 0 ? 0 : re[r - 1];
    if (b != prev_end || e < b || (r == n_rows - 1 && e != n_elems)) atomicOr(flag, 1);
}

}  // namespace b200tok
