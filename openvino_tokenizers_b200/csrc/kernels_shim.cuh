// kernels_shim.cuh — byte-level shims and detokenizer tail:
//   BytesToChars   reference src/bytes_to_chars.cpp:284-339   (GPT-2 byte -> printable-char map, 1 or 2 UTF-8 bytes per byte)
//   CharsToBytes   reference src/chars_to_bytes.cpp:31-68     (its inverse; fuses the ragged dimension: one string per row)
//   FuzeRagged     reference src/fuze.cpp:20-40
//   UTF8Validate   reference src/utf8_validate.cpp:18-137     (replace / drop malformed sequences)
// All are per-string state machines with data-dependent output sizes: lengths pass, cub scan of the lengths, write pass.
// BytesToChars and UTF8Validate are scans of the form "at a start byte: consume c bytes, emit o bytes" and run on the
// warp-per-string kernel of the normalisers (kernels_norm.cuh: 32 positions at a time, ASCII chunks table-driven);
// CharsToBytes (which fuses the ragged dimension) and FuzeRagged keep their one-thread-per-element kernels below.
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

// ---- CharsToBytes: per element lengths, rows take the offsets of their first / last element ----
// src/chars_to_bytes.cpp:52-60: a byte >= 128 consumes the following byte too (even past the element's end, like the reference).
__global__ void c2b_len_kernel(const int32_t* begins, const int32_t* ends, const uint8_t* chars, int64_t n, int32_t* len) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t l = 0;
    for (int32_t k = begins[i]; k < ends[i]; ++k) { if (chars[k] >= 128) ++k; ++l; }
    len[i] = l;
}
__global__ void c2b_write_kernel(const int32_t* begins, const int32_t* ends, const uint8_t* chars, int64_t n, int64_t n_chars,
                                 const uint8_t* pair_map, const int32_t* off, const int32_t* len, uint8_t* out, int64_t cap, int64_t* total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t o = off[i];
    if (i == n - 1) *total = o + len[i];
    if (o + len[i] > cap) return;
    for (int32_t k = begins[i]; k < ends[i]; ++k) {
        const uint8_t f = chars[k];
        if (f < 128) { out[o++] = f; continue; }
        ++k;
        const uint8_t s = k < n_chars ? chars[k] : 128;
        const int fi = (int)f - 194, si = (int)s - 128;
        out[o++] = (fi >= 0 && fi < 4 && si >= 0 && si < 64) ? pair_map[fi * 64 + si] : 0;   // outside the map: 0 (the reference reads out of bounds)
    }
}
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

}  // namespace b200tok
