// kernels_shim.cuh — byte-level shims and detokenizer tail:
//   BytesToChars   reference src/bytes_to_chars.cpp:284-339   (GPT-2 byte -> printable-char map, 1 or 2 UTF-8 bytes per byte)
//   CharsToBytes   reference src/chars_to_bytes.cpp:31-68     (its inverse; fuses the ragged dimension: one string per row)
//   FuzeRagged     reference src/fuze.cpp:20-40
//   UTF8Validate   reference src/utf8_validate.cpp:18-137     (replace / drop malformed sequences)
// All are per-string state machines with data-dependent output sizes: one thread per string computes the output length, a
// cub scan turns the lengths into offsets, the same state machine runs again writing.  HBM-bound byte streams.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200tok {

// GPT-2 byte <-> unicode map (public algorithm of the GPT-2 encoder: printable Latin-1 bytes map to themselves, the other 68
// bytes to U+0100 + n in byte order).  code point of byte b:
__host__ __device__ inline uint32_t gpt2_byte_codepoint(uint32_t b, const uint16_t* shifted /* [256] or null */) {
    return shifted ? shifted[b] : b;
}
// Host-side table builder (tables.cpp): cp[b] for all b.
inline void gpt2_build_byte_codepoints(uint16_t* cp) {
    int n = 0;
    for (int b = 0; b < 256; ++b) {
        const bool keep = (b >= '!' && b <= '~') || (b >= 0xA1 && b <= 0xAC) || (b >= 0xAE && b <= 0xFF);
        cp[b] = keep ? (uint16_t)b : (uint16_t)(256 + n++);
    }
}

struct ByteCharTables {
    const uint16_t* cp;        // [256] code point of every byte (all < 0x800: one or two UTF-8 bytes)
    const uint8_t* pair_map;   // [4 * 64] byte of the 2-byte sequence (first - 194, second - 128); src/chars_to_bytes.cpp:20-29
};

// ---- BytesToChars: one thread per element ----
__global__ void b2c_len_kernel(const int32_t* begins, const int32_t* ends, const uint8_t* chars, const uint8_t* skips, int64_t n,
                               const uint16_t* cp, int32_t* len) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t b = begins[i], e = ends[i];
    int32_t l = e > b ? e - b : 0;
    if (!(skips && skips[i])) for (int32_t k = b; k < e; ++k) l += cp[chars[k]] >= 0x80;      // two bytes for code points >= 0x80
    len[i] = l;
}
__global__ void b2c_write_kernel(const int32_t* begins, const int32_t* ends, const uint8_t* chars, const uint8_t* skips, int64_t n,
                                 const uint16_t* cp, const int32_t* out_begins, const int32_t* len, int32_t* out_ends, uint8_t* out,
                                 int64_t cap, int64_t* total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t o = out_begins[i];
    const int64_t oe = o + len[i];
    out_ends[i] = (int32_t)oe;
    if (i == n - 1) *total = oe;
    if (oe > cap) return;
    const int32_t b = begins[i], e = ends[i];
    if (skips && skips[i]) { for (int32_t k = b; k < e; ++k) out[o++] = chars[k]; return; }
    for (int32_t k = b; k < e; ++k) {
        const uint32_t c = cp[chars[k]];
        if (c < 0x80) out[o++] = (uint8_t)c;
        else { out[o++] = (uint8_t)(0xC0 | (c >> 6)); out[o++] = (uint8_t)(0x80 | (c & 0x3F)); }
    }
}

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

// ---- UTF8Validate: the reference's byte state machine, with the output either counted or written ----
template <bool WRITE>
__device__ __forceinline__ int32_t utf8_validate_string(const uint8_t* bytes, int32_t b, int32_t e, bool replace, uint8_t* out) {
    const uint32_t starts[4] = {0x0, 0x80, 0x800, 0x10000};
    uint32_t cp = 0, to_consume = 0, num = 0;
    int32_t o = 0;
    auto put_repl = [&]() { if (WRITE) { out[o] = 0xEF; out[o + 1] = 0xBF; out[o + 2] = 0xBD; } o += 3; };
    for (int32_t j = b; j < e; ++j) {
        const uint8_t c = bytes[j];
        if (!to_consume) {
            if (c < 128) { if (WRITE) out[o] = c; ++o; }
            else if ((c >> 5) == 0b110) { num = 2; to_consume = 1; cp = (0b11111u & c) << 6; }
            else if ((c >> 4) == 0b1110) { num = 3; to_consume = 2; cp = (0b1111u & c) << 12; }
            else if ((c >> 3) == 0b11110) { num = 4; to_consume = 3; cp = (0b111u & c) << 18; }
            else if (replace) put_repl();
            continue;
        }
        if ((c >> 6) != 0b10) {      // broken continuation: it may still start a new symbol (:96-105)
            --j;
            to_consume = 0;
            if (replace) put_repl();
            continue;
        }
        --to_consume;
        cp |= (0b111111u & c) << (6 * to_consume);
        if (!to_consume) {
            if (cp < starts[num - 1]) {          // overlong (:112-122)
                if (replace) for (uint32_t k = 0; k < num; ++k) put_repl();
            } else {
                if (WRITE) for (uint32_t k = 0; k < num; ++k) out[o + k] = bytes[j + 1 - num + k];
                o += (int32_t)num;
                cp = 0;
            }
        }
    }
    if (replace && to_consume > 0) put_repl();   // unfinished sequence (:131-134)
    return o;
}
__global__ void utf8_len_kernel(const int32_t* begins, const int32_t* ends, const uint8_t* bytes, int64_t n, int replace, int32_t* len) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    len[i] = utf8_validate_string<false>(bytes, begins[i], ends[i], replace != 0, nullptr);
}
__global__ void utf8_write_kernel(const int32_t* begins, const int32_t* ends, const uint8_t* bytes, int64_t n, int replace, int32_t base,
                                  const int32_t* off, const int32_t* len, int32_t* out_begins, int32_t* out_ends, uint8_t* out, int64_t cap,
                                  int64_t* total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t o = (int64_t)base + off[i];     // the reference starts its output cursor at begins[0] (:50)
    out_begins[i] = (int32_t)o;
    out_ends[i] = (int32_t)(o + len[i]);
    if (i == n - 1) *total = o + len[i];
    if (o + len[i] > cap) return;
    utf8_validate_string<true>(bytes, begins[i], ends[i], replace != 0, out + o);
}

}  // namespace b200tok
