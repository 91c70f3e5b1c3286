// kernels_misc.cuh — VocabEncoder lookup, VocabDecoder gather (+ ByteFallback), ByteFallback.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tables.hpp"

namespace b200tok {

// ---- VocabEncoder: exact string -> value, miss -> default (src/vocab_encoder.cpp:88-91) --------
template <typename T>
__global__ void vocab_lookup_kernel(const int32_t* begins, const int32_t* ends, const uint8_t* chars, int64_t n,
                                    const VocabEncSlot* slots, uint32_t mask, const uint8_t* key_bytes,
                                    int64_t default_value, T* out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t b = begins[i], len = ends[i] - b;
    const uint8_t* s = chars + b;
    uint64_t h = 1469598103934665603ull;
    for (int k = 0; k < len; ++k) { h ^= s[k]; h *= 1099511628211ull; }
    uint32_t k = (uint32_t)h & mask;
    int64_t v = default_value;
    for (;;) {
        const VocabEncSlot sl = slots[k];
        if (sl.len < 0) break;
        if (sl.hash == h && sl.len == len) {
            bool eq = true;
            for (int t = 0; t < len; ++t) if (key_bytes[sl.begin + t] != s[t]) { eq = false; break; }
            if (eq) { v = sl.value; break; }
        }
        k = (k + 1) & mask;
    }
    out[i] = (T)v;
}

// ---- ByteFallback test (src/byte_fallback.cpp:37-40 + sentencepiece PieceToByte) -------------
// Returns -1 if the token is copied verbatim, else the single output byte.
__host__ __device__ inline int byte_fallback_value(const uint8_t* t, int len) {
    if (len != 6 || t[0] != '<' || t[5] != '>') return -1;
    for (int k = 1; k < 6; ++k) if (t[k] == '<') return -1;          // rfind("<") == 0
    // rfind(">") == 5 holds because t[5] == '>'
    auto hex = [](uint8_t c) { return (c >= '0' && c <= '9') ? c - '0' : (c >= 'A' && c <= 'F') ? c - 'A' + 10 : -1; };
    if (t[1] != '0' || t[2] != 'x') return 0xFF;                     // PieceToByte -> -1 -> stored as uint8_t
    const int hi = hex(t[3]), lo = hex(t[4]);
    return (hi < 0 || lo < 0) ? 0xFF : hi * 16 + lo;
}

// ---- VocabDecoder pass 1: output length of every token position (src/vocab_decoder.cpp:67-81) --
__global__ void decode_len_kernel(const int32_t* ids, int64_t n, const int32_t* vb, const int32_t* ve, int64_t V,
                                  const int32_t* skip, int32_t n_skip, const int16_t* bf_byte, int32_t* len_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t id = ids[i];
    int32_t len = 0;
    if ((uint64_t)(int64_t)id < (uint64_t)V) {   // negatives fail the unsigned compare, as in the reference
        bool skipped = false;
        for (int k = 0; k < n_skip; ++k) if (skip[k] == id) { skipped = true; break; }
        if (!skipped) len = (bf_byte && bf_byte[id] >= 0) ? 1 : ve[id] - vb[id];
    }
    len_out[i] = len;
}

// pass 2: copy the bytes; begins[] is the exclusive scan of len[]
__global__ void decode_copy_kernel(const int32_t* ids, int64_t n, const int32_t* vb, const uint8_t* vc,
                                   const int16_t* bf_byte, const int32_t* len, const int32_t* begins, int32_t* ends,
                                   uint8_t* out_chars, int64_t cap, int32_t* status, int64_t* total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t l = len[i], o = begins[i];
    ends[i] = o + l;
    if (i == n - 1) { if (total) *total = (int64_t)o + l; }
    if (l == 0) return;
    if ((int64_t)o + l > cap) { atomicOr(&status[0], 1); return; }
    const int32_t id = ids[i];
    if (bf_byte && bf_byte[id] >= 0) { out_chars[o] = (uint8_t)bf_byte[id]; return; }
    const uint8_t* s = vc + vb[id];
    for (int k = 0; k < l; ++k) out_chars[o + k] = s[k];
}

__global__ void decode_ragged_kernel(int64_t batch, int64_t width, int32_t* rb, int32_t* re) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    rb[b] = (int32_t)(b * width);
    re[b] = (int32_t)(b * width + width);
}

// ---- ByteFallback stand-alone -------------------------------------------------------------
__global__ void bytefallback_len_kernel(const int32_t* begins, const int32_t* ends, const uint8_t* chars, int64_t n,
                                        int32_t* len_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t l = ends[i] - begins[i];
    len_out[i] = byte_fallback_value(chars + begins[i], l) >= 0 ? 1 : l;
}
__global__ void bytefallback_copy_kernel(const int32_t* begins, const int32_t* ends, const uint8_t* chars, int64_t n,
                                         const int32_t* out_begins, int32_t* out_ends, uint8_t* out_chars,
                                         int64_t* total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t l = ends[i] - begins[i], o = out_begins[i];
    const uint8_t* s = chars + begins[i];
    const int v = byte_fallback_value(s, l);
    int32_t ol;
    if (v >= 0) { out_chars[o] = (uint8_t)v; ol = 1; }
    else { for (int k = 0; k < l; ++k) out_chars[o + k] = s[k]; ol = l; }
    out_ends[i] = o + ol;
    if (i == n - 1 && total) *total = (int64_t)o + ol;
}

}  // namespace b200tok
