// kernels_norm.cuh — normalisers:
//   RegexNormalization     reference src/regex_normalization.cpp:127-153 (+ src/utils.cpp:315-382, pcre2_substitute), for the
//                          single-character search patterns the converter emits (tables.cpp kNormPatterns)
//   CharsMapNormalization  reference src/charsmap_normalization.cpp:34-69 (sentencepiece Normalizer over a precompiled charsmap)
//   both through evaluate_normalization_helper (src/utils.cpp:178-234): per element, skip-flagged elements pass through.
// The scan "at an active byte, consume c bytes and emit o bytes" is sequential per string.  One warp per string takes 32
// byte positions at a time: every lane evaluates the step that WOULD start at its byte (tok_core.cuh norm_eval), pointer
// jumping over the lanes' jump targets (5 rounds of shuffles) marks the positions the scan really visits, a warp scan of
// their output lengths places the bytes.  Two passes (lengths -> cub scan -> write), like the other byte-stream shims.
// HBM-bound byte streams; the tables (class stage tables, or the double array, 17-260 KB) stay in L1/L2.
#pragma once
#include "kernels.cuh"

namespace b200tok {

template <bool WRITE>
__global__ void __launch_bounds__(256) normalize_kernel(const __grid_constant__ NormRule R, const int32_t* __restrict__ begins,
                                                        const int32_t* __restrict__ ends, const uint8_t* __restrict__ chars,
                                                        const uint8_t* __restrict__ skips, int64_t n, int32_t* __restrict__ len,
                                                        const int32_t* __restrict__ off, int32_t base, int32_t* __restrict__ out_begins,
                                                        int32_t* __restrict__ out_ends, uint8_t* __restrict__ out, int64_t cap, int64_t* total) {
    // off = exclusive scan of len; the strings go to base + off[i] (UTF8Validate starts its cursor at begins[0]); out_begins is
    // written when given (the normalisers scan straight into it and pass nullptr)
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n; i += nwarps) {
        const int b = begins[i], e = ends[i];
        int64_t o0 = 0;
        if (WRITE) {
            o0 = (int64_t)base + off[i];
            const int64_t oe = o0 + len[i];
            if (lane == 0) { if (out_begins) out_begins[i] = (int32_t)o0; out_ends[i] = (int32_t)oe; if (i == n - 1) *total = oe; }
            if (oe > cap) continue;
        }
        if (skips && skips[i]) {          // src/utils.cpp:211: the string is copied unchanged
            if (WRITE) { for (int k = b + lane; k < e; k += 32) out[o0 + (k - b)] = chars[k]; }
            else if (lane == 0) len[i] = e > b ? e - b : 0;
            continue;
        }
        int cur = b;                      // next byte the scan visits
        int o = 0;                        // bytes produced so far
        bool done = false;                // a non-global rule has replaced its match
        for (int c0 = b; c0 < e; c0 += 32) {
            if (cur >= c0 + 32) continue;
            const int pos = c0 + lane;
            const bool valid = pos < e && pos >= cur;
            const uint32_t byte = pos < e ? chars[pos] : 0u;
            NormStep st;
            st.consumed = 32; st.olen = 0; st.src = -2; st.matched = 0;
            // ---- a chunk of ASCII bytes: every byte is a step of its own, straight from the per-byte tables ----
            bool fast = __ballot_sync(FULL, byte >= 0x80u) == 0u;
            uint32_t mapped = byte;
            int folen = 1;                    // output bytes of this lane's byte on the fast path
            if (fast) {
                if (R.kind == NORM_B2C) {
                    folen = reinterpret_cast<const uint16_t*>(R.normalized)[byte] >= 0x80 ? 2 : 1;
                    st.matched = valid && folen == 2;
                } else if (R.kind == NORM_UTF8 || R.kind == NORM_C2B) {
                    st.matched = 0;               // ASCII is copied
                } else {
                    const uint32_t fl = __ldg(R.atab + 128 + byte);
                    if (R.kind == NORM_CHARSMAP) {
                        mapped = __ldg(R.atab + byte);
                        bool slow = (fl & (NA_COMPLEX | NA_ASCII_KIDS)) != 0;
                        if (lane == 31 && (fl & NA_OTHER_KIDS) && pos + 1 < e && chars[pos + 1] >= 0x80u) slow = true;   // a rule may run into the next chunk
                        fast = __ballot_sync(FULL, valid && slow) == 0u;
                    } else {
                        const bool in_class = R.any || (R.literal_cp >= 0 ? (int32_t)byte == R.literal_cp : (fl & R.mask) != 0);
                        st.matched = valid && in_class != (R.negate != 0) && (!R.anchored || pos == b) && (R.global || !done);
                        folen = (int)R.pre_len + (R.keep ? 1 : 0) + (int)R.post_len;
                    }
                }
            }
            if (fast) {
                uint32_t hits = (R.kind == NORM_CLASS || R.kind == NORM_B2C) ? __ballot_sync(FULL, st.matched) : 0u;
                if (hits && !R.global && R.kind == NORM_CLASS) {                 // only the first match of the string is replaced
                    done = true;
                    st.matched = st.matched && lane == __ffs(hits) - 1;
                    hits &= 0u - hits;
                }
                const uint32_t vmask = __ballot_sync(FULL, valid);
                if (!hits) {                             // nothing changes length: byte k goes to slot k
                    if (WRITE && valid) out[o0 + o + __popc(vmask & ((1u << lane) - 1u))] = (uint8_t)mapped;
                    o += __popc(vmask);
                } else {
                    st.consumed = 1; st.src = R.kind == NORM_B2C ? -4 : st.matched ? -1 : -2;
                    st.olen = st.matched ? folen : 1;
                    int incl = valid ? st.olen : 0;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += v; }
                    if (WRITE && valid) norm_emit(R, st, chars, pos, out + o0 + o + (incl - st.olen));
                    o += __shfl_sync(FULL, incl, 31);
                }
                cur = c0 + 32;
                continue;
            }
            st.matched = 0;
            if (valid) st = norm_eval(R, chars, b, pos, e, done);
            // which lanes does the scan visit?  J = where the scan goes after this lane (>= 32: leaves the chunk),
            // M = lanes visited from here; doubling composes them.
            int J = lane + st.consumed;
            uint32_t M = 1u << lane;
#pragma unroll
            for (int r = 0; r < 5; ++r) {
                const int Jj = __shfl_sync(FULL, J, J & 31);
                const uint32_t Mj = __shfl_sync(FULL, M, J & 31);
                if (J < 32) { M |= Mj; J = Jj; }
            }
            const int s0 = cur - c0;
            const uint32_t visited = __shfl_sync(FULL, M, s0);
            cur = c0 + __shfl_sync(FULL, J, s0);
            bool active = ((visited >> lane) & 1u) && pos < e;
            if (!R.global && R.kind == NORM_CLASS && !done) {       // only the first match of the string is replaced
                const uint32_t hits = __ballot_sync(FULL, active && st.matched);
                if (hits) {
                    done = true;
                    if (active && st.matched && lane != __ffs(hits) - 1) { st.matched = 0; st.src = -2; st.olen = st.consumed; }
                }
            }
            int incl = active ? st.olen : 0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += v; }
            if (WRITE && active) norm_emit(R, st, chars, pos, out + o0 + o + (incl - st.olen));
            o += __shfl_sync(FULL, incl, 31);
        }
        if (!WRITE && lane == 0) len[i] = o;
    }
}

// Short elements (pieces after a splitter, the per-token strings of a detokenizer: a handful of bytes each): one thread per
// element runs the sequential scan itself (tok_core.cuh norm_string) — a warp per 5-byte string would idle 27 lanes.
template <bool WRITE>
__global__ void __launch_bounds__(256) normalize_short_kernel(const __grid_constant__ NormRule R, const int32_t* __restrict__ begins,
                                                              const int32_t* __restrict__ ends, const uint8_t* __restrict__ chars,
                                                              const uint8_t* __restrict__ skips, int64_t n, int32_t* __restrict__ len,
                                                              const int32_t* __restrict__ off, int32_t base, int32_t* __restrict__ out_begins,
                                                              int32_t* __restrict__ out_ends, uint8_t* __restrict__ out, int64_t cap, int64_t* total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = begins[i], e = ends[i];
    const bool skip = skips && skips[i];
    if (!WRITE) { len[i] = skip ? (e > b ? e - b : 0) : norm_string(R, chars, b, e, nullptr); return; }
    const int64_t o0 = (int64_t)base + off[i], oe = o0 + len[i];
    if (out_begins) out_begins[i] = (int32_t)o0;
    out_ends[i] = (int32_t)oe;
    if (i == n - 1) *total = oe;
    if (oe > cap) return;
    if (skip) { for (int k = b; k < e; ++k) out[o0 + (k - b)] = chars[k]; }
    else norm_string(R, chars, b, e, out + o0);
}

// Size of a result: offset of the last string + its length.
__global__ void normalize_total_kernel(const int32_t* off, const int32_t* len, int64_t n, int64_t* total) { *total = (int64_t)off[n - 1] + len[n - 1]; }

// ---- a chain of normalisers on all-ASCII strings: one composed byte table ----
// T[a] = what byte a becomes after every op of the chain (api.cu compose_chain): a byte, NT_DEL (dropped) or NT_GENERAL
// (some op does more than map / drop this byte: the string takes the op-by-op path).  A string made only of bytes with a
// simple fate is normalised in ONE pass pair whatever the number of ops; the others ("general") are gathered into a
// sub-list, run op by op, and their results are copied into place by the write pass.

template <bool WRITE>
__global__ void __launch_bounds__(256) compose_kernel(const uint8_t* __restrict__ table, const int32_t* __restrict__ begins,
                                                      const int32_t* __restrict__ ends, const uint8_t* __restrict__ chars,
                                                      const uint8_t* __restrict__ skips, int64_t n, int32_t* __restrict__ len,
                                                      int32_t* __restrict__ general, const int32_t* __restrict__ out_begins,
                                                      int32_t* __restrict__ out_ends, uint8_t* __restrict__ out, const int32_t* __restrict__ sub_index,
                                                      const int32_t* __restrict__ sub_begins, const uint8_t* __restrict__ sub_chars, int64_t* total) {
    __shared__ uint8_t T[256];
    T[threadIdx.x] = threadIdx.x < 128 ? table[threadIdx.x] : (uint8_t)NT_GENERAL;      // non-ASCII bytes are never simple
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n; i += nwarps) {
        const int b = begins[i], e = ends[i];
        const bool skip = skips && skips[i];
        if (!WRITE) {
            int cnt = 0;
            bool gen = false;
            if (skip) cnt = e > b ? e - b : 0;
            else {
                for (int c0 = b; c0 < e && !gen; c0 += 128) {           // 4 chunks in flight
                    uint32_t t[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) { const int pos = c0 + 32 * u + lane; t[u] = pos < e ? T[chars[pos]] : (uint32_t)NT_DEL; }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        gen = gen || __any_sync(FULL, t[u] == NT_GENERAL);
                        cnt += __popc(__ballot_sync(FULL, t[u] < NT_DEL));
                    }
                }
            }
            if (lane == 0) { len[i] = gen ? 0 : cnt; general[i] = gen ? 1 : 0; }
            continue;
        }
        const int64_t o0 = out_begins[i];
        const int l = len[i];
        if (lane == 0) { out_ends[i] = (int32_t)(o0 + l); if (i == n - 1) *total = o0 + l; }
        if (general[i]) {                 // normalised op by op: copy the result into place
            const uint8_t* src = sub_chars + sub_begins[sub_index[i]];
            for (int k = lane; k < l; k += 32) out[o0 + k] = src[k];
        } else if (skip) {
            for (int k = lane; k < l; k += 32) out[o0 + k] = chars[b + k];
        } else {
            int o = 0;
            for (int c0 = b; c0 < e; c0 += 128) {
                uint32_t t[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { const int pos = c0 + 32 * u + lane; t[u] = pos < e ? T[chars[pos]] : (uint32_t)NT_DEL; }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t m = __ballot_sync(FULL, t[u] < NT_DEL);
                    if (t[u] < NT_DEL) out[o0 + o + __popc(m & ((1u << lane) - 1u))] = (uint8_t)t[u];
                    o += __popc(m);
                }
            }
        }
    }
}

// The general strings as a list of their own (their extents still point into the caller's chars).
__global__ void gather_general_kernel(const int32_t* general, const int32_t* sub_index, const int32_t* begins, const int32_t* ends, int64_t n,
                                      int32_t* sub_begins, int32_t* sub_ends) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !general[i]) return;
    sub_begins[sub_index[i]] = begins[i];
    sub_ends[sub_index[i]] = ends[i];
}
__global__ void merge_general_len_kernel(const int32_t* general, const int32_t* sub_index, const int32_t* sub_begins, const int32_t* sub_ends, int64_t n, int32_t* len) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !general[i]) return;
    len[i] = sub_ends[sub_index[i]] - sub_begins[sub_index[i]];
}

}  // namespace b200tok
