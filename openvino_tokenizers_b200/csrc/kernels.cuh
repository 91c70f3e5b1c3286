// kernels.cuh — sm_100a kernels of the tokenizer hot path.
//
// Work decomposition: one warp per row (input string); the row's bytes are staged
// through shared memory in 512-byte windows; the split pattern is evaluated for all byte positions
// of a window in parallel ("what would match if a match started here"), the reference's sequential
// match loop (src/regex_split.cpp:287-309) is then resolved as a pointer chain; the resulting
// pieces are handed round-robin to the 32 lanes, each running the BPE merge loop / WordPiece trie
// walk for its piece with all state in shared memory and the merge-rank hash / tries read through
// the read-only path from L2-resident HBM tables.  Token ids go to a row-local slot of a worst-case
// buffer and are compacted by a scan + copy pass.  Integer/byte work only: no tensor cores.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tok_core.cuh"

namespace b200tok {

constexpr int WIN = 512;            // fresh bytes per window
constexpr int LA = 16;              // look-ahead bytes staged beyond the window
constexpr int WBYTES = WIN + LA + 16;
constexpr int LBK = 16;             // look-back bytes staged before the window (closed-form splitters)
constexpr int NWORDS = (WIN + LA + 31) / 32 + 1;
constexpr int WARPS_PER_BLOCK = 8;
constexpr int BLOCK_THREADS = WARPS_PER_BLOCK * 32;

constexpr int kRowsSmemFixed = 128 + 1024;   // per-CTA tables in front of the per-warp state

constexpr uint16_t F_MATCH = 0x0400, F_DROP = 0x0800, F_UNC = 0x8000, POS_MASK = 0x03FF;

enum : int { OP_BPE = 0, OP_WORDPIECE = 1, OP_SPLIT = 2, OP_SPECIAL = 3 };

// status words (device int32 array)
enum : int { ST_ERROR = 0, ST_NGIANT = 1, ST_TICKET = 2, ST_TOTAL = 3, ST_BASE = 4, ST_POOL_NEED_HI = 5, ST_NREDO = 6, ST_TICKET2 = 7, ST_MINREDO = 8, ST_TICKET3 = 9,
             ST_ALLOC = 32,        // the slot allocator: in its own 128-byte line — it and the ticket counter are both hit once per row
             ST_WORDS = 64 };      // (a multiple of 32 words: consecutive status blocks keep the two counters in separate lines)
enum : int { ERR_TMP_OVERFLOW = 1, ERR_GIANT_LIST = 2, ERR_GIANT_POOL = 4, ERR_HEAP_TIE = 8 };

struct GiantItem { int32_t row, begin, end, slot; };

// Result buffers of every rank, mapped into this device (sharded output); world == 0: not sharded.
struct PeerOut {
    int32_t* ids[8]; int32_t* begins[8]; int32_t* ends[8];
    int32_t world, rank; int64_t slot_capacity, rows_per_rank;
    uint16_t* ids16[8]; int32_t wire16;      // 16-bit wire format: ids go to the peers' u16 staging buffers, widened locally afterwards
    int32_t* ids_mc; int32_t* begins_mc; int32_t* ends_mc;   // NVLS multicast mappings (one store reaches every rank), or null
};
// One store through the NVSwitch multicast address: written into every rank's copy of the buffer.
__device__ __forceinline__ void mc_store(int32_t* p, int32_t v) {
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" :: "l"(p), "f"(__int_as_float(v)) : "memory");
}
__device__ __forceinline__ void mc_store4(int32_t* p, uint4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)),
                 "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w)) : "memory");
}
__device__ __forceinline__ void peer_store(const PeerOut& Q, int p, int64_t o, int32_t v) {
    if (Q.wire16) Q.ids16[p][o] = (uint16_t)v; else Q.ids[p][o] = v;
}

// In-order single-pass emit (fast path): every row obtains its output offset — the exclusive prefix of the token counts of
// all earlier rows — by decoupled look-back over one 64-bit descriptor per row, so the compact (begins, ends, ids) result
// is written exactly once, by the kernel that produced the ids (replaces the reference's serial `ragged_offset++`
// bookkeeping, src/bpe_tokenizer.cpp:146-162, and this library's former capacity scan + slot buffer + compaction pass).
//   descriptor = epoch[63:34] | flag[33:32] | value[31:0];   flag 1 = the row's own count, 2 = inclusive prefix
// Rows take their tickets in row order, so every predecessor of a row is already running: the look-back cannot deadlock.
// The epoch makes descriptors of earlier launches read as "not yet published": the array is never cleared between calls.
struct OrderedOut {
    unsigned long long* desc;     // [n_rows]; nullptr = the slot-buffer path
    uint32_t epoch;               // this launch (30 bits, never 0)
    int32_t* ids; int32_t* begins; int32_t* ends;
    int64_t cap;                  // capacity of ids
    int64_t* total;               // optional device-side total
    void* stage;                  // per-warp staging (global, L2-resident) for rows that span several windows: stage_cap IdT per warp
    int32_t stage_cap;
};
constexpr unsigned long long kDescAgg = 1ull << 32, kDescPrefix = 2ull << 32;
__device__ __forceinline__ unsigned long long desc_load(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void desc_store(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
// Publish a row's own count as soon as it is known (row == first: its inclusive prefix right away).  Lane 0 only.
__device__ __forceinline__ void lookback_publish_count(unsigned long long* desc, uint32_t epoch, int first, int row, uint32_t count, long long seed) {
    const unsigned long long tag = (unsigned long long)epoch << 34;
    desc_store(desc + row, row == first ? (tag | kDescPrefix | (uint32_t)(seed + count)) : (tag | kDescAgg | count));
}
// One look-back step for `row` (> first): examines up to 32 predecessors below `idx` (initially row - 1), adds what has been
// published to `acc`, moves `idx` down.  Returns 1 once a predecessor's inclusive prefix (or the first row) has been reached:
// acc is then the exclusive prefix of `row`; 0 = all 32 had published their counts, keep walking; -1 = stopped at a predecessor
// that has not published yet (try again later: the progress made is kept in idx / acc).  Warp-uniform.
__device__ __forceinline__ int lookback_poll(const unsigned long long* desc, uint32_t epoch, int first, int& idx, long long& acc, int lane) {
    // four chunks of 32 descriptors are fetched at once: a walk is a chain of L2 round trips, and this makes each trip cover 128 rows
    constexpr int U = 4;
    unsigned long long d[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int j = idx - 32 * u - lane;
        d[u] = j >= first ? desc_load(desc + j) : (((unsigned long long)epoch << 34) | kDescPrefix);   // below the first row: prefix 0
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const uint32_t fl = (uint32_t)(d[u] >> 32) & 3u;
        const bool ok = (d[u] >> 34) == epoch && fl != 0u;
        const uint32_t okm = __ballot_sync(0xFFFFFFFFu, ok);
        const int run = okm == 0xFFFFFFFFu ? 32 : __ffs(~okm) - 1;             // predecessors idx, idx-1, ... published without a gap
        if (run == 0) return -1;
        const uint32_t runm = run == 32 ? 0xFFFFFFFFu : ((1u << run) - 1u);
        const uint32_t pm = __ballot_sync(0xFFFFFFFFu, fl == 2u) & runm;
        const int k = pm ? __ffs(pm) - 1 : run - 1;                            // last lane that contributes
        acc += __reduce_add_sync(0xFFFFFFFFu, lane <= k ? (uint32_t)d[u] : 0u);   // (counts and prefixes fit 31 bits: int32 offsets)
        idx -= k + 1;
        if (pm) return 1;
        if (run < 32) return -1;
    }
    return 0;
}
// Blocking form: exclusive prefix of `count` over rows [first, row) (+ `seed` for row == first); publishes both descriptors.
__device__ __forceinline__ long long lookback_exclusive(unsigned long long* desc, uint32_t epoch, int first, int row, uint32_t count, long long seed, int lane) {
    if (lane == 0) lookback_publish_count(desc, epoch, first, row, count, seed);
    if (row == first) return seed;
    int idx = row - 1;
    long long acc = 0;
    for (int r; (r = lookback_poll(desc, epoch, first, idx, acc, lane)) != 1;)
        if (r < 0) __nanosleep(128);
    if (lane == 0) desc_store(desc + row, ((unsigned long long)epoch << 34) | kDescPrefix | (uint32_t)(acc + count));
    return acc;
}

struct RowParams {
    // input ragged strings (device)
    const int32_t* rb; const int32_t* re; int32_t n_rows;
    const int32_t* begins; const int32_t* ends;
    const uint8_t* chars; int32_t n_chars;
    const uint8_t* skips;
    // splitter
    SplitSpec spec; int repeat; int mode; int invert; int max_splits;
    ClassTables cls;
    // tokenizer tables
    BpeTables bpe; int32_t suffix_len;
    WordpieceTables wp; int32_t unk_id;
    // row-local worst-case output slots
    const int32_t* row_base;   // [n_rows] exclusive scan of row capacities
    int32_t* row_ext;          // [n_rows] slots used (tokens incl. reserved holes / pieces)
    int32_t* row_cnt;          // [n_rows] true count
    uint8_t* row_flag;         // [n_rows] 1 = slot range contains holes (-1)
    int32_t* tmp_a;            // tokens, or piece begins
    int32_t* tmp_b;            // piece ends
    uint8_t* tmp_c;            // piece skip flags
    int64_t tmp_cap;
    // giants
    GiantItem* giants; int32_t giants_cap;
    int32_t* status;
    int32_t dbg_flags;         // development switches (env B200TOK_DEBUG_FLAGS), 0 in production
    // contiguous, increasing elements (verified by the host): a row's slot base follows from its first element's byte
    // offset, so the capacity kernel + scan are skipped:  base = begins[rb[row]] - direct_byte0 + (rb[row] - direct_elem0) * extra
    int32_t direct_base; int32_t direct_byte0; int32_t direct_elem0; int32_t direct_extra;
    // fast kernel: a row takes its slot range from a bump allocator (status[ST_ALLOC]) when it starts — no capacity pass, no scan
    int32_t alloc_base;
    // slot path: 1 = software-pipelined input (the next window's bulk copy is issued while the current one is tokenised)
    int32_t prefetch;
    // list mode: process rows row_list[0 .. status[ST_NREDO]) (rows the fast kernel handed back) instead of [0, n_rows)
    const int32_t* row_list;
    // sharded fast path: the emit step stores ids (and row extents) straight into every rank's buffers, rows stay at their
    // worst-case positions inside this rank's slot (a ragged tensor may have gaps), so no scan / compaction pass follows
    PeerOut peer;
    OrderedOut oo;
};

struct __align__(16) WarpSmem {
    uint8_t raw_bytes[LBK + WBYTES];
    __device__ __forceinline__ uint8_t* B() { return raw_bytes + LBK; }   // index 0 = window position
    uint16_t seg[WIN + 4];          // segment list: start | flags, plus an end sentinel
    uint16_t act[WIN / 3 + 8];      // indices of the segments (>= 3 symbols) that still have a mergeable pair
    uint32_t segbits[NWORDS];       // bit per position: a segment starts here
    uint32_t actbits[NWORDS];       // bit per position: the pair (w, w+1) is mergeable
    union {
        struct {
            uint8_t raw_cls[LBK + WBYTES];
            __device__ __forceinline__ uint8_t* K() { return raw_cls + LBK; }
            uint32_t bnd[NWORDS], cs[NWORDS], nl[NWORDS], mm[NWORDS], chain[NWORDS];
            uint16_t entry[32];
            uint16_t nxt[WIN + LA + 8];
            uint16_t exitp[WIN + LA + 8];
        } sp;
        struct {
            int32_t ids[WIN];
            uint32_t key[WIN];
        } bp;
    } u;
};

__device__ __forceinline__ int next_bit(const uint32_t* words, int i, int limit) {
    int wd = (i + 1) >> 5;
    if (wd >= NWORDS) return limit;
    uint32_t m = words[wd] & (0xFFFFFFFFu << ((i + 1) & 31));
    while (!m) {
        if (++wd >= NWORDS) return limit;
        m = words[wd];
    }
    const int r = wd * 32 + __ffs(m) - 1;
    return r < limit ? r : limit;
}
// highest set bit in [lo, e), or -1
__device__ __forceinline__ int prev_bit(const uint32_t* words, int e, int lo) {
    int j = e - 1;
    if (j < lo) return -1;
    int wd = j >> 5;
    uint32_t m = words[wd] & (0xFFFFFFFFu >> (31 - (j & 31)));
    while (!m) {
        if (--wd < 0) return -1;
        m = words[wd];
    }
    const int r = wd * 32 + 31 - __clz(m);
    return r >= lo ? r : -1;
}

// Window context: same interface as ScanCtx (tok_core.cuh), answers from precomputed bitmasks.
struct WinCtx {
    const uint8_t* b; const uint8_t* k;
    const uint32_t* bnd; const uint32_t* cs; const uint32_t* nl;
    int hi;   // known end (window-relative)
    int lm;   // readable limit
    __device__ __forceinline__ int known() const { return hi; }
    __device__ __forceinline__ int lim() const { return lm; }
    __device__ __forceinline__ uint8_t byte(int i) const { return b[i]; }
    __device__ __forceinline__ uint8_t cls(int i) const { return k[i]; }
    __device__ __forceinline__ int next(int i) const { return next_bit(cs, i, lm); }
    __device__ __forceinline__ int run_end(int i) const { return next_bit(bnd, i, hi); }
    __device__ __forceinline__ int last_nl(int i) const { return prev_bit(nl, run_end(i), i); }
};

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

__device__ __forceinline__ bool seg_kept(uint16_t sg, int pat, int mode, int invert) {
    if (pat == PAT_BERT_FUSED) return !(sg & F_DROP);
    if (mode == SPLIT_ISOLATED) return true;
    const bool flag = (sg & F_MATCH) ? !invert : (invert != 0);   // the `invert` argument of add_split
    return !flag;                                                  // REMOVED drops flagged splits
}

// ------------------------------------------------------------------------------------------
// Split phase for one window: fills S.seg[0..ns) (+ sentinel seg[ns]) with the complete segments
// of bytes [0, wlen) and returns ns; `advance` = how far the window position moves (0 => the first
// segment does not fit: "giant").  end_rel = element end relative to the window start.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int split_window(WarpSmem& S, const RowParams& P, const uint8_t* ascii_smem, int lane,
                                            int wlen, int end_rel, int nload, int& advance) {
    auto& sp = S.u.sp;
    ClassTables T = P.cls;
    T.ascii = ascii_smem;
    // pass A: class per byte (continuation bytes copy their owner's class, plus C_CONT)
    for (int w = lane; w < nload; w += 32) {
        const uint8_t b = S.B()[w];
        uint8_t k;
        if (b < 0x80) k = ascii_smem[b];
        else if (is_cont_byte(b) && w > 0) {   // (index 0 is always treated as a character start)
            int j = w - 1;
            while (j >= 0 && j > w - 4 && is_cont_byte(S.B()[j])) --j;
            k = C_CONT;
            if (j >= 0 && j > w - 4 && S.B()[j] >= 0xC0) k |= char_class(S.B(), j, end_rel, T);
        } else k = char_class(S.B(), w, end_rel, T);
        sp.K()[w] = k;
    }
    __syncwarp();
    // pass B: run-boundary / char-start / newline bitmasks
    for (int it = 0; it < NWORDS; ++it) {
        const int w = it * 32 + lane;
        const bool valid = w < nload;
        const uint8_t k = valid ? sp.K()[w] : 0;
        const bool start = valid && !(k & C_CONT);
        bool bnd = false;
        if (start) bnd = (w == 0) || kind_of(k, P.spec.pat, P.spec.class_mask) != kind_of(sp.K()[w - 1], P.spec.pat, P.spec.class_mask);
        const uint32_t mb = __ballot_sync(0xFFFFFFFFu, bnd);
        const uint32_t mc = __ballot_sync(0xFFFFFFFFu, start);
        const uint32_t mn = __ballot_sync(0xFFFFFFFFu, valid && (k & C_NL) && !(k & C_CONT));
        if (lane == 0) { sp.bnd[it] = mb; sp.cs[it] = mc; sp.nl[it] = mn; }
    }
    __syncwarp();
    // pass C: the match that would start at every character position
    WinCtx ctx{S.B(), sp.K(), sp.bnd, sp.cs, sp.nl, wlen, nload};
    const bool hi_is_end = (wlen == end_rel);
    for (int it = 0; it * 32 < wlen; ++it) {
        const int w = it * 32 + lane;
        bool is_m = false;
        if (w < wlen && !(sp.K()[w] & C_CONT)) {
            const Match m = match_rep(ctx, P.spec, P.repeat != 0, w, end_rel);
            const bool unc = !hi_is_end && m.peek > wlen;
            int step = m.len > 0 ? w + m.len : ctx.next(w);
            if (step > wlen + LA) step = wlen + LA;
            is_m = m.len > 0;
            sp.nxt[w] = (uint16_t)(step | (is_m ? F_MATCH : 0) | (m.drop ? F_DROP : 0) | (unc ? F_UNC : 0));
        }
        const uint32_t mmask = __ballot_sync(0xFFFFFFFFu, is_m);
        if (lane == 0) sp.mm[it] = mmask;
    }
    for (int it = (wlen + 31) / 32 + lane; it < NWORDS; it += 32) sp.mm[it] = 0;
    __syncwarp();
    // pass D1: per 16-byte block, where does a chain entering at q leave the block
    {
        const int b0 = lane * 16, b1 = b0 + 16;
        for (int q = 15; q >= 0; --q) {
            const int w = b0 + q;
            if (w < wlen && ((sp.cs[w >> 5] >> (w & 31)) & 1u)) {
                const uint16_t v = sp.nxt[w];
                const int tgt = v & POS_MASK;
                uint16_t ex;
                if (v & F_UNC) ex = (uint16_t)(w | F_UNC);
                else if (tgt >= b1 || tgt >= wlen) ex = (uint16_t)tgt;
                else ex = sp.exitp[tgt];
                sp.exitp[w] = ex;
            }
        }
        sp.entry[lane] = 0xFFFF;
        reinterpret_cast<uint16_t*>(sp.chain)[lane] = 0;
        if (lane < 2 * NWORDS - 32) reinterpret_cast<uint16_t*>(sp.chain)[32 + lane] = 0;
    }
    __syncwarp();
    // pass D2: lane 0 hops block to block
    int stop = 0, unc = 0;
    if (lane == 0) {
        int cur = 0;
        while (cur < wlen) {
            sp.entry[cur >> 4] = (uint16_t)cur;
            const uint16_t ex = sp.exitp[cur];
            if (ex & F_UNC) { unc = 1; cur = ex & POS_MASK; break; }
            cur = ex;
        }
        stop = cur;
    }
    stop = __shfl_sync(0xFFFFFFFFu, stop, 0);
    unc = __shfl_sync(0xFFFFFFFFu, unc, 0);
    __syncwarp();
    // pass D3: every lane marks the chain positions inside its block
    {
        const int b0 = lane * 16, b1 = b0 + 16;
        uint32_t bits = 0;
        int cur = sp.entry[lane];
        if (cur != 0xFFFF) {
            while (cur < b1 && cur < stop) {
                bits |= 1u << (cur - b0);
                const uint16_t v = sp.nxt[cur];
                if (v & F_UNC) break;
                cur = v & POS_MASK;
            }
        }
        reinterpret_cast<uint16_t*>(sp.chain)[lane] = (uint16_t)bits;
    }
    __syncwarp();
    // pass E: segment starts = matches, and gap characters not preceded by a gap character
    int ns = 0;
    for (int it = 0; it * 32 < wlen; ++it) {
        const int w = it * 32 + lane;
        bool st = false;
        uint16_t sg = 0;
        if (w < wlen && w < stop && ((sp.chain[w >> 5] >> (w & 31)) & 1u)) {
            const bool is_m = (sp.mm[w >> 5] >> (w & 31)) & 1u;
            if (is_m || w == 0) st = true;
            else {
                const int pw = prev_bit(sp.cs, w, 0);
                const bool prev_gap = pw >= 0 && ((sp.chain[pw >> 5] >> (pw & 31)) & 1u) && !((sp.mm[pw >> 5] >> (pw & 31)) & 1u);
                st = !prev_gap;
            }
            sg = (uint16_t)(w | (sp.nxt[w] & (F_MATCH | F_DROP)));
        }
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, st);
        if (st) S.seg[ns + __popc(m & ((1u << lane) - 1u))] = sg;
        if (lane == 0) S.segbits[it] = m;
        ns += __popc(m);
    }
    for (int it = (wlen + 31) / 32 + lane; it < NWORDS; it += 32) S.segbits[it] = 0;
    __syncwarp();
    // completeness
    if (hi_is_end && !unc) {
        advance = wlen;
        if (lane == 0) S.seg[ns] = (uint16_t)wlen;
    } else {
        int adv = stop;
        if (ns > 0) {
            const uint16_t last = S.seg[ns - 1];
            if (!(last & F_MATCH)) { --ns; adv = last & POS_MASK; }   // trailing gap may continue: redo from its start
        }
        advance = adv;
        __syncwarp();                               // every lane has read S.seg[ns - 1] before lane 0 reuses that slot (racecheck)
        if (lane == 0) S.seg[ns] = (uint16_t)adv;
    }
    __syncwarp();
    return ns;
}

// Split phase for the GPT-2 byte-level patterns in isolate mode: the closed-form per-position predicate
// (tok_core.cuh gpt2_piece_starts_at) replaces match evaluation + chain resolution.  `lb` look-back bytes of the
// same element are staged before the window so that every position sees its true left context.
__device__ __forceinline__ int split_window_gpt2(WarpSmem& S, const RowParams& P, const uint8_t* ascii_smem, int lane,
                                                 int wlen, int end_rel, int nload, int lb, int& advance) {
    auto& sp = S.u.sp;
    uint8_t* const B = S.B();
    uint8_t* const KC = sp.K();
    ClassTables T = P.cls;
    T.ascii = ascii_smem;
    for (int w = lane - lb; w < nload; w += 32) {
        const uint8_t b = B[w];
        uint8_t k;
        if (b < 0x80) k = ascii_smem[b];
        else if (is_cont_byte(b) && w > -lb) {
            int j = w - 1;
            while (j >= -lb && j > w - 4 && is_cont_byte(B[j])) --j;
            k = C_CONT;
            if (j >= -lb && j > w - 4 && B[j] >= 0xC0) k |= char_class(B, j, end_rel, T);
        } else k = char_class(B, w, end_rel, T);
        KC[w] = k;
    }
    __syncwarp();
    const bool digits = P.spec.pat == PAT_GPT2_DIGITS;
    int ns = 0;
    for (int it = 0; it * 32 < wlen; ++it) {
        const int w = it * 32 + lane;
        const bool st = w < wlen && (w == 0 || gpt2_piece_starts_t(B, ClsArray{KC}, w, -lb, end_rel, digits));
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, st);
        if (st) S.seg[ns + __popc(m & ((1u << lane) - 1u))] = (uint16_t)(w | F_MATCH);
        if (lane == 0) S.segbits[it] = m;
        ns += __popc(m);
    }
    for (int it = (wlen + 31) / 32 + lane; it < NWORDS; it += 32) S.segbits[it] = 0;
    __syncwarp();
    if (wlen == end_rel) {
        advance = wlen;
        if (lane == 0) S.seg[ns] = (uint16_t)wlen;
    } else {
        // the last piece may continue beyond the window: redo it from its start in the next window
        --ns;
        advance = S.seg[ns] & POS_MASK;
        if (lane == 0 && advance > 0) S.segbits[advance >> 5] &= ~(1u << (advance & 31));
    }
    __syncwarp();
    return ns;
}

// Split phase for the fused BERT pre-tokenisation ( \s+ removed, then punctuation / CJK characters isolated ): a closed-form
// per-position predicate replaces match evaluation + chain resolution.  A whitespace run is one dropped segment, every
// punctuation / CJK character is its own segment, everything else forms words that start after whitespace, after a punctuation
// character, or at the element start (same segments as the generic chain for PAT_BERT_FUSED, which the tests compare with the
// two reference RegexSplit ops run back to back).
__device__ __forceinline__ int split_window_bert(WarpSmem& S, const RowParams& P, const uint8_t* ascii_smem, int lane,
                                                 int wlen, int end_rel, int nload, int lb, int& advance) {
    auto& sp = S.u.sp;
    uint8_t* const B = S.B();
    uint8_t* const KC = sp.K();
    ClassTables T = P.cls;
    T.ascii = ascii_smem;
    for (int w = lane - lb; w < nload; w += 32) {
        const uint8_t b = B[w];
        uint8_t k;
        if (b < 0x80) k = ascii_smem[b];
        else if (is_cont_byte(b) && w > -lb) {
            int j = w - 1;
            while (j >= -lb && j > w - 4 && is_cont_byte(B[j])) --j;
            k = C_CONT;
            if (j >= -lb && j > w - 4 && B[j] >= 0xC0) k |= char_class(B, j, end_rel, T);
        } else k = char_class(B, w, end_rel, T);
        KC[w] = k;
    }
    __syncwarp();
    int ns = 0;
    for (int it = 0; it * 32 < wlen; ++it) {
        const int w = it * 32 + lane;
        bool st = false;
        uint16_t sg = 0;
        if (w < wlen) {
            const uint8_t c = KC[w];
            if (!(c & C_CONT)) {
                const uint8_t p = (w > -lb) ? KC[w - 1] : (uint8_t)C_S;       // (element start behaves like "after whitespace")
                if (c & C_S) { st = w == 0 || !(p & C_S); sg = F_MATCH | F_DROP; }
                else if (c & C_BP) { st = true; sg = F_MATCH; }
                else st = w == 0 || (p & (C_S | C_BP)) != 0;
            }
            sg |= (uint16_t)w;
        }
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, st);
        if (st) S.seg[ns + __popc(m & ((1u << lane) - 1u))] = sg;
        if (lane == 0) S.segbits[it] = m;
        ns += __popc(m);
    }
    for (int it = (wlen + 31) / 32 + lane; it < NWORDS; it += 32) S.segbits[it] = 0;
    __syncwarp();
    if (wlen == end_rel) {
        advance = wlen;
        if (lane == 0) S.seg[ns] = (uint16_t)wlen;
    } else {
        // the last segment may continue beyond the window: redo it from its start in the next window
        --ns;
        advance = S.seg[ns] & POS_MASK;
        if (lane == 0 && advance > 0) S.segbits[advance >> 5] &= ~(1u << (advance & 31));
    }
    __syncwarp();
    return ns;
}

// Sequentially (lane 0) find the extent of the segment starting at chars[pos] when it does not fit
// a window.  Returns its length; is_match/drop describe it.
__device__ __noinline__ int giant_segment(const RowParams& P, int pos, int end_rel, int& is_match, int& drop) {
    ScanCtx sc{P.chars + pos, end_rel, P.cls, P.spec.pat, P.spec.class_mask};
    const Match m = match_rep(sc, P.spec, P.repeat != 0, 0, end_rel);
    if (m.len > 0) { is_match = 1; drop = m.drop; return m.len; }
    is_match = 0; drop = 0;
    int q = sc.next(0);
    while (q < end_rel && match_rep(sc, P.spec, P.repeat != 0, q, end_rel).len == 0) q = sc.next(q);
    return q;
}

// any set bit in [a, b)
__device__ __forceinline__ bool range_any(const uint32_t* words, int a, int b) {
    if (b <= a) return false;
    int wa = a >> 5;
    const int wb = (b - 1) >> 5;
    uint32_t m = words[wa] & (0xFFFFFFFFu << (a & 31));
    for (; wa < wb; ++wa, m = words[wa]) if (m) return true;
    return (m & (0xFFFFFFFFu >> (31 - ((b - 1) & 31)))) != 0;
}

// ---------------------------------------------------------------------------------------------------------
// BPE over the kept segments of a window.  Conventions of the shared-memory state (S.u.bp):
//   ids[w]  token id of the symbol living at position w, -1 = no symbol (merged away / dropped segment)
//   key[w]  packed (rank << 12 | birth) of the pair (previous live symbol, symbol at w); kNoKey if none.
//           Initial births are the window positions (< WIN), the pairs created by the m-th merge of a
//           segment get WIN + m: the same order as the reference's push-sequence numbers.
//   S.actbits bit w = key[w] != kNoKey after the initial lookups.
// ---------------------------------------------------------------------------------------------------------

// Position-parallel symbolisation + initial pair keys (every symbol one byte; ranks from the byte-pair table).
// Returns true if the window needs the serial path (a multi-byte symbol or a dropped byte was seen).
__device__ __forceinline__ bool bpe_symbolize_keys(WarpSmem& S, const BpeTables& BT, int lane, int send) {
    auto& bp = S.u.bp;
    const uint8_t* B = S.B();
    bool complex = false;
    for (int it = 0; it * 32 < send; ++it) {
        const int w = it * 32 + lane;
        bool found = false;
        if (w < send) {
            const uint8_t c = B[w];
            int32_t id = BT.byte_sym[c];
            if (id == kSymWalk) {
                const int pe = next_bit(S.segbits, w, send);
                int j = w;
                id = trie_longest(BT.trie, B, j, pe);
                if (id >= 0 && j != w + 1) complex = true;
            }
            if (id < 0) { id = BT.byte_miss[c]; if (id < 0) complex = true; }
            bp.ids[w] = id;
            uint32_t k = kNoKey;
            if (w > 0 && !((S.segbits[w >> 5] >> (w & 31)) & 1u)) {
                const uint32_t r = __ldg(BT.pair_rank + (((uint32_t)B[w - 1] << 8) | c));
                if (r != kNoKey) { found = true; k = (r << kPackedBirthBits) | (uint32_t)w; }
            }
            bp.key[w] = k;
        }
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, found);
        if (lane == 0) S.actbits[it] = m;
    }
    return __any_sync(0xFFFFFFFFu, complex);
}

// Serial per-lane path for windows with multi-byte symbols / dropped bytes.
__device__ __noinline__ void bpe_window_serial(WarpSmem& S, const BpeTables& BT, const RowParams& P, int lane, int ns, bool whole) {
    auto& bp = S.u.bp;
    for (int j = lane; j < ns; j += 32) {
        const uint16_t sg = S.seg[j];
        const int s = sg & POS_MASK, e = S.seg[j + 1] & POS_MASK;
        int c = 0;
        if (whole || seg_kept(sg, P.spec.pat, P.mode, P.invert)) {
            const int n = bpe_symbolize(BT, S.B(), s, e, bp.ids + s);
            bool tie = false;
            c = bpe_merge_packed(BT.merges, bp.ids + s, bp.key + s, n, BT.merges.tie_check ? &tie : nullptr);
            if (tie) atomicOr(&P.status[ST_ERROR], ERR_HEAP_TIE);      // (see bpe_merge_queue)
        }
        for (int t = s + c; t < e; ++t) bp.ids[t] = -1;
    }
}

// Merge phase: segments that own a mergeable pair enter a lane-level work queue; each lane performs one merge
// step per iteration (argmin over the packed keys, lazy deletion of the right operand, two new lookups).
__device__ __forceinline__ void bpe_merge_queue(WarpSmem& S, const BpeTables& BT, const RowParams& P, int lane, int ns, bool whole) {
    auto& bp = S.u.bp;
    const uint32_t lt = (1u << lane) - 1u;
    // segments that need merging -> S.act[]; dropped segments are erased; 2-symbol segments finish here
    int nact = 0;
    for (int j0 = 0; j0 < ns; j0 += 32) {
        const int j = j0 + lane;
        bool act = false;
        if (j < ns) {
            const uint16_t sg = S.seg[j];
            const int s = sg & POS_MASK, e = S.seg[j + 1] & POS_MASK;
            if (!(whole || seg_kept(sg, P.spec.pat, P.mode, P.invert))) {
                for (int t = s; t < e; ++t) bp.ids[t] = -1;
            } else if (e - s >= 2 && range_any(S.actbits, s + 1, e)) {
                if (e - s == 2) { bp.ids[s] = __ldg(BT.merges.rank_newid + (bp.key[s + 1] >> kPackedBirthBits)); bp.ids[s + 1] = -1; }
                else act = true;
            }
        }
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, act);
        if (act) S.act[nact + __popc(m & lt)] = (uint16_t)j;
        nact += __popc(m);
    }
    __syncwarp();
    int head = 0, s = 0, n0 = 0, merges = 0;
    bool have = false;
    for (;;) {
        const uint32_t need = __ballot_sync(0xFFFFFFFFu, !have);
        if (!have) {
            const int qi = head + __popc(need & lt);
            if (qi < nact) {
                const int j = S.act[qi];
                s = S.seg[j] & POS_MASK;
                n0 = (S.seg[j + 1] & POS_MASK) - s;
                merges = 0;
                have = true;
            }
        }
        head += __popc(need);
        if (!__any_sync(0xFFFFFFFFu, have)) break;
        if (have) {
            uint32_t best = kNoKey;
            int bk = -1, live = 0;
            for (int k = 1; k < n0; ++k) {
                const uint32_t q = bp.key[s + k];
                live += (q != kNoKey);
                if (q < best) { best = q; bk = k; }
            }
            if (bk < 0) { have = false; continue; }
            int pl = bk - 1;
            while (bp.ids[s + pl] < 0) --pl;                  // left operand (skip dead slots)
            const int32_t nid = __ldg(BT.merges.rank_newid + (best >> kPackedBirthBits));
            bp.ids[s + pl] = nid;
            bp.ids[s + bk] = -1;
            bp.key[s + bk] = kNoKey;
            --live;
            ++merges;
            const uint32_t birth = (uint32_t)(WIN + merges);
            int pp = pl - 1;
            while (pp >= 0 && bp.ids[s + pp] < 0) --pp;
            int nr = bk + 1;
            while (nr < n0 && bp.ids[s + nr] < 0) ++nr;
            bool fl = false;
            if (pl > 0) {
                live -= (bp.key[s + pl] != kNoKey);
                uint32_t kk = kNoKey;
                int32_t r, v;
                fl = pp >= 0 && merge_find(BT.merges, bp.ids[s + pp], nid, r, v);
                if (fl) { kk = ((uint32_t)r << kPackedBirthBits) | birth; ++live; }
                bp.key[s + pl] = kk;
            }
            if (nr < n0) {
                live -= (bp.key[s + nr] != kNoKey);
                uint32_t kk = kNoKey;
                int32_t r, v;
                if (merge_find(BT.merges, nid, bp.ids[s + nr], r, v)) {
                    kk = ((uint32_t)r << kPackedBirthBits) | birth; ++live;
                    // The merge found its own product on both sides (only with tokens that several merges produce): the two new pairs tie
                    // on (rank, seq) and the reference pops them in std::priority_queue's heap order.  This loop takes the left pair; the
                    // call reports the tie (B200TOK_E_UNSUPPORTED) instead of returning a result that may differ.  (Rows of the fused
                    // GPT-2 / Llama-3 path never get here with such a tie: they are redone exactly, rows_kernel `fits`.)
                    if (BT.merges.tie_check && fl && bp.ids[s + pp] == nid && bp.ids[s + nr] == nid) atomicOr(&P.status[ST_ERROR], ERR_HEAP_TIE);
                }
                bp.key[s + nr] = kk;
            }
            if (live == 0) have = false;                      // nothing left to merge in this segment
        }
    }
}

__device__ __forceinline__ void bpe_window_pieces(WarpSmem& S, const BpeTables& BT, const RowParams& P, int lane,
                                                  int ns, int send, bool whole, bool keys_ready, bool complex) {
    if (!keys_ready) complex = bpe_symbolize_keys(S, BT, lane, send);
    __syncwarp();
    if (complex) bpe_window_serial(S, BT, P, lane, ns, whole);
    else bpe_merge_queue(S, BT, P, lane, ns, whole);
}

// GPT-2 (isolate) split + symbolisation + initial keys in ONE position-parallel pass, for windows whose bytes are all
// ASCII (classes come from the 128-entry table applied to the neighbouring bytes directly).  Returns the segment
// count like split_window_gpt2; `complex` reports that the serial BPE path is needed.
__device__ __forceinline__ int gpt2_ascii_fused_window_v5(WarpSmem& S, const BpeTables& BT, const RowParams& P, const uint8_t* ascii_smem,
                                                       int lane, int wlen, int end_rel, int nload, int lb, int& advance, bool& complex_out) {
    auto& bp = S.u.bp;
    const uint8_t* B = S.B();
    const ClsAsciiLut K{B, ascii_smem};
    const bool digits = P.spec.pat == PAT_GPT2_DIGITS;
    bool complex = false;
    int ns = 0;
    for (int it = 0; it * 32 < wlen; ++it) {
        const int w = it * 32 + lane;
        bool st = false, found = false;
        if (w < wlen) {
            st = (w == 0) || gpt2_piece_starts_t(B, K, w, -lb, end_rel, digits);
            const uint8_t c = B[w];
            int32_t id = BT.byte_sym[c];
            if (id < 0) {
                if (id == kSymWalk) {           // a longer token may start here: if one does (even across a piece
                    int j = w;                  // boundary), the exact piece-limited walk is left to the serial path
                    id = trie_longest(BT.trie, B, j, nload);
                    if (id >= 0 && j != w + 1) complex = true;
                }
                if (id < 0) { id = BT.byte_miss[c]; if (id < 0) complex = true; }
            }
            bp.ids[w] = id;
            uint32_t k = kNoKey;
            if (!st) {
                const uint32_t r = __ldg(BT.pair_rank + (((uint32_t)B[w - 1] << 8) | c));
                if (r != kNoKey) { found = true; k = (r << kPackedBirthBits) | (uint32_t)w; }
            }
            bp.key[w] = k;
        }
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, st);
        const uint32_t ma = __ballot_sync(0xFFFFFFFFu, found);
        if (st) S.seg[ns + __popc(m & ((1u << lane) - 1u))] = (uint16_t)(w | F_MATCH);
        if (lane == 0) { S.segbits[it] = m; S.actbits[it] = ma; }
        ns += __popc(m);
    }
    for (int it = (wlen + 31) / 32 + lane; it < NWORDS; it += 32) S.segbits[it] = 0;
    complex_out = __any_sync(0xFFFFFFFFu, complex);
    __syncwarp();
    if (wlen == end_rel) {
        advance = wlen;
        if (lane == 0) S.seg[ns] = (uint16_t)wlen;
    } else {
        --ns;
        advance = S.seg[ns] & POS_MASK;
    }
    __syncwarp();
    return ns;
}

// ---------------------------------------------------------------------------------------------------------
// Helpers of the bit-mask GPT-2 -> BPE kernel (kernels_fast.cuh).
//   lut32[c]: class bits below | one-byte symbol id << 12
// ---------------------------------------------------------------------------------------------------------
enum : uint32_t { V7_L = 1, V7_N = 2, V7_S = 4, V7_SP = 8, V7_AP = 16, V7_WALK = 32, V7_BAD = 64, V7_CONT = 128, V7_NL = 256 };
constexpr int V7_ID_SHIFT = 12;
constexpr uint32_t FULL = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t v7_lut_entry(const RowParams& P, int c) {
    uint32_t g = c < 128 ? (uint32_t)(P.cls.ascii[c] & (C_L | C_N | C_S)) : 0u;
    if (c == 0x20) g |= V7_SP;
    if (c == '\'') g |= V7_AP;
    if (c < 128 && (P.cls.ascii[c] & C_NL)) g |= V7_NL;
    int32_t id = P.bpe.byte_sym[c];
    if (id == kSymWalk) {                      // a longer token may start with c: the one-byte result is used unless the walk finds one
        g |= V7_WALK;
        id = P.bpe.trie.value[P.bpe.trie.root_child[c]];
    }
    if (id < 0) id = P.bpe.byte_miss[c];
    if (id < 0 || id >= (1 << (32 - V7_ID_SHIFT))) { g |= V7_BAD; id = 0; }
    return g | ((uint32_t)id << V7_ID_SHIFT);
}

// ballot of "g has a bit of mask": and + setp + vote (the C++ form is canonicalised into shift/and/compare per class bit)
__device__ __forceinline__ uint32_t ballot_bits(uint32_t g, uint32_t mask) {
    uint32_t r;
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tand.b32 t, %1, %2;\n\tsetp.ne.u32 p, t, 0;\n\tvote.sync.ballot.b32 %0, p, 0xffffffff;\n\t}"
                 : "=r"(r) : "r"(g), "r"(mask));
    return r;
}

// bits [0, n) of a word whose first position is `base`, for a limit given as a position
__device__ __forceinline__ uint32_t v7_below(int limit, int base) {
    const int r = limit - base;
    return r >= 32 ? FULL : r > 0 ? ((1u << r) - 1u) : 0u;
}

// WordPiece for all kept segments (words) of a window (src/wordpiece_tokenizer.cpp:96-130, tok_core.cuh wordpiece_word).
//   pass A, one lane per word, no loop per word: dropped segments stay dead, over-long / empty words become [unk], ONE-byte words
//           take their token from val1, TWO-byte ASCII words are settled by the two-byte jump table (+ val1 of the ## trie when only
//           the first byte matched) — on split text most words end here (every punctuation mark is a word of its own);
//   pass B, the remaining words: a per-lane state machine driven by a warp work queue — every iteration performs ONE trie step for
//           whatever word a lane holds, and a lane that finishes its word takes the next one, so words of different lengths do not
//           idle the rest of the warp; every walk starts from the jump table (its first two steps in one load).
__device__ __forceinline__ void wordpiece_window_pieces(WarpSmem& S, const RowParams& P, int lane, int ns, bool whole) {
    auto& bp = S.u.bp;
    const uint8_t* B = S.B();
    const WordpieceTables& T = P.wp;
    const uint32_t lt = (1u << lane) - 1u;
    uint16_t* const queue = reinterpret_cast<uint16_t*>(bp.key);           // word indices of pass B (WordPiece has no merge keys)
    {   // every slot starts dead (position-parallel); the lanes then write tokens only
        const int send = S.seg[ns] & POS_MASK;
        for (int w = lane; w < send; w += 32) bp.ids[w] = -1;
        __syncwarp();
    }
    const bool tables = T.root.val1 && T.sub.val1 && T.root.jump2;
    int nq = 0;
    for (int j0 = 0; j0 < ns; j0 += 32) {
        const int j = j0 + lane;
        bool later = false;
        if (j < ns) {
            const uint16_t sg = S.seg[j];
            const int s = sg & POS_MASK, e = S.seg[j + 1] & POS_MASK;
            if (!(whole || seg_kept(sg, P.spec.pat, P.mode, P.invert))) {
                // dropped segment: its slots stay dead
            } else if (e - s > T.max_bytes || e <= s) {                      // :100-103 (and the zero-length word, see tok_core.cuh)
                bp.ids[s] = P.unk_id;
            } else if (tables && e - s == 1) {
                const int32_t v = __ldg(T.root.val1 + B[s]);
                bp.ids[s] = v >= 0 ? v : P.unk_id;
            } else if (tables && e - s == 2 && ((B[s] | B[s + 1]) & 0x80u) == 0u) {
                const uint2 jp = __ldg(reinterpret_cast<const uint2*>(T.root.jump2) + (((uint32_t)B[s] << 7) | B[s + 1]));
                const int len = (int)((jp.y >> 24) & 3u);
                const int32_t v = (int32_t)(jp.y & 0xFFFFFFu) - 1;
                if (len == 2) bp.ids[s] = v;
                else {
                    const int32_t v2 = len == 1 ? __ldg(T.sub.val1 + B[s + 1]) : -1;      // first byte a token: the second must be a ## token
                    if (v2 >= 0) { bp.ids[s] = v; bp.ids[s + 1] = v2; } else bp.ids[s] = P.unk_id;
                }
            } else later = true;
        }
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, later);
        if (later) queue[nq + __popc(m & lt)] = (uint16_t)j;
        nq += __popc(m);
    }
    __syncwarp();
    if (nq == 0) return;
    int head = 0, s = 0, e = 0, i = 0, n = 0, best = 0;
    int32_t node = -1, found = -1;
    bool have = false, sub = false, jumped = false;
    // Start of a longest-match walk at position q of the lane's word (tok_core.cuh rank_trie_longest)
    auto start_walk = [&](const RankTrie& t, int q) {
        found = -1; best = q; jumped = false;
        const uint32_t b0 = B[q];
        if (e - q == 1 && t.val1) {
            found = __ldg(t.val1 + b0);
            best = q + 1; i = q + 1; node = -1;
        } else if (e - q >= 2 && t.jump2 && ((b0 | B[q + 1]) & 0x80u) == 0u) {
            const uint2 j = __ldg(reinterpret_cast<const uint2*>(t.jump2) + ((b0 << 7) | B[q + 1]));
            if (!(j.y & kJumpHas1)) { node = -1; i = q; }
            else {
                found = (int32_t)(j.y & 0xFFFFFFu) - 1;
                best = q + (int)((j.y >> 24) & 3u);
                i = q + 2;
                node = i < e ? (int32_t)j.x : -1;
                jumped = true;
            }
        } else { i = q; node = t.root_child[b0]; }
    };
    while (head < nq || __any_sync(0xFFFFFFFFu, have)) {
        const uint32_t need = __ballot_sync(0xFFFFFFFFu, !have);
        if (!have) {
            const int qi = head + __popc(need & lt);
            if (qi < nq) {
                const int j = queue[qi];
                s = S.seg[j] & POS_MASK; e = S.seg[j + 1] & POS_MASK;
                have = true; sub = false; n = 0;
                start_walk(T.root, s);
            }
        }
        head += __popc(need);
        if (have) {
            if (node >= 0) {                                              // one step of the longest-match walk
                const RankNode& nd = (sub ? T.sub.nodes : T.root.nodes)[node];
                if (!jumped) {
                    ++i;
                    const int32_t v = nd.value;
                    if (v != -1) { found = v; best = i; }
                }
                jumped = false;
                if (i >= e) node = -1;
                else {
                    const uint32_t ch = B[i], wd = ch >> 5, bit = ch & 31u;
                    const uint32_t bw = nd.bits[wd];
                    node = ((bw >> bit) & 1u) ? nd.base + (int32_t)nd.cum[wd] + __popc(bw & ((1u << bit) - 1u)) : -1;
                }
            }
            if (node < 0) {                                               // the walk ended: a token, or the whole word is unknown
                bool done = false;
                if (found < 0) {                                          // :107-112, :116-126: rewind, the word becomes one [unk]
                    for (int t = s + 1; t < s + n; ++t) bp.ids[t] = -1;
                    bp.ids[s] = P.unk_id; n = 1; done = true;
                }
                else {
                    bp.ids[s + n++] = found;
                    if (best >= e) done = true;
                    else { sub = true; start_walk(T.sub, best); }
                }
                if (done) have = false;
            }
        }
    }
}

// Reserve (end-begin)+suffix_len slots for a BPE piece handled by giant_bpe_kernel and queue it.
__device__ __forceinline__ void reserve_giant_bpe(const RowParams& P, int row, int begin, int end, int64_t base, int lane,
                                                  int& emitted, int& holes) {
    const int reserve = (end - begin) + P.suffix_len;
    const int64_t o = base + emitted;
    if (o + reserve <= P.tmp_cap) {
        for (int t = lane; t < reserve; t += 32) P.tmp_a[o + t] = -1;
        if (lane == 0) {
            const int gi = atomicAdd(&P.status[ST_NGIANT], 1);
            if (gi < P.giants_cap) P.giants[gi] = GiantItem{row, begin, end, emitted};
            else atomicOr(&P.status[ST_ERROR], ERR_GIANT_LIST);
        }
    } else if (lane == 0) atomicOr(&P.status[ST_ERROR], ERR_TMP_OVERFLOW);
    emitted += reserve;
    holes += reserve;
}

template <int OP>
__global__ void __launch_bounds__(BLOCK_THREADS, 4) rows_kernel(const __grid_constant__ RowParams P) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* ascii_smem = smem_raw;                                            // [128]
    int32_t* bytesym_smem = reinterpret_cast<int32_t*>(smem_raw + 128);        // [256]
    WarpSmem* warps = reinterpret_cast<WarpSmem*>(smem_raw + kRowsSmemFixed);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    WarpSmem& S = warps[wib];
    const bool listed = P.row_list != nullptr;
    if (listed && P.status[ST_NREDO] == 0) return;
    if (threadIdx.x < 128) ascii_smem[threadIdx.x] = P.cls.ascii[threadIdx.x];
    if (OP == OP_BPE) bytesym_smem[threadIdx.x] = P.bpe.byte_sym[threadIdx.x];
    __syncthreads();
    BpeTables BT = P.bpe;
    BT.byte_sym = bytesym_smem;

    for (;;) {
        int row = 0;
        if (lane == 0) row = atomicAdd(&P.status[listed ? ST_TICKET2 : ST_TICKET], 1);
        row = __shfl_sync(0xFFFFFFFFu, row, 0);
        if (listed) {
            if (row >= P.status[ST_NREDO]) break;
            row = P.row_list[row];
        } else if (row >= P.n_rows) break;
        const int p0 = P.rb[row], p1 = P.re[row];
        int64_t base;
        if (P.direct_base) {
            base = p1 > p0 ? (int64_t)(P.begins[p0] - P.direct_byte0) + (int64_t)(p0 - P.direct_elem0) * P.direct_extra : 0;
            if (lane == 0) const_cast<int32_t*>(P.row_base)[row] = (int32_t)base;     // the compaction pass reads it
        } else base = P.row_base[row];
        int emitted = 0;       // slots used in this row's range
        int holes = 0;         // reserved-but-unfilled slots (giant BPE pieces)
        for (int p = p0; p < p1; ++p) {
            const int eb = P.begins[p], ee = P.ends[p];
            const bool skip = P.skips && P.skips[p];
            if (OP == OP_SPLIT && skip) {   // src/regex_split.cpp:231-234
                if (lane == 0) {
                    const int64_t o = base + emitted;
                    if (o < P.tmp_cap) { P.tmp_a[o] = eb; P.tmp_b[o] = ee; P.tmp_c[o] = 1; }
                    else atomicOr(&P.status[ST_ERROR], ERR_TMP_OVERFLOW);
                }
                ++emitted;
                continue;
            }
            const bool whole = skip || P.spec.pat == PAT_NONE;   // the element is one piece
            SplitEmitter em;
            int last_was_match = 1;
            const bool stateful = OP == OP_SPLIT && (P.mode >= SPLIT_MERGED_PREV || P.max_splits != -1);
            if (OP == OP_SPLIT) em.reset(P.mode, P.invert != 0, P.max_splits, ee - eb);
            if (OP == OP_WORDPIECE && whole && ee <= eb) {   // zero-length word -> [unk] (see tok_core.cuh)
                if (lane == 0) { const int64_t o = base + emitted; if (o < P.tmp_cap) P.tmp_a[o] = P.unk_id; else atomicOr(&P.status[ST_ERROR], ERR_TMP_OVERFLOW); }
                ++emitted;
                continue;
            }
            if (OP == OP_BPE && whole && P.suffix_len > 0 && ee <= eb) {   // "" + end_suffix still yields tokens
                reserve_giant_bpe(P, row, eb, eb, base, lane, emitted, holes);
                continue;
            }
            int pos = eb;
            while (pos < ee) {
                const int end_rel = ee - pos;
                const int wlen = end_rel < WIN ? end_rel : WIN;
                int ns = 0, advance = 0;
                bool keys_ready = false, complex_win = false;
                // BPE with an end_suffix, and the rows the fast kernel handed back under a vocabulary whose merges can tie in the
                // reference's queue (MergeTable::tie_check — the row may have met such a tie): every piece takes the exact heap form
                // (giant_bpe_kernel: std::priority_queue's pop order restated).
                const bool fits = !(whole && end_rel > WIN) && !(OP == OP_BPE && (P.suffix_len > 0 || (listed && P.bpe.merges.tie_check)));
                if (fits) {
                    const int nload = end_rel < wlen + LA ? end_rel : wlen + LA;
                    const int lb = (pos - eb) < LBK ? (pos - eb) : LBK;   // look-back bytes available inside the element
                    uint32_t hibits = 0;
                    const uint8_t* src = P.chars + pos - lb;
                    if (((reinterpret_cast<uintptr_t>(src) | (uintptr_t)lb) & 15) == 0) {
                        // 16-byte vector loads (rows of the benchmark configs start on 16-byte boundaries); the last
                        // quad may read up to 15 bytes past the element — inside the padded chars allocation
                        uint4* dst = reinterpret_cast<uint4*>(S.B() - lb);
                        const int nq = (lb + nload + 15) >> 4;
                        for (int q = lane; q < nq; q += 32) {
                            const uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + q);
                            dst[q] = v;
                            if ((q << 4) + 16 <= lb + nload) hibits |= v.x | v.y | v.z | v.w;
                            else {
                                const uint32_t ww[4] = {v.x, v.y, v.z, v.w};
                                for (int t = 0; t < 16; ++t) if ((q << 4) + t < lb + nload) hibits |= (ww[t >> 2] >> ((t & 3) * 8)) & 0xFFu;
                            }
                        }
                    } else {
                        for (int w = lane - lb; w < nload; w += 32) {
                            const uint8_t bb = __ldg(P.chars + pos + w);
                            S.B()[w] = bb;
                            hibits |= bb;
                        }
                    }
                    const bool all_ascii = !__any_sync(0xFFFFFFFFu, hibits & 0x80808080u);
                    __syncwarp();
                    if (whole) {
                        if (lane == 0) { S.seg[0] = F_MATCH; S.seg[1] = (uint16_t)wlen; }
                        if (lane < NWORDS) S.segbits[lane] = lane == 0 ? 1u : 0u;
                        ns = 1; advance = wlen;
                        __syncwarp();
                    } else {
                        if ((P.spec.pat == PAT_GPT2 || P.spec.pat == PAT_GPT2_DIGITS) && P.mode == SPLIT_ISOLATED && !P.repeat && P.max_splits == -1) {
                            if (OP == OP_BPE && all_ascii) {
                                ns = gpt2_ascii_fused_window_v5(S, BT, P, ascii_smem, lane, wlen, end_rel, nload, lb, advance, complex_win);
                                keys_ready = true;
                            } else
                                ns = split_window_gpt2(S, P, ascii_smem, lane, wlen, end_rel, nload, lb, advance);
                        } else if (P.spec.pat == PAT_BERT_FUSED && OP != OP_SPLIT && !(P.dbg_flags & 16))
                            ns = split_window_bert(S, P, ascii_smem, lane, wlen, end_rel, nload, lb, advance);
                        else
                            ns = split_window(S, P, ascii_smem, lane, wlen, end_rel, nload, advance);
                    }
                }
                if (advance == 0) {
                    // ---- giant: a segment that does not fit the window (or BPE with end_suffix) ----
                    int glen = end_rel, is_m = 1, drop = 0;
                    if (!whole) {
                        if (lane == 0) glen = giant_segment(P, pos, end_rel, is_m, drop);
                        glen = __shfl_sync(0xFFFFFFFFu, glen, 0);
                        is_m = __shfl_sync(0xFFFFFFFFu, is_m, 0);
                        drop = __shfl_sync(0xFFFFFFFFu, drop, 0);
                    }
                    const uint16_t sg = (uint16_t)((is_m ? F_MATCH : 0) | (drop ? F_DROP : 0));
                    const bool kept = whole || seg_kept(sg, P.spec.pat, P.mode, P.invert);
                    if (OP == OP_SPLIT) {
                        if (lane == 0) {
                            int ob, oe;
                            const bool out = stateful || whole ? em.add(pos - eb, pos + glen - eb, is_m ? !em.invert : em.invert, ob, oe) : kept;
                            if (!(stateful || whole)) { ob = pos - eb; oe = pos + glen - eb; }
                            if (out) {
                                const int64_t o = base + emitted;
                                if (o < P.tmp_cap) { P.tmp_a[o] = eb + ob; P.tmp_b[o] = eb + oe; P.tmp_c[o] = 0; }
                                else atomicOr(&P.status[ST_ERROR], ERR_TMP_OVERFLOW);
                            }
                            emitted += out ? 1 : 0;
                        }
                        emitted = __shfl_sync(0xFFFFFFFFu, emitted, 0);
                        last_was_match = is_m;
                    } else if (kept) {
                        if (OP == OP_WORDPIECE) {
                            if (lane == 0) {
                                const int64_t o = base + emitted;
                                int c = 0;
                                if (o + glen <= P.tmp_cap) c = wordpiece_word(P.wp, P.chars, pos, pos + glen, P.unk_id, P.tmp_a + o);
                                else atomicOr(&P.status[ST_ERROR], ERR_TMP_OVERFLOW);
                                emitted += c;
                            }
                            emitted = __shfl_sync(0xFFFFFFFFu, emitted, 0);
                        } else {
                            reserve_giant_bpe(P, row, pos, pos + glen, base, lane, emitted, holes);
                        }
                    }
                    pos += glen;
                    continue;
                }
                // ---- regular window: ns complete segments, S.seg[ns] is the end sentinel ----
                if (OP == OP_SPLIT) {
                    if (stateful) {
                        if (lane == 0) {
                            for (int j = 0; j < ns; ++j) {
                                const uint16_t sg = S.seg[j];
                                const int s = (pos - eb) + (sg & POS_MASK), e = (pos - eb) + (S.seg[j + 1] & POS_MASK);
                                int ob, oe;
                                if (em.add(s, e, (sg & F_MATCH) ? !em.invert : em.invert, ob, oe)) {
                                    const int64_t o = base + emitted;
                                    if (o < P.tmp_cap) { P.tmp_a[o] = eb + ob; P.tmp_b[o] = eb + oe; P.tmp_c[o] = 0; }
                                    else atomicOr(&P.status[ST_ERROR], ERR_TMP_OVERFLOW);
                                    ++emitted;
                                }
                                last_was_match = (sg & F_MATCH) ? 1 : 0;
                            }
                        }
                        emitted = __shfl_sync(0xFFFFFFFFu, emitted, 0);
                        last_was_match = __shfl_sync(0xFFFFFFFFu, last_was_match, 0);
                    } else {
                        for (int j0 = 0; j0 < ns; j0 += 32) {
                            const int j = j0 + lane;
                            bool k = false;
                            uint16_t sg = 0;
                            if (j < ns) { sg = S.seg[j]; k = seg_kept(sg, P.spec.pat, P.mode, P.invert); }
                            const uint32_t m = __ballot_sync(0xFFFFFFFFu, k);
                            if (k) {
                                const int64_t o = base + emitted + __popc(m & ((1u << lane) - 1u));
                                if (o < P.tmp_cap) { P.tmp_a[o] = pos + (sg & POS_MASK); P.tmp_b[o] = pos + (S.seg[j + 1] & POS_MASK); P.tmp_c[o] = 0; }
                                else atomicOr(&P.status[ST_ERROR], ERR_TMP_OVERFLOW);
                            }
                            emitted += __popc(m);
                        }
                    }
                } else {
                    // piece phase: every kept segment becomes tokens in S.u.bp.ids[start..], dead slots = -1
                    const int send = S.seg[ns] & POS_MASK;
                    if (OP == OP_BPE) bpe_window_pieces(S, BT, P, lane, ns, send, whole, keys_ready, complex_win);
                    else wordpiece_window_pieces(S, P, lane, ns, whole);
                    __syncwarp();
                    // output phase: position-parallel compaction of the live tokens into the row slot
                    if (base + emitted + send > P.tmp_cap) {
                        if (lane == 0) atomicOr(&P.status[ST_ERROR], ERR_TMP_OVERFLOW);
                    } else {
                        int32_t* outp = P.tmp_a + base + emitted;
                        const uint32_t ltm = (1u << lane) - 1u;
                        int n_out = 0;
                        for (int w = lane; w - lane < send; w += 128) {       // four words in flight
                            int32_t tok[4];
                            uint32_t m[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) tok[u] = (w + 32 * u) < send ? S.u.bp.ids[w + 32 * u] : -1;
#pragma unroll
                            for (int u = 0; u < 4; ++u) m[u] = __ballot_sync(0xFFFFFFFFu, tok[u] >= 0);
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                if (tok[u] >= 0) outp[n_out + __popc(m[u] & ltm)] = tok[u];
                                n_out += __popc(m[u]);
                            }
                        }
                        emitted += n_out;
                    }
                    __syncwarp();
                }
                pos += advance;
            }
            if (OP == OP_SPLIT && stateful) {   // src/regex_split.cpp:305-309
                if (lane == 0 && last_was_match) {
                    int ob, oe;
                    if (em.finish(ee - eb, ob, oe)) {
                        const int64_t o = base + emitted;
                        if (o < P.tmp_cap) { P.tmp_a[o] = eb + ob; P.tmp_b[o] = eb + oe; P.tmp_c[o] = 0; }
                        else atomicOr(&P.status[ST_ERROR], ERR_TMP_OVERFLOW);
                        ++emitted;
                    }
                }
                emitted = __shfl_sync(0xFFFFFFFFu, emitted, 0);
            }
        }
        if (lane == 0) {
            P.row_ext[row] = emitted;
            P.row_cnt[row] = emitted - holes;
            P.row_flag[row] = (uint8_t)((holes ? 1 : 0) | (P.row_list ? 2 : 0));      // bit 1: a row the fast kernel handed back
        }
    }
}

// Row capacities: tokens (or pieces) a row can produce at most.
//   BPE: sum(len + suffix_len); WordPiece / split: sum(len + 1)
__global__ void row_capacity_kernel(const int32_t* rb, const int32_t* re, const int32_t* begins, const int32_t* ends,
                                    int32_t n_rows, int32_t per_elem_extra, int32_t* cap) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    int64_t c = 0;
    for (int p = rb[r]; p < re[r]; ++p) { const int l = ends[p] - begins[p]; c += (l > 0 ? l : 0) + per_elem_extra; }
    cap[r] = (int32_t)(c > 0x7FFFFFFF ? 0x7FFFFFFF : c);
}

// Very long BPE pieces (and every piece when an end_suffix is configured): one thread per piece,
// heap-ordered merge loop with all state in a global scratch pool.
struct GiantParams {
    const GiantItem* items; const int32_t* status_in; int32_t giants_cap;
    const uint8_t* chars; BpeTables bpe; const uint8_t* suffix; int32_t suffix_len;
    const int32_t* row_base; int32_t* row_cnt; int32_t* tmp; uint8_t* pool; unsigned long long pool_cap;
    unsigned long long* pool_used; int32_t* status;
};
__global__ void giant_bpe_kernel(const GiantParams G) {
    int n_items = G.status_in[ST_NGIANT];
    if (n_items > G.giants_cap) n_items = G.giants_cap;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += gridDim.x * blockDim.x) {
        const GiantItem it = G.items[i];
        const int len = it.end - it.begin, n = len + G.suffix_len;
        // layout: heap [3n x 16 B] | sym_id, sym_prev, sym_next [2n x 4 B each] | bytes [n]; 16-byte granules
        const unsigned long long need = (48ull * n + 24ull * n + (unsigned long long)n + 15ull) & ~15ull;
        const unsigned long long off = atomicAdd(G.pool_used, need);
        if (off + need > G.pool_cap) { atomicOr(&G.status[ST_ERROR], ERR_GIANT_POOL); continue; }
        uint8_t* mem = G.pool + off;
        HeapEntry* heap = reinterpret_cast<HeapEntry*>(mem);
        int32_t* sym_id = reinterpret_cast<int32_t*>(mem + 48ull * n);
        int32_t* sym_prev = sym_id + 2 * n;
        int32_t* sym_next = sym_prev + 2 * n;
        uint8_t* bytes = reinterpret_cast<uint8_t*>(sym_next + 2 * n);
        for (int k = 0; k < len; ++k) bytes[k] = G.chars[it.begin + k];
        for (int k = 0; k < G.suffix_len; ++k) bytes[len + k] = G.suffix[k];
        const int m = bpe_symbolize(G.bpe, bytes, 0, n, sym_id);
        int32_t* out = G.tmp + G.row_base[it.row] + it.slot;
        const int cnt = bpe_merge_heap(G.bpe.merges, m, sym_id, sym_prev, sym_next, heap, out);
        for (int k = cnt; k < n; ++k) out[k] = -1;
        atomicAdd(&G.row_cnt[it.row], cnt);
    }
}

// Copies a chunk's status words into mapped host memory (the pipelined host path polls them after an event).
__global__ void publish_status_kernel(const int32_t* status, int32_t* host_mapped) {
    if (threadIdx.x < 16) host_mapped[threadIdx.x] = status[threadIdx.x];
    __threadfence_system();
}

// out_begins = exclusive scan(row_cnt) is done with cub; this finishes ends + total.
__global__ void finish_offsets_kernel(const int32_t* begins, const int32_t* cnt, int32_t n_rows, int32_t* ends,
                                      int32_t* status, int64_t* total_out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const int e = begins[r] + cnt[r];
    ends[r] = e;
    if (r == n_rows - 1) { status[ST_TOTAL] = e; if (total_out) *total_out = e; }
}

// Pipelined host path, zero-copy variant: chunk totals are chained on the device (base of chunk k = sum of the totals
// of chunks < k), so final row offsets and ids can be stored straight into the caller's pinned host buffers.
__global__ void chunk_base_kernel(const int32_t* begins, const int32_t* cnt, int32_t n_rows, long long* running_total, int32_t* status) {
    const long long base = *running_total;
    const int32_t total = begins[n_rows - 1] + cnt[n_rows - 1];
    status[ST_TOTAL] = total;
    status[ST_BASE] = (int32_t)base;
    *running_total = base + total;
}
__global__ void finish_offsets_host_kernel(const int32_t* begins, const int32_t* cnt, int32_t n_rows, const int32_t* status,
                                           int32_t* host_begins, int32_t* host_ends) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const int32_t b = status[ST_BASE] + begins[r];
    host_begins[r] = b;
    host_ends[r] = b + cnt[r];
}

// Copy every row's slots to their final place (warp per row); rows with holes are filtered.
__global__ void compact_rows_kernel(const int32_t* tmp_a, const int32_t* tmp_b, const uint8_t* tmp_c,
                                    const int32_t* row_base, const int32_t* row_ext, const uint8_t* row_flag,
                                    const int32_t* out_begin, int32_t n_rows,
                                    int32_t* out_a, int32_t* out_b, uint8_t* out_c, int64_t out_cap, int32_t* status,
                                    const int32_t* dst_base, const int32_t* row_cnt, int32_t* out_end, int64_t* total_out = nullptr) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int64_t base = dst_base ? (int64_t)*dst_base : 0;
    for (int r = warp; r < n_rows; r += nwarps) {
        const int64_t src = row_base[r];
        const int ext = row_ext[r];
        int64_t dst = base + out_begin[r];
        if (out_end && lane == 0) {          // folded finish_offsets: row end + chunk total
            const int32_t e = out_begin[r] + row_cnt[r];
            out_end[r] = e;
            if (r == n_rows - 1) { status[ST_TOTAL] = e; if (total_out) *total_out = e; }
        }
        if (!(row_flag[r] & 1)) {
            if (dst + ext > out_cap) { if (lane == 0) atomicOr(&status[ST_ERROR], ERR_TMP_OVERFLOW); continue; }
            if (row_flag[r] & 4) {             // 16-bit slot written by the fast kernel: widen while copying
                const uint16_t* sp = reinterpret_cast<const uint16_t*>(tmp_a + src);
                int32_t* dp = out_a + dst;
                for (int t = lane; t < ext; t += 128) {
                    uint16_t v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) v[u] = (t + 32 * u < ext) ? __ldcs(sp + t + 32 * u) : (uint16_t)0;
#pragma unroll
                    for (int u = 0; u < 4; ++u) if (t + 32 * u < ext) __stcs(dp + t + 32 * u, (int32_t)v[u]);
                }
            } else if (!out_b && !out_c) {     // ids only: four independent loads in flight per lane
                const int32_t* sp = tmp_a + src;
                int32_t* dp = out_a + dst;
                for (int t = lane; t < ext; t += 128) {
                    int32_t v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) v[u] = (t + 32 * u < ext) ? __ldcs(sp + t + 32 * u) : 0;
#pragma unroll
                    for (int u = 0; u < 4; ++u) if (t + 32 * u < ext) dp[t + 32 * u] = v[u];
                }
            } else {
                for (int t = lane; t < ext; t += 32) {
                    out_a[dst + t] = tmp_a[src + t];
                    if (out_b) out_b[dst + t] = tmp_b[src + t];
                    if (out_c) out_c[dst + t] = tmp_c[src + t];
                }
            }
        } else {
            for (int t0 = 0; t0 < ext; t0 += 32) {
                const int t = t0 + lane;
                const int v = t < ext ? tmp_a[src + t] : -1;
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, v >= 0);
                if (v >= 0) {
                    const int64_t o = dst + __popc(m & ((1u << lane) - 1u));
                    if (o < out_cap) out_a[o] = v;
                }
                dst += __popc(m);
            }
        }
    }
}


// Emit fused with the all-gatherv: every row is copied once from its worst-case slot and stored into the result
// buffers of all ranks over NVLink peer memory; row offsets are written shifted into this rank's slot.
__global__ void compact_rows_peer_kernel(const int32_t* tmp_a, const int32_t* row_base, const int32_t* row_ext, const uint8_t* row_flag,
                                         const int32_t* out_begin, const int32_t* row_cnt, int32_t n_rows, const PeerOut Q, int32_t* status,
                                         int64_t* total_out) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int64_t slot = (int64_t)Q.rank * Q.slot_capacity;
    for (int r = warp; r < n_rows; r += nwarps) {
        const int64_t src = row_base[r];
        const int ext = row_ext[r], cnt = row_cnt[r];
        const int64_t dst = slot + out_begin[r];
        if (lane < Q.world) {
            Q.begins[lane][(int64_t)Q.rank * Q.rows_per_rank + r] = (int32_t)dst;
            Q.ends[lane][(int64_t)Q.rank * Q.rows_per_rank + r] = (int32_t)(dst + cnt);
        }
        if (lane == 0 && r == n_rows - 1) { status[ST_TOTAL] = out_begin[r] + cnt; if (total_out) *total_out = out_begin[r] + cnt; }
        if (out_begin[r] + cnt > Q.slot_capacity) { if (lane == 0) atomicOr(&status[ST_ERROR], ERR_TMP_OVERFLOW); continue; }
        if (!(row_flag[r] & 1)) {
            for (int t = lane; t < ext; t += 128) {
                int32_t v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = (t + 32 * u < ext) ? __ldcs(tmp_a + src + t + 32 * u) : 0;
                for (int p = 0; p < Q.world; ++p) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) if (t + 32 * u < ext) peer_store(Q, p, dst + t + 32 * u, v[u]);
                }
            }
        } else {      // rows with holes (giant pieces): filter while copying
            int64_t d = dst;
            for (int t0 = 0; t0 < ext; t0 += 32) {
                const int t = t0 + lane;
                const int v = t < ext ? tmp_a[src + t] : -1;
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, v >= 0);
                if (v >= 0) { const int64_t o = d + __popc(m & ((1u << lane) - 1u)); for (int p = 0; p < Q.world; ++p) peer_store(Q, p, o, v); }
                d += __popc(m);
            }
        }
    }
}


// Sharded fast path: rows that the fast kernel handed back were finished by the generic kernels in the local worst-case buffer;
// copy them (filtering the holes of giant pieces) to the same positions of every rank's slot and publish their extents.
__global__ void peer_redo_rows_kernel(const int32_t* tmp_a, const int32_t* row_base, const int32_t* row_ext, const int32_t* row_list,
                                      const PeerOut Q, int32_t* status) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int n = status[ST_NREDO];
    const int64_t slot = (int64_t)Q.rank * Q.slot_capacity;
    for (int i = warp; i < n; i += nwarps) {
        const int r = row_list[i];
        const int64_t src = row_base[r];
        const int ext = row_ext[r];
        int64_t d = slot + src;
        for (int t0 = 0; t0 < ext; t0 += 32) {
            const int t = t0 + lane;
            const int v = t < ext ? tmp_a[src + t] : -1;
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, v >= 0);
            if (v >= 0) { const int64_t o = d + __popc(m & ((1u << lane) - 1u)); for (int p = 0; p < Q.world; ++p) peer_store(Q, p, o, v); }
            d += __popc(m);
        }
        const int cnt = (int)(d - slot - src);
        if (lane < Q.world) {
            Q.begins[lane][(int64_t)Q.rank * Q.rows_per_rank + r] = (int32_t)(slot + src);
            Q.ends[lane][(int64_t)Q.rank * Q.rows_per_rank + r] = (int32_t)(slot + src + cnt);
        }
        if (lane == 0) atomicAdd(&status[ST_TOTAL], cnt);
    }
}
// 16-bit wire format: widen the rows of every slot of this rank's staging copy into its final i32 buffer (warp per row).
__global__ void peer_expand_kernel(const uint16_t* staging, const int32_t* begins, const int32_t* ends, int64_t n_rows, int32_t* out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n_rows; r += nwarps) {
        const int64_t b = begins[r], e = ends[r];
        for (int64_t t = b + lane; t - lane < e; t += 128) {
            uint16_t v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = (t + 32 * u < e) ? staging[t + 32 * u] : (uint16_t)0;
#pragma unroll
            for (int u = 0; u < 4; ++u) if (t + 32 * u < e) out[t + 32 * u] = (int32_t)v[u];
        }
    }
}
__global__ void publish_total_kernel(const int32_t* status, int64_t* total_out) { if (total_out) *total_out = status[ST_TOTAL]; }

// ---- all-gatherv by PULL over NVLink peer memory ------------------------------------------------------------
// Every rank tokenises its shard with the ordinary one-GPU launch sequence into compact (begins, ends, ids); the ids are then
// packed to 16 bits (when the vocabulary allows) into a peer-mapped staging buffer.  After ONE cross-rank barrier each rank reads the
// staging buffers of all the others with 16-byte loads over NVLink and widens them straight into its own i32 result — the
// transfer and the widening are one pass, the tokenizer kernel runs at its one-GPU speed, and nothing is ever stored remotely.
struct PeerPull {
    int world, rank, wire16, skip_self_ids;
    const uint16_t* src16[8];
    const int32_t* src32[8];
    const int32_t* src_begins[8];
    const int32_t* src_ends[8];
    const int64_t* src_total[8];
    int32_t *ids, *begins, *ends;          // this rank's gathered result: [world * slot_capacity], [world * rows_per_rank]
    int64_t slot_capacity, rows_per_rank;
};
constexpr int kPullTile = 8192;            // ids per tile: 256 threads x 4 x 16-byte loads in flight (16-bit wire)

__device__ __forceinline__ uint4 ld_peer16(const void* p) {      // L2-coherent 16-byte load (peer memory is never cached in L1)
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

// i32 compact ids -> u16 staging (local pass right after the tokenizer: most of the ids are still in L2)
__global__ void peer_pack_kernel(const int32_t* __restrict__ ids, const int64_t* __restrict__ n_dev, int64_t capacity, uint16_t* __restrict__ out) {
    int64_t n = *n_dev;
    if (n > capacity) n = capacity;
    const int64_t n8 = (n + 7) >> 3;       // groups of 8 ids: two 16-byte loads, one 16-byte store (buffers are padded to a multiple of 8)
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n8; g += (int64_t)gridDim.x * blockDim.x) {
        const int4 a = __ldcs(reinterpret_cast<const int4*>(ids) + 2 * g), b = __ldcs(reinterpret_cast<const int4*>(ids) + 2 * g + 1);
        uint4 o;
        o.x = (uint32_t)(a.x & 0xFFFF) | ((uint32_t)a.y << 16); o.y = (uint32_t)(a.z & 0xFFFF) | ((uint32_t)a.w << 16);
        o.z = (uint32_t)(b.x & 0xFFFF) | ((uint32_t)b.y << 16); o.w = (uint32_t)(b.z & 0xFFFF) | ((uint32_t)b.w << 16);
        reinterpret_cast<uint4*>(out)[g] = o;
    }
}

__global__ void __launch_bounds__(256) peer_pull_kernel(const PeerPull Q) {
    __shared__ int64_t tot[8];
    if (threadIdx.x < Q.world) {
        int64_t t = *reinterpret_cast<const volatile int64_t*>(Q.src_total[threadIdx.x]);
        tot[threadIdx.x] = t < 0 ? 0 : (t > Q.slot_capacity ? Q.slot_capacity : t);
    }
    __syncthreads();
    // row extents of every slot (offsets shifted into the slot): B x 8 bytes per rank, spread over the grid
    {
        const int64_t R = Q.rows_per_rank, total_rows = R * Q.world;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_rows; i += (int64_t)gridDim.x * blockDim.x) {
            const int p = (int)(i / R);
            const int64_t r = i - (int64_t)p * R;
            const int32_t shift = (int32_t)((int64_t)p * Q.slot_capacity);
            Q.begins[i] = __ldcg(Q.src_begins[p] + r) + shift;
            Q.ends[i] = __ldcg(Q.src_ends[p] + r) + shift;
        }
    }
    // ids: tiles interleaved over the peers (consecutive CTAs read from different GPUs), starting with the next rank
    const int npeer = Q.skip_self_ids ? Q.world - 1 : Q.world;
    if (npeer <= 0) return;
    int64_t maxt = 0;
    for (int p = 0; p < Q.world; ++p) { const int64_t t = (tot[p] + kPullTile - 1) / kPullTile; maxt = t > maxt ? t : maxt; }
    for (int64_t w = blockIdx.x; w < maxt * npeer; w += gridDim.x) {
        const int j = (int)(w % npeer);
        const int64_t tile = w / npeer;
        const int p = (Q.rank + 1 + j) % Q.world;          // j = world - 1 is this rank itself (only without skip_self_ids)
        const int64_t lo = tile * kPullTile;
        if (lo >= tot[p]) continue;
        int32_t* const dst = Q.ids + (int64_t)p * Q.slot_capacity + lo;
        const int64_t left = tot[p] - lo;                  // ids of this tile (> 0); groups beyond it are not touched
        if (Q.wire16) {
            const uint16_t* const src = Q.src16[p] + lo;
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int g = threadIdx.x + 256 * u; if ((int64_t)g * 8 < left) v[u] = ld_peer16(src + 8 * g); }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int g = threadIdx.x + 256 * u;
                if ((int64_t)g * 8 < left) {
                    int4 a, b;
                    a.x = (int32_t)(v[u].x & 0xFFFFu); a.y = (int32_t)(v[u].x >> 16); a.z = (int32_t)(v[u].y & 0xFFFFu); a.w = (int32_t)(v[u].y >> 16);
                    b.x = (int32_t)(v[u].z & 0xFFFFu); b.y = (int32_t)(v[u].z >> 16); b.z = (int32_t)(v[u].w & 0xFFFFu); b.w = (int32_t)(v[u].w >> 16);
                    if ((int64_t)g * 8 + 8 <= left) {
                        __stcs(reinterpret_cast<int4*>(dst + 8 * g), a);
                        __stcs(reinterpret_cast<int4*>(dst + 8 * g) + 1, b);
                    } else {                                // the slot's last, partial group: never write past the row data (the next slot starts there)
                        const int32_t e[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                        for (int t = 0; t < 8; ++t) if ((int64_t)g * 8 + t < left) dst[8 * g + t] = e[t];
                    }
                }
            }
        } else {
            const int32_t* const src = Q.src32[p] + lo;
#pragma unroll
            for (int h = 0; h < 2; ++h) {                   // 8192 ids = 2 x (256 threads x 4 x 4 ids)
                uint4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { const int g = threadIdx.x + 256 * (u + 4 * h); if ((int64_t)g * 4 < left) v[u] = ld_peer16(src + 4 * g); }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int g = threadIdx.x + 256 * (u + 4 * h);
                    if ((int64_t)g * 4 + 4 <= left) __stcs(reinterpret_cast<uint4*>(dst + 4 * g), v[u]);
                    else if ((int64_t)g * 4 < left) {
                        const uint32_t e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
                        for (int t = 0; t < 4; ++t) if ((int64_t)g * 4 + t < left) dst[4 * g + t] = (int32_t)e[t];
                    }
                }
            }
        }
    }
}

}  // namespace b200tok
