// kernels_fast.cuh — the dedicated kernel of the headline path: RegexSplit(GPT-2 byte-level pattern, isolate) ->
// BPETokenizer (reference src/regex_split.cpp:287-309 + src/bpe_tokenizer.cpp:196-339) for rows whose pieces fit a
// 512-byte window and symbolise byte by byte.  Same work decomposition as rows_kernel (one warp per row, 512-byte
// windows, row-local output slots), but only the bit-mask formulation lives here, so the register budget is spent on
// it alone.  Anything it cannot finish exactly — a window with a multi-byte symbol, a dropped byte, a piece longer
// than a window, a skip-flagged element — makes the warp put the row on a redo list that rows_kernel<OP_BPE>
// (kernels.cuh) processes afterwards from scratch.
//
// Per window:
//   pass 1  position-parallel: byte -> lut32 (class bits | one-byte symbol id); ids[] stored; one warp ballot per class
//           gives the class masks of 32 positions (lane 0 parks them in shared memory); mergeable (previous, this)
//           byte pairs are found with a 2 KB bitmap and only those fetch their rank (initial key) from the L2 table
//   pass 2  one WORD per lane: tok_core.cuh g2_starts evaluates the piece-start predicate for 32 positions at once
//   pass 3  segments owning a mergeable pair: carry chain over the bit-reversed masks; two-symbol segments finish
//           here, longer ones go to a queue
//   pass 4  merge queue, one segment per lane per round; live symbols / live keys are two 32-bit masks
//   emit    position-parallel ballot compaction of the surviving ids into the row slot
#pragma once
#include "kernels.cuh"

namespace b200tok {

constexpr int kMStride = NWORDS + 4;   // pass 1 writes four words per iteration

constexpr int kRawBytes = 16 + LBK + WBYTES + 16;      // one staged window incl. up to 15 bytes of source misalignment on either side
template <class IdT>
struct __align__(16) FastSmem {
    alignas(16) uint8_t raw_pf[2][kRawBytes];        // staged window bytes; two buffers: the slot path prefetches the next window with a TMA bulk copy
    alignas(8) unsigned long long mbar[2];           // one mbarrier per buffer: the bulk copy completes its transaction bytes on it
    int32_t pf_row, pf_rb, pf_re, pf_eb, pf_ee;      // prefetch state: next row and its metadata (filled asynchronously by cp.async)
    int32_t pf_pos, pf_off, pf_n;                    // window the in-flight buffer holds: start byte (-1: none), offset of position 0 in the buffer, bytes copied
    uint32_t segbits[NWORDS], actbits[NWORDS];
    uint32_t m[11][kMStride];                // class masks per word, index = word + 1 (word -1 = look-back): L N S SP A2 A3 CONT MB F NL PG
    IdT ids[WIN];                            // symbol per position; kDead = merged away
    alignas(16) uint32_t key[WIN + 4];       // merge keys; after the merges: staging of the window's compacted ids for wide peer stores
    int32_t pend_row[4], pend_cnt[4], pend_off[4];   // in-order emit: rows tokenised but not yet written out (row | redo << 31, count, ring offset)
    __device__ __forceinline__ uint8_t* B0() { return raw_pf[0] + 16 + LBK; }      // position 0 of a synchronously staged window
    __device__ __forceinline__ uint16_t* act() { return reinterpret_cast<uint16_t*>(&m[0][0]); }   // merge queue; the masks are consumed by then
    static constexpr IdT kDead = (IdT)-1;
};
static_assert(sizeof(uint32_t) * 11 * kMStride >= sizeof(uint16_t) * (WIN / 3 + 10), "merge queue must fit the mask area");
enum : int { M_L = 0, M_N, M_S, M_SP, M_A2, M_A3, M_CONT, M_MB, M_F, M_NL, M_PG };

constexpr int kFastSmemFixed = 128 + 1024;   // ascii classes, lut32
template <class IdT> constexpr size_t fast_smem_bytes() { return kFastSmemFixed + WARPS_PER_BLOCK * sizeof(FastSmem<IdT>); }

// One window.  Returns `send` (how far the window advances; 0 => the first piece does not fit) or -1 when the window
// needs the generic path.  On success S.ids[0 .. send) holds the window's tokens (-1 = merged away).
struct NoHook {       // fast_window calls hook.after_pass1() / after_pass3(): places where a caller can slip in latency-bound work
    __device__ __forceinline__ void after_pass1() {}
    __device__ __forceinline__ void after_pass3() {}
};
// ASCII is a template parameter, not a flag: the all-ASCII and the general form of a window are two straight-line instruction streams
// (fast_window below picks one per window), so a workload runs in the instruction-cache footprint of the one it needs — the kernel is
// ~80 KB of code and `no_instruction` was its top stall on mixed UTF-8 (profiles/r02_fast_kernel_c3_ncu_full.txt).
template <class IdT, bool L3, bool ASCII, class Hook>
__device__ __forceinline__ int fast_window_t(FastSmem<IdT>& S, const RowParams& P, const uint32_t* lut32,
                                             const uint8_t* ascii_smem, int lane, int wlen, int end_rel, int nload, int off,
                                             const uint8_t* B, Hook& hook) {
    constexpr bool ascii = ASCII;
    const uint32_t lt = (1u << lane) - 1u;
    const int lb = off < LBK ? off : LBK;           // look-back bytes staged before the window (off = window start - element start)
    ClassTables T = P.cls;
    T.ascii = ascii_smem;
    // ---- non-ASCII windows: one class byte per position (continuation bytes carry their owner's class plus C_CONT), parked in
    // the bytes of the ids array that pass 1 has not written yet: pass 1 walks the words from the top down and reads the class
    // bytes of an iteration before it stores that iteration's ids, so a class byte is always read before it is overwritten
    // (the class byte of position q lives at byte q + LBK, the id of position p at byte sizeof(IdT) * p >= p + LBK for p >= LBK).
    uint8_t* const KC = reinterpret_cast<uint8_t*>(S.ids) + LBK;
    if (!ascii) {
        // characters are decoded once: ASCII bytes take their class from the table right away; the first bytes of multi-byte
        // characters are first gathered into a dense list (in the not-yet-used key array) so that the UTF-8 decode + two-stage
        // class lookup then runs with every lane busy instead of the few lanes that happen to sit on a lead byte
        uint16_t* const leads = reinterpret_cast<uint16_t*>(S.key);
        int n_leads = 0;
        for (int w0 = -lb; w0 < nload; w0 += 32) {
            const int w = w0 + lane;
            const uint8_t b = w < nload ? B[w] : 0;
            if (w < nload && b < 0x80) KC[w] = ascii_smem[b];
            const bool lead = w < nload && b >= 0x80 && (!is_cont_byte(b) || w == -lb);
            const uint32_t m = __ballot_sync(FULL, lead);
            if (lead) leads[n_leads + __popc(m & lt)] = (uint16_t)(w + LBK);
            n_leads += __popc(m);
        }
        __syncwarp();
        for (int i = lane; i < n_leads; i += 32) {
            const int w = (int)leads[i] - LBK;
            KC[w] = char_class(B, w, end_rel, T);
        }
        __syncwarp();
        for (int w = lane - lb; w < nload; w += 32) {            // continuation bytes copy their owner
            const uint8_t b = B[w];
            if (is_cont_byte(b) && w > -lb) {
                int j = w - 1;                                    // first byte of the character: at most three bytes back
                if (j > -lb && is_cont_byte(B[j])) { --j; if (j > -lb && is_cont_byte(B[j])) --j; }
                uint8_t k = C_CONT;
                if (B[j] >= 0xC0) k |= KC[j] & (uint8_t)~C_CONT;
                KC[w] = k;
            }
        }
        __syncwarp();
    }
    // ---- pass 1: four words (stride 32) per iteration, top word first ----
    bool complex = false;
    const int nw = (nload + 31) >> 5;
    {
        const int it0 = lb > 0 ? -1 : 0;
        if (it0 == 0 && lane < 11) S.m[lane][0] = 0u;
        const uint32_t span = (uint32_t)(nload + lb);
        const uint32_t* const pair_rank = P.bpe.pair_rank;
        const uint16_t* const rank16 = reinterpret_cast<const uint16_t*>(P.bpe.pair_bits + 2560);
        for (int it = it0 + ((nw - 1 - it0) & ~3); it >= it0; it -= 4) {
            const int w0 = it * 32 + lane;
            const uint8_t* bq = B + w0;                           // this lane's byte of word `it`
            uint32_t c[4], g[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                c[u] = bq[32 * u];
                g[u] = lut32[c[u]];
                const bool valid = (uint32_t)(w0 + 32 * u + lb) < span;
                if (!ascii && valid) g[u] |= KC[w0 + 32 * u] & (uint32_t)(C_L | C_N | C_S | C_CONT);
                g[u] = valid ? g[u] : 0u;
            }
            uint32_t* const mm = &S.m[0][it + 1];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t bL = ballot_bits(g[u], V7_L), bN = ballot_bits(g[u], V7_N), bS = ballot_bits(g[u], V7_S), bSP = ballot_bits(g[u], V7_SP);
                if (lane == 0) { mm[M_L * kMStride + u] = bL; mm[M_N * kMStride + u] = bN; mm[M_S * kMStride + u] = bS; mm[M_SP * kMStride + u] = bSP; }
                if (L3) { const uint32_t bNL = ballot_bits(g[u], V7_NL); if (lane == 0) mm[M_NL * kMStride + u] = bNL; }
            }
            const uint32_t gor = g[0] | g[1] | g[2] | g[3];
            if (ballot_bits(gor, V7_AP)) {
                // apostrophes are sparse: only the words that hold one evaluate the contraction forms (one ballot decides per word)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    uint32_t bA2 = 0u, bA3 = 0u;
                    if (ballot_bits(g[u], V7_AP)) {
                        int cl = 0;
                        if (g[u] & V7_AP) cl = L3 ? llama3_contraction_len(B, w0 + 32 * u, nload) : gpt2_contraction_len(B, w0 + 32 * u, nload);
                        bA2 = __ballot_sync(FULL, cl == 2); bA3 = __ballot_sync(FULL, cl == 3);
                    }
                    if (lane == 0) { mm[M_A2 * kMStride + u] = bA2; mm[M_A3 * kMStride + u] = bA3; }
                }
            } else if (lane < 4) { mm[M_A2 * kMStride + lane] = 0u; mm[M_A3 * kMStride + lane] = 0u; }
            if (!ascii) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t bC = ballot_bits(g[u], V7_CONT);
                    bool mb = false;
                    if ((g[u] & V7_S) && c[u] >= 0x80 && !(g[u] & V7_CONT)) {      // multi-byte whitespace: is the next character a non-space?
                        int j = w0 + 32 * u + 1;
                        while (j < nload && is_cont_byte(B[j])) ++j;
                        mb = j < nload && !(KC[j] & C_S);
                    }
                    const uint32_t bMB = __ballot_sync(FULL, mb);
                    if (lane == 0) { mm[M_CONT * kMStride + u] = bC; mm[M_MB * kMStride + u] = bMB; }
                    if (L3) {
                        const int w = w0 + 32 * u;
                        if ((g[u] & V7_N) && c[u] >= 0x80) complex = true;          // digit groups are evaluated for ASCII digits only
                        bool pg = false;                                           // letter after a multi-byte "other" char at which a match starts
                        // (the previous byte carries its character's class: only a multi-byte OTHER char needs the look-back)
                        if ((g[u] & V7_L) && !(g[u] & V7_CONT) && w > -lb && (KC[w - 1] & (C_CONT | C_L | C_N | C_S)) == C_CONT) {
                            int j = w - 1;                       // first byte of that character: at most three bytes back
                            if (j > -lb && is_cont_byte(B[j])) { --j; if (j > -lb && is_cont_byte(B[j])) { --j; if (j > -lb && is_cont_byte(B[j])) --j; } }
                            if (!(KC[j] & (C_L | C_N | C_S)))
                                pg = j == -lb || (B[j - 1] != 0x20 && (KC[j - 1] & (C_L | C_N | C_S)));
                        }
                        const uint32_t bPG = __ballot_sync(FULL, pg);
                        if (lane == 0) mm[M_PG * kMStride + u] = bPG;
                    }
                }
            }
            if (ballot_bits(gor, V7_WALK | V7_BAD)) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (!ballot_bits(g[u], V7_WALK | V7_BAD)) continue;      // (sparse, like the apostrophes)
                    const int w = w0 + 32 * u;
                    if (g[u] & V7_BAD) complex = true;
                    else if ((g[u] & V7_WALK) && w + 1 < nload) {        // could a longer token start here?  second-byte filter, then the walk
                        const uint32_t b1 = bq[32 * u + 1];
                        if ((__ldg(P.bpe.pair_bits + 512 + 8 * c[u] + (b1 >> 5)) >> (b1 & 31u)) & 1u) {
                            int j = w;
                            const int32_t t = trie_longest(P.bpe.trie, B, j, nload);   // (even across a piece boundary -> generic path)
                            if (t >= 0 && j != w + 1) complex = true;
                        }
                    }
                }
            }
            // mergeable (previous byte, this byte) pairs: ASCII pairs from the compact u16 rank table, others from the full one
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int w = w0 + 32 * u;
                bool fb = false;
                if ((uint32_t)w < (uint32_t)wlen) {
                    const uint32_t p = bq[32 * u - 1];
                    uint32_t r;
                    if (ascii || (p | c[u]) < 0x80u) {
                        r = __ldg(rank16 + ((p << 7) | c[u]));
                        if (r >= 0xFFFEu) r = r == 0xFFFFu ? kNoKey : __ldg(pair_rank + ((p << 8) | c[u]));
                    } else r = __ldg(pair_rank + ((p << 8) | c[u]));
                    fb = r != kNoKey;
                    S.key[w] = (r << kPackedBirthBits) | (uint32_t)w;      // (read only where fb: no branch around the store)
                }
                const uint32_t bF = __ballot_sync(FULL, fb);
                if (lane == 0) mm[M_F * kMStride + u] = bF;
            }
            if (!ascii) __syncwarp();                            // every class byte of this iteration has been read
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if ((uint32_t)(w0 + 32 * u) < (uint32_t)wlen) S.ids[w0 + 32 * u] = (IdT)(g[u] >> V7_ID_SHIFT);
        }
    }
    hook.after_pass1();
    if (__any_sync(FULL, complex)) return -1;
    __syncwarp();
    // ---- pass 2: piece starts, lane = word ----
    const int word = lane, base = lane * 32;
    const bool live = word < nw;
    uint32_t found = live ? S.m[M_F][word + 1] : 0u;
    const uint32_t bos = (word == 0 && off == 0) ? 1u : 0u;
    uint32_t start;
    if (!L3) {
        G2Word W{0, 0, 0, 0, 0, 0, 0, 0, 0}, PW{0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (live) {
            W.L = S.m[M_L][word + 1]; W.N = S.m[M_N][word + 1]; W.S = S.m[M_S][word + 1]; W.SP = S.m[M_SP][word + 1];
            W.A2 = S.m[M_A2][word + 1]; W.A3 = S.m[M_A3][word + 1];
            PW.L = S.m[M_L][word]; PW.N = S.m[M_N][word]; PW.S = S.m[M_S][word]; PW.SP = S.m[M_SP][word];
            if (!ascii) { W.CONT = S.m[M_CONT][word + 1]; W.MB = S.m[M_MB][word + 1]; }
            W.X = v7_below(nload, base);
        }
        uint32_t c2, c3;
        g2_contractions(W, g2_ok1(PW), bos, c2, c3);
        uint32_t pc2 = __shfl_up_sync(FULL, c2, 1), pc3 = __shfl_up_sync(FULL, c3, 1);
        if (lane == 0) {      // contractions starting in the look-back word (their apostrophe is at most 3 positions back)
            pc2 = 0; pc3 = 0;
            if (lb > 0) {
                G2Word LBW{PW.L, PW.N, PW.S, PW.SP, S.m[M_A2][0], S.m[M_A3][0], 0, 0, 0};
                g2_contractions(LBW, 0u, off <= LBK ? 1u << (32 - off) : 0u, pc2, pc3);   // (element start inside the look-back)
            }
        }
        const uint32_t nns = __shfl_down_sync(FULL, W.X & ~W.S, 1);
        start = g2_starts(W, PW, c2, c3, pc2, pc3, lane == 31 ? 0u : nns, bos, P.spec.pat == PAT_GPT2_DIGITS);
    } else {
        L3Word W{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, PW{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (live) {
            W.L = S.m[M_L][word + 1]; W.N = S.m[M_N][word + 1]; W.S = S.m[M_S][word + 1]; W.SP = S.m[M_SP][word + 1];
            W.NL = S.m[M_NL][word + 1]; W.A2 = S.m[M_A2][word + 1]; W.A3 = S.m[M_A3][word + 1];
            PW.L = S.m[M_L][word]; PW.N = S.m[M_N][word]; PW.S = S.m[M_S][word]; PW.SP = S.m[M_SP][word]; PW.NL = S.m[M_NL][word];
            if (!ascii) { W.CONT = S.m[M_CONT][word + 1]; W.MB = S.m[M_MB][word + 1]; W.PG = S.m[M_PG][word + 1]; }
            W.X = v7_below(nload, base);
            PW.X = word > 0 ? FULL : (lb > 0 ? FULL << (32 - lb) : 0u);
        }
        const uint32_t mso = l3_mso(W, PW), d2 = W.A2 & mso, d3 = W.A3 & mso;
        uint32_t p_mso = __shfl_up_sync(FULL, mso, 1), pd2 = __shfl_up_sync(FULL, d2, 1), pd3 = __shfl_up_sync(FULL, d3, 1);
        if (lane == 0) {      // the look-back word's own contribution (its predecessor is unknown: treated as empty)
            p_mso = 0; pd2 = 0; pd3 = 0;
            if (lb > 0) {
                L3Word LBW = PW, Z{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
                LBW.A2 = S.m[M_A2][0]; LBW.A3 = S.m[M_A3][0];
                if (!ascii) LBW.CONT = S.m[M_CONT][0];
                p_mso = l3_mso(LBW, Z); pd2 = LBW.A2 & p_mso; pd3 = LBW.A3 & p_mso;
            }
        }
        // newline runs right after an other char belong to its piece: fill them upwards, carrying across words
        uint32_t lead;
        {
            const uint32_t seeds = l3_lead_seeds(W, PW);
            lead = l3_fill_up(W.NL, seeds);
            const uint32_t G = __ballot_sync(FULL, (lead >> 31) != 0u), Pm = __ballot_sync(FULL, W.NL == FULL);
            const uint32_t np = ~Pm & lt;
            const int j = np ? 31 - __clz(np) : 0;
            if ((G & lt & ~((1u << j) - 1u)) && (W.NL & 1u)) lead = l3_fill_up(W.NL, seeds | 1u);
        }
        uint32_t p_lead = __shfl_up_sync(FULL, lead, 1);
        if (lane == 0) p_lead = 0;      // (a window never starts inside such a run: its newlines are not piece starts)
        // the tail of every whitespace run (non-newline chars after its last newline, up to a non-space): fill downwards
        uint32_t tail;
        {
            const uint32_t nS0 = __shfl_down_sync(FULL, W.S, 1);
            const uint32_t Tb = W.S & ~W.NL, seeds = l3_tail_seeds(W, lane == 31 ? 0u : nS0);
            tail = l3_fill_down(Tb, seeds);
            const uint32_t G = __ballot_sync(FULL, (tail & 1u) != 0u), Pm = __ballot_sync(FULL, Tb == FULL);
            const uint32_t above = lane == 31 ? 0u : (FULL << (lane + 1));
            const uint32_t np = ~Pm & above;
            const int j = np ? __ffs(np) - 1 : 31;
            if ((G & above & (FULL >> (31 - j))) && (Tb >> 31)) tail = l3_fill_down(Tb, seeds | 0x80000000u);
        }
        // every third digit from the start of its run; a run entering the word from below needs its digit count so far
        int phase = 0;
        if (word > 0 && (W.N & 1u) && (PW.N >> 31))
            for (int p = base - 1; p >= 0 && (lut32[B[p]] & V7_N); --p) ++phase;
        const uint32_t nst = l3_number_starts(W.N, PW.N >> 31, phase, word == 0);
        const uint32_t nns = __shfl_down_sync(FULL, W.X & ~W.S, 1);
        start = l3_starts(W, PW, mso, p_mso, d2, d3, pd2, pd3, lead, p_lead, tail, nst, lane == 31 ? 0u : nns, bos);
        if (nload < end_rel) {      // the loaded bytes end inside the element: a whitespace run touching their end is undecided
            const int lastp = nload - 1;
            if ((S.m[M_S][(lastp >> 5) + 1] >> (lastp & 31)) & 1u) {
                int a = 0;
                for (int wi = lastp >> 5, top = lastp & 31; wi >= 0; --wi, top = 31) {
                    const uint32_t zeros = ~S.m[M_S][wi + 1] & (top == 31 ? FULL : ((2u << top) - 1u));
                    if (zeros) { a = wi * 32 + 32 - __clz(zeros); break; }
                }
                start &= v7_below(a + 1, base);      // keep the run's first piece start at most; the rest is redone in the next window
            }
        }
    }
    start &= v7_below(wlen, base);
    if (word == 0) start |= 1u;
    int send = wlen;
    if (wlen != end_rel) {        // the last piece may continue beyond the window: redo it from its start in the next window
        const int hb = start ? base + 31 - __clz(start) : -1;
        send = __reduce_max_sync(FULL, hb);
        start &= v7_below(send, base);
        if (send <= 0) return 0;
    }
    found &= ~start & v7_below(send, base);
    if (word < NWORDS) { S.segbits[word] = start; S.actbits[word] = found; }
    // ---- pass 3: segments that own a mergeable pair.  In the bit-reversed word a segment's start is its top bit; adding the
    // found bits to "all non-start bits" carries every found bit up to (and only to) its segment's start.
    const uint32_t sr = __brev(start), fr = __brev(found), kk = ~sr;
    const bool gen = (uint32_t)(fr + kk) < fr;                          // carry out of the word (towards lower positions)
    const uint32_t gm = __ballot_sync(FULL, gen), pm = __ballot_sync(FULL, start == 0u);
    uint32_t cin = 0;
    if (lane < 31) {                                                    // carry in = a generating word above, reached through start-less words
        const uint32_t up = gm >> (lane + 1), pr = pm >> (lane + 1);
        const int t = __ffs(~pr) - 1;
        cin = (up & (FULL >> (31 - t))) != 0u;
    }
    const uint32_t actst = __brev((fr + kk + cin) & sr);
    uint32_t stx = start;                                               // starts + the end sentinel at `send`
    if ((send >> 5) == word) stx |= 1u << (send & 31);
    const uint32_t nstx0 = __shfl_down_sync(FULL, stx, 1);
    const uint32_t nstx = lane == 31 ? 0u : nstx0;
    const uint32_t len2 = start & ~g2_fsr(stx, nstx, 1) & g2_fsr(stx, nstx, 2);
    const int32_t* const rank_newid = P.bpe.merges.rank_newid;
    __syncwarp();
    const int32_t nb = P.bpe.newid_base;
    for (uint32_t a = actst & len2; a; a &= a - 1u) {                   // two symbols, one pair: done here
        const int s = base + __ffs(a) - 1;
        const uint32_t r = S.key[s + 1] >> kPackedBirthBits;
        S.ids[s] = (IdT)(nb >= 0 ? nb + (int32_t)r : __ldg(rank_newid + r));
        S.ids[s + 1] = S.kDead;
    }
    const uint32_t a3 = actst & ~len2;
    const int cnt = __popc(a3);
    const int incl = warp_incl_scan(cnt, lane);
    const int nact = __shfl_sync(FULL, incl, 31);
    {
        int off = incl - cnt;
        uint16_t* act = S.act();
        for (uint32_t a = a3; a; a &= a - 1u) act[off++] = (uint16_t)(base + __ffs(a) - 1);
    }
    __syncwarp();
    hook.after_pass3();
    // ---- pass 4: merge queue, one segment per lane, one merge per iteration ----
    const MergeTable MT = P.bpe.merges;
    int head = 0, s = 0, merges = 0;
    uint32_t alive = 0, km = 0;
    bool have = false;
    while (head < nact || __any_sync(FULL, have)) {
        const uint32_t need = __ballot_sync(FULL, !have);
        if (!have) {
            const int qi = head + __popc(need & lt);
            if (qi < nact) {
                s = S.act()[qi];
                const int e = next_bit(S.segbits, s, send), n = e - s;
                if (n > 32) {                                           // a run longer than the 32-bit masks: serial loop over the whole segment
                    const int c = bpe_merge_packed(MT, S.ids + s, S.key + s, n, MT.tie_check ? &complex : nullptr);
                    for (int t = s + c; t < e; ++t) S.ids[t] = S.kDead;
                } else {
                    const uint32_t mask = n == 32 ? FULL : ((1u << n) - 1u);
                    km = g2_fsr(S.actbits[s >> 5], S.actbits[(s >> 5) + 1], s & 31) & mask;
                    alive = mask;
                    merges = 0;
                    have = true;
                }
            }
        }
        head += __popc(need);
        if (have) {
            uint32_t best = kNoKey;
            int bk = 0;
            for (uint32_t m = km; m; m &= m - 1u) {
                const int k = __ffs(m) - 1;
                const uint32_t q = S.key[s + k];
                if (q < best) { best = q; bk = k; }
            }
            const int pl = 31 - __clz(alive & ((1u << bk) - 1u));       // left operand = nearest live symbol below
            const int32_t nid = nb >= 0 ? nb + (int32_t)(best >> kPackedBirthBits) : __ldg(rank_newid + (best >> kPackedBirthBits));
            S.ids[s + pl] = (IdT)nid;
            S.ids[s + bk] = S.kDead;
            alive &= ~(1u << bk);
            km &= ~((1u << bk) | (1u << pl));
            ++merges;
            const uint32_t birth = (uint32_t)(WIN + merges);
            const uint32_t below = alive & ((1u << pl) - 1u);
            bool fl = false;
            if (below) {
                int32_t r, v;
                fl = merge_find(MT, (int32_t)S.ids[s + 31 - __clz(below)], nid, r, v);
                if (fl) { S.key[s + pl] = ((uint32_t)r << kPackedBirthBits) | birth; km |= 1u << pl; }
            }
            const uint32_t above = alive & ~((2u << bk) - 1u);
            if (above) {
                const int nr = __ffs(above) - 1;
                km &= ~(1u << nr);
                int32_t r, v;
                if (merge_find(MT, nid, (int32_t)S.ids[s + nr], r, v)) {
                    S.key[s + nr] = ((uint32_t)r << kPackedBirthBits) | birth; km |= 1u << nr;
                    // the merge found its own product on both sides: the two new pairs tie on (rank, seq) and the reference pops them in
                    // heap order — the row goes to the exact path (only vocabularies with doubly produced tokens get here)
                    if (MT.tie_check && fl && (int32_t)S.ids[s + 31 - __clz(below)] == nid && (int32_t)S.ids[s + nr] == nid) complex = true;
                }
            }
            if (!km) have = false;
        }
    }
    __syncwarp();
    if (__any_sync(FULL, complex)) return -1;
    return send;
}


template <class IdT, bool L3, class Hook>
__device__ __forceinline__ int fast_window(FastSmem<IdT>& S, const RowParams& P, const uint32_t* lut32,
                                           const uint8_t* ascii_smem, int lane, int wlen, int end_rel, int nload, int off,
                                           bool ascii, const uint8_t* B, Hook& hook) {
    return ascii ? fast_window_t<IdT, L3, true, Hook>(S, P, lut32, ascii_smem, lane, wlen, end_rel, nload, off, B, hook)
                 : fast_window_t<IdT, L3, false, Hook>(S, P, lut32, ascii_smem, lane, wlen, end_rel, nload, off, B, hook);
}

// Stage the bytes of one window (+ look-back / look-ahead) into shared memory; returns true if every byte is ASCII.
template <class IdT>
__device__ __forceinline__ bool stage_window(FastSmem<IdT>& S, const RowParams& P, int pos, int lb, int nload, int lane, uint8_t* Bw) {
    uint32_t hibits = 0;
    const uint8_t* src = P.chars + pos - lb;
    if (((reinterpret_cast<uintptr_t>(src) | (uintptr_t)lb) & 15) == 0) {
        // 16-byte vector loads; the last quad may read up to 15 bytes past the element (padded chars allocation)
        uint4* dst = reinterpret_cast<uint4*>(Bw - lb);
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
            Bw[w] = bb;
            hibits |= bb;
        }
    }
    const bool all_ascii = !__any_sync(FULL, hibits & 0x80808080u);
    __syncwarp();
    return all_ascii;
}

// Compact the live ids of S.ids[0 .. send) into dst[0 ..) (global staging ring).  Returns the number of live ids.
template <class IdT>
__device__ __forceinline__ int compact_window(FastSmem<IdT>& S, int send, int lane, IdT* __restrict__ dst) {
    const uint32_t ltm = (1u << lane) - 1u;
    int n_out = 0;
    for (int w = lane; w - lane < send; w += 128) {
        IdT tok[4];
        uint32_t m[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) tok[u] = (w + 32 * u) < send ? S.ids[w + 32 * u] : S.kDead;
#pragma unroll
        for (int u = 0; u < 4; ++u) m[u] = __ballot_sync(FULL, tok[u] != S.kDead);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (tok[u] != S.kDead) dst[n_out + __popc(m[u] & ltm)] = tok[u];
            n_out += __popc(m[u]);
        }
    }
    return n_out;
}

// The row loop of the in-order single-pass emit (OrderedOut, kernels.cuh).  A warp tokenises a row into its private staging ring
// (global memory, L2-resident), publishes the row's count at once and goes on to the next row; up to four rows wait in the ring
// until the look-back over the counts of all earlier rows yields their output offset, then they are copied — widened to i32 — to
// their final place.  (Writing a row only when its offset is known but never WAITING for it while there is other work is what
// keeps the warps from marching in lock step behind the slowest one.)  A row the bit-mask path cannot finish reserves its worst
// case (count <= bytes, src/bpe_tokenizer.cpp:135) and goes on the redo list; ordered_recompact_kernel closes those gaps.
constexpr int kPend = 4;
template <class IdT, bool L3>
__device__ __forceinline__ void ordered_rows(FastSmem<IdT>& S, const RowParams& P, const uint32_t* lut32_smem, const uint8_t* ascii_smem,
                                             int lane, int32_t* __restrict__ redo_rows) {
    const OrderedOut& O = P.oo;
    IdT* const ring = reinterpret_cast<IdT*>(O.stage) + (size_t)(blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5)) * (size_t)O.stage_cap;
    int n_pend = 0, head = 0, ring_used = 0;
    int lb_idx = -1;             // look-back progress of the oldest waiting row (-1: not started)
    long long lb_acc = 0;

    // Write out the oldest waiting row if its offset can be resolved (blocking: wait until it can).  Warp-uniform.
    auto retire = [&](bool blocking) -> bool {
        const int prow = S.pend_row[head];
        const int row = prow & 0x7FFFFFFF;
        const bool redo = prow < 0;
        const int count = S.pend_cnt[head], off = S.pend_off[head];
        long long excl = 0;
        if (row > 0) {
            if (lb_idx < 0) { lb_idx = row - 1; lb_acc = 0; }
            for (int r; (r = lookback_poll(O.desc, O.epoch, 0, lb_idx, lb_acc, lane)) != 1;) {
                if (r < 0) {                 // a predecessor is still being tokenised
                    if (!blocking) return false;
                    __nanosleep(256);
                }
            }
            excl = lb_acc;
            if (lane == 0) desc_store(O.desc + row, ((unsigned long long)O.epoch << 34) | kDescPrefix | (uint32_t)(excl + count));
        }
        lb_idx = -1;
        const bool fits = excl + (long long)count <= O.cap;
        if (!fits && lane == 0) atomicOr(&P.status[ST_ERROR], ERR_TMP_OVERFLOW);
        if (redo) {
            if (lane == 0) {
                redo_rows[atomicAdd(&P.status[ST_NREDO], 1)] = row;
                atomicMax(&P.status[ST_MINREDO], P.n_rows - row);
                const_cast<int32_t*>(P.row_base)[row] = (int32_t)excl;       // the generic kernels fill tmp_a[excl ..)
                O.begins[row] = (int32_t)excl; O.ends[row] = (int32_t)excl;
                P.row_flag[row] = 2;
            }
        } else {
            if (fits) {
                int32_t* const outp = O.ids + excl;
                const IdT* const src = ring + off;
                for (int t = lane; t < count; t += 128) {
                    IdT v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) v[u] = (t + 32 * u < count) ? __ldcg(src + t + 32 * u) : (IdT)0;
#pragma unroll
                    for (int u = 0; u < 4; ++u) if (t + 32 * u < count) __stcs(outp + t + 32 * u, (int32_t)v[u]);
                }
            }
            if (lane == 0) { O.begins[row] = (int32_t)excl; O.ends[row] = (int32_t)(excl + count); P.row_flag[row] = 0; }
        }
        if (row == P.n_rows - 1 && lane == 0) {
            P.status[ST_TOTAL] = (int32_t)(excl + count);
            if (O.total) *O.total = excl + count;
        }
        head = (head + 1) & (kPend - 1);
        if (--n_pend == 0) ring_used = 0;
        return true;
    };

    bool exhausted = false;
    for (;;) {
        // the one place rows are written out.  Forced (blocking) when the queue is full, the ring is more than half full, or no rows
        // are left — a warp never blocks while it holds a ticket, or the rows behind it would wait for a row nobody works on;
        // otherwise a row is tried once it has a successor in the queue: by then its predecessors have usually published.
        while (n_pend > 0) {
            const bool must = exhausted || n_pend == kPend || ring_used > (O.stage_cap >> 1);
            if (!must && n_pend < 2) break;
            if (!retire(must)) break;
        }
        if (exhausted) break;
        int row = 0;
        if (lane == 0) row = atomicAdd(&P.status[ST_TICKET], 1);
        row = __shfl_sync(FULL, row, 0);
        if (row >= P.n_rows) { exhausted = true; continue; }
        const int p0 = P.rb[row], p1 = P.re[row];
        long long need = 0;          // upper bound of the row's ids: one per byte (src/bpe_tokenizer.cpp:135)
        for (int p = p0 + lane; p < p1; p += 32) { const int l = P.ends[p] - P.begins[p]; need += l > 0 ? l : 0; }
#pragma unroll
        for (int o = 16; o; o >>= 1) need += __shfl_xor_sync(FULL, need, o);
        int emitted = 0;
        bool redo = ring_used + need > O.stage_cap;  // does not fit the ring (rows longer than half of it may not): the generic path takes the row
        for (int p = p0; p < p1 && !redo; ++p) {
            const int eb = P.begins[p], ee = P.ends[p];
            if (P.skips && P.skips[p]) { redo = true; break; }
            int pos = eb;
            while (pos < ee) {
                const int end_rel = ee - pos;
                const int wlen = end_rel < WIN ? end_rel : WIN;
                const int nload = end_rel < wlen + LA ? end_rel : wlen + LA;
                const int lb = (pos - eb) < LBK ? (pos - eb) : LBK;   // look-back bytes available inside the element
                const bool all_ascii = stage_window(S, P, pos, lb, nload, lane, S.B0());
                NoHook nh;
                const int send = fast_window<IdT, L3>(S, P, lut32_smem, ascii_smem, lane, wlen, end_rel, nload, pos - eb, all_ascii, S.B0(), nh);
                if (send <= 0) { redo = true; break; }
                emitted += compact_window(S, send, lane, ring + ring_used + emitted);
                __syncwarp();
                pos += send;
            }
        }
        const uint32_t count = redo ? (uint32_t)(need > 0x7FFFFFFF ? 0x7FFFFFFF : need) : (uint32_t)emitted;   // handed back: reserve the worst case
        if (lane == 0) {
            lookback_publish_count(O.desc, O.epoch, 0, row, count, 0);
            const int slot = (head + n_pend) & (kPend - 1);
            S.pend_row[slot] = row | (redo ? (int)0x80000000 : 0);
            S.pend_cnt[slot] = (int32_t)count;
            S.pend_off[slot] = ring_used;
        }
        __syncwarp();
        ++n_pend;
        if (!redo) ring_used += emitted;
    }
}


// ---- TMA bulk copy / mbarrier / cp.async wrappers (PTX ISA: cp.async.bulk, mbarrier, cp.async) ----------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    } while (!ok);
}
// global -> shared bulk copy by the TMA engine; completes `bytes` transaction bytes on the mbarrier (16-byte aligned, size % 16 == 0)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Metadata of the warp's NEXT row, fetched while the current window is being tokenised: fast_window calls the two hooks between its
// passes; each issues the loads whose addresses the previous step delivered (ticket -> rb/re -> begins/ends), as cp.async copies into
// shared memory, so no register is held across a pass and no pass waits for them.
template <class IdT>
struct RowPrefetch {
    FastSmem<IdT>& S;
    const RowParams& P;
    int lane;
    int t_next;          // lane 0: the next row's ticket (result of the atomicAdd issued when this row started)
    int stage;           // 0: nothing issued, 1: rb/re in flight, 2: begins/ends in flight (or not applicable)
    __device__ __forceinline__ void after_pass1() {
        if (stage != 0) return;
        const int nrow = __shfl_sync(FULL, t_next, 0);
        if (lane == 0) S.pf_row = nrow;
        if (nrow < P.n_rows) {
            if (lane == 0) cp_async4(&S.pf_rb, P.rb + nrow);
            if (lane == 1) cp_async4(&S.pf_re, P.re + nrow);
            cp_async_commit();
        }
        __syncwarp();
        stage = 1;
    }
    __device__ __forceinline__ void after_pass3() {
        if (stage != 1) return;
        cp_async_wait_all();
        __syncwarp();
        if (S.pf_row < P.n_rows && S.pf_re == S.pf_rb + 1) {      // one string per row (what the converter's add_ragged_dimension produces)
            if (lane == 0) cp_async4(&S.pf_eb, P.begins + S.pf_rb);
            if (lane == 1) cp_async4(&S.pf_ee, P.ends + S.pf_rb);
            cp_async_commit();
        }
        stage = 2;
    }
};

// Row loop of the slot path (one GPU): bump-allocated worst-case slot per row, ids stored 16 bits wide when they fit, and
// software-pipelined input: while a window is tokenised, the warp's next window — the next 512 bytes of the same string, or the
// first window of the row it will take next — is fetched into the second staging buffer by ONE cp.async.bulk (TMA) issued by
// lane 0 and signalled through an mbarrier, so a window never starts with the ticket -> offsets -> bytes chain of dependent loads.
template <class IdT, bool L3>
__device__ __forceinline__ void slot_rows(FastSmem<IdT>& S, const RowParams& P, const uint32_t* lut32_smem, const uint8_t* ascii_smem,
                                          int lane, int32_t* __restrict__ redo_rows) {
    const uint32_t ltm = (1u << lane) - 1u;
    if (lane == 0) {
        mbar_init(&S.mbar[0], 1);
        mbar_init(&S.mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        S.pf_pos = -1;
    }
    __syncwarp();
    int buf = 0;                 // staging buffer of the current window
    uint32_t par = 0;            // phase parity of the two mbarriers (bit b = next phase to wait for on mbar[b])
    bool inflight = false;       // a bulk copy into raw_pf[buf ^ 1] has been issued and not yet waited for
    bool meta_ok = false;        // S.pf_rb / pf_re / pf_eb / pf_ee describe `row`
    const bool can_prefetch = P.prefetch == 1 && P.skips == nullptr;
    int row = 0;
    if (lane == 0) row = atomicAdd(&P.status[ST_TICKET], 1);
    row = __shfl_sync(FULL, row, 0);
    while (row < P.n_rows) {
        RowPrefetch<IdT> hook{S, P, lane, 0, can_prefetch ? 0 : 3};           // (stage 3: hooks off — the next ticket is taken when this row is done)
        if (can_prefetch && lane == 0) hook.t_next = atomicAdd(&P.status[ST_TICKET], 1);      // looked at after pass 1 of the first window
        int p0, p1, eb0 = 0, ee0 = 0;
        if (meta_ok) { p0 = S.pf_rb; p1 = S.pf_re; eb0 = S.pf_eb; ee0 = S.pf_ee; }
        else { p0 = P.rb[row]; p1 = P.re[row]; }
        __syncwarp();            // (the hooks overwrite pf_* with the next row's metadata from here on)
        // the row's worst-case slot range (one id per byte, src/bpe_tokenizer.cpp:135) from the bump allocator; answer picked up at emit time
        int cap = 0;
        if (meta_ok) cap = ee0 > eb0 ? ee0 - eb0 : 0;
        else {
            for (int p = p0 + lane; p < p1; p += 32) { const int l = P.ends[p] - P.begins[p]; cap += l > 0 ? l : 0; }
            cap = (int)__reduce_add_sync(FULL, (unsigned)cap);
        }
        int alloc_b0 = 0;
        if (lane == 0) alloc_b0 = atomicAdd(&P.status[ST_ALLOC], cap);
        int64_t base = -1;
        int emitted = 0;
        bool redo = false;
        for (int p = p0; p < p1 && !redo; ++p) {
            const int eb = (meta_ok && p == p0) ? eb0 : P.begins[p], ee = (meta_ok && p == p0) ? ee0 : P.ends[p];
            if (P.skips && P.skips[p]) { redo = true; break; }
            int pos = eb;
            while (pos < ee) {
                const int end_rel = ee - pos;
                const int wlen = end_rel < WIN ? end_rel : WIN;
                const int nload = end_rel < wlen + LA ? end_rel : wlen + LA;
                const int lb = (pos - eb) < LBK ? (pos - eb) : LBK;   // look-back bytes available inside the element
                // ---- the window's bytes: already on their way (TMA prefetch), or staged now ----
                if (inflight) {
                    mbar_wait(&S.mbar[buf ^ 1], (par >> (buf ^ 1)) & 1u);
                    par ^= 1u << (buf ^ 1);
                    inflight = false;
                }
                bool all_ascii;
                uint8_t* Bw;
                {
                if (S.pf_pos != pos) {
                    // not prefetched (first window of the warp, multi-string rows): the same TMA bulk copy, waited for now
                    __syncwarp();
                    const uint8_t* src = P.chars + pos - lb;
                    const int head = (int)(reinterpret_cast<uintptr_t>(src) & 15);
                    const uint32_t bytes = (uint32_t)((head + lb + nload + 15) & ~15);
                    if (lane == 0) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        mbar_expect_tx(&S.mbar[buf ^ 1], bytes);
                        bulk_g2s(S.raw_pf[buf ^ 1], src - head, bytes, &S.mbar[buf ^ 1]);
                        S.pf_off = head + lb;
                        S.pf_n = (int32_t)bytes;
                    }
                    __syncwarp();
                    mbar_wait(&S.mbar[buf ^ 1], (par >> (buf ^ 1)) & 1u);
                    par ^= 1u << (buf ^ 1);
                }
                buf ^= 1;
                Bw = S.raw_pf[buf] + S.pf_off;
                {
                    // ASCII test over every 16-byte quad the copy brought in — up to 15 neighbouring bytes on either side included: a
                    // non-ASCII neighbour only sends an ASCII window down the general (equally exact) path
                    uint32_t hibits = 0;
                    const uint4* q4 = reinterpret_cast<const uint4*>(S.raw_pf[buf]);
                    const int nq = S.pf_n >> 4;
                    for (int q = lane; q < nq; q += 32) { const uint4 v = q4[q]; hibits |= v.x | v.y | v.z | v.w; }
                    all_ascii = !__any_sync(FULL, hibits & 0x80808080u);
                }
                __syncwarp();
                }
                if (lane == 0) S.pf_pos = -1;
                const int send = fast_window<IdT, L3>(S, P, lut32_smem, ascii_smem, lane, wlen, end_rel, nload, pos - eb, all_ascii, Bw, hook);
                hook.after_pass1();            // (windows that left early: the metadata steps still have to happen, in order)
                hook.after_pass3();
                if (send <= 0) redo = true;
                // ---- issue the prefetch of the warp's next window ----
                if (can_prefetch) {
                    int tpos = -1, teb = 0, tee = 0;
                    if (!redo && pos + send < ee) { tpos = pos + send; teb = eb; tee = ee; }
                    else if (p == p1 - 1) {                                   // the row ends here: first window of the next row
                        cp_async_wait_all();
                        __syncwarp();
                        if (S.pf_row < P.n_rows && S.pf_re == S.pf_rb + 1 && S.pf_ee > S.pf_eb) { tpos = teb = S.pf_eb; tee = S.pf_ee; }
                    }
                    if (tpos >= 0) {
                        const int lbt = (tpos - teb) < LBK ? (tpos - teb) : LBK;
                        const int relt = tee - tpos;
                        const int wl = relt < WIN ? relt : WIN;
                        const int nl = relt < wl + LA ? relt : wl + LA;
                        const uint8_t* src = P.chars + tpos - lbt;
                        const int head = (int)(reinterpret_cast<uintptr_t>(src) & 15);
                        const uint32_t bytes = (uint32_t)((head + lbt + nl + 15) & ~15);
                        if (lane == 0) {
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                            mbar_expect_tx(&S.mbar[buf ^ 1], bytes);
                            bulk_g2s(S.raw_pf[buf ^ 1], src - head, bytes, &S.mbar[buf ^ 1]);
                            S.pf_pos = tpos;
                            S.pf_off = head + lbt;
                            S.pf_n = (int32_t)bytes;
                        }
                        inflight = true;
                    }
                    __syncwarp();
                }
                if (redo) break;
                // ---- emit: ballot compaction of the surviving ids into the row's slot ----
                if (base < 0) {                       // bump-allocated slot: pick the allocator's answer up now
                    base = __shfl_sync(FULL, alloc_b0, 0);
                    if (lane == 0) const_cast<int32_t*>(P.row_base)[row] = (int32_t)base;     // the compaction pass reads it
                }
                if (base + emitted + send > P.tmp_cap) {
                    if (lane == 0) atomicOr(&P.status[ST_ERROR], ERR_TMP_OVERFLOW);
                } else {
                    int n_out = 0;
                    for (int w = lane; w - lane < send; w += 128) {
                        int32_t tok[4];
                        uint32_t m[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) tok[u] = ((w + 32 * u) < send && S.ids[w + 32 * u] != S.kDead) ? (int32_t)S.ids[w + 32 * u] : -1;
#pragma unroll
                        for (int u = 0; u < 4; ++u) m[u] = __ballot_sync(FULL, tok[u] >= 0);
                        // 16-bit ids stay 16 bits wide in the row slot (the first half of its i32 range): half the bytes written here and
                        // read by the compaction pass, which widens them
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (tok[u] >= 0) {
                                if constexpr (sizeof(IdT) == 2) reinterpret_cast<uint16_t*>(P.tmp_a + base)[emitted + n_out + __popc(m[u] & ltm)] = (uint16_t)tok[u];
                                else P.tmp_a[base + emitted + n_out + __popc(m[u] & ltm)] = tok[u];
                            }
                            n_out += __popc(m[u]);
                        }
                    }
                    emitted += n_out;
                }
                __syncwarp();
                pos += send;
            }
        }
        hook.after_pass1();       // rows without a window (empty): the next row's ticket and metadata are still due
        hook.after_pass3();
        cp_async_wait_all();
        __syncwarp();
        if (base < 0) {                               // nothing was emitted (handed back, or empty): the row still owns its slot range
            base = __shfl_sync(FULL, alloc_b0, 0);
            if (lane == 0) const_cast<int32_t*>(P.row_base)[row] = (int32_t)base;
        }
        if (lane == 0) {
            if (redo) redo_rows[atomicAdd(&P.status[ST_NREDO], 1)] = row;
            else { P.row_ext[row] = emitted; P.row_cnt[row] = emitted; P.row_flag[row] = sizeof(IdT) == 2 ? 4 : 0; }      // bit 2: 16-bit slot
        }
        if (can_prefetch) {
            row = S.pf_row;
            meta_ok = row < P.n_rows && S.pf_re == S.pf_rb + 1;
        } else {
            if (lane == 0) row = atomicAdd(&P.status[ST_TICKET], 1);
            row = __shfl_sync(FULL, row, 0);
        }
    }
    if (inflight) mbar_wait(&S.mbar[buf ^ 1], (par >> (buf ^ 1)) & 1u);      // no copy may be in flight when the CTA retires
}

// MODE 0: row slots + TMA-prefetched input (opt-in); 1: in-order single-pass emit (opt-in); 2: row slots, plain loop (the default);
// 3: the loop of 2 with the sharded emit — ids stored into every rank's buffers (kept out of the one-GPU kernel's instruction stream)
template <class IdT, int CTAS, bool L3, int MODE>
__global__ void __launch_bounds__(BLOCK_THREADS, CTAS) gpt2_bpe_fast_kernel(const __grid_constant__ RowParams P, int32_t* __restrict__ redo_rows) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* ascii_smem = smem_raw;                                            // [128]
    uint32_t* lut32_smem = reinterpret_cast<uint32_t*>(smem_raw + 128);        // [256]
    FastSmem<IdT>* warps = reinterpret_cast<FastSmem<IdT>*>(smem_raw + kFastSmemFixed);
    const int lane = threadIdx.x & 31;
    FastSmem<IdT>& S = warps[threadIdx.x >> 5];
    if (threadIdx.x < 128) ascii_smem[threadIdx.x] = P.cls.ascii[threadIdx.x];
    lut32_smem[threadIdx.x] = v7_lut_entry(P, threadIdx.x);
    __syncthreads();
    const uint32_t ltm = (1u << lane) - 1u;

    if constexpr (MODE == 1) {
        ordered_rows<IdT, L3>(S, P, lut32_smem, ascii_smem, lane, redo_rows);
        return;
    } else if constexpr (MODE == 0) {
        slot_rows<IdT, L3>(S, P, lut32_smem, ascii_smem, lane, redo_rows);
        return;
    } else {
        for (;;) {
            int row = 0;
            if (lane == 0) row = atomicAdd(&P.status[ST_TICKET], 1);
            row = __shfl_sync(FULL, row, 0);
            if (row >= P.n_rows) break;
            const int p0 = P.rb[row], p1 = P.re[row];
            int64_t base;
            int alloc_b0 = 0;
            if (P.alloc_base) {          // take the row's worst-case slot range (one id per byte, src/bpe_tokenizer.cpp:135) from the bump allocator
                int c = 0;
                for (int p = p0 + lane; p < p1; p += 32) { const int l = P.ends[p] - P.begins[p]; c += l > 0 ? l : 0; }
                c = (int)__reduce_add_sync(FULL, (unsigned)c);
                if (lane == 0) alloc_b0 = atomicAdd(&P.status[ST_ALLOC], c);       // (the result is first looked at when the row emits: no wait here)
                base = -1;
            } else if (P.direct_base) {
                base = p1 > p0 ? (int64_t)(P.begins[p0] - P.direct_byte0) + (int64_t)(p0 - P.direct_elem0) * P.direct_extra : 0;
                if (lane == 0) const_cast<int32_t*>(P.row_base)[row] = (int32_t)base;     // the compaction pass reads it
            } else base = P.row_base[row];
            int emitted = 0;
            bool redo = false;
            for (int p = p0; p < p1 && !redo; ++p) {
                const int eb = P.begins[p], ee = P.ends[p];
                if (P.skips && P.skips[p]) { redo = true; break; }
                int pos = eb;
                while (pos < ee) {
                    const int end_rel = ee - pos;
                    const int wlen = end_rel < WIN ? end_rel : WIN;
                    const int nload = end_rel < wlen + LA ? end_rel : wlen + LA;
                    const int lb = (pos - eb) < LBK ? (pos - eb) : LBK;   // look-back bytes available inside the element
                    const bool all_ascii = stage_window(S, P, pos, lb, nload, lane, S.B0());
                    NoHook nh;
                    const int send = fast_window<IdT, L3>(S, P, lut32_smem, ascii_smem, lane, wlen, end_rel, nload, pos - eb, all_ascii, S.B0(), nh);
                    if (send <= 0) { redo = true; break; }
                    if (base < 0) {                       // bump-allocated slot: pick the allocator's answer up now
                        base = __shfl_sync(FULL, alloc_b0, 0);
                        if (lane == 0) const_cast<int32_t*>(P.row_base)[row] = (int32_t)base;     // the compaction pass reads it
                    }
                    if (base + emitted + send > P.tmp_cap) {
                        if (lane == 0) atomicOr(&P.status[ST_ERROR], ERR_TMP_OVERFLOW);
                    } else {
                        const int nP = MODE == 3 ? P.peer.world : 0;  // > 0: sharded output, store into every rank's slot
                        const int64_t o0 = (nP ? (int64_t)P.peer.rank * P.peer.slot_capacity : 0) + base + emitted;
                        int32_t* outp = P.tmp_a + o0;
                        int n_out = 0;
                        for (int w = lane; w - lane < send; w += 128) {
                            int32_t tok[4];
                            uint32_t m[4];
    #pragma unroll
                            for (int u = 0; u < 4; ++u) tok[u] = ((w + 32 * u) < send && S.ids[w + 32 * u] != S.kDead) ? (int32_t)S.ids[w + 32 * u] : -1;
    #pragma unroll
                            for (int u = 0; u < 4; ++u) m[u] = __ballot_sync(FULL, tok[u] >= 0);
                            if (!nP) {
                                // 16-bit ids stay 16 bits wide in the row slot (the first half of its i32 range): half the bytes written
                                // here and read by the compaction pass, which widens them
    #pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    if (tok[u] >= 0) {
                                        if constexpr (sizeof(IdT) == 2) reinterpret_cast<uint16_t*>(P.tmp_a + base)[emitted + n_out + __popc(m[u] & ltm)] = (uint16_t)tok[u];
                                        else outp[n_out + __popc(m[u] & ltm)] = tok[u];
                                    }
                                    n_out += __popc(m[u]);
                                }
                            } else {
                                // sharded: compact into shared memory first (same 16-byte misalignment as the destination), stored wide below
                                const int shift = P.peer.wire16 ? (int)(o0 & 7) : (int)(o0 & 3);
    #pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    const int i = shift + n_out + __popc(m[u] & ltm);
                                    if (tok[u] >= 0) { if (P.peer.wire16) reinterpret_cast<uint16_t*>(S.key)[i] = (uint16_t)tok[u]; else reinterpret_cast<int32_t*>(S.key)[i] = tok[u]; }
                                    n_out += __popc(m[u]);
                                }
                            }
                        }
                        if (nP) {
                            __syncwarp();
                            // 16-byte chunks of the staged ids go to every rank with one vector store each (remote stores are
                            // transaction-bound: few wide stores instead of many 2-4 byte ones); ragged head / tail element-wise
                            const int per = P.peer.wire16 ? 8 : 4;
                            const int shift = (int)(o0 & (per - 1)), total = shift + n_out;
                            const int64_t o_al = o0 - shift;                          // 16-byte aligned destination element
                            const uint4* sv = reinterpret_cast<const uint4*>(S.key);
                            for (int c = lane; c * per < total; c += 32) {
                                const int lo = c * per, hi = lo + per;
                                if (P.peer.ids_mc && !P.peer.wire16) {           // NVLS multicast: one store reaches every rank
                                    int32_t* mc = P.peer.ids_mc + o_al;
                                    if (lo >= shift && hi <= total) mc_store4(mc + lo, sv[c]);
                                    else for (int i = lo < shift ? shift : lo; i < (hi < total ? hi : total); ++i) mc_store(mc + i, reinterpret_cast<const int32_t*>(S.key)[i]);
                                } else if (lo >= shift && hi <= total) {
                                    const uint4 v = sv[c];
                                    for (int p = 0; p < nP; ++p) {
                                        uint4* dp = P.peer.wire16 ? reinterpret_cast<uint4*>(P.peer.ids16[p] + o_al) : reinterpret_cast<uint4*>(P.peer.ids[p] + o_al);
                                        dp[c] = v;
                                    }
                                } else {
                                    for (int i = lo < shift ? shift : lo; i < (hi < total ? hi : total); ++i)
                                        for (int p = 0; p < nP; ++p) {
                                            if (P.peer.wire16) P.peer.ids16[p][o_al + i] = reinterpret_cast<const uint16_t*>(S.key)[i];
                                            else P.peer.ids[p][o_al + i] = reinterpret_cast<const int32_t*>(S.key)[i];
                                        }
                                }
                            }
                            __syncwarp();
                        }
                        emitted += n_out;
                    }
                    __syncwarp();
                    pos += send;
                }
            }
            if (base < 0) {                               // nothing was emitted (handed back, or empty): the row still owns its slot range
                base = __shfl_sync(FULL, alloc_b0, 0);
                if (lane == 0) const_cast<int32_t*>(P.row_base)[row] = (int32_t)base;
            }
            if (lane == 0) {
                if (redo) redo_rows[atomicAdd(&P.status[ST_NREDO], 1)] = row;
                else { P.row_ext[row] = emitted; P.row_cnt[row] = emitted; P.row_flag[row] = (sizeof(IdT) == 2 && MODE != 3) ? 4 : 0; }      // bit 2: 16-bit slot
            }
            if (MODE == 3 && P.peer.world && !redo) {          // sharded: publish the row's extent to every rank
                const int64_t o0 = (int64_t)P.peer.rank * P.peer.slot_capacity + base;
                if (P.peer.begins_mc) {
                    if (lane == 0) {
                        mc_store(P.peer.begins_mc + (int64_t)P.peer.rank * P.peer.rows_per_rank + row, (int32_t)o0);
                        mc_store(P.peer.ends_mc + (int64_t)P.peer.rank * P.peer.rows_per_rank + row, (int32_t)(o0 + emitted));
                    }
                } else if (lane < P.peer.world) {
                    P.peer.begins[lane][(int64_t)P.peer.rank * P.peer.rows_per_rank + row] = (int32_t)o0;
                    P.peer.ends[lane][(int64_t)P.peer.rank * P.peer.rows_per_rank + row] = (int32_t)(o0 + emitted);
                }
                if (lane == 0) atomicAdd(&P.status[ST_TOTAL], emitted);
            }
        }
    }
}


// ---- closing the gaps of handed-back rows (both kernels return at once when the fast kernel finished every row) ----------
// 1. rows after the first handed-back row are parked in tmp_a at their (gapped) offsets, next to the ids the generic kernels
//    produced for the handed-back rows themselves;
__global__ void ordered_stash_kernel(const __grid_constant__ RowParams P) {
    if (P.status[ST_NREDO] == 0) return;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int first = P.n_rows - P.status[ST_MINREDO];
    for (int r = first + warp; r < P.n_rows; r += nwarps) {
        if (P.row_flag[r] & 2) continue;
        const int b = P.oo.begins[r], e = P.oo.ends[r];
        if ((long long)e > P.oo.cap || (long long)e > P.tmp_cap) continue;
        for (int t = b + lane; t < e; t += 32) P.tmp_a[t] = P.oo.ids[t];
    }
}
// 2. a second look-back chain over the true counts moves every row to its final offset and rewrites begins / ends / total.
__global__ void ordered_recompact_kernel(const __grid_constant__ RowParams P) {
    if (P.status[ST_NREDO] == 0) return;
    const int lane = threadIdx.x & 31;
    const uint32_t ltm = (1u << lane) - 1u;
    const int first = P.n_rows - P.status[ST_MINREDO];
    for (;;) {
        int r = 0;
        if (lane == 0) r = first + atomicAdd(&P.status[ST_TICKET3], 1);
        r = __shfl_sync(FULL, r, 0);
        if (r >= P.n_rows) break;
        const uint8_t flag = P.row_flag[r];
        const int gb = (flag & 2) ? P.row_base[r] : P.oo.begins[r];
        const int cnt = (flag & 2) ? P.row_cnt[r] : P.oo.ends[r] - gb;
        const int ext = (flag & 2) ? P.row_ext[r] : cnt;
        const long long excl = lookback_exclusive(P.oo.desc, P.oo.epoch + 1u, first, r, (uint32_t)cnt, gb, lane);
        if (excl + cnt <= P.oo.cap && (long long)gb + ext <= P.tmp_cap) {
            int32_t* outp = P.oo.ids + excl;
            if (!(flag & 1)) {
                for (int t = lane; t < ext; t += 32) outp[t] = P.tmp_a[gb + t];
            } else {                                         // holes of giant pieces are filtered while copying
                int d = 0;
                for (int t0 = 0; t0 < ext; t0 += 32) {
                    const int t = t0 + lane;
                    const int v = t < ext ? P.tmp_a[gb + t] : -1;
                    const uint32_t m = __ballot_sync(FULL, v >= 0);
                    if (v >= 0) outp[d + __popc(m & ltm)] = v;
                    d += __popc(m);
                }
            }
        } else if (lane == 0) atomicOr(&P.status[ST_ERROR], ERR_TMP_OVERFLOW);
        if (lane == 0) {
            P.oo.begins[r] = (int32_t)excl; P.oo.ends[r] = (int32_t)(excl + cnt);
            if (r == P.n_rows - 1) { P.status[ST_TOTAL] = (int32_t)(excl + cnt); if (P.oo.total) *P.oo.total = excl + cnt; }
        }
    }
}

}  // namespace b200tok
