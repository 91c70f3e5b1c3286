// regex_vm.cuh — a small backtracking regex machine for GENERAL split patterns (the reference compiles any pattern with PCRE2:
// src/regex_split.cpp:147-151, src/utils.cpp:256-272).  The named tokenizer patterns keep their closed forms (tok_core.cuh); every
// other pattern inside the supported syntax is compiled on the host (regex_compile.cpp) into this machine's program and evaluated
// on the device "for every start position in parallel" by the same split framework (match_at -> split_window).
//
// Semantics reproduced: PCRE2_UTF | PCRE2_UCP matching anchored at a given position — alternatives tried in order, greedy
// quantifiers with backtracking (character by character), (?i:...) for ASCII letters plus the two non-ASCII characters that fold
// onto ASCII under PCRE2 caseless UTF (U+017F -> s, U+212A -> k), single-character look-aheads, ^ and $.
// A program is a DAG (no backward jumps), so it always terminates; its backtrack stack is bounded (kVmStack), and a program that
// could overflow it is refused at compile time.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define VM_HD __host__ __device__ __forceinline__
#else
#define VM_HD inline
#endif

namespace b200tok {

enum : uint32_t {
    VM_SET = 1,      // a: set index — the current character must be in the set; advance
    VM_SPLIT = 2,    // a: first choice, b: second choice (tried on backtrack)
    VM_JMP = 3,      // a: target (always forward)
    VM_MATCH = 4,
    VM_LOOP = 5,     // a: set index, b: min | max << 12 (max 0xFFF = unbounded) | possessive << 24: greedy X{min,max} over one set
    VM_LA = 6,       // a: set index — look-ahead: the next character is in the set (fails at the end of the subject)
    VM_NLA = 7,      // a: set index — negative look-ahead (succeeds at the end of the subject)
    VM_BOL = 8,      // ^
    VM_EOL = 9,      // $ (end of the subject, or before a final \n — PCRE2 without DOLLAR_ENDONLY)
};
constexpr int kVmStack = 24;

struct VmInst { uint32_t op_a; uint32_t b; };        // op in the low 8 bits, a in the upper 24
struct VmSet {
    uint32_t gc_mask;      // bit g: general category g (unicode_gc_ranges.inc order) is in the set
    uint32_t range_off;    // first (lo, hi) code-point pair in the ranges array
    uint16_t n_ranges;
    uint8_t cls_mask;      // C_S / C_W bits of the PCRE2-generated class table (\s, \w are not unions of general categories)
    uint8_t negate;
};
struct VmProgram {
    const VmInst* code;
    const VmSet* sets;
    const uint32_t* ranges;
    const uint16_t* gc1;   // two-stage general-category table: gc1[cp >> 8] -> block, gc2[block * 256 + (cp & 255)]
    const uint8_t* gc2;
    int32_t n_code;
};

VM_HD uint32_t vm_gc(const VmProgram& V, uint32_t cp) {
    if (cp >= 0x110000u) return 29;
    return V.gc2[(uint32_t)V.gc1[cp >> 8] * 256u + (cp & 255u)];
}
VM_HD bool vm_in_set(const VmProgram& V, uint32_t set, uint32_t cp, uint8_t cls) {
    const VmSet s = V.sets[set];
    bool in = ((s.gc_mask >> vm_gc(V, cp)) & 1u) || (cls & s.cls_mask);
    for (uint32_t r = 0; !in && r < s.n_ranges; ++r) in = cp >= V.ranges[s.range_off + 2 * r] && cp <= V.ranges[s.range_off + 2 * r + 1];
    return in != (s.negate != 0);
}

// Decodes the character at i of context c (element end `end`); false = no character can be read there: either the subject ends
// (i >= end) or the context cannot see that far (peek then exceeds what is known and the caller defers the decision).
template <class C>
VM_HD bool vm_char(const C& c, int i, int end, int& peek, uint32_t& cp, int& len, uint8_t& cls) {
    if (i >= end) return false;
    if (i + 1 > peek) peek = i + 1;
    if (i >= c.lim()) return false;
    const uint8_t b0 = c.byte(i);
    cls = c.cls(i);
    if (b0 < 0x80) { cp = b0; len = 1; return true; }
    const int need = b0 >= 0xF0 ? 3 : b0 >= 0xE0 ? 2 : b0 >= 0xC0 ? 1 : 0;
    cp = 0xFFFFFFFFu; len = 1;                                    // malformed (out of contract for the reference): one "other" byte
    if (need == 0 || b0 >= 0xF8 || i + need >= end) return true;
    if (i + need + 1 > peek) peek = i + need + 1;
    if (i + need >= c.lim()) return false;
    uint32_t v = need == 1 ? (b0 & 0x1Fu) : need == 2 ? (b0 & 0x0Fu) : (b0 & 0x07u);
    for (int k = 1; k <= need; ++k) {
        const uint8_t b = c.byte(i + k);
        if ((b & 0xC0) != 0x80) return true;
        v = (v << 6) | (b & 0x3Fu);
    }
    cp = v; len = need + 1;
    return true;
}

// The match of program V anchored at p: length (0 = no match), 1 + highest byte index examined.
template <class C>
VM_HD void vm_match(const C& c, const VmProgram& V, int p, int end, int& out_len, int& out_peek) {
    int bt_pc[kVmStack], bt_pos[kVmStack], bt_lo[kVmStack];      // bt_lo >= 0: a greedy loop that can still give characters back
    int sp = 0, pc = 0, pos = p, peek = p + 1;
    out_len = 0;
    for (;;) {
        bool fail = false;
        const VmInst I = V.code[pc];
        const uint32_t op = I.op_a & 0xFFu, a = I.op_a >> 8;
        uint32_t cp; int len; uint8_t cls;
        switch (op) {
        case VM_SET:
            if (vm_char(c, pos, end, peek, cp, len, cls) && vm_in_set(V, a, cp, cls)) { pos += len; ++pc; } else fail = true;
            break;
        case VM_LOOP: {
            const int mn = (int)(I.b & 0xFFFu), mx = (int)((I.b >> 12) & 0xFFFu);
            const bool possessive = (I.b >> 24) & 1u;
            int n = 0, lo = pos;
            while ((mx == 0xFFF || n < mx) && vm_char(c, pos, end, peek, cp, len, cls) && vm_in_set(V, a, cp, cls)) {
                pos += len; ++n;
                if (n == mn) lo = pos;
            }
            if (n < mn) { fail = true; break; }
            // (mn == 0: lo is the loop's start)  A greedy loop leaves ONE entry that hands characters back one at a time
            if (!possessive && pos > lo && sp < kVmStack) { bt_pc[sp] = pc + 1; bt_pos[sp] = pos; bt_lo[sp] = lo; ++sp; }
            ++pc;
            break;
        }
        case VM_SPLIT:
            if (sp < kVmStack) { bt_pc[sp] = (int)I.b; bt_pos[sp] = pos; bt_lo[sp] = -1; ++sp; }
            pc = (int)a;
            break;
        case VM_JMP: pc = (int)a; break;
        case VM_LA:
            if (vm_char(c, pos, end, peek, cp, len, cls) && vm_in_set(V, a, cp, cls)) ++pc; else fail = true;
            break;
        case VM_NLA:
            if (vm_char(c, pos, end, peek, cp, len, cls) && vm_in_set(V, a, cp, cls)) fail = true; else ++pc;
            break;
        case VM_BOL: if (pos == 0) ++pc; else fail = true; break;
        case VM_EOL:
            if (pos >= end) ++pc;
            else if (vm_char(c, pos, end, peek, cp, len, cls) && cp == '\n' && pos + 1 >= end) ++pc;
            else fail = true;
            break;
        case VM_MATCH:
            if (pos > p) { out_len = pos - p; out_peek = peek; return; }
            fail = true;                          // an empty match ends the reference's loop like "no match" (src/regex_split.cpp:156)
            break;
        default: fail = true; break;
        }
        if (!fail) continue;
        // backtrack
        for (;;) {
            if (sp == 0) { out_len = 0; out_peek = peek; return; }
            const int t = sp - 1;
            if (bt_lo[t] < 0) { pc = bt_pc[t]; pos = bt_pos[t]; sp = t; break; }
            // a greedy loop gives back its last character
            int q = bt_pos[t] - 1;
            while (q > bt_lo[t] && (c.byte(q) & 0xC0) == 0x80) --q;
            if (q < bt_lo[t]) { sp = t; continue; }
            pc = bt_pc[t]; pos = q;
            if (q > bt_lo[t]) bt_pos[t] = q; else sp = t;
            break;
        }
    }
}

}  // namespace b200tok
