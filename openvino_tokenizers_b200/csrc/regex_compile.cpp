// regex_compile.cpp — host-side compiler of general split patterns into the program of regex_vm.cuh.
// Supported syntax (what the split patterns of the reference's model list use; anything else is refused with
// B200TOK_E_UNSUPPORTED, never approximated): literals (UTF-8), escapes \s \S \w \W \d \D \p{..} \P{..} (general categories)
// \r \n \t \f \x{H..} \xHH and escaped punctuation, `.`, bracket sets with ranges / escapes / negation, groups ( ) (?: ) (?i: ),
// a leading (?i), alternation, quantifiers ? * + {m} {m,} {m,n} (greedy; possessive ?+ *+ ++ on a single character set),
// look-aheads (?= ) (?! ) over ONE character set, ^ and $.
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/b200tok.h"
#include "tables.hpp"

namespace b200tok {

namespace {
struct GcRun { uint32_t cp; uint8_t gc; };
const GcRun kGcRuns[] = {
#include "unicode_gc_ranges.inc"
};
const char* const kGcNames[30] = {"Lu", "Ll", "Lt", "Lm", "Lo", "Mn", "Mc", "Me", "Nd", "Nl", "No", "Pc", "Pd", "Ps", "Pe", "Pi", "Pf", "Po",
                                  "Sm", "Sc", "Sk", "So", "Zs", "Zl", "Zp", "Cc", "Cf", "Cs", "Co", "Cn"};

struct SetSpec {              // a character set before it is interned
    uint32_t gc_mask = 0;
    uint8_t cls_mask = 0;
    bool negate = false;
    std::vector<std::pair<uint32_t, uint32_t>> ranges;
    bool simple_positive() const { return !negate; }
};

struct Node {
    enum Kind { SET, CAT, ALT, REPEAT, LOOK, BOL, EOL, EMPTY } kind = EMPTY;
    int set = -1;                    // SET / LOOK
    bool neg = false;                // LOOK
    int mn = 1, mx = 1;              // REPEAT (mx < 0: unbounded)
    bool possessive = false;
    std::vector<std::unique_ptr<Node>> kids;
};
using NodeP = std::unique_ptr<Node>;

struct Parser {
    const std::string& s;
    size_t i = 0;
    HostVm& out;
    std::vector<SetSpec> specs;
    std::string err;
    explicit Parser(const std::string& pat, HostVm& o) : s(pat), out(o) {}

    bool fail(const std::string& why) { if (err.empty()) err = why; return false; }
    bool eof() const { return i >= s.size(); }

    bool decode(uint32_t& cp) {      // one UTF-8 character of the pattern
        const uint8_t b0 = (uint8_t)s[i];
        if (b0 < 0x80) { cp = b0; ++i; return true; }
        const int need = b0 >= 0xF0 ? 3 : b0 >= 0xE0 ? 2 : b0 >= 0xC0 ? 1 : -1;
        if (need < 0 || i + (size_t)need >= s.size()) return fail("malformed UTF-8 in the pattern");
        uint32_t v = need == 1 ? (b0 & 0x1Fu) : need == 2 ? (b0 & 0x0Fu) : (b0 & 0x07u);
        for (int k = 1; k <= need; ++k) {
            const uint8_t b = (uint8_t)s[i + k];
            if ((b & 0xC0) != 0x80) return fail("malformed UTF-8 in the pattern");
            v = (v << 6) | (b & 0x3Fu);
        }
        cp = v; i += (size_t)need + 1;
        return true;
    }
    static void add_char(SetSpec& S, uint32_t cp, bool ci) {
        S.ranges.emplace_back(cp, cp);
        if (!ci) return;
        if (cp >= 'a' && cp <= 'z') S.ranges.emplace_back(cp - 32, cp - 32);
        else if (cp >= 'A' && cp <= 'Z') S.ranges.emplace_back(cp + 32, cp + 32);
        if (cp == 's' || cp == 'S') S.ranges.emplace_back(0x17F, 0x17F);      // PCRE2 caseless UTF: LATIN SMALL LETTER LONG S folds to s
        if (cp == 'k' || cp == 'K') S.ranges.emplace_back(0x212A, 0x212A);    // KELVIN SIGN folds to k
        if (cp == 0x17F) { S.ranges.emplace_back('s', 's'); S.ranges.emplace_back('S', 'S'); }
        if (cp == 0x212A) { S.ranges.emplace_back('k', 'k'); S.ranges.emplace_back('K', 'K'); }
    }
    bool property(SetSpec& S, bool& negated) {        // after \p or \P: {Name} or a single letter
        std::string name;
        if (!eof() && s[i] == '{') {
            const size_t e = s.find('}', i);
            if (e == std::string::npos) return fail("unterminated \\p{");
            name = s.substr(i + 1, e - i - 1);
            i = e + 1;
        } else if (!eof()) name = std::string(1, s[i++]);
        if (!name.empty() && name[0] == '^') { negated = !negated; name.erase(0, 1); }
        uint32_t mask = 0;
        if (name.size() == 1 || name == "L&") {
            for (int g = 0; g < 30; ++g) {
                if (name == "L&") { if (g <= 2) mask |= 1u << g; }
                else if (kGcNames[g][0] == name[0]) mask |= 1u << g;
            }
        } else
            for (int g = 0; g < 30; ++g) if (name == kGcNames[g]) mask |= 1u << g;
        if (!mask) return fail("unsupported Unicode property \\p{" + name + "} (general categories only)");
        S.gc_mask |= mask;
        return true;
    }
    // An escape; `in_set`: inside [...].  Adds to S; negated = the escape is a complemented class (\S \W \D \P{..}).
    bool escape(SetSpec& S, bool ci, bool& negated) {
        negated = false;
        if (eof()) return fail("dangling backslash");
        const char c = s[i++];
        switch (c) {
        case 's': S.cls_mask |= C_S; return true;
        case 'S': S.cls_mask |= C_S; negated = true; return true;
        case 'w': S.cls_mask |= C_W; return true;
        case 'W': S.cls_mask |= C_W; negated = true; return true;
        case 'd': S.gc_mask |= 1u << 8; return true;
        case 'D': S.gc_mask |= 1u << 8; negated = true; return true;
        case 'p': return property(S, negated);
        case 'P': negated = true; return property(S, negated);
        case 'r': add_char(S, '\r', false); return true;
        case 'n': add_char(S, '\n', false); return true;
        case 't': add_char(S, '\t', false); return true;
        case 'f': add_char(S, '\f', false); return true;
        case 'a': add_char(S, 7, false); return true;
        case 'e': add_char(S, 27, false); return true;
        case '0': add_char(S, 0, false); return true;
        case 'x': {
            uint32_t v = 0;
            auto hex = [](char h) { return h >= '0' && h <= '9' ? h - '0' : h >= 'a' && h <= 'f' ? h - 'a' + 10 : h >= 'A' && h <= 'F' ? h - 'A' + 10 : -1; };
            if (!eof() && s[i] == '{') {
                ++i;
                int n = 0;
                while (!eof() && s[i] != '}') { const int h = hex(s[i++]); if (h < 0) return fail("bad \\x{...}"); v = v * 16 + (uint32_t)h; ++n; }
                if (eof() || n == 0) return fail("bad \\x{...}");
                ++i;
            } else {
                for (int n = 0; n < 2 && !eof() && hex(s[i]) >= 0; ++n) v = v * 16 + (uint32_t)hex(s[i++]);
            }
            add_char(S, v, ci);
            return true;
        }
        default:
            if ((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '1' && c <= '9'))
                return fail(std::string("unsupported escape \\") + c);
            add_char(S, (uint8_t)c, false);          // escaped punctuation
            return true;
        }
    }
    int intern(const SetSpec& S) {
        specs.push_back(S);
        return (int)specs.size() - 1;
    }
    NodeP set_node(const SetSpec& S) {
        auto n = std::make_unique<Node>();
        n->kind = Node::SET;
        n->set = intern(S);
        return n;
    }
    bool bracket(SetSpec& S, bool ci) {          // after '['
        if (!eof() && s[i] == '^') { S.negate = true; ++i; }
        bool first = true;
        for (;;) {
            if (eof()) return fail("unterminated [");
            if (s[i] == ']' && !first) { ++i; return true; }
            first = false;
            if (s[i] == '[' && i + 1 < s.size() && s[i + 1] == ':') return fail("POSIX classes are not supported");
            uint32_t lo;
            bool is_class = false;
            if (s[i] == '\\') {
                ++i;
                SetSpec T;
                bool neg = false;
                if (!escape(T, ci, neg)) return false;
                if (neg) return fail("a complemented class (\\S \\W \\D \\P) inside [...] is not supported");
                if (T.gc_mask || T.cls_mask) { S.gc_mask |= T.gc_mask; S.cls_mask |= T.cls_mask; is_class = true; }
                if (is_class) continue;
                lo = T.ranges.front().first;       // a single escaped character (its case variants are re-added below)
            } else if (!decode(lo)) return false;
            uint32_t hi = lo;
            if (i + 1 < s.size() && s[i] == '-' && s[i + 1] != ']') {
                ++i;
                if (s[i] == '\\') {
                    ++i;
                    SetSpec T;
                    bool neg = false;
                    if (!escape(T, false, neg)) return false;
                    if (neg || T.gc_mask || T.cls_mask || T.ranges.empty()) return fail("bad range end in [...]");
                    hi = T.ranges.front().first;
                } else if (!decode(hi)) return false;
                if (hi < lo) return fail("range out of order in [...]");
            }
            if (lo == hi) add_char(S, lo, ci);
            else {
                S.ranges.emplace_back(lo, hi);
                if (ci) {      // the other case of the ASCII letters inside the range
                    const uint32_t a = std::max<uint32_t>(lo, 'a'), b = std::min<uint32_t>(hi, 'z');
                    if (a <= b) S.ranges.emplace_back(a - 32, b - 32);
                    const uint32_t A = std::max<uint32_t>(lo, 'A'), B = std::min<uint32_t>(hi, 'Z');
                    if (A <= B) S.ranges.emplace_back(A + 32, B + 32);
                }
            }
        }
    }
    // atom := set | group | look-ahead | anchor
    NodeP atom(bool ci) {
        if (eof()) return nullptr;
        const char c = s[i];
        if (c == '(') {
            ++i;
            bool look = false, neg = false, gci = ci;
            if (!eof() && s[i] == '?') {
                ++i;
                if (eof()) { fail("unterminated group"); return nullptr; }
                if (s[i] == ':') ++i;
                else if (s[i] == '=') { look = true; ++i; }
                else if (s[i] == '!') { look = true; neg = true; ++i; }
                else if (s[i] == 'i' && i + 1 < s.size() && s[i + 1] == ':') { gci = true; i += 2; }
                else { fail("unsupported group construct (?" + std::string(1, s[i]) + " (only (?: (?i: (?= (?! are)"); return nullptr; }
            }
            NodeP inner = alternation(gci);
            if (!inner) return nullptr;
            if (eof() || s[i] != ')') { fail("unterminated group"); return nullptr; }
            ++i;
            if (look) {
                int set = single_set(*inner);
                if (set < 0) { fail("look-aheads over more than one character set are not supported"); return nullptr; }
                auto n = std::make_unique<Node>();
                n->kind = Node::LOOK; n->set = set; n->neg = neg;
                return n;
            }
            return inner;
        }
        if (c == '[') { ++i; SetSpec S; if (!bracket(S, ci)) return nullptr; return set_node(S); }
        if (c == '.') { ++i; SetSpec S; S.negate = true; S.ranges.emplace_back('\n', '\n'); return set_node(S); }
        if (c == '^') { ++i; auto n = std::make_unique<Node>(); n->kind = Node::BOL; return n; }
        if (c == '$') { ++i; auto n = std::make_unique<Node>(); n->kind = Node::EOL; return n; }
        if (c == '\\') {
            ++i;
            if (!eof() && (s[i] == 'b' || s[i] == 'B' || s[i] == 'A' || s[i] == 'z' || s[i] == 'Z' || s[i] == 'G' || s[i] == 'K' || s[i] == 'R' || s[i] == 'X' || s[i] == 'h' || s[i] == 'H' || s[i] == 'v' || s[i] == 'V' || s[i] == 'N')) {
                fail(std::string("unsupported escape \\") + s[i]);
                return nullptr;
            }
            SetSpec S;
            bool neg = false;
            if (!escape(S, ci, neg)) return nullptr;
            S.negate = neg;
            return set_node(S);
        }
        if (c == '*' || c == '+' || c == '?' || c == '{' || c == ')' || c == '|') return nullptr;
        uint32_t cp;
        if (!decode(cp)) return nullptr;
        SetSpec S;
        add_char(S, cp, ci);
        return set_node(S);
    }
    // the set index if the node is one character set (or an alternation of positive sets: merged), else -1
    int single_set(const Node& n) {
        if (n.kind == Node::SET) return n.set;
        if (n.kind == Node::CAT && n.kids.size() == 1) return single_set(*n.kids[0]);
        if (n.kind == Node::ALT) {
            SetSpec U;
            for (auto& k : n.kids) {
                const int si = single_set(*k);
                if (si < 0 || specs[(size_t)si].negate) return -1;
                U.gc_mask |= specs[(size_t)si].gc_mask; U.cls_mask |= specs[(size_t)si].cls_mask;
                U.ranges.insert(U.ranges.end(), specs[(size_t)si].ranges.begin(), specs[(size_t)si].ranges.end());
            }
            return intern(U);
        }
        return -1;
    }
    NodeP quantified(bool ci) {
        NodeP a = atom(ci);
        if (!a) return nullptr;
        for (;;) {
            if (eof()) return a;
            int mn, mx;
            const char c = s[i];
            if (c == '*') { mn = 0; mx = -1; ++i; }
            else if (c == '+') { mn = 1; mx = -1; ++i; }
            else if (c == '?') { mn = 0; mx = 1; ++i; }
            else if (c == '{') {
                size_t j = i + 1;
                auto num = [&](int& v) { if (j >= s.size() || s[j] < '0' || s[j] > '9') return false; v = 0; while (j < s.size() && s[j] >= '0' && s[j] <= '9') v = v * 10 + (s[j++] - '0'); return true; };
                if (!num(mn)) return a;                      // a literal '{'... PCRE2 treats it so; we only get here for real quantifiers
                mx = mn;
                if (j < s.size() && s[j] == ',') { ++j; if (j < s.size() && s[j] == '}') mx = -1; else if (!num(mx)) { fail("bad {m,n}"); return nullptr; } }
                if (j >= s.size() || s[j] != '}') { fail("bad {m,n}"); return nullptr; }
                i = j + 1;
                if (mx >= 0 && mx < mn) { fail("bad {m,n}"); return nullptr; }
            } else return a;
            bool possessive = false;
            if (!eof() && s[i] == '+') { possessive = true; ++i; }
            else if (!eof() && s[i] == '?') { fail("lazy quantifiers are not supported"); return nullptr; }
            if (a->kind == Node::BOL || a->kind == Node::EOL || a->kind == Node::LOOK) { fail("quantified assertion"); return nullptr; }
            auto r = std::make_unique<Node>();
            r->kind = Node::REPEAT; r->mn = mn; r->mx = mx; r->possessive = possessive;
            r->kids.push_back(std::move(a));
            a = std::move(r);
        }
    }
    NodeP sequence(bool ci) {
        auto n = std::make_unique<Node>();
        n->kind = Node::CAT;
        while (!eof() && s[i] != '|' && s[i] != ')') {
            NodeP q = quantified(ci);
            if (!q) { if (err.empty()) fail("unexpected character in the pattern"); return nullptr; }
            n->kids.push_back(std::move(q));
        }
        return n;
    }
    NodeP alternation(bool ci) {
        auto n = std::make_unique<Node>();
        n->kind = Node::ALT;
        for (;;) {
            NodeP q = sequence(ci);
            if (!q) return nullptr;
            n->kids.push_back(std::move(q));
            if (!eof() && s[i] == '|') { ++i; continue; }
            break;
        }
        if (n->kids.size() == 1) return std::move(n->kids[0]);
        return n;
    }

    // ---- code generation ----
    void emit(uint32_t op, uint32_t a, uint32_t b = 0) { out.code.push_back(VmInst{op | (a << 8), b}); }
    bool gen(const Node& n) {
        switch (n.kind) {
        case Node::EMPTY: return true;
        case Node::SET: emit(VM_SET, (uint32_t)n.set); return true;
        case Node::LOOK: emit(n.neg ? VM_NLA : VM_LA, (uint32_t)n.set); return true;
        case Node::BOL: emit(VM_BOL, 0); return true;
        case Node::EOL: emit(VM_EOL, 0); return true;
        case Node::CAT:
            for (auto& k : n.kids) if (!gen(*k)) return false;
            return true;
        case Node::ALT: {
            std::vector<size_t> jumps;
            for (size_t k = 0; k < n.kids.size(); ++k) {
                size_t split = 0;
                const bool last = k + 1 == n.kids.size();
                if (!last) { split = out.code.size(); emit(VM_SPLIT, (uint32_t)out.code.size() + 1, 0); }
                if (!gen(*n.kids[k])) return false;
                if (!last) {
                    jumps.push_back(out.code.size());
                    emit(VM_JMP, 0);
                    out.code[split].b = (uint32_t)out.code.size();
                }
            }
            for (size_t j : jumps) out.code[j].op_a = VM_JMP | ((uint32_t)out.code.size() << 8);
            return true;
        }
        case Node::REPEAT: {
            const Node& c = *n.kids[0];
            const int set = single_set(c);
            if (set >= 0) {
                if (n.mn > 0xFFE || n.mx > 0xFFE) return fail("repeat count too large");
                emit(VM_LOOP, (uint32_t)set, (uint32_t)n.mn | ((uint32_t)(n.mx < 0 ? 0xFFF : n.mx) << 12) | (n.possessive ? 1u << 24 : 0u));
                return true;
            }
            if (n.possessive) return fail("possessive quantifiers are supported on a single character set only");
            if (n.mx < 0) return fail("* and + over a group that is not a single character set are not supported");
            if (n.mx > 8) return fail("{m,n} over a group: n > 8");
            for (int k = 0; k < n.mn; ++k) if (!gen(c)) return false;
            std::vector<size_t> splits;                       // the optional copies, nested: (x(x(x)?)?)?
            for (int k = n.mn; k < n.mx; ++k) {
                splits.push_back(out.code.size());
                emit(VM_SPLIT, (uint32_t)out.code.size() + 1, 0);
                if (!gen(c)) return false;
            }
            for (size_t sidx : splits) out.code[sidx].b = (uint32_t)out.code.size();
            return true;
        }
        }
        return false;
    }
};
}  // namespace

const HostGcTables& host_gc_tables() {
    static const HostGcTables tables = [] {
        HostGcTables t;
        std::vector<uint8_t> flat(0x110000);
        const size_t n = sizeof(kGcRuns) / sizeof(kGcRuns[0]);
        for (size_t i = 0; i < n; ++i) {
            const uint32_t a = kGcRuns[i].cp, b = (i + 1 < n) ? kGcRuns[i + 1].cp : 0x110000u;
            for (uint32_t cp = a; cp < b; ++cp) flat[cp] = kGcRuns[i].gc;
        }
        t.stage1.resize(0x1100);
        std::map<std::vector<uint8_t>, uint16_t> seen;
        for (uint32_t blk = 0; blk < 0x1100; ++blk) {
            std::vector<uint8_t> v(flat.begin() + blk * 256, flat.begin() + blk * 256 + 256);
            auto it = seen.find(v);
            if (it == seen.end()) {
                it = seen.emplace(v, (uint16_t)seen.size()).first;
                t.stage2.insert(t.stage2.end(), v.begin(), v.end());
            }
            t.stage1[blk] = it->second;
        }
        return t;
    }();
    return tables;
}

int compile_regex(const std::string& pattern, HostVm& out, std::string& err) {
    out = HostVm{};
    Parser P(pattern, out);
    bool ci = false;
    if (pattern.compare(0, 4, "(?i)") == 0) { ci = true; P.i = 4; }
    NodeP root = P.alternation(ci);
    if (!root || !P.eof()) {
        err = "RegexSplit: the pattern is outside the syntax the GPU splitter compiles (" + (P.err.empty() ? std::string("unbalanced parenthesis") : P.err) + "): " + pattern;
        return B200TOK_E_UNSUPPORTED;
    }
    if (!P.gen(*root)) { err = "RegexSplit: the pattern is outside the syntax the GPU splitter compiles (" + P.err + "): " + pattern; return B200TOK_E_UNSUPPORTED; }
    P.emit(VM_MATCH, 0);
    for (const SetSpec& S : P.specs) {
        VmSet v{};
        v.gc_mask = S.gc_mask; v.cls_mask = S.cls_mask; v.negate = S.negate ? 1 : 0;
        v.range_off = (uint32_t)out.ranges.size();
        v.n_ranges = (uint16_t)S.ranges.size();
        for (auto& r : S.ranges) { out.ranges.push_back(r.first); out.ranges.push_back(r.second); }
        out.sets.push_back(v);
    }
    if (out.ranges.empty()) out.ranges.push_back(0);
    // backtrack-stack bound: the program is a DAG, so the deepest stack over all paths is computable
    std::vector<int> depth(out.code.size(), -1);
    std::function<int(size_t)> need = [&](size_t pc) -> int {
        if (pc >= out.code.size()) return 0;
        if (depth[pc] >= 0) return depth[pc];
        const uint32_t op = out.code[pc].op_a & 0xFFu, a = out.code[pc].op_a >> 8;
        int d;
        if (op == VM_MATCH) d = 0;
        else if (op == VM_SPLIT) d = std::max(1 + need(a), need(out.code[pc].b));
        else if (op == VM_JMP) d = need(a);
        else if (op == VM_LOOP) d = ((out.code[pc].b >> 24) & 1u ? 0 : 1) + need(pc + 1);
        else d = need(pc + 1);
        return depth[pc] = d;
    };
    if (need(0) > kVmStack) { err = "RegexSplit: the pattern nests more alternatives / quantifiers than the GPU matcher's backtrack stack holds: " + pattern; return B200TOK_E_UNSUPPORTED; }
    return B200TOK_OK;
}

}  // namespace b200tok
