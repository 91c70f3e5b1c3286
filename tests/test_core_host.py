"""CPU tier: the product's algorithmic core (csrc/tok_core.cuh + csrc/tables.cpp, the code the kernels run)
compiled for the host by tests/harness and compared with the oracle."""
import numpy as np
import pytest

import cases
import hostcore as H
from openvino_tokenizers_b200 import assets as A
from openvino_tokenizers_b200.strings import pack_strings

SPLIT_CASES = [
    (A.GPT2_PATTERN, "isolate", False), (A.GPT2_DIGITS_PATTERN, "isolate", False), (A.LLAMA3_PATTERN, "isolate", False),
    (A.LLAMA3_PATTERN, "contiguous", False), (r"\s+", "remove", False), (A.BERT_PUNCT_PATTERN, "isolate", False),
    (r"\w+|[^\w\s]+", "remove", True), (".", "isolate", False), ("▁", "mergedwithnext", False),
    ("▁", "mergedwithprevious", False), (r"\p{N}", "isolate", False), (r"\p{P}", "contiguous", False),
    (r"\s+", "mergedwithprevious", True), (r"\s+", "mergedwithnext", True), (r"\p{Nd}|\p{Nl}|\p{No}", "remove", False),
]
ALPHA = ["a", "s", "t", "'", "l", "r", "e", "v", "1", "2", " ", " ", "\n", "\r", "\t", "!", "?", "é", "ſ", " ", "測", "😁",
         "▁", "_", "S", "L", "٣"]


def _oracle_split(o, data):
    rb, re_, b, e, c = cases.batch_from_strings([data])
    r = o(rb, re_, b, e, c)
    return list(zip(r[2].tolist(), r[3].tolist()))


@pytest.mark.parametrize("pattern,behaviour,invert", SPLIT_CASES)
def test_matcher_equals_pcre2(oracle_mod, pattern, behaviour, invert):
    o = oracle_mod.SplitOracle(pattern, behaviour, invert)
    rng = np.random.default_rng(0)
    for _ in range(1500):
        s = "".join(rng.choice(ALPHA, size=int(rng.integers(1, 14)))).encode()
        assert H.split(pattern, behaviour, invert, -1, s) == _oracle_split(o, s), s
    for s in cases.EDGE_STRINGS:
        if s:
            assert H.split(pattern, behaviour, invert, -1, s.encode()) == _oracle_split(o, s.encode()), s


def test_max_splits_quirk(oracle_mod):
    for ms in (1, 2, 3):
        o = oracle_mod.SplitOracle(r"\s+", "remove", False, ms)
        for s in (b"a b c d e", b" a  b ", b"abc"):
            assert H.split(r"\s+", "remove", False, ms, s) == _oracle_split(o, s)


def test_unknown_pattern_rejected():
    with pytest.raises(ValueError):
        H.split(r"(foo|bar)+baz", "isolate", False, -1, b"abc")


@pytest.mark.parametrize("name", ["gpt2_synth", "llama3_synth"])
def test_bpe_merge_loops_equal_oracle(oracle_mod, name):
    a = A.load_bpe(name)
    v, ml, mr, ad, aid = a.tensors()
    o = oracle_mod.BpeOracle(v, ml, mr, ad, aid, use_cache=False)
    h = H.HostBpe(v, ml, mr, ad, aid)
    assert h.info(0) == 0          # no two merges produce the same token: the (rank, birth) tie is unreachable
    rng = np.random.default_rng(1)
    words = cases.long_prompts()[0].encode().split()
    pieces = [b" " + w for w in words] + words
    pieces += [bytes(rng.integers(0x20, 0x7F, size=int(rng.integers(1, 40)), dtype=np.uint8)) for _ in range(1500)]
    pieces += ["Тест".encode(), " 測試".encode(), "😁😁".encode(), b"a" * 300, b" " * 256, b"<|endoftext|>",
               b"<|endoftext", b"ab<|endoftext|>cd", bytes(range(256)), b"ab" * 500]
    b, e, c = pack_strings(pieces)
    rb = np.arange(len(pieces), dtype=np.int32)
    ob, oe, ids = o(rb, rb + 1, b, e, c)
    for i, p in enumerate(pieces):
        exp = ids[ob[i]:oe[i]].tolist()
        assert h.piece(p, 0) == exp and h.piece(p, 1) == exp and h.piece(p, 2) == exp, p[:30]


def test_bpe_suffix_unk_fallback_tables(oracle_mod):
    vocab = ["<unk>", "a", "b", "c", "</w>", "ab", "abc", "c</w>", "bc</w>", "<0x64>", "<0x65>", "d</w>", "ab</w>"]
    merges = ["a b", "ab c", "c </w>", "b c</w>", "ab </w>"]
    v, mg = pack_strings(vocab), pack_strings(merges)
    words = [b"abc", b"ab", b"abcd", b"xabc", b"de", b"cab", b"abab" * 20]
    b, e, c = pack_strings(words)
    rb = np.arange(len(words), dtype=np.int32)
    for bf in (False, True):
        for unk in (b"<unk>", b""):
            o = oracle_mod.BpeOracle(v, mg, None, unk_token=unk, end_suffix=b"</w>", byte_fallback=bf, use_cache=False)
            ob, oe, ids = o(rb, rb + 1, b, e, c)
            h = H.HostBpe(v, mg, None, unk_token=unk, end_suffix=b"</w>", byte_fallback=bf)
            for i, w in enumerate(words):
                for mode in (0, 1, 2):
                    assert h.piece(w, mode) == ids[ob[i]:oe[i]].tolist(), (w, bf, unk, mode)


def _tie_vocab(seed):
    """A vocabulary over {a, b} in which many tokens are the product of several merges: every string up to six
    characters is a token, the merges are a random selection of (left, right) splits in random order."""
    import itertools
    rng = np.random.default_rng(seed)
    toks = ["".join(t) for n in range(1, 7) for t in itertools.product("ab", repeat=n)]
    splits = [(t[:k], t[k:]) for t in toks for k in range(1, len(t))]
    order = rng.permutation(len(splits))[: int(len(splits) * 0.7)]
    return toks, [splits[i][0] + " " + splits[i][1] for i in order]


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_heap_ties_pop_like_std_priority_queue(oracle_mod, seed):
    """Two merges producing one token let queue entries tie on (rank, seq); the reference then follows libstdc++'s heap layout
    (src/bpe_tokenizer.cpp:166-183).  bpe_merge_heap restates std::push_heap / std::pop_heap operation for operation and must give the
    oracle's result (the oracle runs std::priority_queue itself; tests/test_reference_pin.py ties it to the reference's compiled code);
    the packed loop breaks ties left pair first and must REPORT the tie wherever its result differs."""
    toks, merges = _tie_vocab(seed)
    v, mg = pack_strings(toks), pack_strings(merges)
    o = oracle_mod.BpeOracle(v, mg, None, use_cache=False)
    h = H.HostBpe(v, mg, None)
    assert h.info(0) > 0                       # tokens with more than one producing merge: tie_check is on for this vocabulary
    rng = np.random.default_rng(100 + seed)
    words = [bytes(rng.choice([97, 98], size=int(rng.integers(2, 48))).astype(np.uint8)) for _ in range(4000)]
    words += [b"a" * n for n in range(2, 40)] + [b"ab" * n for n in range(1, 24)] + [b"aab" * n for n in range(1, 16)]
    b, e, c = pack_strings(words)
    rb = np.arange(len(words), dtype=np.int32)
    ob, oe, ids = o(rb, rb + 1, b, e, c)
    differ = 0
    for i, w in enumerate(words):
        exp = ids[ob[i]:oe[i]].tolist()
        assert h.piece(w, 1) == exp, w         # exact, ties included
        packed, tie = h.piece_packed_with_tie(w)
        if packed != exp:
            differ += 1
            assert tie, w                      # a left-first result that differs from the reference's is always flagged
    if refops_available():
        import refops
        r = refops.bpe(v, mg)                  # the reference's own compiled BPETokenizer on the same pieces
        rb_, re_, rids = r(rb, rb + 1, b, e, c)
        assert np.array_equal(rids, ids) and np.array_equal(rb_, ob) and np.array_equal(re_, oe)
    print(f"seed {seed}: left-first differs from the reference on {differ} of {len(words)} pieces")


def refops_available():
    try:
        import refops
        return refops.available()
    except Exception:
        return False


def test_birth_order_differs_from_position_order(oracle_mod):
    """equal ranks are ordered by push sequence, not by position."""
    vocab = ["u", "v", "q", "p", "X", "XX", "pq"]
    vocab = ["u", "v", "q", "uv", "uvq", "uvquvq"]
    merges = ["u v", "uv q", "uvq uvq"]
    v, mg = pack_strings(vocab), pack_strings(merges)
    o = oracle_mod.BpeOracle(v, mg, None, use_cache=False)
    h = H.HostBpe(v, mg, None)
    for w in (b"uvquvquvq", b"uvquvq", b"uvquvquvquvq", b"quvquvquv"):
        b, e, c = pack_strings([w])
        z = np.zeros(1, np.int32)
        exp = o(z, z + 1, b, e, c)[2].tolist()
        for mode in (0, 1, 2):
            assert h.piece(w, mode) == exp, (w, mode)


def test_wordpiece_word_equals_oracle(oracle_mod):
    a = A.load_wordpiece("bert_synth")
    v = pack_strings(a.vocab)
    o = oracle_mod.WordpieceOracle(v, a.suffix_indicator, a.max_bytes_per_word)
    h = H.HostWordpiece(v, a.suffix_indicator, a.max_bytes_per_word)
    words = [w.lower() for w in cases.long_prompts()[0].encode().split()]
    words += [b"a" * 100, b"a" * 101, b"unaffable", b"xyzzyqq", "тест".encode(), b"##", b"#"]
    b, e, c = pack_strings(words)
    rb = np.arange(len(words), dtype=np.int32)
    ob, oe, ids = o(rb, rb + 1, b, e, c, a.unk_token_id)
    for i, w in enumerate(words):
        assert h.word(w, a.unk_token_id) == ids[ob[i]:oe[i]].tolist(), w


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_wordpiece_jump_tables_on_random_vocabularies(oracle_mod, seed):
    """The walks start from val1 / the two-byte jump table (tok_core.cuh rank_trie_longest): random vocabularies over a small alphabet
    that includes non-ASCII bytes (which must fall back to the root lookup), with and without one-, two- and three-byte tokens, every
    word up to four bytes exhaustively plus random longer ones, against the oracle and the reference's compiled WordpieceTokenizer."""
    import itertools
    rng = np.random.default_rng(seed)
    alpha = [b"a", b"b", b"c", b"\xc3", b"\xa9", b"1"]
    pool = [b"".join(t) for n in range(1, 5) for t in itertools.product(alpha, repeat=n)]
    keep = rng.random(len(pool)) < [0.9, 0.5, 0.25, 0.1][seed]
    root = [t for t, k in zip(pool, keep) if k]
    keep2 = rng.random(len(pool)) < 0.35
    sub = [b"##" + t for t, k in zip(pool, keep2) if k]
    vocab = [b"[UNK]"] + root + sub
    rng.shuffle(vocab)
    unk = vocab.index(b"[UNK]")
    v = pack_strings(vocab)
    o = oracle_mod.WordpieceOracle(v, b"##", 100)
    h = H.HostWordpiece(v, b"##", 100)
    words = pool + [b"".join(rng.choice(alpha, size=int(rng.integers(5, 14)))) for _ in range(1500)]
    b, e, c = pack_strings(words)
    rb = np.arange(len(words), dtype=np.int32)
    ob, oe, ids = o(rb, rb + 1, b, e, c, unk)
    for i, w in enumerate(words):
        assert h.word(w, unk) == ids[ob[i]:oe[i]].tolist(), w
    if refops_available():
        import refops
        r = refops.wordpiece(v, unk, b"##", 100)
        rb_, re_, rids = r(rb, rb + 1, b, e, c)
        assert np.array_equal(rids, ids) and np.array_equal(rb_, ob) and np.array_equal(re_, oe)


def test_class_table_spot_checks():
    L, N, S, P, W, BP, NL = 1, 2, 4, 8, 16, 32, 64
    assert H.lib().hz_char_class(ord("a")) & L
    assert H.lib().hz_char_class(ord("7")) & N
    assert H.lib().hz_char_class(0x20) & S and H.lib().hz_char_class(0xA0) & S and H.lib().hz_char_class(0x3000) & S
    assert H.lib().hz_char_class(ord("\n")) & NL and H.lib().hz_char_class(ord("\n")) & S
    assert H.lib().hz_char_class(ord("!")) & BP and H.lib().hz_char_class(0x4E2D) & BP and not H.lib().hz_char_class(ord("a")) & BP
    assert H.lib().hz_char_class(ord("$")) & BP and not H.lib().hz_char_class(ord("$")) & P   # BERT's ASCII ranges beyond \p{P}
    assert H.lib().hz_char_class(0x0416) & L and H.lib().hz_char_class(0x0663) & N and H.lib().hz_char_class(ord("_")) & W


@pytest.mark.parametrize("digits", [False, True])
def test_gpt2_closed_form_equals_pcre2(oracle_mod, digits):
    """The per-position piece-start predicate the GPT-2 window splitter uses, against PCRE2: exhaustive over short
    strings of a small alphabet plus random strings of a wide one."""
    import itertools
    pat = A.GPT2_DIGITS_PATTERN if digits else A.GPT2_PATTERN
    o = oracle_mod.SplitOracle(pat, "isolate")
    small = ["a", "'", "s", "r", "e", "l", " ", "\n", "1", "!"]
    for L in range(1, 5):
        for tup in itertools.product(small, repeat=L):
            s = "".join(tup).encode()
            assert H.gpt2_closed_form(s, digits) == _oracle_split(o, s), s
    rng = np.random.default_rng(3)
    wide = ALPHA + ["m", "d", "'", "'"]
    for _ in range(5000):
        s = "".join(rng.choice(wide, size=int(rng.integers(1, 18)))).encode()
        assert H.gpt2_closed_form(s, digits) == _oracle_split(o, s), s
    for s in cases.EDGE_STRINGS + cases.long_prompts():
        if s:
            assert H.gpt2_closed_form(s.encode(), digits) == _oracle_split(o, s.encode()), s[:40]


@pytest.mark.parametrize("digits", [False, True])
def test_gpt2_neighbour_form_equals_closed_form(digits):
    """The branch-free neighbour form the fused ASCII window pass evaluates == the closed form (== PCRE2, above)."""
    import itertools
    small = ["a", "'", "s", "r", "e", "l", " ", "\n", "1", "!", "v", "d"]
    for L in range(1, 5):
        for tup in itertools.product(small, repeat=L):
            s = "".join(tup).encode()
            assert H.gpt2_neighbour_form(s, digits) == H.gpt2_closed_form(s, digits), s
    rng = np.random.default_rng(5)
    for _ in range(5000):
        s = bytes(rng.integers(0x09, 0x7F, size=int(rng.integers(1, 40)), dtype=np.uint8))
        assert H.gpt2_neighbour_form(s, digits) == H.gpt2_closed_form(s, digits), s


@pytest.mark.parametrize("digits", [False, True])
def test_gpt2_word_form_equals_closed_form(digits):
    """The word (bit-mask) form the window kernel evaluates == the closed form (== PCRE2, above); subjects longer than one
    32-position word exercise the carries, UTF-8 subjects the continuation-byte and multi-byte-whitespace handling."""
    import itertools
    small = ["a", "'", "s", "r", "e", "l", " ", "\n", "1", "!", "v", "d"]
    for L in range(1, 5):
        for tup in itertools.product(small, repeat=L):
            s = "".join(tup).encode()
            assert H.gpt2_word_form(s, digits) == H.gpt2_closed_form(s, digits), s
    rng = np.random.default_rng(7)
    for _ in range(4000):
        s = bytes(rng.integers(0x09, 0x7F, size=int(rng.integers(1, 200)), dtype=np.uint8))
        assert H.gpt2_word_form(s, digits) == H.gpt2_closed_form(s, digits), s
    dense = ["a", "b", "'", "'", "s", "t", "m", "d", "r", "e", "v", "l", " ", " ", "\n", "\t", "1", "2", "!", "?"]
    uni = dense + [chr(0xA0), chr(0x2003), chr(0x3000), chr(0x416), chr(0x4E2D), chr(0x1F600), chr(0x663), chr(0xE9), chr(0x85), chr(0x1680)]
    for alpha in (dense, uni):
        for _ in range(4000):
            s = "".join(rng.choice(alpha, size=int(rng.integers(1, 120)))).encode()
            assert H.gpt2_word_form(s, digits) == H.gpt2_closed_form(s, digits), s
    for s in cases.EDGE_STRINGS + cases.long_prompts():
        if s:
            assert H.gpt2_word_form(s.encode(), digits) == H.gpt2_closed_form(s.encode(), digits), s[:40]


def test_llama3_word_form_equals_pcre2(oracle_mod):
    """The word (bit-mask) form of the Llama-3 / cl100k pattern that the window kernel evaluates, against PCRE2: exhaustive over
    short strings of a small alphabet, random strings of wider ones (long enough to cross 32-position words), UTF-8 subjects."""
    import itertools
    o = oracle_mod.SplitOracle(A.LLAMA3_PATTERN, "isolate")
    small = ["a", "'", "s", "L", " ", "\n", "\t", "1", "!", "\r"]
    for L in range(1, 6):
        for tup in itertools.product(small, repeat=L):
            s = "".join(tup).encode()
            assert H.llama3_word_form(s) == _oracle_split(o, s), s
    rng = np.random.default_rng(17)
    dense = ["a", "b", "'", "'", "s", "t", "m", "d", "r", "e", "v", "l", "S", "T", "R", "E", "L", " ", " ", " ", "\n", "\n", "\t", "\r",
             "1", "2", "3", "!", "?", "-"]
    uni = dense + [chr(0xA0), chr(0x2003), chr(0x3000), chr(0x416), chr(0x4E2D), chr(0x1F600), chr(0xE9), chr(0x85), chr(0x17F), chr(0xFF0C), chr(0x2028)]
    for alpha in (dense, uni):
        for _ in range(6000):
            s = "".join(rng.choice(alpha, size=int(rng.integers(1, 150)))).encode()
            assert H.llama3_word_form(s) == _oracle_split(o, s), s
    for _ in range(3000):
        s = bytes(rng.integers(0x09, 0x7F, size=int(rng.integers(1, 200)), dtype=np.uint8))
        assert H.llama3_word_form(s) == _oracle_split(o, s), s
    for _ in range(300):      # long runs: digits, newlines, spaces crossing several words
        parts = [rng.choice(["7", "\n", " ", "a", "!", "\t"]) * int(rng.integers(1, 90)) for _ in range(int(rng.integers(1, 8)))]
        s = "".join(parts).encode()
        assert H.llama3_word_form(s) == _oracle_split(o, s), s
    for s in cases.EDGE_STRINGS + cases.long_prompts():
        if s:
            got = H.llama3_word_form(s.encode())
            if got is not None:      # (None: a non-ASCII digit — such rows go to the generic kernel)
                assert got == _oracle_split(o, s.encode()), s[:40]


def test_contiguous_batch_predicate():
    """tables.cpp contiguous_batch (the qualification of the pipelined host-buffer path, vectorised) against the plain predicate."""
    import ctypes as C
    L = H.lib()
    i32p = C.POINTER(C.c_int32)

    def plain(rb, re_, eb, ee, N):
        B, E = len(rb), len(eb)
        if rb[0] != 0 or re_[B - 1] != E or re_[B - 1] < rb[B - 1]:
            return False
        if np.any(rb[1:] != re_[:-1]) or np.any(re_[:-1] < rb[:-1]):
            return False
        return not (np.any(eb < 0) or np.any(ee < eb) or np.any(ee > N) or np.any(eb[1:] < ee[:-1]))

    def got(rb, re_, eb, ee, N):
        a = [np.ascontiguousarray(x, np.int32) for x in (rb, re_, eb, ee)]
        return bool(L.hz_contiguous_batch(*(x.ctypes.data_as(i32p) for x in a), C.c_int64(len(rb)), C.c_int64(len(eb)), C.c_int64(N)))

    rng = np.random.default_rng(3)
    for trial in range(300):
        B = int(rng.integers(1, 70))
        per = rng.integers(0, 4, size=B)                       # elements per row (rows may be empty)
        E = int(per.sum())
        re_ = np.cumsum(per).astype(np.int32)
        rb = (re_ - per).astype(np.int32)
        lens = rng.integers(0, 9, size=E)
        gaps = rng.integers(0, 3, size=E)                      # gaps between elements are fine
        eb = (np.cumsum(lens + gaps) - lens).astype(np.int32)
        ee = (eb + lens).astype(np.int32)
        N = int(ee[-1]) + int(rng.integers(0, 5)) if E else int(rng.integers(0, 5))
        assert got(rb, re_, eb, ee, N) == plain(rb, re_, eb, ee, N) == True
        if E == 0:
            continue
        for _ in range(6):                                     # one perturbation at a time: both must agree (mostly: reject)
            arrs = [rb.copy(), re_.copy(), eb.copy(), ee.copy()]
            k = int(rng.integers(0, 4))
            i = int(rng.integers(0, len(arrs[k])))
            arrs[k][i] += int(rng.choice([-9, -1, 1, 9, 1000]))
            assert got(*arrs, N) == plain(*arrs, N), (trial, k, i)
