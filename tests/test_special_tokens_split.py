"""SpecialTokensSplit.  CPU tier: the oracle (system PCRE2) against the reference's own known-answer vectors
(tests/golden/special_tokens_split_layer_tests.json, from the reference's tests/layer_tests.py:405-457, patterns built by the
reference's converter code) and the pattern builder restatement.  GPU tier: the CUDA op through the C ABI against the oracle."""
import json
from pathlib import Path

import numpy as np
import pytest

import cases
from openvino_tokenizers_b200.strings import add_ragged_dimension, pack_strings, unpack_strings

GOLDEN = json.loads((Path(__file__).resolve().parent / "golden" / "special_tokens_split_layer_tests.json").read_text())


def _pieces(res, chars):
    return [p.decode() for p in unpack_strings(res[2], res[3], chars)], [int(x) for x in res[-1]]


def test_oracle_reference_vectors(oracle_mod):
    for c in GOLDEN["cases"]:
        assert oracle_mod.special_tokens_pattern(c["tokens"]) == c["pattern"]
        b, e, ch = pack_strings([c["text"]])
        rb, re_ = add_ragged_dimension(b, e)
        r = oracle_mod.SpecialTokensSplitOracle(c["pattern"])(rb, re_, b, e, ch)
        assert _pieces(r, ch) == (c["expected"], c["expected_skips"]), c["text"]


def test_empty_strings_yield_no_pieces(oracle_mod):
    b, e, ch = pack_strings(["", "<s>", ""])
    rb, re_ = add_ragged_dimension(b, e)
    r = oracle_mod.SpecialTokensSplitOracle(oracle_mod.special_tokens_pattern([("<s>", False, False)]))(rb, re_, b, e, ch)
    assert r[0].tolist() == [0, 0, 1] and r[1].tolist() == [0, 1, 1] and r[4].tolist() == [1]


@pytest.fixture(scope="module")
def ops():
    from openvino_tokenizers_b200 import ops as O
    return O


def _run_both(ops, oracle_mod, pattern, texts, skips=None, rows=None):
    b, e, ch = pack_strings(texts)
    if rows is None:
        rb, re_ = add_ragged_dimension(b, e)
    else:
        rb, re_ = rows
    exp = oracle_mod.SpecialTokensSplitOracle(pattern)(rb, re_, b, e, ch, skips)
    pat = np.frombuffer(pattern.encode(), np.uint8)
    ins = [rb, re_, b, e, ch] + ([np.asarray(skips, bool)] if skips is not None else []) + [pat]
    got = ops.SpecialTokensSplit().evaluate(ins)
    for k in (0, 1, 2, 3):
        assert np.array_equal(got[k], exp[k]), (k, texts[:3])
    assert np.array_equal(got[5], exp[4].astype(bool))
    return got


@pytest.mark.gpu
def test_gpu_reference_vectors(ops, oracle_mod):
    for c in GOLDEN["cases"]:
        got = _run_both(ops, oracle_mod, c["pattern"], [c["text"]])
        assert _pieces(got, got[4]) == (c["expected"], c["expected_skips"]), c["text"]


@pytest.mark.gpu
def test_gpu_random_texts_and_token_sets(ops, oracle_mod):
    rng = np.random.default_rng(21)
    token_sets = [
        [("<|endoftext|>", False, False), ("<|im_start|>", False, False), ("<|im_end|>", False, True)],
        [("<s>", False, False), ("</s>", True, False), ("<s>x", False, False), ("<unk>", True, True), ("[MASK]", True, False)],
        [("    ", False, False), ("def", True, True), (" ", True, False)],              # whitespace tokens: full backtracking
        [("▁", False, False), ("<｜begin▁of▁sentence｜>", False, True), ("　　", True, False)],
        [("a", False, False), ("ab", False, False), ("abc", True, False), ("b", False, True)],   # prefixes: first alternative wins
    ]
    frag = ["<|endoftext|>", "<|im_start|>", "<|im_end|>", "<s>", "</s>", "<s>x", "<unk>", "[MASK]", "    ", "def", " ", "  ", "\n", "\t",
            "▁", "<｜begin▁of▁sentence｜>", "　", " ", "a", "b", "c", "abc", "<", "|", ">", "x", "hello", "Ж", "\U0001F600"]
    for toks in token_sets:
        pattern = oracle_mod.special_tokens_pattern(toks)
        texts = ["".join(rng.choice(frag, size=int(rng.integers(0, 40)))) for _ in range(300)] + ["", " ", "   "]
        _run_both(ops, oracle_mod, pattern, texts)
    # long rows and the edge corpus
    pattern = oracle_mod.special_tokens_pattern(token_sets[0] + token_sets[1])
    _run_both(ops, oracle_mod, pattern, [s for s in cases.EDGE_STRINGS] + cases.long_prompts())
    _run_both(ops, oracle_mod, pattern, ["x" * 700 + "<s>" + " " * 300 + "</s>" + "y" * 1000])


@pytest.mark.gpu
def test_gpu_skips_and_multi_element_rows(ops, oracle_mod):
    pattern = oracle_mod.special_tokens_pattern([("<s>", False, False), ("</s>", True, True)])
    texts = ["a<s>b", "<s>", "keep <s> whole", "", "x </s> y", "tail"]
    rb, re_ = np.array([0, 2, 2, 5], np.int32), np.array([2, 2, 5, 6], np.int32)        # rows of 2, 0, 3 and 1 elements
    skips = np.array([0, 0, 1, 0, 0, 1], bool)
    got = _run_both(ops, oracle_mod, pattern, texts, skips, (rb, re_))
    assert got[0].tolist() == [0, 4, 4, 8] and got[1].tolist() == [4, 4, 8, 9]
    _run_both(ops, oracle_mod, pattern, texts, None, (rb, re_))


@pytest.mark.gpu
def test_gpu_unsupported_pattern_is_an_error(ops):
    from openvino_tokenizers_b200._capi import B200TokError, E_UNSUPPORTED
    for bad in [r"(\d+)", r"(a)(b)", r"(?:\s*)(a)+", r"(a||b)", ""]:
        with pytest.raises(B200TokError) as ei:
            ops.SpecialTokensSplit().with_pattern(bad)
        assert ei.value.code == E_UNSUPPORTED


def test_host_matcher_equals_pcre2(oracle_mod):
    """CPU tier: the product's pattern parser + per-position matcher (tok_core.cuh, compiled for the host by tests/harness)
    against the oracle's PCRE2 capture-group scan — the reference vectors, exhaustive short strings over a small alphabet for
    every strip combination, and random texts with whitespace / prefix-related / multi-byte tokens."""
    import itertools
    import hostcore as H

    def oracle_pieces(o, s):
        b, e, ch = pack_strings([s])
        rb, re_ = add_ragged_dimension(b, e)
        r = o(rb, re_, b, e, ch)
        return list(zip(r[2].tolist(), r[3].tolist(), r[4].tolist()))

    for c in GOLDEN["cases"]:
        o = oracle_mod.SpecialTokensSplitOracle(c["pattern"])
        assert H.special_split(c["pattern"], c["text"].encode()) == oracle_pieces(o, c["text"].encode()), c["text"]
    small = ["a", "b", " ", "\n", "<", ">", "x"]
    for sl, sr in itertools.product((False, True), repeat=2):
        pattern = oracle_mod.special_tokens_pattern([("<a>", sl, sr), ("ab", sl, sr), ("a", sl, sr), (" b", sl, sr), ("x", not sl, sr)])
        o = oracle_mod.SpecialTokensSplitOracle(pattern)
        for L in range(0, 6):
            for tup in itertools.product(small, repeat=L):
                s = "".join(tup).encode()
                assert H.special_split(pattern, s) == oracle_pieces(o, s), (pattern, s)
    rng = np.random.default_rng(77)
    sets = [
        [("<|endoftext|>", False, False), ("<|im_start|>", True, False), ("<|im_end|>", False, True), ("<|im", True, True)],
        [("    ", False, False), ("def", True, True), (" ", True, False), ("\n\n", False, True)],
        [("▁", False, False), ("<｜begin▁of▁sentence｜>", True, True), ("　　", True, False), ("é", False, True)],
    ]
    frag = ["<|endoftext|>", "<|im_start|>", "<|im_end|>", "<|im", "    ", "def", " ", "  ", "\n", "\n\n", "\t", "▁", "<｜begin▁of▁sentence｜>", "　", "é",
            "a", "b", "<", "|", ">", "hello", "Ж", "\U0001F600", " "]
    for toks in sets:
        pattern = oracle_mod.special_tokens_pattern(toks)
        o = oracle_mod.SpecialTokensSplitOracle(pattern)
        for _ in range(3000):
            s = "".join(rng.choice(frag, size=int(rng.integers(0, 30)))).encode()
            assert H.special_split(pattern, s) == oracle_pieces(o, s), (pattern, s)
