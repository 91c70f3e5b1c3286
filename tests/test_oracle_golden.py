"""CPU tier: pin the oracle — against the reference's own known-answer vectors (RegexSplit) and against
HuggingFace `tokenizers` outputs on the frozen vocabularies (BPE / WordPiece), both committed under tests/golden."""
import json
from pathlib import Path

import numpy as np
import pytest

import cases
from openvino_tokenizers_b200 import assets as A
from openvino_tokenizers_b200.strings import pack_strings, unpack_strings

GOLDEN = Path(__file__).resolve().parent / "golden"


def test_regex_split_reference_vectors(oracle_mod):
    g = json.loads((GOLDEN / "regex_split_layer_tests.json").read_text())
    assert len(g["cases"]) == 33
    for case in g["cases"]:
        o = oracle_mod.SplitOracle(case["pattern"], case["behaviour"], case["invert"], case["max_splits"])
        rb, re_, b, e, c = cases.batch_from_strings([case["text"]])
        r = o(rb, re_, b, e, c)
        assert [p.decode() for p in unpack_strings(r[2], r[3], c)] == case["expected"], case


@pytest.mark.parametrize("name", ["gpt2_synth", "llama3_synth"])
def test_bpe_oracle_matches_hf(oracle_mod, name):
    a = A.load_bpe(name)
    v, ml, mr, ad, aid = a.tensors()
    g = json.loads((GOLDEN / f"hf_{name}.json").read_text())
    split = oracle_mod.SplitOracle(a.split_pattern, "isolate")
    for use_cache in (True, False):
        bpe = oracle_mod.BpeOracle(v, ml, mr, ad, aid, cache_capacity=a.cache_capacity, use_cache=use_cache)
        batch = cases.batch_from_strings(g["texts"])
        s = split(*batch)
        ob, oe, ids = bpe(s[0], s[1], s[2], s[3], batch[4])
        for i, exp in enumerate(g["ids"]):
            assert ids[ob[i]:oe[i]].tolist() == exp, g["texts"][i][:50]


def test_wordpiece_oracle_matches_hf(oracle_mod):
    a = A.load_wordpiece("bert_synth")
    g = json.loads((GOLDEN / "hf_bert_synth.json").read_text())
    v = pack_strings(a.vocab)
    s1 = oracle_mod.SplitOracle(A.BERT_WHITESPACE_PATTERN, "remove")
    s2 = oracle_mod.SplitOracle(A.BERT_PUNCT_PATTERN, "isolate")
    wp = oracle_mod.WordpieceOracle(v, a.suffix_indicator, a.max_bytes_per_word)
    texts = [t for t in g["texts"] if t != ""] or ["x"]
    batch = cases.batch_from_strings(g["texts"])
    if len(batch[4]) == 0:
        pytest.skip("empty corpus")
    r1 = s1(*batch)
    r2 = s2(r1[0], r1[1], r1[2], r1[3], batch[4])
    ob, oe, ids = wp(r2[0], r2[1], r2[2], r2[3], batch[4], a.unk_token_id)
    for i, exp in enumerate(g["ids"]):
        assert ids[ob[i]:oe[i]].tolist() == exp, g["texts"][i][:50]


def test_multithreaded_oracle_is_identical(oracle_mod):
    a = A.load_bpe("gpt2_synth")
    v, ml, mr, ad, aid = a.tensors()
    split = oracle_mod.SplitOracle(a.split_pattern, "isolate")
    bpe = oracle_mod.BpeOracle(v, ml, mr, ad, aid)
    batch = cases.random_ascii_batch(512, 128)
    s1 = split(*batch)
    s4 = split(*batch, threads=4)
    for x, y in zip(s1, s4):
        assert np.array_equal(x, y)
    r1 = bpe(s1[0], s1[1], s1[2], s1[3], batch[4])
    r4 = bpe(s1[0], s1[1], s1[2], s1[3], batch[4], threads=4)
    assert cases.ragged_rows_equal(r1, r4)


def test_vocab_decoder_and_byte_fallback_hand_vectors(oracle_mod):
    """No reference test exists for these ops ("parity unpinned"); hand-computed vectors."""
    vocab = pack_strings([b"<unk>", b"<0x41>", b"hi", b"", "▁x".encode()])
    ids = np.array([[2, 1, 0, 7, -1, 4]], np.int32)
    rb, re_, b, e, c = oracle_mod.vocab_decoder(ids, vocab, [0])
    assert rb.tolist() == [0] and re_.tolist() == [6]
    assert unpack_strings(b, e, c) == [b"hi", b"<0x41>", b"", b"", b"", "▁x".encode()]
    b2, e2, c2 = oracle_mod.byte_fallback(b, e, c)
    assert unpack_strings(b2, e2, c2) == [b"hi", b"A", b"", b"", b"", "▁x".encode()]
    t = pack_strings([b"<0xZZ>", b"<0x4a>", b"<0x4A>", b"<<x41>"])
    b3, e3, c3 = oracle_mod.byte_fallback(*t)
    assert unpack_strings(b3, e3, c3) == [b"\xff", b"\xff", b"J", b"<<x41>"]
