"""CPU tier: pin the oracle to the REFERENCE'S OWN CODE.  oracle/_ref/libovtok_ref.so is the reference's op sources
(src/bpe_tokenizer.cpp, wordpiece_tokenizer.cpp, vocab_encoder.cpp, vocab_decoder.cpp, byte_fallback.cpp, regex_split.cpp,
utils.cpp, special_tokens_split.cpp, truncate.cpp, combine_segments.cpp, ragged_to_dense.cpp, fuze.cpp, bytes_to_chars.cpp,
chars_to_bytes.cpp, utf8_validate.cpp, regex_normalization.cpp) compiled unmodified against a stand-in OpenVINO API
(oracle/Makefile target `_ref`); every test here runs the reference's evaluate() and the restatement (oracle/oracle.cpp)
on the same inputs and demands identical tensors.  The GPU parity tests compare with the restatement, so this closes the
chain GPU == oracle.cpp == reference code."""
import json
from pathlib import Path

import numpy as np
import pytest

import cases
import refops
from openvino_tokenizers_b200 import assets as A
from openvino_tokenizers_b200.strings import pack_strings, unpack_strings

GOLDEN = Path(__file__).resolve().parent / "golden"
pytestmark = pytest.mark.skipif(not refops.available(), reason="oracle/_ref not built and /root/reference absent")


def eq(a, b, what=""):
    assert len(a) == len(b), what
    for i, (x, y) in enumerate(zip(a, b)):
        if x is None or y is None:
            continue
        assert np.array_equal(np.asarray(x), np.asarray(y)), f"{what}: output {i} differs"


def corpus():
    return cases.EDGE_STRINGS + cases.long_prompts()


# ------------------------------------------------------------------ RegexSplit
def test_reference_regex_split_golden_vectors():
    """The reference's own 33 known-answer vectors through the reference's own compiled RegexSplit (sanity of the build)."""
    g = json.loads((GOLDEN / "regex_split_layer_tests.json").read_text())
    for case in g["cases"]:
        run = refops.regex_split(case["pattern"], case["behaviour"], case["invert"], case["max_splits"], with_skips=False)
        rb, re_, b, e, c = cases.batch_from_strings([case["text"]])
        r = run(rb, re_, b, e, c)
        assert [p.decode() for p in unpack_strings(r[2], r[3], c)] == case["expected"], case


PATTERNS = [A.GPT2_PATTERN, A.GPT2_DIGITS_PATTERN, A.LLAMA3_PATTERN, A.BERT_WHITESPACE_PATTERN, A.BERT_PUNCT_PATTERN, r"\w+|[^\w\s]+", r"\.", "ab"]


@pytest.mark.parametrize("behaviour", ["remove", "isolate", "contiguous", "mergedwithprevious", "mergedwithnext"])
@pytest.mark.parametrize("invert", [False, True])
def test_regex_split_oracle_equals_reference(oracle_mod, behaviour, invert):
    batch = cases.batch_from_strings(corpus())
    rnd = cases.mixed_utf8_batch(48, 192, seed=5)
    for pat in PATTERNS:
        for max_splits in (-1, 3):
            if max_splits != -1 and behaviour == "contiguous":
                continue
            o = oracle_mod.SplitOracle(pat, behaviour, invert, max_splits)
            r = refops.regex_split(pat, behaviour, invert, max_splits)
            for bt in (batch, rnd):
                sk = (np.arange(len(bt[2])) % 5 == 2)
                eq(o(*bt, skips=sk.astype(np.uint8)), r(*bt, skips=sk), f"{pat!r} {behaviour} invert={invert} max={max_splits}")


def test_regex_split_whole_batch_empty(oracle_mod):
    bt = cases.batch_from_strings(["", "", ""])
    o = oracle_mod.SplitOracle(A.GPT2_PATTERN, "isolate")(*bt)
    r = refops.regex_split(A.GPT2_PATTERN, "isolate", with_skips=False)(*bt)
    assert r[0].shape == (1,) and o[0].shape == (1,) and r[0][0] == 0 and r[1][0] == 0      # src/regex_split.cpp:129-143


# ------------------------------------------------------------------ BPE
def _bpe_inputs(oracle_mod, a, batch):
    s = oracle_mod.SplitOracle(a.split_pattern, "isolate")(*batch)
    return s[0], s[1], s[2], s[3], batch[4]


@pytest.mark.parametrize("name", ["gpt2_synth", "llama3_synth"])
def test_bpe_oracle_equals_reference(oracle_mod, name):
    a = A.load_bpe(name)
    v, ml, mr, ad, aid = a.tensors()
    o = oracle_mod.BpeOracle(v, ml, mr, ad, aid, cache_capacity=a.cache_capacity)
    r = refops.bpe(v, ml, mr, ad, aid, cache_capacity=a.cache_capacity)
    batches = [cases.batch_from_strings(corpus()), cases.random_ascii_batch(256, 128),                       # edge corpus, C0
               cases.random_ascii_batch(2048, 512, seed=99), cases.mixed_utf8_batch(512, 1024, seed=3),       # C1 / C3 slices
               cases.english_like_batch(512, 512)]
    for bt in batches:
        ins = _bpe_inputs(oracle_mod, a, bt)
        eq(o(*ins), r(*ins), name)
    # special-token pieces arrive whole (skip-flagged upstream) and must hit the added-token entry as one symbol
    sp = cases.batch_from_strings(["<|endoftext|>", "a<|endoftext|>b", "<|endoftext|><|endoftext|>"])
    eq(o(*sp), r(*sp), name + " special")


def test_bpe_attribute_forms_oracle_equals_reference(oracle_mod):
    """Non-byte-level forms: "L R" merge strings (11 / 15 inputs), unk token, byte_fallback, end_suffix, fuse_unk."""
    vocab = [b"<unk>", b"a", b"b", b"c", b"ab", b"abc", b"bc", b"</w>", b"c</w>", b"<0x64>", b"<0x0A>", b"d", b"ab</w>", b"\xe2\x96\x81", b"\xe2\x96\x81a"]
    merges = [(b"a", b"b"), (b"ab", b"c"), (b"b", b"c"), (b"c", b"</w>"), (b"ab", b"</w>"), (b"\xe2\x96\x81", b"a")]
    v = pack_strings(vocab)
    ml, mr = pack_strings([m[0] for m in merges]), pack_strings([m[1] for m in merges])
    mtxt = pack_strings([m[0] + b" " + m[1] for m in merges])
    texts = ["abc", "abcabc", "abd", "xyz", "a\nb", "", "cab", "dab", "▁a▁ab", "abcd" * 40, "é", "ab ab"]
    bt = cases.batch_from_strings(texts)
    added = pack_strings([b"<special>", b"ab"])
    aid = np.array([15, 4], np.int32)
    for kw in ({}, {"unk_token": b"<unk>"}, {"unk_token": b"<unk>", "fuse_unk": True}, {"byte_fallback": True}, {"unk_token": b"<unk>", "byte_fallback": True},
               {"end_suffix": b"</w>"}, {"end_suffix": b"</w>", "unk_token": b"<unk>"}, {"cache_capacity": 0}, {"cache_capacity": 2}):
        for form in ("pairs", "text", "pairs+added", "text+added"):
            okw = {k: v_ for k, v_ in kw.items()}
            if "text" in form:
                o = oracle_mod.BpeOracle(v, mtxt, None, added if "added" in form else None, aid if "added" in form else None, **okw)
                r = refops.bpe(v, mtxt, None, added if "added" in form else None, aid if "added" in form else None, **kw)
            else:
                o = oracle_mod.BpeOracle(v, ml, mr, added if "added" in form else None, aid if "added" in form else None, **okw)
                r = refops.bpe(v, ml, mr, added if "added" in form else None, aid if "added" in form else None, **kw)
            for _ in range(2):      # second pass goes through the result cache
                eq(o(*bt), r(*bt), f"{kw} {form}")


# ------------------------------------------------------------------ WordPiece
def test_wordpiece_oracle_equals_reference(oracle_mod):
    a = A.load_wordpiece("bert_synth")
    v = pack_strings(a.vocab)
    s1 = oracle_mod.SplitOracle(A.BERT_WHITESPACE_PATTERN, "remove")
    s2 = oracle_mod.SplitOracle(A.BERT_PUNCT_PATTERN, "isolate")
    o = oracle_mod.WordpieceOracle(v, a.suffix_indicator, a.max_bytes_per_word)
    r = refops.wordpiece(v, a.unk_token_id, a.suffix_indicator, a.max_bytes_per_word)
    for bt in (cases.batch_from_strings([t.lower() for t in corpus()]), cases.random_ascii_batch(2048, 256, lower=True),      # edge corpus, C2 slice
               cases.english_like_batch(512, 256)):
        r1 = s1(*bt)
        r2 = s2(r1[0], r1[1], r1[2], r1[3], bt[4])
        keep = (r2[3] - r2[2]) > 0          # zero-length words are UB in the reference: none arise here, assert it
        assert keep.all()
        ins = (r2[0], r2[1], r2[2], r2[3], bt[4])
        eq(o(*ins, a.unk_token_id), r(*ins), "wordpiece")
    o2 = oracle_mod.WordpieceOracle(v, a.suffix_indicator, 5)
    rr = refops.wordpiece(v, a.unk_token_id, a.suffix_indicator, 5)
    bt = cases.batch_from_strings(["hello", "tokenization", "a", "unbelievable!"])
    eq(o2(*bt, a.unk_token_id), rr(*bt), "wordpiece max_bytes")


# ------------------------------------------------------------------ VocabEncoder / VocabDecoder / ByteFallback
def test_vocab_encoder_oracle_equals_reference(oracle_mod):
    rng = np.random.default_rng(11)
    keys = [bytes(rng.integers(97, 123, size=rng.integers(0, 9), dtype=np.uint8)) for _ in range(3000)] + [b"", "ключ".encode(), b"dup", b"dup"]
    kt = pack_strings(keys)
    probes = keys[::3] + [b"missing", b"", b"du", b"dupx"] + [bytes(rng.integers(97, 123, size=4, dtype=np.uint8)) for _ in range(500)]
    pt = pack_strings(probes)
    for dt in (np.int32, np.int64):
        vals = (np.arange(len(keys)) * 7 - 5).astype(dt)
        o = oracle_mod.VocabEncoderOracle(kt, vals)
        r = refops.vocab_encoder(kt, vals, -3)
        got = r(*pt)
        assert got.dtype == dt
        assert np.array_equal(o(*pt, -3), got.astype(np.int64))


def test_vocab_decoder_and_byte_fallback_oracle_equals_reference(oracle_mod):
    toks = A.load_detok_vocab()
    v = pack_strings(toks)
    rng = np.random.default_rng(5)
    for shape in ((64, 256), (1, 1), (7, 0), (3, 33)):
        ids = rng.integers(-3, len(toks) + 4, size=shape).astype(np.int32)
        for skip in ((0, 1, 2), (), (5, 5, 700)):
            o = oracle_mod.vocab_decoder(ids, v, skip)
            eq(o, refops.vocab_decoder(v, skip, as_input=True)(ids), f"vocab_decoder {shape} {skip}")
            eq(o, refops.vocab_decoder(v, skip, as_input=False)(ids), f"vocab_decoder attr {shape} {skip}")
            if shape[1]:
                bf = refops.simple("ByteFallback", [o[2], o[3], o[4]])(o[2], o[3], o[4])
                eq(oracle_mod.byte_fallback(o[2], o[3], o[4]), bf, "byte_fallback")
    t = pack_strings([b"<0xZZ>", b"<0x4a>", b"<0x4A>", b"<<x41>", b"<0x00>", b"<0xFF>", b"plain", b"", b"<0x4A>x"])
    eq(oracle_mod.byte_fallback(*t), refops.simple("ByteFallback", list(t))(*t), "byte_fallback odd tokens")


# ------------------------------------------------------------------ SpecialTokensSplit
def test_special_tokens_split_oracle_equals_reference(oracle_mod):
    g = json.loads((GOLDEN / "special_tokens_split_layer_tests.json").read_text())
    for case in g["cases"]:
        bt = cases.batch_from_strings([case["text"]] if "text" in case else case["texts"])
        o = oracle_mod.SpecialTokensSplitOracle(case["pattern"])(*bt)
        r = refops.special_tokens_split(case["pattern"])(*bt)
        eq(o, r, str(case)[:80])
    pat = oracle_mod.special_tokens_pattern([("<s>", False, False), ("</s>", True, True), ("<mask>", True, False), ("<|x|>", False, True)])
    bt = cases.batch_from_strings(corpus() + ["a <s> b</s>  c <mask>d<|x|>  e", "<s><s></s>", " <mask> ", "no specials"])
    o = oracle_mod.SpecialTokensSplitOracle(pat)(*bt)
    eq(o, refops.special_tokens_split(pat)(*bt), "special random")
    sk = (np.arange(len(o[2])) % 3 == 1)
    eq(oracle_mod.SpecialTokensSplitOracle(pat)(o[0], o[1], o[2], o[3], bt[4], skips=sk.astype(np.uint8)),
       refops.special_tokens_split(pat, with_skips=True)(o[0], o[1], o[2], o[3], bt[4], skips=sk), "special with skips")


# ------------------------------------------------------------------ post-tokenizer tail
def _ragged_ids(rng, n, lo, hi):
    lens = rng.integers(lo, hi, size=n)
    e = np.cumsum(lens).astype(np.int32)
    b = (e - lens).astype(np.int32)
    return b, e, rng.integers(0, 30000, size=int(e[-1]) if n else 0).astype(np.int32)


def test_truncate_oracle_equals_reference(oracle_mod):
    rng = np.random.default_rng(3)
    for n in (1, 17, 200):
        b0, e0, x0 = _ragged_ids(rng, n, 0, 40)
        b1, e1, x1 = _ragged_ids(rng, n, 0, 40)
        for max_len in (0, 1, 7, 16, 33, 100):
            for side in ("left", "right"):
                r1 = refops.simple("Truncate", [b0, e0, x0, np.int32(0), b"right", b"longest_first"], m_num_inputs=1)(b0.copy(), e0.copy(), x0, np.int32(max_len), side.encode(), b"longest_first")
                o1 = oracle_mod.truncate([(b0, e0)], max_len, side, "longest_first")
                eq([o1[0][0], o1[0][1]], r1[:2], f"truncate1 {n} {max_len} {side}")
                for mode in ("only_first", "only_second", "longest_first"):
                    r2 = refops.simple("Truncate", [b0, e0, x0, b1, e1, x1, np.int32(0), b"right", b"longest_first"], m_num_inputs=2)(
                        b0.copy(), e0.copy(), x0, b1.copy(), e1.copy(), x1, np.int32(max_len), side.encode(), mode.encode())
                    o2 = oracle_mod.truncate([(b0, e0), (b1, e1)], max_len, side, mode)
                    eq([o2[0][0], o2[0][1], o2[1][0], o2[1][1]], [r2[0], r2[1], r2[3], r2[4]], f"truncate2 {n} {max_len} {side} {mode}")


def test_combine_segments_and_ragged_to_dense_oracle_equals_reference(oracle_mod):
    rng = np.random.default_rng(4)
    n = 37
    seg_a = _ragged_ids(rng, n, 0, 20)
    seg_b = _ragged_ids(rng, n, 0, 20)
    cls = (np.zeros(1, np.int32), np.ones(1, np.int32), np.array([101], np.int32))
    sep = (np.zeros(1, np.int32), np.ones(1, np.int32), np.array([102], np.int32))
    segs = [cls, seg_a, sep, seg_b, sep]
    ids = np.array([0, 0, 0, 1, 1], np.int32)
    # single-token segments arrive with SCALAR begins / ends in converted IRs (the reference takes the output shape from the
    # last input of rank > 0, src/combine_segments.cpp:56-59)
    flat = [(x.reshape(()) if (k < 2 and len(s_[0]) == 1) else x) for s_ in segs for k, x in enumerate(s_)] + [ids]
    r = refops.simple("CombineSegments", flat)(*flat)
    o = oracle_mod.combine_segments(segs, ids)
    eq([o[0], o[1], o[2]], r[0:3], "combine elems")
    eq([o[0], o[1], o[3]], r[3:6], "combine ids")
    for target, pad_right, pad_max in ((64, True, False), (8, True, False), (8, False, False), (64, False, True), (70, True, True)):
        ins = [o[0], o[1], o[2], np.int32(target), np.int32(-7)]
        rr = refops.simple("RaggedToDense", ins, pad_right=pad_right, m_pad_max_length=pad_max)(*ins)
        if pad_max:
            continue        # pad_max_length reads beyond short rows (reference :140-160): covered by the golden vectors only
        od, om = oracle_mod.ragged_to_dense(o[0], o[1], o[2], target, -7, pad_right, pad_max)
        assert np.array_equal(od, rr[0]) and np.array_equal(om.astype(bool), rr[1].astype(bool)), (target, pad_right, pad_max)


# ------------------------------------------------------------------ byte-level shims
def test_shim_ops_oracle_equals_reference(oracle_mod):
    bt = cases.batch_from_strings(corpus())
    s = oracle_mod.SplitOracle(A.GPT2_PATTERN, "isolate")(*bt)
    ins = (s[0], s[1], s[2], s[3], bt[4])
    o = oracle_mod.bytes_to_chars(*ins)
    r = refops.simple("BytesToChars", list(ins))(*ins)
    eq(o, r[2:5], "bytes_to_chars")
    sk = (np.arange(len(s[2])) % 4 == 1)
    r = refops.simple("BytesToChars", list(ins) + [sk])(*ins, sk)
    eq(oracle_mod.bytes_to_chars(*ins, skips=sk.astype(np.uint8)), r[2:5], "bytes_to_chars skips")
    back = (s[0], s[1], o[0], o[1], o[2])
    eq(oracle_mod.chars_to_bytes(*back), refops.simple("CharsToBytes", list(back))(*back), "chars_to_bytes")
    fz = (s[0], s[1], s[2], s[3])
    eq(oracle_mod.fuze_ragged(*fz), refops.simple("FuzeRagged", list(fz))(*fz), "fuze_ragged")
    g = json.loads((GOLDEN / "shim_ops_layer_tests.json").read_text())
    raw = sorted({bytes.fromhex(c["input_hex"]) for c in g["utf8_validate"]})
    rng = np.random.default_rng(8)
    raw += [bytes(rng.integers(0, 256, size=rng.integers(0, 40), dtype=np.uint8)) for _ in range(300)] + [t.encode() for t in corpus()]
    t = pack_strings(raw)
    for mode in (False, True):
        o = oracle_mod.utf8_validate(*t, mode)
        r = refops.simple("UTF8Validate", list(t), replace_mode=mode)(*t)
        eq([o[0], o[1]], r[:2], f"utf8_validate {mode} offsets")
        assert np.array_equal(o[2][: int(o[1][-1])], r[2][: int(r[1][-1])]), f"utf8_validate {mode} chars"


def test_regex_normalization_oracle_equals_reference(oracle_mod):
    g = json.loads((GOLDEN / "normalization_layer_tests.json").read_text())
    t = pack_strings([x.encode() for x in corpus()] + [b"  lead", b"trail  ", b"\x00ctl\x07", "Ünï".encode()])
    seen = set()
    for case in g["regex_normalization"]:
        key = (case["search"], case["replace"], case.get("global_replace", True))
        if key in seen:
            continue
        seen.add(key)
        sp, rp, gl = key
        ins = [t[0], t[1], t[2], sp.encode(), rp.encode()]
        r = ref_regex_norm(ins, gl)
        o = oracle_mod.regex_normalize(sp, rp, gl, *t)
        eq(o, r[:3], f"regex_normalization {key}")


def ref_regex_norm(ins, gl):
    from oracle import ref
    op = ref.RefOp("RegexNormalization", ins, constants={3: ins[3], 4: ins[4]}, global_replace=gl)
    return op(*ins)
