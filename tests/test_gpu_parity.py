"""GPU parity tests proper: the CUDA path, called through the C ABI (ops.py -> libb200tok.so), against the
CPU oracle on the same inputs.  Bit-exact (integer / byte / index work)."""
import json
from pathlib import Path

import numpy as np
import pytest

import cases
from openvino_tokenizers_b200 import assets as A
from openvino_tokenizers_b200.strings import pack_strings, ragged_rows, unpack_strings

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def ops():
    from openvino_tokenizers_b200 import ops as O
    return O


@pytest.fixture(scope="module")
def gpt2(ops, oracle_mod):
    a = A.load_bpe("gpt2_synth")
    v, ml, mr, ad, aid = a.tensors()
    consts = [*v, *ml, *mr] + ([*ad, aid] if ad is not None else [])
    return dict(assets=a, consts=consts,
                bpe=ops.BPETokenizer().with_constants(consts),
                split=ops.RegexSplit("isolate").with_pattern(a.split_pattern),
                o_bpe=oracle_mod.BpeOracle(v, ml, mr, ad, aid, cache_capacity=a.cache_capacity),
                o_split=oracle_mod.SplitOracle(a.split_pattern, "isolate"))


@pytest.fixture(scope="module")
def llama3(ops, oracle_mod):
    a = A.load_bpe("llama3_synth")
    v, ml, mr, ad, aid = a.tensors()
    consts = [*v, *ml, *mr] + ([*ad, aid] if ad is not None else [])
    return dict(assets=a, consts=consts,
                bpe=ops.BPETokenizer().with_constants(consts),
                split=ops.RegexSplit("isolate").with_pattern(a.split_pattern),
                o_bpe=oracle_mod.BpeOracle(v, ml, mr, ad, aid, cache_capacity=a.cache_capacity),
                o_split=oracle_mod.SplitOracle(a.split_pattern, "isolate"))


@pytest.fixture(scope="module")
def bert(ops, oracle_mod):
    a = A.load_wordpiece("bert_synth")
    v = pack_strings(a.vocab)
    return dict(assets=a, vocab=v,
                wp=ops.WordpieceTokenizer(a.suffix_indicator, a.max_bytes_per_word).with_constants(v),
                s1=ops.RegexSplit("remove").with_pattern(A.BERT_WHITESPACE_PATTERN),
                s2=ops.RegexSplit("isolate").with_pattern(A.BERT_PUNCT_PATTERN),
                o_wp=oracle_mod.WordpieceOracle(v, a.suffix_indicator, a.max_bytes_per_word),
                o_s1=oracle_mod.SplitOracle(A.BERT_WHITESPACE_PATTERN, "remove"),
                o_s2=oracle_mod.SplitOracle(A.BERT_PUNCT_PATTERN, "isolate"))


def oracle_chain_bpe(m, batch, threads=1):
    s = m["o_split"](*batch, threads=threads)
    return m["o_bpe"](s[0], s[1], s[2], s[3], batch[4], threads=threads)


def host_threads():
    import os
    return max(1, min(32, os.cpu_count() or 1))


def check_bpe_model(ops, m, batch):
    rb, re_, b, e, c = batch
    pat = np.frombuffer(m["assets"].split_pattern.encode(), np.uint8)
    exp_split = m["o_split"](rb, re_, b, e, c)
    exp = m["o_bpe"](exp_split[0], exp_split[1], exp_split[2], exp_split[3], c)
    # 1. stand-alone RegexSplit op (6-input form)
    got_split = m["split"].evaluate([rb, re_, b, e, c, pat])
    for k in range(4):
        assert np.array_equal(got_split[k], exp_split[k]), f"RegexSplit output {k} differs"
    # 2. stand-alone BPETokenizer op on the split result
    got = m["bpe"].evaluate([*got_split[:5], *m["consts"]])
    assert cases.ragged_rows_equal(got, exp), "BPETokenizer differs"
    # 3. fused split+BPE
    got_f = ops.split_bpe(m["split"], m["bpe"], [rb, re_, b, e, c])
    assert cases.ragged_rows_equal(got_f, exp), "fused split+BPE differs"
    return exp


# ------------------------------------------------------------------------------------------------
def test_library_loads_and_sees_gpu():
    from openvino_tokenizers_b200 import _capi as K
    assert K.lib().b200tok_device_count() >= 1


@pytest.mark.parametrize("model", ["gpt2", "llama3"])
def test_bpe_edge_corpus(ops, model, request):
    m = request.getfixturevalue(model)
    strings = cases.EDGE_STRINGS + cases.long_prompts()
    check_bpe_model(ops, m, cases.batch_from_strings(strings))
    # every string alone too (row boundaries / single-row batches)
    for s in strings[:12] + [" " * 256, ""]:
        check_bpe_model(ops, m, cases.batch_from_strings([s]))


def test_bpe_c0_random_ascii(ops, gpt2):
    """BASELINE config 0: 256 x 128 B printable ASCII."""
    exp = check_bpe_model(ops, gpt2, cases.random_ascii_batch(256, 128))
    assert len(exp[2]) > 0


def test_bpe_c1_slice_random_ascii(ops, gpt2):
    """BASELINE config 1 shape (512-byte docs), 4096 rows."""
    check_bpe_model(ops, gpt2, cases.random_ascii_batch(4096, 512))


def test_bpe_english_like(ops, gpt2, llama3):
    batch = cases.english_like_batch(2048, 512)
    check_bpe_model(ops, gpt2, batch)
    check_bpe_model(ops, llama3, batch)


def test_bpe_c3_slice_mixed_utf8(ops, llama3, gpt2):
    """BASELINE config 3 shape (1 KiB mixed UTF-8 docs), 1024 rows."""
    batch = cases.mixed_utf8_batch(1024, 1024)
    check_bpe_model(ops, llama3, batch)
    check_bpe_model(ops, gpt2, batch)


def test_bpe_ragged_lengths_and_multi_element_rows(ops, gpt2):
    rng = np.random.default_rng(5)
    strings = []
    for i in range(600):
        L = int(rng.integers(0, 1500))
        strings.append(bytes(rng.integers(0x20, 0x7F, size=L, dtype=np.uint8)))
    b, e, c = pack_strings(strings)
    # rows of 0..4 elements
    cuts = np.sort(rng.integers(0, len(strings) + 1, size=299))
    rb = np.concatenate([[0], cuts]).astype(np.int32)
    re_ = np.concatenate([cuts, [len(strings)]]).astype(np.int32)
    check_bpe_model(ops, gpt2, (rb, re_, b, e, c))


def test_bpe_skips_pass_special_tokens_through(ops, gpt2, oracle_mod):
    """7-input RegexSplit form: skip-flagged elements are not split and reach BPE whole."""
    parts = [b"Hello world", b"<|endoftext|>", b" more text here", b"<|endoftext|>", b""]
    b, e, c = pack_strings(parts)
    rb, re_ = np.array([0, 3], np.int32), np.array([3, 5], np.int32)
    skips = np.array([0, 1, 0, 1, 0], np.uint8)
    pat = np.frombuffer(gpt2["assets"].split_pattern.encode(), np.uint8)
    exp_s = gpt2["o_split"](rb, re_, b, e, c, skips=skips)
    got_s = gpt2["split"].evaluate([rb, re_, b, e, c, skips.astype(bool), pat])
    for k in range(4):
        assert np.array_equal(got_s[k], exp_s[k])
    assert np.array_equal(got_s[5].astype(np.uint8), exp_s[4])
    exp = gpt2["o_bpe"](exp_s[0], exp_s[1], exp_s[2], exp_s[3], c)
    got = ops.split_bpe(gpt2["split"], gpt2["bpe"], [rb, re_, b, e, c, skips])
    assert cases.ragged_rows_equal(got, exp)
    eot = gpt2["assets"].vocab.index(b"<|endoftext|>")
    assert eot in got[2]


def test_hf_golden_ids(ops, gpt2, llama3):
    """Second oracle: ids HuggingFace `tokenizers` produced for the same vocab (tests/golden, made by tools/make_golden.py)."""
    for name, m in (("gpt2_synth", gpt2), ("llama3_synth", llama3)):
        g = json.loads((GOLDEN / f"hf_{name}.json").read_text())
        got = ops.split_bpe(m["split"], m["bpe"], list(cases.batch_from_strings(g["texts"])))
        for i, ids in enumerate(g["ids"]):
            assert got[2][got[0][i]:got[1][i]].tolist() == ids, (name, g["texts"][i][:40])


def test_regex_split_golden_vectors(ops):
    """The reference's own known-answer vectors (tests/layer_tests.py:331-389)."""
    g = json.loads((GOLDEN / "regex_split_layer_tests.json").read_text())
    n_run = 0
    for case in g["cases"]:
        op = ops.RegexSplit(case["behaviour"], case["invert"], case["max_splits"])
        rb, re_, b, e, c = cases.batch_from_strings([case["text"]])
        pat = np.frombuffer(case["pattern"].encode(), np.uint8)
        out = op.evaluate([rb, re_, b, e, c, pat])          # every pattern of the reference's vectors runs on the GPU (CLIP through the regex machine)
        pieces = [p.decode() for p in unpack_strings(out[2], out[3], c)]
        if case["text"] == "":
            assert out[0].tolist() == [0] and out[1].tolist() == [0]   # shape-[1] shortcut
        else:
            assert pieces == case["expected"], case
        n_run += 1
    assert n_run == 33


VM_PATTERNS = {
    "clip": r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+",
    "o200k": (r"[^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]*[\p{Ll}\p{Lm}\p{Lo}\p{M}]+(?i:'s|'t|'re|'ve|'m|'ll|'d)?|"
              r"[^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]+[\p{Ll}\p{Lm}\p{Lo}\p{M}]*(?i:'s|'t|'re|'ve|'m|'ll|'d)?|\p{N}{1,3}|"
              r" ?[^\s\p{L}\p{N}]+[\r\n/]*|\s*[\r\n]+|\s+(?!\S)|\s+"),
    "qwen2": r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+",
    "cl100k_possessive": r"'(?i:[sdmt]|ll|ve|re)|[^\r\n\p{L}\p{N}]?+\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]++[\r\n]*|\s*[\r\n]|\s+(?!\S)|\s+",
    "deepseek_digits": r"\p{N}{1,3}",
    "deepseek_cjk": "[一-龥\u3040-ゟ゠-ヿ]+",
    "deepseek_main": (r"[!\"#$%&'()*+,\-./:;<=>?@\[\\\]^_`{|}~][A-Za-z]+|[^\r\n\p{L}\p{P}\p{S}]?[\p{L}\p{M}]+| ?[\p{P}\p{S}]+[\r\n]*|"
                      r"\s*[\r\n]+|\s+(?!\S)|\s+"),
}

SPLIT_CASES = [
    (A.GPT2_PATTERN, "isolate", False), (A.GPT2_DIGITS_PATTERN, "isolate", False), (A.LLAMA3_PATTERN, "isolate", False),
    (A.LLAMA3_PATTERN, "contiguous", False), (r"\s+", "remove", False), (A.BERT_PUNCT_PATTERN, "isolate", False),
    (r"\w+|[^\w\s]+", "remove", True), (".", "isolate", False), ("▁", "mergedwithnext", False),
    ("▁", "mergedwithprevious", False), (r"\p{N}", "isolate", False), (r"\p{P}", "contiguous", False),
    (r"\s+", "mergedwithprevious", True), (r"\s+", "mergedwithnext", True), (r"\p{Nd}|\p{Nl}|\p{No}", "remove", False),
    (r"\s+", "isolate", True),
    # general patterns, compiled for the regex machine (csrc/regex_vm.cuh): CLIP, gpt-4o (o200k), Qwen2, cl100k with possessive
    # quantifiers, the three DeepSeek-V3 splitters, and a few that exercise groups / anchors / counted repeats
    (VM_PATTERNS["clip"], "isolate", True), (VM_PATTERNS["o200k"], "isolate", False), (VM_PATTERNS["qwen2"], "isolate", False),
    (VM_PATTERNS["cl100k_possessive"], "isolate", False), (VM_PATTERNS["deepseek_digits"], "isolate", False),
    (VM_PATTERNS["deepseek_cjk"], "isolate", False), (VM_PATTERNS["deepseek_main"], "isolate", False),
    (VM_PATTERNS["o200k"], "contiguous", False), (VM_PATTERNS["qwen2"], "remove", True), (r"(ab|a)(c|bcd)?", "mergedwithnext", False),
    (r"a.c|^x|y$", "mergedwithprevious", False), (r"\s?\w{2,4}", "isolate", False), (r"(?i)straße|[a-f]+", "remove", False),
]


@pytest.mark.parametrize("pattern,behaviour,invert", SPLIT_CASES)
def test_regex_split_random(ops, oracle_mod, pattern, behaviour, invert):
    alpha = ["a", "s", "t", "'", "l", "r", "e", "v", "1", "2", " ", " ", "\n", "\r", "\t", "!", "?", "é", "ſ", " ",
             "測", "😁", "▁", "_", "S", "L", "٣", "word", "  ", "...", "12345"]
    rng = np.random.default_rng(11)
    strings = ["".join(rng.choice(alpha, size=int(rng.integers(0, 40)))) for _ in range(1500)]
    strings += ["".join(rng.choice(alpha, size=int(rng.integers(300, 900)))) for _ in range(40)]
    strings += cases.EDGE_STRINGS
    if all(s == "" for s in strings):
        strings.append("x")
    batch = cases.batch_from_strings(strings)
    o = oracle_mod.SplitOracle(pattern, behaviour, invert)
    exp = o(*batch)
    op = ops.RegexSplit(behaviour, invert)
    got = op.evaluate([*batch, np.frombuffer(pattern.encode(), np.uint8)])
    for k in range(4):
        assert np.array_equal(got[k], exp[k]), f"output {k} differs"


def test_regex_split_max_splits(ops, oracle_mod):
    batch = cases.batch_from_strings(["a b c d e f", "no-space", " lead", "x  y", ""])
    for ms in (1, 2, 4):
        o = oracle_mod.SplitOracle(r"\s+", "remove", False, ms)
        exp = o(*batch)
        got = ops.RegexSplit("remove", False, ms).evaluate([*batch, np.frombuffer(rb"\s+", np.uint8)])
        for k in range(4):
            assert np.array_equal(got[k], exp[k])


def test_regex_split_legacy_skip_tokens(ops, oracle_mod):
    """The legacy 9-input form (reference src/regex_split.cpp:98-113, 164-178, 231-238): elements equal to a skip token pass through
    unsplit.  Checked against the reference's own compiled RegexSplit (one element per row: its stand-in skips tensor has one entry
    per ROW, :196-197) and, for rows of several elements, against the 7-input oracle with the equal elements flagged."""
    import refops
    from oracle import ref
    from openvino_tokenizers_b200.strings import pack_strings
    tokens = ["<|endoftext|>", "<s>", "hello world", "a", "...", "  "]
    tb, te, tc = pack_strings(tokens)
    rng = np.random.default_rng(5)
    alpha = ["a", "b", " ", "  ", ".", "<s>", "<|endoftext|>", "hello world", "...", "é", "1"]
    strings = [t for t in tokens] + ["".join(rng.choice(alpha, size=int(rng.integers(0, 12)))) for _ in range(600)] + ["hello world!", "<s> ", ""]
    pat = A.GPT2_PATTERN
    pat_u8 = np.frombuffer(pat.encode(), np.uint8)
    for behaviour in ("isolate", "remove", "mergedwithprevious"):
        op = ops.RegexSplit(behaviour)
        # (i) one element per row, against the reference
        batch = cases.batch_from_strings(strings)
        got = op.evaluate([*batch, pat_u8, tb, te, tc])
        assert len(got) == 5
        if refops.available():
            protos = [np.zeros(1, np.int32)] * 4 + [np.zeros(1, np.uint8), pat_u8, np.zeros(1, np.int32), np.zeros(1, np.int32), np.zeros(1, np.uint8)]
            rop = ref.RefOp("RegexSplit", protos, constants={5: pat_u8, 6: tb, 7: te, 8: tc}, behaviour=behaviour, invert=False, max_splits=-1)
            exp = rop(*batch, pat_u8, tb, te, tc)
            for k in range(4):
                assert np.array_equal(got[k], exp[k]), (behaviour, k)
        # (ii) several elements per row, against the oracle with the equal elements flagged as skips
        rb, re_, b, e, c = batch
        n = len(b)
        rb2 = np.arange(0, n, 3, dtype=np.int32)
        re2 = np.minimum(rb2 + 3, n).astype(np.int32)
        flags = np.array([bytes(c[b[i]:e[i]]).decode() in tokens for i in range(n)], np.uint8)
        exp2 = oracle_mod.SplitOracle(pat, behaviour)(rb2, re2, b, e, c, flags)
        got2 = op.evaluate([rb2, re2, b, e, c, pat_u8, tb, te, tc])
        for k in range(4):
            assert np.array_equal(got2[k], exp2[k]), (behaviour, "rows of three", k)
    with pytest.raises(ValueError):
        ops.RegexSplit("isolate").evaluate([*batch, pat_u8, tb, te])          # 8 inputs


def test_bpe_with_doubly_produced_tokens(ops, oracle_mod, gpt2):
    """A vocabulary in which tokens are the product of more than one merge (tiktoken-derived merge lists are like that; SURVEY App. B
    item 1): MergeTable::tie_check is on — the fast kernel watches for merges that meet their own product on both sides, and every row
    it hands back (here: rows with an added token inside the text) is redone through the exact heap form of the loop
    (std::priority_queue's pop order restated).  Results must equal the oracle's, fused and as separate ops."""
    a = gpt2["assets"]
    vocab = list(a.vocab)
    index = {t: i for i, t in enumerate(vocab)}
    merges = list(a.merges)
    have = set(merges)
    extra = []
    for l, r in merges[:4000]:
        t = l + r
        for k in range(1, len(t)):
            l2, r2 = t[:k], t[k:]
            if (l2, r2) not in have and l2 in index and r2 in index:
                extra.append((l2, r2)); have.add((l2, r2))
                break
        if len(extra) >= 300:
            break
    assert extra
    allm = merges + extra
    v = pack_strings(vocab)
    ml, mr = pack_strings([m[0] for m in allm]), pack_strings([m[1] for m in allm])
    _, _, _, ad, aid = a.tensors()
    consts = [*v, *ml, *mr] + ([*ad, aid] if ad is not None else [])
    bpe = ops.BPETokenizer().with_constants(consts)
    o_bpe = oracle_mod.BpeOracle(v, ml, mr, ad, aid, use_cache=False)
    rng = np.random.default_rng(77)
    strings = [s.encode() for s in cases.EDGE_STRINGS] + [p.encode() for p in cases.long_prompts()]
    strings += [bytes(rng.integers(0x20, 0x7F, size=int(n), dtype=np.uint8)) for n in rng.integers(0, 700, size=1500)]
    strings += [b"aaaa" * 100, b"abab" * 150, b"the the the the " * 40, b"x<|endoftext|>y" * 10]
    batch = cases.batch_from_strings(strings)
    sp = gpt2["o_split"](*batch)
    exp = o_bpe(sp[0], sp[1], sp[2], sp[3], batch[4])
    got = ops.split_bpe(gpt2["split"], bpe, list(batch))
    assert cases.ragged_rows_equal(got, exp)
    got2 = bpe.evaluate([sp[0], sp[1], sp[2], sp[3], batch[4], *consts])       # separate ops: the generic kernel (exact, or an explicit error)
    assert cases.ragged_rows_equal(got2, exp)


def test_regex_split_unknown_pattern_is_an_error(ops):
    for pat in (r"(foo|bar)+baz", r"\bword\b", r"a*?b", r"(?<=x)y", r"\p{Han}+", r"(?=ab)a", r"[[:alpha:]]+", r"(a)\1"):
        with pytest.raises(ops.B200TokError) as ei:         # outside the compiled syntax: refused, never approximated
            ops.RegexSplit("isolate").with_pattern(pat)
        assert ei.value.code == -4, pat
    with pytest.raises(ops.B200TokError):
        ops.RegexSplit("nonsense").with_pattern(r"\s+")
    with pytest.raises(ops.B200TokError):
        ops.RegexSplit("remove", max_splits=0).with_pattern(r"\s+")


def test_bpe_end_suffix_unk_and_byte_fallback(ops, oracle_mod):
    """Non-byte-level BPE (11-input "L R" merges form) with end_suffix, unk token and byte_fallback tokens."""
    vocab = ["<unk>", "a", "b", "c", "</w>", "ab", "abc", "c</w>", "bc</w>", "<0x64>", "<0x65>", "d</w>", "ab</w>"]
    merges = ["a b", "ab c", "c </w>", "b c</w>", "ab </w>"]
    v, mg = pack_strings(vocab), pack_strings(merges)
    words = [b"abc", b"ab", b"abcd", b"xabc", b"de", b"", b"cab", b"abab" * 50]
    b, e, c = pack_strings(words)
    rb = np.arange(len(words), dtype=np.int32)
    for bf in (False, True):
        for unk in ("<unk>", ""):
            o = oracle_mod.BpeOracle(v, mg, None, unk_token=unk.encode(), end_suffix=b"</w>", byte_fallback=bf)
            exp = o(rb, rb + 1, b, e, c)
            op = ops.BPETokenizer(unk_token=unk, end_suffix="</w>", byte_fallback=bf)
            got = op.evaluate([rb, rb + 1, b, e, c, *v, *mg])
            assert cases.ragged_rows_equal(got, exp), (bf, unk)


def test_bpe_missing_merge_token_is_an_error(ops):
    v, mg = pack_strings(["a", "b"]), pack_strings(["a b"])
    with pytest.raises(ops.B200TokError) as ei:
        ops.BPETokenizer().with_constants([*v, *mg])
    assert ei.value.code == -5


def test_bpe_giant_pieces(ops, gpt2):
    strings = ["a" * 5000, " " * 3000, "ab" * 4000 + " tail", "x" * 513, "y" * 512, "z" * 511, "q" * 1024 + " " + "w" * 1025]
    check_bpe_model(ops, gpt2, cases.batch_from_strings(strings))


def test_bpe_empty_inputs(ops, gpt2):
    batch = cases.batch_from_strings(["", "", ""])
    got = gpt2["bpe"].evaluate([*batch, *gpt2["consts"]])
    assert got[0].tolist() == [0, 0, 0] and got[1].tolist() == [0, 0, 0] and len(got[2]) == 0
    got = ops.split_bpe(gpt2["split"], gpt2["bpe"], list(batch))
    assert len(got[2]) == 0
    z = np.zeros(0, np.int32)
    got = gpt2["bpe"].evaluate([z, z, z, z, np.zeros(0, np.uint8), *gpt2["consts"]])
    assert len(got[0]) == 0 and len(got[2]) == 0


# ------------------------------------------------------------------------------------------------
def oracle_chain_wp(m, batch, threads=1):
    s1 = m["o_s1"](*batch, threads=threads)
    s2 = m["o_s2"](s1[0], s1[1], s1[2], s1[3], batch[4], threads=threads)
    return s2, m["o_wp"](s2[0], s2[1], s2[2], s2[3], batch[4], m["assets"].unk_token_id, threads=threads)


def check_wp(ops, m, batch):
    unk = m["assets"].unk_token_id
    words, exp = oracle_chain_wp(m, batch)
    # stand-alone ops chained like the BERT IR: RegexSplit -> RegexSplit -> WordpieceTokenizer
    p1 = np.frombuffer(A.BERT_WHITESPACE_PATTERN.encode(), np.uint8)
    p2 = np.frombuffer(A.BERT_PUNCT_PATTERN.encode(), np.uint8)
    g1 = m["s1"].evaluate([*batch, p1])
    g2 = m["s2"].evaluate([*g1[:5], p2])
    for k in range(4):
        assert np.array_equal(g2[k], words[k]), f"BERT split output {k} differs"
    got = m["wp"].evaluate([*g2[:5], *m["vocab"], np.array(unk, np.int32)])
    assert cases.ragged_rows_equal(got, exp), "WordpieceTokenizer differs"
    got_f = ops.split_wordpiece(m["s1"], m["s2"], m["wp"], list(batch), unk)
    assert cases.ragged_rows_equal(got_f, exp), "fused split+WordPiece differs"
    return exp


def test_wordpiece_edge_corpus(ops, bert):
    strings = [s.lower() for s in cases.EDGE_STRINGS + cases.long_prompts()]
    if all(len(s) == 0 for s in strings):
        strings.append("x")
    check_wp(ops, bert, cases.batch_from_strings(strings))


def test_wordpiece_c2_slice(ops, bert):
    """BASELINE config 2 shape: 256-byte lower-cased printable ASCII docs, 4096 rows."""
    exp = check_wp(ops, bert, cases.random_ascii_batch(4096, 256, lower=True))
    assert len(exp[2]) > 0


def test_wordpiece_english_like(ops, bert):
    rb, re_, b, e, c = cases.english_like_batch(2048, 256)
    up = (c >= 0x41) & (c <= 0x5A)
    c = np.where(up, c + 32, c).astype(np.uint8)
    check_wp(ops, bert, (rb, re_, b, e, c))


def test_wordpiece_long_and_unknown_words(ops, bert, oracle_mod):
    words = [b"a" * 100, b"a" * 101, b"a" * 700, b"unaffable", b"\xe6\xb8\xac", b"zzzzqqqqxxxx", b"the"]
    b, e, c = pack_strings(words)
    rb = np.arange(len(words), dtype=np.int32)
    unk = bert["assets"].unk_token_id
    exp = bert["o_wp"](rb, rb + 1, b, e, c, unk)
    got = bert["wp"].evaluate([rb, rb + 1, b, e, c, *bert["vocab"], np.array(unk, np.int32)])
    assert cases.ragged_rows_equal(got, exp)


# ------------------------------------------------------------------------------------------------
def test_vocab_encoder(ops, oracle_mod):
    rng = np.random.default_rng(3)
    keys = [f"tok{i}".encode() for i in range(5000)] + [b"", b"dup", b"dup", "測試".encode()]
    for dt in (np.int32, np.int64):
        values = rng.integers(-5, 1 << 20, size=len(keys)).astype(dt)
        k = pack_strings(keys)
        queries = [keys[int(i)] for i in rng.integers(0, len(keys), size=3000)] + [b"missing", b"tok", b"", b"dup"]
        q = pack_strings(queries)
        exp = oracle_mod.VocabEncoderOracle(k, values)(q[0], q[1], q[2], -7)
        got = ops.VocabEncoder().evaluate([*q, *k, values, np.array(-7, dt)])[0]
        assert got.dtype == dt
        assert np.array_equal(got.astype(np.int64), exp)


def test_vocab_decoder_and_byte_fallback(ops, oracle_mod):
    vocab_list = A.load_detok_vocab()
    v = pack_strings(vocab_list)
    rng = np.random.default_rng(9)
    ids = rng.integers(-3, len(vocab_list) + 3, size=(64, 257)).astype(np.int32)
    for skip in ([], [0, 1, 2]):
        exp = oracle_mod.vocab_decoder(ids, v, skip)
        got = ops.VocabDecoder(skip_tokens=skip).evaluate([ids, *v])
        for k in range(5):
            assert np.array_equal(got[k], exp[k]), f"VocabDecoder output {k} differs"
        got5 = ops.VocabDecoder().evaluate([ids, *v, np.asarray(skip, np.int32)])
        for k in range(5):
            assert np.array_equal(got5[k], exp[k])
        # ByteFallback stand-alone on the decoder's strings, and fused into the decoder
        exp_bf = oracle_mod.byte_fallback(exp[2], exp[3], exp[4])
        got_bf = ops.ByteFallback().evaluate([got[2], got[3], got[4]])
        for k in range(3):
            assert np.array_equal(got_bf[k], exp_bf[k]), f"ByteFallback output {k} differs"
        fused = ops.VocabDecoder(skip_tokens=skip, byte_fallback=True).evaluate([ids, *v])
        assert np.array_equal(fused[2], exp_bf[0]) and np.array_equal(fused[3], exp_bf[1]) and np.array_equal(fused[4], exp_bf[2])
    # seq == 0 special case (src/vocab_decoder.cpp:46-47,61-65)
    z = np.zeros((3, 0), np.int32)
    exp = oracle_mod.vocab_decoder(z, v, [])
    got = ops.VocabDecoder().evaluate([z, *v])
    for k in range(5):
        assert np.array_equal(got[k], exp[k])


def test_byte_fallback_odd_tokens(ops, oracle_mod):
    toks = [b"<0x41>", b"<0xZZ>", b"<0x4g>", b"<<0x4>", b"<0x41", b"abcdef", b"<abcd>", b"", b"<0xff>", b"<0xFF>", b"x"]
    t = pack_strings(toks)
    exp = oracle_mod.byte_fallback(*t)
    got = ops.ByteFallback().evaluate(list(t))
    for k in range(3):
        assert np.array_equal(got[k], exp[k])


def test_c4_detokenize_full_size(ops, oracle_mod):
    """BASELINE config 4 at full size: 1024 x 1024 ids, VocabDecoder + ByteFallback fused, vs the oracle."""
    vocab_list = A.load_detok_vocab()
    v = pack_strings(vocab_list)
    ids = np.random.default_rng(1234).integers(0, len(vocab_list), size=(1024, 1024)).astype(np.int32)
    exp = oracle_mod.vocab_decoder(ids, v, [0, 1, 2])
    exp_bf = oracle_mod.byte_fallback(exp[2], exp[3], exp[4])
    got = ops.VocabDecoder(skip_tokens=[0, 1, 2], byte_fallback=True).evaluate([ids, *v])
    assert np.array_equal(got[2], exp_bf[0]) and np.array_equal(got[3], exp_bf[1]) and np.array_equal(got[4], exp_bf[2])


def test_c1_full_size_properties(ops, gpt2):
    """BASELINE config 1 at full size (65 536 x 512 B): size-independent properties + the whole batch against the oracle."""
    batch = cases.random_ascii_batch(65536, 512)
    rb, re_, b, e, c = batch
    got = ops.split_bpe(gpt2["split"], gpt2["bpe"], list(batch))
    ob, oe, ids = got
    assert ob[0] == 0 and np.array_equal(ob[1:], oe[:-1]) and oe[-1] == len(ids)
    assert ids.min() >= 0 and ids.max() < len(gpt2["assets"].vocab)
    # decode(ids) reproduces the input bytes exactly (byte-level BPE is lossless)
    lens = np.array([len(t) for t in gpt2["assets"].vocab], dtype=np.int64)
    assert int(lens[ids].sum()) == len(c)
    row_bytes = np.add.reduceat(lens[ids], ob.astype(np.int64))
    assert np.all(row_bytes == 512)
    vb, ve, vc = pack_strings(gpt2["assets"].vocab)
    sample = np.r_[0:64, 30000:30064, 65472:65536]
    for r in sample:
        dec = b"".join(gpt2["assets"].vocab[t] for t in ids[ob[r]:oe[r]])
        assert dec == bytes(c[b[r]:e[r]])
    # EVERY row against the oracle (rows sharded over the host threads; the oracle is pinned to the reference's own code by
    # tests/test_reference_pin.py)
    exp = oracle_chain_bpe(gpt2, batch, threads=host_threads())
    assert cases.ragged_rows_equal(got, exp), "C1 full batch differs from the oracle"
    # idempotence: same call, same answer
    again = ops.split_bpe(gpt2["split"], gpt2["bpe"], list(batch))
    assert cases.ragged_rows_equal(again, got)


def test_host_pipeline_equals_device_path_and_oracle(ops, gpt2):
    """Large host-buffer calls are cut into row chunks that overlap H2D / kernels / D2H; the result must equal the
    single-launch device-resident path and (on sampled rows) the oracle.  Ragged row lengths, ~12 MB."""
    import torch
    from openvino_tokenizers_b200 import runtime as R
    rng = np.random.default_rng(21)
    lens = rng.integers(0, 700, size=36000)
    strings = [bytes(rng.integers(0x20, 0x7F, size=int(n), dtype=np.uint8)) for n in lens]
    batch = cases.batch_from_strings(strings)
    assert len(batch[4]) > (8 << 20)
    got = ops.split_bpe(gpt2["split"], gpt2["bpe"], list(batch))          # host path (pipelined)
    pipe = R.TokenizerPipeline("bpe", "gpt2_synth")
    db = R.to_device(batch, torch.device("cuda", 0))
    o = pipe.run_device(db)
    torch.cuda.synchronize()
    n = int(o["n"].item())
    assert n == len(got[2])
    assert np.array_equal(o["ids"][:n].cpu().numpy(), got[2])
    assert np.array_equal(o["begins"].cpu().numpy(), got[0]) and np.array_equal(o["ends"].cpu().numpy(), got[1])
    sample = np.r_[0:50, 17990:18040, 35950:36000]
    sub = cases.batch_from_strings([strings[i] for i in sample])
    exp = oracle_chain_bpe(gpt2, sub)
    for i, r in enumerate(sample):
        assert np.array_equal(got[2][got[0][r]:got[1][r]], exp[2][exp[0][i]:exp[1][i]])


def _sample_rows(batch, rows, L):
    rb, re_, b, e, c = batch
    chars = np.concatenate([c[b[r]:e[r]] for r in rows])
    return cases.uniform_batch(chars, len(rows), L)


def test_c3_full_shard_properties(ops, llama3):
    """BASELINE config 3, one GPU's shard at full size (32 768 x 1 KiB mixed UTF-8, Llama-3 pattern on the bit-mask kernel):
    size-independent properties (lossless decode of every row, contiguous row extents, idempotence) + sampled rows vs the oracle."""
    batch = cases.mixed_utf8_batch(32768, 1024)
    rb, re_, b, e, c = batch
    got = ops.split_bpe(llama3["split"], llama3["bpe"], list(batch))
    ob, oe, ids = got
    assert ob[0] == 0 and np.array_equal(ob[1:], oe[:-1]) and oe[-1] == len(ids)
    vocab = llama3["assets"].vocab
    assert ids.min() >= 0 and ids.max() < len(vocab)
    lens = np.array([len(t) for t in vocab], dtype=np.int64)
    assert int(lens[ids].sum()) == len(c)
    assert np.all(np.add.reduceat(lens[ids], ob.astype(np.int64)) == 1024)
    sample = np.r_[0:48, 16000:16048, 32720:32768]
    for r in sample:
        assert b"".join(vocab[t] for t in ids[ob[r]:oe[r]]) == bytes(c[b[r]:e[r]])
    exp = oracle_chain_bpe(llama3, batch, threads=host_threads())        # every row of the shard
    assert cases.ragged_rows_equal(got, exp), "C3 full shard differs from the oracle"
    again = ops.split_bpe(llama3["split"], llama3["bpe"], list(batch))
    assert cases.ragged_rows_equal(again, got)


def test_c2_full_size_properties(ops, bert):
    """BASELINE config 2 at full size (65 536 x 256 B lower-cased ASCII through the fused BERT splitter + WordPiece):
    contiguous row extents, ids in range, idempotence, sampled rows vs the oracle chain of the three reference ops."""
    batch = cases.random_ascii_batch(65536, 256, lower=True)
    unk = bert["assets"].unk_token_id
    got = ops.split_wordpiece(bert["s1"], bert["s2"], bert["wp"], list(batch), unk)
    ob, oe, ids = got
    assert ob[0] == 0 and np.array_equal(ob[1:], oe[:-1]) and oe[-1] == len(ids)
    assert ids.min() >= 0 and ids.max() < len(bert["assets"].vocab)
    _, exp = oracle_chain_wp(bert, batch, threads=host_threads())         # every row
    assert cases.ragged_rows_equal(got, exp), "C2 full batch differs from the oracle"
    again = ops.split_wordpiece(bert["s1"], bert["s2"], bert["wp"], list(batch), unk)
    assert cases.ragged_rows_equal(again, got)


def test_tokenizer_pipeline_with_special_tokens_and_dense_tail(ops, gpt2, oracle_mod):
    """The op chain of a converted GPT-2-style tokenizer IR, every op on the GPU through the C ABI:
    SpecialTokensSplit -> RegexSplit + BPETokenizer (skip-flagged pieces pass through whole) -> Truncate -> CombineSegments
    (bos + tokens) -> RaggedToDense; against the same chain of oracle ops."""
    a = gpt2["assets"]
    specials = [t.decode() for t in (a.added_tokens or [])][:3] or ["<|endoftext|>"]
    rng = np.random.default_rng(5)
    texts = []
    for i in range(400):
        parts = []
        for _ in range(int(rng.integers(1, 5))):
            parts.append(bytes(rng.integers(0x20, 0x7F, size=int(rng.integers(0, 120)), dtype=np.uint8)).decode())
            if rng.random() < 0.6:
                parts.append(specials[int(rng.integers(0, len(specials)))])
        texts.append("".join(parts))
    b, e, c = pack_strings(texts)
    rb, re_ = np.arange(len(texts), dtype=np.int32), np.arange(1, len(texts) + 1, dtype=np.int32)
    pattern = oracle_mod.special_tokens_pattern([(s, False, False) for s in specials])
    # oracle chain
    o_sp = oracle_mod.SpecialTokensSplitOracle(pattern)(rb, re_, b, e, c)
    o_rs = gpt2["o_split"](o_sp[0], o_sp[1], o_sp[2], o_sp[3], c, o_sp[4])
    o_ids = gpt2["o_bpe"](o_rs[0], o_rs[1], o_rs[2], o_rs[3], c)
    # GPU chain
    g_sp = ops.SpecialTokensSplit().evaluate([rb, re_, b, e, c, np.frombuffer(pattern.encode(), np.uint8)])
    g_ids = ops.split_bpe(gpt2["split"], gpt2["bpe"], [g_sp[0], g_sp[1], g_sp[2], g_sp[3], c, g_sp[5]])
    assert cases.ragged_rows_equal(g_ids, o_ids)
    max_len, target, bos, pad = 48, 50, 50256, 0
    o_t = oracle_mod.truncate([(o_ids[0], o_ids[1])], max_len, "right")[0]
    one = (np.array([0], np.int32), np.array([1], np.int32), np.array([bos], np.int32))
    o_c = oracle_mod.combine_segments([one, (o_t[0], o_t[1], o_ids[2])], [0, 0])
    o_d = oracle_mod.ragged_to_dense(o_c[0], o_c[1], o_c[2], target, pad, True)
    u8 = lambda s: np.frombuffer(s.encode(), np.uint8)
    g_t = ops.Truncate(1).evaluate([g_ids[0], g_ids[1], g_ids[2], np.int32(max_len), u8("right"), u8("longest_first")])
    g_c = ops.CombineSegments().evaluate([*one, g_t[0], g_t[1], g_t[2], np.array([0, 0], np.int32)])
    g_d = ops.RaggedToDense(pad_right=True).evaluate([g_c[0], g_c[1], g_c[2], np.int32(target), np.int32(pad)])
    assert np.array_equal(g_d[0], o_d[0]) and np.array_equal(g_d[1], o_d[1].astype(bool))
    fused, fmask = ops.post_dense(g_ids[0], g_ids[1], g_ids[2], max_len, target, pad, prefix=[bos])
    assert np.array_equal(fused, o_d[0]) and np.array_equal(fmask, o_d[1].astype(bool))


def test_sharded_peer_store_emit_world1(ops, gpt2):
    """b200tok_split_bpe_run_sharded (emit fused with the all-gatherv over peer memory) on a one-rank group: the peer-store
    compaction kernel must produce the rows of the ordinary path (the multi-rank exchange is checked by
    tools/peer_gather_check.py under torchrun)."""
    import os
    import socket
    import torch
    import torch.distributed as dist
    from openvino_tokenizers_b200 import runtime as R
    from openvino_tokenizers_b200.sharded import PeerGather
    own = not dist.is_initialized()
    if own:
        with socket.socket() as s_:
            s_.bind(("127.0.0.1", 0))
            port = s_.getsockname()[1]
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        rng = np.random.default_rng(8)
        strings = [bytes(rng.integers(0x20, 0x7F, size=int(n), dtype=np.uint8)) for n in rng.integers(0, 900, size=3000)]
        strings[17] = b"x" * 2000                                    # a piece longer than a window: giant path, row with holes
        batch = cases.batch_from_strings(strings)
        exp = ops.split_bpe(gpt2["split"], gpt2["bpe"], list(batch))
        dev = torch.device("cuda", 0)
        pipe = R.TokenizerPipeline("bpe", "gpt2_synth")
        db = R.to_device(batch, dev)
        pg = PeerGather(db.n_rows, db.n_chars + db.n_elems, dev)
        b, e, ids = pg.run(pipe, db)
        torch.cuda.synchronize()
        assert int(pg.n.item()) == len(exp[2])
        gb, ge, gi = b.cpu().numpy(), e.cpu().numpy(), ids.cpu().numpy()
        assert np.array_equal(ge - gb, exp[1] - exp[0])                 # same row lengths; rows may sit at their worst-case positions
        assert np.all(gb[1:] >= ge[:-1])                                # ... in order, without overlap
        for r in range(len(gb)):
            assert np.array_equal(gi[gb[r]:ge[r]], exp[2][exp[0][r]:exp[1][r]]), r
    finally:
        if own:
            dist.destroy_process_group()


@pytest.mark.parametrize("wire16", [True, False], ids=["wire16", "wire32"])
def test_sharded_pull_gather_world1(ops, gpt2, wire16):
    """sharded.PullGather (all-gatherv by pull: b200tok_peer_pack_run + b200tok_peer_pull_run) on a one-rank group, several steps
    (both parities of the double-buffered source buffers): compact rows, equal to the ordinary path, inside the rank's slot."""
    import os
    import socket
    import torch
    import torch.distributed as dist
    from openvino_tokenizers_b200 import runtime as R
    from openvino_tokenizers_b200.sharded import PullGather
    own = not dist.is_initialized()
    if own:
        with socket.socket() as s_:
            s_.bind(("127.0.0.1", 0))
            port = s_.getsockname()[1]
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        dev = torch.device("cuda", 0)
        pipe = R.TokenizerPipeline("bpe", "gpt2_synth")
        rng = np.random.default_rng(18)
        pgs = {}
        for step, n_rows in enumerate((3000, 3000, 3000, 41)):
            strings = [bytes(rng.integers(0x20, 0x7F, size=int(n), dtype=np.uint8)) for n in rng.integers(0, 900, size=n_rows)]
            if n_rows > 100:
                strings[17 + step] = b"x" * 2000                          # a piece longer than a window: giant path
            batch = cases.batch_from_strings(strings)
            exp = ops.split_bpe(gpt2["split"], gpt2["bpe"], list(batch))
            db = R.to_device(batch, dev)
            cap = 3000 * 900 + 4096
            pg = pgs.setdefault(n_rows, PullGather(db.n_rows, cap, dev, wire16=wire16))
            b, e, ids = pg.run(pipe, db)
            torch.cuda.synchronize()
            assert int(pg.n.item()) == len(exp[2])
            gb, ge, gi = b.cpu().numpy(), e.cpu().numpy(), ids.cpu().numpy()
            assert np.array_equal(gb, exp[0]) and np.array_equal(ge, exp[1])       # rank 0's slot starts at 0: the compact offsets themselves
            assert np.array_equal(gi[: len(exp[2])], exp[2])
    finally:
        if own:
            dist.destroy_process_group()


@pytest.mark.parametrize("vocab,pattern", [("gpt2", "llama3"), ("llama3", "gpt2"), ("gpt2", "gpt2_digits")])
def test_fast_kernel_every_instantiation(ops, oracle_mod, gpt2, llama3, vocab, pattern):
    """The dedicated kernel is instantiated per (id width, split pattern): cross the vocabularies and patterns so that
    <u16, llama3> and <i32, gpt2 / gpt2-digits> run too (the natural pairs are covered by the other tests)."""
    m = {"gpt2": gpt2, "llama3": llama3}[vocab]
    pat = {"gpt2": A.GPT2_PATTERN, "llama3": A.LLAMA3_PATTERN, "gpt2_digits": A.GPT2_DIGITS_PATTERN}[pattern]
    split = ops.RegexSplit("isolate").with_pattern(pat)
    o_split = oracle_mod.SplitOracle(pat, "isolate")
    rng = np.random.default_rng(41)
    strings = [bytes(rng.integers(0x20, 0x7F, size=int(n), dtype=np.uint8)) for n in rng.integers(0, 1500, size=300)]
    strings += [s.encode() for s in cases.EDGE_STRINGS if "<|" not in s] + [p.encode() for p in cases.long_prompts()]
    strings += [("12345678901234567890 " * 40).encode(), ("\n\n  \t" * 300).encode(), ("x" * 40 + " ") * 30 and (("x" * 40 + " ") * 30).encode()]
    batch = cases.batch_from_strings(strings)
    mix = cases.mixed_utf8_batch(256, 1024, seed=9)
    for bt in (batch, mix):
        s = o_split(*bt)
        exp = m["o_bpe"](s[0], s[1], s[2], s[3], bt[4])
        got = ops.split_bpe(split, m["bpe"], list(bt))
        assert cases.ragged_rows_equal(got, exp), (vocab, pattern)


def test_sharded_wordpiece_world1_and_concurrent_calls(ops, bert, gpt2):
    """(1) The sharded WordPiece entry point (compaction storing into the peer slots) on a one-rank group equals the ordinary
    path.  (2) A handle may be used from several host threads at once (the reference's evaluate() is const and concurrent):
    four threads, two handles, interleaved calls, every result identical to the single-threaded one."""
    import os
    import socket
    import threading
    import torch
    import torch.distributed as dist
    from openvino_tokenizers_b200 import runtime as R
    from openvino_tokenizers_b200.sharded import PeerGather
    own = not dist.is_initialized()
    if own:
        with socket.socket() as s_:
            s_.bind(("127.0.0.1", 0))
            port = s_.getsockname()[1]
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        batch = cases.random_ascii_batch(3000, 256, seed=3, lower=True)
        unk = bert["assets"].unk_token_id
        exp = ops.split_wordpiece(bert["s1"], bert["s2"], bert["wp"], list(batch), unk)
        dev = torch.device("cuda", 0)
        pipe = R.TokenizerPipeline("wordpiece", "bert_synth")
        db = R.to_device(batch, dev)
        pg = PeerGather(db.n_rows, db.n_chars + db.n_elems, dev)
        b, e, ids = pg.run(pipe, db)
        torch.cuda.synchronize()
        gb, ge, gi = b.cpu().numpy(), e.cpu().numpy(), ids.cpu().numpy()
        assert int(pg.n.item()) == len(exp[2]) and np.array_equal(ge - gb, exp[1] - exp[0])
        for r in range(0, len(gb), 7):
            assert np.array_equal(gi[gb[r]:ge[r]], exp[2][exp[0][r]:exp[1][r]])
    finally:
        if own:
            dist.destroy_process_group()
    # concurrent evaluate() calls
    b1 = cases.random_ascii_batch(2000, 300, seed=11)
    b2 = cases.random_ascii_batch(1500, 256, seed=12, lower=True)
    ref1 = ops.split_bpe(gpt2["split"], gpt2["bpe"], list(b1))
    ref2 = ops.split_wordpiece(bert["s1"], bert["s2"], bert["wp"], list(b2), unk)
    errors = []

    def worker(k):
        try:
            for _ in range(6):
                if k % 2 == 0:
                    assert cases.ragged_rows_equal(ops.split_bpe(gpt2["split"], gpt2["bpe"], list(b1)), ref1)
                else:
                    assert cases.ragged_rows_equal(ops.split_wordpiece(bert["s1"], bert["s2"], bert["wp"], list(b2), unk), ref2)
        except Exception as ex:      # noqa: BLE001
            errors.append(repr(ex))
    threads = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


# ------------------------------------------------------------------------------------------------
# In-order single-pass emit (csrc/kernels.cuh OrderedOut): the fast kernel writes the compact result itself; rows it hands
# back reserve their worst case and the gaps are closed afterwards.  Every mix must give the reference's contiguous tensors.
def _device_run(pipe, batch, capacity=None):
    import ctypes as C
    import torch
    from openvino_tokenizers_b200 import _capi as K
    from openvino_tokenizers_b200 import runtime as R
    dev = torch.device("cuda", 0)
    db = R.to_device(batch, dev)
    cap = int(capacity if capacity is not None else db.n_chars + db.n_elems)
    o = pipe.alloc_device_out(db.n_rows, cap)
    rin = K.RaggedStrings(db.rb.data_ptr(), db.re.data_ptr(), db.n_rows, db.begins.data_ptr(), db.ends.data_ptr(),
                          db.n_elems, db.chars.data_ptr(), db.n_chars, None, K.MEM_DEVICE)
    out = K.RaggedIds(o["begins"].data_ptr(), o["ends"].data_ptr(), o["ids"].data_ptr(), cap, 0, o["n"].data_ptr(), K.MEM_DEVICE)
    pipe._call(rin, out, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    n = int(o["n"].item())
    return o["begins"].cpu().numpy(), o["ends"].cpu().numpy(), o["ids"][:n].cpu().numpy()


def _mixed_rows(rng, n, special="<|endoftext|>"):
    """Rows of every kind the emit has to order: one-window rows, multi-window rows, rows longer than the staging area, empty rows,
    rows the bit-mask path hands back (an added token in the text, a piece longer than a window)."""
    rows = []
    for i in range(n):
        k = rng.integers(0, 100)
        if k < 55:
            rows.append(bytes(rng.integers(0x20, 0x7F, size=int(rng.integers(1, 513)), dtype=np.uint8)).decode())
        elif k < 70:
            rows.append(bytes(rng.integers(0x20, 0x7F, size=int(rng.integers(513, 4000)), dtype=np.uint8)).decode())
        elif k < 75:
            rows.append("")
        elif k < 83:
            rows.append("ab " + special + bytes(rng.integers(0x20, 0x7F, size=int(rng.integers(0, 300)), dtype=np.uint8)).decode())
        elif k < 90:
            rows.append("x" * int(rng.integers(600, 1500)) + " tail")                       # one piece longer than a window
        elif k < 95:
            rows.append(bytes(rng.integers(0x20, 0x7F, size=int(rng.integers(9000, 14000)), dtype=np.uint8)).decode())   # > staging area
        else:
            rows.append("Тест 測試 😁 " * int(rng.integers(1, 60)))
    return rows


@pytest.mark.parametrize("model", ["gpt2", "llama3"])
def test_ordered_emit_mixed_rows(ops, model, request):
    from openvino_tokenizers_b200 import runtime as R
    m = request.getfixturevalue(model)
    pipe = R.TokenizerPipeline("bpe", "gpt2_synth" if model == "gpt2" else "llama3_synth")
    rng = np.random.default_rng(77)
    for n, first_special in ((700, False), (300, True), (1, False), (1, True), (40, False)):
        rows = _mixed_rows(rng, n)
        if first_special:
            rows[0] = "<|endoftext|>" + rows[0]
        batch = cases.batch_from_strings(rows)
        exp = oracle_chain_bpe(m, batch, threads=host_threads())
        assert exp[0][0] == 0 and np.array_equal(exp[0][1:], exp[1][:-1])
        got_d = _device_run(pipe, batch)
        assert cases.ragged_rows_equal(got_d, exp), f"{model}: device-resident in-order emit differs (n={n})"
        got_h = ops.split_bpe(m["split"], m["bpe"], list(batch))
        assert cases.ragged_rows_equal(got_h, exp), f"{model}: host path differs (n={n})"
        again = _device_run(pipe, batch)                 # the descriptor array is reused under a new epoch
        assert cases.ragged_rows_equal(again, exp)


def test_ordered_emit_last_rows_handed_back_and_tight_capacity(ops, gpt2):
    from openvino_tokenizers_b200 import runtime as R
    pipe = R.TokenizerPipeline("bpe", "gpt2_synth")
    rows = ["plain text row %d" % i for i in range(50)] + ["<|endoftext|>", "y" * 900, "<|endoftext|> end"]
    batch = cases.batch_from_strings(rows)
    exp = oracle_chain_bpe(gpt2, batch)
    assert cases.ragged_rows_equal(_device_run(pipe, batch), exp)
    # a caller buffer that holds the result but not the worst case of a handed-back row: the slot-buffer path takes over
    tight = _device_run(pipe, batch, capacity=len(exp[2]) + 8)
    assert cases.ragged_rows_equal(tight, exp)
    # multi-element rows (several strings per row) with skip flags
    b, e, c = pack_strings(rows)
    rb = np.arange(0, len(rows), 3, dtype=np.int32)
    re_ = np.minimum(rb + 3, len(rows)).astype(np.int32)
    sk = np.zeros(len(rows), np.uint8)
    sk[50] = 1
    s = gpt2["o_split"](rb, re_, b, e, c, skips=sk)
    exp2 = gpt2["o_bpe"](s[0], s[1], s[2], s[3], c)
    got2 = ops.split_bpe(gpt2["split"], gpt2["bpe"], [rb, re_, b, e, c, sk])
    assert cases.ragged_rows_equal(got2, exp2)


def test_sharded_exchange_two_ranks():
    """Every rank's gathered slots must equal the single-GPU result of the shard that produced them, for both wire formats
    (tools/peer_gather_check.py under torchrun; needs two GPUs)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = Path(__file__).resolve().parent.parent
    for wire16 in ("0", "1"):
        env = dict(__import__("os").environ, B200TOK_WIRE16=wire16, MASTER_ADDR="127.0.0.1")
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                            "--master-port", "29577", str(root / "tools" / "peer_gather_check.py"), "8192"], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        assert "on every rank: True" in r.stdout, r.stdout[-2000:]       # (covers the peer-store emit and the pull gather, both wire widths)


@pytest.mark.parametrize("env", [{"B200TOK_TMA": "1"}, {"B200TOK_TMA": "2"}, {"B200TOK_SLOT_ALLOC": "1"}, {"B200TOK_ORDERED_EMIT": "1"}, {"B200TOK_GRAPHS": "0"}],
                         ids=["tma-prefetch", "tma-sync", "slot-alloc", "ordered-emit", "no-graphs"])
def test_fast_kernel_variants(env):
    """The opt-in row loops of the fast kernel (TMA bulk-copy staging with / without prefetch, bump-allocated slots, in-order
    single-pass emit) and the path without CUDA-graph replay give the same tensors as the default: tools/variant_check.py in a
    process of its own (the library reads the switches once)."""
    import os
    import subprocess
    import sys
    root = Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, str(root / "tools" / "variant_check.py")], capture_output=True, text=True, env=dict(os.environ, **env), timeout=900)
    assert r.returncode == 0 and "variant ok" in r.stdout, (r.stdout[-1500:], r.stderr[-1500:])


def test_fused_split_bpe_with_a_general_pattern(ops, llama3, oracle_mod):
    """RegexSplit with a pattern only the regex machine knows (Qwen2's, gpt-4o's), fused with BPETokenizer and as separate ops."""
    batch = cases.batch_from_strings(cases.EDGE_STRINGS + cases.long_prompts())
    rnd = cases.mixed_utf8_batch(512, 1024, seed=12)
    for name in ("qwen2", "o200k", "clip"):
        pat = VM_PATTERNS[name]
        split = ops.RegexSplit("isolate").with_pattern(pat)
        o_split = oracle_mod.SplitOracle(pat, "isolate")
        for bt in (batch, rnd):
            s = o_split(*bt)
            exp = llama3["o_bpe"](s[0], s[1], s[2], s[3], bt[4])
            assert cases.ragged_rows_equal(ops.split_bpe(split, llama3["bpe"], list(bt)), exp), name
            g = split.evaluate([*bt, np.frombuffer(pat.encode(), np.uint8)])
            assert cases.ragged_rows_equal(llama3["bpe"].evaluate([*g[:5], *llama3["consts"]]), exp), name
