"""The ov::Op shim (openvino_tokenizers_b200/csrc/ov_shim/ov_extension_b200.cpp) compiled against the stand-in OpenVINO API.

CPU tier: it compiles, exports the extension entry points, registers every op name of the hot path, and loading an IR-like chain
through its extensions fuses RegexSplit -> BPETokenizer / RegexSplit -> RegexSplit -> WordpieceTokenizer into one layer.
GPU tier: evaluate() of the shim's classes against evaluate() of the REFERENCE'S OWN classes (oracle/_ref) on the same tensors —
op by op and for whole chains loaded both ways."""
import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import cases
import refops
import shimlib
from openvino_tokenizers_b200 import assets as A
from openvino_tokenizers_b200.strings import pack_strings
from oracle import ref

ROOT = Path(__file__).resolve().parent.parent
OPS = ["RegexSplit", "BPETokenizer", "WordpieceTokenizer", "VocabEncoder", "VocabDecoder", "ByteFallback", "SpecialTokensSplit", "Truncate",
       "CombineSegments", "RaggedToDense", "RegexNormalization", "CharsMapNormalization", "BytesToChars", "CharsToBytes", "FuzeRagged", "UTF8Validate"]


def test_shim_compiles_and_exports_the_extension_entry_points():
    lib = shimlib.build()
    names = subprocess.run(["nm", "-D", "--defined-only", str(lib)], capture_output=True, text=True, check=True).stdout
    assert " create_extensions" in names            # OPENVINO_CREATE_EXTENSIONS (reference src/ov_extension.cpp:72)
    assert " create_tokenizer_node" in names        # GenAI's factory (reference src/tokenizers_factory.hpp:32-33)
    L = shimlib.lib()
    L.ovshim_registered.argtypes = [C.c_char_p]
    for op in OPS:
        assert L.ovshim_registered(op.encode()) == 1, f"{op} is not registered by the shim"


I32, U8, BOOL = np.zeros(1, np.int32), np.zeros(1, np.uint8), np.zeros(1, np.bool_)


def bpe_chain(G, a, with_skips=True):
    """Parameters -> RegexSplit -> BPETokenizer, as the converter emits them (7-input RegexSplit, 18-input BPETokenizer)."""
    v, ml, mr, ad, aid = a.tensors()
    params = [G.parameter(x) for x in ([I32, I32, I32, I32, U8] + ([BOOL] if with_skips else []))]
    pat = G.constant(a.split_pattern.encode())
    rs = G.layer("RegexSplit", params + [pat], behaviour="isolate", invert=False, max_splits=-1)
    consts = [G.constant(x) for x in ([*v, *ml, *mr] + ([*ad, np.asarray(aid, np.int32)] if ad is not None else []))]
    out = G.layer("BPETokenizer", rs[:5] + consts, cache_capacity=a.cache_capacity)
    return params, rs, out


def wp_chain(G, a):
    v = pack_strings(a.vocab)
    params = [G.parameter(x) for x in [I32, I32, I32, I32, U8, BOOL]]
    s1 = G.layer("RegexSplit", params + [G.constant(A.BERT_WHITESPACE_PATTERN.encode())], behaviour="remove", invert=False, max_splits=-1)
    s2 = G.layer("RegexSplit", s1[:6] + [G.constant(A.BERT_PUNCT_PATTERN.encode())], behaviour="isolate", invert=False, max_splits=-1)
    out = G.layer("WordpieceTokenizer", s2[:5] + [G.constant(x) for x in v] + [G.constant(np.asarray(a.unk_token_id, np.int32))],
                  suffix_indicator=a.suffix_indicator, max_bytes_per_word=a.max_bytes_per_word)
    return params, (s1, s2), out


def test_loading_a_chain_fuses_split_and_tokenizer():
    a = A.load_bpe("gpt2_synth")
    for with_skips in (True, False):
        G = shimlib.ShimGraph()
        _, rs, out = bpe_chain(G, a, with_skips)
        assert G.producer_type(rs[0]) == "RegexSplit"
        assert G.producer_type(out[0]) == "B200SplitBPE" and len(out) == 3        # same three outputs as BPETokenizer
    G = shimlib.ShimGraph()
    _, _, out = wp_chain(G, A.load_wordpiece("bert_synth"))
    assert G.producer_type(out[0]) == "B200SplitWordpiece" and len(out) == 3
    # unsupported producer configurations stay unfused: merged-with-next splitter, max_splits, a non-RegexSplit producer
    G = shimlib.ShimGraph()
    v, ml, mr, ad, aid = a.tensors()
    consts = [*v, *ml, *mr] + ([*ad, np.asarray(aid, np.int32)] if ad is not None else [])
    params = [G.parameter(x) for x in [I32, I32, I32, I32, U8]]
    rs = G.layer("RegexSplit", params + [G.constant(b"\\s+")], behaviour="mergedwithnext", invert=False, max_splits=-1)
    assert G.producer_type(G.layer("BPETokenizer", rs[:5] + [G.constant(x) for x in consts])[0]) == "BPETokenizer"
    rs = G.layer("RegexSplit", params + [G.constant(b"\\s+")], behaviour="isolate", invert=False, max_splits=2)
    assert G.producer_type(G.layer("BPETokenizer", rs[:5] + [G.constant(x) for x in consts])[0]) == "BPETokenizer"
    assert G.producer_type(G.layer("BPETokenizer", params + [G.constant(x) for x in consts])[0]) == "BPETokenizer"


def test_legacy_nine_input_regex_split_loads_and_stays_a_layer_of_its_own():
    """The legacy RegexSplit form (reference src/regex_split.cpp:98-113: pattern at input 5, skip-token strings at 6..8, five outputs)
    is accepted by the shim's validate_and_infer_types, 8 inputs are not, and the load-time fusion leaves such a splitter alone."""
    a = A.load_bpe("gpt2_synth")
    G = shimlib.ShimGraph()
    params = [G.parameter(x) for x in [I32, I32, I32, I32, U8]]
    toks = pack_strings(["<|endoftext|>", "<s>"])
    rs = G.layer("RegexSplit", params + [G.constant(a.split_pattern.encode())] + [G.constant(x) for x in toks], behaviour="isolate", invert=False, max_splits=-1)
    assert len(rs) == 5 and G.producer_type(rs[0]) == "RegexSplit"
    v, ml, mr, ad, aid = a.tensors()
    consts = [*v, *ml, *mr] + ([*ad, np.asarray(aid, np.int32)] if ad is not None else [])
    out = G.layer("BPETokenizer", rs[:5] + [G.constant(x) for x in consts])
    assert G.producer_type(out[0]) == "BPETokenizer"
    with pytest.raises(RuntimeError):
        G.layer("RegexSplit", params + [G.constant(a.split_pattern.encode())] + [G.constant(x) for x in toks[:2]], behaviour="isolate", invert=False, max_splits=-1)


def test_fusion_can_be_switched_off():
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import shimlib, test_ov_shim as T\nfrom openvino_tokenizers_b200 import assets as A\n"
            "G = shimlib.ShimGraph(); _, _, out = T.bpe_chain(G, A.load_bpe('gpt2_synth')); print(G.producer_type(out[0]))" % (str(ROOT), str(ROOT / "tests")))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, B200TOK_FUSE="0"), timeout=300)
    assert r.stdout.strip().endswith("BPETokenizer"), (r.stdout, r.stderr[-1000:])


# ------------------------------------------------------------------------------------------------ GPU tier
needs_ref = pytest.mark.skipif(not refops.available(), reason="oracle/_ref not built and /root/reference absent")


def eq(a, b, what):
    assert len(a) == len(b), what
    for i, (x, y) in enumerate(zip(a, b)):
        if x is None or y is None:
            continue
        assert x.shape == y.shape and np.array_equal(x, y), f"{what}: output {i} differs"


def corpus_batch():
    return cases.batch_from_strings(cases.EDGE_STRINGS + cases.long_prompts())


@pytest.mark.gpu
@needs_ref
def test_shim_chains_equal_reference_chains():
    """The same IR-like chain loaded through the reference's extensions (separate RegexSplit / tokenizer layers, the reference's
    evaluate()) and through the shim's (one fused layer on the GPU): identical output tensors."""
    a = A.load_bpe("gpt2_synth")
    for batch in (corpus_batch(), cases.random_ascii_batch(2048, 512, seed=3), cases.mixed_utf8_batch(256, 1024, seed=4)):
        sk = np.zeros(len(batch[2]), np.bool_)
        sk[::7] = True
        for with_skips in (True, False):
            GR, GS = ref.StubGraph(), shimlib.ShimGraph()
            _, _, out_r = bpe_chain(GR, a, with_skips)
            _, _, out_s = bpe_chain(GS, a, with_skips)
            ins = list(batch) + ([sk] if with_skips else [])
            eq(GR.run(out_r, ins), GS.run(out_s, ins), f"bpe chain skips={with_skips}")
            assert GS.evaluated_nodes == 1 and GR.evaluated_nodes == 2          # one fused layer instead of two
    w = A.load_wordpiece("bert_synth")
    for batch in (cases.batch_from_strings([t.lower() for t in cases.EDGE_STRINGS if t]), cases.random_ascii_batch(2048, 256, seed=5, lower=True)):
        sk = np.zeros(len(batch[2]), np.bool_)
        GR, GS = ref.StubGraph(), shimlib.ShimGraph()
        _, _, out_r = wp_chain(GR, w)
        _, _, out_s = wp_chain(GS, w)
        eq(GR.run(out_r, list(batch) + [sk]), GS.run(out_s, list(batch) + [sk]), "wordpiece chain")
        assert GS.evaluated_nodes == 1 and GR.evaluated_nodes == 3


@pytest.mark.gpu
@needs_ref
def test_shim_ops_equal_reference_ops():
    """Every op class of the shim, evaluate() against the reference class of the same name on the same tensors."""
    batch = corpus_batch()
    rb, re_, b, e, c = batch
    a = A.load_bpe("gpt2_synth")
    v, ml, mr, ad, aid = a.tensors()
    pat = a.split_pattern.encode()
    sk = (np.arange(len(b)) % 5 == 2)
    # RegexSplit, 7- and 6-input forms, several behaviours
    for beh, p in (("isolate", pat), ("remove", rb"\s+"), ("mergedwithprevious", rb"\s+"), ("mergedwithnext", rb"\s+"), ("contiguous", rb"\w")):
        for ins in ([rb, re_, b, e, c, sk, p], [rb, re_, b, e, c, p]):
            kw = dict(constants={len(ins) - 1: p}, behaviour=beh, invert=False, max_splits=-1)
            r, s = ref.RefOp("RegexSplit", ins, **kw)(*ins), shimlib.ShimOp("RegexSplit", ins, **kw)(*ins)
            eq(r[:4] + r[5:], s[:4] + s[5:], f"RegexSplit {beh} {len(ins)}")
    split = ref.RefOp("RegexSplit", [rb, re_, b, e, c, pat], constants={5: pat}, behaviour="isolate")(rb, re_, b, e, c, pat)
    # BPETokenizer, 18 inputs
    consts = [*v, *ml, *mr] + ([*ad, np.asarray(aid, np.int32)] if ad is not None else [])
    ins = list(split[:4]) + [c] + consts
    kw = dict(constants={5 + i: x for i, x in enumerate(consts)}, cache_capacity=a.cache_capacity)
    eq(ref.RefOp("BPETokenizer", ins, **kw)(*ins), shimlib.ShimOp("BPETokenizer", ins, **kw)(*ins), "BPETokenizer")
    # WordpieceTokenizer
    w = A.load_wordpiece("bert_synth")
    wv = pack_strings(w.vocab)
    lb = cases.batch_from_strings([t.lower() for t in cases.EDGE_STRINGS if t.strip()])
    s1 = ref.RefOp("RegexSplit", [*lb, rb"\s+"], constants={5: rb"\s+"}, behaviour="remove")(*lb, rb"\s+")
    unk = np.asarray(w.unk_token_id, np.int32)
    ins = list(s1[:4]) + [lb[4], *wv, unk]
    kw = dict(constants={5: wv[0], 6: wv[1], 7: wv[2], 8: unk}, suffix_indicator=w.suffix_indicator, max_bytes_per_word=w.max_bytes_per_word)
    eq(ref.RefOp("WordpieceTokenizer", ins, **kw)(*ins), shimlib.ShimOp("WordpieceTokenizer", ins, **kw)(*ins), "WordpieceTokenizer")
    # VocabEncoder (i32 and i64 values)
    keys = pack_strings([b"alpha", b"beta", b"", b"gamma", b"beta"])
    probes = pack_strings([b"beta", b"delta", b"", b"alph", b"gamma"])
    for dt in (np.int32, np.int64):
        vals, dflt = np.arange(5).astype(dt) * 3, np.asarray(-1, dt)
        ins = [*probes, *keys, vals, dflt]
        kw = dict(constants={3: keys[0], 4: keys[1], 5: keys[2], 6: vals, 7: dflt})
        eq(ref.RefOp("VocabEncoder", ins, **kw)(*ins), shimlib.ShimOp("VocabEncoder", ins, **kw)(*ins), f"VocabEncoder {dt}")
    # VocabDecoder (5-input form) + ByteFallback
    toks = pack_strings(A.load_detok_vocab())
    ids = np.random.default_rng(2).integers(-2, len(toks[0]) + 3, size=(16, 37)).astype(np.int32)
    skip = np.array([0, 1, 2], np.int32)
    ins = [ids, *toks, skip]
    kw = dict(constants={1: toks[0], 2: toks[1], 3: toks[2]})
    d_r, d_s = ref.RefOp("VocabDecoder", ins, **kw)(*ins), shimlib.ShimOp("VocabDecoder", ins, **kw)(*ins)
    eq(d_r, d_s, "VocabDecoder")
    eq(ref.RefOp("ByteFallback", d_r[2:5])(*d_r[2:5]), shimlib.ShimOp("ByteFallback", d_r[2:5])(*d_r[2:5]), "ByteFallback")
    # byte-level shims
    ins = list(split[:4]) + [c]
    b2c_r, b2c_s = ref.RefOp("BytesToChars", ins)(*ins), shimlib.ShimOp("BytesToChars", ins)(*ins)
    eq(b2c_r, b2c_s, "BytesToChars")
    ins6 = ins + [np.zeros(len(split[2]), np.bool_)]
    eq(ref.RefOp("BytesToChars", ins6)(*ins6), shimlib.ShimOp("BytesToChars", ins6)(*ins6), "BytesToChars with skips")
    back = [split[0], split[1], b2c_r[2], b2c_r[3], b2c_r[4]]
    eq(ref.RefOp("CharsToBytes", back)(*back), shimlib.ShimOp("CharsToBytes", back)(*back), "CharsToBytes")
    fz = list(split[:4])
    eq(ref.RefOp("FuzeRagged", fz)(*fz), shimlib.ShimOp("FuzeRagged", fz)(*fz), "FuzeRagged")
    raw = pack_strings([bytes(np.random.default_rng(i).integers(0, 256, size=i % 23, dtype=np.uint8)) for i in range(200)])
    for mode in (False, True):
        r, s = ref.RefOp("UTF8Validate", list(raw), replace_mode=mode)(*raw), shimlib.ShimOp("UTF8Validate", list(raw), replace_mode=mode)(*raw)
        eq(r[:2], s[:2], f"UTF8Validate {mode}")
        n = int(r[1][-1])
        assert np.array_equal(r[2][:n], s[2][:n])
    # SpecialTokensSplit, Truncate, CombineSegments, RaggedToDense
    import oracle
    sp_pat = oracle.special_tokens_pattern([("<|endoftext|>", False, False), ("<s>", True, True)]).encode()
    ins = [rb, re_, b, e, c, sp_pat]
    eq(ref.RefOp("SpecialTokensSplit", ins, constants={5: sp_pat})(*ins), shimlib.ShimOp("SpecialTokensSplit", ins, constants={5: sp_pat})(*ins), "SpecialTokensSplit")
    tok = ref.RefOp("BPETokenizer", list(split[:4]) + [c] + consts, constants={5 + i: x for i, x in enumerate(consts)})(*split[:4], c, *consts)
    tr = [tok[0].copy(), tok[1].copy(), tok[2], np.int32(9), b"right", b"longest_first"]
    tr2 = [tok[0].copy(), tok[1].copy(), tok[2], np.int32(9), b"right", b"longest_first"]
    eq(ref.RefOp("Truncate", tr, m_num_inputs=1)(*tr)[:2], shimlib.ShimOp("Truncate", tr2, m_num_inputs=1)(*tr2)[:2], "Truncate")
    t_r = ref.RefOp("Truncate", tr, m_num_inputs=1)(*tr)
    segs = [np.int32(0), np.int32(1), np.array([50256], np.int32), t_r[0], t_r[1], tok[2], np.array([0, 1], np.int32)]
    cs_r, cs_s = ref.RefOp("CombineSegments", segs)(*segs), shimlib.ShimOp("CombineSegments", segs)(*segs)
    eq(cs_r, cs_s, "CombineSegments")
    dn = [cs_r[0], cs_r[1], cs_r[2], np.int32(16), np.int32(0)]
    for pad_right in (True, False):
        r, s = ref.RefOp("RaggedToDense", dn, pad_right=pad_right, m_pad_max_length=False)(*dn), shimlib.ShimOp("RaggedToDense", dn, pad_right=pad_right, m_pad_max_length=False)(*dn)
        assert np.array_equal(r[0], s[0]) and np.array_equal(r[1].astype(bool), s[1].astype(bool)), f"RaggedToDense {pad_right}"
    # RegexNormalization (a BERT step)
    nr = [b, e, c, rb"\s", b" "]
    kw = dict(constants={3: rb"\s", 4: b" "}, global_replace=True)
    eq(ref.RefOp("RegexNormalization", nr, **kw)(*nr)[:3], shimlib.ShimOp("RegexNormalization", nr, **kw)(*nr)[:3], "RegexNormalization")
