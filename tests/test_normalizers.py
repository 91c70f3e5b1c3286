"""Normalisers: RegexNormalization (single-character patterns) and CharsMapNormalization.

CPU tier: the oracle against the reference's known-answer vectors and against the real sentencepiece; the product's host
parser + scan (tok_core.cuh norm_eval, the function the kernel runs per byte position) against the oracle.
GPU tier: the ops through the C ABI against the oracle."""
import json
from pathlib import Path

import numpy as np
import pytest

import oracle
import normcases as NC
from openvino_tokenizers_b200 import _capi as K

GOLDEN = json.loads((Path(__file__).parent / "golden" / "normalization_layer_tests.json").read_text())
STEPS = GOLDEN["bert_steps"] + GOLDEN["other_steps"] + [
    dict(name="legacy_prepend_1", search=r"(^)(.)", replace="▁\\2", global_replace=True),
    dict(name="legacy_prepend_2", search=r"(^)(.+)", replace="▁$2", global_replace=True),
    dict(name="first_space_only", search=" ", replace="_", global_replace=False),
    dict(name="first_ws_only", search=r"\s", replace="<$0>", global_replace=False),
    dict(name="han_braces", search=r"([\p{Han}])", replace="[${1}]$$", global_replace=True),
]
UNSUPPORTED = [(r"^\s*|\s*$", ""), (r" ([\\.\\?\\!,])| ('[ms])| (') | ('[rv]e)| (n't)", r"\1"), (r"\s+", " "), (r"\s", "$1"), (r"([\p{Han}])", "$1$1"),
               (r"\s", "x" * 17), (r".", "x")]


def _enc(s):
    return s.encode() if isinstance(s, str) else bytes(s)


# ------------------------------------------------------------------ CPU tier
@pytest.mark.skipif(not oracle.pcre2_available(), reason="libpcre2-8 not present")
@pytest.mark.parametrize("case", GOLDEN["regex_normalization"], ids=lambda c: repr(c["text"])[:20])
def test_oracle_regex_normalization_reference_vectors(case):
    """reference tests/layer_tests.py:253-290 (every row, including the legacy clean-up pattern the GPU path refuses)."""
    b, e, c = NC.pack([case["text"].encode()])
    ob, oe, oc = oracle.regex_normalize(case["search"], case["replace"], case["global_replace"], b, e, c)
    assert NC.unpack(ob, oe, oc) == [case["expected"].encode()]


def test_oracle_charsmap_is_sentencepiece():
    """The restated Normalizer == the installed sentencepiece 0.2.1 (the pinned version) on compiled-in and generated blobs,
    for every flag combination, malformed UTF-8 included.  tools/pin_charsmap_oracle.py repeats this with the reference's
    own blobs where /root/reference exists."""
    raw = NC.corpus(seed=11, n=500, malformed=150)
    b, e, c = NC.pack(raw)
    for blob in (NC.builtin_blob("nfkc_cf"), NC.builtin_blob("nmt_nfkc"), NC.unicodedata_blob("NFD", True, limit=0x3000)):
        for adp, rew, esc in ((0, 0, 0), (1, 0, 1), (0, 1, 0), (1, 1, 1)):
            sp = NC.sp_normalizer(blob, adp, rew, esc)
            ob, oe, oc = oracle.charsmap_normalize(blob, b, e, c, add_dummy_prefix=adp, remove_extra_whitespaces=rew, escape_whitespaces=esc)
            got = NC.unpack(ob, oe, oc)
            for r, g in zip(raw, got):
                exp = sp.normalize(r)
                exp = exp.encode("utf-8", "surrogatepass") if isinstance(exp, str) else exp
                assert g == exp, (r, adp, rew, esc)


def test_casefold_reference_vectors():
    """reference tests/layer_tests.py:226-250 (utf-8 rows) with a case-fold charsmap compiled from unicodedata."""
    blob = NC.unicodedata_blob(None, True)
    raw = [c["text"].encode() for c in GOLDEN["casefold_utf8"]]
    b, e, c = NC.pack(raw)
    ob, oe, oc = oracle.charsmap_normalize(blob, b, e, c)
    assert NC.unpack(ob, oe, oc) == [c_["expected"].encode() for c_ in GOLDEN["casefold_utf8"]]


@pytest.mark.skipif(not oracle.pcre2_available(), reason="libpcre2-8 not present")
@pytest.mark.parametrize("step", STEPS, ids=lambda s: s["name"])
def test_host_regex_scan_vs_pcre2(step):
    import hostcore
    raw = NC.corpus(seed=3, n=800)
    b, e, c = NC.pack(raw)
    skips = (np.arange(len(raw)) % 7 == 3).astype(np.uint8)
    exp = oracle.regex_normalize(step["search"], step["replace"], step["global_replace"], b, e, c, skips)
    got = hostcore.hz_normalize(0, _enc(step["search"]), _enc(step["replace"]), step["global_replace"], b, e, c, skips)
    assert NC.unpack(*got) == NC.unpack(*exp)
    assert (got[0] == exp[0]).all() and (got[1] == exp[1]).all()


@pytest.mark.parametrize("which", ["nfkc_cf", "nfd_cf", "nfd", "casefold", "empty", "custom"])
def test_host_charsmap_scan_vs_oracle(which):
    import hostcore
    blob = {"custom": NC.custom_blob, "nfkc_cf": lambda: NC.builtin_blob("nfkc_cf"), "nfd_cf": lambda: NC.unicodedata_blob("NFD", True), "nfd": lambda: NC.unicodedata_blob("NFD", False),
            "casefold": lambda: NC.unicodedata_blob(None, True), "empty": lambda: b""}[which]()
    raw = NC.corpus(seed=4, n=800, malformed=200)
    b, e, c = NC.pack(raw)
    exp = oracle.charsmap_normalize(blob, b, e, c)
    got = hostcore.hz_normalize(1, blob, b"", 0, b, e, c)
    assert NC.unpack(*got) == NC.unpack(*exp)


def test_host_scan_property_random_unicode():
    """Property test (hypothesis): for arbitrary Unicode strings the host scan of every supported RegexNormalization rule
    equals PCRE2's substitute, and the charsmap scan equals the sentencepiece restatement — and, for arbitrary BYTES
    (malformed UTF-8 included), the charsmap scan and UTF8Validate still equal their oracles."""
    import hostcore
    from hypothesis import given, settings, strategies as st

    alphabet = st.one_of(st.characters(min_codepoint=1, max_codepoint=0x7F), st.characters(min_codepoint=0x80, max_codepoint=0x2FFF, exclude_categories=("Cs",)),
                         st.sampled_from([chr(c) for c in NC.INTERESTING]))
    blob = NC.unicodedata_blob("NFD", True)

    @settings(max_examples=120, deadline=None, derandomize=True)
    @given(st.lists(st.text(alphabet, max_size=80), min_size=1, max_size=8))
    def text_case(strings):
        raw = [s_.encode() for s_ in strings]
        b, e, c = NC.pack(raw)
        if not c.size:
            return
        for step in STEPS[:9]:
            if oracle.pcre2_available():
                exp = oracle.regex_normalize(step["search"], step["replace"], step["global_replace"], b, e, c)
                got = hostcore.hz_normalize(0, _enc(step["search"]), _enc(step["replace"]), step["global_replace"], b, e, c)
                assert NC.unpack(*got) == NC.unpack(*exp), (step["name"], raw)
        assert NC.unpack(*hostcore.hz_normalize(1, blob, b"", 0, b, e, c)) == NC.unpack(*oracle.charsmap_normalize(blob, b, e, c))

    @settings(max_examples=150, deadline=None, derandomize=True)
    @given(st.lists(st.binary(max_size=60), min_size=1, max_size=8))
    def bytes_case(raw):
        b, e, c = NC.pack(raw)
        if not c.size:
            return
        assert NC.unpack(*hostcore.hz_normalize(1, blob, b"", 0, b, e, c)) == NC.unpack(*oracle.charsmap_normalize(blob, b, e, c))
        for mode in (False, True):
            assert NC.unpack(*hostcore.hz_normalize(3, b"", b"", mode, b, e, c)) == NC.unpack(*oracle.utf8_validate(b, e, c, mode))

    text_case()
    bytes_case()


def test_charsmap_ascii_shortcut_tables_agree_with_the_trie():
    import hostcore
    for blob in (NC.builtin_blob("nfkc"), NC.builtin_blob("nfkc_cf"), NC.builtin_blob("nmt_nfkc_cf"), NC.unicodedata_blob("NFD", True),
                 NC.unicodedata_blob("NFC", False), NC.unicodedata_blob(None, True), NC.custom_blob()):
        assert hostcore.hz_charsmap_ascii_table_check(blob) == 0


def _chain_steps(chain):
    bs = {s["name"]: s for s in GOLDEN["bert_steps"] + GOLDEN["other_steps"]}
    R = lambda name: ("regex", bs[name]["search"], bs[name]["replace"], bs[name]["global_replace"])
    return {
        "bert": lambda: [R("del_control_chars_regex"), R("replace_whitespace_regex"), R("handle_chinese_chars_regex"), ("charsmap", NC.unicodedata_blob("NFD", False), None, None),
                         R("strip_accents_regex"), ("charsmap", NC.unicodedata_blob(None, True), None, None)],
        "custom": lambda: [R("replace_whitespace_regex"), ("charsmap", NC.custom_blob(), None, None), R("del_control_chars_regex"), ("charsmap", NC.custom_blob(), None, None)],
        "anchored": lambda: [R("replace_whitespace_regex"), R("add_prefix_whitespace_regex"), ("charsmap", NC.unicodedata_blob(None, True), None, None)],
        "expanding": lambda: [("regex", r"\s", "<$0>", True), ("charsmap", NC.builtin_blob("nfkc_cf"), None, None), R("replace_spaces_metaspace")],
    }[chain]()


def _oracle_chain(steps, ins, skips=None):
    cur = list(ins)
    for kind, a, b_, g in steps:
        if kind == "regex":
            cur = list(oracle.regex_normalize(a, b_, g, *cur, skips))
        else:
            cur = list(oracle.charsmap_normalize(a, *cur, skips))
    return cur


@pytest.mark.skipif(not oracle.pcre2_available(), reason="libpcre2-8 not present")
@pytest.mark.parametrize("chain", ["bert", "custom", "anchored", "expanding"])
def test_composed_chain_table_vs_oracle(chain):
    """The composed per-byte table of a chain (what all-ASCII strings take on the GPU) against the ops applied one after
    the other by the oracle: every string whose bytes all have a simple fate must come out as the table says."""
    import hostcore
    steps = _chain_steps(chain)
    T = hostcore.hz_chain_table([(0, _enc(a), _enc(b_), g) if kind == "regex" else (1, a, b"", 0) for kind, a, b_, g in steps])
    if chain == "anchored":
        assert T is None
        return
    assert T is not None
    if chain == "bert":      # lower-casing, tabs / newlines -> space, controls dropped, everything else itself
        assert T[ord("A")] == ord("a") and T[9] == 32 and T[10] == 32 and T[1] == 0xFE and T[0x7F] == 0xFE and T[ord("~")] == ord("~")
        assert not (T == 0xFF).any()
    if chain == "expanding":
        assert T[32] == 0xFF and T[9] == 0xFF and T[ord("Q")] == ord("q")
    raw = [bytes([a]) for a in range(128)] + [r for r in NC.ascii_corpus(seed=41, n=1500) if r.isascii()]
    ins = NC.pack(raw)
    exp = NC.unpack(*_oracle_chain(steps, ins)[:3])
    checked = 0
    for r, x in zip(raw, exp):
        fates = T[np.frombuffer(r, np.uint8)] if r else np.zeros(0, np.uint8)
        if (fates == 0xFF).any():
            continue
        assert bytes(fates[fates != 0xFE]) == x, r
        checked += 1
    assert checked > 100


@pytest.mark.parametrize("search,replace", UNSUPPORTED)
def test_unsupported_patterns_are_refused(search, replace):
    """No CPU fallback: a pattern outside the single-character set fails at create, with or without a GPU."""
    import ctypes as C
    h = C.c_void_p()
    s, r = _enc(search), _enc(replace)
    rc = K.lib().b200tok_regexnorm_create(s, C.c_int64(len(s)), r, C.c_int64(len(r)), 1, 0, C.byref(h))
    assert rc == K.E_UNSUPPORTED and not h
    assert K.lib().b200tok_last_error()


def test_charsmap_flags_and_bad_blobs_are_refused():
    import ctypes as C
    h = C.c_void_p()
    blob = NC.builtin_blob("nfkc")
    assert K.lib().b200tok_charsmap_create(blob, C.c_int64(len(blob)), 1, 0, 0, 0, C.byref(h)) == K.E_UNSUPPORTED
    assert K.lib().b200tok_charsmap_create(blob, C.c_int64(3), 0, 0, 0, 0, C.byref(h)) == K.E_INVALID
    bad = (2 ** 31).to_bytes(4, "little") + blob[4:]
    assert K.lib().b200tok_charsmap_create(bad, C.c_int64(len(bad)), 0, 0, 0, 0, C.byref(h)) == K.E_INVALID
    assert K.lib().b200tok_normalize_run(None, None, None, C.c_int64(0), None, C.c_int64(0), None, None, None, None, C.c_int64(0), None, 0, None) == K.E_INVALID


# ------------------------------------------------------------------ GPU tier
def _strings_in(raw):
    b, e, c = NC.pack(raw)
    return [b, e, c]


@pytest.mark.gpu
@pytest.mark.parametrize("step", STEPS, ids=lambda s: s["name"])
def test_gpu_regex_normalization_vs_oracle(step, norm_path):
    from openvino_tokenizers_b200 import ops
    raw = NC.corpus(seed=21, n=3000, max_len=90) + NC.ascii_corpus(seed=24, n=600)
    ins = _strings_in(raw)
    skips = (np.arange(len(raw)) % 5 == 1)
    op = ops.RegexNormalization(step["global_replace"])
    pat = [np.frombuffer(_enc(step["search"]), np.uint8), np.frombuffer(_enc(step["replace"]), np.uint8)]
    exp = oracle.regex_normalize(step["search"], step["replace"], step["global_replace"], *ins)
    got = op.evaluate(ins + pat)
    assert NC.unpack(*got[:3]) == NC.unpack(*exp)
    assert (got[0] == exp[0]).all() and (got[1] == exp[1]).all()
    exp = oracle.regex_normalize(step["search"], step["replace"], step["global_replace"], *ins, skips)
    got = op.evaluate(ins + [skips] + pat)
    assert len(got) == 4 and got[3] is skips
    assert NC.unpack(*got[:3]) == NC.unpack(*exp)
    assert op.launches >= 6          # two calls, each at least lengths + size + write


@pytest.mark.gpu
def test_gpu_regex_normalization_reference_vectors():
    from openvino_tokenizers_b200 import ops
    ran = 0
    for case in GOLDEN["regex_normalization"]:
        ins = _strings_in([case["text"].encode()])
        pat = [np.frombuffer(_enc(case["search"]), np.uint8), np.frombuffer(_enc(case["replace"]), np.uint8)]
        op = ops.RegexNormalization(case["global_replace"])
        if case["search"].startswith(" ([") :
            with pytest.raises(ops.B200TokError):
                op.evaluate(ins + pat)
            continue
        got = op.evaluate(ins + pat)
        assert NC.unpack(*got[:3]) == [case["expected"].encode()], case
        ran += 1
    assert ran == len(GOLDEN["regex_normalization"]) - 1


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["nfkc_cf", "nmt_nfkc", "nfd_cf", "nfd", "nfc", "casefold", "empty", "custom"])
def test_gpu_charsmap_normalization_vs_oracle(which, norm_path):
    from openvino_tokenizers_b200 import ops
    blob = {"custom": NC.custom_blob, "nfc": lambda: NC.unicodedata_blob("NFC", False), "nfkc_cf": lambda: NC.builtin_blob("nfkc_cf"), "nmt_nfkc": lambda: NC.builtin_blob("nmt_nfkc"), "nfd_cf": lambda: NC.unicodedata_blob("NFD", True),
            "nfd": lambda: NC.unicodedata_blob("NFD", False), "casefold": lambda: NC.unicodedata_blob(None, True), "empty": lambda: b""}[which]()
    raw = NC.corpus(seed=22, n=3000, malformed=600, max_len=90) + NC.ascii_corpus(seed=23, n=600)
    ins = _strings_in(raw)
    skips = (np.arange(len(raw)) % 4 == 2)
    exp = oracle.charsmap_normalize(blob, *ins)
    got = ops.CharsMapNormalization().evaluate(ins + [np.frombuffer(blob, np.uint8)])            # 4-input form: blob as a tensor
    assert NC.unpack(*got[:3]) == NC.unpack(*exp)
    exp = oracle.charsmap_normalize(blob, *ins, skips)
    got = ops.CharsMapNormalization(precompiled_charsmap=blob).evaluate(ins + [skips])           # attribute form + skips
    assert len(got) == 4 and NC.unpack(*got[:3]) == NC.unpack(*exp)
    if which == "casefold":
        raw = [c["text"].encode() for c in GOLDEN["casefold_utf8"]]
        got = ops.CharsMapNormalization(precompiled_charsmap=blob).evaluate(_strings_in(raw))
        assert NC.unpack(*got[:3]) == [c["expected"].encode() for c in GOLDEN["casefold_utf8"]]


@pytest.mark.gpu
@pytest.mark.parametrize("chain", ["bert", "custom", "anchored", "expanding"])
def test_gpu_chain_mixed_vs_oracle(chain, norm_path):
    """Chains over a batch that mixes all-ASCII strings (one composed byte table), strings the ops must run one by one
    (non-ASCII, bytes with multi-byte rules) and skip-flagged strings; with an anchored op the chain is not composable."""
    from openvino_tokenizers_b200 import ops
    steps = _chain_steps(chain)
    prepared = [ops.RegexNormalization(g).prepare(a, b_) if kind == "regex" else ops.CharsMapNormalization().prepare(a) for kind, a, b_, g in steps]
    # mostly non-ASCII strings (the ops run over the whole batch) and mostly ASCII ones (composed table + a gathered sub-list)
    for n_mixed, n_ascii in ((2500, 1500), (700, 3000)):
        raw = NC.corpus(seed=31, n=n_mixed, max_len=90) + NC.ascii_corpus(seed=32, n=n_ascii) + [b"", b"plain ascii only", b"TAB\tand\x01ctl"]
        ins = _strings_in(raw)
        for skips in (None, (np.arange(len(raw)) % 6 == 2)):
            exp = _oracle_chain(steps, ins, skips)
            got = ops.normalize_chain(prepared, ins + ([skips] if skips is not None else []))
            assert NC.unpack(*got[:3]) == NC.unpack(*exp[:3])
            assert (got[0] == exp[0]).all() and (got[1] == exp[1]).all()
    # all-ASCII batch: no string takes the op-by-op path
    raw = [r for r in NC.ascii_corpus(seed=33, n=800) if r.isascii()]
    ins = _strings_in(raw)
    got = ops.normalize_chain(prepared, ins)
    assert NC.unpack(*got[:3]) == NC.unpack(*_oracle_chain(steps, ins)[:3])


@pytest.mark.gpu
def test_gpu_normalizers_edges():
    from openvino_tokenizers_b200 import ops
    pat = [np.frombuffer(br"\s", np.uint8), np.frombuffer(b" ", np.uint8)]
    op = ops.RegexNormalization(True)
    got = op.evaluate([np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.uint8)] + pat)         # no strings
    assert [len(x) for x in got] == [0, 0, 0]
    got = op.evaluate(_strings_in([b"", b"", b""]) + pat)                                                   # only empty strings
    assert NC.unpack(*got) == [b"", b"", b""]
    big = [("中" * 4000).encode(), b"a\tb" * 3000]                                                      # growth beyond the first guess
    han = [np.frombuffer(r"([\p{Han}])".encode(), np.uint8), np.frombuffer(b" $1 ", np.uint8)]
    exp = oracle.regex_normalize(r"([\p{Han}])", " $1 ", True, *_strings_in(big))
    assert NC.unpack(*ops.RegexNormalization(True).evaluate(_strings_in(big) + han)) == NC.unpack(*exp)
    # non-contiguous / overlapping element extents are allowed on input (each element is read independently)
    c = np.frombuffer(b"Hello\tWorld", np.uint8)
    b, e = np.array([6, 0, 0], np.int32), np.array([11, 5, 11], np.int32)
    got = op.evaluate([b, e, c] + pat)
    assert NC.unpack(*got) == [b"World", b"Hello", b"Hello World"]
    # 2-D shapes keep their shape (src/utils.cpp:187-188)
    got = op.evaluate([b.reshape(3, 1), e.reshape(3, 1), c] + pat)
    assert got[0].shape == (3, 1)


@pytest.mark.gpu
def test_gpu_bert_normalizer_chain_c2_size():
    """The BERT normaliser (hf_parser.py:84-102) as its five ops on the C2 batch (65 536 x 256 B), device results checked
    against the oracle on a sample of rows and by invariants on all of them."""
    import time
    from openvino_tokenizers_b200 import ops
    rng = np.random.default_rng(1234)
    B, L = 65536, 256
    chars = rng.integers(0x20, 0x7F, size=B * L, dtype=np.uint8)
    ctl = rng.random(B * L)
    chars[ctl < 0.01] = 0x09
    chars[(ctl >= 0.01) & (ctl < 0.015)] = 0x01
    chars[(ctl >= 0.015) & (ctl < 0.02)] = 0x0A
    b = (np.arange(B, dtype=np.int32) * L)
    e = b + L
    fold = NC.unicodedata_blob(None, True)
    nfd = NC.unicodedata_blob("NFD", False)
    chain = []
    for s in GOLDEN["bert_steps"][:3]:
        chain.append((ops.RegexNormalization(s["global_replace"]), [np.frombuffer(_enc(s["search"]), np.uint8), np.frombuffer(_enc(s["replace"]), np.uint8)], s))
    chain.append((ops.CharsMapNormalization(precompiled_charsmap=nfd), [], "nfd"))
    s = GOLDEN["bert_steps"][3]
    chain.append((ops.RegexNormalization(s["global_replace"]), [np.frombuffer(_enc(s["search"]), np.uint8), np.frombuffer(_enc(s["replace"]), np.uint8)], s))
    chain.append((ops.CharsMapNormalization(precompiled_charsmap=fold), [], "fold"))
    cur = [b, e, chars]
    for op, extra, _ in chain:
        cur = op.evaluate(cur + extra)
    fused = ops.normalize_chain([op for op, _, _ in chain], [b, e, chars])                 # one call, intermediates stay on the device
    t0 = time.perf_counter()
    fused = ops.normalize_chain([op for op, _, _ in chain], [b, e, chars])
    dt = time.perf_counter() - t0
    print(f"BERT normaliser chain, one call, host buffers, {B * L / 1e6:.1f} MB: {dt * 1e3:.1f} ms")
    assert all(np.array_equal(x, y) for x, y in zip(fused, cur))
    sample = np.arange(0, B, 257)
    ref = [b[sample], e[sample], chars]
    for _, _, s in chain:
        if s == "nfd": ref = list(oracle.charsmap_normalize(nfd, *ref))
        elif s == "fold": ref = list(oracle.charsmap_normalize(fold, *ref))
        else: ref = list(oracle.regex_normalize(s["search"], s["replace"], s["global_replace"], *ref))
    got = NC.unpack(cur[0][sample], cur[1][sample], cur[2])
    assert got == NC.unpack(*ref)
    out = cur[2]
    assert not ((out >= 0x41) & (out <= 0x5A)).any() and not (out < 0x20).any()           # lower-cased, no controls left
    assert int(cur[1][-1]) == out.size == B * L - int((chars == 0x01).sum())               # only the control bytes were dropped
