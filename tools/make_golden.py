#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/ (run in the build container, where
/root/reference and HuggingFace `tokenizers` are available; the fixtures travel, this script is provenance).

  regex_split_layer_tests.json  the reference's own known-answer vectors for RegexSplit, extracted by evaluating
                                the parametrize list of tests/layer_tests.py:331-389 against the RegexSplitStep
                                factory methods of python/openvino_tokenizers/tokenizer_pipeline.py:354-470
  post_ops_layer_tests.json     the reference's known-answer vectors for RaggedToDense and CombineSegments
                                (tests/layer_tests.py:497-644)
  special_tokens_split_layer_tests.json  the reference's known-answer vectors for SpecialTokensSplit (tests/layer_tests.py:405-457)
  normalization_layer_tests.json  RegexNormalization / case-fold known-answer vectors (tests/layer_tests.py:226-290) and the
                                normaliser pattern set of the converter
  shim_ops_layer_tests.json     UTF8Validate known-answer strings (tests/layer_tests.py:84-139) and the byte -> char table
  hf_<vocab>.json               ids produced by HuggingFace `tokenizers` for the frozen synthetic vocabularies
                                (second oracle; the reference reports 100 % agreement with HF for these families)
"""
import ast
import json
import re
import sys
from dataclasses import dataclass, field  # noqa: F401  (used by the exec'd reference snippet)
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
REF = Path("/root/reference")
GOLDEN = ROOT / "tests" / "golden"


def reference_regex_cases():
    src = (REF / "python/openvino_tokenizers/tokenizer_pipeline.py").read_text()
    a = src.index("@dataclass\nclass RegexSplitStep")
    b = src.index("    def get_ov_subgraph", a)
    ns = {"dataclass": dataclass, "field": field, "PreTokenizatinStep": object}
    exec(src[a:b], ns)
    RegexSplitStep = ns["RegexSplitStep"]
    lt = (REF / "tests/layer_tests.py").read_text()
    a = lt.index("clip_regex_pattern = (")
    b = lt.index("@pytest.mark.parametrize", a)
    ns2 = {"re": re, "RegexSplitStep": RegexSplitStep}
    exec(lt[a:b], ns2)
    tree = ast.parse(lt)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "test_regex_split":
            lst = node.decorator_list[0].args[1]
            cases = eval(compile(ast.Expression(lst), "layer_tests", "eval"), ns2)
            return cases
    raise RuntimeError("test_regex_split not found")


def make_regex_golden():
    import hostcore
    from oracle import SplitOracle
    from openvino_tokenizers_b200.strings import add_ragged_dimension, pack_strings, unpack_strings
    out = []
    for text, expected, layer in reference_regex_cases():
        pat, beh, inv, ms = layer.split_pattern, layer.behaviour, layer.invert, layer.max_splits
        try:
            hostcore.split(pat, beh, inv, ms, b"x")
            supported = True
        except ValueError:
            supported = False
        b, e, c = pack_strings([text])
        rb, re_ = add_ragged_dimension(b, e)
        r = SplitOracle(pat, beh, inv, ms)(rb, re_, b, e, c)
        got = [p.decode() for p in unpack_strings(r[2], r[3], c)]
        assert got == list(expected), (text, expected, got)   # the oracle reproduces the reference's vector
        out.append(dict(text=text, expected=list(expected), pattern=pat, behaviour=beh, invert=bool(inv),
                        max_splits=ms, gpu_supported=supported))
    (GOLDEN / "regex_split_layer_tests.json").write_text(json.dumps(
        dict(source="reference tests/layer_tests.py:331-389", cases=out), ensure_ascii=False, indent=1))
    print("regex cases", len(out), "gpu-supported", sum(c["gpu_supported"] for c in out))


def reference_parametrize(func_name: str):
    """The (evaluated) argvalues list of the @pytest.mark.parametrize decorator of a test in the reference's layer_tests.py."""
    import numpy as np  # noqa: F401
    lt = (REF / "tests/layer_tests.py").read_text()
    tree = ast.parse(lt)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == func_name:
            lst = node.decorator_list[0].args[1]
            return eval(compile(ast.Expression(lst), "layer_tests", "eval"), {"np": __import__("numpy")})
    raise RuntimeError(func_name + " not found")


def make_post_golden():
    """Known-answer vectors of RaggedToDense (reference tests/layer_tests.py:497-591) and CombineSegments (:594-644);
    each is checked against the oracle before it is written.  (The reference holds no Truncate vectors.)"""
    import numpy as np
    import oracle
    r2d = []
    for inp, expected in reference_parametrize("test_ragged_to_dense"):
        pad_right = inp["pad_right"] if "pad_right" in inp else inp["padding_side"] == "right"   # input [5] has priority over the attribute
        got, _ = oracle.ragged_to_dense(inp["begins"], inp["ends"], inp["data"], inp["padding_size"], inp["value"], pad_right)
        assert got.tolist() == expected, (inp, expected, got)
        r2d.append(dict(inputs=inp, expected=expected))
    comb = []
    for segs, expected in reference_parametrize("test_combine_segments"):
        got = oracle.combine_segments([(s["begins"], s["ends"], s["data"]) for s in segs], np.arange(len(segs)))
        assert got[0].tolist() == expected["begins"] and got[1].tolist() == expected["ends"] and got[2].tolist() == expected["data"]
        comb.append(dict(segments=segs, expected=expected))
    (GOLDEN / "post_ops_layer_tests.json").write_text(json.dumps(
        dict(source="reference tests/layer_tests.py:497-644", ragged_to_dense=r2d, combine_segments=comb), indent=1))
    print("ragged_to_dense cases", len(r2d), "combine_segments cases", len(comb))


def make_special_golden():
    """The reference's SpecialTokensSplit known-answer vectors (tests/layer_tests.py:405-457).  The split pattern of every
    case is built by the reference's own code (SpecialToken / SpecialTokensSplit.get_ov_subgraph of
    python/openvino_tokenizers/tokenizer_pipeline.py:80-158, exec'd up to the node construction) and the oracle must
    reproduce the expected pieces and skip flags before the case is written."""
    from collections import defaultdict
    import oracle
    from openvino_tokenizers_b200.strings import add_ragged_dimension, pack_strings, unpack_strings
    utils_src = (REF / "python/openvino_tokenizers/utils.py").read_text()
    a = utils_src.index("def quote_meta(")
    b = utils_src.index("\n\n\n", a)
    ns = {"Union": __import__("typing").Union}
    exec(utils_src[a:b], ns)
    src = (REF / "python/openvino_tokenizers/tokenizer_pipeline.py").read_text()
    a = src.index("@dataclass(frozen=True, order=True)\nclass SpecialToken:")
    b = src.index("@dataclass\nclass SpecialTokensSplit", a)
    ns.update({"dataclass": dataclass, "field": field})
    exec(src[a:b], ns)
    SpecialToken = ns["SpecialToken"]
    a = src.index("        grouped_tokens = defaultdict(list)", b)
    b2 = src.index("        input_nodes.extend(create_string_constant_node(split_pattern))", a)
    body = "def build_pattern(self):\n" + src[a:b2] + "        return split_pattern\n"
    ns.update({"defaultdict": defaultdict})
    exec(body, ns)

    class Step:
        def __init__(self, toks):
            self.special_tokens = sorted(toks, reverse=True)
    out = []
    lt = (REF / "tests/layer_tests.py").read_text()
    tree = ast.parse(lt)
    cases_ = None
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "test_special_tokens_split":
            cases_ = eval(compile(ast.Expression(node.decorator_list[0].args[1]), "layer_tests", "eval"), {"SpecialToken": SpecialToken})
    for toks, text, expected, expected_skips in cases_:
        pattern = ns["build_pattern"](Step(toks))
        assert pattern == oracle.special_tokens_pattern([(t.text, t.strip_left, t.strip_right) for t in toks]), pattern
        bb, ee, cc = pack_strings([text])
        rb, re_ = add_ragged_dimension(bb, ee)
        r = oracle.SpecialTokensSplitOracle(pattern)(rb, re_, bb, ee, cc)
        got = [p.decode() for p in unpack_strings(r[2], r[3], cc)]
        assert got == list(expected) and r[4].tolist() == list(expected_skips), (text, got, r[4])
        out.append(dict(tokens=[[t.text, t.strip_left, t.strip_right] for t in toks], pattern=pattern, text=text,
                        expected=list(expected), expected_skips=list(expected_skips)))
    (GOLDEN / "special_tokens_split_layer_tests.json").write_text(json.dumps(
        dict(source="reference tests/layer_tests.py:405-457", cases=out), ensure_ascii=False, indent=1))
    print("special tokens split cases", len(out))


def make_shim_golden():
    """UTF8Validate: the reference's own string list (tests/layer_tests.py:84-117) with the expectation its test uses
    (`bytes.decode(errors=mode)`, :131-139), checked against the oracle.  BytesToChars: the reference's literal byte -> char
    table (src/bytes_to_chars.cpp:11-268) as hex, checked against the oracle's generated map and against the converter's
    unicode_to_bytes (python/openvino_tokenizers/utils.py:196-210)."""
    import oracle
    from openvino_tokenizers_b200.strings import add_ragged_dimension, pack_strings
    lt = (REF / "tests/layer_tests.py").read_text()
    a = lt.index("utf8_validate_strings = [")
    b = lt.index("\n]\n", a) + 3
    ns = {}
    exec(lt[a:b], ns)
    out = []
    for sbytes in ns["utf8_validate_strings"]:
        for mode in ("ignore", "replace"):
            exp = sbytes.decode(errors=mode).encode()
            bb, ee, cc = pack_strings([sbytes])
            got = oracle.utf8_validate(bb, ee, cc, mode == "replace")
            assert bytes(got[2]) == exp, (sbytes, mode)
            out.append(dict(input_hex=sbytes.hex(), mode=mode, expected_hex=exp.hex()))
    cpp = (REF / "src/bytes_to_chars.cpp").read_text()
    body = cpp[cpp.index("create_bytes_to_chars_map"):cpp.index("}};")]
    table = [bytes(int(x) for x in e.replace(" ", "").split(",") if x) for e in re.findall(r"\{\s*([\d,\s]+)\}", body)][-256:]
    assert len(table) == 256
    bb, ee, cc = pack_strings([bytes(range(256))])
    rb, re_ = add_ragged_dimension(bb, ee)
    assert bytes(oracle.bytes_to_chars(rb, re_, bb, ee, cc)[2]) == b"".join(table)
    (GOLDEN / "shim_ops_layer_tests.json").write_text(json.dumps(
        dict(source="reference tests/layer_tests.py:84-139, src/bytes_to_chars.cpp:11-268", utf8_validate=out,
             bytes_to_chars_table_hex=[t.hex() for t in table]), indent=1))
    print("utf8_validate cases", len(out), "byte map entries", len(table))


def make_hf_golden():
    import numpy as np
    import cases
    from tokenizers import Tokenizer
    from openvino_tokenizers_b200 import assets as A
    rng = np.random.default_rng(99)
    rand = [bytes(rng.integers(0x20, 0x7F, size=int(rng.integers(1, 300)), dtype=np.uint8)).decode() for _ in range(60)]
    # special-token strings need SpecialTokensSplit upstream (out of scope); HF would extract them
    texts = [s for s in cases.EDGE_STRINGS if len(s) < 600 and "<|" not in s] + [p[:1500] for p in cases.long_prompts()] + rand
    for name in ("gpt2_synth", "llama3_synth"):
        a = A.load_bpe(name)
        hf = Tokenizer.from_str(a.hf_json)
        ids = [hf.encode(t, add_special_tokens=False).ids for t in texts]
        (GOLDEN / f"hf_{name}.json").write_text(json.dumps(dict(texts=texts, ids=ids), ensure_ascii=False))
        print(name, len(texts), "texts", sum(map(len, ids)), "ids")
    w = A.load_wordpiece("bert_synth")
    hf = Tokenizer.from_str(w.hf_json)
    btexts = [t.lower() for t in texts if t.isascii() and all(c >= " " or c in "\n\t\r" for c in t)]
    ids = [hf.encode(t, add_special_tokens=False).ids for t in btexts]
    (GOLDEN / "hf_bert_synth.json").write_text(json.dumps(dict(texts=btexts, ids=ids), ensure_ascii=False))
    print("bert_synth", len(btexts), "texts")


def make_norm_golden():
    """RegexNormalization known-answer vectors (reference tests/layer_tests.py:253-290): the parametrize list is evaluated
    against the reference's own RegexNormalizationStep dataclass (its classmethods define the patterns); the case-fold
    vectors (:226-250, utf-8 rows).  Also records the pattern set of the BERT normaliser (hf_parser.py:84-102)."""
    src = (REF / "python/openvino_tokenizers/tokenizer_pipeline.py").read_text()
    a = src.index("@dataclass\nclass RegexNormalizationStep")
    b = src.index("    def get_ov_subgraph", a)
    ns = {"dataclass": dataclass, "field": field, "NormalizationStep": object}
    exec(src[a:b], ns)
    Step = ns["RegexNormalizationStep"]
    lt = (REF / "tests/layer_tests.py").read_text()
    tree = ast.parse(lt)
    regex_cases, casefold = [], []
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "test_regex_normalization":
            for text, expected, layer in eval(compile(ast.Expression(node.decorator_list[0].args[1]), "layer_tests", "eval"), {"RegexNormalizationStep": Step}):
                regex_cases.append(dict(text=text, expected=expected, search=layer.regex_search_pattern, replace=layer.replace_term,
                                        global_replace=layer.global_replace))
        if isinstance(node, ast.FunctionDef) and node.name == "test_casefold_normalization":
            for text, expected, is_utf8 in eval(compile(ast.Expression(node.decorator_list[0].args[1]), "layer_tests", "eval"), {}):
                if is_utf8:
                    casefold.append(dict(text=text, expected=expected))
    bert = [dict(name=n, search=getattr(Step, n)().regex_search_pattern, replace=getattr(Step, n)().replace_term,
                 global_replace=getattr(Step, n)().global_replace)
            for n in ("del_control_chars_regex", "replace_whitespace_regex", "handle_chinese_chars_regex", "strip_accents_regex")]
    other = [dict(name=n, search=getattr(Step, n)().regex_search_pattern, replace=getattr(Step, n)().replace_term,
                  global_replace=getattr(Step, n)().global_replace)
             for n in ("add_prefix_whitespace_regex", "add_prefix_whitespace_to_not_whitespace_regex", "replace_spaces_metaspace")]
    p = Step.prepend_regex("\u2581")
    other.append(dict(name="prepend_regex", search=p.regex_search_pattern, replace=p.replace_term, global_replace=p.global_replace))
    (GOLDEN / "normalization_layer_tests.json").write_text(json.dumps(
        dict(source="reference tests/layer_tests.py:226-290, python/openvino_tokenizers/tokenizer_pipeline.py:223-278",
             regex_normalization=regex_cases, casefold_utf8=casefold, bert_steps=bert, other_steps=other), ensure_ascii=False, indent=1))
    print("regex normalization cases", len(regex_cases), "casefold", len(casefold), "steps", len(bert) + len(other))


if __name__ == "__main__":
    GOLDEN.mkdir(parents=True, exist_ok=True)
    which = sys.argv[1:] or ["regex", "hf", "post", "special", "shim", "norm"]
    if "regex" in which:
        make_regex_golden()
    if "hf" in which:
        make_hf_golden()
    if "post" in which:
        make_post_golden()
    if "special" in which:
        make_special_golden()
    if "shim" in which:
        make_shim_golden()
    if "norm" in which:
        make_norm_golden()
