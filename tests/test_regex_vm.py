"""CPU tier: the regex machine for general split patterns (csrc/regex_vm.cuh + csrc/regex_compile.cpp), compiled for the host by
tests/harness, against PCRE2 (the oracle's RegexSplit) on the split patterns of the reference's model list — CLIP
(tests/layer_tests.py:311-315), gpt-4o / Qwen2 / DeepSeek (tests/tokenizers_test.py:105-137) — in all five split behaviours."""
import json
from pathlib import Path

import numpy as np
import pytest

import cases
import hostcore
from test_gpu_parity import VM_PATTERNS

GOLDEN = Path(__file__).resolve().parent / "golden"
EXTRA = [r"\d+|\D+", r"a.c|^x|y$", r"(ab|a)(c|bcd)?", r"\s?\w{2,4}", r"(?i)straße|[a-f]+", r"[\x{4E00}-\x{9FFF}]|\p{Lu}\p{Ll}*", r"x{2}y{0,2}(z|w){1,3}", r"[^a-c\s]+|\s"]


def texts():
    t = cases.EDGE_STRINGS + cases.long_prompts() + ["<|startoftext|>a photo of a cat<|endoftext|>", "He'LL ſay K'S 'ſ 'Kelvin", "abc abcd ac abcdbcd",
                                                     "x  y\nz y", "١٢٣٤٥٦٧ 12345678", "天気がいい日 カタカナ ひらがな 漢字", "STRASSE Straße ABCdef", "xxyyzwz xxw xyz"]
    rng = np.random.default_rng(1)
    t += [bytes(rng.integers(0x20, 0x7F, size=int(rng.integers(1, 300)), dtype=np.uint8)).decode() for _ in range(120)]
    b = cases.mixed_utf8_batch(48, 256, seed=9)
    t += [bytes(b[4][b[2][i]:b[3][i]]).decode() for i in range(48)]
    return [x for x in t if x]


@pytest.mark.parametrize("pattern", list(VM_PATTERNS.values()) + EXTRA, ids=list(VM_PATTERNS) + [f"extra{i}" for i in range(len(EXTRA))])
def test_regex_machine_equals_pcre2(oracle_mod, pattern):
    for behaviour in ("isolate", "remove", "mergedwithprevious", "mergedwithnext", "contiguous"):
        o = oracle_mod.SplitOracle(pattern, behaviour)
        for t in texts():
            exp = o(*cases.batch_from_strings([t]))
            assert hostcore.split(pattern, behaviour, False, -1, t.encode()) == list(zip(exp[2].tolist(), exp[3].tolist())), (behaviour, t[:60])


def test_clip_golden_vectors_through_the_regex_machine():
    """The 10 CLIP vectors of the reference's RegexSplit tests (tests/layer_tests.py:311-389) — refused in round 1."""
    g = json.loads((GOLDEN / "regex_split_layer_tests.json").read_text())
    n = 0
    for case in g["cases"]:
        if "startoftext" not in case["pattern"] or not case["text"]:
            continue
        data = case["text"].encode()
        pieces = hostcore.split(case["pattern"], case["behaviour"], case["invert"], case["max_splits"], data)
        assert [data[b:e].decode() for b, e in pieces] == case["expected"], case
        n += 1
    assert n >= 9


def test_unsupported_syntax_is_refused():
    for pat in (r"(foo|bar)+baz", r"\bword\b", r"a*?b", r"(?<=x)y", r"\p{Han}+", r"(?=ab)a", r"[[:alpha:]]+", r"(a)\1", r"[\S\d]", r"(", r"a{3,2}"):
        with pytest.raises(ValueError):
            hostcore.split(pat, "isolate", False, -1, b"some text")
