"""CPU tier: the C-ABI library loads, exports every symbol include/b200tok.h declares, and fails loudly
(no CPU fallback) when there is no GPU."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def K():
    from openvino_tokenizers_b200 import _capi, build
    build.build()
    return _capi


def test_header_symbols_are_exported(K):
    header = (ROOT / "include" / "b200tok.h").read_text()
    declared = sorted(set(re.findall(r"B200TOK_API\s+[\w\s\*]+?\b(b200tok_\w+)\s*\(", header)))
    assert declared == sorted(K.EXPORTED_SYMBOLS)
    lib = K.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.b200tok_version() >= 100


def test_no_oracle_in_product():
    """The product package must not import or link the oracle."""
    for p in (ROOT / "openvino_tokenizers_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".cpp", ".hpp"):
            text = p.read_text(errors="ignore")
            assert "liboracle" not in text and "import oracle" not in text and "from oracle" not in text, p


def test_pattern_and_vocab_errors_do_not_need_a_gpu(K):
    lib = K.lib()
    h = C.c_void_p()
    d = K.RegexSplitDesc(b"(a|b)+c", 7, b"isolate", 0, -1, 0)
    assert lib.b200tok_regexsplit_create(C.byref(d), C.byref(h)) == K.E_UNSUPPORTED
    assert b"pattern" in lib.b200tok_last_error()
    d = K.RegexSplitDesc(rb"\s+", 3, b"sideways", 0, -1, 0)
    assert lib.b200tok_regexsplit_create(C.byref(d), C.byref(h)) == K.E_INVALID


def test_create_fails_loudly_without_gpu(K):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = K.lib()
    assert lib.b200tok_device_count() == 0
    h = C.c_void_p()
    d = K.RegexSplitDesc(rb"\s+", 3, b"remove", 0, -1, 0)
    assert lib.b200tok_regexsplit_create(C.byref(d), C.byref(h)) == K.E_CUDA
    assert b"no CPU fallback" in lib.b200tok_last_error()
    from openvino_tokenizers_b200 import ops
    with pytest.raises(ops.B200TokError):
        ops.RegexSplit("remove").with_pattern(r"\s+")
