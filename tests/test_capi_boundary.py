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
    d = K.RegexSplitDesc(b"(foo|bar)+baz", 13, b"isolate", 0, -1, 0)      # a group repeated without bound: outside the compiled syntax
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


def test_ctypes_mirrors_match_the_header_layout(K, tmp_path):
    """The ctypes structures in _capi.py must have the size the C compiler gives the structs of include/b200tok.h (a plain-C
    translation unit: the header is a C ABI, no C++ needed)."""
    import subprocess
    pairs = [("b200tok_strings", K.Strings), ("b200tok_ragged_strings", K.RaggedStrings), ("b200tok_ragged_strings_out", K.RaggedStringsOut),
             ("b200tok_ragged_ids", K.RaggedIds), ("b200tok_regexsplit_desc", K.RegexSplitDesc), ("b200tok_bpe_desc", K.BpeDesc),
             ("b200tok_wordpiece_desc", K.WordpieceDesc), ("b200tok_vocabenc_desc", K.VocabEncDesc), ("b200tok_vocabdec_desc", K.VocabDecDesc),
             ("b200tok_decoded", K.Decoded), ("b200tok_ragged_i32", K.RaggedI32), ("b200tok_post_desc", K.PostDesc), ("b200tok_peer_out", K.PeerOut),
             ("b200tok_peer_pull", K.PeerPull)]
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "b200tok.h"\nint main(void) {\n' +
                   "".join(f'  printf("%zu\\n", sizeof({c}));\n' for c, _ in pairs) + "  return 0;\n}\n")
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I", str(ROOT / "include"), "-o", str(exe), str(src)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    for (cname, ct), n in zip(pairs, sizes):
        assert C.sizeof(ct) == n, (cname, C.sizeof(ct), n)


def test_peer_pull_and_pack_reject_bad_layouts_before_touching_a_device(K):
    """Argument errors of the all-gatherv-by-pull entry points come back as B200TOK_E_INVALID with a message, GPU or not."""
    lib = K.lib()
    buf = np.zeros(64, np.int32)
    q = K.PeerPull()
    q.world, q.rank, q.wire16, q.skip_self_ids, q.slot_capacity, q.rows_per_rank = 2, 0, 1, 1, 16, 4
    q.ids = q.begins = q.ends = buf.ctypes.data
    assert lib.b200tok_peer_pull_run(0, C.byref(q), None) == K.E_INVALID and b"missing source buffer" in lib.b200tok_last_error()
    q.slot_capacity = 12                                        # not a multiple of 8
    assert lib.b200tok_peer_pull_run(0, C.byref(q), None) == K.E_INVALID and b"multiple of 8" in lib.b200tok_last_error()
    q.slot_capacity, q.world = 16, 9
    assert lib.b200tok_peer_pull_run(0, C.byref(q), None) == K.E_INVALID
    q.world, q.rank = 2, 2
    assert lib.b200tok_peer_pull_run(0, C.byref(q), None) == K.E_INVALID
    assert lib.b200tok_peer_pull_run(0, None, None) == K.E_INVALID
    assert lib.b200tok_peer_pack_run(0, None, None, C.c_int64(16), None, None) == K.E_INVALID
    assert lib.b200tok_peer_pack_run(0, C.c_void_p(buf.ctypes.data + 4), C.c_void_p(buf.ctypes.data), C.c_int64(16), C.c_void_p(buf.ctypes.data), None) == K.E_INVALID
    assert b"16-byte aligned" in lib.b200tok_last_error()
    st = K.make_strings((np.zeros(1, np.int32), np.ones(1, np.int32), b"x"), [])
    assert lib.b200tok_regexsplit_set_skip_tokens(None, C.byref(st)) == K.E_INVALID
