"""Byte-level shims and detokenizer tail: BytesToChars, CharsToBytes, FuzeRagged, UTF8Validate.
CPU tier: the oracle against the reference's known-answer material (tests/golden/shim_ops_layer_tests.json: the UTF8Validate
string list of the reference's tests/layer_tests.py:84-139 with the expectation its test uses, and the reference's literal
byte -> char table).  GPU tier: the CUDA ops through the C ABI against the oracle, bit-exact."""
import json
from pathlib import Path

import numpy as np
import pytest

import cases
from openvino_tokenizers_b200.strings import add_ragged_dimension, pack_strings, unpack_strings

GOLDEN = json.loads((Path(__file__).resolve().parent / "golden" / "shim_ops_layer_tests.json").read_text())


def test_oracle_utf8_validate_reference_vectors(oracle_mod):
    for c in GOLDEN["utf8_validate"]:
        s = bytes.fromhex(c["input_hex"])
        b, e, ch = pack_strings([s])
        got = oracle_mod.utf8_validate(b, e, ch, c["mode"] == "replace")
        assert bytes(got[2]).hex() == c["expected_hex"], (s, c["mode"])


def test_oracle_byte_char_map_is_the_reference_table(oracle_mod):
    table = [bytes.fromhex(h) for h in GOLDEN["bytes_to_chars_table_hex"]]
    b, e, ch = pack_strings([bytes(range(256))])
    rb, re_ = add_ragged_dimension(b, e)
    ob, oe, oc = oracle_mod.bytes_to_chars(rb, re_, b, e, ch)
    assert bytes(oc) == b"".join(table)
    back = oracle_mod.chars_to_bytes(rb, re_, ob, oe, oc)
    assert bytes(back[2]) == bytes(range(256)) and back[0].tolist() == [0] and back[1].tolist() == [256]


def test_host_scan_utf8_validate_and_bytes_to_chars_vs_oracle(oracle_mod):
    """UTF8Validate and BytesToChars run on the warp-per-string scan kernel of the normalisers; its per-position step
    (tok_core.cuh norm_eval, compiled for the host) against the oracle's restatement of the reference loops."""
    import hostcore
    for c in GOLDEN["utf8_validate"]:
        s = bytes.fromhex(c["input_hex"])
        b, e, ch = pack_strings([s])
        got = hostcore.hz_normalize(3, b"", b"", c["mode"] == "replace", b, e, ch)
        assert bytes(got[2]).hex() == c["expected_hex"], (s, c["mode"])
    rng = np.random.default_rng(35)
    alphabet = [b"a", b" ", "é".encode(), "€".encode(), "😁".encode(), b"\x80", b"\xc3", b"\xe2\x82", b"\xf0\x9f", b"\xc0\x80", b"\xff", b"\xed\xa0\x80",
                b"\xe0\x80\x80", b"\xf0\x80\x80\x80", b"\xf8", b"\xf4\x90\x80\x80"]
    strings = [b"".join(alphabet[i] for i in rng.integers(0, len(alphabet), size=int(rng.integers(0, 50)))) for _ in range(3000)]
    strings += [s.encode() for s in cases.EDGE_STRINGS] + [b"", b"\xe2", b"\xf0\x9f\x98"]
    b, e, ch = pack_strings(strings)
    for mode in (False, True):
        exp = oracle_mod.utf8_validate(b, e, ch, mode)
        got = hostcore.hz_normalize(3, b"", b"", mode, b, e, ch)
        assert [bytes(x) for x in unpack_strings(*got)] == [bytes(x) for x in unpack_strings(*exp)], mode
    strings = [bytes(rng.integers(0, 256, size=int(rng.integers(0, 70)), dtype=np.uint8)) for _ in range(2000)] + [b"", bytes(range(256))]
    b, e, ch = pack_strings(strings)
    rb, re_ = add_ragged_dimension(b, e)
    skips = rng.integers(0, 2, size=len(strings)).astype(np.uint8)
    exp = oracle_mod.bytes_to_chars(rb, re_, b, e, ch, skips.astype(bool))
    got = hostcore.hz_normalize(2, b"", b"", 0, b, e, ch, skips)
    assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1]) and np.array_equal(got[2], exp[2])
    # CharsToBytes, element level (one element per row), on the characters BytesToChars produced; every 5th element is cut
    # right after a lead byte, so its last character reads the follower from beyond the element (src/chars_to_bytes.cpp:52-60)
    cb, ce, cc = oracle_mod.bytes_to_chars(rb, re_, b, e, ch)
    cb, ce = cb.copy(), ce.copy()
    for k in range(0, len(cb), 5):
        if ce[k] - cb[k] >= 2 and cc[ce[k] - 1] < 0xC0 and cc[ce[k] - 2] >= 0xC0:      # ends with a 2-byte character: drop its follower
            ce[k] -= 1
    exp = oracle_mod.chars_to_bytes(rb, re_, cb, ce, cc)
    got = hostcore.hz_normalize(4, b"", b"", int(cc.size), cb, ce, cc)
    assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1]) and np.array_equal(got[2], exp[2])


@pytest.fixture(scope="module")
def ops():
    from openvino_tokenizers_b200 import ops as O
    return O


def _ragged(rng, strings, rows):
    """Pack strings as elements and cut them into `rows` contiguous rows (some empty)."""
    b, e, ch = pack_strings(strings)
    cuts = np.sort(rng.integers(0, len(strings) + 1, size=rows - 1)) if rows > 1 else np.array([], np.int64)
    rb = np.concatenate([[0], cuts]).astype(np.int32)
    re_ = np.concatenate([cuts, [len(strings)]]).astype(np.int32)
    return rb, re_, b, e, ch


@pytest.mark.gpu
def test_gpu_bytes_to_chars_and_back(ops, oracle_mod, norm_path):
    rng = np.random.default_rng(31)
    strings = [bytes(rng.integers(0, 256, size=int(rng.integers(0, 40)), dtype=np.uint8)) for _ in range(3000)] + [b"", bytes(range(256))]
    for rows in (1, 7, 500):
        rb, re_, b, e, ch = _ragged(rng, strings, rows)
        for skips in (None, rng.integers(0, 2, size=len(strings)).astype(bool)):
            exp = oracle_mod.bytes_to_chars(rb, re_, b, e, ch, skips)
            got = ops.BytesToChars().evaluate([rb, re_, b, e, ch] + ([skips] if skips is not None else []))
            assert np.array_equal(got[2], exp[0]) and np.array_equal(got[3], exp[1]) and np.array_equal(got[4], exp[2])
        # ... and back: CharsToBytes of the unskipped result restores the bytes, one string per row
        exp = oracle_mod.bytes_to_chars(rb, re_, b, e, ch)
        back_exp = oracle_mod.chars_to_bytes(rb, re_, exp[0], exp[1], exp[2])
        back = ops.CharsToBytes().evaluate([rb, re_, exp[0], exp[1], exp[2]])
        for k in range(3):
            assert np.array_equal(back[k], back_exp[k])
        rows_bytes = [b"".join(strings[int(rb[r]):int(re_[r])]) for r in range(rows)]
        assert [bytes(x) for x in unpack_strings(back[0], back[1], back[2])] == rows_bytes
        # elements cut right after a lead byte: the follower is read from beyond the element (src/chars_to_bytes.cpp:52-60)
        cb, ce, cc = exp[0].copy(), exp[1].copy(), exp[2]
        for k in range(0, len(cb), 5):
            if ce[k] - cb[k] >= 2 and cc[ce[k] - 1] < 0xC0 and cc[ce[k] - 2] >= 0xC0:
                ce[k] -= 1
        back_exp = oracle_mod.chars_to_bytes(rb, re_, cb, ce, cc)
        back = ops.CharsToBytes().evaluate([rb, re_, cb, ce, cc])
        for k in range(3):
            assert np.array_equal(back[k], back_exp[k])


@pytest.mark.gpu
def test_gpu_fuze_ragged(ops, oracle_mod):
    rng = np.random.default_rng(32)
    strings = [b"x" * int(rng.integers(0, 9)) for _ in range(1000)]
    for rows in (1, 13, 400):
        rb, re_, b, e, _ = _ragged(rng, strings, rows)
        keep = re_ > rb          # the reference reads element re[r] for an empty row: keep the test inside the buffer
        rb2, re2 = rb[keep], re_[keep]
        exp = oracle_mod.fuze_ragged(rb2, re2, b, e)
        got = ops.FuzeRagged().evaluate([rb2, re2, b, e])
        assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1])


@pytest.mark.gpu
def test_gpu_utf8_validate(ops, oracle_mod, norm_path):
    for c in GOLDEN["utf8_validate"]:
        s = bytes.fromhex(c["input_hex"])
        b, e, ch = pack_strings([s])
        got = ops.UTF8Validate(c["mode"] == "replace").evaluate([b, e, ch])
        assert bytes(got[2]).hex() == c["expected_hex"], (s, c["mode"])
    rng = np.random.default_rng(33)
    alphabet = [b"a", b" ", "é".encode(), "€".encode(), "😁".encode(), b"\x80", b"\xc3", b"\xe2\x82", b"\xf0\x9f", b"\xc0\x80", b"\xff", b"\xed\xa0\x80"]
    strings = [b"".join(alphabet[i] for i in rng.integers(0, len(alphabet), size=int(rng.integers(0, 30)))) for _ in range(4000)]
    strings += [s.encode() for s in cases.EDGE_STRINGS] + [b"", b"\xe2", b"\xf0\x9f\x98"]
    b, e, ch = pack_strings(strings)
    for mode in (False, True):
        exp = oracle_mod.utf8_validate(b, e, ch, mode)
        got = ops.UTF8Validate(mode).evaluate([b, e, ch])
        for k in range(3):
            assert np.array_equal(got[k], exp[k]), (k, mode)
    # offsets that do not start at 0: the reference's cursor starts at begins[0]
    exp = oracle_mod.utf8_validate(b[5:], e[5:], ch, True)
    got = ops.UTF8Validate(True).evaluate([b[5:], e[5:], ch])
    assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1])
    assert np.array_equal(got[2][int(b[5]):], exp[2][int(b[5]):])


@pytest.mark.gpu
def test_chars_to_bytes_device_buffers_validate_row_partition():
    """Device-resident inputs get the same contract check as host inputs: rows that do not cover the elements contiguously and
    in order are refused (E_UNSUPPORTED) instead of yielding extents taken from the wrong elements."""
    import ctypes as C
    import torch
    from openvino_tokenizers_b200 import _capi as K
    dev = torch.device("cuda", 0)
    b, e, c = pack_strings([b"ab", b"c", b"def", b"g"])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    db, de, dc = t(b), t(e), t(np.concatenate([c, np.zeros(64, np.uint8)]))
    ob, oe, oc = torch.empty(4, dtype=torch.int32, device=dev), torch.empty(4, dtype=torch.int32, device=dev), torch.empty(64, dtype=torch.uint8, device=dev)

    def run(rb, re_):
        drb, dre = t(np.asarray(rb, np.int32)), t(np.asarray(re_, np.int32))
        rin = K.RaggedStrings(drb.data_ptr(), dre.data_ptr(), len(rb), db.data_ptr(), de.data_ptr(), 4, dc.data_ptr(), len(c), None, K.MEM_DEVICE)
        n = C.c_int64(0)
        rcs = []
        for fn in (K.lib().b200tok_chars_to_bytes_run, K.lib().b200tok_bytes_to_chars_run):
            rcs.append(fn(0, C.byref(rin), C.c_void_p(ob.data_ptr()), C.c_void_p(oe.data_ptr()), C.c_void_p(oc.data_ptr()), C.c_int64(64), C.byref(n), None))
        return rcs
    assert run([0, 2], [2, 4]) == [0, 0]
    for rb, re_ in (([0, 3], [2, 4]), ([2, 0], [4, 2]), ([0, 2], [2, 3]), ([1, 2], [2, 4])):      # gap / reordered / short / late start
        assert run(rb, re_) == [K.E_UNSUPPORTED, K.E_UNSUPPORTED], (rb, re_)
