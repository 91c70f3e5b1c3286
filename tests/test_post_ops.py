"""Post-tokenizer tail: Truncate, CombineSegments, RaggedToDense and their fusion.

CPU tier: the oracle against the reference's own known-answer vectors (tests/golden/post_ops_layer_tests.json, extracted
from the reference's tests/layer_tests.py:497-644) and hand-derived Truncate cases (the reference holds no Truncate
vectors: parity for it is pinned by the restatement of src/truncate.cpp:37-147 only).
GPU tier: the CUDA ops through the C ABI against the oracle (bit-exact) on the golden vectors and on random ragged input.
"""
import json
from pathlib import Path

import numpy as np
import pytest

GOLDEN = json.loads((Path(__file__).resolve().parent / "golden" / "post_ops_layer_tests.json").read_text())


def _pad_right(inp):
    return inp["pad_right"] if "pad_right" in inp else inp["padding_side"] == "right"


def _random_ragged(rng, rows, max_len, gaps=True):
    lens = rng.integers(0, max_len + 1, size=rows)
    gap = rng.integers(0, 3, size=rows) if gaps else np.zeros(rows, np.int64)
    begins = np.cumsum(np.concatenate([[0], (lens + gap)[:-1]])) + gap
    ends = begins + lens
    n = int(ends.max()) + 2 if rows else 0
    return begins.astype(np.int32), ends.astype(np.int32), rng.integers(0, 50000, size=n).astype(np.int32)


# ---------------------------------------------------------------- CPU tier: the oracle
def test_oracle_ragged_to_dense_reference_vectors(oracle_mod):
    for c in GOLDEN["ragged_to_dense"]:
        i = c["inputs"]
        out, mask = oracle_mod.ragged_to_dense(i["begins"], i["ends"], i["data"], i["padding_size"], i["value"], _pad_right(i))
        assert out.tolist() == c["expected"]
        lens = np.minimum(np.array(i["ends"]) - np.array(i["begins"]), i["padding_size"])
        assert mask.sum(axis=1).tolist() == lens.tolist()


def test_oracle_combine_segments_reference_vectors(oracle_mod):
    for c in GOLDEN["combine_segments"]:
        segs = [(s["begins"], s["ends"], s["data"]) for s in c["segments"]]
        b, e, x, ids = oracle_mod.combine_segments(segs, np.arange(len(segs)))
        assert (b.tolist(), e.tolist(), x.tolist()) == (c["expected"]["begins"], c["expected"]["ends"], c["expected"]["data"])
        assert len(ids) == len(x)


TRUNCATE_CASES = [
    # (pairs, max_length, side, mode, expected pairs) — derived by hand from src/truncate.cpp:57-144
    ([([0, 3], [3, 8])], 2, "right", "longest_first", [([0, 3], [2, 5])]),
    ([([0, 3], [3, 8])], 2, "left", "longest_first", [([1, 6], [3, 8])]),
    ([([0], [9]), ([0], [2])], 10, "right", "longest_first", [([0], [8]), ([0], [2])]),      # the comment's own example (:101-102)
    ([([0], [2]), ([0], [9])], 10, "right", "longest_first", [([0], [2]), ([0], [8])]),      # (:105-106)
    ([([0], [9]), ([0], [9])], 9, "right", "longest_first", [([0], [5]), ([0], [4])]),       # odd max_length: remainder to the first (:86-87)
    ([([0], [8]), ([0], [9])], 9, "left", "longest_first", [([4], [8]), ([4], [9])]),        # ... to the longer second
    ([([0], [20]), ([0], [3])], 10, "right", "only_first", [([0], [10]), ([0], [3])]),
    ([([0], [20]), ([0], [30])], 10, "left", "only_second", [([0], [20]), ([20], [30])]),
    ([([0], [4]), ([0], [5])], 10, "right", "longest_first", [([0], [4]), ([0], [5])]),      # fits: untouched (:83)
]


def test_oracle_truncate_hand_vectors(oracle_mod):
    for pairs, ml, side, mode, expected in TRUNCATE_CASES:
        got = oracle_mod.truncate(pairs, ml, side, mode)
        assert [(b.tolist(), e.tolist()) for b, e in got] == [(list(b), list(e)) for b, e in expected], (pairs, ml, side, mode)
    with pytest.raises(ValueError):
        oracle_mod.truncate([([0], [1])], 3, "middle", "longest_first")


# ---------------------------------------------------------------- GPU tier: CUDA ops vs the oracle
@pytest.fixture(scope="module")
def ops():
    from openvino_tokenizers_b200 import ops as O
    return O


@pytest.mark.gpu
def test_gpu_ragged_to_dense(ops, oracle_mod):
    for c in GOLDEN["ragged_to_dense"]:
        i = c["inputs"]
        ins = [np.array(i[k], np.int32) for k in ("begins", "ends", "data", "padding_size", "value")]
        if "pad_right" in i:
            ins.append(np.array(i["pad_right"], bool))
        out, mask = ops.RaggedToDense(pad_right=i["padding_side"] == "right").evaluate(ins)
        assert out.tolist() == c["expected"]
    rng = np.random.default_rng(11)
    for rows, max_len, target in [(1, 0, 4), (7, 9, 5), (300, 40, 33), (4096, 600, 512), (5, 3, 0)]:
        b, e, x = _random_ragged(rng, rows, max_len)
        for pad_right in (True, False):
            for pml in (False, True):
                exp = oracle_mod.ragged_to_dense(b, e, x, target, -7, pad_right, pml) if not pml else None
                got = ops.RaggedToDense(pad_right=pad_right, pad_max_length=pml).evaluate([b, e, x, np.int32(target), np.int32(-7)])
                if exp is not None:
                    assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1].astype(bool))
                else:   # pad_max_length copies target_dim elements from every row start (reads past short rows stay inside elems here)
                    idx = b[:, None].astype(np.int64) + np.arange(target)[None, :]
                    ref = np.where(idx < len(x), x[np.minimum(idx, len(x) - 1)], -7)
                    assert np.array_equal(got[0], ref) and got[1].all()


@pytest.mark.gpu
def test_gpu_combine_segments(ops, oracle_mod):
    for c in GOLDEN["combine_segments"]:
        ins = []
        for s in c["segments"]:
            ins += [np.array(s[k], np.int32) for k in ("begins", "ends", "data")]
        ins.append(np.arange(len(c["segments"]), dtype=np.int32))
        out = ops.CombineSegments().evaluate(ins)
        assert (out[0].tolist(), out[1].tolist(), out[2].tolist()) == (c["expected"]["begins"], c["expected"]["ends"], c["expected"]["data"])
    rng = np.random.default_rng(12)
    for rows in (1, 2, 513, 20000):
        tok = _random_ragged(rng, rows, 70)
        tok2 = _random_ragged(rng, rows, 9)
        bos = (np.array([0], np.int32), np.array([1], np.int32), np.array([101], np.int32))        # broadcast constant segments
        sep = (np.array([1], np.int32), np.array([3], np.int32), np.array([0, 102, 103], np.int32))
        segs, ids = [bos, tok, sep, tok2, sep], np.array([0, 0, 0, 1, 1], np.int32)
        exp = oracle_mod.combine_segments(segs, ids)
        got = ops.CombineSegments().evaluate([a for s in segs for a in s] + [ids])
        assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1])
        assert np.array_equal(got[2], exp[2]) and np.array_equal(got[5], exp[3])
        assert np.array_equal(got[3], exp[0]) and np.array_equal(got[4], exp[1])
    with pytest.raises(Exception):   # 3 rows against 2 rows: neither equal nor broadcastable
        ops.CombineSegments().evaluate([np.zeros(3, np.int32), np.zeros(3, np.int32), np.zeros(1, np.int32),
                                        np.zeros(2, np.int32), np.zeros(2, np.int32), np.zeros(1, np.int32), np.arange(2, dtype=np.int32)])


@pytest.mark.gpu
def test_gpu_truncate(ops, oracle_mod):
    u8 = lambda s: np.frombuffer(s.encode(), np.uint8)
    for pairs, ml, side, mode, expected in TRUNCATE_CASES:
        ins = []
        for b, e in pairs:
            ins += [np.array(b, np.int32), np.array(e, np.int32), np.zeros(1, np.int32)]
        out = ops.Truncate(len(pairs)).evaluate(ins + [np.int32(ml), u8(side), u8(mode)])
        assert [(out[3 * k].tolist(), out[3 * k + 1].tolist()) for k in range(len(pairs))] == [(list(b), list(e)) for b, e in expected]
    rng = np.random.default_rng(13)
    for rows in (1, 1000, 70000):
        a = _random_ragged(rng, rows, 40)
        b = _random_ragged(rng, rows, 40)
        for ml in (0, 1, 7, 16, 33, 100):
            for side in ("left", "right"):
                exp = oracle_mod.truncate([a[:2]], ml, side)
                got = ops.Truncate(1).evaluate([*a, np.int32(ml), u8(side), u8("longest_first")])
                assert np.array_equal(got[0], exp[0][0]) and np.array_equal(got[1], exp[0][1])
                for mode in ("only_first", "only_second", "longest_first"):
                    exp = oracle_mod.truncate([a[:2], b[:2]], ml, side, mode)
                    got = ops.Truncate(2).evaluate([*a, *b, np.int32(ml), u8(side), u8(mode)])
                    for k in range(2):
                        assert np.array_equal(got[3 * k], exp[k][0]) and np.array_equal(got[3 * k + 1], exp[k][1]), (ml, side, mode)
    with pytest.raises(Exception):
        ops.Truncate(1).evaluate([*a, np.int32(4), u8("middle"), u8("longest_first")])


@pytest.mark.gpu
def test_gpu_fused_tail_equals_the_three_ops(ops, oracle_mod):
    """b200tok_post_dense_run == Truncate -> CombineSegments(prefix, tokens, suffix) -> RaggedToDense run one after the other
    (oracle chain), for both truncation and padding sides."""
    rng = np.random.default_rng(14)
    for rows, max_len, ml, target in [(1, 5, 3, 8), (257, 90, 30, 32), (5000, 700, 510, 512), (64, 10, 100, 16)]:
        b, e, x = _random_ragged(rng, rows, max_len)
        for prefix, suffix in [((), ()), ((101,), (102,)), ((1, 2, 3), (4, 5))]:
            for tleft in (False, True):
                for pad_right in (True, False):
                    tb, te = oracle_mod.truncate([(b, e)], ml, "left" if tleft else "right")[0]
                    segs, ids = [], []
                    if prefix:
                        segs.append((np.array([0], np.int32), np.array([len(prefix)], np.int32), np.array(prefix, np.int32)))
                    segs.append((tb, te, x))
                    if suffix:
                        segs.append((np.array([0], np.int32), np.array([len(suffix)], np.int32), np.array(suffix, np.int32)))
                    if rows == 1:   # a one-row token segment is not a broadcast: give the constants one row as well (they have)
                        pass
                    cb, ce, cx, _ = oracle_mod.combine_segments(segs, np.arange(len(segs)))
                    exp, emask = oracle_mod.ragged_to_dense(cb, ce, cx, target, 0, pad_right)
                    got, gmask = ops.post_dense(b, e, x, ml, target, 0, prefix, suffix, tleft, pad_right)
                    assert np.array_equal(got, exp) and np.array_equal(gmask, emask.astype(bool)), (rows, prefix, tleft, pad_right)


@pytest.mark.gpu
def test_gpu_capacity_and_argument_errors(ops):
    """Errors come back as codes + message (rethrown by the ov::Op shim as ov::Exception); nothing is written past a too-small buffer."""
    import ctypes as C
    from openvino_tokenizers_b200 import _capi as K
    b, e, x = np.array([0, 3], np.int32), np.array([3, 8], np.int32), np.arange(8, dtype=np.int32)
    segs = (K.RaggedI32 * 1)(K.RaggedI32(b.ctypes.data, e.ctypes.data, 2, x.ctypes.data, 8))
    ids = np.zeros(1, np.int32)
    ob, oe, ox, oi = np.empty(2, np.int32), np.empty(2, np.int32), np.full(4, -9, np.int32), np.full(4, -9, np.int32)
    n = C.c_int64(0)
    rc = K.lib().b200tok_combine_segments_run(0, segs, 1, C.c_void_p(ids.ctypes.data), C.c_void_p(ob.ctypes.data), C.c_void_p(oe.ctypes.data),
                                              C.c_void_p(ox.ctypes.data), C.c_void_p(oi.ctypes.data), C.c_int64(4), C.byref(n), K.MEM_HOST, None)
    assert rc == K.E_CAPACITY and n.value == 8 and np.all(ox == -9)
    assert b"capacity" in K.lib().b200tok_last_error()
    with pytest.raises(K.B200TokError):
        ops.post_dense(b, e, x, 4, 8, 0, prefix=list(range(9)))          # more than 8 prefix ids
    out1 = np.empty(1, np.int32)
    rc = K.lib().b200tok_ragged_to_dense_run(0, C.c_void_p(b.ctypes.data), C.c_void_p(e.ctypes.data), C.c_int64(2), C.c_void_p(x.ctypes.data),
                                             C.c_int64(8), C.c_int32(-1), C.c_int32(0), 1, 0, C.c_void_p(out1.ctypes.data), None, K.MEM_HOST, None)
    assert rc == K.E_INVALID                                              # negative target dimension
