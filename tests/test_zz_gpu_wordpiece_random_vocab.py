"""GPU tier, collected last: the table-driven WordPiece window pass (kernels.cuh wordpiece_window_pieces: val1, the two-byte jump
table, the work queue) on RANDOM vocabularies — missing root children, missing one- / two-byte tokens, non-ASCII bytes that must take
the root lookup — against the oracle (reference src/wordpiece_tokenizer.cpp:96-130).  The host form of the same walk is checked in the
CPU tier (tests/test_core_host.py::test_wordpiece_jump_tables_on_random_vocabularies)."""
import itertools

import numpy as np
import pytest

import cases
from openvino_tokenizers_b200.strings import pack_strings

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_wordpiece_random_vocabulary(seed):
    import oracle
    from openvino_tokenizers_b200 import ops
    rng = np.random.default_rng(seed)
    alpha = [b"a", b"b", b"c", b"\xc3", b"\xa9", b"1"]
    pool = [b"".join(t) for n in range(1, 5) for t in itertools.product(alpha, repeat=n)]
    root = [t for t, k in zip(pool, rng.random(len(pool)) < [0.9, 0.5, 0.25, 0.1][seed]) if k]
    sub = [b"##" + t for t, k in zip(pool, rng.random(len(pool)) < 0.35) if k]
    vocab = [b"[UNK]"] + root + sub
    rng.shuffle(vocab)
    unk = vocab.index(b"[UNK]")
    v = pack_strings(vocab)
    words = pool + [b"".join(rng.choice(alpha, size=int(rng.integers(5, 14)))) for _ in range(4000)] + [b"a" * 100, b"a" * 101, b"ab" * 300]
    rng.shuffle(words)
    b, e, c = pack_strings(words)
    n = len(words)
    rb = np.arange(0, n, 7, dtype=np.int32)                 # rows of seven words (the last one shorter)
    re_ = np.minimum(rb + 7, n).astype(np.int32)
    exp = oracle.WordpieceOracle(v, b"##", 100)(rb, re_, b, e, c, unk)
    got = ops.WordpieceTokenizer(b"##", 100).with_constants(v).evaluate([rb, re_, b, e, c, *v, np.array(unk, np.int32)])
    assert cases.ragged_rows_equal(got, exp)
