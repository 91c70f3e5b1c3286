"""Decomposed string tensors — the (begins:i32, ends:i32, chars:u8) layout every reference op uses
(reference src/utils.cpp:84-102, python/openvino_tokenizers/utils.py:436-458)."""
from __future__ import annotations

import numpy as np


def pack_strings(strings):
    """list[bytes|str] -> (begins, ends, chars) exactly like ``create_unpacked_string``."""
    bs = [s.encode("utf-8") if isinstance(s, str) else bytes(s) for s in strings]
    lens = np.fromiter((len(b) for b in bs), dtype=np.int64, count=len(bs))
    ends = np.cumsum(lens, dtype=np.int64)
    begins = ends - lens
    chars = np.frombuffer(b"".join(bs), dtype=np.uint8).copy()
    return begins.astype(np.int32), ends.astype(np.int32), chars


def unpack_strings(begins, ends, chars):
    """(begins, ends, chars) -> list[bytes]."""
    buf = np.asarray(chars, dtype=np.uint8).tobytes()
    return [buf[b:e] for b, e in zip(np.asarray(begins).tolist(), np.asarray(ends).tolist())]


def add_ragged_dimension(begins, ends):
    """One element per row, as ``TokenizerPipeline.add_ragged_dimension`` does with Range ops
    (reference python/openvino_tokenizers/tokenizer_pipeline.py:1668-1676)."""
    n = len(begins)
    return np.arange(0, n, dtype=np.int32), np.arange(1, n + 1, dtype=np.int32)


def ragged_rows(rb, re_, begins, ends, chars):
    """Ragged string tensor -> list[list[bytes]]."""
    flat = unpack_strings(begins, ends, chars)
    return [flat[a:b] for a, b in zip(np.asarray(rb).tolist(), np.asarray(re_).tolist())]
