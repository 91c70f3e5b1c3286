"""ctypes mirror of include/b200tok.h and the loader of libb200tok.so.

There is deliberately no fallback: if the CUDA library is missing or cannot be loaded, importing the
ops raises — the product path never computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

LIB_PATH = Path(__file__).resolve().parent / "csrc" / "libb200tok.so"

MEM_HOST, MEM_DEVICE = 0, 1
E_INVALID, E_CUDA, E_CAPACITY, E_UNSUPPORTED, E_VOCAB = -1, -2, -3, -4, -5

i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)
u8p = C.POINTER(C.c_uint8)


class Strings(C.Structure):
    _fields_ = [("begins", i32p), ("ends", i32p), ("chars", u8p), ("n", C.c_int64), ("n_chars", C.c_int64)]


class RaggedStrings(C.Structure):
    _fields_ = [("ragged_begins", C.c_void_p), ("ragged_ends", C.c_void_p), ("n_rows", C.c_int64),
                ("begins", C.c_void_p), ("ends", C.c_void_p), ("n_elems", C.c_int64),
                ("chars", C.c_void_p), ("n_chars", C.c_int64), ("skips", C.c_void_p), ("mem", C.c_int)]


class RaggedStringsOut(C.Structure):
    _fields_ = [("ragged_begins", C.c_void_p), ("ragged_ends", C.c_void_p), ("begins", C.c_void_p),
                ("ends", C.c_void_p), ("skips", C.c_void_p), ("capacity", C.c_int64), ("n_elems", C.c_int64),
                ("n_rows", C.c_int64), ("mem", C.c_int)]


class RaggedIds(C.Structure):
    _fields_ = [("begins", C.c_void_p), ("ends", C.c_void_p), ("ids", C.c_void_p), ("capacity", C.c_int64),
                ("n_ids", C.c_int64), ("n_ids_device", C.c_void_p), ("mem", C.c_int)]


class RegexSplitDesc(C.Structure):
    _fields_ = [("pattern", C.c_char_p), ("pattern_len", C.c_int64), ("behaviour", C.c_char_p), ("invert", C.c_int),
                ("max_splits", C.c_int), ("device", C.c_int)]


class BpeDesc(C.Structure):
    _fields_ = [("vocab", Strings), ("merges_left", Strings), ("merges_right", Strings), ("added_tokens", Strings),
                ("added_ids", i32p),
                ("unk_token", C.c_char_p), ("unk_token_len", C.c_int64),
                ("suffix_indicator", C.c_char_p), ("suffix_indicator_len", C.c_int64),
                ("end_suffix", C.c_char_p), ("end_suffix_len", C.c_int64),
                ("fuse_unk", C.c_int), ("byte_fallback", C.c_int), ("cache_capacity", C.c_int64), ("device", C.c_int)]


class WordpieceDesc(C.Structure):
    _fields_ = [("vocab", Strings), ("suffix_indicator", C.c_char_p), ("suffix_indicator_len", C.c_int64),
                ("max_bytes_per_word", C.c_int), ("device", C.c_int)]


class VocabEncDesc(C.Structure):
    _fields_ = [("keys", Strings), ("values", C.c_void_p), ("values_are_i64", C.c_int), ("device", C.c_int)]


class VocabDecDesc(C.Structure):
    _fields_ = [("vocab", Strings), ("device", C.c_int)]


class Decoded(C.Structure):
    _fields_ = [("ragged_begins", C.c_void_p), ("ragged_ends", C.c_void_p), ("begins", C.c_void_p),
                ("ends", C.c_void_p), ("chars", C.c_void_p), ("chars_capacity", C.c_int64), ("n_chars", C.c_int64),
                ("mem", C.c_int)]


class RaggedI32(C.Structure):
    _fields_ = [("begins", C.c_void_p), ("ends", C.c_void_p), ("n", C.c_int64), ("elems", C.c_void_p), ("n_elems", C.c_int64)]


class PostDesc(C.Structure):
    _fields_ = [("max_length", C.c_int32), ("truncate_left", C.c_int), ("prefix", i32p), ("n_prefix", C.c_int32),
                ("suffix", i32p), ("n_suffix", C.c_int32), ("target_dim", C.c_int32), ("pad_value", C.c_int32),
                ("pad_right", C.c_int)]


class PeerOut(C.Structure):
    _fields_ = [("world", C.c_int), ("rank", C.c_int), ("ids", C.c_void_p * 8), ("begins", C.c_void_p * 8), ("ends", C.c_void_p * 8),
                ("slot_capacity", C.c_int64), ("rows_per_rank", C.c_int64), ("wire16", C.c_int), ("ids16", C.c_void_p * 8),
                ("ids_mc", C.c_void_p), ("begins_mc", C.c_void_p), ("ends_mc", C.c_void_p)]


class PeerPull(C.Structure):
    _fields_ = [("world", C.c_int), ("rank", C.c_int), ("wire16", C.c_int), ("skip_self_ids", C.c_int),
                ("src_ids16", C.c_void_p * 8), ("src_ids", C.c_void_p * 8), ("src_begins", C.c_void_p * 8), ("src_ends", C.c_void_p * 8),
                ("src_total", C.c_void_p * 8), ("ids", C.c_void_p), ("begins", C.c_void_p), ("ends", C.c_void_p),
                ("slot_capacity", C.c_int64), ("rows_per_rank", C.c_int64)]


def make_strings(triple, keep: list) -> Strings:
    """(begins, ends, chars) numpy triple -> Strings struct; arrays are appended to `keep` to stay alive."""
    if triple is None:
        return Strings(None, None, None, 0, 0)
    b = np.ascontiguousarray(triple[0], dtype=np.int32)
    e = np.ascontiguousarray(triple[1], dtype=np.int32)
    c = triple[2]
    if isinstance(c, (bytes, bytearray)):
        c = np.frombuffer(bytes(c), dtype=np.uint8)
    c = np.ascontiguousarray(c, dtype=np.uint8)
    if c.size == 0:
        c = np.zeros(1, np.uint8)
        n_chars = 0
    else:
        n_chars = c.size
    keep.extend([b, e, c])
    return Strings(b.ctypes.data_as(i32p), e.ctypes.data_as(i32p), c.ctypes.data_as(u8p), len(b), n_chars)


_lib = None


class B200TokError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"b200tok error {code}: {msg}")
        self.code = code


def lib():
    """Load libb200tok.so (built by __graft_entry__.build()).  Raises if it is missing — no fallback."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(f"{LIB_PATH} not found: build the CUDA extension first "
                              "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
        L = C.CDLL(str(LIB_PATH))
        L.b200tok_last_error.restype = C.c_char_p
        L.b200tok_launch_count.restype = C.c_int64
        L.b200tok_launch_count.argtypes = [C.c_void_p]
        L.b200tok_destroy.argtypes = [C.c_void_p]
        L.b200tok_set_timing.argtypes = [C.c_void_p, C.c_int]
        L.b200tok_last_kernel_ms.argtypes = [C.c_void_p]
        L.b200tok_last_kernel_ms.restype = C.c_float
        L.b200tok_vocabdec_max_chars.restype = C.c_int64
        L.b200tok_peer_pack_run.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.b200tok_peer_pull_run.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.b200tok_vocabdec_max_chars.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise B200TokError(rc, (lib().b200tok_last_error() or b"").decode("utf-8", "replace"))


EXPORTED_SYMBOLS = [
    "b200tok_version", "b200tok_last_error", "b200tok_device_count", "b200tok_destroy", "b200tok_launch_count",
    "b200tok_set_timing", "b200tok_last_kernel_ms",
    "b200tok_regexsplit_create", "b200tok_regexsplit_run", "b200tok_regexsplit_set_skip_tokens", "b200tok_specialsplit_create", "b200tok_specialsplit_run",
    "b200tok_bpe_create", "b200tok_bpe_run", "b200tok_split_bpe_run", "b200tok_split_bpe_run_sharded", "b200tok_peer_expand_run",
    "b200tok_peer_pack_run", "b200tok_peer_pull_run",
    "b200tok_wordpiece_create", "b200tok_wordpiece_run", "b200tok_split_wordpiece_run", "b200tok_split_wordpiece_run_sharded",
    "b200tok_vocabenc_create", "b200tok_vocabenc_run",
    "b200tok_vocabdec_create", "b200tok_vocabdec_run", "b200tok_vocabdec_max_chars",
    "b200tok_bytefallback_run",
    "b200tok_bytes_to_chars_run", "b200tok_chars_to_bytes_run", "b200tok_fuze_ragged_run", "b200tok_utf8_validate_run",
    "b200tok_truncate_run", "b200tok_combine_segments_run", "b200tok_ragged_to_dense_run", "b200tok_post_dense_run",
    "b200tok_regexnorm_create", "b200tok_charsmap_create", "b200tok_normalize_run", "b200tok_normalize_chain_run",
]
