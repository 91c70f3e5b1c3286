"""Loader for tests/harness/libharness.so — the product's algorithmic core (csrc/tok_core.cuh +
csrc/tables.cpp) compiled for the host so that it can be compared with the oracle without a GPU."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from openvino_tokenizers_b200 import _capi as K

HERE = Path(__file__).resolve().parent / "harness"
CSRC = HERE.parent.parent / "openvino_tokenizers_b200" / "csrc"
_lib = None


def lib():
    global _lib
    if _lib is None:
        so = HERE / "libharness.so"
        srcs = [HERE / "harness.cpp", CSRC / "tables.cpp", CSRC / "regex_compile.cpp", CSRC / "regex_vm.cuh", CSRC / "tables.hpp", CSRC / "tok_core.cuh"]
        if not so.exists() or so.stat().st_mtime < max(s.stat().st_mtime for s in srcs):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-o", str(so),
                                   str(HERE / "harness.cpp"), str(CSRC / "tables.cpp"), str(CSRC / "regex_compile.cpp")])
        _lib = C.CDLL(str(so))
        _lib.hz_split.restype = C.c_int64
        _lib.hz_bpe_create.restype = C.c_void_p
        _lib.hz_wp_create.restype = C.c_void_p
        _lib.hz_bpe_piece.restype = C.c_int64
        _lib.hz_wp_word.restype = C.c_int64
        _lib.hz_bpe_info.restype = C.c_int64
    return _lib


def split(pattern, behaviour, invert, max_splits, data: bytes):
    pat = pattern.encode() if isinstance(pattern, str) else pattern
    arr = np.frombuffer(data, np.uint8) if len(data) else np.zeros(1, np.uint8)
    cap = len(data) + 2
    ob, oe = np.empty(cap, np.int32), np.empty(cap, np.int32)
    n = lib().hz_split(pat, C.c_int64(len(pat)), behaviour.encode(), int(invert), int(max_splits),
                       arr.ctypes.data_as(K.u8p), C.c_int64(len(data)), ob.ctypes.data_as(K.i32p),
                       oe.ctypes.data_as(K.i32p), C.c_int64(cap))
    if n < 0:
        raise ValueError(f"hz_split error {n}")
    return list(zip(ob[:n].tolist(), oe[:n].tolist()))


class HostBpe:
    def __init__(self, vocab, ml, mr=None, added=None, added_ids=None, unk_token=b"", end_suffix=b"",
                 byte_fallback=False, fuse_unk=False):
        self._keep = []
        d = K.BpeDesc()
        d.vocab = K.make_strings(vocab, self._keep)
        d.merges_left = K.make_strings(ml, self._keep)
        d.merges_right = K.make_strings(mr, self._keep)
        d.added_tokens = K.make_strings(added, self._keep)
        if added_ids is not None:
            aid = np.ascontiguousarray(added_ids, np.int32)
            self._keep.append(aid)
            d.added_ids = aid.ctypes.data_as(K.i32p)
        d.unk_token, d.unk_token_len = unk_token, len(unk_token)
        d.end_suffix, d.end_suffix_len = end_suffix, len(end_suffix)
        d.suffix_indicator, d.suffix_indicator_len = b"", 0
        d.byte_fallback, d.fuse_unk, d.cache_capacity, d.device = int(byte_fallback), int(fuse_unk), 0, 0
        self.h = C.c_void_p(lib().hz_bpe_create(C.byref(d)))
        if not self.h:
            raise ValueError("hz_bpe_create failed")

    def info(self, what):
        return lib().hz_bpe_info(self.h, what)

    def piece(self, data: bytes, mode=0):
        arr = np.frombuffer(data, np.uint8) if len(data) else np.zeros(1, np.uint8)
        out = np.empty(len(data) + 64, np.int32)
        n = lib().hz_bpe_piece(self.h, arr.ctypes.data_as(K.u8p), C.c_int64(len(data)), mode, out.ctypes.data_as(K.i32p))
        return out[:n].tolist()


    def piece_packed_with_tie(self, data: bytes):
        """The packed-key merge loop (left pair first on ties) and its tie report."""
        arr = np.frombuffer(data, np.uint8) if len(data) else np.zeros(1, np.uint8)
        out = np.empty(len(data) + 64, np.int32)
        tie = C.c_int32(0)
        L = lib()
        L.hz_bpe_piece_tie.restype = C.c_int64
        n = L.hz_bpe_piece_tie(self.h, arr.ctypes.data_as(K.u8p), C.c_int64(len(data)), out.ctypes.data_as(K.i32p), C.byref(tie))
        return out[:n].tolist(), bool(tie.value)


class HostWordpiece:
    def __init__(self, vocab, suffix=b"##", max_bytes=100):
        self._keep = []
        d = K.WordpieceDesc()
        d.vocab = K.make_strings(vocab, self._keep)
        d.suffix_indicator, d.suffix_indicator_len = suffix, len(suffix)
        d.max_bytes_per_word, d.device = max_bytes, 0
        self.h = C.c_void_p(lib().hz_wp_create(C.byref(d)))

    def word(self, data: bytes, unk: int):
        arr = np.frombuffer(data, np.uint8) if len(data) else np.zeros(1, np.uint8)
        out = np.empty(len(data) + 4, np.int32)
        n = lib().hz_wp_word(self.h, arr.ctypes.data_as(K.u8p), C.c_int64(len(data)), C.c_int32(unk), out.ctypes.data_as(K.i32p))
        return out[:n].tolist()


def gpt2_closed_form(data: bytes, single_digits=False):
    """Piece (begin, end) list from the closed-form GPT-2 predicate."""
    if not data:
        return []
    arr = np.frombuffer(data, np.uint8)
    out = np.empty(len(data) + 1, np.int32)
    lib().hz_gpt2_closed_form.restype = C.c_int64
    n = lib().hz_gpt2_closed_form(arr.ctypes.data_as(K.u8p), C.c_int64(len(data)), int(single_digits), out.ctypes.data_as(K.i32p))
    b = out[:n].tolist()
    return list(zip(b, b[1:] + [len(data)]))


def gpt2_neighbour_form(data: bytes, single_digits=False):
    """Piece (begin, end) list from the branch-free neighbour form (ASCII subjects only)."""
    if not data:
        return []
    arr = np.frombuffer(data, np.uint8)
    out = np.empty(len(data) + 1, np.int32)
    lib().hz_gpt2_neighbour_form.restype = C.c_int64
    n = lib().hz_gpt2_neighbour_form(arr.ctypes.data_as(K.u8p), C.c_int64(len(data)), int(single_digits), out.ctypes.data_as(K.i32p))
    b = out[:n].tolist()
    return list(zip(b, b[1:] + [len(data)]))


def gpt2_word_form(data: bytes, single_digits=False):
    """Piece (begin, end) list from the word (bit-mask) form of the predicate; any valid UTF-8 subject."""
    if not data:
        return []
    arr = np.frombuffer(data, np.uint8)
    out = np.empty(len(data) + 1, np.int32)
    lib().hz_gpt2_word_form.restype = C.c_int64
    n = lib().hz_gpt2_word_form(arr.ctypes.data_as(K.u8p), C.c_int64(len(data)), int(single_digits), out.ctypes.data_as(K.i32p))
    b = out[:n].tolist()
    return list(zip(b, b[1:] + [len(data)]))


def llama3_word_form(data: bytes):
    """Piece (begin, end) list from the word (bit-mask) form of the Llama-3 predicate, or None if the subject holds a non-ASCII
    digit (such rows are handed to the generic kernel)."""
    if not data:
        return []
    arr = np.frombuffer(data, np.uint8)
    out = np.empty(len(data) + 1, np.int32)
    lib().hz_llama3_word_form.restype = C.c_int64
    n = lib().hz_llama3_word_form(arr.ctypes.data_as(K.u8p), C.c_int64(len(data)), out.ctypes.data_as(K.i32p))
    if n == -2:
        return None
    b = out[:n].tolist()
    return list(zip(b, b[1:] + [len(data)]))


def special_split(pattern: str, data: bytes):
    """SpecialTokensSplit of one element by the product's host-compiled parser + matcher: [(begin, end, skip)]."""
    pat = pattern.encode()
    arr = np.frombuffer(data, np.uint8) if len(data) else np.zeros(1, np.uint8)
    cap = len(data) + 2
    ob, oe, osk = np.empty(cap, np.int32), np.empty(cap, np.int32), np.empty(cap, np.uint8)
    lib().hz_special_split.restype = C.c_int64
    n = lib().hz_special_split(pat, C.c_int64(len(pat)), arr.ctypes.data_as(K.u8p), C.c_int64(len(data)), ob.ctypes.data_as(K.i32p),
                               oe.ctypes.data_as(K.i32p), osk.ctypes.data_as(K.u8p), C.c_int64(cap))
    if n < 0:
        raise ValueError(f"hz_special_split error {n}")
    return list(zip(ob[:n].tolist(), oe[:n].tolist(), osk[:n].tolist()))


def hz_normalize(kind, a, b, flag, begins, ends, chars, skips=None, expand=24):
    """Normalise strings with the product's host-compiled parser + scan.  kind 0: RegexNormalization(a=search, b=replace,
    flag=global_replace); kind 1: CharsMapNormalization(a=blob).  Returns (begins, ends, chars); raises ValueError(code)."""
    begins, ends = np.ascontiguousarray(begins, np.int32), np.ascontiguousarray(ends, np.int32)
    chars = np.ascontiguousarray(chars, np.uint8) if len(chars) else np.zeros(1, np.uint8)
    sk = None if skips is None else np.ascontiguousarray(skips, np.uint8)
    n = len(begins)
    cap = expand * int(chars.size) + 64 * n + 64
    ob, oe, oc = np.empty(max(n, 1), np.int32), np.empty(max(n, 1), np.int32), np.zeros(cap, np.uint8)
    a, b = bytes(a), bytes(b)
    lib().hz_normalize.restype = C.c_int64
    t = lib().hz_normalize(int(kind), a, C.c_int64(len(a)), b, C.c_int64(len(b)), int(flag), begins.ctypes.data_as(K.i32p),
                           ends.ctypes.data_as(K.i32p), chars.ctypes.data_as(K.u8p), None if sk is None else sk.ctypes.data_as(K.u8p),
                           C.c_int64(n), ob.ctypes.data_as(K.i32p), oe.ctypes.data_as(K.i32p), oc.ctypes.data_as(K.u8p), C.c_int64(cap))
    if t < 0:
        raise ValueError(int(t))
    return ob[:n], oe[:n], oc[:t].copy()


def hz_charsmap_ascii_table_check(blob):
    """Violations of the ASCII shortcut tables of a charsmap against the general step (0 = consistent)."""
    blob = bytes(blob)
    lib().hz_charsmap_ascii_table_check.restype = C.c_int64
    return int(lib().hz_charsmap_ascii_table_check(blob, C.c_int64(len(blob))))


def hz_chain_table(steps):
    """Composed ASCII byte table of a chain: steps = [(kind, a, b, flag)] as for hz_normalize.  Returns a 128-entry uint8
    array (0xFE dropped, 0xFF general) or None when the chain is not composable."""
    lib().hz_chain_reset()
    for kind, a, b, flag in steps:
        a, b = bytes(a), bytes(b)
        rc = lib().hz_chain_add(int(kind), a, C.c_int64(len(a)), b, C.c_int64(len(b)), int(flag))
        if rc:
            raise ValueError(rc)
    T = np.zeros(128, np.uint8)
    rc = lib().hz_chain_table(T.ctypes.data_as(K.u8p))
    if rc < 0:
        raise ValueError(rc)
    return T if rc == 1 else None
