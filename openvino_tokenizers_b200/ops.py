"""Host-side mirror of the reference's hot-path operators, over the C ABI (include/b200tok.h).

Each class has the reference op's name, attributes and *input list* (the same tensors, in the same
order, that `ov::Op::evaluate(outputs, inputs)` receives), and `evaluate(inputs)` returns the op's
output list — so the parity tests read like the reference's own op tests.

  RegexSplit          src/regex_split.{hpp,cpp}        6 / 7 inputs   -> 5 / 6 outputs
  BPETokenizer        src/bpe_tokenizer.{hpp,cpp}      11/14/15/18    -> 3 outputs
  WordpieceTokenizer  src/wordpiece_tokenizer.{hpp,cpp} 9 inputs      -> 3 outputs
  VocabEncoder        src/vocab_encoder.{hpp,cpp}      8 inputs       -> 1 output
  VocabDecoder        src/vocab_decoder.{hpp,cpp}      4 / 5 inputs   -> 5 outputs
  ByteFallback        src/byte_fallback.{hpp,cpp}      3 inputs       -> 3 outputs
and the ops either side of that path: SpecialTokensSplit, Truncate, CombineSegments, RaggedToDense,
BytesToChars, CharsToBytes, FuzeRagged, UTF8Validate, RegexNormalization (single-character patterns), CharsMapNormalization.

All compute happens in libb200tok.so on the GPU; this module only marshals pointers.  Tables are
built on the first `evaluate` from the Constant inputs (as the reference does under call_once).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as K

__all__ = ["RegexSplit", "BPETokenizer", "WordpieceTokenizer", "VocabEncoder", "VocabDecoder", "ByteFallback",
           "SpecialTokensSplit", "BytesToChars", "CharsToBytes", "FuzeRagged", "UTF8Validate", "Truncate", "CombineSegments",
           "RaggedToDense", "RegexNormalization", "CharsMapNormalization", "post_dense", "normalize_chain",
           "split_bpe", "split_wordpiece", "B200TokError"]

B200TokError = K.B200TokError


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _u8(a):
    if isinstance(a, (bytes, bytearray)):
        a = np.frombuffer(bytes(a), dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8)


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _as_text(a) -> str:
    """A u8 string-scalar input (what string attributes passed as tensors look like) or a Python str/bytes."""
    if isinstance(a, str):
        return a
    if isinstance(a, (bytes, bytearray)):
        return bytes(a).decode()
    return bytes(np.asarray(a, dtype=np.uint8).reshape(-1)).decode()


class _Handle:
    def __init__(self):
        self._h = C.c_void_p()

    @property
    def handle(self):
        return self._h

    @property
    def launches(self) -> int:
        return int(K.lib().b200tok_launch_count(self._h)) if self._h else 0

    def close(self):
        if self._h:
            K.lib().b200tok_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _ragged_in(rb, re_, begins, ends, chars, skips=None, keep=None):
    rb, re_, begins, ends, chars = _i32(rb), _i32(re_), _i32(begins), _i32(ends), _u8(chars)
    sk = None if skips is None else _u8(np.asarray(skips, dtype=np.uint8))
    keep.extend([rb, re_, begins, ends, chars, sk])
    if len(begins) != len(ends) or len(rb) != len(re_):
        raise ValueError("ragged string tensor: begins/ends length mismatch")
    r = K.RaggedStrings(_ptr(rb), _ptr(re_), len(rb), _ptr(begins), _ptr(ends), len(begins), _ptr(chars), chars.size,
                        _ptr(sk) if sk is not None else None, K.MEM_HOST)
    return r


def _ids_out(n_rows, capacity):
    ob, oe = np.empty(n_rows, np.int32), np.empty(n_rows, np.int32)
    ids = np.empty(max(capacity, 1), np.int32)
    out = K.RaggedIds(_ptr(ob), _ptr(oe), _ptr(ids), capacity, 0, None, K.MEM_HOST)
    return out, ob, oe, ids


class RegexSplit(_Handle):
    """RegexSplit(behaviour, invert, max_splits); inputs: ragged strings [0..4], optional skips [5], pattern; legacy form:
    ragged strings [0..4], pattern [5], skip-token strings [6..8] (src/regex_split.cpp:98-113)."""

    def __init__(self, behaviour="remove", invert=False, max_splits=-1, device=0):
        super().__init__()
        self.behaviour, self.invert, self.max_splits, self.device = behaviour, bool(invert), int(max_splits), device
        self._pattern = None

    def _ensure(self, pattern: bytes):
        if self._h:
            return
        self._pattern = bytes(pattern)
        d = K.RegexSplitDesc(self._pattern, len(self._pattern), self.behaviour.lower().encode(), int(self.invert),
                             self.max_splits, self.device)
        K.check(K.lib().b200tok_regexsplit_create(C.byref(d), C.byref(self._h)))

    def with_pattern(self, pattern):
        self._ensure(pattern.encode() if isinstance(pattern, str) else pattern)
        return self

    def with_skip_tokens(self, tokens):
        """Legacy 9-input form: `tokens` = (begins, ends, chars) of the skip-token strings (inputs [6..8])."""
        keep = []
        st = K.make_strings(tokens, keep)
        K.check(K.lib().b200tok_regexsplit_set_skip_tokens(self._h, C.byref(st)))
        self._skip_tokens_set = True
        return self

    def evaluate(self, inputs):
        if len(inputs) not in (6, 7, 9):
            raise ValueError("Incorrect number of inputs passed to RegexSplit: %d; try to reconvert tokenizer with newer version of "
                             "OpenVINO Tokenizers" % len(inputs))
        has_skips = len(inputs) == 7
        self._ensure(_u8(inputs[5 + has_skips]).tobytes())
        if len(inputs) == 9 and not getattr(self, "_skip_tokens_set", False) and len(inputs[6]) > 0:      # src/regex_split.cpp:166
            self.with_skip_tokens((inputs[6], inputs[7], inputs[8]))
        keep = []
        rin = _ragged_in(*inputs[:5], skips=inputs[5] if has_skips else None, keep=keep)
        n_rows, cap = rin.n_rows, rin.n_chars + rin.n_elems
        orb, ore = np.empty(max(n_rows, 1), np.int32), np.empty(max(n_rows, 1), np.int32)
        ob, oe = np.empty(max(cap, 1), np.int32), np.empty(max(cap, 1), np.int32)
        osk = np.empty(max(cap, 1), np.uint8) if has_skips else None
        out = K.RaggedStringsOut(_ptr(orb), _ptr(ore), _ptr(ob), _ptr(oe), _ptr(osk) if has_skips else None, cap, 0, 0,
                                 K.MEM_HOST)
        K.check(K.lib().b200tok_regexsplit_run(self._h, C.byref(rin), C.byref(out), None))
        P, R = out.n_elems, out.n_rows
        res = [orb[:R].copy(), ore[:R].copy(), ob[:P].copy(), oe[:P].copy(), keep[4]]
        if has_skips:
            res.append(osk[:P].astype(bool))
        return res


class SpecialTokensSplit(_Handle):
    """SpecialTokensSplit; inputs: ragged strings [0..4], optional skips [5], split pattern (reference
    src/special_tokens_split.cpp:61-162).  Outputs: ragged strings [0..4] and skips [5] (always)."""

    def __init__(self, device=0):
        super().__init__()
        self.device = device

    def with_pattern(self, pattern):
        if not self._h:
            p = pattern.encode() if isinstance(pattern, str) else bytes(pattern)
            K.check(K.lib().b200tok_specialsplit_create(p, C.c_int64(len(p)), self.device, C.byref(self._h)))
        return self

    def evaluate(self, inputs):
        if len(inputs) not in (6, 7):
            raise ValueError("Incorrect number of inputs passed to SpecialTokensSplit: %d; try to reconvert tokenizer with newer "
                             "version of OpenVINO Tokenizers" % len(inputs))
        has_skips = len(inputs) == 7
        self.with_pattern(_u8(inputs[5 + has_skips]).tobytes())
        keep = []
        rin = _ragged_in(*inputs[:5], skips=inputs[5] if has_skips else None, keep=keep)
        n_rows, cap = rin.n_rows, rin.n_chars + rin.n_elems
        orb, ore = np.empty(max(n_rows, 1), np.int32), np.empty(max(n_rows, 1), np.int32)
        ob, oe, osk = np.empty(max(cap, 1), np.int32), np.empty(max(cap, 1), np.int32), np.empty(max(cap, 1), np.uint8)
        out = K.RaggedStringsOut(_ptr(orb), _ptr(ore), _ptr(ob), _ptr(oe), _ptr(osk), cap, 0, 0, K.MEM_HOST)
        K.check(K.lib().b200tok_specialsplit_run(self._h, C.byref(rin), C.byref(out), None))
        P = out.n_elems
        return [orb[:n_rows].copy(), ore[:n_rows].copy(), ob[:P].copy(), oe[:P].copy(), keep[4], osk[:P].astype(bool)]


class BPETokenizer(_Handle):
    """BPETokenizer(unk_token, fuse_unk, suffix_indicator, end_suffix, byte_fallback, cache_capacity)."""

    def __init__(self, unk_token="", fuse_unk=False, suffix_indicator="", end_suffix="", byte_fallback=False,
                 cache_capacity=20000, device=0):
        super().__init__()
        enc = lambda s: s.encode() if isinstance(s, str) else bytes(s)
        self.unk_token, self.suffix_indicator, self.end_suffix = enc(unk_token), enc(suffix_indicator), enc(end_suffix)
        self.fuse_unk, self.byte_fallback, self.cache_capacity, self.device = bool(fuse_unk), bool(byte_fallback), int(cache_capacity), device

    def _ensure(self, consts):
        if self._h:
            return
        n = len(consts) + 5
        if n not in (11, 14, 15, 18):
            raise ValueError("Incorrect number of inputs passed to BPETokenizer, try to reconvert tokenizer with newer "
                             "version of OpenVINO Tokenizers")
        keep = []
        d = K.BpeDesc()
        d.vocab = K.make_strings(consts[0:3], keep)
        d.merges_left = K.make_strings(consts[3:6], keep)
        pairs = n in (14, 18)
        d.merges_right = K.make_strings(consts[6:9], keep) if pairs else K.Strings(None, None, None, 0, 0)
        if n in (15, 18):
            a = consts[-4:]
            d.added_tokens = K.make_strings(a[0:3], keep)
            aid = _i32(a[3])
            keep.append(aid)
            d.added_ids = aid.ctypes.data_as(K.i32p)
        else:
            d.added_tokens = K.Strings(None, None, None, 0, 0)
        d.unk_token, d.unk_token_len = self.unk_token, len(self.unk_token)
        d.suffix_indicator, d.suffix_indicator_len = self.suffix_indicator, len(self.suffix_indicator)
        d.end_suffix, d.end_suffix_len = self.end_suffix, len(self.end_suffix)
        d.fuse_unk, d.byte_fallback, d.cache_capacity, d.device = int(self.fuse_unk), int(self.byte_fallback), self.cache_capacity, self.device
        K.check(K.lib().b200tok_bpe_create(C.byref(d), C.byref(self._h)))

    def with_constants(self, consts):
        self._ensure(list(consts))
        return self

    def evaluate(self, inputs):
        self._ensure(list(inputs[5:]))
        keep = []
        rin = _ragged_in(*inputs[:5], keep=keep)
        out, ob, oe, ids = _ids_out(rin.n_rows, rin.n_chars + rin.n_elems * len(self.end_suffix))
        K.check(K.lib().b200tok_bpe_run(self._h, C.byref(rin), C.byref(out), None))
        return [ob, oe, ids[:out.n_ids].copy()]


class WordpieceTokenizer(_Handle):
    """WordpieceTokenizer(suffix_indicator, max_bytes_per_word); inputs [0..4] words, [5..7] vocab, [8] unk id."""

    def __init__(self, suffix_indicator="##", max_bytes_per_word=100, device=0):
        super().__init__()
        self.suffix_indicator = suffix_indicator.encode() if isinstance(suffix_indicator, str) else bytes(suffix_indicator)
        self.max_bytes_per_word, self.device = int(max_bytes_per_word), device

    def _ensure(self, vocab):
        if self._h:
            return
        keep = []
        d = K.WordpieceDesc(K.make_strings(vocab, keep), self.suffix_indicator, len(self.suffix_indicator),
                            self.max_bytes_per_word, self.device)
        K.check(K.lib().b200tok_wordpiece_create(C.byref(d), C.byref(self._h)))

    def with_constants(self, vocab):
        self._ensure(vocab)
        return self

    def evaluate(self, inputs):
        if len(inputs) != 9:
            raise ValueError("WordpieceTokenizer expects 9 inputs")
        self._ensure(inputs[5:8])
        unk = int(np.asarray(inputs[8]).reshape(-1)[0])
        keep = []
        rin = _ragged_in(*inputs[:5], keep=keep)
        out, ob, oe, ids = _ids_out(rin.n_rows, rin.n_chars + rin.n_elems)
        K.check(K.lib().b200tok_wordpiece_run(self._h, C.byref(rin), C.c_int32(unk), C.byref(out), None))
        return [ob, oe, ids[:out.n_ids].copy()]


class VocabEncoder(_Handle):
    """VocabEncoder; inputs [0..2] strings, [3..5] keys, [6] values (i32|i64), [7] default."""

    def __init__(self, device=0):
        super().__init__()
        self.device = device
        self._dtype = None

    def evaluate(self, inputs):
        if len(inputs) != 8:
            raise ValueError("VocabEncoder expects 8 inputs")
        values = np.asarray(inputs[6])
        if values.dtype not in (np.int32, np.int64):
            raise ValueError("VocabEncoder: unsupported element type: %s" % values.dtype)
        if not self._h:
            keep = []
            vals = np.ascontiguousarray(values)
            d = K.VocabEncDesc(K.make_strings(inputs[3:6], keep), C.c_void_p(vals.ctypes.data),
                               int(vals.dtype == np.int64), self.device)
            K.check(K.lib().b200tok_vocabenc_create(C.byref(d), C.byref(self._h)))
            self._dtype = vals.dtype
        b, e, c = _i32(inputs[0]), _i32(inputs[1]), _u8(inputs[2])
        out = np.empty(len(b), self._dtype)
        default = int(np.asarray(inputs[7]).reshape(-1)[0])
        K.check(K.lib().b200tok_vocabenc_run(self._h, _ptr(b), _ptr(e), C.c_int64(len(b)), _ptr(c), C.c_int64(c.size),
                                             C.c_int64(default), _ptr(out), K.MEM_HOST, None))
        return [out]


class VocabDecoder(_Handle):
    """VocabDecoder(skip_tokens); inputs [0] ids i32[B,S], [1..3] vocab, optional [4] skip tokens.
    `byte_fallback=True` fuses the ByteFallback op that follows it in detokenizer IRs."""

    def __init__(self, skip_tokens=(), device=0, byte_fallback=False):
        super().__init__()
        self.skip_tokens, self.device, self.byte_fallback = list(skip_tokens), device, bool(byte_fallback)

    def evaluate(self, inputs):
        if len(inputs) not in (4, 5):
            raise ValueError("Too few inputs passed to VocabDecoder, it means it is not converted properly or it is "
                             "not used in the supported pattern")
        if not self._h:
            keep = []
            d = K.VocabDecDesc(K.make_strings(inputs[1:4], keep), self.device)
            K.check(K.lib().b200tok_vocabdec_create(C.byref(d), C.byref(self._h)))
        ids = _i32(inputs[0])
        if ids.ndim != 2:
            raise ValueError("VocabDecoder expects ids of shape [batch, seq]")
        B, S = ids.shape
        skip = _i32(inputs[4] if len(inputs) == 5 else np.asarray(self.skip_tokens, dtype=np.int32))
        n = B * max(S, 1)
        cap = max(int(K.lib().b200tok_vocabdec_max_chars(self._h, B, S)), 1)
        rb, re_ = np.empty(B, np.int32), np.empty(B, np.int32)
        ob, oe = np.empty(max(n, 1), np.int32), np.empty(max(n, 1), np.int32)
        oc = np.empty(cap, np.uint8)
        out = K.Decoded(_ptr(rb), _ptr(re_), _ptr(ob), _ptr(oe), _ptr(oc), cap, 0, K.MEM_HOST)
        K.check(K.lib().b200tok_vocabdec_run(self._h, _ptr(ids), C.c_int64(B), C.c_int64(S), _ptr(skip),
                                             C.c_int64(len(skip)), int(self.byte_fallback), C.byref(out), K.MEM_HOST, None))
        return [rb, re_, ob[:n].copy(), oe[:n].copy(), oc[:out.n_chars].copy()]


class ByteFallback:
    """ByteFallback; inputs [0..2] strings -> strings."""

    def __init__(self, device=0):
        self.device = device

    def evaluate(self, inputs):
        b, e, c = _i32(inputs[0]), _i32(inputs[1]), _u8(inputs[2])
        ob, oe = np.empty(len(b), np.int32), np.empty(len(b), np.int32)
        oc = np.empty(max(c.size, 1), np.uint8)
        n = C.c_int64(0)
        K.check(K.lib().b200tok_bytefallback_run(self.device, _ptr(b), _ptr(e), C.c_int64(len(b)), _ptr(c), C.c_int64(c.size),
                                                 _ptr(ob), _ptr(oe), _ptr(oc), C.byref(n), K.MEM_HOST, None))
        return [ob, oe, oc[:n.value].copy()]


class BytesToChars:
    """BytesToChars; inputs ragged strings [0..4], optional skips [5] (reference src/bytes_to_chars.cpp:284-339)."""

    def __init__(self, device=0):
        self.device = device

    def evaluate(self, inputs):
        if len(inputs) not in (5, 6):
            raise ValueError("supported input sizes are 5 or 6")
        keep = []
        rin = _ragged_in(*inputs[:5], skips=inputs[5] if len(inputs) == 6 else None, keep=keep)
        ob, oe = np.empty(max(rin.n_elems, 1), np.int32), np.empty(max(rin.n_elems, 1), np.int32)
        cap = 2 * rin.n_chars
        oc = np.empty(max(cap, 1), np.uint8)
        n = C.c_int64(0)
        K.check(K.lib().b200tok_bytes_to_chars_run(self.device, C.byref(rin), _ptr(ob), _ptr(oe), _ptr(oc), C.c_int64(cap), C.byref(n), None))
        res = [keep[0], keep[1], ob[:rin.n_elems].copy(), oe[:rin.n_elems].copy(), oc[:n.value].copy()]
        if len(inputs) == 6:
            res.append(np.asarray(inputs[5]))
        return res


class CharsToBytes:
    """CharsToBytes; inputs ragged strings [0..4]; outputs one string per row (reference src/chars_to_bytes.cpp:31-68)."""

    def __init__(self, device=0):
        self.device = device

    def evaluate(self, inputs):
        if len(inputs) != 5:
            raise ValueError("CharsToBytes expects 5 inputs")
        keep = []
        rin = _ragged_in(*inputs[:5], keep=keep)
        ob, oe = np.empty(max(rin.n_rows, 1), np.int32), np.empty(max(rin.n_rows, 1), np.int32)
        cap = rin.n_chars
        oc = np.empty(max(cap, 1), np.uint8)
        n = C.c_int64(0)
        K.check(K.lib().b200tok_chars_to_bytes_run(self.device, C.byref(rin), _ptr(ob), _ptr(oe), _ptr(oc), C.c_int64(cap), C.byref(n), None))
        return [ob[:rin.n_rows].copy(), oe[:rin.n_rows].copy(), oc[:n.value].copy()]


class FuzeRagged:
    """FuzeRagged; inputs ragged_begins, ragged_ends, begins, ends (reference src/fuze.cpp:20-40)."""

    def __init__(self, device=0):
        self.device = device

    def evaluate(self, inputs):
        if len(inputs) != 4:
            raise ValueError("FuzeRagged expects 4 inputs")
        rb, re_, b, e = (_i32(x).reshape(-1) for x in inputs)
        ob, oe = np.empty(len(rb), np.int32), np.empty(len(rb), np.int32)
        K.check(K.lib().b200tok_fuze_ragged_run(self.device, _ptr(rb), _ptr(re_), C.c_int64(len(rb)), _ptr(b), _ptr(e), C.c_int64(len(b)),
                                                _ptr(ob), _ptr(oe), K.MEM_HOST, None))
        return [ob, oe]


class UTF8Validate:
    """UTF8Validate(replace_mode); inputs strings [0..2] (reference src/utf8_validate.cpp:18-137)."""

    def __init__(self, replace_mode=False, device=0):
        self.replace_mode, self.device = bool(replace_mode), device

    def evaluate(self, inputs):
        b, e, c = _i32(inputs[0]).reshape(-1), _i32(inputs[1]).reshape(-1), _u8(inputs[2]).reshape(-1)
        cap = 3 * c.size + (int(b[0]) if len(b) else 0)
        ob, oe = np.empty(max(len(b), 1), np.int32), np.empty(max(len(b), 1), np.int32)
        oc = np.zeros(max(cap, 1), np.uint8)
        n = C.c_int64(0)
        K.check(K.lib().b200tok_utf8_validate_run(self.device, _ptr(b), _ptr(e), C.c_int64(len(b)), _ptr(c) if c.size else None, C.c_int64(c.size),
                                                  int(self.replace_mode), _ptr(ob), _ptr(oe), _ptr(oc), C.c_int64(cap), C.byref(n), K.MEM_HOST, None))
        return [ob[:len(b)].copy(), oe[:len(b)].copy(), oc[:n.value].copy()]


class Truncate:
    """Truncate(num_inputs): inputs [3i..3i+2] ragged i32 (begins, ends, elems) per sequence, then max_length (i32 scalar),
    truncation side ("left"|"right") and mode ("only_first"|"only_second"|"longest_first") as u8 strings
    (reference src/truncate.cpp:37-147).  Outputs: the inputs with edited begins / ends."""

    def __init__(self, num_inputs=1, device=0):
        self.num_inputs, self.device = int(num_inputs), device

    def evaluate(self, inputs):
        n_in = self.num_inputs
        if len(inputs) != 3 * n_in + 3:
            raise ValueError("Truncate expects 3 * num_inputs + 3 inputs")
        max_length = int(np.asarray(inputs[-3]).reshape(-1)[0])
        side, mode = _as_text(inputs[-2]), _as_text(inputs[-1])
        arr = [_i32(inputs[3 * i + k]).copy() for i in range(n_in) for k in (0, 1)]
        n = len(arr[0])
        if any(len(a) != n for a in arr):
            raise ValueError("Begin and end tensors should have the same size")
        ptrs = [_ptr(a) for a in arr] + [None] * (4 - len(arr))
        K.check(K.lib().b200tok_truncate_run(self.device, n_in, *ptrs, C.c_int64(n), C.c_int32(max_length), side.encode(),
                                             mode.encode(), K.MEM_HOST, None))
        out = []
        for i in range(n_in):
            out += [arr[2 * i], arr[2 * i + 1], np.asarray(inputs[3 * i + 2])]
        return out


class CombineSegments:
    """CombineSegments: inputs [3j..3j+2] ragged i32 segments, last input = one id per segment (reference
    src/combine_segments.cpp:36-134).  Outputs: (begins, ends, elems) and (begins, ends, ids)."""

    def __init__(self, device=0):
        self.device = device

    def evaluate(self, inputs):
        num = (len(inputs) - 1) // 3
        ids = _i32(inputs[-1]).reshape(-1)
        if num < 1 or len(inputs) != 3 * num + 1 or len(ids) != num:
            raise ValueError("CombineSegments expects 3 * n + 1 inputs with one id per segment")
        segs = (K.RaggedI32 * num)()
        keep, cap, rows = [], 0, 0
        for j in range(num):
            b, e, x = _i32(inputs[3 * j]).reshape(-1), _i32(inputs[3 * j + 1]).reshape(-1), _i32(inputs[3 * j + 2]).reshape(-1)
            keep += [b, e, x]
            segs[j] = K.RaggedI32(b.ctypes.data, e.ctypes.data, len(b), x.ctypes.data if x.size else None, x.size)
            rows = max(rows, len(b))
        for j in range(num):   # the reference's flat_out_size estimate (:65-71)
            cap += (rows if segs[j].n == 1 else 1) * segs[j].n_elems
        ob, oe = np.empty(rows, np.int32), np.empty(rows, np.int32)
        ox, oi = np.empty(max(cap, 1), np.int32), np.empty(max(cap, 1), np.int32)
        n = C.c_int64(0)
        K.check(K.lib().b200tok_combine_segments_run(self.device, segs, num, _ptr(ids), _ptr(ob), _ptr(oe), _ptr(ox), _ptr(oi),
                                                     C.c_int64(cap), C.byref(n), K.MEM_HOST, None))
        return [ob, oe, ox[:n.value].copy(), ob.copy(), oe.copy(), oi[:n.value].copy()]


class RaggedToDense:
    """RaggedToDense(pad_right, pad_max_length): inputs begins, ends, elems (i32), target_dim, default value, optional
    pad_right bool (reference src/ragged_to_dense.cpp:70-174).  Outputs: dense i32[n, target_dim], mask bool."""

    def __init__(self, pad_right=True, pad_max_length=False, device=0):
        self.pad_right, self.pad_max_length, self.device = bool(pad_right), bool(pad_max_length), device

    def evaluate(self, inputs):
        if len(inputs) not in (5, 6):
            raise ValueError("RaggedToDense expects 5 or 6 inputs")
        b, e, x = _i32(inputs[0]).reshape(-1), _i32(inputs[1]).reshape(-1), _i32(inputs[2]).reshape(-1)
        target = int(np.asarray(inputs[3]).reshape(-1)[0])
        default = int(np.asarray(inputs[4]).reshape(-1)[0])
        pad_right = bool(np.asarray(inputs[5]).reshape(-1)[0]) if len(inputs) == 6 else self.pad_right
        out = np.empty((len(b), target), np.int32)
        mask = np.empty((len(b), target), np.uint8)
        K.check(K.lib().b200tok_ragged_to_dense_run(self.device, _ptr(b), _ptr(e), C.c_int64(len(b)), _ptr(x) if x.size else None,
                                                    C.c_int64(x.size), C.c_int32(target), C.c_int32(default), int(pad_right),
                                                    int(self.pad_max_length), _ptr(out), _ptr(mask), K.MEM_HOST, None))
        return [out, mask.astype(bool)]


def post_dense(begins, ends, ids, max_length, target_dim, pad_value, prefix=(), suffix=(), truncate_left=False, pad_right=True, device=0):
    """Fused Truncate -> CombineSegments(prefix, tokens, suffix) -> RaggedToDense on ragged ids (host arrays)."""
    b, e, x = _i32(begins), _i32(ends), _i32(ids)
    pre, suf = _i32(np.asarray(list(prefix), np.int32)), _i32(np.asarray(list(suffix), np.int32))
    d = K.PostDesc(int(max_length), int(truncate_left), pre.ctypes.data_as(K.i32p) if pre.size else None, pre.size,
                   suf.ctypes.data_as(K.i32p) if suf.size else None, suf.size, int(target_dim), int(pad_value), int(pad_right))
    out = np.empty((len(b), target_dim), np.int32)
    mask = np.empty((len(b), target_dim), np.uint8)
    K.check(K.lib().b200tok_post_dense_run(device, C.byref(d), _ptr(b), _ptr(e), C.c_int64(len(b)), _ptr(x) if x.size else None,
                                           C.c_int64(x.size), _ptr(out), _ptr(mask), K.MEM_HOST, None))
    return out, mask.astype(bool)


def split_bpe(split: RegexSplit, bpe: BPETokenizer, inputs):
    """Fused RegexSplit -> BPETokenizer on ragged strings `inputs[0..4]` (+ optional skips [5])."""
    keep = []
    rin = _ragged_in(*inputs[:5], skips=inputs[5] if len(inputs) > 5 else None, keep=keep)
    out, ob, oe, ids = _ids_out(rin.n_rows, rin.n_chars + rin.n_elems * len(bpe.end_suffix))
    K.check(K.lib().b200tok_split_bpe_run(split.handle, bpe.handle, C.byref(rin), C.byref(out), None))
    return [ob, oe, ids[:out.n_ids].copy()]


def split_wordpiece(split1: RegexSplit, split2, wp: WordpieceTokenizer, inputs, unk_token_id: int):
    """Fused RegexSplit [-> RegexSplit] -> WordpieceTokenizer (BERT pre-tokenisation chain)."""
    keep = []
    rin = _ragged_in(*inputs[:5], skips=inputs[5] if len(inputs) > 5 else None, keep=keep)
    out, ob, oe, ids = _ids_out(rin.n_rows, rin.n_chars + rin.n_elems)
    K.check(K.lib().b200tok_split_wordpiece_run(split1.handle, split2.handle if split2 is not None else None, wp.handle,
                                                C.byref(rin), C.c_int32(int(unk_token_id)), C.byref(out), None))
    return [ob, oe, ids[:out.n_ids].copy()]


class _Normalizer(_Handle):
    """Shared evaluate of the normalisers (reference evaluate_normalization_helper, src/utils.cpp:178-234):
    inputs strings [0..2] (+ skips [3] when `has_skips`); outputs strings (+ the skips tensor unchanged)."""

    device = 0
    _expand = 4

    def _run(self, inputs, has_skips):
        return _normalize([self], inputs, has_skips)


def _normalize(ops_, inputs, has_skips):
    b, e, c = _i32(inputs[0]).reshape(-1), _i32(inputs[1]).reshape(-1), _u8(inputs[2]).reshape(-1)
    sk = np.ascontiguousarray(inputs[3], np.uint8).reshape(-1) if has_skips else None
    n = len(b)
    ob, oe = np.empty(max(n, 1), np.int32), np.empty(max(n, 1), np.int32)
    cap = int(max(o._expand for o in ops_) * c.size) + 64
    hs = (C.c_void_p * len(ops_))(*[o.handle for o in ops_])
    for _ in range(2):
        oc = np.empty(max(cap, 1), np.uint8)
        got = C.c_int64(0)
        rc = K.lib().b200tok_normalize_chain_run(hs, len(ops_), _ptr(b), _ptr(e), C.c_int64(n), _ptr(c) if c.size else None, C.c_int64(c.size),
                                                 _ptr(sk) if sk is not None and n else None, _ptr(ob), _ptr(oe), _ptr(oc), C.c_int64(cap),
                                                 C.byref(got), K.MEM_HOST, None)
        if rc == K.E_CAPACITY and got.value > cap:      # the call reports the size it needs
            cap = got.value
            continue
        K.check(rc)
        break
    shape = np.asarray(inputs[0]).shape
    out = [ob[:n].copy().reshape(shape), oe[:n].copy().reshape(shape), oc[:got.value].copy()]
    if has_skips:
        out.append(inputs[3])
    return out


def normalize_chain(steps, inputs):
    """Run several normaliser ops back to back on the device (b200tok_normalize_chain_run): `steps` are prepared
    RegexNormalization / CharsMapNormalization objects (see their .prepare), `inputs` the strings [0..2] and optional skips [3]
    — the subgraph the converter emits for one HF normaliser, without a host round trip between the ops."""
    if len(inputs) not in (3, 4):
        raise ValueError("normalize_chain expects the strings and optionally the skips tensor")
    for s_ in steps:
        if not s_.handle:
            raise ValueError("normalize_chain: every step must be prepared (RegexNormalization.prepare / CharsMapNormalization.prepare)")
    return _normalize(list(steps), inputs, len(inputs) == 4)


class RegexNormalization(_Normalizer):
    """RegexNormalization(global_replace): inputs strings [0..2], optional skips [3], then search pattern and replace
    pattern as u8 strings (reference src/regex_normalization.cpp:59-153).  The handle is built from the pattern inputs
    on the first evaluate, like the reference's lazily compiled PCRE2 object; patterns outside the single-character set
    raise B200TokError(E_UNSUPPORTED)."""

    def __init__(self, global_replace=True, device=0):
        super().__init__()
        self.global_replace, self.device = bool(global_replace), device
        self._key = None

    def evaluate(self, inputs):
        if len(inputs) not in (5, 6):
            raise ValueError(f"supported input sizes are 5 or 6, got {len(inputs)}")
        has_skips = len(inputs) == 6
        self.prepare(_as_text(inputs[3 + has_skips]), _as_text(inputs[4 + has_skips]))
        return self._run(inputs, has_skips)

    def prepare(self, search_pattern, replace_pattern):
        search = search_pattern.encode() if isinstance(search_pattern, str) else bytes(search_pattern)
        replace = replace_pattern.encode() if isinstance(replace_pattern, str) else bytes(replace_pattern)
        if self._key != (search, replace):
            self.close()
            K.check(K.lib().b200tok_regexnorm_create(search, C.c_int64(len(search)), replace, C.c_int64(len(replace)),
                                                     int(self.global_replace), self.device, C.byref(self._h)))
            self._key = (search, replace)
            self._expand = 2 + len(replace)
        return self


class CharsMapNormalization(_Normalizer):
    """CharsMapNormalization: inputs strings [0..2], optional skips, then the precompiled charsmap (u8) — the 4/5-input
    form of the reference op (src/charsmap_normalization.cpp:13-69).  For the attribute form (`normalization_form`,
    `case_fold`) the caller passes the blob the reference's get_precompiled_charsmap() returns as `precompiled_charsmap`."""

    def __init__(self, precompiled_charsmap=None, add_dummy_prefix=False, remove_extra_whitespaces=False, escape_whitespaces=False, device=0):
        super().__init__()
        self.flags = (bool(add_dummy_prefix), bool(remove_extra_whitespaces), bool(escape_whitespaces))
        self.device = device
        self._blob = None if precompiled_charsmap is None else bytes(precompiled_charsmap)
        self._built = None
        self._expand = 3          # first guess; the call reports the size it needs if this is too small

    def evaluate(self, inputs):
        if len(inputs) not in (3, 4, 5):
            raise ValueError("CharsMapNormalization supports input sizes 3, 4 or 5.")
        has_skips = len(inputs) == 5 or (self._blob is not None and len(inputs) == 4)     # charsmap_normalization.cpp:35
        self.prepare(self._blob if self._blob is not None else bytes(_u8(inputs[3 + has_skips]).reshape(-1).tobytes()))
        return self._run(inputs, has_skips)

    def prepare(self, blob=None):
        blob = self._blob if blob is None else bytes(blob)
        if blob is None:
            raise ValueError("CharsMapNormalization: no precompiled charsmap")
        if self._built != blob:
            self.close()
            K.check(K.lib().b200tok_charsmap_create(blob, C.c_int64(len(blob)), *[int(f) for f in self.flags], self.device, C.byref(self._h)))
            self._built = blob
        return self
