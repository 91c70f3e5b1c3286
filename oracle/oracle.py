"""ctypes wrapper over liboracle.so — TEST INFRASTRUCTURE ONLY (see oracle.cpp header).

All arrays are numpy, laid out exactly as the reference ops see them (decomposed strings:
``begins:i32, ends:i32, chars:u8``; ragged: ``ragged_begins:i32[B], ragged_ends:i32[B]`` in front,
reference src/utils.cpp:84-102).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "liboracle.so"

__all__ = [
    "build", "lib", "pcre2_available", "BpeOracle", "WordpieceOracle", "SplitOracle", "VocabEncoderOracle",
    "vocab_decoder", "byte_fallback", "truncate", "combine_segments", "ragged_to_dense", "SpecialTokensSplitOracle", "special_tokens_pattern", "bytes_to_chars", "chars_to_bytes", "fuze_ragged", "utf8_validate", "regex_normalize", "charsmap_normalize",
]


def build(force: bool = False) -> Path:
    """Compile liboracle.so with the committed Makefile (gcc only; no reference sources involved)."""
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < (_HERE / "oracle.cpp").stat().st_mtime:
        subprocess.check_call(["make", "-C", str(_HERE), "-s", "-B", "liboracle.so"])
    return _LIB_PATH


_lib = None
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_u8p = C.POINTER(C.c_uint8)


def lib():
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            build()
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.orc_bpe_create.restype = C.c_void_p
        _lib.orc_wp_create.restype = C.c_void_p
        _lib.orc_split_create.restype = C.c_void_p
        _lib.orc_venc_create.restype = C.c_void_p
        for f in ("orc_bpe_run", "orc_wp_run", "orc_split_run", "orc_vdec_run", "orc_bytefallback_run"):
            getattr(_lib, f).restype = C.c_int64
    return _lib


def _p(a, typ):
    if a is None:
        return C.cast(None, typ)
    return a.ctypes.data_as(typ)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _u8(a):
    if isinstance(a, (bytes, bytearray)):
        a = np.frombuffer(bytes(a), dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8)


def pcre2_available() -> bool:
    return bool(lib().orc_pcre2_available())


class BpeOracle:
    """BPETokenizer (reference src/bpe_tokenizer.cpp).  ``merges_right=None`` selects the "L R" string form."""

    def __init__(self, vocab, merges_left, merges_right=None, added=None, added_ids=None, *, unk_token=b"",
                 fuse_unk=False, end_suffix=b"", byte_fallback=False, cache_capacity=20000, use_cache=True):
        L = lib()
        vb, ve, vc = (_i32(vocab[0]), _i32(vocab[1]), _u8(vocab[2]))
        lb, le, lc = (_i32(merges_left[0]), _i32(merges_left[1]), _u8(merges_left[2]))
        if merges_right is not None:
            rb, re_, rc = (_i32(merges_right[0]), _i32(merges_right[1]), _u8(merges_right[2]))
        else:
            rb = re_ = rc = None
        if added is not None:
            ab, ae, ac = (_i32(added[0]), _i32(added[1]), _u8(added[2]))
            aid = _i32(added_ids)
            A = len(aid)
        else:
            ab = ae = ac = aid = None
            A = 0
        unk = _u8(unk_token)
        es = _u8(end_suffix)
        self._keep = (vb, ve, vc, lb, le, lc, rb, re_, rc, ab, ae, ac, aid, unk, es)
        self._h = C.c_void_p(L.orc_bpe_create(
            _p(vb, _i32p), _p(ve, _i32p), _p(vc, _u8p), C.c_int64(len(vb)),
            _p(lb, _i32p), _p(le, _i32p), _p(lc, _u8p),
            _p(rb, _i32p), _p(re_, _i32p), _p(rc, _u8p), C.c_int64(len(lb)),
            _p(ab, _i32p), _p(ae, _i32p), _p(ac, _u8p), _p(aid, _i32p), C.c_int64(A),
            _p(unk, _u8p), C.c_int64(len(unk)), C.c_int(int(fuse_unk)),
            _p(es, _u8p), C.c_int64(len(es)), C.c_int(int(byte_fallback)),
            C.c_int64(cache_capacity), C.c_int(int(use_cache))))
        if not self._h:
            raise ValueError("oracle: BPE table construction failed (merge refers to a missing token)")

    def clear_cache(self):
        lib().orc_bpe_clear_cache(self._h)

    def __call__(self, rb, re_, begins, ends, chars, threads=1):
        rb, re_, begins, ends, chars = _i32(rb), _i32(re_), _i32(begins), _i32(ends), _u8(chars)
        B = len(rb)
        cap = max(int(np.sum(np.maximum(ends - begins, 0))) * 2 + 16, len(chars) + 16)
        ob, oe = np.empty(B, np.int32), np.empty(B, np.int32)
        ids = np.empty(cap, np.int32)
        T = lib().orc_bpe_run(self._h, _p(rb, _i32p), _p(re_, _i32p), C.c_int64(B), _p(begins, _i32p), _p(ends, _i32p),
                              _p(chars, _u8p), _p(ob, _i32p), _p(oe, _i32p), _p(ids, _i32p), C.c_int64(cap), C.c_int(threads))
        assert T >= 0
        return ob, oe, ids[:T].copy()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_bpe_destroy(self._h)
            self._h = None


class WordpieceOracle:
    """WordpieceTokenizer (reference src/wordpiece_tokenizer.cpp:49-133)."""

    def __init__(self, vocab, suffix_indicator=b"##", max_bytes_per_word=100):
        vb, ve, vc = (_i32(vocab[0]), _i32(vocab[1]), _u8(vocab[2]))
        sfx = _u8(suffix_indicator)
        self._h = C.c_void_p(lib().orc_wp_create(_p(vb, _i32p), _p(ve, _i32p), _p(vc, _u8p), C.c_int64(len(vb)),
                                                 _p(sfx, _u8p), C.c_int64(len(sfx)), C.c_int(max_bytes_per_word)))

    def __call__(self, rb, re_, begins, ends, chars, unk_id, threads=1):
        rb, re_, begins, ends, chars = _i32(rb), _i32(re_), _i32(begins), _i32(ends), _u8(chars)
        B = len(rb)
        cap = len(chars) + len(begins) + 16
        ob, oe = np.empty(B, np.int32), np.empty(B, np.int32)
        ids = np.empty(cap, np.int32)
        T = lib().orc_wp_run(self._h, _p(rb, _i32p), _p(re_, _i32p), C.c_int64(B), _p(begins, _i32p), _p(ends, _i32p),
                             _p(chars, _u8p), C.c_int32(unk_id), _p(ob, _i32p), _p(oe, _i32p), _p(ids, _i32p),
                             C.c_int64(cap), C.c_int(threads))
        assert T >= 0
        return ob, oe, ids[:T].copy()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_wp_destroy(self._h)
            self._h = None


class SplitOracle:
    """RegexSplit (reference src/regex_split.cpp:124-324) on the system PCRE2 (UTF|UCP, JIT)."""

    def __init__(self, pattern, behaviour="remove", invert=False, max_splits=-1, skip_tokens=None):
        if not pcre2_available():
            raise RuntimeError("oracle: libpcre2-8.so.0 not found")
        pat = _u8(pattern.encode() if isinstance(pattern, str) else pattern)
        if skip_tokens is not None:
            sb, se, sc = _i32(skip_tokens[0]), _i32(skip_tokens[1]), _u8(skip_tokens[2])
            ns = len(sb)
        else:
            sb = se = sc = None
            ns = 0
        self._h = C.c_void_p(lib().orc_split_create(_p(pat, _u8p), C.c_int64(len(pat)), behaviour.lower().encode(),
                                                    C.c_int(int(invert)), C.c_int(max_splits),
                                                    _p(sb, _i32p), _p(se, _i32p), _p(sc, _u8p), C.c_int64(ns)))
        if not self._h:
            raise ValueError(f"oracle: cannot compile pattern / unknown behaviour {behaviour!r}")

    def __call__(self, rb, re_, begins, ends, chars, skips=None, threads=1):
        """Returns (rb', re', begins', ends', skips').  chars is passed through by the op (aliased)."""
        rb, re_, begins, ends, chars = _i32(rb), _i32(re_), _i32(begins), _i32(ends), _u8(chars)
        B = len(rb)
        if len(chars) == 0:  # regex_split.cpp:129-143
            z = np.zeros(1, np.int32)
            return z, z.copy(), begins, ends, (None if skips is None else np.asarray(skips, np.uint8))
        sk = None if skips is None else _u8(np.asarray(skips, dtype=np.uint8))
        cap = len(chars) + len(begins) + 16
        orb, ore = np.empty(B, np.int32), np.empty(B, np.int32)
        ob, oe, osk = np.empty(cap, np.int32), np.empty(cap, np.int32), np.empty(cap, np.uint8)
        P = lib().orc_split_run(self._h, _p(rb, _i32p), _p(re_, _i32p), C.c_int64(B), _p(begins, _i32p), _p(ends, _i32p),
                                _p(chars, _u8p), C.c_int64(len(chars)), _p(sk, _u8p),
                                _p(orb, _i32p), _p(ore, _i32p), _p(ob, _i32p), _p(oe, _i32p), _p(osk, _u8p),
                                C.c_int64(cap), C.c_int(threads))
        assert P >= 0
        return orb, ore, ob[:P].copy(), oe[:P].copy(), osk[:P].copy()

    def fullmatch_cp(self, cp: int) -> bool:
        return bool(lib().orc_regex_fullmatch_cp(self._h, C.c_uint32(cp)))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_split_destroy(self._h)
            self._h = None


def special_tokens_pattern(tokens):
    """The split pattern the reference converter builds for a list of special tokens — a restatement of
    SpecialTokensSplit.get_ov_subgraph (reference python/openvino_tokenizers/tokenizer_pipeline.py:138-158) and quote_meta
    (utils.py:421-429).  tokens: iterable of (text, strip_left, strip_right); sorted descending like __post_init__ (:95-97)."""
    def quote_meta(t):
        return "".join(("" if (ch.isalnum() or ch in ("_", "\u2581", "\uff5c")) else "\\") + ch for ch in t)
    toks = sorted({(str(t[0]), bool(t[1]), bool(t[2])) for t in tokens}, reverse=True)
    groups = {}
    for text, sl, sr in toks:
        groups.setdefault((sl, sr), []).append(text)
    return "|".join(r"(?:\s*)" * sl + "(" + "|".join(quote_meta(t) for t in ts) + ")" + r"(?:\s*)" * sr for (sl, sr), ts in groups.items())


class SpecialTokensSplitOracle(SplitOracle):
    """SpecialTokensSplit (reference src/special_tokens_split.cpp:61-162) on the system PCRE2."""

    def __init__(self, pattern):
        super().__init__(pattern, "isolate")

    def __call__(self, rb, re_, begins, ends, chars, skips=None):
        rb, re_, begins, ends, chars = _i32(rb), _i32(re_), _i32(begins), _i32(ends), _u8(chars)
        B = len(rb)
        sk = None if skips is None else _u8(np.asarray(skips, dtype=np.uint8))
        cap = len(chars) + len(begins) + 16
        orb, ore = np.empty(B, np.int32), np.empty(B, np.int32)
        ob, oe, osk = np.empty(cap, np.int32), np.empty(cap, np.int32), np.empty(cap, np.uint8)
        lib().orc_special_split_run.restype = C.c_int64
        P = lib().orc_special_split_run(self._h, _p(rb, _i32p), _p(re_, _i32p), C.c_int64(B), _p(begins, _i32p), _p(ends, _i32p),
                                        _p(chars, _u8p), _p(sk, _u8p), _p(orb, _i32p), _p(ore, _i32p), _p(ob, _i32p), _p(oe, _i32p),
                                        _p(osk, _u8p), C.c_int64(cap))
        assert P >= 0
        return orb, ore, ob[:P].copy(), oe[:P].copy(), osk[:P].copy()


class VocabEncoderOracle:
    """VocabEncoder (reference src/vocab_encoder.cpp:56-94)."""

    def __init__(self, keys, values):
        vb, ve, vc = _i32(keys[0]), _i32(keys[1]), _u8(keys[2])
        vals = np.ascontiguousarray(values, dtype=np.int64)
        self._h = C.c_void_p(lib().orc_venc_create(_p(vb, _i32p), _p(ve, _i32p), _p(vc, _u8p), _p(vals, _i64p),
                                                   C.c_int64(len(vb))))

    def __call__(self, begins, ends, chars, default_value):
        begins, ends, chars = _i32(begins), _i32(ends), _u8(chars)
        out = np.empty(len(begins), np.int64)
        lib().orc_venc_run(self._h, _p(begins, _i32p), _p(ends, _i32p), _p(chars, _u8p), C.c_int64(len(begins)),
                           C.c_int64(default_value), _p(out, _i64p))
        return out

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_venc_destroy(self._h)
            self._h = None


def vocab_decoder(ids, vocab, skip_tokens=()):
    """VocabDecoder (reference src/vocab_decoder.cpp:23-87).  ids: i32[B,S]."""
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    B, S = ids.shape
    vb, ve, vc = _i32(vocab[0]), _i32(vocab[1]), _u8(vocab[2])
    skip = _i32(np.asarray(list(skip_tokens), dtype=np.int32))
    n = B * max(S, 1)
    maxlen = int((ve - vb).max()) if len(vb) else 0
    cap = n * maxlen + 16
    rb, re_ = np.empty(B, np.int32), np.empty(B, np.int32)
    ob, oe = np.empty(n, np.int32), np.empty(n, np.int32)
    oc = np.empty(cap, np.uint8)
    N = lib().orc_vdec_run(_p(ids, _i32p), C.c_int64(B), C.c_int64(S), _p(vb, _i32p), _p(ve, _i32p), _p(vc, _u8p),
                           C.c_int64(len(vb)), _p(skip, _i32p), C.c_int64(len(skip)),
                           _p(rb, _i32p), _p(re_, _i32p), _p(ob, _i32p), _p(oe, _i32p), _p(oc, _u8p), C.c_int64(cap))
    assert N >= 0
    return rb, re_, ob, oe, oc[:N].copy()


def byte_fallback(begins, ends, chars):
    """ByteFallback (reference src/byte_fallback.cpp:16-50)."""
    begins, ends, chars = _i32(begins), _i32(ends), _u8(chars)
    ob, oe = np.empty(len(begins), np.int32), np.empty(len(begins), np.int32)
    oc = np.empty(len(chars) + 16, np.uint8)
    N = lib().orc_bytefallback_run(_p(begins, _i32p), _p(ends, _i32p), _p(chars, _u8p), C.c_int64(len(begins)),
                                   _p(ob, _i32p), _p(oe, _i32p), _p(oc, _u8p))
    return ob, oe, oc[:N].copy()


def truncate(pairs, max_length, side="right", mode="longest_first"):
    """Truncate (reference src/truncate.cpp:37-147).  pairs: [(begins, ends)] for one or two ragged inputs; returns new
    [(begins, ends)] (the reference edits its aliased outputs in place)."""
    out = [(_i32(b).copy(), _i32(e).copy()) for b, e in pairs]
    b0, e0 = out[0]
    b1, e1 = out[1] if len(out) > 1 else (None, None)
    rc = lib().orc_truncate(len(out), _p(b0, _i32p), _p(e0, _i32p), _p(b1, _i32p), _p(e1, _i32p), C.c_int64(len(b0)),
                            C.c_int32(max_length), side.encode(), mode.encode())
    if rc:
        raise ValueError(f"Unknown truncation side/mode: {side}/{mode}")
    return out


def combine_segments(segments, ids):
    """CombineSegments (reference src/combine_segments.cpp:36-134), i32 elements.  segments: [(begins, ends, elems)];
    a segment with one row is broadcast.  Returns (begins, ends, elems, ids) of the combined ragged tensor."""
    segs = [(_i32(b), _i32(e), _i32(x)) for b, e, x in segments]
    ids = _i32(ids)
    num = len(segs)
    rows = max(len(s[0]) for s in segs)
    total = sum((int((s[1] - s[0]).sum()) if len(s[0]) > 1 or rows == 1 else int(s[1][0] - s[0][0]) * rows) for s in segs)
    PP = _i32p * num
    nn = (C.c_int64 * num)(*[len(s[0]) for s in segs])
    ob, oe = np.empty(rows, np.int32), np.empty(rows, np.int32)
    ox, oi = np.empty(total + 1, np.int32), np.empty(total + 1, np.int32)
    lib().orc_combine_segments.restype = C.c_int64
    n = lib().orc_combine_segments(num, PP(*[_p(s[0], _i32p) for s in segs]), PP(*[_p(s[1], _i32p) for s in segs]), nn,
                                   PP(*[_p(s[2], _i32p) for s in segs]), _p(ids, _i32p), _p(ob, _i32p), _p(oe, _i32p),
                                   _p(ox, _i32p), _p(oi, _i32p))
    assert n == total
    return ob, oe, ox[:n].copy(), oi[:n].copy()


def ragged_to_dense(begins, ends, elems, target_dim, default_value, pad_right=True, pad_max_length=False):
    """RaggedToDense (reference src/ragged_to_dense.cpp:70-174), i32 elements.  Returns (dense i32[n, target], mask u8)."""
    begins, ends, elems = _i32(begins), _i32(ends), _i32(elems)
    n = len(begins)
    out = np.empty((n, target_dim), np.int32)
    mask = np.empty((n, target_dim), np.uint8)
    lib().orc_ragged_to_dense(_p(begins, _i32p), _p(ends, _i32p), C.c_int64(n), _p(elems, _i32p), C.c_int32(target_dim),
                              C.c_int32(default_value), int(bool(pad_right)), int(bool(pad_max_length)), _p(out, _i32p), _p(mask, _u8p))
    return out, mask


def bytes_to_chars(rb, re_, begins, ends, chars, skips=None):
    """BytesToChars (reference src/bytes_to_chars.cpp:284-339).  Returns (begins, ends, chars)."""
    rb, re_, begins, ends, chars = _i32(rb), _i32(re_), _i32(begins), _i32(ends), _u8(chars)
    sk = None if skips is None else _u8(np.asarray(skips, dtype=np.uint8))
    ob, oe = np.zeros(len(begins), np.int32), np.zeros(len(begins), np.int32)
    oc = np.empty(2 * len(chars) + 16, np.uint8)
    lib().orc_bytes_to_chars.restype = C.c_int64
    n = lib().orc_bytes_to_chars(_p(rb, _i32p), _p(re_, _i32p), C.c_int64(len(rb)), _p(begins, _i32p), _p(ends, _i32p), _p(chars, _u8p),
                                 _p(sk, _u8p), _p(ob, _i32p), _p(oe, _i32p), _p(oc, _u8p))
    return ob, oe, oc[:n].copy()


def chars_to_bytes(rb, re_, begins, ends, chars):
    """CharsToBytes (reference src/chars_to_bytes.cpp:31-68).  Returns one string per row: (begins, ends, chars)."""
    rb, re_, begins, ends, chars = _i32(rb), _i32(re_), _i32(begins), _i32(ends), _u8(chars)
    ob, oe = np.zeros(len(rb), np.int32), np.zeros(len(rb), np.int32)
    oc = np.empty(len(chars) + 16, np.uint8)
    lib().orc_chars_to_bytes.restype = C.c_int64
    n = lib().orc_chars_to_bytes(_p(rb, _i32p), _p(re_, _i32p), C.c_int64(len(rb)), _p(begins, _i32p), _p(ends, _i32p), _p(chars, _u8p),
                                 _p(ob, _i32p), _p(oe, _i32p), _p(oc, _u8p))
    return ob, oe, oc[:n].copy()


def fuze_ragged(rb, re_, begins, ends):
    """FuzeRagged (reference src/fuze.cpp:20-40)."""
    rb, re_, begins, ends = _i32(rb), _i32(re_), _i32(begins), _i32(ends)
    ob, oe = np.empty(len(rb), np.int32), np.empty(len(rb), np.int32)
    lib().orc_fuze_ragged(_p(rb, _i32p), _p(re_, _i32p), C.c_int64(len(rb)), _p(begins, _i32p), _p(ends, _i32p), _p(ob, _i32p), _p(oe, _i32p))
    return ob, oe


def utf8_validate(begins, ends, chars, replace_mode):
    """UTF8Validate (reference src/utf8_validate.cpp:18-137).  Returns (begins, ends, chars[: extent])."""
    begins, ends, chars = _i32(begins), _i32(ends), _u8(chars)
    ob, oe = np.empty(len(begins), np.int32), np.empty(len(begins), np.int32)
    oc = np.zeros(3 * len(chars) + 16 + (int(begins[0]) if len(begins) else 0), np.uint8)
    lib().orc_utf8_validate.restype = C.c_int64
    n = lib().orc_utf8_validate(_p(begins, _i32p), _p(ends, _i32p), C.c_int64(len(begins)), _p(chars, _u8p), int(bool(replace_mode)),
                                _p(ob, _i32p), _p(oe, _i32p), _p(oc, _u8p))
    return ob, oe, oc[:n].copy()


def _normalize_call(fn, head, begins, ends, chars, skips, expand):
    begins, ends, chars = _i32(begins), _i32(ends), _u8(chars)
    sk = None if skips is None else np.ascontiguousarray(skips, np.uint8)
    n = len(begins)
    ob, oe = np.empty(n, np.int32), np.empty(n, np.int32)
    cap = int(expand * max(int((ends - begins).clip(min=0).sum()), 0)) + 64 * n + 64
    oc = np.zeros(cap, np.uint8)
    fn.restype = C.c_int64
    total = fn(*head, _p(begins, _i32p), _p(ends, _i32p), _p(chars, _u8p), None if sk is None else _p(sk, _u8p), C.c_int64(n),
               _p(ob, _i32p), _p(oe, _i32p), _p(oc, _u8p), C.c_int64(cap))
    assert total <= cap
    return ob, oe, oc[:total].copy()


def regex_normalize(search_pattern, replace_pattern, global_replace, begins, ends, chars, skips=None):
    """RegexNormalization (reference src/regex_normalization.cpp:127-153) through PCRE2's own pcre2_substitute."""
    sp = search_pattern.encode() if isinstance(search_pattern, str) else bytes(search_pattern)
    rp = replace_pattern.encode() if isinstance(replace_pattern, str) else bytes(replace_pattern)
    head = (sp, C.c_int64(len(sp)), rp, C.c_int64(len(rp)), int(bool(global_replace)))
    return _normalize_call(lib().orc_regex_normalize, head, begins, ends, chars, skips, 4 + 4 * len(rp))


def charsmap_normalize(blob, begins, ends, chars, skips=None, add_dummy_prefix=False, remove_extra_whitespaces=False, escape_whitespaces=False):
    """CharsMapNormalization (reference src/charsmap_normalization.cpp:34-69): sentencepiece Normalizer over a precompiled charsmap."""
    blob = bytes(blob)
    head = (blob, C.c_int64(len(blob)), int(add_dummy_prefix), int(remove_extra_whitespaces), int(escape_whitespaces))
    return _normalize_call(lib().orc_charsmap_normalize, head, begins, ends, chars, skips, 20)
