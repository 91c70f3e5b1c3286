"""Helpers that build the reference's op nodes (oracle/ref.py: the reference's own classes compiled from
/root/reference/src) with the input lists the converter emits — shared by the pinning tests and bench.py."""
from __future__ import annotations

import numpy as np

from oracle import ref

_I32 = np.zeros(1, np.int32)
_U8 = np.zeros(1, np.uint8)
_BOOL = np.zeros(1, np.bool_)
_RAGGED = [_I32, _I32, _I32, _I32, _U8]
_STR = [_I32, _I32, _U8]


def available() -> bool:
    return ref.available()


def regex_split(pattern, behaviour="remove", invert=False, max_splits=-1, with_skips=True):
    pat = pattern.encode() if isinstance(pattern, str) else bytes(pattern)
    protos = _RAGGED + ([_BOOL] if with_skips else []) + [pat]
    op = ref.RefOp("RegexSplit", protos, constants={len(protos) - 1: pat}, behaviour=behaviour, invert=invert, max_splits=max_splits)

    def run(rb, re_, b, e, c, skips=None):
        ins = [rb, re_, b, e, c] + ([np.zeros(len(b), np.bool_) if skips is None else np.asarray(skips, np.bool_)] if with_skips else []) + [pat]
        out = op(*ins)
        return out[0], out[1], out[2], out[3], (out[5].astype(np.uint8) if with_skips else None)
    run.op = op
    return run


def special_tokens_split(pattern, with_skips=False):
    pat = pattern.encode() if isinstance(pattern, str) else bytes(pattern)
    protos = _RAGGED + ([_BOOL] if with_skips else []) + [pat]
    op = ref.RefOp("SpecialTokensSplit", protos, constants={len(protos) - 1: pat})

    def run(rb, re_, b, e, c, skips=None):
        ins = [rb, re_, b, e, c] + ([np.asarray(skips, np.bool_)] if with_skips else []) + [pat]
        out = op(*ins)
        return out[0], out[1], out[2], out[3], out[5].astype(np.uint8)
    run.op = op
    return run


def bpe(vocab, merges_left, merges_right=None, added=None, added_ids=None, **attrs):
    """11 / 14 / 15 / 18-input BPETokenizer (src/bpe_tokenizer.cpp:18-21)."""
    consts = [*vocab, *merges_left] + ([*merges_right] if merges_right is not None else []) + ([*added, np.asarray(added_ids, np.int32)] if added is not None else [])
    attrs = {k: v for k, v in attrs.items()}
    op = ref.RefOp("BPETokenizer", _RAGGED + consts, constants={5 + i: a for i, a in enumerate(consts)}, **attrs)

    def run(rb, re_, b, e, c):
        return tuple(op(rb, re_, b, e, c, *consts))
    run.op = op
    return run


def wordpiece(vocab, unk_id, suffix_indicator=b"##", max_bytes_per_word=100):
    unk = np.asarray(unk_id, np.int32)
    op = ref.RefOp("WordpieceTokenizer", _RAGGED + [*vocab, unk], constants={5: vocab[0], 6: vocab[1], 7: vocab[2], 8: unk},
                   suffix_indicator=suffix_indicator, max_bytes_per_word=max_bytes_per_word)

    def run(rb, re_, b, e, c):
        return tuple(op(rb, re_, b, e, c, *vocab, unk))
    run.op = op
    return run


def vocab_encoder(keys, values, default):
    values = np.asarray(values)
    default = np.asarray(default, values.dtype)
    op = ref.RefOp("VocabEncoder", _STR + [*keys, values, default], constants={3: keys[0], 4: keys[1], 5: keys[2], 6: values, 7: default})

    def run(b, e, c):
        return op(b, e, c, *keys, values, default)[0]
    return run


def vocab_decoder(vocab, skip_tokens, as_input=True):
    skip = np.asarray(list(skip_tokens), np.int32)
    ids_proto = np.zeros((1, 1), np.int32)
    if as_input:
        op = ref.RefOp("VocabDecoder", [ids_proto, *vocab, skip], constants={1: vocab[0], 2: vocab[1], 3: vocab[2]})
        return lambda ids: tuple(op(np.ascontiguousarray(ids, np.int32), *vocab, skip))
    op = ref.RefOp("VocabDecoder", [ids_proto, *vocab], constants={1: vocab[0], 2: vocab[1], 3: vocab[2]}, skip_tokens=skip.tolist())
    return lambda ids: tuple(op(np.ascontiguousarray(ids, np.int32), *vocab))


def simple(name, protos, **attrs):
    op = ref.RefOp(name, protos, **attrs)
    return lambda *ins: tuple(op(*ins))


class RefChain:
    """The op chain of a converted tokenizer IR on the reference's own code: RegexSplit -> BPETokenizer, or RegexSplit ->
    RegexSplit -> WordpieceTokenizer.  ``chain(batch, threads)`` shards the rows over `threads` concurrent evaluate() calls on
    the SAME op instances (shared tables and result cache, like several infer requests on one compiled model) and stitches
    the ragged results; threads = 1 is one plain evaluate() per op, the reference's serial path."""

    def __init__(self, kind, assets):
        from concurrent.futures import ThreadPoolExecutor
        from openvino_tokenizers_b200 import assets as A
        from openvino_tokenizers_b200.strings import pack_strings
        self.kind, self._pool_cls = kind, ThreadPoolExecutor
        if kind == "bpe":
            v, ml, mr, ad, aid = assets.tensors()
            self.pat = [assets.split_pattern.encode()]
            self.splits = [regex_split(assets.split_pattern, "isolate").op]
            self.consts = [*v, *ml, *mr] + ([*ad, np.asarray(aid, np.int32)] if ad is not None else [])
            self.tok = bpe(v, ml, mr, ad, aid, cache_capacity=assets.cache_capacity).op
        else:
            v = pack_strings(assets.vocab)
            self.pat = [A.BERT_WHITESPACE_PATTERN.encode(), A.BERT_PUNCT_PATTERN.encode()]
            self.splits = [regex_split(A.BERT_WHITESPACE_PATTERN, "remove").op, regex_split(A.BERT_PUNCT_PATTERN, "isolate").op]
            self.consts = [*v, np.asarray(assets.unk_token_id, np.int32)]
            self.tok = wordpiece(v, assets.unk_token_id, assets.suffix_indicator, assets.max_bytes_per_word).op
        self._shared = {}

    def _ops(self, t):
        if t not in self._shared:
            self._shared[t] = ([s if t == 0 else s.share() for s in self.splits], self.tok if t == 0 else self.tok.share())
        return self._shared[t]

    def _one(self, t, rb, re_, b, e, c):
        splits, tok = self._ops(t)
        cur = (rb, re_, b, e)
        for s, pat in zip(splits, self.pat):
            o = s(cur[0], cur[1], cur[2], cur[3], c, np.zeros(len(cur[2]), np.bool_), pat)
            cur = (o[0], o[1], o[2], o[3])
        return tok(cur[0], cur[1], cur[2], cur[3], c, *self.consts)

    def __call__(self, batch, threads=1):
        rb, re_, b, e, c = batch
        n = len(rb)
        threads = max(1, min(threads, n))
        if threads == 1:
            return tuple(self._one(0, rb, re_, b, e, c))
        cuts = [n * t // threads for t in range(threads + 1)]
        with self._pool_cls(threads) as ex:
            parts = list(ex.map(lambda t: self._one(t, rb[cuts[t]:cuts[t + 1]], re_[cuts[t]:cuts[t + 1]], b, e, c), range(threads)))
        offs = np.cumsum([0] + [len(p[2]) for p in parts])
        return (np.concatenate([p[0] + np.int32(o) for p, o in zip(parts, offs)]), np.concatenate([p[1] + np.int32(o) for p, o in zip(parts, offs)]),
                np.concatenate([p[2] for p in parts]))
