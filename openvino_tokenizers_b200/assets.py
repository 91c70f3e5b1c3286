"""Turn a HuggingFace ``tokenizer.json`` into the exact Constant tensors the reference converter
feeds the hot-path ops (host-side logic only; no OpenVINO needed).

Mirrors, without importing it, what the reference converter does:
  * vocab list with id = position and gaps filled with ""  (python/openvino_tokenizers/tokenizer_pipeline.py:518-530)
  * byte-level BPE: vocab and merges rewritten to raw bytes, merges become (left,right) pairs
    => 18-input BPETokenizer form  (tokenizer_pipeline.py:677-694,788-805)
  * added tokens = every ``added_tokens`` entry with a non-zero id  (tokenizer_pipeline.py:713)
  * GPT-2 byte<->unicode table  (python/openvino_tokenizers/utils.py:198-223)

Real assets are looked up first under ``$B200TOK_ASSETS/<name>/tokenizer.json``; otherwise the frozen
synthetic stand-ins under ``assets/`` (made by tools/make_assets.py) are used.
"""
from __future__ import annotations

import gzip
import json
import os
from dataclasses import dataclass, field
from functools import lru_cache
from pathlib import Path

from .strings import pack_strings

ASSET_DIR = Path(__file__).resolve().parent.parent / "assets"

GPT2_PATTERN = r"'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+"
GPT2_DIGITS_PATTERN = r"'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+|\p{N}| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+"
LLAMA3_PATTERN = (r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*"
                  r"|\s*[\r\n]+|\s+(?!\S)|\s+")
BERT_WHITESPACE_PATTERN = r"\s+"
BERT_PUNCT_PATTERN = "|".join([
    r"[!-/]", r"[:-@]", r"[\[-`]", r"[{-~]", r"[\p{P}]",
    r"[\x{4E00}-\x{9FFF}]", r"[\x{3400}-\x{4DBF}]", r"[\x{20000}-\x{2A6DF}]", r"[\x{2A700}-\x{2B73F}]",
    r"[\x{2B740}-\x{2B81F}]", r"[\x{2B820}-\x{2CEAF}]", r"[\x{F900}-\x{FAFF}]", r"[\x{2F800}-\x{2FA1F}]",
])


@lru_cache()
def unicode_to_bytes():
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAD)) + list(range(0xAE, 0x100))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return {chr(c): b for c, b in zip(cs, bs)}


def _to_bytes(token: str, keep_corrupted: bool = False) -> bytes:
    table = unicode_to_bytes()
    try:
        return bytes(table[ch] for ch in token)
    except KeyError:
        return token.encode() if keep_corrupted else b""


def _vocab_list(vocab: dict) -> list:
    out = []
    for tok, idx in sorted(vocab.items(), key=lambda kv: kv[1]):
        while len(out) < idx:
            out.append("")
        if len(out) == idx:
            out.append(tok)
        else:
            out[idx] = tok
    return out


def _find(name: str, suffix: str) -> Path:
    root = os.environ.get("B200TOK_ASSETS")
    if root:
        p = Path(root) / name / suffix
        if p.exists():
            return p
    p = ASSET_DIR / f"{name}.{suffix}.gz"
    if not p.exists():
        raise FileNotFoundError(f"no asset {name}: run tools/make_assets.py or set B200TOK_ASSETS")
    return p


def _read_json(path: Path):
    if path.suffix == ".gz":
        with gzip.open(path, "rb") as fh:
            return json.loads(fh.read().decode("utf-8"))
    return json.loads(path.read_text(encoding="utf-8"))


@dataclass
class BpeAssets:
    """The Constant inputs + attributes of one BPETokenizer node."""
    vocab: list                      # list[bytes], id = index
    merges: list                     # list[(bytes, bytes)]
    added_tokens: dict = field(default_factory=dict)   # bytes -> id
    unk_token: bytes = b""
    fuse_unk: bool = False
    end_suffix: bytes = b""
    byte_fallback: bool = False
    cache_capacity: int = 20000
    split_pattern: str = GPT2_PATTERN
    split_behaviour: str = "isolate"
    source: str = ""
    hf_json: str = ""                # raw tokenizer.json (for HF cross-checks in tests)

    def tensors(self):
        """(vocab, merges_left, merges_right, added, added_ids) as decomposed string tensors."""
        import numpy as np
        v = pack_strings(self.vocab)
        ml = pack_strings([m[0] for m in self.merges])
        mr = pack_strings([m[1] for m in self.merges])
        if self.added_tokens:
            a = pack_strings(list(self.added_tokens.keys()))
            aid = np.asarray(list(self.added_tokens.values()), dtype=np.int32)
        else:
            a, aid = None, None
        return v, ml, mr, a, aid


def load_bpe(name: str = "gpt2_synth") -> BpeAssets:
    path = _find(name, "tokenizer.json")
    tj = _read_json(path)
    model = tj["model"]
    assert model["type"] == "BPE", model["type"]
    pre = json.dumps(tj.get("pre_tokenizer") or {})
    byte_level = "ByteLevel" in pre
    vocab = _vocab_list(model["vocab"])
    merges = [tuple(m.split(" ")) if isinstance(m, str) else tuple(m) for m in model["merges"]]
    added = {t["content"]: t["id"] for t in tj.get("added_tokens", []) if t["id"]}
    if byte_level:
        vocab_b = [_to_bytes(t) for t in vocab]
        merges_b = [(_to_bytes(a), _to_bytes(b)) for a, b in merges]
    else:
        vocab_b = [t.encode() for t in vocab]
        merges_b = [(a.encode(), b.encode()) for a, b in merges]
    if added:
        need = max(added.values()) - len(vocab_b) + 1
        if need > 0:
            vocab_b.extend(b"" for _ in range(need))
    added_b = {}
    for tok, idx in added.items():
        tb = _to_bytes(tok, keep_corrupted=True) if byte_level else tok.encode()
        vocab_b[idx] = tb
        added_b[tok.encode()] = idx   # the Constant holds raw UTF-8 strings (tokenizer_pipeline.py:799-805)
    pattern = LLAMA3_PATTERN if "Split" in pre else GPT2_PATTERN
    return BpeAssets(
        vocab=vocab_b, merges=merges_b, added_tokens=added_b,
        unk_token=(model.get("unk_token") or "").encode(), fuse_unk=bool(model.get("fuse_unk")),
        end_suffix=(model.get("end_of_word_suffix") or "").encode(), byte_fallback=bool(model.get("byte_fallback")),
        cache_capacity=max(int(len(vocab) * 0.2), 20000), split_pattern=pattern, source=str(path),
        hf_json=json.dumps(tj),
    )


@dataclass
class WordpieceAssets:
    vocab: list                      # list[bytes]
    unk_token_id: int
    suffix_indicator: bytes = b"##"
    max_bytes_per_word: int = 100
    source: str = ""
    hf_json: str = ""


def load_wordpiece(name: str = "bert_synth") -> WordpieceAssets:
    path = _find(name, "tokenizer.json")
    tj = _read_json(path)
    model = tj["model"]
    assert model["type"] == "WordPiece", model["type"]
    vocab = [t.encode() for t in _vocab_list(model["vocab"])]
    unk = model.get("unk_token", "[UNK]")
    return WordpieceAssets(vocab=vocab, unk_token_id=model["vocab"][unk],
                           suffix_indicator=model.get("continuing_subword_prefix", "##").encode(),
                           max_bytes_per_word=int(model.get("max_input_chars_per_word", 100)),
                           source=str(path), hf_json=json.dumps(tj))


def load_detok_vocab(name: str = "llama2_detok_synth") -> list:
    path = _find(name, "vocab.json")
    return [t.encode() for t in _read_json(path)]
