#!/usr/bin/env python
"""Train the deterministic synthetic vocabularies used when no real assets are present.

There are no real tokenizer files (gpt2 / bert-base-uncased / Llama-3) on disk and no network
(SURVEY App. C), so parity and throughput are measured on stand-ins of identical size and shape,
trained offline with HuggingFace `tokenizers` on the text that ships inside this image (Python
sources + docs under site-packages) plus seeded pseudo-text for non-Latin scripts.  The outputs are
frozen under assets/ and committed, so the oracle, HF and the CUDA path all read identical bytes;
this script is the provenance record.

    python tools/make_assets.py [gpt2] [bert] [llama3]

Outputs  assets/<name>.tokenizer.json.gz  (HF format; loaded by openvino_tokenizers_b200.assets).
"""
from __future__ import annotations

import gzip
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ASSETS = ROOT / "assets"
CORPUS_ROOT = Path("/opt/prime-rl/.venv/lib/python3.12/site-packages")
CORPUS_BYTES = int(os.environ.get("B200TOK_CORPUS_BYTES", 160 << 20))

GPT2_PATTERN = r"'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+"
LLAMA3_PATTERN = (r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*"
                  r"|\s*[\r\n]+|\s+(?!\S)|\s+")


def corpus_files():
    files = []
    for dp, dn, fn in os.walk(CORPUS_ROOT):
        dn.sort()
        for f in sorted(fn):
            if f.endswith((".py", ".md", ".rst", ".txt", ".pyi", ".h", ".cuh", ".json")):
                files.append(os.path.join(dp, f))
    return files


def corpus_iter(limit=CORPUS_BYTES, lower=False):
    total = 0
    for path in corpus_files():
        try:
            with open(path, "r", encoding="utf-8") as fh:
                text = fh.read()
        except (UnicodeDecodeError, OSError):
            continue
        if len(text) > (1 << 20):
            text = text[: 1 << 20]
        total += len(text)
        yield text.lower() if lower else text
        if total >= limit:
            break
    yield from pseudo_multilingual()


def pseudo_multilingual(seed=1234, n_lines=60000):
    """Seeded Zipfian pseudo-words in Cyrillic / CJK / emoji so non-Latin bytes get merges too."""
    rng = np.random.default_rng(seed)
    cyr = [chr(c) for c in range(0x0410, 0x0450)]
    words = ["".join(rng.choice(cyr, size=int(rng.integers(2, 10)))) for _ in range(4000)]
    cjk = [chr(c) for c in range(0x4E00, 0x4E00 + 3000)]
    emo = [chr(c) for c in range(0x1F600, 0x1F650)]
    zw = 1.0 / np.arange(1, len(words) + 1)
    zw /= zw.sum()
    zc = 1.0 / np.arange(1, len(cjk) + 1)
    zc /= zc.sum()
    for i in range(n_lines):
        kind = i % 3
        if kind == 0:
            ws = rng.choice(len(words), size=12, p=zw)
            yield " ".join(words[j] for j in ws) + (". " if i % 2 else "! ")
        elif kind == 1:
            cs = rng.choice(len(cjk), size=24, p=zc)
            yield "".join(cjk[j] for j in cs) + "。"
        else:
            ws = rng.choice(len(words), size=4, p=zw)
            yield " ".join(words[j] for j in ws) + " " + "".join(rng.choice(emo, size=2)) + " ok 123\n"


def save(tok, name):
    ASSETS.mkdir(exist_ok=True)
    raw = tok.to_str().encode("utf-8")
    with gzip.GzipFile(ASSETS / f"{name}.tokenizer.json.gz", "wb", mtime=0) as fh:
        fh.write(raw)
    print(name, "vocab", tok.get_vocab_size(), "json bytes", len(raw))


def train_gpt2():
    from tokenizers import Tokenizer, decoders, models, pre_tokenizers, trainers
    tok = Tokenizer(models.BPE())
    tok.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=True)
    tok.decoder = decoders.ByteLevel()
    tr = trainers.BpeTrainer(vocab_size=50257, special_tokens=["<|endoftext|>"], show_progress=False,
                             initial_alphabet=pre_tokenizers.ByteLevel.alphabet())
    tok.train_from_iterator(corpus_iter(), tr)
    save(tok, "gpt2_synth")


def train_llama3():
    from tokenizers import Regex, Tokenizer, decoders, models, pre_tokenizers, trainers
    tok = Tokenizer(models.BPE(ignore_merges=False))
    tok.pre_tokenizer = pre_tokenizers.Sequence([
        pre_tokenizers.Split(Regex(LLAMA3_PATTERN), behavior="isolated", invert=False),
        pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False),
    ])
    tok.decoder = decoders.ByteLevel()
    specials = ["<|begin_of_text|>", "<|end_of_text|>"] + [f"<|reserved_special_token_{i}|>" for i in range(254)]
    tr = trainers.BpeTrainer(vocab_size=128000, show_progress=False,
                             initial_alphabet=pre_tokenizers.ByteLevel.alphabet())
    tok.train_from_iterator(corpus_iter(limit=2 * CORPUS_BYTES), tr)
    tok.add_special_tokens(specials)  # ids 128000.. like Llama-3
    save(tok, "llama3_synth")


def train_bert():
    from tokenizers import Tokenizer, decoders, models, normalizers, pre_tokenizers, trainers
    tok = Tokenizer(models.WordPiece(unk_token="[UNK]", max_input_chars_per_word=100))
    tok.normalizer = normalizers.BertNormalizer(lowercase=True)
    tok.pre_tokenizer = pre_tokenizers.BertPreTokenizer()
    tok.decoder = decoders.WordPiece(prefix="##")
    tr = trainers.WordPieceTrainer(vocab_size=30522, show_progress=False,
                                   special_tokens=["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"])
    tok.train_from_iterator(corpus_iter(limit=CORPUS_BYTES // 2), tr)
    save(tok, "bert_synth")


def make_llama2_detok():
    """C4 detokenizer vocab: 32 000 entries, <unk>/<s>/</s> at 0-2, 256 <0xHH> byte
    tokens at ids 3..258, the rest sentencepiece-style pieces derived from the gpt2_synth vocab."""
    import json
    from openvino_tokenizers_b200 import assets as A
    src = A.load_bpe("gpt2_synth")
    pieces = []
    seen = set()
    for tokb in src.vocab:
        try:
            s = tokb.decode("utf-8")
        except UnicodeDecodeError:
            continue
        if not s or s in seen or s.startswith("<0x"):
            continue
        seen.add(s)
        pieces.append(s.replace(" ", "▁"))
        if len(pieces) == 32000 - 259:
            break
    vocab = ["<unk>", "<s>", "</s>"] + [f"<0x{i:02X}>" for i in range(256)] + pieces
    assert len(vocab) == 32000, len(vocab)
    with gzip.GzipFile(ASSETS / "llama2_detok_synth.vocab.json.gz", "wb", mtime=0) as fh:
        fh.write(json.dumps(vocab, ensure_ascii=False).encode("utf-8"))
    print("llama2_detok_synth vocab", len(vocab))


if __name__ == "__main__":
    sys.path.insert(0, str(ROOT))
    which = sys.argv[1:] or ["gpt2", "bert", "llama3", "detok"]
    if "gpt2" in which:
        train_gpt2()
    if "bert" in which:
        train_bert()
    if "llama3" in which:
        train_llama3()
    if "detok" in which:
        make_llama2_detok()
