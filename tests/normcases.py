"""Shared inputs of the normaliser tests: corpora and charsmap blobs.

Charsmap blobs are never taken from the reference tree: they are compiled HERE by the installed sentencepiece package —
either one of its built-in rule sets, or a rule TSV generated from Python's unicodedata (NFD, case folding, both), which is
what the reference's generated header holds for its `nfd` / `identity+case_fold` forms up to the Unicode version."""
import io
import random
import tempfile
import unicodedata
from pathlib import Path

import numpy as np

_cache = {}


def _train_blob(**kw):
    import sentencepiece as spm
    from sentencepiece import sentencepiece_model_pb2 as pb
    buf = io.BytesIO()
    words = ["alpha", "beta", "gamma", "delta", "epsilon", "zeta", "eta", "theta", "iota", "kappa", "lambda", "mu"]
    sents = [" ".join(words[(i + j) % len(words)] for j in range(6)) for i in range(200)]
    spm.SentencePieceTrainer.train(sentence_iterator=iter(sents), model_writer=buf, vocab_size=30, model_type="unigram",
                                   hard_vocab_limit=False, minloglevel=2, **kw)
    mp = pb.ModelProto()
    mp.ParseFromString(buf.getvalue())
    return bytes(mp.normalizer_spec.precompiled_charsmap)


def builtin_blob(rule_name):
    """A charsmap compiled into the sentencepiece library (nfkc, nfkc_cf, nmt_nfkc, nmt_nfkc_cf)."""
    key = ("builtin", rule_name)
    if key not in _cache:
        _cache[key] = _train_blob(normalization_rule_name=rule_name)
    return _cache[key]


def unicodedata_blob(form=None, case_fold=False, limit=0x30000):
    """A charsmap for `form` (NFD/NFC/NFKD/NFKC or None) followed by case folding, rules from unicodedata."""
    key = ("ucd", form, case_fold)
    if key not in _cache:
        lines = []
        for cp in range(1, limit):
            if 0xD800 <= cp <= 0xDFFF:
                continue
            ch = chr(cp)
            t = unicodedata.normalize(form, ch) if form else ch
            if case_fold:
                t = t.casefold()
                if form:
                    t = unicodedata.normalize(form, t)
            if t != ch and "\0" not in t and t:
                lines.append(" ".join(f"{ord(c):X}" for c in ch) + "\t" + " ".join(f"{ord(c):X}" for c in t))
        with tempfile.TemporaryDirectory() as d:
            p = Path(d) / "rules.tsv"
            p.write_text("\n".join(lines) + "\n")
            _cache[key] = _train_blob(normalization_rule_tsv=str(p))
    return _cache[key]


def custom_blob():
    """Rules that stress the ASCII shortcuts: ASCII -> ASCII, ASCII -> multi-byte, ASCII sequences, ASCII + mark, deletions."""
    key = ("custom",)
    if key not in _cache:
        rules = [("41", "61"), ("42", "C9"), ("43 44", "78"), ("45 301", "C9"), ("46", "66 66"), ("71", ""), ("20 20", "20"),
                 ("7A 7A 7A", "5A"), ("C9", "65"), ("301", ""), ("31", "32"), ("32", "31")]
        with tempfile.TemporaryDirectory() as d:
            p = Path(d) / "rules.tsv"
            p.write_text("".join(f"{a}\t{b}\n" for a, b in rules))
            _cache[key] = _train_blob(normalization_rule_tsv=str(p))
    return _cache[key]


def ascii_corpus(seed=7, n=400, max_len=200):
    """ASCII-only strings (the kernels' fast chunks), with tabs / controls, lengths around the 32-byte chunk size, and a
    sprinkle of strings whose only non-ASCII character sits right after a chunk boundary."""
    rng = random.Random(seed)
    out = []
    for i in range(n):
        ln = rng.choice([0, 1, 31, 32, 33, 63, 64, 65, rng.randint(0, max_len)])
        s = bytearray(rng.choice(b"ABCDEFzzq  12 \t\x01abcdefghijklmnop") for _ in range(ln))
        if i % 5 == 0 and ln >= 32:
            s[32:32] = "\u0301".encode()
        elif i % 7 == 0 and ln >= 33:          # (never both: the strings stay well-formed UTF-8)
            s[31:33] = b"E" + "\u0301".encode()
        out.append(bytes(s))
    return out


def sp_normalizer(blob, add_dummy_prefix=False, remove_extra_whitespaces=False, escape_whitespaces=False):
    """The real sentencepiece Normalizer over a blob (what CharsMapNormalization::evaluate calls)."""
    import sentencepiece as spm
    from sentencepiece import sentencepiece_model_pb2 as pb
    mp = pb.ModelProto()
    mp.normalizer_spec.precompiled_charsmap = blob
    for p, t in [("<unk>", 2), ("<s>", 3), ("</s>", 3), ("a", 1)]:
        x = mp.pieces.add()
        x.piece, x.score, x.type = p, 0.0, t
    return spm.SentencePieceNormalizer(model_proto=mp.SerializeToString(), add_dummy_prefix=bool(add_dummy_prefix),
                                       escape_whitespaces=bool(escape_whitespaces), remove_extra_whitespaces=bool(remove_extra_whitespaces))


INTERESTING = [0x300, 0x301, 0x308, 0x327, 0x323, 0x1100, 0x1161, 0xAC00, 0xFB01, 0x2126, 0x212B, 0x1E9B,      # marks, Hangul, ligatures
               0x4E2D, 0x3400, 0x20000, 0x3042, 0xFF21, 0x2460, 0xDF, 0x130, 0x1E9E, 0x1F600,                # Han, kana, wide, sharp s, emoji
               0xAD, 0x200B, 0x200D, 0x2060, 0xFEFF, 0x600, 0x61C, 0x85, 0x9F, 0x7F, 0x1, 0x8, 0xB, 0xC, 0xE, 0x1F,   # Cf / control
               0x9, 0xA, 0xD, 0x20, 0xA0, 0x1680, 0x2000, 0x2003, 0x2028, 0x2029, 0x202F, 0x205F, 0x3000,       # whitespace
               0x483, 0x591, 0x64B, 0x93C, 0x20D0, 0xFE20]                                                    # more Mn


def corpus(seed=5, n=1500, malformed=0, max_len=60):
    """Byte strings: ASCII, Latin, Greek/Cyrillic and the INTERESTING code points; some long; optionally malformed UTF-8."""
    rng = random.Random(seed)

    def rand_str(length):
        out = []
        for _ in range(length):
            r = rng.random()
            if r < .45: out.append(chr(rng.randint(0x20, 0x7E)))
            elif r < .55: out.append(" ")
            elif r < .68: out.append(chr(rng.randint(0xA0, 0x24F)))
            elif r < .88: out.append(chr(rng.choice(INTERESTING)))
            else: out.append(chr(rng.randint(0x370, 0x52F)))
        return "".join(out)

    raw = [rand_str(rng.randint(0, max_len)).encode() for _ in range(n)]
    raw += [rand_str(rng.randint(100, 400)).encode() for _ in range(max(n // 50, 4))]
    raw += [b"", b"a", b"A", b"  a  b  ", b" ", b"\t", b"   ", "中".encode(), "é".encode(), b"x" * 31 + "é".encode(),
            b"y" * 30 + "中中".encode(), b"z" * 29 + "\U0001F600".encode() + b"!", b"Hello World!", "Ю Σ".encode()]
    for _ in range(malformed):
        b = bytearray(rng.choice(raw) or b"x")
        for _ in range(rng.randint(1, 3)):
            b.insert(rng.randint(0, len(b)), rng.choice([0x80, 0xC0, 0xE2, 0xF0, 0xFF, 0xED, 0xA0, 0xC1]))
        raw.append(bytes(b))
    return raw


def pack(raw):
    ends = np.cumsum([len(r) for r in raw]).astype(np.int32) if raw else np.zeros(0, np.int32)
    begins = np.concatenate([[0], ends[:-1]]).astype(np.int32) if raw else np.zeros(0, np.int32)
    return begins, ends, np.frombuffer(b"".join(raw), np.uint8)


def unpack(b, e, c):
    c = np.asarray(c, np.uint8).tobytes()
    return [c[int(x):int(y)] for x, y in zip(b, e)]
