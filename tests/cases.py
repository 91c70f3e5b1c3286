"""Shared test inputs: the edge-case corpus (short strings after reference tests/tokenizers_test.py:27-75:
English with punctuation/digits/whitespace runs, multilingual, emoji incl. ZWJ sequences, empty / control /
256 spaces) and the seeded synthetic batches of BASELINE.md (C0..C4)."""
from __future__ import annotations

from pathlib import Path

import numpy as np

from openvino_tokenizers_b200.strings import add_ragged_dimension, pack_strings

ROOT = Path(__file__).resolve().parent.parent

EDGE_STRINGS = [
    "Eng... test, string?!",
    "Multiline\nstring!\nWow!",
    "A lot\t w!",
    "A lot\t\tof whitespaces!",
    "\n\n\n\t\t   A    lot\t\tof\twhitespaces\n!\n\n\n\t\n\n",
    "Eng, but with d1gits: 123; 0987654321, stop.0987654321 - eng, but with d1gits: 123",
    "USER: <image>\nWhat is in the image? ASSISTANT:",
    "What is OpenVINO?",
    "If I have 100 million dollars, what kinds of projects should I invest to maximize my benefits?",
    "He'll say it's fine, they've said we're done; I'd go, I'm sure you'll see. DON'T SHOUT 'LL 'RE",
    "Тестовая строка!",
    "Testzeichenfolge?",
    "Tester, la chaîne...",
    "測試字符串",
    "سلسلة الاختبار",
    "מחרוזת בדיקה",
    "Сынақ жолы á",
    "رشته تست",
    "介绍下清华大学",
    "若我有一亿美元，在人工智能盛行的今天，我怎样投资才能收益最大化？",
    "😀",
    "😁😁",
    "🤣🤣🤣😁😁😁😁",
    "🫠",
    "🤷‍♂️",
    "🤦🏼‍♂️",
    "",
    "\x06",
    " ",
    " " * 10,
    " " * 256,
    "\n",
    " \t\n",
    "<|endoftext|>",
    "a<|endoftext|>b <|endoftext",
    "x y  z　w v",
    "tab\tnew\r\nline\r\n\r\n  indent",
    "1234567890 12 345 6789٣٤",
    "'s't're've'm'll'd 'S 'ſ",
    "a" * 300,
    "ab" * 700,
    " " * 2000 + "x",
    "word " * 400,
    "!?" * 520,
    "9" * 1500,
]


def long_prompts():
    """A few multi-KB English prompts (the reference corpus has three); cut from this repo's SURVEY.md."""
    text = (ROOT / "SURVEY.md").read_text(encoding="utf-8")
    return [text[0:4096], text[10000:13000], text[20000:28192]]


def batch_from_strings(strings):
    b, e, c = pack_strings(strings)
    rb, re_ = add_ragged_dimension(b, e)
    return rb, re_, b, e, c


def uniform_batch(chars: np.ndarray, B: int, L: int):
    b = (np.arange(B, dtype=np.int64) * L).astype(np.int32)
    e = b + np.int32(L)
    rb, re_ = add_ragged_dimension(b, e)
    return rb, re_, b, e, chars


def random_ascii_batch(B: int, L: int, seed: int = 1234, lower: bool = False):
    """C0/C1/C2: bytes uniform on printable ASCII 0x20..0x7E."""
    rng = np.random.default_rng(seed)
    chars = rng.integers(0x20, 0x7F, size=B * L, dtype=np.uint8)
    if lower:
        up = (chars >= 0x41) & (chars <= 0x5A)
        chars = np.where(up, chars + 32, chars).astype(np.uint8)
    return uniform_batch(chars, B, L)


def mixed_utf8_batch(B: int, L: int, seed: int = 1234):
    """C3: 70 % printable ASCII / 15 % U+0400-04FF / 10 % U+4E00-9FFF / 5 % U+1F600-1F64F, rows of exactly L bytes
    (tail padded with 'a'); always valid UTF-8."""
    rng = np.random.default_rng(seed)
    n_cp = B * L  # upper bound on code points needed
    kind = rng.choice(4, size=n_cp, p=[0.70, 0.15, 0.10, 0.05]).astype(np.uint8)
    u = rng.random(n_cp)
    cp = np.where(kind == 0, 0x20 + (u * 95).astype(np.int64),
                  np.where(kind == 1, 0x400 + (u * 256).astype(np.int64),
                           np.where(kind == 2, 0x4E00 + (u * (0x9FFF - 0x4E00 + 1)).astype(np.int64),
                                    0x1F600 + (u * 0x50).astype(np.int64))))
    nbytes = np.where(kind == 0, 1, np.where(kind == 1, 2, np.where(kind == 2, 3, 4))).astype(np.int64)
    out = np.full(B * L, ord("a"), dtype=np.uint8)
    # greedy fill per row, vectorised across rows: walk code points with a running cursor per row
    cum = np.cumsum(nbytes)
    start = 0
    row_starts = np.zeros(B, dtype=np.int64)
    pos = 0
    idx = 0
    # sequential over rows but vectorised inside via searchsorted
    base = 0
    for r in range(B):
        # take code points idx.. while they fit in L bytes
        lim = base + L
        j = int(np.searchsorted(cum, lim, side="right"))
        n_take = j - idx
        if n_take > 0:
            offs = (cum[idx:j] - nbytes[idx:j]) - base + r * L
            cps = cp[idx:j]
            nb = nbytes[idx:j]
            m1 = nb == 1
            out[offs[m1]] = cps[m1]
            m2 = nb == 2
            out[offs[m2]] = 0xC0 | (cps[m2] >> 6)
            out[offs[m2] + 1] = 0x80 | (cps[m2] & 63)
            m3 = nb == 3
            out[offs[m3]] = 0xE0 | (cps[m3] >> 12)
            out[offs[m3] + 1] = 0x80 | ((cps[m3] >> 6) & 63)
            out[offs[m3] + 2] = 0x80 | (cps[m3] & 63)
            m4 = nb == 4
            out[offs[m4]] = 0xF0 | (cps[m4] >> 18)
            out[offs[m4] + 1] = 0x80 | ((cps[m4] >> 12) & 63)
            out[offs[m4] + 2] = 0x80 | ((cps[m4] >> 6) & 63)
            out[offs[m4] + 3] = 0x80 | (cps[m4] & 63)
        idx = j
        base = int(cum[j - 1]) if j > 0 else 0
    return uniform_batch(out, B, L)


def english_like_batch(B: int, L: int, seed: int = 7):
    """Rows of L bytes cut from repeated real text (this repo's SURVEY.md) at seeded offsets."""
    text = (ROOT / "SURVEY.md").read_bytes()
    text = bytes(b if b < 0x80 else 0x20 for b in text)  # ASCII only so that any cut is valid UTF-8
    rng = np.random.default_rng(seed)
    offs = rng.integers(0, len(text) - L, size=B)
    arr = np.frombuffer(text, np.uint8)
    chars = np.concatenate([arr[o:o + L] for o in offs])
    return uniform_batch(chars, B, L)


def ragged_rows_equal(a, b):
    """Compare two ragged id results (begins, ends, ids)."""
    return (np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]))
