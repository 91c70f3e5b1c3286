"""Development soak: the scan kernels (both forms) against the oracle on freshly seeded corpora — more chunk alignments and
string mixes than the fixed-seed tests.  usage: python tools/norm_soak.py [first_seed] [n_seeds]"""
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import normcases as NC
import oracle
from openvino_tokenizers_b200 import ops

G = json.loads((ROOT / "tests" / "golden" / "normalization_layer_tests.json").read_text())
steps = G["bert_steps"] + G["other_steps"]
first, count = int(sys.argv[1]) if len(sys.argv) > 1 else 100, int(sys.argv[2]) if len(sys.argv) > 2 else 6
blobs = [NC.unicodedata_blob("NFD", True), NC.builtin_blob("nfkc_cf"), NC.custom_blob()]
bad = 0
for seed in range(first, first + count):
    rng = np.random.default_rng(seed)
    raw = NC.corpus(seed=seed, n=1200, malformed=0, max_len=int(rng.integers(20, 200))) + NC.ascii_corpus(seed=seed + 1, n=600, max_len=300)
    rawm = raw + NC.corpus(seed=seed + 2, n=300, malformed=300)[-300:]
    for path in ("warp", "thread"):
        os.environ["B200TOK_NORM_PATH"] = path
        b, e, c = NC.pack(raw)
        for s in steps:
            exp = oracle.regex_normalize(s["search"], s["replace"], s["global_replace"], b, e, c)
            got = ops.RegexNormalization(s["global_replace"]).evaluate([b, e, c, np.frombuffer(s["search"].encode(), np.uint8), np.frombuffer(s["replace"].encode(), np.uint8)])
            if NC.unpack(*got[:3]) != NC.unpack(*exp):
                bad += 1; print("MISMATCH regex", s["name"], seed, path)
        b, e, c = NC.pack(rawm)
        for blob in blobs:
            exp = oracle.charsmap_normalize(blob, b, e, c)
            got = ops.CharsMapNormalization(precompiled_charsmap=blob).evaluate([b, e, c])
            if NC.unpack(*got[:3]) != NC.unpack(*exp):
                bad += 1; print("MISMATCH charsmap", seed, path)
        for mode in (False, True):
            exp = oracle.utf8_validate(b, e, c, mode)
            got = ops.UTF8Validate(mode).evaluate([b, e, c])
            if not all(np.array_equal(x, y) for x, y in zip(got, exp)):
                bad += 1; print("MISMATCH utf8", seed, path, mode)
print("soak done, mismatches:", bad)
sys.exit(1 if bad else 0)
