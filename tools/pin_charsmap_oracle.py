#!/usr/bin/env python
"""Pin oracle.charsmap_normalize (the restated sentencepiece Normalizer) against the real thing:
the installed sentencepiece 0.2.1 (the version the reference pins, src/CMakeLists.txt:77) driven with the reference's own
precompiled charsmaps (parsed out of /root/reference/src/precompiled_charsmap.hpp at run time; nothing is copied) and
with sentencepiece's built-in rule sets.  Runs only where /root/reference exists (this container).

    python tools/pin_charsmap_oracle.py            # prints mismatch counts; exit code 1 on any mismatch
"""
import random
import re
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402

REF_HEADER = Path("/root/reference/src/precompiled_charsmap.hpp")


def reference_blob(name, _cache={}):
    if "src" not in _cache:
        _cache["src"] = REF_HEADER.read_bytes().decode("latin1")
    m = re.search(r"const std::string precompiled_charsmap_%s = std::string\((.*?), (\d+)\);" % name, _cache["src"], re.S)
    body, n = m.group(1), int(m.group(2))
    out = bytearray()
    for lit in re.findall(r'"((?:[^"\\]|\\.)*)"', body):
        i = 0
        while i < len(lit):
            if lit[i] == "\\":
                out.append(int(lit[i + 2:i + 4], 16))
                i += 4
            else:
                out.append(ord(lit[i]))
                i += 1
    assert len(out) == n
    return bytes(out)


def builtin_blob(rule_name, _cache={}):
    """A charsmap compiled into the sentencepiece library itself: train a throw-away model with that rule and read the blob back."""
    if rule_name not in _cache:
        import io
        import sentencepiece as spm
        from sentencepiece import sentencepiece_model_pb2 as pb
        buf = io.BytesIO()
        words = ["alpha", "beta", "gamma", "delta", "epsilon", "zeta", "eta", "theta", "iota", "kappa", "lambda", "mu"]
        sents = [" ".join(words[(i + j) % len(words)] for j in range(6)) for i in range(200)]
        spm.SentencePieceTrainer.train(sentence_iterator=iter(sents), model_writer=buf, vocab_size=30, model_type="unigram", hard_vocab_limit=False,
                                       normalization_rule_name=rule_name, minloglevel=2)
        mp = pb.ModelProto()
        mp.ParseFromString(buf.getvalue())
        _cache[rule_name] = bytes(mp.normalizer_spec.precompiled_charsmap)
    return _cache[rule_name]


def sp_normalizer(blob, adp=False, rew=False, esc=False):
    import sentencepiece as spm
    from sentencepiece import sentencepiece_model_pb2 as pb
    mp = pb.ModelProto()
    mp.normalizer_spec.precompiled_charsmap = blob
    mp.normalizer_spec.add_dummy_prefix = adp
    mp.normalizer_spec.remove_extra_whitespaces = rew
    mp.normalizer_spec.escape_whitespaces = esc
    for p, t in [("<unk>", 2), ("<s>", 3), ("</s>", 3), ("a", 1)]:
        x = mp.pieces.add()
        x.piece, x.score, x.type = p, 0.0, t
    return spm.SentencePieceNormalizer(model_proto=mp.SerializeToString(), add_dummy_prefix=adp, escape_whitespaces=esc,
                                       remove_extra_whitespaces=rew)


def corpus(seed=5, n=3000, malformed=500):
    rng = random.Random(seed)
    special = [0x300, 0x301, 0x308, 0x327, 0x1100, 0x1161, 0xAC00, 0xFB01, 0x2126, 0x212B, 0x1E9B, 0x323]
    other = [0x4E2D, 0x1F600, 0x3042, 0xFF21, 0x2460, 0xDF, 0x130, 0x1E9E]

    def rand_str():
        out = []
        for _ in range(rng.randint(0, 40)):
            r = rng.random()
            if r < .4: out.append(chr(rng.randint(0x20, 0x7E)))
            elif r < .5: out.append(" ")
            elif r < .7: out.append(chr(rng.randint(0xA0, 0x24F)))
            elif r < .8: out.append(chr(rng.choice(special)))
            elif r < .9: out.append(chr(rng.randint(0x370, 0x52F)))
            else: out.append(chr(rng.choice(other)))
        return "".join(out)

    raw = [rand_str().encode() for _ in range(n)] + [b"", b"a", b"  a  b  ", b" ", b"\t", b"   "]
    for _ in range(malformed):
        b = bytearray(rng.choice(raw) or b"x")
        for _ in range(rng.randint(1, 3)):
            b.insert(rng.randint(0, len(b)), rng.choice([0x80, 0xC0, 0xE2, 0xF0, 0xFF, 0xED, 0xA0, 0xC1]))
        raw.append(bytes(b))
    return raw


def pack(raw):
    ends = np.cumsum([len(r) for r in raw]).astype(np.int32)
    begins = np.concatenate([[0], ends[:-1]]).astype(np.int32)
    return begins, ends, np.frombuffer(b"".join(raw), np.uint8)


def main():
    raw = corpus()
    b, e, c = pack(raw)
    blobs = {"builtin:" + r: builtin_blob(r) for r in ("nfkc", "nfkc_cf", "nmt_nfkc", "nmt_nfkc_cf")}
    if REF_HEADER.exists():
        blobs.update({"reference:" + r: reference_blob(r) for r in ("casefold", "nfd", "nfd_casefold", "nfc", "nfkc", "nfkd_casefold")})
    total_bad = 0
    for name, blob in blobs.items():
        for kw in ({}, {"adp": True, "esc": True}, {"rew": True}, {"adp": True, "rew": True, "esc": True}):
            n = sp_normalizer(blob, **kw)
            ob, oe, oc = oracle.charsmap_normalize(blob, b, e, c, add_dummy_prefix=kw.get("adp", False),
                                                   remove_extra_whitespaces=kw.get("rew", False), escape_whitespaces=kw.get("esc", False))
            bad = 0
            for i, r in enumerate(raw):
                exp = n.normalize(r)
                if isinstance(exp, str):
                    exp = exp.encode("utf-8", "surrogatepass")
                got = bytes(oc[ob[i]:oe[i]])
                if exp != got:
                    bad += 1
                    if bad < 3:
                        print("  ", name, kw, repr(r)[:80], repr(exp)[:80], repr(got)[:80])
            print(f"{name:28s} {str(kw):50s} mismatches {bad} of {len(raw)}")
            total_bad += bad
    return 1 if total_bad else 0


if __name__ == "__main__":
    sys.exit(main())
