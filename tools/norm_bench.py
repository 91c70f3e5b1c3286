"""Development probe: time the normaliser ops device-resident on the C2 batch shape (65 536 x 256 B of ASCII with
some tabs / controls) and on a mixed-Unicode batch; prints ms per call (two kernels + a cub scan + one 8-byte D2H)."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import normcases as NC
from openvino_tokenizers_b200 import _capi as K

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
L = int(sys.argv[2]) if len(sys.argv) > 2 else 256
kind = sys.argv[3] if len(sys.argv) > 3 else "ascii"
rng = np.random.default_rng(1234)
if kind == "ascii":
    chars = rng.integers(0x20, 0x7F, size=B * L, dtype=np.uint8)
    r = rng.random(B * L)
    chars[r < 0.01] = 0x09
    chars[(r >= 0.01) & (r < 0.015)] = 0x01
    b = np.arange(B, dtype=np.int32) * L
    e = b + L
else:
    raw = NC.corpus(seed=9, n=B, max_len=L // 2)[:B]
    b, e, chars = NC.pack(raw)
    chars = chars.copy()
N = chars.size
dev = torch.device("cuda:0")
db, de = torch.from_numpy(b).to(dev), torch.from_numpy(e).to(dev)
dc = torch.from_numpy(np.concatenate([chars, np.zeros(64, np.uint8)])).to(dev)
cap = 3 * N + 64
ob = torch.empty(len(b), dtype=torch.int32, device=dev); oe = torch.empty_like(ob); oc = torch.empty(cap, dtype=torch.uint8, device=dev)
lib = K.lib()
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); st = C.c_void_p(stream.cuda_stream)


def handle(kind_, *a):
    h = C.c_void_p()
    if kind_ == "regex":
        s, r, g = a
        K.check(lib.b200tok_regexnorm_create(s, C.c_int64(len(s)), r, C.c_int64(len(r)), g, 0, C.byref(h)))
    else:
        K.check(lib.b200tok_charsmap_create(a[0], C.c_int64(len(a[0])), 0, 0, 0, 0, C.byref(h)))
    return h


ops = [("del_control", handle("regex", rb"([\x00-\x08\x0B\x0C\x0E-\x1F\x7F-\x9F\p{Cf}])", b"", 1)),
       ("whitespace", handle("regex", rb"\s", b" ", 1)),
       ("han", handle("regex", rb"([\p{Han}])", b" $1 ", 1)),
       ("nfd", handle("charsmap", NC.unicodedata_blob("NFD", False))),
       ("strip_accents", handle("regex", rb"\p{Mn}", b"", 1)),
       ("casefold", handle("charsmap", NC.unicodedata_blob(None, True)))]
got = C.c_int64(0)
for name, h in ops:
    def run():
        K.check(lib.b200tok_normalize_run(h, C.c_void_p(db.data_ptr()), C.c_void_p(de.data_ptr()), C.c_int64(len(b)), C.c_void_p(dc.data_ptr()), C.c_int64(N),
                                          None, C.c_void_p(ob.data_ptr()), C.c_void_p(oe.data_ptr()), C.c_void_p(oc.data_ptr()), C.c_int64(cap), C.byref(got),
                                          K.MEM_DEVICE, st))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(10):
        run()
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 10
    alg = 2 * N + got.value + 16 * len(b)     # the text is read by both passes, the result written once, offsets in / out
    print(f"{name:14s} {ms:.3f} ms/call  {N / 1e6 / (ms / 1e3):9.1f} MB/s text  algorithmic {alg / 1e9 / (ms / 1e3):7.1f} GB/s  out {got.value} B")

hs = (C.c_void_p * len(ops))(*[h for _, h in ops])


def run_chain():
    K.check(lib.b200tok_normalize_chain_run(hs, len(ops), C.c_void_p(db.data_ptr()), C.c_void_p(de.data_ptr()), C.c_int64(len(b)), C.c_void_p(dc.data_ptr()),
                                            C.c_int64(N), None, C.c_void_p(ob.data_ptr()), C.c_void_p(oe.data_ptr()), C.c_void_p(oc.data_ptr()), C.c_int64(cap),
                                            C.byref(got), K.MEM_DEVICE, st))


for _ in range(3):
    run_chain()
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(10):
    run_chain()
ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / 10
print(f"BERT chain (6 ops, one call) {ms:.3f} ms  {N / 1e6 / (ms / 1e3):9.1f} MB/s text")
