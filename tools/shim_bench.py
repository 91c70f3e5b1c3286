"""Development probe: device-resident timing of the byte-level shims (BytesToChars, CharsToBytes, UTF8Validate) on a
65 536 x 512 B batch (ASCII or mixed UTF-8)."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import cases
from openvino_tokenizers_b200 import _capi as K

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
L = int(sys.argv[2]) if len(sys.argv) > 2 else 512
kind = sys.argv[3] if len(sys.argv) > 3 else "ascii"
rb, re_, b, e, c = cases.random_ascii_batch(B, L) if kind == "ascii" else cases.mixed_utf8_batch(B, L)
N = c.size
dev = torch.device("cuda:0")
d = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (rb, re_, b, e)]
dc = torch.from_numpy(np.concatenate([c, np.zeros(64, np.uint8)])).to(dev)
cap = 3 * N + 64
ob = torch.empty(B, dtype=torch.int32, device=dev); oe = torch.empty_like(ob); oc = torch.empty(cap, dtype=torch.uint8, device=dev)
lib = K.lib()
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); st = C.c_void_p(stream.cuda_stream)
rin = K.RaggedStrings(d[0].data_ptr(), d[1].data_ptr(), B, d[2].data_ptr(), d[3].data_ptr(), B, dc.data_ptr(), N, None, K.MEM_DEVICE)
got = C.c_int64(0)
P = lambda t: C.c_void_p(t.data_ptr())


def timeit(name, fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(10):
        fn()
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 10
    print(f"{name:16s} {ms:.3f} ms/call  {N / 1e6 / (ms / 1e3):9.1f} MB/s text  out {got.value} B")


timeit("BytesToChars", lambda: K.check(lib.b200tok_bytes_to_chars_run(0, C.byref(rin), P(ob), P(oe), P(oc), C.c_int64(cap), C.byref(got), st)))
timeit("UTF8Validate", lambda: K.check(lib.b200tok_utf8_validate_run(0, P(d[2]), P(d[3]), C.c_int64(B), P(dc), C.c_int64(N), 1, P(ob), P(oe), P(oc), C.c_int64(cap), C.byref(got), K.MEM_DEVICE, st)))
# CharsToBytes on the BytesToChars output
K.check(lib.b200tok_bytes_to_chars_run(0, C.byref(rin), P(ob), P(oe), P(oc), C.c_int64(cap), C.byref(got), st))
torch.cuda.synchronize()
n2 = got.value
b2, e2, c2 = ob.clone(), oe.clone(), oc[: n2 + 64].clone()
rin2 = K.RaggedStrings(d[0].data_ptr(), d[1].data_ptr(), B, b2.data_ptr(), e2.data_ptr(), B, c2.data_ptr(), n2, None, K.MEM_DEVICE)
timeit("CharsToBytes", lambda: K.check(lib.b200tok_chars_to_bytes_run(0, C.byref(rin2), P(ob), P(oe), P(oc), C.c_int64(cap), C.byref(got), st)))
# short elements (the realistic shape for BytesToChars / CharsToBytes: pieces and per-token strings): 8-byte elements
E8 = N // 8
b8 = torch.arange(0, E8, dtype=torch.int32, device=dev) * 8
e8 = b8 + 8
r8b = torch.arange(0, E8, dtype=torch.int32, device=dev); r8e = r8b + 1
ob8 = torch.empty(E8, dtype=torch.int32, device=dev); oe8 = torch.empty_like(ob8)
rin8 = K.RaggedStrings(r8b.data_ptr(), r8e.data_ptr(), E8, b8.data_ptr(), e8.data_ptr(), E8, dc.data_ptr(), N, None, K.MEM_DEVICE)
timeit("B2C 8-byte elems", lambda: K.check(lib.b200tok_bytes_to_chars_run(0, C.byref(rin8), P(ob8), P(oe8), P(oc), C.c_int64(cap), C.byref(got), st)))
