"""Development probe: time the fused split+BPE path on C1-shaped input (device-resident and host-to-host)."""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import cases
from openvino_tokenizers_b200 import _capi as K
from openvino_tokenizers_b200 import assets as A
from openvino_tokenizers_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
L = int(sys.argv[2]) if len(sys.argv) > 2 else 512
kind = sys.argv[3] if len(sys.argv) > 3 else "ascii"
a = A.load_bpe("gpt2_synth")
v, ml, mr, ad, aid = a.tensors()
consts = [*v, *ml, *mr] + ([*ad, aid] if ad is not None else [])
split = ops.RegexSplit("isolate").with_pattern(a.split_pattern)
bpe = ops.BPETokenizer().with_constants(consts)
batch = cases.random_ascii_batch(B, L) if kind == "ascii" else cases.english_like_batch(B, L)
rb, re_, b, e, c = batch
NOHOST = "nohost" in sys.argv
got = None
if not NOHOST:
    t0 = time.time(); got = ops.split_bpe(split, bpe, list(batch)); t1 = time.time()
    print("first host call", t1 - t0, "s; tokens", len(got[2]))
for _ in range(0 if NOHOST else 3):
    t0 = time.time(); got = ops.split_bpe(split, bpe, list(batch)); t1 = time.time()
    print("host-to-host MB/s", B * L / 1e6 / (t1 - t0))
dev = torch.device("cuda:0")
d = [torch.from_numpy(x).to(dev) for x in (rb, re_, b, e)]
dc = torch.from_numpy(np.concatenate([c, np.zeros(64, np.uint8)])).to(dev)
ob = torch.empty(B, dtype=torch.int32, device=dev); oe = torch.empty(B, dtype=torch.int32, device=dev)
ids = torch.empty(B * L, dtype=torch.int32, device=dev); nid = torch.zeros(1, dtype=torch.int64, device=dev)
rin = K.RaggedStrings(d[0].data_ptr(), d[1].data_ptr(), B, d[2].data_ptr(), d[3].data_ptr(), B, dc.data_ptr(), B * L, None, K.MEM_DEVICE)
out = K.RaggedIds(ob.data_ptr(), oe.data_ptr(), ids.data_ptr(), B * L, 0, nid.data_ptr(), K.MEM_DEVICE)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
st = stream.cuda_stream
lib = K.lib()
for _ in range(3):
    K.check(lib.b200tok_split_bpe_run(split.handle, bpe.handle, C.byref(rin), C.byref(out), C.c_void_p(st)))
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K_STEPS = 10
ev0.record()
for _ in range(K_STEPS):
    K.check(lib.b200tok_split_bpe_run(split.handle, bpe.handle, C.byref(rin), C.byref(out), C.c_void_p(st)))
ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / K_STEPS
T = int(nid.item())
print(f"device-resident: {ms:.3f} ms/step, {B*L/1e6/(ms/1e3):.1f} MB/s text, T={T}, algorithmic GB/s {(B*L+16*B+4*T)/1e9/(ms/1e3):.1f}")
assert got is None or np.array_equal(ids[:T].cpu().numpy(), got[2])
