"""Development probe: timeline of the host-buffer pipeline of one C1 call (B200TOK_PIPE_TRACE=1 makes the library print when each
chunk's kernels were seen finished by the host and when the call ended).  python tools/pipe_trace.py [plan]   e.g. "2,2,4" (MiB per chunk)"""
import os
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
os.environ["B200TOK_PIPE_TRACE"] = "1"
if len(sys.argv) > 1:
    os.environ["B200TOK_PIPE_PLAN"] = sys.argv[1]
import cases
from openvino_tokenizers_b200 import runtime as R
pipe = R.TokenizerPipeline("bpe", "gpt2_synth")
batch = cases.random_ascii_batch(65536, 512, 1234)
hb = R.to_pinned(batch)
ho = pipe.alloc_host_out(65536, 65536 * 512 + 65536)
for _ in range(3):
    pipe.run_host(hb, ho)
t0 = time.perf_counter()
for _ in range(5):
    n = pipe.run_host(hb, ho)
print(f"plan {os.environ.get('B200TOK_PIPE_PLAN', 'default 2,2,4')}: {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms per call, {n} ids")
