"""Parity of one row-loop variant of the fast kernel (selected by environment variables that the library reads once, hence a
process of its own): mixed fast / handed-back / multi-window rows through the device-resident and the host-buffer path against
the CPU oracle.  Used by tests/test_gpu_parity.py::test_fast_kernel_variants.  Exit code 0 = identical."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import cases  # noqa: E402
import oracle  # noqa: E402
import test_gpu_parity as T  # noqa: E402
from openvino_tokenizers_b200 import assets as A  # noqa: E402
from openvino_tokenizers_b200 import ops  # noqa: E402
from openvino_tokenizers_b200 import runtime as R  # noqa: E402

bad = 0
for name in ("gpt2_synth", "llama3_synth"):
    a = A.load_bpe(name)
    v, ml, mr, ad, aid = a.tensors()
    consts = [*v, *ml, *mr] + ([*ad, aid] if ad is not None else [])
    m = dict(o_split=oracle.SplitOracle(a.split_pattern, "isolate"), o_bpe=oracle.BpeOracle(v, ml, mr, ad, aid))
    split, bpe = ops.RegexSplit("isolate").with_pattern(a.split_pattern), ops.BPETokenizer().with_constants(consts)
    pipe = R.TokenizerPipeline("bpe", name)
    rng = np.random.default_rng(3)
    batches = [cases.batch_from_strings(T._mixed_rows(rng, 500)), cases.batch_from_strings(cases.EDGE_STRINGS + cases.long_prompts()),
               cases.random_ascii_batch(2048, 512, seed=4), cases.mixed_utf8_batch(768, 1024, seed=5),
               cases.batch_from_strings([bytes(rng.integers(0x20, 0x7F, size=int(n), dtype=np.uint8)).decode() for n in rng.integers(0, 1300, size=900)])]
    for i, batch in enumerate(batches):
        exp = T.oracle_chain_bpe(m, batch, threads=T.host_threads())
        for tag, got in (("device", T._device_run(pipe, batch)), ("host", ops.split_bpe(split, bpe, list(batch))), ("device again", T._device_run(pipe, batch))):
            if not cases.ragged_rows_equal(got, exp):
                print(f"MISMATCH {name} batch {i} {tag}")
                bad += 1
print("variant ok" if not bad else f"{bad} mismatches")
sys.exit(1 if bad else 0)
