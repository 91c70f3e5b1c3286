"""Device-resident and pinned-host drivers over the C ABI, using torch only for memory and streams.

`TokenizerPipeline` bundles the handles of one converted tokenizer (RegexSplit + BPETokenizer or
RegexSplit x2 + WordpieceTokenizer) and exposes the two ways the hot path is driven:

  * run_device(batch)  — inputs already in HBM, outputs stay in HBM, fully asynchronous on the current torch stream
                         (b200tok_ragged_ids.n_ids_device): what `value` / the roofline measure.
  * run_host(batch)    — host buffers in, host buffers out through B200TOK_MEM_HOST, i.e. what an
                         ov::Op::evaluate() shim calls; with pinned buffers the copies run at PCIe speed: `e2e`.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _capi as K
from . import assets as A
from . import ops
from .strings import pack_strings


@dataclass
class DeviceBatch:
    rb: torch.Tensor
    re: torch.Tensor
    begins: torch.Tensor
    ends: torch.Tensor
    chars: torch.Tensor     # padded by >= 64 bytes
    n_chars: int

    @property
    def n_rows(self):
        return self.rb.numel()

    @property
    def n_elems(self):
        return self.begins.numel()


def to_device(batch, device) -> DeviceBatch:
    rb, re_, b, e, c = batch
    pad = np.zeros(64, np.uint8)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(device)
    return DeviceBatch(t(rb), t(re_), t(b), t(e), t(np.concatenate([c, pad])), int(len(c)))


def split_rows(batch, parts: int):
    """Cut a host batch (rb, re, begins, ends, chars) with one element per row and contiguous chars into `parts` row blocks
    (each rebased to its own chars slice) — the unit the multi-GPU step pipelines: tokenise block k+1 while block k is gathered."""
    rb, re_, b, e, c = batch
    n = len(rb)
    out = []
    for k in range(parts):
        lo, hi = n * k // parts, n * (k + 1) // parts
        p0, p1 = int(rb[lo]), int(re_[hi - 1])
        c0, c1 = int(b[p0]), int(e[p1 - 1])
        out.append((rb[lo:hi] - p0, re_[lo:hi] - p0, b[p0:p1] - c0, e[p0:p1] - c0, c[c0:c1]))
    return out


def to_pinned(batch):
    rb, re_, b, e, c = batch
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
    return [t(rb), t(re_), t(b), t(e), t(c)]


class TokenizerPipeline:
    def __init__(self, kind: str, name: str, device: int = 0):
        self.kind, self.device = kind, device
        if kind == "bpe":
            a = A.load_bpe(name)
            v, ml, mr, ad, aid = a.tensors()
            consts = [*v, *ml, *mr] + ([*ad, aid] if ad is not None else [])
            self.assets = a
            self.tok = ops.BPETokenizer(device=device).with_constants(consts)
            self.split1 = ops.RegexSplit("isolate", device=device).with_pattern(a.split_pattern)
            self.split2 = None
            self.unk = 0
        elif kind == "wordpiece":
            a = A.load_wordpiece(name)
            self.assets = a
            self.tok = ops.WordpieceTokenizer(a.suffix_indicator, a.max_bytes_per_word, device=device).with_constants(pack_strings(a.vocab))
            self.split1 = ops.RegexSplit("remove", device=device).with_pattern(A.BERT_WHITESPACE_PATTERN)
            self.split2 = ops.RegexSplit("isolate", device=device).with_pattern(A.BERT_PUNCT_PATTERN)
            self.unk = a.unk_token_id
        else:
            raise ValueError(kind)
        self._out = None

    # -- bookkeeping ---------------------------------------------------------------------------
    @property
    def launches(self) -> int:
        return self.tok.launches

    @property
    def dominant_kernel(self) -> str:
        """Name of the kernel b200tok_last_kernel_ms() times (csrc/api.cu launch_chunk)."""
        if self.kind == "bpe" and self.assets.split_pattern in (A.GPT2_PATTERN, A.GPT2_DIGITS_PATTERN, A.LLAMA3_PATTERN) and not self.assets.end_suffix:
            pat = "llama3" if self.assets.split_pattern == A.LLAMA3_PATTERN else "gpt2"
            return "gpt2_bpe_fast_kernel" + ("<u16,5," if len(self.assets.vocab) < 0xFFFF else "<i32,4,") + pat + ">"
        return f"rows_kernel<{self.kind}>"

    def set_timing(self, on: bool):
        K.lib().b200tok_set_timing(self.tok.handle, int(on))

    def last_kernel_ms(self) -> float:
        return float(K.lib().b200tok_last_kernel_ms(self.tok.handle))

    def _call(self, rin, out, stream):
        L = K.lib()
        if self.kind == "bpe":
            K.check(L.b200tok_split_bpe_run(self.split1.handle, self.tok.handle, C.byref(rin), C.byref(out), stream))
        else:
            K.check(L.b200tok_split_wordpiece_run(self.split1.handle, self.split2.handle, self.tok.handle, C.byref(rin),
                                                  C.c_int32(self.unk), C.byref(out), stream))

    # -- device-resident -------------------------------------------------------------------------
    def alloc_device_out(self, n_rows: int, capacity: int):
        dev = torch.device("cuda", self.device)
        self._out = dict(begins=torch.empty(n_rows, dtype=torch.int32, device=dev),
                         ends=torch.empty(n_rows, dtype=torch.int32, device=dev),
                         ids=torch.empty(capacity, dtype=torch.int32, device=dev),
                         n=torch.zeros(1, dtype=torch.int64, device=dev), cap=capacity)
        return self._out

    def run_device(self, db: DeviceBatch, out=None):
        """Asynchronous on the current torch stream; results in `out` (from alloc_device_out) or self._out (ids[:n])."""
        o = out if out is not None else self._out
        if o is None or o["begins"].numel() != db.n_rows or o["cap"] < db.n_chars + db.n_elems:
            o = self.alloc_device_out(db.n_rows, db.n_chars + db.n_elems)
        rin = K.RaggedStrings(db.rb.data_ptr(), db.re.data_ptr(), db.n_rows, db.begins.data_ptr(), db.ends.data_ptr(),
                              db.n_elems, db.chars.data_ptr(), db.n_chars, None, K.MEM_DEVICE)
        out = K.RaggedIds(o["begins"].data_ptr(), o["ends"].data_ptr(), o["ids"].data_ptr(), o["cap"], 0,
                          o["n"].data_ptr(), K.MEM_DEVICE)
        self._call(rin, out, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        return o

    # -- host to host ----------------------------------------------------------------------------
    def alloc_host_out(self, n_rows: int, capacity: int, pinned: bool = True):
        mk = (lambda n: torch.empty(n, dtype=torch.int32).pin_memory()) if pinned else (lambda n: torch.empty(n, dtype=torch.int32))
        return dict(begins=mk(n_rows), ends=mk(n_rows), ids=mk(capacity), cap=capacity)

    def run_host(self, hb, ho):
        """hb: list of 5 host tensors (rb, re, begins, ends, chars); ho: dict from alloc_host_out.  Synchronous."""
        rin = K.RaggedStrings(hb[0].data_ptr(), hb[1].data_ptr(), hb[0].numel(), hb[2].data_ptr(), hb[3].data_ptr(),
                              hb[2].numel(), hb[4].data_ptr(), hb[4].numel(), None, K.MEM_HOST)
        out = K.RaggedIds(ho["begins"].data_ptr(), ho["ends"].data_ptr(), ho["ids"].data_ptr(), ho["cap"], 0, None, K.MEM_HOST)
        self._call(rin, out, None)
        return int(out.n_ids)
