#!/usr/bin/env python
"""bench.py — MB/s of input text tokenized (bit-exact ids) on the BASELINE.json headline workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c1|c2|c3|c4|norm]

A "step" is one pass of the fused RegexSplit -> BPETokenizer hot path over one synthetic batch
(C1: gpt2-shaped BPE, 65 536 x 512-byte printable-ASCII docs per GPU; weak scaling: every rank tokenises its
own 65 536-row shard and, for N > 1, the ragged id rows are all-gathered over NCCL).
  value       whole-job MB/s with inputs resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e         the same metric through the host-buffer C-ABI call an ov::Op::evaluate() shim makes
              (pinned host buffers; H2D and D2H copies inside the timed region)
  roofline    dominant kernel (gpt2_bpe_fast_kernel on C1/C3, rows_kernel<wordpiece> on C2) algorithmic bytes / its CUDA-event
              duration vs the measured HBM peak
  cpu_baseline  the reference's own op code (oracle/_ref: src/*.cpp of the reference compiled against a stand-in OpenVINO API;
                the oracle port where that library is absent), 1 thread like the reference's serial evaluate(), on a bounded sample
After the timed regions the ids of the timed configuration are compared with the CPU result: the whole batch at N = 1 (device-resident
and host-buffer paths), and at N > 1 every rank checks a remote rank's slot of the gathered result against its own tokenisation of
that shard.
`--impl reference` times the reference's own CPU evaluate() chain (oracle/_ref) on the FULL batch, with the fastest thread
count (concurrent evaluate() calls on one op instance, rows sharded — what several infer requests would do).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

PULL_FROM_WORLD = 8      # N > 1 exchange: all-gatherv by pull from this many ranks on (16-bit wire), emit fused with peer stores below (measured)

WORKLOADS = {
    "c1": dict(kind="bpe", vocab="gpt2_synth", rows=65536, row_bytes=512, gen="ascii",
               name="C1: gpt2-shaped byte-level BPE (50 257 vocab / 50 000 merges, synthetic stand-in gpt2_synth), "
                    "65 536 x 512 B printable-ASCII docs per GPU, fused RegexSplit->BPETokenizer"),
    "c2": dict(kind="wordpiece", vocab="bert_synth", rows=65536, row_bytes=256, gen="ascii_lower",
               name="C2: bert-shaped WordPiece (30 522 vocab, synthetic stand-in bert_synth), 65 536 x 256 B lower-cased "
                    "printable-ASCII docs per GPU, fused RegexSplit x2->WordpieceTokenizer"),
    "c4": dict(kind="detok", vocab="llama2_detok_synth", rows=1024, row_bytes=1024, gen="ids",
               name="C4: detokenize, VocabDecoder + ByteFallback fused, 1 024 x 1 024 token ids (Llama-2-shaped 32 000 vocab with 256 <0xHH> "
                    "tokens, synthetic stand-in), skip_tokens {0,1,2}"),
    "norm": dict(kind="norm", vocab="-", rows=65536, row_bytes=256, gen="ascii_ctl",
                 name="BERT normaliser (SURVEY 8f.4): RegexNormalization x4 + CharsMapNormalization NFD + case fold, the six ops of "
                      "hf_parser.py:84-102 in one b200tok_normalize_chain_run, 65 536 x 256 B printable ASCII with 1 % tabs / 0.5 % control bytes"),
    "c3": dict(kind="bpe", vocab="llama3_synth", rows=32768, row_bytes=1024, gen="utf8",
               name="C3 shard: Llama-3-shaped BPE (128 256 vocab, synthetic stand-in llama3_synth), 32 768 x 1 KiB "
                    "mixed-UTF-8 docs per GPU (262 144 rows at 8 GPUs), fused RegexSplit->BPETokenizer"),
}


def make_batch(w, seed):
    import cases
    if w["gen"] == "ascii":
        return cases.random_ascii_batch(w["rows"], w["row_bytes"], seed)
    if w["gen"] == "ascii_lower":
        return cases.random_ascii_batch(w["rows"], w["row_bytes"], seed, lower=True)
    return cases.mixed_utf8_batch(w["rows"], w["row_bytes"], seed)


class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        busy = sm[len(sm) // 2:] if sm else []     # the upper half of the samples = under load
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def profiled(workload, key):
    """A figure of profiles/roofline_traffic.json (ncu pass over `bench.py --device-only`, tools/step_metrics.py) or None."""
    try:
        return (json.loads((ROOT / "profiles" / "roofline_traffic.json").read_text()).get(workload) or {}).get(key)
    except Exception:
        return None


def cpu_sample(w, batch, rows):
    rb, re_, b, e, c = batch
    n = min(rows, len(rb))
    L = w["row_bytes"]
    return (rb[:n], re_[:n], b[:n], e[:n], c[: n * L])


def cpu_runner(w, cached=True):
    """Returns (f(batch, threads) -> (begins, ends, ids), kind).  kind "reference": the reference's own op classes
    (oracle/_ref/libovtok_ref.so, built from /root/reference/src by oracle/Makefile); "port": the oracle restatement."""
    from openvino_tokenizers_b200 import assets as A
    a = A.load_bpe(w["vocab"]) if w["kind"] == "bpe" else A.load_wordpiece(w["vocab"])
    if cached:
        import refops
        if refops.available():
            return refops.RefChain(w["kind"], a), "reference"
    import oracle
    from openvino_tokenizers_b200.strings import pack_strings
    if w["kind"] == "bpe":
        v, ml, mr, ad, aid = a.tensors()
        sp = oracle.SplitOracle(a.split_pattern, "isolate")
        bpe = oracle.BpeOracle(v, ml, mr, ad, aid, cache_capacity=a.cache_capacity, use_cache=cached)

        def run(batch, threads):
            s = sp(*batch, threads=threads)
            return bpe(s[0], s[1], s[2], s[3], batch[4], threads=threads)
        return run, "port"
    s1 = oracle.SplitOracle(A.BERT_WHITESPACE_PATTERN, "remove")
    s2 = oracle.SplitOracle(A.BERT_PUNCT_PATTERN, "isolate")
    wp = oracle.WordpieceOracle(pack_strings(a.vocab), a.suffix_indicator, a.max_bytes_per_word)

    def run(batch, threads):
        r1 = s1(*batch, threads=threads)
        r2 = s2(r1[0], r1[1], r1[2], r1[3], batch[4], threads=threads)
        return wp(r2[0], r2[1], r2[2], r2[3], batch[4], a.unk_token_id, threads=threads)
    return run, "port"


def checker(w):
    """The fastest exact CPU result for verification: the oracle restatement (equal to the reference's code, tests/test_reference_pin.py)
    with its result cache off, so that rows shard over all host threads without the cache mutex."""
    return cpu_runner(w, cached=False)[0]


def time_cpu(run, batch, threads, repeats=3):
    run(batch, threads)   # warm-up: builds the tables on first use and fills the reference's 20 000-entry BPE cache
    best = 1e30
    for _ in range(repeats):
        t0 = time.perf_counter()
        run(batch, threads)
        best = min(best, time.perf_counter() - t0)
    return len(batch[4]) / 1e6 / best


def main_reference(args, w, rank, world):
    if rank != 0:
        return
    batch = make_batch(w, 1234)
    cores = os.cpu_count() or 1
    run, kind = cpu_runner(w)
    # The reference's evaluate() is a serial loop and its BPE result cache sits behind one shared_mutex
    # (src/bpe_tokenizer.cpp:196-205,331-338), so more threads are not always faster: probe 1 .. all cores on a small
    # sample (concurrent evaluate() calls on one op instance, rows sharded) and time the steps with the best count.
    probe = cpu_sample(w, batch, 2048)
    cands = sorted({1, 2, 4, 8, cores} & set(range(1, cores + 1)))
    best_t, best_v = 1, 0.0
    for t in cands:
        v = time_cpu(run, probe, t, repeats=1)
        if v > best_v:
            best_t, best_v = t, v
    for _ in range(args.warmup):
        run(batch, best_t)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run(batch, best_t)
    dt = (time.perf_counter() - t0) / args.steps
    mbs = len(batch[4]) / 1e6 / dt
    what = ("the reference's own op classes (oracle/_ref: RegexSplit / BPETokenizer / WordpieceTokenizer evaluate() compiled from the reference sources)"
            if kind == "reference" else "CPU oracle port of the reference ops (oracle/_ref not present)")
    desc = (f"all {len(batch[0])} rows x {w['row_bytes']} B per step (the full batch); {best_t} thread(s) = the fastest of {cands} "
            f"on this host ({cores} cores)")
    print(json.dumps({
        "impl": "reference", "metric": "input text tokenized (bit-exact ids)", "value": mbs, "unit": "MB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": w["name"], "rows_per_gpu": len(batch[0]), "row_bytes": w["row_bytes"], "note": what},
        "cpu_baseline": {"value": mbs, "unit": "MB/s", "cores": best_t, "kind": kind, "sample": desc},
        "e2e": {"value": mbs, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main_detok(args, w, rank, world, local_rank):
    """C4: metric = MB/s of detokenized text produced (bit-exact bytes); a step = one VocabDecoder+ByteFallback pass over 1 Mi ids."""
    import ctypes as C
    import torch
    import oracle
    from openvino_tokenizers_b200 import _capi as K
    from openvino_tokenizers_b200 import assets as A
    from openvino_tokenizers_b200 import ops
    from openvino_tokenizers_b200.strings import pack_strings
    if rank != 0:
        return
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    vocab = pack_strings(A.load_detok_vocab(w["vocab"]))
    Bn, Sn = w["rows"], w["row_bytes"]
    ids = np.random.default_rng(1234).integers(0, len(vocab[0]), size=(Bn, Sn)).astype(np.int32)
    skip = np.array([0, 1, 2], np.int32)
    dec = ops.VocabDecoder(skip_tokens=[0, 1, 2], byte_fallback=True, device=local_rank)
    ref = dec.evaluate([ids, *vocab])                      # creates the handle; host path
    n_out = int(len(ref[4]))
    L = K.lib()
    cap = int(L.b200tok_vocabdec_max_chars(dec.handle, Bn, Sn))
    d_ids = torch.from_numpy(ids).to(dev)
    d_skip = torch.from_numpy(skip).to(dev)
    d_rb, d_re = torch.empty(Bn, dtype=torch.int32, device=dev), torch.empty(Bn, dtype=torch.int32, device=dev)
    d_b, d_e = torch.empty(Bn * Sn, dtype=torch.int32, device=dev), torch.empty(Bn * Sn, dtype=torch.int32, device=dev)
    d_c = torch.empty(cap, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    def step_device():
        out = K.Decoded(d_rb.data_ptr(), d_re.data_ptr(), d_b.data_ptr(), d_e.data_ptr(), d_c.data_ptr(), cap, 0, K.MEM_DEVICE)
        K.check(L.b200tok_vocabdec_run(dec.handle, C.c_void_p(d_ids.data_ptr()), C.c_int64(Bn), C.c_int64(Sn), C.c_void_p(d_skip.data_ptr()),
                                       C.c_int64(3), 1, C.byref(out), K.MEM_DEVICE, C.c_void_p(stream.cuda_stream)))
        return out.n_chars
    for _ in range(args.warmup):
        step_device(); dec.evaluate([ids, *vocab])
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = dec.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record()
        n_dev = step_device()
        ev[k][1].record()
        torch.cuda.synchronize()
    launches = dec.launches - launches0
    clocks = sampler.stop()
    if args.device_only:
        print(json.dumps({"device_only": True, "steps": args.steps}))
        return
    assert n_dev == n_out and bytes(d_c[:n_out].cpu().numpy()) == bytes(ref[4])
    ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    # end to end: pinned host buffers through the same C-ABI call (H2D of the ids, D2H of offsets + bytes inside the call)
    h_ids, h_skip = torch.from_numpy(ids).pin_memory(), torch.from_numpy(skip).pin_memory()
    h_rb, h_re = torch.empty(Bn, dtype=torch.int32).pin_memory(), torch.empty(Bn, dtype=torch.int32).pin_memory()
    h_b, h_e = torch.empty(Bn * Sn, dtype=torch.int32).pin_memory(), torch.empty(Bn * Sn, dtype=torch.int32).pin_memory()
    h_c = torch.empty(cap, dtype=torch.uint8).pin_memory()

    def step_host():
        out = K.Decoded(h_rb.data_ptr(), h_re.data_ptr(), h_b.data_ptr(), h_e.data_ptr(), h_c.data_ptr(), cap, 0, K.MEM_HOST)
        K.check(L.b200tok_vocabdec_run(dec.handle, C.c_void_p(h_ids.data_ptr()), C.c_int64(Bn), C.c_int64(Sn), C.c_void_p(h_skip.data_ptr()),
                                       C.c_int64(3), 1, C.byref(out), K.MEM_HOST, None))
        return out.n_chars
    for _ in range(3):
        step_host()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        n_host = step_host()
    e2e_s = (time.perf_counter() - t0) / args.steps
    assert n_host == n_out and bytes(h_c[:n_out].numpy()) == bytes(ref[4])
    peak, peak_src = measured_peak()
    algo = 4 * Bn * Sn + 8 * Bn * Sn + n_out + 8 * Bn            # ids in, per-token offsets + bytes + row extents out
    t0 = time.perf_counter()
    for _ in range(3):
        o = oracle.vocab_decoder(ids, vocab, [0, 1, 2]); oracle.byte_fallback(o[2], o[3], o[4])
    cpu_s = (time.perf_counter() - t0) / 3
    print(json.dumps({
        "metric": "detokenized text produced (bit-exact bytes)", "value": n_out / 1e6 / (ms / 1e3), "unit": "MB/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": w["name"], "ids": Bn * Sn, "out_bytes": n_out, "l2": "256 MiB buffer zeroed between timed steps (L2 flush)",
                   "note": "the device-resident call returns the byte count to the host: one stream synchronisation is inside the step"},
        "e2e": {"value": n_out / 1e6 / e2e_s, "unit": "MB/s", "h2d_bytes_per_step": 4 * Bn * Sn, "d2h_bytes_per_step": 8 * Bn * Sn + n_out + 8 * Bn,
                "ms_per_step": e2e_s * 1e3, "path": "b200tok_vocabdec_run with B200TOK_MEM_HOST on pinned buffers"},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": algo / 1e9 / (ms / 1e3), "peak": peak, "unit": "GB/s", "frac": algo / 1e9 / (ms / 1e3) / peak, "traffic": profiled("c4", "step_dram_bytes"),
                     "kernel": "decode_len_kernel + cub scan + decode_copy_kernel (whole step)", "algorithmic_bytes_per_launch": algo, "peak_source": peak_src},
        "cpu_baseline": {"value": n_out / 1e6 / cpu_s, "unit": "MB/s", "cores": 1, "kind": "port", "sample": "the full 1 Mi ids, mean of 3"},
    }))


def main_norm(args, w, rank, world, local_rank):
    """BERT normaliser chain: metric = MB/s of input text normalised (bit-exact bytes); a step = the six ops over the batch."""
    import ctypes as C
    import torch
    import oracle
    sys.path.insert(0, str(ROOT / "tests"))
    import normcases as NC
    from openvino_tokenizers_b200 import _capi as K
    from openvino_tokenizers_b200 import ops
    if rank != 0:
        return
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    golden = json.loads((ROOT / "tests" / "golden" / "normalization_layer_tests.json").read_text())
    bs = golden["bert_steps"]
    nfd, fold = NC.unicodedata_blob("NFD", False), NC.unicodedata_blob(None, True)      # compiled by the installed sentencepiece from unicodedata rules
    steps = [("regex", bs[0]), ("regex", bs[1]), ("regex", bs[2]), ("charsmap", nfd), ("regex", bs[3]), ("charsmap", fold)]
    chain = [ops.RegexNormalization(x["global_replace"], device=local_rank).prepare(x["search"], x["replace"]) if k == "regex"
             else ops.CharsMapNormalization(device=local_rank).prepare(x) for k, x in steps]
    Bn, Ln = w["rows"], w["row_bytes"]
    rng = np.random.default_rng(1234)
    chars = rng.integers(0x20, 0x7F, size=Bn * Ln, dtype=np.uint8)
    r = rng.random(Bn * Ln)
    chars[r < 0.01] = 0x09
    chars[(r >= 0.01) & (r < 0.015)] = 0x01
    b = np.arange(Bn, dtype=np.int32) * Ln
    e = b + Ln
    N = chars.size

    def oracle_chain(ins):
        cur = list(ins)
        for k, x in steps:
            cur = list(oracle.regex_normalize(x["search"], x["replace"], x["global_replace"], *cur)) if k == "regex" else list(oracle.charsmap_normalize(x, *cur))
        return cur
    ref = ops.normalize_chain(chain, [b, e, chars])                      # host path (creates nothing new; also the e2e call)
    n_out = int(ref[2].size)
    sample = slice(0, 2048)
    exp = oracle_chain([b[sample], e[sample], chars])
    assert bytes(ref[2][: int(ref[1][sample][-1])]) == bytes(exp[2]), "device result differs from the oracle"
    L = K.lib()
    hb, he = torch.from_numpy(b).pin_memory(), torch.from_numpy(e).pin_memory()
    hc = torch.from_numpy(chars).pin_memory()
    db, de, dc = hb.to(dev), he.to(dev), torch.cat([hc, torch.zeros(64, dtype=torch.uint8)]).to(dev)
    cap = N + 64
    ob, oe = torch.empty(Bn, dtype=torch.int32, device=dev), torch.empty(Bn, dtype=torch.int32, device=dev)
    oc = torch.empty(cap, dtype=torch.uint8, device=dev)
    hob, hoe, hoc = torch.empty(Bn, dtype=torch.int32).pin_memory(), torch.empty(Bn, dtype=torch.int32).pin_memory(), torch.empty(cap, dtype=torch.uint8).pin_memory()
    hs = (C.c_void_p * len(chain))(*[o.handle for o in chain])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    got = C.c_int64(0)

    def step(mem, tb, te, tc, tob, toe, toc):
        K.check(L.b200tok_normalize_chain_run(hs, len(chain), C.c_void_p(tb.data_ptr()), C.c_void_p(te.data_ptr()), C.c_int64(Bn), C.c_void_p(tc.data_ptr()),
                                              C.c_int64(N), None, C.c_void_p(tob.data_ptr()), C.c_void_p(toe.data_ptr()), C.c_void_p(toc.data_ptr()),
                                              C.c_int64(cap), C.byref(got), mem, C.c_void_p(stream.cuda_stream)))
    for _ in range(args.warmup):
        step(K.MEM_DEVICE, db, de, dc, ob, oe, oc); step(K.MEM_HOST, hb, he, hc, hob, hoe, hoc)
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = sum(o.launches for o in chain)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record()
        step(K.MEM_DEVICE, db, de, dc, ob, oe, oc)
        ev[k][1].record()
        torch.cuda.synchronize()
    launches = sum(o.launches for o in chain) - launches0
    assert got.value == n_out and bytes(oc[:n_out].cpu().numpy()) == bytes(ref[2])
    if args.device_only:
        sampler.stop()
        print(json.dumps({"device_only": True, "steps": args.steps}))
        return
    ms = sum(a.elapsed_time(b_) for a, b_ in ev) / args.steps
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(K.MEM_HOST, hb, he, hc, hob, hoe, hoc)
    e2e_s = (time.perf_counter() - t0) / args.steps
    clocks = sampler.stop()
    assert bytes(hoc[:n_out].numpy()) == bytes(ref[2])
    peak, peak_src = measured_peak()
    algo = 2 * N + n_out + 16 * Bn                      # the text is read by the lengths pass and the write pass, the result written once
    rows = 8192
    t0 = time.perf_counter()
    oracle_chain([b[:rows], e[:rows], chars])
    cpu_s = time.perf_counter() - t0
    print(json.dumps({
        "metric": "input text normalised (bit-exact bytes)", "value": N / 1e6 / (ms / 1e3), "unit": "MB/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": w["name"], "rows": Bn, "row_bytes": Ln, "out_bytes": n_out, "l2": "256 MiB buffer zeroed between timed steps (L2 flush)",
                   "note": "the device-resident call sizes its result on the host: two stream synchronisations are inside the step"},
        "e2e": {"value": N / 1e6 / e2e_s, "unit": "MB/s", "h2d_bytes_per_step": N + 8 * Bn, "d2h_bytes_per_step": n_out + 8 * Bn,
                "ms_per_step": e2e_s * 1e3, "path": "b200tok_normalize_chain_run with B200TOK_MEM_HOST on pinned buffers"},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": algo / 1e9 / (ms / 1e3), "peak": peak, "unit": "GB/s", "frac": algo / 1e9 / (ms / 1e3) / peak, "traffic": profiled("norm", "step_dram_bytes"),
                     "kernel": "compose_kernel<lengths> + cub scan + compose_kernel<write> (whole step, host round trips included)",
                     "algorithmic_bytes_per_launch": algo, "peak_source": peak_src},
        "cpu_baseline": {"value": rows * Ln / 1e6 / cpu_s, "unit": "MB/s", "cores": 1, "kind": "port",
                         "sample": f"first {rows} of {Bn} rows, one pass: PCRE2 pcre2_substitute x4 + the restated sentencepiece Normalizer x2"},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c1", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--device-only", action="store_true", help="profiling runs: only the device-resident step (no host-buffer calls, no JSON contract)")
    ap.add_argument("--no-unfused", action="store_true", help="skip the separate-ops (unfused) host-to-host measurement")
    ap.add_argument("--no-verify", action="store_true", help="skip the full-batch CPU-oracle comparison after the timed regions")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if w["kind"] == "norm":
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "the reference arm is defined for the tokenize workloads; the normaliser workload reports its CPU port inline"}))
            return
        return main_norm(args, w, rank, world, local_rank)
    if w["kind"] == "detok":
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "the reference arm is defined for the tokenize workloads; C4 reports its CPU port inline"}))
            return
        return main_detok(args, w, rank, world, local_rank)
    if args.impl == "reference":
        return main_reference(args, w, rank, world)

    import torch
    import torch.distributed as dist
    from openvino_tokenizers_b200 import runtime as R
    from openvino_tokenizers_b200.sharded import allgather_ragged_slots

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    pipe = R.TokenizerPipeline(w["kind"], w["vocab"], device=local_rank)
    batch = make_batch(w, 1234 + rank)                    # weak scaling: every rank has its own shard
    n_bytes = int(len(batch[4]))
    db = R.to_device(batch, dev)
    hb = R.to_pinned(batch)
    ho = pipe.alloc_host_out(db.n_rows, db.n_chars + db.n_elems)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    # N > 1: the emit step stores every id row straight into the result buffers of all ranks over NVLink peer memory
    # (sharded.PeerGather: torch symmetric memory + b200tok_split_bpe_run_sharded), i.e. the all-gatherv is fused into the
    # compaction kernel.  B200TOK_BENCH_EXCHANGE=nccl selects the NCCL slot all-gather instead (row blocks pipelined).
    n_blocks = 1
    exchange = "single GPU"
    if world > 1:
        mode = os.environ.get("B200TOK_BENCH_EXCHANGE", "auto")
        if mode == "auto":      # measured (profiles/r02_bench_*_n8_*.json): pull wins from 8 ranks on when the ids fit the 16-bit wire
            mode = "pull" if (world >= PULL_FROM_WORLD and len(pipe.assets.vocab) < 0xFFFF) else "peer"
        pg = None
        if mode == "pull":
            try:
                from openvino_tokenizers_b200.sharded import PullGather
                wire16 = len(pipe.assets.vocab) < 0xFFFF and os.environ.get("B200TOK_WIRE16", "auto") != "0"
                pg = PullGather(db.n_rows, db.n_chars + (db.n_elems if pipe.kind != "bpe" else 0), dev, wire16=wire16)
                exchange = ("row shards; all-gatherv by pull over NVLink peer memory: one-GPU tokenisation into peer-mapped source buffers, one "
                            "symmetric-memory barrier, then every rank reads all peers with 16-byte loads and widens into its own i32 result"
                            + (" (16-bit ids on the wire)" if wire16 else " (32-bit ids on the wire)"))
            except Exception as ex:
                print(f"[bench] pull exchange unavailable ({type(ex).__name__}: {ex}); trying the peer-store emit", file=sys.stderr)
                mode = "peer"
        if mode == "peer":
            try:
                from openvino_tokenizers_b200.sharded import PeerGather
                wire16 = os.environ.get("B200TOK_WIRE16", "auto")
                wire16 = (world >= 4 and len(pipe.assets.vocab) < 0xFFFF) if wire16 == "auto" else wire16 == "1"
                pg = PeerGather(db.n_rows, db.n_chars, dev, wire16=wire16)
                exchange = "row shards; emit fused with the all-gatherv: the tokenizer kernel stores id rows into every rank's buffers over NVLink peer memory" + (" (16-bit ids on the wire, widened locally)" if wire16 else "") + (" (NVLS multicast stores)" if pg.multicast else "")
            except Exception as ex:      # symmetric memory unavailable on this box: fall back to the NCCL exchange (still GPU-only)
                print(f"[bench] peer-memory exchange unavailable ({type(ex).__name__}: {ex}); using NCCL", file=sys.stderr)
        if pg is None:
            n_blocks = int(os.environ.get("B200TOK_BENCH_BLOCKS", "2"))
            blocks = [R.to_device(bk, dev) for bk in R.split_rows(batch, n_blocks)]
            block_out = [pipe.alloc_device_out(bk.n_rows, bk.n_chars + bk.n_elems) for bk in blocks]
            gathered = [torch.empty(world * bk.n_chars, dtype=torch.int32, device=dev) for bk in blocks]
            exchange = ("row shards in %d row block(s); all-gatherv of block k's ragged id rows over NCCL (fixed-capacity slots, no host sync) "
                        "overlaps the tokenisation of block k+1" % n_blocks)

    def step_device():
        if world == 1:
            return pipe.run_device(db)
        if pg is not None:
            pg.run(pipe, db)
            return [{"n": pg.n}]
        works = []
        for bk, bo, g in zip(blocks, block_out, gathered):
            o = pipe.run_device(bk, bo)
            # ids <= bytes (src/bpe_tokenizer.cpp:135): gather the first n_chars slots of every rank's buffer, no host sync
            works.append(allgather_ragged_slots(o["ids"][: bk.n_chars], o["begins"], o["ends"], g, async_op=True))
        for wk in works:
            wk()
        return block_out

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- warm-up ----
    for _ in range(args.warmup):
        step_device()
        if not args.device_only:
            pipe.run_host(hb, ho)
    barrier()
    if args.device_only:
        for _ in range(args.steps):
            flush.zero_()
            step_device()
        barrier()
        print(json.dumps({"device_only": True, "steps": args.steps}))
        return

    # ---- device-resident timed region: K steps, one CUDA-event pair per step, L2 flushed between steps ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = pipe.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record()
        o = step_device()
        ev[k][1].record()
        torch.cuda.synchronize()
    barrier()
    launches = pipe.launches - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    ms_per_step = total_ms / args.steps
    n_ids = int(o["n"].item()) if world == 1 else int(sum(int(x["n"].item()) for x in o))
    # ---- the dominant kernel alone: the same K steps again with the library's own CUDA-event pair around that kernel, on the stream
    # it is launched on (b200tok_set_timing; the timed steps above replay a captured graph, which cannot carry timing events)
    pipe.set_timing(True)
    kernel_ms = []
    for k in range(args.steps):
        flush.zero_()
        step_device()
        torch.cuda.synchronize()
        km = pipe.last_kernel_ms()
        if km > 0:
            kernel_ms.append(km)
    pipe.set_timing(False)
    clocks = sampler.stop()
    barrier()

    # ---- end to end: host (pinned) buffers -> C ABI -> host buffers, every step ----
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        n_host = pipe.run_host(hb, ho)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item()) / args.steps
    h2d = n_bytes + 8 * db.n_rows + 8 * db.n_elems
    d2h = 4 * n_ids + 8 * db.n_rows
    e2e_path = "b200tok_split_*_run with B200TOK_MEM_HOST on pinned buffers"
    e2e_local = None
    if world > 1 and pg is not None:
        # N > 1: the SAME job as `value` (tokenise the shard + all-gatherv on the devices), fed from and drained to pinned host buffers:
        # H2D of the shard, the sharded step, then D2H of this rank's slot of the gathered ids and of the row extents of ALL ranks' rows
        # (the host ends up with the whole gathered result across the ranks' buffers, every id crossing PCIe once)
        e2e_local = {"value": n_bytes * world / 1e6 / e2e_s, "unit": "MB/s", "ms_per_step": e2e_s * 1e3,
                     "path": "every rank's own b200tok_split_*_run host -> host, no exchange (the N = 1 path per rank)"}
        g_rows = world * db.n_rows
        h_gb, h_ge = torch.empty(g_rows, dtype=torch.int32).pin_memory(), torch.empty(g_rows, dtype=torch.int32).pin_memory()
        slot0 = rank * pg.cap

        def gathered_host_step():
            for dst, src in zip((db.rb, db.re, db.begins, db.ends), hb[:4]):
                dst.copy_(src, non_blocking=True)
            db.chars[:n_bytes].copy_(hb[4], non_blocking=True)
            pg.run(pipe, db)
            n_mine = int(pg.n.item())                                   # (one synchronisation: the size of the id copy)
            # the pull gathers compact rows; the fused emit leaves the rows at their worst-case positions inside the slot: copy its whole extent
            span = n_mine if getattr(pg, "compact_slots", False) else min(pg.cap, ho["ids"].numel())
            ho["ids"][:span].copy_(pg.ids[slot0:slot0 + span], non_blocking=True)
            h_gb.copy_(pg.begins, non_blocking=True)
            h_ge.copy_(pg.ends, non_blocking=True)
            torch.cuda.synchronize()
            return n_mine
        for _ in range(2):
            gathered_host_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            n_mine = gathered_host_step()
        e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        e2e_s = float(e2e_s.item()) / args.steps
        d2h = 4 * (n_mine if getattr(pg, "compact_slots", False) else min(pg.cap, ho["ids"].numel())) + 8 * g_rows
        e2e_path = ("pinned host shard -> H2D -> tokenise + all-gatherv on the devices (the step `value` times) -> D2H of this rank's id slot "
                    "and of all row extents; per-rank bytes")
        n_host = pipe.run_host(hb, ho)       # (restore the plain host-path result for the checks below)

    # ---- the same host-to-host job through the SEPARATE ops (what an IR costs when the load-time fusion of the ov::Op shim is off, or
    # for split patterns it cannot fuse): RegexSplit host -> host, then BPETokenizer / WordpieceTokenizer host -> host; the piece
    # offsets (8 bytes per ~1.8-byte piece) cross PCIe twice
    unfused = None
    if world == 1 and not args.no_unfused:
        import ctypes as C
        from openvino_tokenizers_b200 import _capi as K
        Lb = K.lib()
        cap_p = db.n_chars + db.n_elems
        pin = lambda n, dt=torch.int32: torch.empty(max(n, 1), dtype=dt).pin_memory()
        s_rb, s_re, s_b, s_e = pin(db.n_rows), pin(db.n_rows), pin(cap_p), pin(cap_p)
        s2_rb, s2_re, s2_b, s2_e = pin(db.n_rows), pin(db.n_rows), pin(cap_p), pin(cap_p)

        def split_host(handle, rb_t, re_t, b_t, e_t, n_el, orb, ore, ob, oe):
            rin = K.RaggedStrings(rb_t.data_ptr(), re_t.data_ptr(), db.n_rows, b_t.data_ptr(), e_t.data_ptr(), n_el, hb[4].data_ptr(), hb[4].numel(), None, K.MEM_HOST)
            out = K.RaggedStringsOut(orb.data_ptr(), ore.data_ptr(), ob.data_ptr(), oe.data_ptr(), None, cap_p, 0, 0, K.MEM_HOST)
            K.check(Lb.b200tok_regexsplit_run(handle, C.byref(rin), C.byref(out), None))
            return int(out.n_elems)

        def unfused_step():
            n1 = split_host(pipe.split1.handle, hb[0], hb[1], hb[2], hb[3], hb[2].numel(), s_rb, s_re, s_b, s_e)
            rb_t, re_t, b_t, e_t = s_rb, s_re, s_b, s_e
            if pipe.split2 is not None:
                n1 = split_host(pipe.split2.handle, s_rb, s_re, s_b, s_e, n1, s2_rb, s2_re, s2_b, s2_e)
                rb_t, re_t, b_t, e_t = s2_rb, s2_re, s2_b, s2_e
            rin = K.RaggedStrings(rb_t.data_ptr(), re_t.data_ptr(), db.n_rows, b_t.data_ptr(), e_t.data_ptr(), n1, hb[4].data_ptr(), hb[4].numel(), None, K.MEM_HOST)
            out = K.RaggedIds(ho["begins"].data_ptr(), ho["ends"].data_ptr(), ho["ids"].data_ptr(), ho["cap"], 0, None, K.MEM_HOST)
            if pipe.kind == "bpe":
                K.check(Lb.b200tok_bpe_run(pipe.tok.handle, C.byref(rin), C.byref(out), None))
            else:
                K.check(Lb.b200tok_wordpiece_run(pipe.tok.handle, C.byref(rin), C.c_int32(pipe.unk), C.byref(out), None))
            return n1, int(out.n_ids)
        unfused_step()
        k_un = max(3, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(k_un):
            n_pieces, n_un = unfused_step()
        un_s = (time.perf_counter() - t0) / k_un
        assert n_un == n_host, "separate ops and fused call disagree on the id count"
        unfused = {"value": n_bytes / 1e6 / un_s, "unit": "MB/s", "ms_per_step": un_s * 1e3, "pieces": n_pieces,
                   "path": "b200tok_regexsplit_run -> b200tok_bpe_run / b200tok_wordpiece_run, each host -> host (pinned buffers)"}
        n_host = pipe.run_host(hb, ho)       # (restore the fused result in the host buffers for the check below)

    # ---- verification of the timed configuration (outside every timed region) ----
    # N = 1: the whole batch, device-resident result and host-buffer result, against the CPU oracle.
    # N > 1: every rank re-tokenises the shard of rank (rank + 1) % N on its own GPU and compares it with that rank's slot of the
    #        gathered result it holds; rank 0 also checks its own shard against the CPU oracle.
    verified = {}
    own = pipe.run_device(db)
    torch.cuda.synchronize()
    n_own = int(own["n"].item())
    own_np = (own["begins"].cpu().numpy(), own["ends"].cpu().numpy(), own["ids"][:n_own].cpu().numpy())
    host_np = (ho["begins"].numpy(), ho["ends"].numpy(), ho["ids"][:n_host].numpy())
    assert n_host == n_own and all(np.array_equal(x, y) for x, y in zip(own_np, host_np)), "host-buffer path differs from the device-resident path"
    if rank == 0 and not args.no_verify:
        exp = checker(w)(batch, os.cpu_count() or 1)
        assert all(np.array_equal(x, y) for x, y in zip(own_np, exp)), "GPU ids differ from the CPU oracle on the timed batch"
        verified["rank0_rows_vs_cpu_oracle"] = int(db.n_rows)
    if world > 1:
        peer = (rank + 1) % world
        pb = R.to_device(make_batch(w, 1234 + peer), dev)
        po = pipe.run_device(pb, pipe.alloc_device_out(pb.n_rows, pb.n_chars + pb.n_elems))
        step_device()
        torch.cuda.synchronize()
        n_p = int(po["n"].item())
        pb_, pe_, pids = po["begins"].cpu().numpy().astype(np.int64), po["ends"].cpu().numpy().astype(np.int64), po["ids"][:n_p].cpu().numpy()
        if pg is not None:
            gb, ge, gids = (t.cpu().numpy() for t in (pg.begins, pg.ends, pg.ids))
            rows = slice(peer * db.n_rows, (peer + 1) * db.n_rows)
            gb, ge = gb[rows].astype(np.int64), ge[rows].astype(np.int64)
            assert np.array_equal(ge - gb, pe_ - pb_), "gathered row lengths of the peer slot differ"
            lens = ge - gb
            idx = np.repeat(gb - (np.cumsum(lens) - lens), lens) + np.arange(int(lens.sum()))
            assert np.array_equal(gids[idx], pids), "gathered ids of the peer slot differ"
            verified["peer_slot_rows_vs_local_gpu"] = int(db.n_rows)
        barrier()
    n_ids = n_own if world == 1 else n_ids

    total_bytes = n_bytes * world
    value = total_bytes / 1e6 / (ms_per_step / 1e3)
    e2e_value = total_bytes / 1e6 / e2e_s

    if rank == 0:
        peak, peak_src = measured_peak()
        algo_bytes = (n_bytes + 16 * db.n_rows + 4 * n_ids) // n_blocks     # per launch: N > 1 launches the kernel once per row block
        k_ms = statistics.mean(kernel_ms) if kernel_ms else None
        achieved = algo_bytes / 1e9 / (k_ms / 1e3) if k_ms else None
        # ncu-measured DRAM traffic and warp-instruction counts of this workload's step (profiles/roofline_traffic.json, written by
        # tools/step_metrics.py from an `ncu --metrics dram__bytes_*,smsp__inst_executed.sum` pass over `bench.py --device-only`)
        traffic, prof = None, {}
        tp = ROOT / "profiles" / "roofline_traffic.json"
        if tp.exists():
            try:
                prof = json.loads(tp.read_text()).get(args.workload) or {}
                traffic = prof.get("kernel_dram_bytes")
            except Exception:
                prof = {}
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                    "traffic": traffic, "kernel": pipe.dominant_kernel, "kernel_ms": k_ms,
                    "algorithmic_bytes_per_launch": algo_bytes, "peak_source": peak_src,
                    "kernel_share_of_step": (k_ms * n_blocks / ms_per_step) if k_ms else None, "launches_per_step": n_blocks,
                    "step_traffic": prof.get("step_dram_bytes"), "traffic_source": prof.get("source")}
        if prof.get("kernel_warp_inst") and k_ms:
            # the bound this integer kernel really runs against: one warp instruction per SM sub-partition per cycle
            sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
            floor_ms = prof["kernel_warp_inst"] / (148 * 4) / sm_hz * 1e3
            roofline["issue"] = {"warp_inst_per_launch": prof["kernel_warp_inst"], "warp_inst_per_input_byte": prof["kernel_warp_inst"] / n_bytes,
                                 "issue_floor_ms": floor_ms, "frac_of_issue_peak": floor_ms / k_ms,
                                 "note": "warp instructions (ncu smsp__inst_executed.sum) / (148 SMs x 4 schedulers x SM clock) over the kernel's event-timed duration"}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            run, kind = cpu_runner(w)
            cores = os.cpu_count() or 1
            s1 = cpu_sample(w, batch, 4096)
            v1 = time_cpu(run, s1, 1)
            sN = cpu_sample(w, batch, 16384)
            vN = time_cpu(run, sN, cores, repeats=1)
            cpu = {"value": v1, "unit": "MB/s", "cores": 1, "kind": kind,
                   "sample": f"first 4096 of {w['rows']} rows x {w['row_bytes']} B, best of 3 after 1 warm-up pass; the reference's "
                             "evaluate() for RegexSplit/BPE is a serial loop, hence 1 thread",
                   "all_cores": {"value": vN, "cores": cores, "sample": "first 16384 rows, rows sharded over concurrent evaluate() calls on one op instance"}}
        print(json.dumps({
            "metric": "input text tokenized (bit-exact ids)", "value": value, "unit": "MB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": w["name"], "rows_per_gpu": db.n_rows, "row_bytes": w["row_bytes"], "tokens_per_gpu": n_ids,
                       "l2": "256 MiB buffer zeroed between timed steps (L2 flush), outside the per-step event pair",
                       "multi_gpu": exchange},
            "e2e": {"value": e2e_value, "unit": "MB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s * 1e3, "path": e2e_path},
            "e2e_no_exchange": e2e_local, "e2e_unfused": unfused,
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "verified": verified,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
