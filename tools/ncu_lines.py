"""Per-source-line instruction / stall profile of one kernel from an .ncu-rep captured with `--set full --import-source on`:
joins the SASS page of the report (instructions executed, stall samples per instruction) with the line table `nvdisasm -g`
prints for the same cubin (the .so must be the build that was profiled).
   python tools/ncu_lines.py <report.ncu-rep> <mangled-kernel-substring> [top]"""
import csv
import re
import subprocess
import sys
import tempfile
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
rep, ksub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
with tempfile.TemporaryDirectory() as td:
    subprocess.check_call(["cuobjdump", "-xelf", "all", str(ROOT / "openvino_tokenizers_b200/csrc/libb200tok.so")], cwd=td, stdout=subprocess.DEVNULL)
    sass = subprocess.run(["nvdisasm", "-g", "-c", str(Path(td) / "api.sm_100a.cubin")], capture_output=True, text=True).stdout.splitlines()
line_of = {}
cur, infn = None, False
for ln in sass:
    if ln.startswith(".text."):
        infn = ksub in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (Path(m.group(1)).name, int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(raw))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]
ci = {n: H.index(n) for n in ("Address", "Source", "# Samples", "Instructions Executed", "Thread Instructions Executed", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_mio", "stall_lg", "stall_barrier", "stall_branch_resolving")}
body = rows[hdr + 1:]
base = int(body[0][ci["Address"]], 16)
agg = defaultdict(lambda: [0, 0, 0, defaultdict(int)])
tot_i = tot_s = tot_t = 0
for r in body:
    if len(r) < len(H):
        continue
    off = int(r[ci["Address"]], 16) - base
    key, _ = line_of.get(off, (("?", 0), ""))
    n_i, n_s, n_t = int(r[ci["Instructions Executed"]]), int(r[ci["# Samples"]]), int(r[ci["Thread Instructions Executed"]])
    a = agg[key]
    a[0] += n_i; a[1] += n_s; a[2] += n_t
    for st in ("stall_long_sb", "stall_short_sb", "stall_wait", "stall_mio", "stall_lg", "stall_barrier", "stall_branch_resolving"):
        a[3][st] += int(r[ci[st]])
    tot_i += n_i; tot_s += n_s; tot_t += n_t
print(f"total warp instructions {tot_i}  samples {tot_s}  avg active lanes {tot_t / max(tot_i, 1):.1f}")
src_cache = {}
def src(key):
    f, n = key
    for d in (ROOT / "openvino_tokenizers_b200/csrc",):
        p = d / f
        if p.exists():
            if p not in src_cache:
                src_cache[p] = p.read_text().splitlines()
            L = src_cache[p]
            return L[n - 1].strip()[:110] if 0 < n <= len(L) else ""
    return ""
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ", ".join(f"{k[6:]}={v}" for k, v in sorted(a[3].items(), key=lambda kv: -kv[1])[:3] if v)
    print(f"{a[0] * 100 / tot_i:5.1f}% inst {a[1] * 100 / max(tot_s, 1):5.1f}% smp lanes {a[2] / max(a[0], 1):4.1f}  {key[0]}:{key[1]:<5d} {src(key)}   [{st}]")
if len(sys.argv) > 4:      # region summary: "name:file:lo-hi,..."
    print("--- regions")
    regs = []
    for spec in sys.argv[4].split(","):
        nm, f, rng = spec.split(":")
        lo, hi = map(int, rng.split("-"))
        regs.append((nm, f, lo, hi))
    acc = defaultdict(lambda: [0, 0, 0])
    for key, a in agg.items():
        nm = next((r[0] for r in regs if r[1] == key[0] and r[2] <= key[1] <= r[3]), "other:" + key[0])
        acc[nm][0] += a[0]; acc[nm][1] += a[1]; acc[nm][2] += a[2]
    for nm, a in sorted(acc.items(), key=lambda kv: -kv[1][0]):
        print(f"{a[0] * 100 / tot_i:5.1f}% inst {a[1] * 100 / max(tot_s, 1):5.1f}% smp lanes {a[2] / max(a[0], 1):4.1f}  {nm}")
