"""bench.py contract, CPU tier: the reference arm (`--impl reference`: the reference's own op code from oracle/_ref, or the oracle port, on the host cores, the one leg of
bench.py that needs no GPU) prints ONE JSON line with the keys the driver reads."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_the_contract_line():
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] and d["unit"] == "MB/s" and d["value"] > 0
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["ms_per_step"] > 0
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "u8" and d["data"] == "synthetic" and d["config"]["workload"].startswith("C1")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"] and cb["unit"] == "MB/s"
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == "MB/s" and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_is_defined_for_the_tokenize_workloads_only():
    for w in ("c4", "norm"):
        p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", w], capture_output=True, text=True,
                           timeout=120, cwd=str(ROOT))
        assert p.returncode == 0
        d = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
        assert d["impl"] == "reference" and "unavailable" in d
