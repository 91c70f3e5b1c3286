"""One-line digest of a bench.py JSON line (for GPU-session logs)."""
import json
import sys
for path in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(path) if l.startswith("{")][-1])
    except Exception as ex:
        print(path, "unreadable:", ex)
        continue
    r, e = d.get("roofline") or {}, d.get("e2e") or {}
    print(f"{path}: value {d.get('value', 0):.0f} {d.get('unit')} ({d.get('ms_per_step', 0):.4f} ms/step, n_gpus {d.get('n_gpus')}), kernel {r.get('kernel')} "
          f"{(r.get('kernel_ms') or 0):.4f} ms frac {(r.get('frac') or 0):.4f} share {(r.get('kernel_share_of_step') or 0):.2f}, e2e {e.get('value', 0):.0f} ({e.get('ms_per_step', 0):.3f} ms), "
          f"launches {d.get('gpu_launches')}, cpu {(d.get('cpu_baseline') or {}).get('value')}, verified {d.get('verified')}, clocks {d.get('clocks')}")
