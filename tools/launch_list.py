#!/usr/bin/env python
"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv <command>`):
per-kernel totals with their share of the GPU time, and optionally the launches of one call in order.

    python tools/launch_list.py gpurun_out/launches.csv [out.txt] [--first KERNEL_SUBSTRING --count N]
"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    out = open(sys.argv[2], "w") if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else sys.stdout
    rows = []
    with open(path, newline="") as fh:
        for r in csv.reader(fh):
            if len(r) > 10 and r[0].isdigit():
                try:
                    rows.append((r[4], float(r[-1].replace(",", "")), r[-2]))
                except ValueError:
                    pass
    if not rows:
        print("no launches found", file=out)
        return
    unit = rows[0][2]
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3 if max(t for _, t, _ in rows) > 1e5 else 1.0)
    tot = collections.defaultdict(lambda: [0, 0.0])
    for name, t, _ in rows:
        tot[name][0] += 1
        tot[name][1] += t * scale
    total = sum(v[1] for v in tot.values())
    print(f"# {len(rows)} launches, {total:.1f} us of GPU time (ncu: serialised, cold cache — compare shares, not absolutes)", file=out)
    print(" count   total us   share  kernel", file=out)
    for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{n:6d} {t:10.1f} {100 * t / total:6.1f}%  {name[:110]}", file=out)
    if "--first" in sys.argv:
        key = sys.argv[sys.argv.index("--first") + 1]
        cnt = int(sys.argv[sys.argv.index("--count") + 1]) if "--count" in sys.argv else 12
        idx = next((i for i, (n, _, _) in enumerate(rows) if key in n), None)
        if idx is not None:
            seg = rows[idx:idx + cnt]
            st = sum(t * scale for _, t, _ in seg)
            print(f"\n## {cnt} launches in order from the first `{key}`", file=out)
            for n, t, _ in seg:
                print(f"{t * scale:9.2f} {100 * t * scale / st:6.1f}%  {n[:110]}", file=out)


if __name__ == "__main__":
    main()
