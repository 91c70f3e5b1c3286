#!/usr/bin/env python
"""Per-kernel DRAM traffic / instruction counts of ONE step from an ncu metrics pass
(`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,...,gpu__time_duration.sum --csv --log-file X.csv
python bench.py --workload W --steps 2 --warmup 3 --device-only`): the launches of the LAST step (everything after the last L2-flush
fill kernel) are listed in order with their DRAM bytes, warp instructions and duration, and summed — the whole-step traffic figure
bench.py reports as roofline.step_traffic, and the instruction-issue roofline of the dominant kernel.

    python tools/step_metrics.py gpurun_out/stepmetrics_c1.csv [out.txt] [--json KEY profiles/roofline_traffic.json]
"""
import csv
import json
import sys

SM_COUNT, SMSP_PER_SM = 148, 4        # B200: one warp instruction per SMSP per cycle is the issue peak


def load(path):
    launches = {}
    with open(path, newline="") as fh:
        for r in csv.reader(fh):
            if len(r) < 15 or not r[0].isdigit():
                continue
            d = launches.setdefault(int(r[0]), {"name": r[4]})
            try:
                d[r[12]] = float(r[14].replace(",", ""))
            except ValueError:
                pass
            d["unit:" + r[12]] = r[13]
    return [launches[k] for k in sorted(launches)]


def ns(d):
    v, u = d.get("gpu__time_duration.sum", 0.0), d.get("unit:gpu__time_duration.sum", "ns")
    return v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6}.get(u, 1.0)


def nbytes(d, key):
    v, u = d.get(key, 0.0), d.get("unit:" + key, "byte")
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)


def main():
    path = sys.argv[1]
    out = open(sys.argv[2], "w") if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else sys.stdout
    L = load(path)
    fills = [i for i, d in enumerate(L) if "FillFunctor<unsigned char>" in d["name"]]
    step = L[fills[-1] + 1:] if fills else L
    step = [d for d in step if "at::" not in d["name"]]
    tot_t = sum(ns(d) for d in step)
    tot_r = sum(nbytes(d, "dram__bytes_read.sum") for d in step)
    tot_w = sum(nbytes(d, "dram__bytes_write.sum") for d in step)
    tot_i = sum(d.get("smsp__inst_executed.sum", 0.0) for d in step)
    print(f"# {path}: the launches of one step in order (ncu: serialised, cold caches — shares, not absolutes)", file=out)
    print("      us  share   DRAM rd MB  DRAM wr MB  warp inst M  issue%  lanes  kernel", file=out)
    for d in step:
        print(f"{ns(d) / 1e3:8.1f} {100 * ns(d) / tot_t:5.1f}% {nbytes(d, 'dram__bytes_read.sum') / 1e6:11.2f} {nbytes(d, 'dram__bytes_write.sum') / 1e6:11.2f} "
              f"{d.get('smsp__inst_executed.sum', 0) / 1e6:12.2f} {d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0):6.1f} "
              f"{d.get('smsp__thread_inst_executed_per_inst_executed.ratio', 0):6.1f}  {d['name'][:100]}", file=out)
    print(f"{tot_t / 1e3:8.1f} 100.0% {tot_r / 1e6:11.2f} {tot_w / 1e6:11.2f} {tot_i / 1e6:12.2f}                whole step ({len(step)} launches)", file=out)
    dom = max(step, key=ns)
    inst = dom.get("smsp__inst_executed.sum", 0.0)
    print(f"\ndominant kernel: {dom['name'][:100]}", file=out)
    print(f"  DRAM traffic {nbytes(dom, 'dram__bytes_read.sum') / 1e6:.2f} MB read + {nbytes(dom, 'dram__bytes_write.sum') / 1e6:.2f} MB written per launch", file=out)
    print(f"  instruction-issue roofline: {inst / 1e6:.1f} M warp instructions / ({SM_COUNT} SMs x {SMSP_PER_SM} issue slots) = "
          f"{inst / (SM_COUNT * SMSP_PER_SM):.0f} issue cycles minimum; at the SM clock of the bench run this is the floor of the kernel's time "
          f"(1.965 GHz: {inst / (SM_COUNT * SMSP_PER_SM) / 1.965e3:.1f} us; ncu duration of this launch {ns(dom) / 1e3:.1f} us)", file=out)
    if "--json" in sys.argv:
        key, jp = sys.argv[sys.argv.index("--json") + 1], sys.argv[sys.argv.index("--json") + 2]
        try:
            J = json.load(open(jp))
        except Exception:
            J = {}
        J[key] = {"kernel": dom["name"][:80], "kernel_dram_bytes": int(nbytes(dom, "dram__bytes_read.sum") + nbytes(dom, "dram__bytes_write.sum")),
                  "kernel_dram_read": int(nbytes(dom, "dram__bytes_read.sum")), "kernel_dram_write": int(nbytes(dom, "dram__bytes_write.sum")),
                  "kernel_warp_inst": int(inst), "step_dram_bytes": int(tot_r + tot_w), "step_dram_read": int(tot_r), "step_dram_write": int(tot_w),
                  "step_warp_inst": int(tot_i), "step_launches": len(step), "source": path.split("/")[-1]}
        json.dump(J, open(jp, "w"), indent=1)


if __name__ == "__main__":
    main()
