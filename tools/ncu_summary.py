#!/usr/bin/env python
"""Summarise an .ncu-rep (captured under gpurun with --set full --import-source on) into a short text file:
key launch metrics, stall mix, and instruction share per source line (top N).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [profiles/out.txt] [--top 40]
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    out_path = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else None
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    lines = [f"# ncu summary of {rep}"]
    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        name = row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        lines.append(f"\n## kernel: {name[:100]}")
        for i, h in enumerate(hdr):
            if h in KEYS or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
                try:
                    v = float(row[i])
                except ValueError:
                    continue
                if "issue_stalled" in h and v < 0.05:
                    continue
                lines.append(f"{h:90s} {row[i]:>16s} {units[i]}")
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    hdr = next((r for r in src if r and r[0] == "Line No"), None)
    if hdr:
        i_inst, i_thr, i_samp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        agg = collections.defaultdict(lambda: [0, 0, 0, ""])
        fname = "?"
        for r in src:
            if len(r) >= 2 and r[0] == "File Name":
                fname = r[1].split("/")[-1]
                continue
            try:
                ln, inst, thr, samp = int(r[0]), int(r[i_inst] or 0), int(r[i_thr] or 0), int(r[i_samp] or 0)
            except (ValueError, IndexError):
                continue
            a = agg[(r[1].strip()[:100], ln)]
            a[0] += inst; a[1] += thr; a[2] += samp
        tot = sum(v[0] for v in agg.values()) or 1
        tots = sum(v[2] for v in agg.values()) or 1
        lines.append(f"\n## instruction share per source line (total warp instructions {tot}, samples {tots})")
        lines.append("  inst%  samp%  thr/inst  line  source")
        for (text, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
            lines.append(f"  {100 * v[0] / tot:5.1f}  {100 * v[2] / tots:5.1f}  {v[1] / max(v[0], 1):8.1f}  {ln:4d}  {text}")
    text = "\n".join(lines) + "\n"
    if out_path:
        open(out_path, "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
