#!/bin/bash
# Short GPU session for a kernel change: full GPU tests, C1 / C3 bench lines, warp-instruction + DRAM counts of one step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-q}
WLS=${2:-"c1 c3"}
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -4 gpurun_out/pytest_gpu_$TAG.log
M=dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum
for wl in $WLS; do
  echo "== bench $wl"; timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-unfused > gpurun_out/bench_${wl}_$TAG.json 2> gpurun_out/bench_${wl}_$TAG.err; python tools/bench_brief.py gpurun_out/bench_${wl}_$TAG.json; tail -3 gpurun_out/bench_${wl}_$TAG.err
  timeout 600 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/stepmetrics_${wl}_$TAG.csv python bench.py --workload $wl --steps 2 --warmup 3 --device-only > gpurun_out/b_ncu3_$TAG.log 2>&1
  python tools/step_metrics.py gpurun_out/stepmetrics_${wl}_$TAG.csv | tail -12
done
if [ -n "$EXTRA" ]; then
  for wl in c4 norm; do
    timeout 600 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/stepmetrics_${wl}_$TAG.csv python bench.py --workload $wl --steps 2 --warmup 3 --device-only > gpurun_out/b_ncu3_$TAG.log 2>&1
    python tools/step_metrics.py gpurun_out/stepmetrics_${wl}_$TAG.csv | tail -8
  done
  echo "== pcie duplex probe"; timeout 120 python tools/pcie_duplex_probe.py 2>&1 | tail -4 | tee gpurun_out/pcie_duplex_$TAG.txt
fi
