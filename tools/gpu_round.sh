#!/bin/bash
# One GPU-box session: smoke, GPU tests, bench lines, launch list, per-kernel traffic / instruction counts of one step, full ncu
# capture of the dominant kernel.  Everything is wrapped in `timeout` so that a hung kernel cannot hold the box.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi_$TAG.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
if [ -z "$SKIP_TESTS" ]; then echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -14 gpurun_out/pytest_gpu_$TAG.log; fi
echo "== bench c1"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c1_$TAG.json 2> gpurun_out/bench_c1_$TAG.err; python tools/bench_brief.py gpurun_out/bench_c1_$TAG.json; tail -5 gpurun_out/bench_c1_$TAG.err
echo "== bench c3"; timeout 600 python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err; python tools/bench_brief.py gpurun_out/bench_c3_$TAG.json
if [ -z "$SKIP_C2" ]; then echo "== bench c2"; timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err; python tools/bench_brief.py gpurun_out/bench_c2_$TAG.json; fi
if [ -z "$SKIP_C4" ]; then echo "== bench c4"; timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 > gpurun_out/bench_c4_$TAG.json 2> gpurun_out/bench_c4_$TAG.err; python tools/bench_brief.py gpurun_out/bench_c4_$TAG.json; fi
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c1_$TAG.csv python bench.py --steps 3 --warmup 3 --device-only > gpurun_out/b_ncu_$TAG.log 2>&1; tail -2 gpurun_out/b_ncu_$TAG.log | cut -c1-300
M=dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum
for wl in c1 c2 c3 c4 norm; do
  echo "== step metrics $wl"; timeout 600 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/stepmetrics_${wl}_$TAG.csv python bench.py --workload $wl --steps 2 --warmup 3 --device-only > gpurun_out/b_ncu3_$TAG.log 2>&1; tail -1 gpurun_out/b_ncu3_$TAG.log | cut -c1-200
done
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:gpt2_bpe_fast_kernel -s 4 -c 1 -o gpurun_out/prof_${TAG}_fast -f python bench.py --steps 3 --warmup 3 --device-only > gpurun_out/b_ncu2_$TAG.log 2>&1; ls -la gpurun_out/prof_${TAG}_fast.ncu-rep
if [ -n "$NCU_C2" ]; then timeout 900 ncu --set full --clock-control none --import-source on -k regex:rows_kernel -s 4 -c 1 -o gpurun_out/prof_${TAG}_c2 -f python bench.py --workload c2 --steps 3 --warmup 3 --device-only > gpurun_out/b_ncu4_$TAG.log 2>&1; ls -la gpurun_out/prof_${TAG}_c2.ncu-rep; fi
