#!/bin/bash
# One GPU-box session: smoke, GPU tests, bench lines, launch list, full ncu capture of the dominant kernel.  Everything is
# wrapped in `timeout` so that a hung kernel cannot hold the box.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi_$TAG.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== ordered tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "ordered or edge or c0 or c1_slice" 2>&1 | tail -15
echo "== bench c1"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c1_$TAG.json 2> gpurun_out/bench_c1_$TAG.err; python tools/bench_brief.py gpurun_out/bench_c1_$TAG.json; tail -5 gpurun_out/bench_c1_$TAG.err
if [ -z "$SKIP_TESTS" ]; then echo "== gpu tests"; timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -4 gpurun_out/pytest_gpu_$TAG.log; fi
echo "== bench c3"; timeout 600 python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err; python tools/bench_brief.py gpurun_out/bench_c3_$TAG.json
if [ -z "$SKIP_C2" ]; then echo "== bench c2"; timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err; python tools/bench_brief.py gpurun_out/bench_c2_$TAG.json; fi
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c1_$TAG.csv python bench.py --steps 3 --warmup 3 --device-only > gpurun_out/b_ncu_$TAG.log 2>&1; tail -2 gpurun_out/b_ncu_$TAG.log | cut -c1-300
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:gpt2_bpe_fast_kernel -s 4 -c 1 -o gpurun_out/prof_${TAG}_fast -f python bench.py --steps 3 --warmup 3 --device-only > gpurun_out/b_ncu2_$TAG.log 2>&1; ls -la gpurun_out/prof_${TAG}_fast.ncu-rep
