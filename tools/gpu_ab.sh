#!/bin/bash
# A/B of environment switches on the C1 / C3 bench, all on ONE box (numbers from different boxes differ by several percent):
#   tools/gpu_ab.sh TAG "VAR=a VAR2=b" "VAR=c" ...
cd "$(dirname "$0")/.."
TAG=$1; shift
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "ordered or edge or c0 or c1_slice or c3_slice or ragged" 2>&1 | tail -3
i=0
for kv in "$@"; do
  i=$((i+1))
  name="v$i"
  echo "== $name: $kv"
  env $kv timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-unfused > "gpurun_out/bench_c1_${TAG}_$name.json" 2> "gpurun_out/bench_c1_${TAG}_$name.err"
  python tools/bench_brief.py "gpurun_out/bench_c1_${TAG}_$name.json" | cut -c1-260
  if [ -z "$SKIP_C3" ]; then
  env $kv timeout 600 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --no-unfused > "gpurun_out/bench_c3_${TAG}_$name.json" 2>> "gpurun_out/bench_c1_${TAG}_$name.err"
  python tools/bench_brief.py "gpurun_out/bench_c3_${TAG}_$name.json" | cut -c1-260
  fi
done
