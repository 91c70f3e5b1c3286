#!/bin/bash
# Multi-GPU session: exchange correctness + timings (tools/peer_gather_check.py) and bench lines per exchange mode.
#   tools/gpu_multi.sh TAG NGPUS "mode1 mode2 ..." [workloads]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-m}; N=${2:-2}; MODES=${3:-"peer pull"}; WLS=${4:-"c1"}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ -z "$SKIP_CHECK" ]; then echo "== peer_gather_check N=$N"; B200TOK_WIRE16=1 timeout 600 $RUN --master-port 29544 tools/peer_gather_check.py 65536 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -4; fi
for wl in $WLS; do for m in $MODES; do
  echo "== bench $wl N=$N exchange=$m"
  B200TOK_BENCH_EXCHANGE=$m timeout 900 $RUN --master-port 29545 bench.py --gpus $N --workload $wl --steps 20 --warmup 5 > gpurun_out/bench_${wl}_n${N}_${m}_$TAG.json 2> gpurun_out/bench_${wl}_n${N}_${m}_$TAG.err
  python tools/bench_brief.py gpurun_out/bench_${wl}_n${N}_${m}_$TAG.json | cut -c1-420; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_${wl}_n${N}_${m}_$TAG.err | tail -3
done; done
