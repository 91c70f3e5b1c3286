#!/bin/bash
# quick instruction-count / duration probe of the rows kernel (box-independent optimisation metric)
# usage: tools/ncu_quick.sh [ascii|english] [rows] [len]
KIND=${1:-ascii}; ROWS=${2:-65536}; LEN=${3:-512}
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio \
    --clock-control none -k regex:^(rows_kernel|gpt2_bpe_fast) -s 1 -c 1 python tools/quick_bench.py $ROWS $LEN $KIND nohost 2>&1 | grep -E "inst_executed|time_duration|issue_active" 
