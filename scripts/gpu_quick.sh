#!/bin/bash
# quick match-kernel A/B on the GPU box: cluster width of the first launch x pass budget
OUT=gpurun_out; mkdir -p $OUT
for n in ${SIZES:-592}; do for c in ${ENGINES:-1 2}; do for b in ${BUDGETS:-0}; do echo -n "CTAS=$c BUDGET=$b "; CTAS=$c BUDGET=$b timeout 300 python scripts/bench_match.py $n 3 2>&1 | tail -1 | cut -c1-330; done; done; done > $OUT/engine_ab.log
cat $OUT/engine_ab.log
