#!/bin/bash
# quick match-kernel A/B on the GPU box: straggler-stage cluster width x pass budget
OUT=gpurun_out; mkdir -p $OUT
for n in ${SIZES:-592}; do for g in ${G2S:-0 8 4}; do for b in ${BUDGETS:-0}; do echo -n "G2=$g BUDGET=$b "; if [ $g != 0 ]; then export NDTB_G2=$g; else unset NDTB_G2; fi; BUDGET=$b timeout 300 python scripts/bench_match.py $n 3 2>&1 | tail -1 | cut -c1-330; done; done; done > $OUT/engine_ab.log
cat $OUT/engine_ab.log
