#!/bin/bash
# quick engine A/B on the GPU box: parity tests + match micro-benchmark at a few batch sizes
OUT=gpurun_out; mkdir -p $OUT
(timeout 900 python -m pytest tests -m gpu -x -q > $OUT/tests.log 2>&1; echo "tests exit $?" >> $OUT/tests.log)
tail -5 $OUT/tests.log
for n in ${SIZES:-296 1 32 148}; do for c in ${ENGINES:--1 1}; do CTAS=$c timeout 300 python scripts/bench_match.py $n 3 2>&1 | tail -1 | cut -c1-330; done; done > $OUT/engine_ab.log
cat $OUT/engine_ab.log
