#!/bin/bash
# One GPU-box round: parity tests, bench, engine A/B.  Outputs -> gpurun_out/
OUT=gpurun_out; mkdir -p $OUT
(timeout 900 python -m pytest tests -m gpu -x -q > $OUT/tests.log 2>&1; echo "tests exit $?" >> $OUT/tests.log)
tail -15 $OUT/tests.log
(timeout 600 python bench.py ${BENCH_ARGS:-} > $OUT/bench.log 2>&1; echo "bench exit $?" >> $OUT/bench.log)
tail -2 $OUT/bench.log
for c in -1 1; do CTAS=$c timeout 300 python scripts/bench_match.py 296 3 2>&1 | tail -1; done > $OUT/engine_ab.log
for n in 1 8 32 74 148; do CTAS=-1 timeout 300 python scripts/bench_match.py $n 3 2>&1 | tail -1; CTAS=0 timeout 300 python scripts/bench_match.py $n 3 2>&1 | tail -1; done >> $OUT/engine_ab.log
cat $OUT/engine_ab.log
