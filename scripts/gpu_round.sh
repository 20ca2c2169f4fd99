#!/bin/bash
# One GPU-box round: parity tests, bench, host-phase timing, pass-budget sweep.  Outputs -> gpurun_out/
OUT=gpurun_out; mkdir -p $OUT
(timeout 900 python -m pytest tests -m gpu -x -q > $OUT/tests.log 2>&1; echo "tests exit $?" >> $OUT/tests.log)
tail -4 $OUT/tests.log
(timeout 600 python bench.py > $OUT/bench.log 2>&1; echo "bench exit $?" >> $OUT/bench.log)
tail -2 $OUT/bench.log
(NDTB_PROFILE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu > $OUT/phases.log 2>&1)
grep "ndtb" $OUT/phases.log | tail -40
for b in 32 40 48 64 96; do BUDGET=$b timeout 300 python scripts/bench_match.py 296 3 2>&1 | tail -1; done > $OUT/budget_sweep.log
cat $OUT/budget_sweep.log
