#!/bin/bash
# One GPU-box round: parity tests, bench, launch list of one bench step.  Outputs -> gpurun_out/
OUT=gpurun_out; mkdir -p $OUT
(timeout 900 python -m pytest tests -m gpu -x -q > $OUT/tests.log 2>&1; echo "tests exit $?" >> $OUT/tests.log)
tail -15 $OUT/tests.log
(timeout 600 python bench.py ${BENCH_ARGS:-} > $OUT/bench.log 2>&1; echo "bench exit $?" >> $OUT/bench.log)
tail -2 $OUT/bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > $OUT/launches.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/launches.csv') if l.startswith('"'))]
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value')
half=(len(rows)-1)//2
tot=0
for r in rows[1+half:]:
    ms=float(r[iv])/1e6; tot+=ms
    print(f"{r[ik].split('(')[0]:40s} {ms:8.3f} ms")
print('step total', round(tot,3))
PY
