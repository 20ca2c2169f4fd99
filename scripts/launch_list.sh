#!/bin/bash
# per-kernel device times of one bench step (ncu launch list); usage: scripts/launch_list.sh <tag>
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-extra --lanes 1 > $OUT/launches_$TAG.log 2>&1
python - <<PY
import csv
from collections import OrderedDict
rows=[r for r in csv.reader(l for l in open('$OUT/launches_$TAG.csv') if l.startswith('"'))]
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value')
names=[r[ik].split('(')[0].replace('ndtb::','') for r in rows[1:]]
vals=[float(r[iv])/1e6 for r in rows[1:]]
starts=[i for i,n in enumerate(names) if n.endswith('k_centroid_chunks')]
s,e=starts[4],starts[5]
agg=OrderedDict()
for n,v in zip(names[s:e],vals[s:e]): agg[n]=agg.get(n,0)+v
tot=sum(agg.values())
for k,v in agg.items(): print(f"{k:32s} {v:8.3f} ms {100*v/tot:5.1f}%")
print(f"total {tot:.3f} ms, {e-s} launches")
PY
