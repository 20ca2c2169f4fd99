#!/bin/bash
# Runs on the GPU box (under gpurun): launch list + full ncu captures of one bench step.  Outputs -> gpurun_out/
# usage: scripts/profile_gpu.sh <tag> [pairs]
set -u
TAG=${1:-r01}
PAIRS=${2:-296}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 1 --warmup 1 --pairs $PAIRS --no-cpu --no-e2e"
# 1) every launch of warm-up + timed step with its device time
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $BENCH > $OUT/launches_$TAG.log 2>&1
# 2) full capture of the registration kernel (timed step = 2nd and later launches)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:match_kernel -s 1 -c 2 -f -o $OUT/match_$TAG $BENCH > $OUT/match_$TAG.log 2>&1
# 3) full capture of the map-build + covariance kernels of the timed step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^k_|cov_' -s 13 -c 13 -f -o $OUT/build_$TAG $BENCH > $OUT/build_$TAG.log 2>&1
ls -la $OUT
