#!/bin/bash
# Runs on the GPU box (under gpurun): launch list + ncu captures of one bench step.  Outputs -> gpurun_out/ (CSV exports;
# only the match-kernel report itself is kept, gpurun brings back at most 64 MiB).
# usage: scripts/profile_gpu.sh <tag> [pairs]
set -u
TAG=${1:-r01}
PAIRS=${2:-592}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 1 --warmup 1 --pairs $PAIRS --no-cpu --no-e2e --no-extra --lanes 1"
# 1) every launch of warm-up + timed step with its device time
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $BENCH > $OUT/launches_$TAG.log 2>&1
# 2) full capture of the registration kernel: first (1 CTA per registration) and finishing (8-CTA clusters) launch of a step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:match_kernel -s 2 -c 2 -f -o $OUT/match_$TAG $BENCH > $OUT/match_$TAG.log 2>&1
ncu -i $OUT/match_$TAG.ncu-rep --page raw --csv > $OUT/match_${TAG}_raw.csv 2>/dev/null
ncu -i $OUT/match_$TAG.ncu-rep --page source --csv > $OUT/match_${TAG}_source.csv 2>/dev/null
# 3) lighter capture of the map-build + covariance kernels of one step
timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats \
  --clock-control none -k regex:'k_|cov_' -s 20 -c 20 -f -o $OUT/build_$TAG $BENCH > $OUT/build_$TAG.log 2>&1
ncu -i $OUT/build_$TAG.ncu-rep --page raw --csv > $OUT/build_${TAG}_raw.csv 2>/dev/null
rm -f $OUT/build_$TAG.ncu-rep
# 4) SASS evidence of the TMA bulk copy + mbarrier in the registration kernel
cuobjdump -sass ndt_feature_graph_b200/lib/libndtb.so 2>/dev/null | grep -n -E "Function : .*match_kernel|UBLKCP|SYNCS|MBARRIER|ARRIVES" | head -40 > $OUT/match_${TAG}_sass_tma.txt
gzip -f $OUT/match_${TAG}_source.csv
ls -la $OUT
du -sh $OUT
