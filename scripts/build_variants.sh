#!/bin/bash
# Builds kernel variants of libndtb.so for A/B runs on the GPU box:  scripts/build_variants.sh name "-DFLAG=.. -DFLAG2=.." ...
set -e
cd "$(dirname "$0")/../ndt_feature_graph_b200/csrc"
OUT=../lib/variants
mkdir -p $OUT
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="$ARCH -ccbin /usr/bin/g++ -O3 -std=c++17 -lineinfo -Xcompiler -fPIC"
[ -f $OUT/map_build.o ] && [ $OUT/map_build.o -nt map_build.cu ] || nvcc $COMMON -fmad=false -c map_build.cu -o $OUT/map_build.o
[ -f $OUT/api.o ] && [ $OUT/api.o -nt api.cu ] || nvcc $COMMON -c api.cu -o $OUT/api.o
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc $COMMON $flags -Xptxas -v -c d2d.cu -o $OUT/d2d_$name.o 2> $OUT/d2d_$name.log
  nvcc $ARCH -ccbin /usr/bin/g++ -shared -o $OUT/libndtb_$name.so $OUT/map_build.o $OUT/d2d_$name.o $OUT/api.o
  echo "$name: $(grep -A2 match_kernel $OUT/d2d_$name.log | grep -E -o 'Used [0-9]+ registers|[0-9]+ bytes spill stores' | tr '\n' ' ')"
done
