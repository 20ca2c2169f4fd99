#!/bin/bash
# Builds kernel variants of libndtb.so for A/B runs on the GPU box (select with NDTB_LIB=.../libndtb_<name>.so):
#   scripts/build_variants.sh <d2d|map_build> name "-DFLAG=.." [name "-DFLAG=.."] ...
set -e
cd "$(dirname "$0")/../ndt_feature_graph_b200/csrc"
UNIT=$1; shift
OUT=../lib/variants
mkdir -p $OUT
make -s
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="$ARCH -ccbin /usr/bin/g++ -O3 -std=c++17 -lineinfo -Xcompiler -fPIC"
EXTRA=""; [ "$UNIT" = map_build ] && EXTRA="-fmad=false"
OBJS=""; for u in map_build d2d api fuser jff formats; do [ $u != $UNIT ] && OBJS="$OBJS ../lib/obj/$u.o"; done
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc $COMMON $EXTRA $flags -Xptxas -v -c $UNIT.cu -o $OUT/${UNIT}_$name.o 2> $OUT/${UNIT}_$name.log
  nvcc $ARCH -ccbin /usr/bin/g++ -shared -o $OUT/libndtb_$name.so $OBJS $OUT/${UNIT}_$name.o -ldl
  echo "$name: built"
done
