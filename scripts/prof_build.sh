#!/bin/bash
# ncu --set full of the map-build kernels of one bench step (with source counters): scripts/prof_build.sh <tag>
TAG=${1:-r02}
OUT=gpurun_out; mkdir -p $OUT
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-extra --lanes 1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^k_(centroid_chunks|centroid|extent|mark|blockscan|count|rs_hist|rs_scan|rs_scatter|seg_bounds|cells|eigen|eigen_hard|gscan|gfill)$' -s 54 -c 18 -f -o $OUT/build_$TAG $BENCH > $OUT/build_$TAG.log 2>&1
tail -2 $OUT/build_$TAG.log
ncu -i $OUT/build_$TAG.ncu-rep --page raw --csv > $OUT/build_${TAG}_raw.csv 2>/dev/null
ncu -i $OUT/build_$TAG.ncu-rep --page source --csv > $OUT/build_${TAG}_source.csv 2>/dev/null
LIB=${NDTB_LIB:-ndt_feature_graph_b200/lib/libndtb.so}
rm -rf /tmp/cubx; mkdir /tmp/cubx; (cd /tmp/cubx && cuobjdump -xelf all $OLDPWD/$LIB > /dev/null 2>&1; for f in *.cubin; do nvdisasm -g $f >> all.sass 2>/dev/null; done)
python - <<PY
import csv
rows=list(csv.reader(open('$OUT/build_${TAG}_raw.csv')))
h=rows[0]; ik=h.index('Kernel Name')
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_active','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__inst_executed.sum']
for r in rows[2:]:
    print(r[ik].split('(')[0])
    for w in want:
        if w in h: print('   ', w, rows[1][h.index(w)], r[h.index(w)])
    st=[(float(r[i].replace(',','')),c) for i,c in enumerate(h) if c.startswith('smsp__average_warps_issue_stalled') and c.endswith('per_issue_active.ratio') and 'not_issued' not in c and r[i] not in ('','n/a')]
    st.sort(reverse=True)
    print('    stalls', [(round(v,2),c.split('stalled_')[1].split('_per')[0]) for v,c in st[:6]])
PY
gzip -dc < /dev/null; for spec in "k_extent 2" "k_mark 3" "k_count 5" "k_rs_scatter 8" "k_cells 13" "k_gfill 17"; do set -- $spec; echo "== $1"; python scripts/ncu_by_line.py $OUT/build_${TAG}_source.csv /tmp/cubx/all.sass $1 $2 14; done > $OUT/build_${TAG}_by_line.txt 2>&1
gzip -f $OUT/build_${TAG}_raw.csv
gzip -f $OUT/build_${TAG}_source.csv
rm -f $OUT/build_$TAG.ncu-rep
