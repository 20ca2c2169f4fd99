#!/bin/bash
# ncu --set full of the registration kernel alone (maps built once, second batched match captured): scripts/prof_match.sh <tag>
TAG=${1:-r02}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:match_kernel -s 2 -c 2 -f -o $OUT/match_$TAG python scripts/bench_match.py 592 1 > $OUT/match_$TAG.log 2>&1
tail -2 $OUT/match_$TAG.log
ncu -i $OUT/match_$TAG.ncu-rep --page raw --csv > $OUT/match_${TAG}_raw.csv 2>/dev/null
ncu -i $OUT/match_$TAG.ncu-rep --page source --csv > $OUT/match_${TAG}_source.csv 2>/dev/null
LIB=${NDTB_LIB:-ndt_feature_graph_b200/lib/libndtb.so}
rm -rf /tmp/cubx; mkdir /tmp/cubx; (cd /tmp/cubx && cuobjdump -xelf all $OLDPWD/$LIB > /dev/null 2>&1; for f in *.cubin; do nvdisasm -g $f >> all.sass 2>/dev/null; done)
python scripts/ncu_by_line.py $OUT/match_${TAG}_source.csv /tmp/cubx/all.sass match_kernel 0 45 > $OUT/match_${TAG}_by_line.txt 2>&1
python scripts/ncu_by_line.py $OUT/match_${TAG}_source.csv /tmp/cubx/all.sass match_kernel 1 25 >> $OUT/match_${TAG}_by_line.txt 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('$OUT/match_${TAG}_raw.csv')))
h=rows[0]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__inst_executed.sum','sm__cycles_elapsed.max','smsp__cycles_active.avg','sm__throughput.avg.pct_of_peak_sustained_elapsed','launch__shared_mem_per_block_dynamic','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active','l1tex__throughput.avg.pct_of_peak_sustained_active','smsp__warps_eligible.avg.per_cycle_active','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__average_warp_latency_issue_stalled_wait.ratio','local_load_requests','l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_local_op_st.sum']
for w in want:
    if w in h:
        i=h.index(w); print(w, rows[1][i], [r[i] for r in rows[2:]])
PY
rm -f $OUT/match_${TAG}_source.csv
gzip -f $OUT/match_${TAG}_raw.csv
ls -la $OUT | tail -5
