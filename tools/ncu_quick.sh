#!/bin/bash
# usage: tools/ncu_quick.sh <tag> [ENV=VAL ...] -- quick ncu counters of one pathtrace launch on C2 (or $CFG)
tag=$1; shift
for kv in "$@"; do export "$kv"; done
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__warps_eligible.avg.per_cycle_active,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum
ncu --metrics $M --clock-control none -k regex:pathtrace -c 1 --csv python tools/tune.py ${CFG:-C2} 2>/dev/null | python -c "
import csv,sys
rows=[r for r in csv.reader(sys.stdin) if len(r)>10]
h=rows[0]; 
for r in rows[1:]:
    d=dict(zip(h,r)); print('$tag', d['Metric Name'], d['Metric Value'])
"
