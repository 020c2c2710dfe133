#!/bin/bash
# usage (on the GPU box, through gpurun): tools/ncu_full.sh <out-name> <kernel-regex> <tune.py args...>
# One `--set full` capture (+ the SASS FP-op counters and L2 bytes SURVEY §8d asks for) of the first matching launch.
out=$1; shift; re=$1; shift
X=smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fp32_pred_on.sum,lts__t_bytes.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_red.sum,lts__t_sectors_op_write.sum,smsp__thread_inst_executed_per_inst_executed.ratio
timeout -s KILL 600 ncu --set full --metrics $X --clock-control none --import-source on -k regex:$re -c 1 -f -o gpurun_out/$out python tools/tune.py "$@" 2>&1 | tail -3
