#!/bin/bash
# ncu captures for profiles/r02 (run under gpurun, 1 GPU).  Each capture follows a plain run of the same command.
#   tools/ncu_capture.sh <tag> [rounds]
set -u
TAG=${1:-r02}
ROUNDS=${2:-10}
OUT=gpurun_out
CMD="python tools/ncu_workload.py --rounds $ROUNDS"
METRICS=gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_xu.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active,sm__issue_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,launch__grid_size,launch__block_size,sm__warps_active.avg.pct_of_peak_sustained_active
$CMD > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/${TAG}_plain.log; exit 1; }
ncu --metrics $METRICS --clock-control none --csv --log-file $OUT/${TAG}_metrics.csv $CMD > $OUT/${TAG}_ncu_metrics.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'rollout_cost_kernel|weight_philox_kernel|step_tp_kernel|step_fused_kernel|weighted_noise_kernel' \
    -o $OUT/${TAG}_full $CMD > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT | tail -8
