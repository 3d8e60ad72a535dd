#!/bin/bash
# ncu captures for profiles/r02 (run under gpurun, 1 GPU).  Each capture follows a plain run of the same command.
#   tools/ncu_capture.sh <tag> <rounds> [full]
# Writes gpurun_out/<tag>_metrics.csv (every launch: duration, FP32 op counters, pipe utilisation, DRAM bytes) and, with
# "full", gpurun_out/<tag>_full.ncu-rep (--set full + source, the first 8 launches of the step kernels) and
# <tag>_tp.ncu-rep (the time-parallel kernel).  gpurun copies back at most 64 MiB: keep the reports small.
set -u
TAG=${1:-r02}
ROUNDS=${2:-7}
FULL=${3:-}
OUT=gpurun_out
CMD="python tools/ncu_workload.py --rounds $ROUNDS"
METRICS=gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_xu.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__issue_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,launch__grid_size,launch__block_size,sm__warps_active.avg.pct_of_peak_sustained_active
$CMD > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/${TAG}_plain.log; exit 1; }
ncu --metrics $METRICS --clock-control none --csv --log-file $OUT/${TAG}_metrics.csv $CMD > $OUT/${TAG}_ncu_metrics.log 2>&1
if [ "$FULL" = "full" ]; then
  ncu --set full --clock-control none --import-source on -k regex:'rollout_cost_kernel|weight_philox_kernel|weighted_noise_kernel' -c 8 \
      -o $OUT/${TAG}_full $CMD > $OUT/${TAG}_ncu_full.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:'step_tp_kernel' -c 1 \
      -o $OUT/${TAG}_tp $CMD > $OUT/${TAG}_ncu_tp.log 2>&1
fi
ls -la $OUT | grep ${TAG}
