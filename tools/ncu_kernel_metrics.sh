#!/bin/bash
# One ncu pass per bench config: executed fp instructions, fp64-pipe cycles, DRAM bytes and duration of every pair /
# dense kernel launch.  Output: gpurun_out/r02_ncu_<cfg>.csv -> tools/gen_kernel_metrics.py -> profiles/kernel_metrics.json
# usage (on the GPU box, ONE ncu tool per gpurun call): bash tools/ncu_kernel_metrics.sh c3 c2 c4 ...
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
M=$M,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum
M=$M,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum
M=$M,smsp__thread_inst_executed_pipe_fp64_pred_on.sum,smsp__thread_inst_executed_pipe_xu_pred_on.sum,smsp__pipe_fp64_cycles_active.sum,smsp__cycles_elapsed.sum
M=$M,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active
mkdir -p gpurun_out
for cfg in "$@"; do
  A="--config $cfg --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
  python bench.py $A > gpurun_out/r02_plain_$cfg.log 2>&1 || { echo "plain run of $cfg failed"; tail -5 gpurun_out/r02_plain_$cfg.log; continue; }
  ncu --metrics $M --clock-control none -k regex:'ks_' --csv --log-file gpurun_out/r02_ncu_$cfg.csv python bench.py $A > gpurun_out/r02_ncu_$cfg.log 2>&1
  echo "$cfg: ncu rc=$? rows=$(wc -l < gpurun_out/r02_ncu_$cfg.csv)"
done
