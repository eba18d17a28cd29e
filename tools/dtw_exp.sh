#!/bin/bash
# DTW kernel A/B on the GPU box.  VCB_DTW_STREAM: 1 persistent stream kernel (default), 0 barrier kernel.
# Usage: tools/dtw_exp.sh "1 0" [ncu]    Logs: gpurun_out/dtw_exp.log
mkdir -p gpurun_out
L=gpurun_out/dtw_exp.log
: > $L
modes="${1:-1 0}"
for pipe in $modes; do
  echo "== parity, VCB_DTW_STREAM=$pipe" >> $L
  VCB_DTW_STREAM=$pipe timeout -k 10 300 python -m pytest tests/test_gpu_dtw.py -m gpu -q --tb=short -x 2>&1 | tail -n 15 >> $L
done
for pipe in $modes $modes; do
  echo "== timing, VCB_DTW_STREAM=$pipe" >> $L
  VCB_DTW_STREAM=$pipe timeout -k 10 300 python tools/run_path.py dtw 10 2>&1 | tail -n 2 >> $L
done
M=gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,lts__t_sectors_srcunit_tex.sum,l1tex__t_sector_hit_rate.pct,launch__registers_per_thread,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
if [ -n "$2" ]; then
for pipe in $modes; do
  echo "== ncu, VCB_DTW_STREAM=$pipe" >> $L
  VCB_DTW_STREAM=$pipe timeout -k 10 600 ncu --clock-control none --metrics $M -k regex:"dtw_(fused|pipe|stream)" -c 1 python tools/run_path.py dtw 1 2>&1 | grep -vE "^==PROF==" | tail -n 32 >> $L
done
fi
cat $L
