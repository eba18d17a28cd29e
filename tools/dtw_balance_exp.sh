mkdir -p gpurun_out; L=gpurun_out/dtw_balance.log; : > $L
M=gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
for r in 576,576 608,608 640,640 672,672 550,650; do
  echo "== S_RANGE=$r" >> $L
  S_RANGE=$r timeout 300 python tools/run_path.py dtw 5 2>&1 | tail -n 1 >> $L
  S_RANGE=$r timeout 300 ncu --clock-control none --metrics $M -k regex:dtw_stream -s 1 -c 1 python tools/run_path.py dtw 1 2>&1 | grep -E "duration|fp64|long_score" >> $L
done
cat $L
