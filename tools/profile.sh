#!/bin/bash
# ncu evidence runs (GPU box, repo root).  Outputs land in gpurun_out/; summaries are copied to
# profiles/ by tools/summarize_profiles.py on the build box.   usage: tools/profile.sh r02
mkdir -p gpurun_out
R=${1:-r02}
# 1. launch list of one short bench run (cold-cache, serialised: compare SHARES only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_$R.csv python bench.py --steps 2 --warmup 3 > gpurun_out/launches_$R.log 2>&1
# 2. full captures of the hot kernels (one launch each, after warm-up launches)
cap() { # name regex skip cmd...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c 1 \
      -f -o gpurun_out/${name}_$R "$@" > gpurun_out/${name}_$R.log 2>&1
}
cap prof_fbf_tc   gmm_tc_kernel        2 python tools/run_path.py fbf 1
cap prof_traj     traj_solve_warp      1 python tools/run_path.py traj 1
cap prof_argmax   gmm_tc_kernel        1 python tools/run_path.py traj 1
cap prof_group    group_panel          1 python tools/run_path.py traj 1
cap prof_recheck  recheck_panel_kernel 1 python tools/run_path.py traj 1
cap prof_dtw      dtw_stream_kernel    1 python tools/run_path.py dtw 1
cap prof_dtwbarrier dtw_fused_kernel   1 env VCB_DTW_STREAM=0 python tools/run_path.py dtw 1
ls -la gpurun_out | grep _$R
