#!/bin/bash
# ncu evidence runs (GPU box, repo root).  Outputs land in gpurun_out/; summaries are copied to
# profiles/ by tools/summarize_profiles.py on the build box.
mkdir -p gpurun_out
R=${1:-r01}
# 1. launch list of one short bench run (cold-cache, serialised: compare SHARES only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_$R.csv python bench.py --steps 2 --warmup 3 > gpurun_out/launches_$R.log 2>&1
# 2. full captures of the three hot kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gmm_tc_kernel -s 3 -c 1 \
    -f -o gpurun_out/prof_fbf_tc_$R python bench.py --steps 2 --warmup 3 --skip-extras > gpurun_out/prof_fbf_tc_$R.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gmm_simt_kernel -s 3 -c 1 \
    -f -o gpurun_out/prof_fbf_simt_$R python bench.py --steps 2 --warmup 3 --skip-extras --variant 1 > gpurun_out/prof_fbf_simt_$R.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"traj_solve_kernel|dtw_fused_kernel" -s 2 -c 2 \
    -f -o gpurun_out/prof_traj_dtw_$R python bench.py --steps 1 --warmup 3 --frames 131072 > gpurun_out/prof_traj_dtw_$R.log 2>&1
ls -la gpurun_out
