#!/bin/bash
# Short end-of-session run (GPU box, repo root): full GPU parity suite, smoke, the three bench lines, one ncu capture of the band solver.
R=${1:-r02e}
mkdir -p gpurun_out
S=gpurun_out/final_$R.log
: > $S
echo "== pytest -m gpu" >> $S
timeout -k 10 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -n 4 >> $S
echo "== smoke" >> $S
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2 >> $S
for p in fbf traj dtw; do
  timeout -k 10 900 python bench.py --path $p > gpurun_out/bench_${p}_$R.json 2> gpurun_out/bench_${p}_$R.err
  echo "== bench --path $p: $(tail -c 200 gpurun_out/bench_${p}_$R.json)" >> $S
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:traj_solve_warp -s 1 -c 1 -f -o gpurun_out/prof_traj_$R python tools/run_path.py traj 1 > gpurun_out/prof_traj_$R.log 2>&1
cat $S
