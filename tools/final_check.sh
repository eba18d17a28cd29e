#!/bin/bash
# End-of-session evidence run on the GPU box (repo root): full GPU parity suite, the three bench lines,
# launch list + ncu captures of the DTW kernels, compute-sanitizer on the DTW tests.  Outputs: gpurun_out/.
R=${1:-r02d}
mkdir -p gpurun_out
S=gpurun_out/final_$R.log
: > $S
echo "== pytest -m gpu" >> $S
timeout -k 10 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -n 8 >> $S
for p in fbf traj dtw; do
  echo "== bench --path $p" >> $S
  timeout -k 10 900 python bench.py --path $p > gpurun_out/bench_${p}_$R.json 2> gpurun_out/bench_${p}_$R.err
  tail -c 600 gpurun_out/bench_${p}_$R.json >> $S; echo >> $S
done
echo "== ncu" >> $S
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/launches_$R.log 2>&1
cap() { local name=$1 rx=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c 1 -f -o gpurun_out/${name}_$R "$@" > gpurun_out/${name}_$R.log 2>&1; }
cap prof_dtw        dtw_stream_kernel 1 python tools/run_path.py dtw 1
cap prof_dtwbarrier dtw_fused_kernel  1 env VCB_DTW_STREAM=0 python tools/run_path.py dtw 1
ls -la gpurun_out | grep "_$R" >> $S
echo "== compute-sanitizer" >> $S
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck synccheck racecheck; do
  out=gpurun_out/san_${tool}_dtw_$R.txt
  echo "# compute-sanitizer --tool $tool python -m pytest tests/test_gpu_dtw.py -k 'not full and not long'" > $out
  timeout 900 $CS --tool $tool --print-limit 8 python -m pytest tests/test_gpu_dtw.py -m gpu -x -q --tb=line -k "not full and not long" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|at .*dtw" | head -24 >> $out
  echo "-- $tool" >> $S; tail -n 6 $out >> $S
done
cat $S
