#!/bin/bash
# compute-sanitizer evidence runs (GPU box, repo root); summaries go to gpurun_out/san_<tool>_<what>_<round>.txt
R=${1:-r02}
CS=/usr/local/cuda/bin/compute-sanitizer
run() { # tool what timeout pytest-args...
  local tool=$1 what=$2 t=$3; shift 3
  local out=gpurun_out/san_${tool}_${what}_$R.txt
  echo "# compute-sanitizer --tool $tool  python -m pytest $*" > $out
  timeout $t $CS --tool $tool --print-limit 5 python -m pytest "$@" -m gpu -x -q --tb=line 2>&1 | grep -E "passed|failed|error|ERROR SUMMARY|Error|Race|Hazard|========= [A-Z]" | grep -v "Warning" | head -40 >> $out
  echo "exit=$?" >> $out
  tail -3 $out
}
run memcheck  tc      900 tests/test_gpu_gmmmap.py -k "tcgen05 and not full"
run synccheck tc      900 tests/test_gpu_gmmmap.py -k "tcgen05 and not full"
run racecheck tc      900 tests/test_gpu_gmmmap.py -k "tcgen05 and not full and not stress"
run memcheck  traj    900 tests/test_gpu_traj.py -k "not full"
run racecheck traj    900 tests/test_gpu_traj.py -k "tcgen05 and (golden or fvconvert or ragged) or c4_shaped"
VCB_TRAJ_PAIR=1 run racecheck trajpair 900 tests/test_gpu_traj.py -k "golden or stiff or test_static_dimensions"
VCB_TRAJ_PAIR=1 run synccheck trajpair 900 tests/test_gpu_traj.py -k "golden or stiff or test_static_dimensions"
run memcheck  dtw     900 tests/test_gpu_dtw.py -k "not full"
VCB_DTW_PIPE=1 run racecheck dtwpipe 600 tests/test_gpu_dtw.py -k "c3_shaped or ragged"
run memcheck  full    1200 tests -k "full"
