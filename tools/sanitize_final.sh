R=r02b
CS=/usr/local/cuda/bin/compute-sanitizer
run() { tool=$1; what=$2; t=$3; shift 3; out=gpurun_out/san_${tool}_${what}_$R.txt; echo "# compute-sanitizer --tool $tool  python -m pytest $*" > $out; timeout $t $CS --tool $tool --print-limit 5 python -m pytest "$@" -m gpu -x -q --tb=line 2>&1 | grep -E "passed|failed|ERROR SUMMARY|SUMMARY" | head -5 >> $out; tail -2 $out; }
run memcheck traj 600 tests/test_gpu_traj.py -k "not full"
run memcheck gv 600 tests/test_gpu_gv.py
run memcheck dtw 600 tests/test_gpu_dtw.py -k "not full"
run memcheck aux 600 tests/test_gpu_aux.py
run synccheck traj 600 tests/test_gpu_traj.py -k "not full and (golden or c4 or ragged or static)"
run memcheck tc 600 tests/test_gpu_gmmmap.py -k "tcgen05 and not full"
