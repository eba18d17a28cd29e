#!/bin/bash
# Runs the GPU parity suites in isolated processes (a hang in one kernel cannot take the others
# down), then short bench runs.  Usage (on the GPU box, from the repo root): tools/gpu_check.sh [stage...]
# Logs go to gpurun_out/.
mkdir -p gpurun_out
stages="${@:-dtw simt traj_simt aux tc traj_tc full bench}"
run() { # name timeout cmd...
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.log
  timeout -k 10 "$t" "$@" > "gpurun_out/$name.log" 2>&1
  local rc=$?
  echo "rc=$rc $(tail -n 3 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-300)" | tee -a gpurun_out/summary.log
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.log 2>&1
for s in $stages; do
  case $s in
    dtw)       run t_dtw 600 python -m pytest tests/test_gpu_dtw.py -m gpu -q --tb=short -k "not full" ;;
    simt)      run t_gmm_simt 900 python -m pytest tests/test_gpu_gmmmap.py -m gpu -q --tb=short -k "not tcgen05 and not full" ;;
    traj_simt) run t_traj_simt 900 python -m pytest tests/test_gpu_traj.py -m gpu -q --tb=short -k "not tcgen05 and not full" ;;
    gv)        run t_gv 600 python -m pytest tests/test_gpu_gv.py -m gpu -q --tb=short ;;
    aux)       run t_aux 300 python -m pytest tests/test_gpu_aux.py -m gpu -q --tb=short ;;
    tc)        run t_gmm_tc 600 python -m pytest tests/test_gpu_gmmmap.py -m gpu -q --tb=short -k "tcgen05" ;;
    traj_tc)   run t_traj_tc 600 python -m pytest tests/test_gpu_traj.py -m gpu -q --tb=short -k "tcgen05" ;;
    full)      run t_full 1200 python -m pytest tests -m gpu -q --tb=short -k "full" ;;
    smoke)     run t_smoke 600 python -c "import __graft_entry__ as g; g.smoke()" ;;
    bench)     run b_simt 600 python bench.py --steps 5 --warmup 3 --variant 1 --skip-extras
               run b_tc 600 python bench.py --steps 5 --warmup 3 --variant 2 --skip-extras
               run b_full 900 python bench.py ;;
    all)       run t_all 2400 python -m pytest tests -m gpu -q --tb=short ;;
  esac
done
cat gpurun_out/summary.log
