#!/bin/bash
# Multi-GPU evidence run (one box, N GPUs visible).  usage: tools/multigpu_evidence.sh r02 8
R=${1:-r02}; N=${2:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${N}gpu_$R.txt 2>&1
tools/micro/copy_probe 200 10 0 > gpurun_out/copy_probe_$R.json 2>&1
tools/micro/copy_probe 200 10 1 >> gpurun_out/copy_probe_$R.json 2>&1
timeout 300 python -m pytest tests/test_gpu_aux.py -m gpu -x -q --tb=short -k multi_device > gpurun_out/multidev_test_$R.log 2>&1; tail -2 gpurun_out/multidev_test_$R.log
COUNTS=1,2,4,8 timeout 300 python tools/multidev_e2e.py fbf 2>&1 | grep -v Warn > gpurun_out/multidev_e2e_$R.json
COUNTS=1,$N timeout 400 python tools/multidev_e2e.py traj dtw 2>&1 | grep -v Warn >> gpurun_out/multidev_e2e_$R.json
for n in $N 4; do
  [ $n -le $N ] || continue
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n bench.py --gpus $n --path traj --steps 5 > gpurun_out/bench_traj_n${n}_$R.json 2> gpurun_out/bench_traj_n${n}_$R.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --path dtw --steps 10 > gpurun_out/bench_dtw_n${N}_$R.json 2> gpurun_out/bench_dtw_n${N}_$R.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29578 bench.py --gpus $N --steps 10 --skip-extras > gpurun_out/bench_fbf_n${N}_$R.json 2> gpurun_out/bench_fbf_n${N}_$R.err
cat gpurun_out/copy_probe_$R.json gpurun_out/multidev_e2e_$R.json
for f in gpurun_out/bench_*_n*_$R.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["n_gpus"], "value %.4g" % d["value"], "ms %.3f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], d.get("gather"), d.get("host"))
except Exception as ex:
    print(sys.argv[1], "unreadable", ex)
PY
done
