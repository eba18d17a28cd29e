#!/usr/bin/env python
"""Per-call wall time of the trajectory host path (vcb_traj_vc_batch, pinned buffers), C2 or C4-shard shape.
env: N_UTT (1000), MIX (64), LIMIT (500), VCB_TRAJ_SLICE_CHUNKS"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vcb200 as vcb
vcb.set_device(0)
n = int(os.environ.get("N_UTT", 1000)); M = int(os.environ.get("MIX", 64)); limit = int(os.environ.get("LIMIT", 500))
gm, fm, off = vcb.synth.config_c2(n, 500, M=M) if M != 64 else vcb.synth.config_c2(n, 500)
tj = vcb.TrajectoryGMMMap(vcb.GMMMap(*gm), limit)
hin = torch.from_numpy(np.ascontiguousarray(fm.T)).pin_memory()
hout = torch.empty((fm.shape[1], 25), dtype=torch.float64).pin_memory()
a, b = hin.numpy().T, hout.numpy().T
ts = []
for _ in range(10):
    t0 = time.perf_counter(); vcb.vc_batch(tj, a, off, _split=False, out=b); ts.append((time.perf_counter() - t0) * 1e3)
print("n=%d M=%d limit=%d slice=%s: " % (n, M, limit, os.environ.get("VCB_TRAJ_SLICE_CHUNKS", "1036")) + " ".join(f"{t:.2f}" for t in ts))
