#!/usr/bin/env python
"""Trajectory-solver experiments: times the stages of vc_batch for a few batch shapes in one process
(library switches are read from the environment once per process).  usage: traj_exp.py [tag]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vcb200 as vcb

tag = sys.argv[1] if len(sys.argv) > 1 else ""
vcb.set_device(0)
cases = [c.split(":") for c in os.environ.get("CASES", "64:1000:500,64:1000:100,64:4096:500,128:1024:500").split(",")]
models = {}
for M, n, limit in cases:
    M, n, limit = int(M), int(n), int(limit)
    if M not in models:
        gm = vcb.synth.random_joint_gmm(1002 if M == 64 else 1004, M, 96)
        models[M] = (gm, vcb.GMMMap(*gm))
    gm, g = models[M]
    fm, off = vcb.synth.c4_utterances(gm, np.arange(n), 500, seed=1002)
    t = vcb.TrajectoryGMMMap(g, limit)
    d = torch.from_numpy(np.ascontiguousarray(fm.T)).cuda()
    for _ in range(2):
        vcb.vc_batch(t, d, off, _split=False)
    torch.cuda.synchronize()
    vcb.stage_timing(True)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        vcb.vc_batch(t, d, off, _split=False)
    e.record(); torch.cuda.synchronize()
    st = vcb.stage_times(); vcb.stage_timing(False)
    ms = s.elapsed_time(e) / 5
    print(f"{tag} M={M} n={n} limit={limit}: total {ms:.3f} ms  stages " + " ".join(f"{x:.3f}" for x in st) +
          f"  frac28224={28224*n*500/(ms*1e-3)/6452.2e9:.3f}", flush=True)
    del d
