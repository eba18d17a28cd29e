#!/usr/bin/env python
"""Runs one path of the library a few times (for ncu / quick timing): traj | dtw | fbf | argmax."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vcb200 as vcb

which = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
vcb.set_device(0)
def timeit(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps
if which == "traj":
    n = int(os.environ.get("N_UTT", 1000)); limit = int(os.environ.get("LIMIT", 500))
    M = int(os.environ.get("MIX", 64))
    if "DS" in os.environ:       # other static dimensions (e.g. DS=40: the order-40 fixture shape, 64-thread CTA solver)
        gm = vcb.synth.random_joint_gmm(1002, M, 4 * int(os.environ["DS"]))
        fm, off = vcb.synth.trajectory_utterances(gm, n, 500, 1002)
    else:
        gm, fm, off = vcb.synth.config_c2(n, 500, M=M) if M != 64 else vcb.synth.config_c2(n, 500)
    t = vcb.TrajectoryGMMMap(vcb.GMMMap(*gm), limit)
    d = torch.from_numpy(np.ascontiguousarray(fm.T)).cuda()
    ms = timeit(lambda: vcb.vc_batch(t, d, off, _split=False))
    print(f"traj C2 n={n} limit={limit}: {ms:.3f} ms  {n*500/ms*1e3:.3e} frames/s")
elif which == "gv":
    n = int(os.environ.get("N_UTT", 1000)); limit = int(os.environ.get("LIMIT", 500)); ep = int(os.environ.get("EPOCHS", 100))
    gm, fm, off = vcb.synth.config_c2(n, 500)
    rng = np.random.default_rng(0)
    a = rng.standard_normal((24, 24))
    tg = vcb.TrajectoryGVGMMMap(vcb.TrajectoryGMMMap(vcb.GMMMap(*gm), limit), rng.uniform(0.2, 1.0, 24), a @ a.T + 24 * np.eye(24))
    d = torch.from_numpy(np.ascontiguousarray(fm.T)).cuda()
    ms = timeit(lambda: vcb.vc_batch(tg, d, off, _split=False, epochs=ep))
    print(f"traj+GV C2 n={n} limit={limit} epochs={ep}: {ms:.3f} ms  {n*500/ms*1e3:.3e} frames/s")
elif which == "dtw":
    if "S_RANGE" in os.environ:      # e.g. S_RANGE=640,640: every template / sequence length in that range
        lo, hi = (int(x) for x in os.environ["S_RANGE"].split(","))
        tm, to, sq, so = vcb.synth.dtw_pairs(int(os.environ.get("N_PAIRS", 1000)), 24, (lo, hi), 1003)
    else:
        tm, to, sq, so = vcb.synth.config_c3(int(os.environ.get("N_PAIRS", 1000)))
    a = torch.from_numpy(np.ascontiguousarray(tm.T)).cuda(); b = torch.from_numpy(np.ascontiguousarray(sq.T)).cuda()
    d = vcb.DTWs.DTW(fstep=0, bstep=2)
    ms = timeit(lambda: vcb.DTWs.fit_batch(d, a, to, b, so))
    cells = float(np.sum(np.diff(to).astype(float) * np.diff(so)))
    print(f"dtw C3: {ms:.3f} ms  {cells/ms*1e3:.3e} cells/s")
elif which in ("fbf", "fbf_simt"):
    vcb.set_kernel_variant(1 if which == "fbf_simt" else 0)
    if "DIM" in os.environ:     # other shapes: DIM = feature dimension, MIX = mixtures (e.g. the order-40, 32-mixture fixture shape)
        gm = vcb.synth.random_joint_gmm(1001, int(os.environ.get("MIX", 64)), 2 * int(os.environ["DIM"]))
        fm = vcb.synth.fbf_feature_matrix(gm, int(os.environ.get("FRAMES", 1000000)), 1001)
    else:
        gm, fm = vcb.synth.config_c1(int(os.environ.get("FRAMES", 1000000)))
    g = vcb.GMMMap(*gm)
    d = torch.from_numpy(np.ascontiguousarray(fm.T)).cuda()
    ms = timeit(lambda: vcb.vc(g, d))
    print(f"{which} C1: {ms:.3f} ms  {fm.shape[1]/ms*1e3:.3e} frames/s")
