#!/usr/bin/env python
"""End-to-end throughput of the in-library multi-device mode (vcb_init): ONE process, host Float64
buffers through the C ABI, sharded over 1/2/4/8 GPUs by the library.  Prints one JSON line per device
count and path.  usage: multidev_e2e.py [fbf] [traj] [dtw]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vcb200 as vcb

paths = sys.argv[1:] or ["fbf", "traj", "dtw"]
ndev = vcb.device_count()
counts = [n for n in (int(c) for c in os.environ.get("COUNTS", "1,2,4,8").split(",")) if n <= ndev]


def pinned(a_T):
    t = torch.from_numpy(np.ascontiguousarray(a_T)).pin_memory()
    return t, t.numpy().T


def timed(fn, reps=5):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


if "fbf" in paths:
    per = 1_000_000
    gm, fm1 = vcb.synth.config_c1(per)
    for n in counts:
        fm = np.asfortranarray(np.tile(fm1, (1, n)))              # n million frames, sharded by the library
        _, hfm = pinned(fm.T)
        hout_t = torch.empty((fm.shape[1], fm.shape[0]), dtype=torch.float64).pin_memory()
        hout = hout_t.numpy().T
        vcb.init(n)
        g = vcb.GMMMap(*gm)
        s = timed(lambda: vcb.vc(g, hfm, out=hout))
        ok = bool(np.array_equal(hout[:, :per], hout[:, (n - 1) * per:]))
        print(json.dumps({"path": "fbf C1 x n", "gpus": n, "frames": fm.shape[1], "ms": s * 1e3, "frames_per_s": fm.shape[1] / s,
                          "gb_per_s_each_direction": fm.size * 8 / s / 1e9, "shards_agree": ok}), flush=True)
        del g
if "traj" in paths:
    gm = vcb.synth.random_joint_gmm(1004, 128, 96)
    for n in counts:
        fm, off = vcb.synth.c4_utterances(gm, np.arange(1024 * n), 500)
        _, hfm = pinned(fm.T)
        hout_t = torch.empty((fm.shape[1], 25), dtype=torch.float64).pin_memory()
        hout = hout_t.numpy().T
        vcb.init(n)
        t = vcb.TrajectoryGMMMap(vcb.GMMMap(*gm), 500)
        s = timed(lambda: vcb.vc_batch(t, hfm, off, _split=False, out=hout), reps=3)
        vcb.init(1)
        ref, = vcb.vc_batch(vcb.TrajectoryGMMMap(vcb.GMMMap(*gm), 500), np.asfortranarray(fm[:, :5000]), off[:11], _split=False)
        print(json.dumps({"path": "traj C4 (1024 utt per GPU)", "gpus": n, "frames": fm.shape[1], "ms": s * 1e3,
                          "frames_per_s": fm.shape[1] / s, "matches_single_device": bool(np.array_equal(ref, hout[:, :5000]))}), flush=True)
        del t
if "dtw" in paths:
    for n in counts:
        tm, to, sq, so = vcb.synth.config_c3(1000 * n)
        _, htm = pinned(tm.T); _, hsq = pinned(sq.T)
        vcb.init(n)
        d = vcb.DTWs.DTW(fstep=0, bstep=2)
        s = timed(lambda: vcb.DTWs.fit_batch(d, htm, to, hsq, so), reps=3)
        cells = float(np.sum(np.diff(to).astype(float) * np.diff(so)))
        print(json.dumps({"path": "dtw C3 (1000 pairs per GPU)", "gpus": n, "cells": cells, "ms": s * 1e3, "cells_per_s": cells / s}), flush=True)
vcb.init(1)
