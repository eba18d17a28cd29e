#!/usr/bin/env python
"""Per-call wall time of the DTW host path (vcb_dtw_fit_batch, pinned inputs): shows warm-up effects."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vcb200 as vcb
vcb.set_device(0)
tm, to, sq, so = vcb.synth.config_c3(1000)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).pin_memory()     # (rows, T) column-major view, as bench.py does
ptm, psq = pin(tm), pin(sq)
htm, hsq = ptm.numpy().T, psq.numpy().T
d = vcb.DTWs.DTW(fstep=0, bstep=2)
ts = []
for _ in range(12):
    t0 = time.perf_counter(); vcb.DTWs.fit_batch(d, htm, to, hsq, so); ts.append((time.perf_counter() - t0) * 1e3)
print("stream=%s slice=%s: " % (os.environ.get("VCB_DTW_STREAM", "1"), os.environ.get("VCB_DTW_SLICE", "148")) + " ".join(f"{t:.2f}" for t in ts))
