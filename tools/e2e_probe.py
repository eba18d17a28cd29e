#!/usr/bin/env python
"""End-to-end timing of vc(g, fm) through the host C ABI with pinned buffers (C1 workload)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vcb200 as vcb
vcb.set_device(0)
gm, fm = vcb.synth.config_c1(1_000_000)
g = vcb.GMMMap(*gm)
hin = torch.from_numpy(np.ascontiguousarray(fm.T)).pin_memory(); hout = torch.empty_like(hin).pin_memory()
a, b = hin.numpy().T, hout.numpy().T
for _ in range(3): vcb.vc(g, a, out=b)
t0 = time.perf_counter()
for _ in range(10): vcb.vc(g, a, out=b)
dt = (time.perf_counter() - t0) / 10
print(f"e2e pinned: {dt*1e3:.3f} ms  {1e6/dt:.3e} frames/s  ({0.4/dt:.1f} GB/s both directions)")
fp = np.asfortranarray(fm.copy()); op = np.empty_like(fp, order="F")
for _ in range(2): vcb.vc(g, fp, out=op)
t0 = time.perf_counter()
for _ in range(3): vcb.vc(g, fp, out=op)
dt = (time.perf_counter() - t0) / 3
print(f"e2e pageable: {dt*1e3:.3f} ms  {1e6/dt:.3e} frames/s")
