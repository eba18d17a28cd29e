#!/usr/bin/env python
"""Repeated trajectory / frame-by-frame / DTW calls with varying sizes (hang and determinism check)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vcb200 as vcb
vcb.set_device(0)
rng = np.random.default_rng(0)
gm, fm, off = vcb.synth.config_c2(64, 500)
g = vcb.GMMMap(*gm)
d = torch.from_numpy(np.ascontiguousarray(fm.T)).cuda()
gm1, fm1 = vcb.synth.config_c1(200000)
g1 = vcb.GMMMap(*gm1)
d1 = torch.from_numpy(np.ascontiguousarray(fm1.T)).cuda()
tmd, tod, sqd, sod = vcb.synth.dtw_pairs(400, 24, (20, 672), 5)          # persistent DTW kernel: several pairs per CTA
dtm = torch.from_numpy(np.ascontiguousarray(tmd.T)).cuda(); dsq = torch.from_numpy(np.ascontiguousarray(sqd.T)).cuda()
dtw = vcb.DTWs.DTW(fstep=0, bstep=2)
dref = None
ref = None
t0 = time.time()
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 150):
    n = int(rng.integers(1, 65)); lim = int(rng.choice([500, 100, 37, 7]))
    t = vcb.TrajectoryGMMMap(g, lim)
    out, = vcb.vc_batch(t, d[: off[n]], off[: n + 1], _split=False)
    T1 = int(rng.integers(1, 200001))
    o1 = vcb.vc(g1, d1[:T1])
    npair = int(rng.integers(1, 401))
    pp, pc = vcb.DTWs.fit_batch(dtw, dtm[: tod[npair]], tod[: npair + 1], dsq[: sod[npair]], sod[: npair + 1])
    if it % 10 == 0:
        pa, ca = vcb.DTWs.fit_batch(dtw, dtm, tod, dsq, sod)
        if dref is None: dref = (pa.clone(), ca.clone())
        assert torch.equal(pa, dref[0]) and torch.equal(ca, dref[1]), "DTW result changed between identical calls"
        assert torch.equal(pp, dref[0][: sod[npair]]), "DTW paths depend on the batch a pair is in"
        t2 = vcb.TrajectoryGMMMap(g, 500)
        full, = vcb.vc_batch(t2, d, off, _split=False)
        if ref is None: ref = full.clone()
        assert torch.equal(full, ref), "trajectory result changed between identical calls"
        assert bool(torch.isfinite(o1).all())
torch.cuda.synchronize()
print("stress OK: %.1f s" % (time.time() - t0))
