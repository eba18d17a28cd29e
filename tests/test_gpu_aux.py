"""GPU parity of the callers either side of the path: push_delta and align."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_push_delta(vcb, oracle):
    rng = np.random.default_rng(2)
    lens = [1, 2, 3, 50, 7]
    off = np.concatenate([[0], np.cumsum(lens)])
    src = rng.standard_normal((5, off[-1]))
    out = vcb.push_delta(src, off)
    ref = np.concatenate([oracle.push_delta(np.asfortranarray(src[:, off[i]:off[i + 1]])) for i in range(5)], 1)
    assert np.array_equal(out, ref)
    assert np.array_equal(vcb.push_delta(src[:, :50]), oracle.push_delta(np.asfortranarray(src[:, :50])))


def test_align(vcb, oracle):
    tm, to, sq, so = vcb.synth.dtw_pairs(7, 12, (60, 120), 5, noise=0.1)
    newtgt, paths = vcb.align_batch(tm, to, sq, so)
    for p in range(7):
        s, nt, path = oracle.align(np.asfortranarray(tm[:, to[p]:to[p + 1]]), np.asfortranarray(sq[:, so[p]:so[p + 1]]))
        assert np.array_equal(paths[so[p]:so[p + 1]], path)
        assert np.array_equal(newtgt[:, to[p]:to[p + 1]], nt)
    s, nt = vcb.align(tm[:, :to[1]], sq[:, :so[1]])
    assert np.array_equal(nt, newtgt[:, :to[1]])
    with pytest.raises(vcb.DimensionMismatch):
        vcb.align(np.zeros((3, 4)), np.zeros((2, 4)))                 # src/align.jl:11-13


def test_vc_static_batch_fuses_push_delta(vcb, oracle):
    """bin/vc.jl:76-82: src = [src[1,:]; push_delta(src[2:end,:])]; vc(mapper, src) -- in one call."""
    import torch
    Ds = 24
    gm = vcb.synth.random_joint_gmm(41, 8, 4 * Ds)
    fm, off = vcb.synth.trajectory_utterances(gm, 4, (20, 70), 41)
    static = np.asfortranarray(fm[:1 + Ds])                         # power row + static features
    full = np.asfortranarray(np.concatenate(
        [np.concatenate([static[:1, off[i]:off[i + 1]], oracle.push_delta(static[1:, off[i]:off[i + 1]])], axis=0)
         for i in range(4)], axis=1))
    ref = oracle.vc_traj_batch(oracle.GMMMap(*gm), 30, full, off)
    t = vcb.TrajectoryGMMMap(vcb.GMMMap(*gm), 30)
    out = vcb.vc_static_batch(t, static, off)
    assert out.shape == ref.shape and np.array_equal(out[0], static[0])
    assert np.abs(out - ref).max() <= 1e-4 * np.abs(ref).max()
    two_step = np.concatenate(vcb.vc_batch(vcb.TrajectoryGMMMap(vcb.GMMMap(*gm), 30), full, off), axis=1)
    assert np.array_equal(out, two_step)                            # same kernels, same inputs
    d = torch.from_numpy(np.ascontiguousarray(static.T)).cuda()
    outd = vcb.vc_static_batch(vcb.TrajectoryGMMMap(vcb.GMMMap(*gm), 30), d, off).cpu().numpy().T
    assert np.array_equal(outd, out)
    with pytest.raises(vcb.DimensionMismatch):
        vcb.vc_static_batch(t, fm, off)                             # already has delta rows


def test_multi_device_mode_matches_single_device(vcb, oracle):
    """vcb_init(n): the host batch entry points shard over the visible devices (one host thread and
    pipeline per device, replicated model) and return exactly what one device returns.  With a single
    visible GPU the call degenerates to single-device mode."""
    ndev = vcb.device_count()
    try:
        assert vcb.init(0) == max(1, ndev)
        gm, fm = vcb.synth.config_c1(ndev * 65536 + 777)
        g = vcb.GMMMap(*gm)
        multi = vcb.vc(g, fm)
        gm2, fm2, off2 = vcb.synth.config_c2(4 * ndev + 3, 90)
        # batches below the sharding threshold stay on one device; use a long one as well
        big = np.asfortranarray(np.tile(fm2, (1, 48)))
        boff = np.concatenate([[0], np.cumsum(np.tile(np.diff(off2), 48))]).astype(np.int64)
        t = vcb.TrajectoryGMMMap(vcb.GMMMap(*gm2), 40)
        tmulti, = vcb.vc_batch(t, big, boff, _split=False)
        tm, to, sq, so = vcb.synth.dtw_pairs(8 * ndev + 1, 24, (60, 120), 5)
        pmulti, cmulti = vcb.DTWs.fit_batch(vcb.DTWs.DTW(fstep=0, bstep=2), tm, to, sq, so)
        amulti, _ = vcb.align_batch(tm, to, sq, so)
        assert vcb.init(1) == 1
        # the shards cut the frame axis elsewhere than one device's slices do, and the frames of a partial
        # 128-frame tile take a differently rounded path: equal within the parity bar, not bit for bit
        single = vcb.vc(g, fm)
        assert np.array_equal(multi[0], fm[0]) and np.abs(multi - single).max() <= 1e-5 * np.abs(single[1:]).max()
        tsingle, = vcb.vc_batch(vcb.TrajectoryGMMMap(vcb.GMMMap(*gm2), 40), big, boff, _split=False)
        assert np.array_equal(tmulti, tsingle)
        psingle, csingle = vcb.DTWs.fit_batch(vcb.DTWs.DTW(fstep=0, bstep=2), tm, to, sq, so)
        assert np.array_equal(pmulti, psingle) and np.array_equal(cmulti, csingle)
        assert np.array_equal(amulti, vcb.align_batch(tm, to, sq, so)[0])
        ref, _ = oracle.dtw_fit_batch(tm, to, sq, so, 0, 2)
        assert np.array_equal(pmulti, ref)
    finally:
        vcb.init(1)


def test_pinned_arrays(vcb):
    """pinned_empty returns page-locked column-major arrays (vcb_host_alloc); the host paths accept them
    as inputs and result buffers and free them with their last view."""
    import gc
    gm, fm = vcb.synth.config_c1(5000)
    g = vcb.GMMMap(*gm)
    a = vcb.pinned_empty(fm.shape)
    assert a.flags.f_contiguous and a.dtype == np.float64
    a[...] = fm
    out = vcb.pinned_empty(fm.shape)
    vcb.vc(g, a, out=out)
    assert np.array_equal(out, vcb.vc(g, fm))
    view = out[:, 10:20]
    del out, a
    gc.collect()
    assert np.isfinite(view).all()
