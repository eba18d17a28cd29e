"""GPU parity of TrajectoryGMMMap conversion against the oracle."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, tol_for

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[1, 2], ids=["simt", "tcgen05"])
def variant(request, vcb):
    vcb.set_kernel_variant(request.param)
    yield request.param
    vcb.set_kernel_variant(0)


def test_golden_small(vcb, variant):
    z = np.load(os.path.join(GOLDEN, "traj_small.npz"))
    seed, M, jd = [int(v) for v in z["seed"]]
    g = vcb.GMMMap(*vcb.synth.random_joint_gmm(seed, M, jd))
    t = vcb.TrajectoryGMMMap(g, int(z["limit"]))
    assert len(t) == 40 and vcb.dim(t) == 12 and vcb.ncomponents(t) == 8 and t.size == (12, 40)
    outs = vcb.vc_batch(t, z["fm"], z["offsets"])
    out = np.concatenate(outs, axis=1)
    assert out.shape == z["out"].shape
    assert np.array_equal(out[0], z["fm"][0])
    assert np.abs(out - z["out"]).max() <= tol_for(z["out"][1:])


@pytest.mark.parametrize("T", [1, 2, 3, 4, 50])
def test_fvconvert_aux_outputs(vcb, oracle, variant, T):
    gm = vcb.synth.random_joint_gmm(31, 6, 32)                       # Ds = 8
    fm, off = vcb.synth.trajectory_utterances(gm, 1, max(T, 3), 32)
    X = np.asfortranarray(fm[1:, :T])
    g, o = vcb.GMMMap(*gm), oracle.GMMMap(*gm)
    t, ot = vcb.TrajectoryGMMMap(g, 100), oracle.TrajectoryGMMMap(o, 100)
    Y, mh, Ey = vcb.fvconvert(t, X, return_aux=True)
    Yr, mhr, Eyr = ot.fvconvert(X, True)
    assert np.array_equal(mh, mhr)                                    # arg-max sequence is exact
    assert np.abs(Ey - Eyr).max() <= 1e-9 * max(1.0, np.abs(Eyr).max())
    assert np.abs(Y - Yr).max() <= tol_for(Yr)
    assert len(t) == T == len(ot)                                     # quirk Q3
    assert np.abs(t.Dy - ot.Dy).max() <= 1e-6 * np.abs(ot.Dy).max()


@pytest.mark.parametrize("limit", [500, 100, 7])
def test_c2_shaped_batch(vcb, oracle, variant, limit):
    gm, fm, off = vcb.synth.config_c2(12, 500)
    o = oracle.GMMMap(*gm)
    ref = oracle.vc_traj_batch(o, limit, fm, off, nthreads=oracle.max_threads())
    t = vcb.TrajectoryGMMMap(vcb.GMMMap(*gm), limit)
    outs = vcb.vc_batch(t, fm, off)
    out = np.concatenate(outs, axis=1)
    assert out.shape == (25, 6000) and np.array_equal(out[0], fm[0])
    assert np.abs(out - ref).max() <= tol_for(ref[1:])


@pytest.mark.parametrize("limit", [500, 100])
def test_c4_shaped(vcb, oracle, variant, limit):
    """BASELINE config 5 (C4) shape: 128 mixtures, static dimension 24, utterances of 500 frames;
    chunk limits 500 (one solve per utterance) and 100 (the CLI default, bin/vc.jl:18).  The
    arg-max mixture sequence must be exact, y within the 1e-4 bar (src/common.jl:31-63,
    src/trajectory_gmmmap.jl:65-110)."""
    gm, fm, off = vcb.synth.config_c2(16, 500, M=128, seed=1004)
    g, o = vcb.GMMMap(*gm), oracle.GMMMap(*gm)
    ref = oracle.vc_traj_batch(o, limit, fm, off, nthreads=oracle.max_threads())
    t = vcb.TrajectoryGMMMap(g, limit)
    out = np.concatenate(vcb.vc_batch(t, fm, off), axis=1)
    assert out.shape == (25, 8000) and np.array_equal(out[0], fm[0])
    assert np.abs(out - ref).max() <= tol_for(ref[1:])
    # arg-max sequence of two whole utterances through fvconvert's side output
    X = np.asfortranarray(fm[1:, :1000])
    _, mh, _ = vcb.fvconvert(vcb.TrajectoryGMMMap(g, 1000), X, return_aux=True)
    assert np.array_equal(mh, o.predict(X))
    # the sharded indexed generator used by bench.py --path traj produces distinct utterances
    fa, oa = vcb.synth.c4_utterances(gm, np.array([3, 5000]), 500)
    fb, _ = vcb.synth.c4_utterances(gm, np.array([5000]), 500)
    assert np.array_equal(fa[:, 500:], fb) and not np.array_equal(fa[:, :500], fb)


def test_indefinite_precision_is_reported(vcb):
    """A joint covariance whose yy block is indefinite gives an indefinite Dy: the reference's
    sparse `\\` would still return something (LU); the band Cholesky reports PosDefException
    instead of returning NaN (host path) and raises from vcb_traj_status after device calls."""
    import torch
    gm = vcb.synth.random_joint_gmm(77, 3, 16)                 # dim(g) = 8 = [static 4; delta 4]
    cov = gm.covars.copy()
    cov[8:, 8:, 1] -= 3.0 * np.eye(8) * np.abs(cov[8:, 8:, 1]).max()             # Syy indefinite, Sxx stays PD
    g = vcb.GMMMap(gm.weights, gm.means, cov)
    t = vcb.TrajectoryGMMMap(g, 50)
    fm, off = vcb.synth.trajectory_utterances(gm, 2, 50, 78)
    mh = vcb.predict(g, np.asfortranarray(fm[1:]))
    assert (mh == 2).any()                                 # the broken mixture is actually used
    with pytest.raises(vcb.PosDefException):
        vcb.vc_batch(t, fm, off)
    dfm = torch.from_numpy(np.ascontiguousarray(fm.T)).cuda()
    vcb.vc_batch(t, dfm, off)                              # enqueues; no synchronisation, no error yet
    with pytest.raises(vcb.PosDefException):
        vcb.traj_status(t)
    vcb.traj_status(t)                                     # the flag is cleared by the query
    # a healthy model on the same handle type stays clean
    t2 = vcb.TrajectoryGMMMap(vcb.GMMMap(*gm), 50)
    vcb.vc_batch(t2, dfm, off)
    vcb.traj_status(t2)


def test_ragged_batch_and_state(vcb, oracle):
    gm = vcb.synth.random_joint_gmm(41, 5, 24)                        # Ds = 6
    fm, off = vcb.synth.trajectory_utterances(gm, 9, (1, 70), 42)
    o = oracle.GMMMap(*gm)
    ref = oracle.vc_traj_batch(o, 16, fm, off)
    t = vcb.TrajectoryGMMMap(vcb.GMMMap(*gm), 16)
    mats = [np.asfortranarray(fm[:, off[i]:off[i + 1]]) for i in range(9)]
    outs = vcb.vc_batch(t, mats)
    assert np.abs(np.concatenate(outs, 1) - ref).max() <= tol_for(ref[1:])
    lastT = int(off[-1] - off[-2])
    assert len(t) == (lastT % 16 or min(16, lastT))
    # single-utterance vc == the reference driver incl. the mutable chunk length
    t2 = vcb.TrajectoryGMMMap(vcb.GMMMap(*gm), 16)
    ot = oracle.TrajectoryGMMMap(o, 16)
    for i in (0, 3):
        a, b = vcb.vc(t2, mats[i]), ot.vc(mats[i])
        assert a.shape == b.shape and np.abs(a - b).max() <= tol_for(ref[1:])
        assert len(t2) == len(ot)


def test_device_path_and_errors(vcb, oracle):
    import torch
    gm, fm, off = vcb.synth.config_c2(6, 120)
    ref = oracle.vc_traj_batch(oracle.GMMMap(*gm), 50, fm, off, nthreads=4)
    t = vcb.TrajectoryGMMMap(vcb.GMMMap(*gm), 50)
    dfm = torch.from_numpy(np.ascontiguousarray(fm.T)).cuda()
    outs = vcb.vc_batch(t, dfm, off)
    torch.cuda.synchronize()
    out = torch.cat(outs, 0).cpu().numpy().T
    assert np.array_equal(out[0], fm[0]) and np.abs(out - ref).max() <= tol_for(ref[1:])
    with pytest.raises(vcb.DimensionMismatch):
        vcb.fvconvert(t, np.zeros((10, 4)))                           # src/trajectory_gmmmap.jl:67-68
    with pytest.raises(vcb.DimensionMismatch):
        vcb.TrajectoryGMMMap(vcb.GMMMap(*vcb.synth.random_joint_gmm(1, 2, 6)), 10)   # odd dim(g)


def test_stiff_precisions(vcb, oracle):
    # conditioning like the survey's probe: fp32 band solve fails here, the FP64 solve must pass
    gm = vcb.synth.random_joint_gmm(51, 8, 48, lam_lo=1e-6, lam_hi=1.0)
    fm, off = vcb.synth.trajectory_utterances(gm, 3, 100, 52)
    ref = oracle.vc_traj_batch(oracle.GMMMap(*gm), 100, fm, off, nthreads=3)
    out = np.concatenate(vcb.vc_batch(vcb.TrajectoryGMMMap(vcb.GMMMap(*gm), 100), fm, off), 1)
    assert np.abs(out - ref).max() <= tol_for(ref[1:])


def test_full_c2_properties(vcb, oracle):
    """BASELINE C2 at full size (1000 utterances x 500 frames): spot-check utterances against the
    oracle and verify batch-independence (an utterance converts the same alone or in the batch)."""
    gm, fm, off = vcb.synth.config_c2(1000, 500)
    g = vcb.GMMMap(*gm)
    t = vcb.TrajectoryGMMMap(g, 500)
    outs = vcb.vc_batch(t, fm, off)
    assert len(outs) == 1000 and all(o.shape == (25, 500) for o in outs)
    o = oracle.GMMMap(*gm)
    pick = [0, 333, 999]
    sub = np.asfortranarray(np.concatenate([fm[:, off[i]:off[i + 1]] for i in pick], 1))
    ref = oracle.vc_traj_batch(o, 500, sub, np.array([0, 500, 1000, 1500]), nthreads=3)
    got = np.concatenate([outs[i] for i in pick], 1)
    assert np.abs(got - ref).max() <= tol_for(ref[1:])
    alone = vcb.vc(vcb.TrajectoryGMMMap(g, 500), np.asfortranarray(fm[:, off[333]:off[334]]))
    assert np.array_equal(alone, outs[333])


@pytest.mark.parametrize("Ds", [1, 5, 13, 16, 23, 25, 32, 47])
def test_static_dimensions_cover_both_solvers(vcb, oracle, Ds):
    """Odd, padded and exact static dimensions: 1..24 run the warp-per-chunk solver (padded tiles,
    scalar loads when Ds is odd), 25..48 the 64-thread CTA solver; the arg-max kernel pads 2Ds to 8."""
    gm = vcb.synth.random_joint_gmm(60 + Ds, 3, 4 * Ds)
    fm, off = vcb.synth.trajectory_utterances(gm, 3, (9, 40), 60 + Ds)
    ref = oracle.vc_traj_batch(oracle.GMMMap(*gm), 16, fm, off)
    t = vcb.TrajectoryGMMMap(vcb.GMMMap(*gm), 16)
    out = np.concatenate(vcb.vc_batch(t, fm, off), axis=1)
    assert out.shape == ref.shape and np.array_equal(out[0], fm[0])
    assert np.abs(out - ref).max() <= tol_for(ref[1:])
    # the static-input entry point goes through the same kernels
    if Ds > 1:
        stat = np.asfortranarray(fm[:1 + Ds])
        full = np.asfortranarray(np.concatenate(
            [np.concatenate([stat[:1, off[i]:off[i + 1]], oracle.push_delta(stat[1:, off[i]:off[i + 1]])], axis=0)
             for i in range(3)], axis=1))
        two = np.concatenate(vcb.vc_batch(vcb.TrajectoryGMMMap(vcb.GMMMap(*gm), 16), full, off), axis=1)
        assert np.array_equal(vcb.vc_static_batch(vcb.TrajectoryGMMMap(vcb.GMMMap(*gm), 16), stat, off), two)
