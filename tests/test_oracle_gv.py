"""Oracle restatements of the GV / differential-model helpers:
VarianceScaling (src/gv.jl:6-21), TrajectoryGVGMMMap (src/trajectory_gmmmap.jl:112-189) and
diffgmm (src/diffgmm.jl:9-25).  The reference has no tests for any of them ("TODO: tests"), so the
oracle is checked against independent NumPy formulations and defining properties -- parity unpinned."""
import numpy as np
import pytest


def _setup(oracle, vcb, seed=5, M=4, Ds=3, T=14):
    gm = vcb.synth.random_joint_gmm(seed, M, 4 * Ds)
    fm, off = vcb.synth.trajectory_utterances(gm, 1, T, seed)
    return gm, np.asfortranarray(fm), oracle.GMMMap(*gm)


def test_variance_scaling_matches_numpy(oracle):
    rng = np.random.default_rng(0)
    src = np.asfortranarray(rng.standard_normal((5, 40)) * np.arange(1, 6)[:, None] + 3.0)
    s2 = rng.uniform(0.5, 2.0, 5)
    out = oracle.fvpostf(s2, src)
    mu = src.mean(axis=1, keepdims=True)
    ref = np.sqrt(s2[:, None] / src.var(axis=1, ddof=1, keepdims=True)) * (src - mu) + mu   # src/gv.jl:13
    assert np.abs(out - ref).max() < 1e-12
    assert np.allclose(out.var(axis=1, ddof=1), s2, rtol=1e-12)      # the filter's defining property
    assert np.allclose(out.mean(axis=1), src.mean(axis=1), rtol=1e-12)


def test_diffgmm_matches_definition_and_property(oracle, vcb):
    gm = vcb.synth.random_joint_gmm(11, 3, 8)
    D = 4
    mo, so = oracle.diffgmm(gm.means, gm.covars)
    for m in range(3):
        S = gm.covars[:, :, m]
        Sxx, Sxy, Syx, Syy = S[:D, :D], S[:D, D:], S[D:, :D], S[D:, D:]
        assert np.array_equal(mo[:D, m], gm.means[:D, m])
        assert np.array_equal(mo[D:, m], gm.means[D:, m] - gm.means[:D, m])          # eq. (6)
        assert np.array_equal(so[:D, :D, m], Sxx)
        assert np.array_equal(so[:D, D:, m], Sxy - Sxx)                                # eq. (7)
        assert np.array_equal(so[D:, :D, m], (Sxy - Sxx).T)
        assert np.array_equal(so[D:, D:, m], Sxx + Syy - Sxy - Syx)                    # eq. (8)
    # converting with the differential model gives E[y - x | x] = E[y | x] - x
    g, gd = oracle.GMMMap(*gm), oracle.GMMMap(gm.weights, mo, so)
    x = np.random.default_rng(1).standard_normal(D)
    assert np.abs(gd.fvconvert(x) - (g.fvconvert(x) - x)).max() < 1e-9


def _numpy_gv(oracle, tg, X, mu_v, S_vv, epochs, alpha):
    """The reference's update written with dense NumPy matrices (src/trajectory_gmmmap.jl:146-171)."""
    Ds, T = X.shape[0] // 2, X.shape[1]
    y0, mh, Ey = tg.fvconvert(X, True)
    r, c, v = oracle.constructW(Ds, T)
    W = np.zeros((2 * Ds * T, Ds * T)); W[r, c] = v
    Dinv = np.zeros((2 * Ds * T, 2 * Ds * T))
    for t in range(T):
        Dinv[2 * Ds * t:2 * Ds * (t + 1), 2 * Ds * t:2 * Ds * (t + 1)] = tg.Dy[:, :, mh[t] - 1]
    mu = y0.mean(axis=1, keepdims=True)
    y = np.sqrt(mu_v[:, None] / y0.var(axis=1, ddof=1, keepdims=True)) * (y0 - mu) + mu
    pv = np.linalg.inv(S_vv)
    WtD = W.T @ Dinv
    om = 1.0 / (2 * T)
    for _ in range(epochs):
        gv = y.var(axis=1, ddof=1)
        grad = -2.0 / T * (pv.T @ (gv - mu_v))[:, None] * (y - y.mean(axis=1, keepdims=True))
        d = om * (-WtD @ W @ y.reshape(-1, order="F") + WtD @ Ey.reshape(-1, order="F")) + grad.reshape(-1, order="F")
        y = y + alpha * d.reshape(Ds, T, order="F")
    return y


@pytest.mark.parametrize("epochs,alpha", [(0, 1e-5), (5, 1e-5), (20, 1e-3)])
def test_trajgv_matches_numpy(oracle, vcb, epochs, alpha):
    gm, fm, g = _setup(oracle, vcb)
    Ds, T = 3, fm.shape[1]
    tg = oracle.TrajectoryGMMMap(g, T)
    rng = np.random.default_rng(2)
    mu_v = rng.uniform(0.2, 1.0, Ds)
    a = rng.standard_normal((Ds, Ds))
    S_vv = a @ a.T + Ds * np.eye(Ds)
    tgv = oracle.TrajectoryGVGMMMap(tg, mu_v, S_vv)
    X = np.asfortranarray(fm[1:])
    y = tgv.fvconvert(X, epochs=epochs, alpha=alpha)
    ref = _numpy_gv(oracle, oracle.TrajectoryGMMMap(g, T), X, mu_v, S_vv, epochs, alpha)
    assert np.abs(y - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())
    if epochs == 0:     # eq. (58): the initial value carries exactly the target global variance
        assert np.allclose(y.var(axis=1, ddof=1), mu_v, rtol=1e-10)


def test_trajgv_vc_chunks_and_errors(oracle, vcb):
    gm, fm, g = _setup(oracle, vcb)
    Ds = 3
    tg = oracle.TrajectoryGMMMap(g, 6)
    mu_v, S_vv = np.full(Ds, 0.5), np.eye(Ds)
    tgv = oracle.TrajectoryGVGMMMap(tg, mu_v, S_vv)
    out = tgv.vc(fm, epochs=3)
    assert out.shape == (1 + Ds, fm.shape[1]) and np.array_equal(out[0], fm[0])
    ref = np.concatenate([oracle.TrajectoryGVGMMMap(oracle.TrajectoryGMMMap(g, 6), mu_v, S_vv)
                          .fvconvert(np.asfortranarray(fm[1:, b:min(b + 6, fm.shape[1])]), epochs=3)
                          for b in range(0, fm.shape[1], 6)], axis=1)
    assert np.array_equal(out[1:], ref)
    with pytest.raises(oracle.OracleError):
        oracle.TrajectoryGVGMMMap(tg, np.array([0.5, -0.1, 0.5]), S_vv)          # :124 @assert
    with pytest.raises(oracle.OracleError):
        oracle.TrajectoryGVGMMMap(tg, mu_v, np.zeros((Ds, Ds)))                  # inv of a singular matrix
