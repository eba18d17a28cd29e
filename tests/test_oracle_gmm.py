"""Oracle GMMMap / predict_proba / fvconvert / vc vs independent NumPy/SciPy formulations and the
reference's accessor test (test/gmmmap.jl) on the real model."""
import os

import numpy as np
import pytest
from scipy.stats import multivariate_normal

from conftest import GOLDEN


def _np_fvconvert(w, mu, sg, x):
    D = mu.shape[0] // 2
    M = len(w)
    lp = np.array([multivariate_normal(mu[:D, m], sg[:D, :D, m]).logpdf(x) + np.log(w[m]) for m in range(M)])
    p = np.exp(lp - np.logaddexp.reduce(lp))
    E = np.stack([mu[D:, m] + sg[D:, :D, m] @ np.linalg.solve(sg[:D, :D, m], x - mu[:D, m]) for m in range(M)], 1)
    return E @ p, p


def test_accessors_on_real_model(oracle, fixture_model):
    w, mu, sg = fixture_model
    g = oracle.GMMMap(w, mu, sg)                  # test/gmmmap.jl:8
    D = mu.shape[0] // 2
    assert len(g) == 1                            # :9
    assert g.dim == D == 40                       # :12
    assert g.ncomponents == len(w) == 32          # :13
    assert g.size == (D, 1)                       # :14


def test_fvconvert_matches_scipy_on_real_model(oracle, fixture_model):
    w, mu, sg = fixture_model
    z = np.load(os.path.join(GOLDEN, "fbf_c0.npz"))
    g = oracle.GMMMap(w, mu, sg)
    for t in range(0, 96, 8):
        x = z["fm"][1:, t]
        y, p = _np_fvconvert(w, mu, sg, x)
        assert np.abs(g.fvconvert(x) - y).max() < 1e-9
        po = g.predict_proba(x)
        assert np.abs(po - p).max() < 1e-9 and abs(po.sum() - 1) < 1e-12
        assert g.predict(x)[0] == int(np.argmax(po)) + 1


def test_golden_c0(oracle, fixture_model):
    z = np.load(os.path.join(GOLDEN, "fbf_c0.npz"))
    g = oracle.GMMMap(*fixture_model)
    out = g.vc(z["fm"])
    assert np.array_equal(out, z["out"])
    assert np.array_equal(out[0], z["fm"][0])                 # power row, src/common.jl:23
    assert np.array_equal(g.vc(z["fm"], nthreads=3), z["out"])


def test_single_mixture_is_linear_regression(oracle, vcb):
    gm = vcb.synth.random_joint_gmm(4, 1, 12)
    g = oracle.GMMMap(*gm)
    x = np.random.default_rng(0).standard_normal(6)
    A = gm.covars[6:, :6, 0] @ np.linalg.inv(gm.covars[:6, :6, 0])
    assert np.allclose(g.fvconvert(x), gm.means[6:, 0] + A @ (x - gm.means[:6, 0]), atol=1e-12)


def test_swap_equals_permuted_joint(oracle, vcb):
    gm = vcb.synth.random_joint_gmm(5, 3, 10)
    perm = np.r_[5:10, 0:5]
    g1 = oracle.GMMMap(gm.weights, gm.means, gm.covars, swap=True)       # src/gmmmap.jl:74-78
    g2 = oracle.GMMMap(gm.weights, gm.means[perm], gm.covars[np.ix_(perm, perm)])
    x = np.random.default_rng(1).standard_normal(5)
    assert np.allclose(g1.fvconvert(x), g2.fvconvert(x), atol=1e-12)


def test_errors(oracle, vcb):
    gm = vcb.synth.random_joint_gmm(6, 3, 8)
    g = oracle.GMMMap(*gm)
    with pytest.raises(oracle.OracleError) as e:
        g.fvconvert(np.zeros(5))                                         # src/gmmmap.jl:102
    assert e.value.code == oracle.EDIM
    bad = gm.covars.copy(); bad[:4, :4, 1] = -np.eye(4)
    with pytest.raises(oracle.OracleError) as e:
        oracle.GMMMap(gm.weights, gm.means, bad)                         # PosDefException
    assert e.value.code == oracle.ENOTPD
    w0 = gm.weights.copy(); w0[0] += w0[1]; w0[1] = 0.0
    g0 = oracle.GMMMap(w0, gm.means, gm.covars)
    with pytest.raises(oracle.OracleError) as e:                         # quirk Q1: zero weight -> throws
        g0.fvconvert(np.zeros(4))
    assert e.value.code == oracle.EDIM
    with pytest.raises(oracle.OracleError):
        oracle.GMMMap(gm.weights * 2, gm.means, gm.covars)               # not a probability vector


def test_vs_sklearn(oracle, fixture_model):
    """Third, independent pin of predict_proba / predict (src/gmm.jl:24-58): scikit-learn's
    GaussianMixture -- the library bin/train_gmm.jl:84-89 trains the reference's models with -- given
    the fixture model's source marginal.  Its posterior comes from its own precision-Cholesky code
    path (no shared code with the oracle or with scipy.stats)."""
    from sklearn.mixture import GaussianMixture
    from sklearn.mixture._gaussian_mixture import _compute_precision_cholesky
    w, mu, sg = fixture_model
    D, M = mu.shape[0] // 2, len(w)
    z = np.load(os.path.join(GOLDEN, "fbf_c0.npz"))
    X = np.asfortranarray(z["fm"][1:, :200])
    gm = GaussianMixture(n_components=M, covariance_type="full")
    gm.weights_ = w
    gm.means_ = np.ascontiguousarray(mu[:D].T)
    gm.covariances_ = np.ascontiguousarray(np.transpose(sg[:D, :D], (2, 0, 1)))
    gm.precisions_cholesky_ = _compute_precision_cholesky(gm.covariances_, "full")
    post = gm.predict_proba(X.T)                                  # (T, M)
    g = oracle.GMMMap(w, mu, sg)
    ours = np.stack([g.predict_proba(X[:, t]) for t in range(X.shape[1])])
    assert np.abs(ours - post).max() < 1e-9
    assert np.array_equal(g.predict(X), gm.predict(X.T) + 1)      # first maximum, 1-based
    # and the conversion itself from sklearn's posterior: E[y|x] = sum_m p_m (mu_y + Syx Sxx^-1 (x - mu_x))
    E = np.stack([mu[D:, m][:, None] + sg[D:, :D, m] @ np.linalg.solve(sg[:D, :D, m], X - mu[:D, m][:, None]) for m in range(M)])
    y = np.einsum("tm,mdt->dt", post, E)
    assert np.abs(g.vc(np.asfortranarray(z["fm"][:, :200]))[1:] - y).max() < 1e-9 * max(1.0, np.abs(y).max())
