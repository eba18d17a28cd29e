"""Oracle trajectory conversion: W structure pinned by the reference (test/trajectory_gmmmap.jl),
normal equations vs a dense NumPy solve, chunked vc semantics."""
import os

import numpy as np
import pytest

from conftest import GOLDEN


def _dense_W(oracle, D, T):
    r, c, v = oracle.constructW(D, T)
    W = np.zeros((2 * D * T, D * T))
    W[r, c] = v
    return W


@pytest.mark.parametrize("D,T", [(30, 40), (3, 1), (2, 2)])
def test_constructW_structure(oracle, D, T):
    """test/trajectory_gmmmap.jl:1-34, statement by statement."""
    W = _dense_W(oracle, D, T)
    assert W.shape == (2 * D * T, D * T)
    I, Z = np.eye(D), np.zeros((D, D))
    for t in range(1, T + 1):
        s = 2 * D * (t - 1)
        rows = slice(s, s + D)
        assert np.array_equal(W[rows, (t - 1) * D:t * D], I)
        for i in range(1, T + 1):
            if i != t:
                assert np.array_equal(W[rows, (i - 1) * D:i * D], Z)
        rows = slice(s + D, s + 2 * D)
        if t >= 2:
            assert np.array_equal(W[rows, (t - 2) * D:(t - 1) * D], -0.5 * I)
        if t < T:
            assert np.array_equal(W[rows, t * D:(t + 1) * D], 0.5 * I)
        for i in range(1, T + 1):
            if i != t - 1 and i != t + 1:
                assert np.array_equal(W[rows, (i - 1) * D:i * D], Z)


def test_python_constructW_matches(oracle, vcb):
    assert np.array_equal(vcb.constructW(4, 7).toarray(), _dense_W(oracle, 4, 7))


def _setup(oracle, vcb, seed=3, M=4, Ds=4, T=12):
    gm = vcb.synth.random_joint_gmm(seed, M, 4 * Ds)
    fm, off = vcb.synth.trajectory_utterances(gm, 2, T, seed)
    return gm, fm, off, oracle.GMMMap(*gm)


@pytest.mark.parametrize("T", [1, 2, 3, 12])
def test_fvconvert_matches_dense_solve(oracle, vcb, T):
    gm, fm, off, g = _setup(oracle, vcb, T=max(T, 3))
    Ds = 4
    tg = oracle.TrajectoryGMMMap(g, T)
    assert len(tg) == T and tg.dim == 2 * Ds and tg.size == (2 * Ds, T)     # test/trajectory_gmmmap.jl:44-49
    X = np.asfortranarray(fm[1:, :T])
    Y, mh, Ey = tg.fvconvert(X, True)
    assert np.array_equal(mh, g.predict(X))                                  # :82
    W = _dense_W(oracle, Ds, T)
    Dinv = np.zeros((2 * Ds * T, 2 * Ds * T))
    for t in range(T):
        Dinv[2 * Ds * t:2 * Ds * (t + 1), 2 * Ds * t:2 * Ds * (t + 1)] = tg.Dy[:, :, mh[t] - 1]
    R = W.T @ Dinv @ W
    y = np.linalg.solve(R, W.T @ Dinv @ Ey.reshape(-1, order="F")).reshape(Ds, T, order="F")
    assert np.abs(y - Y).max() < 1e-10 * max(1.0, np.abs(y).max())
    # Dy as the reference computes it (:24-28)
    m = 1
    A = gm.covars[8:, :8, m] @ np.linalg.inv(gm.covars[:8, :8, m])
    assert np.allclose(tg.Dy[:, :, m], np.linalg.inv(gm.covars[8:, 8:, m] - A @ gm.covars[:8, 8:, m]), rtol=1e-8)


def test_vc_chunking_and_state(oracle, vcb):
    gm, fm, off, g = _setup(oracle, vcb)
    fm1 = np.asfortranarray(fm[:, :12])
    tg = oracle.TrajectoryGMMMap(g, 5)
    out = tg.vc(fm1)                                                         # src/common.jl:31-63
    assert out.shape == (1 + 4, 12)                                          # quirk Q5
    assert np.array_equal(out[0], fm1[0])                                    # :60
    ref = np.concatenate([oracle.TrajectoryGMMMap(g, 5).fvconvert(np.asfortranarray(fm1[1:, b:min(b + 5, 12)]))
                          for b in range(0, 12, 5)], axis=1)
    assert np.array_equal(out[1:], ref)
    assert len(tg) == 2                                                      # quirk Q3: last chunk was 2 frames
    with pytest.raises(oracle.OracleError) as e:
        tg.fvconvert(np.zeros((6, 4)))                                       # :67-68
    assert e.value.code == oracle.EDIM


def test_golden_traj(oracle, vcb):
    z = np.load(os.path.join(GOLDEN, "traj_small.npz"))
    seed, M, jd = [int(v) for v in z["seed"]]
    g = oracle.GMMMap(*vcb.synth.random_joint_gmm(seed, M, jd))
    out = oracle.vc_traj_batch(g, int(z["limit"]), z["fm"], z["offsets"], nthreads=2)
    assert np.array_equal(out, z["out"])


def test_vs_sparse_spsolve_at_c2_size(oracle, vcb):
    """Literal SciPy-sparse transcription of src/trajectory_gmmmap.jl:82-109 at the C2/C4 chunk size
    (static dimension 24, T = 500): W from the independent Python constructW, D^-1 = block_diag of the
    LU-inverse Dy[:,:,mhat_t], and a general sparse direct solve of (W' D^-1 W) y = W' D^-1 E (SuperLU,
    like the LU branch Julia's `\\` takes for this not exactly Hermitian matrix)."""
    import scipy.sparse as sp
    from scipy.sparse.linalg import spsolve
    Ds, T, M = 24, 500, 8
    gm = vcb.synth.random_joint_gmm(1002, M, 4 * Ds)
    fm, off = vcb.synth.trajectory_utterances(gm, 1, T, 1002)
    X = np.asfortranarray(fm[1:, :T])
    g = oracle.GMMMap(*gm)
    tg = oracle.TrajectoryGMMMap(g, T)
    Y = tg.fvconvert(X)
    # --- transcription
    D2 = 2 * Ds
    mu, sg = gm.means, gm.covars
    mhat = g.predict(X) - 1                                                        # :82
    A = [sg[D2:, :D2, m] @ np.linalg.inv(sg[:D2, :D2, m]) for m in range(M)]       # src/gmmmap.jl:34-36
    Dy = [np.linalg.inv(sg[D2:, D2:, m] - A[m] @ sg[:D2, D2:, m]) for m in range(M)]   # :24-28
    E = np.concatenate([mu[D2:, m] + A[m] @ (X[:, t] - mu[:D2, m]) for t, m in enumerate(mhat)])   # :85-91
    Dinv = sp.block_diag([sp.csc_matrix(Dy[m]) for m in mhat], format="csc")      # :95-96
    W = vcb.constructW(Ds, T)                                                      # :39-61
    WtD = W.T @ Dinv                                                               # :103
    y = spsolve((WtD @ W).tocsc(), WtD @ E)                                        # :105
    ref = y.reshape(Ds, T, order="F")                                              # :109
    assert np.abs(Y - ref).max() <= 1e-9 * np.abs(ref).max()
