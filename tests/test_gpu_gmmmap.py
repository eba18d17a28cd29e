"""GPU parity of GMMMap conversion against the oracle: max abs error <= 1e-4 x feature scale
(BASELINE.json).  Runs for both kernels (CUDA-core fp32 and tcgen05 3xTF32)."""
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


def test_accessors_and_params(vcb, oracle, fixture_model):
    w, mu, sg = fixture_model
    g = vcb.GMMMap(w, mu, sg)                       # test/gmmmap.jl:8-14
    assert len(g) == 1 and vcb.dim(g) == 40 and vcb.ncomponents(g) == 32 and g.size == (40, 1)
    o = oracle.GMMMap(w, mu, sg)
    A = o.param(2, (40, 40, 32))
    assert np.allclose(g.params.SyxSxxinv, A, rtol=1e-6, atol=1e-9 * np.abs(A).max())
    assert np.array_equal(g.params.mux, mu[:40]) and np.array_equal(g.params.Syy, sg[40:, 40:])
    gs = vcb.GMMMap(w, mu, sg, swap=True)           # src/gmmmap.jl:74-78
    assert np.array_equal(gs.params.mux, mu[40:]) and np.array_equal(gs.params.Sxy, sg[40:, :40])


def test_c0_real_model(vcb, fixture_model, variant):
    z = np.load(os.path.join(GOLDEN, "fbf_c0.npz"))
    g = vcb.GMMMap(*fixture_model)
    out = vcb.vc(g, z["fm"])
    assert out.shape == z["out"].shape and np.isfinite(out).all()     # test/vc.jl:26
    assert np.array_equal(out[0], z["fm"][0])                         # power row untouched
    assert np.abs(out[1:] - z["out"][1:]).max() <= tol_for(z["out"][1:])
    # single-vector fvconvert, the reference call shape
    y = vcb.fvconvert(g, z["fm"][1:, 3])
    assert y.shape == (40,) and np.abs(y - z["out"][1:, 3]).max() <= tol_for(z["out"][1:])


@pytest.mark.parametrize("T", [1, 127, 128, 129, 1000, 4099])
def test_c1_shapes_and_tails(vcb, oracle, variant, T):
    gm, fm = vcb.synth.config_c1(T)
    g, o = vcb.GMMMap(*gm), oracle.GMMMap(*gm)
    ref = o.vc(fm, nthreads=oracle.max_threads())
    out = vcb.vc(g, fm)
    assert np.array_equal(out[0], fm[0])
    assert np.abs(out - ref).max() <= tol_for(ref[1:])


def test_stress_conditioning(vcb, oracle, variant):
    gm, fm = vcb.synth.config_c1(3000, stress=True)                   # lambda in [1e-7, 2]
    ref = oracle.GMMMap(*gm).vc(fm, nthreads=oracle.max_threads())
    out = vcb.vc(vcb.GMMMap(*gm), fm)
    assert np.abs(out - ref).max() <= tol_for(ref[1:])


@pytest.mark.parametrize("M,D", [(1, 4), (3, 5), (7, 13), (33, 40), (128, 24), (16, 64)])
def test_odd_shapes(vcb, oracle, variant, M, D):
    gm = vcb.synth.random_joint_gmm(100 + M + D, M, 2 * D, mean_scale=0.3)
    fm = vcb.synth.fbf_feature_matrix(gm, 700, 5)
    ref = oracle.GMMMap(*gm).vc(fm, nthreads=oracle.max_threads())
    try:
        out = vcb.vc(vcb.GMMMap(*gm), fm)
    except vcb.VCBError as e:
        if variant == 2 and e.code == vcb._lib.EUNSUPPORTED:
            pytest.skip("shape outside the tensor-core kernel's shared-memory plan (auto mode uses the CUDA-core kernel)")
        raise
    assert np.abs(out - ref).max() <= tol_for(ref[1:])


def test_overlapping_mixtures_soft_posteriors(vcb, oracle, variant):
    gm = vcb.synth.random_joint_gmm(8, 16, 16, lam_lo=0.3, lam_hi=1.0, mean_scale=0.2)
    fm = vcb.synth.fbf_feature_matrix(gm, 2000, 6)
    o = oracle.GMMMap(*gm)
    post = np.stack([o.predict_proba(fm[1:, t]) for t in range(200)], 1)
    assert np.median(post.max(0)) < 0.9                               # genuinely soft
    ref = o.vc(fm, nthreads=oracle.max_threads())
    out = vcb.vc(vcb.GMMMap(*gm), fm)
    assert np.abs(out - ref).max() <= tol_for(ref[1:])


def test_strided_convert_and_device_path(vcb, oracle, variant):
    import torch
    gm, fm = vcb.synth.config_c1(5000)
    g = vcb.GMMMap(*gm)
    ref = oracle.GMMMap(*gm).vc(fm, nthreads=oracle.max_threads())
    Y = vcb.fvconvert(g, np.asfortranarray(fm[1:]))                   # (D, T) matrix form
    assert np.abs(Y - ref[1:]).max() <= tol_for(ref[1:])
    dfm = torch.from_numpy(np.ascontiguousarray(fm.T)).cuda()         # frame-major (T, rows)
    dout = vcb.vc(g, dfm)
    torch.cuda.synchronize()
    out = dout.cpu().numpy().T
    assert np.array_equal(out[0], fm[0]) and np.abs(out - ref).max() <= tol_for(ref[1:])
    dY = vcb.fvconvert(g, dfm[:, 1:].contiguous())
    torch.cuda.synchronize()
    assert np.abs(dY.cpu().numpy().T - ref[1:]).max() <= tol_for(ref[1:])


def test_predict_and_proba(vcb, oracle, fixture_model):
    z = np.load(os.path.join(GOLDEN, "fbf_c0.npz"))
    g = vcb.GMMMap(*fixture_model)
    post = vcb.predict_proba(g, np.asfortranarray(z["fm"][1:]))
    assert post.shape == (32, 96) and np.abs(post - z["post"]).max() < 1e-9
    assert np.abs(post.sum(0) - 1).max() < 1e-12
    for v in (1, 2):
        vcb.set_kernel_variant(v)
        mh = vcb.predict(g, np.asfortranarray(z["fm"][1:]))
        assert np.array_equal(mh, z["post"].argmax(0) + 1)
    vcb.set_kernel_variant(0)


def test_errors(vcb):
    gm = vcb.synth.random_joint_gmm(6, 3, 8)
    g = vcb.GMMMap(*gm)
    with pytest.raises(vcb.DimensionMismatch):
        vcb.fvconvert(g, np.zeros(5))                                 # src/gmmmap.jl:102
    with pytest.raises(vcb.DimensionMismatch):
        vcb.vc(g, np.zeros((7, 3)))
    bad = gm.covars.copy(); bad[:4, :4, 1] = -np.eye(4)
    with pytest.raises(vcb.PosDefException):
        vcb.GMMMap(gm.weights, gm.means, bad)
    w0 = gm.weights.copy(); w0[0] += w0[1]; w0[1] = 0.0
    with pytest.raises(vcb.ArgumentError):
        vcb.GMMMap(w0, gm.means, gm.covars)                           # quirk Q1
    sing = gm.covars.copy(); sing[:, :, 0] = 0.0
    with pytest.raises((vcb.SingularException, vcb.PosDefException)):
        vcb.GMMMap(gm.weights, gm.means, sing)
    assert vcb.vc(g, np.zeros((5, 0))).shape == (5, 0)                # empty input


def test_full_c1_properties(vcb, oracle):
    """BASELINE C1 at full size (1M frames): spot-check against the oracle + permutation
    equivariance (every frame is converted independently, src/common.jl:17-19)."""
    gm, fm = vcb.synth.config_c1(1_000_000)
    g = vcb.GMMMap(*gm)
    out = vcb.vc(g, fm)
    assert np.isfinite(out).all() and np.array_equal(out[0], fm[0])
    idx = np.random.default_rng(0).choice(1_000_000, 4000, replace=False)
    sub = np.asfortranarray(fm[:, idx])
    ref = oracle.GMMMap(*gm).vc(sub, nthreads=oracle.max_threads())
    assert np.abs(out[:, idx] - ref).max() <= tol_for(ref[1:])
    out_perm = vcb.vc(g, sub)
    assert np.array_equal(out_perm, out[:, idx])


def test_host_calls_are_reentrant_across_threads(vcb, oracle):
    """Handles are immutable and the host pipeline keeps one cached (streams, staging) context per
    concurrent caller: four threads converting different matrices with one model get their own results."""
    import threading
    gm = vcb.synth.random_joint_gmm(23, 16, 48)
    g = vcb.GMMMap(*gm)
    fms = [vcb.synth.fbf_feature_matrix(gm, 20000 + 3000 * i, 50 + i) for i in range(4)]
    outs, errs = [None] * 4, []

    def work(i):
        try:
            for _ in range(3):
                outs[i] = vcb.vc(g, fms[i])
        except Exception as e:       # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs
    og = oracle.GMMMap(*gm)
    for i in range(4):
        ref = og.vc(np.asfortranarray(fms[i][:, :1500]))
        assert np.array_equal(outs[i][0], fms[i][0])
        assert np.abs(outs[i][:, :1500] - ref).max() <= tol_for(ref)
        assert np.array_equal(outs[i], vcb.vc(g, fms[i]))      # and the same as a serial call
