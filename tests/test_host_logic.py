"""Host-side logic that needs no GPU: sharding, the JLD model reader, synthetic workloads."""
import os

import numpy as np
import pytest


def test_partition_contiguous_balances(vcb):
    rng = np.random.default_rng(0)
    costs = rng.integers(100, 900, size=1000)
    parts = vcb.shard.partition_contiguous(costs, 8)
    assert parts[0][0] == 0 and parts[-1][1] == 1000
    assert all(parts[i][1] == parts[i + 1][0] for i in range(7))
    loads = [costs[b:e].sum() for b, e in parts]
    assert max(loads) / (costs.sum() / 8) < 1.02
    assert vcb.shard.partition_contiguous([5, 5], 4)[-1][1] == 2          # more parts than units
    assert vcb.shard.partition_contiguous([], 3) == [(0, 0)] * 3


def test_frame_range_and_ragged(vcb):
    spans = [vcb.shard.frame_range(1_000_003, r, 8) for r in range(8)]
    assert spans[0][0] == 0 and spans[-1][1] == 1_000_003
    assert all(spans[i][1] == spans[i + 1][0] for i in range(7))
    off = np.concatenate([[0], np.cumsum(np.arange(1, 21))])
    covered = []
    for r in range(4):
        (b, e), rel = vcb.shard.shard_ragged(off, r, 4)
        assert rel[0] == 0 and len(rel) == e - b + 1
        covered.extend(range(b, e))
    assert covered == list(range(20))


def test_synth_is_deterministic(vcb):
    g1, fm1 = vcb.synth.config_c1(256)
    g2, fm2 = vcb.synth.config_c1(256)
    assert np.array_equal(fm1, fm2) and np.array_equal(g1.covars, g2.covars)
    assert fm1.shape == (25, 256) and g1.means.shape == (48, 64) and abs(g1.weights.sum() - 1) < 1e-12
    for m in range(0, 64, 16):
        ev = np.linalg.eigvalsh(g1.covars[:, :, m])
        assert ev.min() > 5e-5 and ev.max() < 1.5
    gm, fm, off = vcb.synth.config_c2(3, 50)
    assert fm.shape == (49, 150) and off.tolist() == [0, 50, 100, 150]
    # delta rows follow push_delta incl. the boundary copy (src/datasets.jl:8-11)
    assert np.array_equal(fm[25:, 0], fm[1:25, 0])
    assert np.allclose(fm[25:, 1], -0.5 * fm[1:25, 0] + 0.5 * fm[1:25, 2])
    tm, to, sq, so = vcb.synth.dtw_pairs(3, 24, (550, 650), 1003)
    assert tm.shape == (24, to[-1]) and sq.shape == (24, so[-1]) and np.all(np.diff(to) >= 550)


REF_MODEL = "/root/reference/test/models/clb_to_slt_gmm32_order40_diff.jld"


@pytest.mark.skipif(not os.path.exists(REF_MODEL), reason="reference checkout not present on this box")
def test_jld_reader_on_reference_model(vcb, fixture_model):
    d = vcb.jld.load(REF_MODEL)
    assert d["diff"] is True and d["n_components"] == 32
    w, mu, sg = fixture_model
    assert np.array_equal(d["weights"], w) and np.array_equal(d["means"], mu) and np.array_equal(d["covars"], sg)
    d2 = vcb.jld.load(REF_MODEL.replace("clb_to_slt_gmm32_order40_diff", "clb_and_slt_gmm32_order40"))
    assert d2["diff"] is False and d2["covars"].shape == (80, 80, 32)


def test_jld_writer_round_trip(vcb, fixture_model, tmp_path):
    """jld.save writes the bin/train_gmm.jl:106-113 schema as the HDF5 subset of JLD v0.1; jld.load reads
    it back bit for bit (both models of the reference's test/models when the checkout is present)."""
    models = [("fixture", fixture_model, True)]
    if os.path.exists(REF_MODEL):
        for name in ("clb_to_slt_gmm32_order40_diff", "clb_and_slt_gmm32_order40"):
            d = vcb.jld.load(REF_MODEL.replace("clb_to_slt_gmm32_order40_diff", name))
            models.append((name, (d["weights"], d["means"], d["covars"]), d["diff"]))
    for name, (w, mu, sg), diff in models:
        p = tmp_path / (name + ".jld")
        vcb.jld.save(str(p), w, mu, sg, diff=diff)
        raw = p.read_bytes()
        assert raw.startswith(b"Julia data file (HDF5), version 0.1.0") and raw[512:520] == b"\x89HDF\r\n\x1a\n"
        e = vcb.jld.load(str(p))
        assert np.array_equal(e["weights"], w) and np.array_equal(e["means"], mu) and np.array_equal(e["covars"], sg)
        assert e["diff"] is bool(diff) and e["n_components"] == len(w)
        assert e["means"].flags.f_contiguous and e["covars"].shape == (mu.shape[0], mu.shape[0], len(w))
    # a random model with other sizes, and the schema check of the writer
    gm = vcb.synth.random_joint_gmm(5, 3, 10)
    p = tmp_path / "small.jld"
    vcb.jld.save(str(p), *gm, diff=False, n_components=3)
    e = vcb.jld.load(str(p))
    assert np.array_equal(e["covars"], gm.covars) and e["diff"] is False
    with pytest.raises(vcb.jld.JLDFormatError):
        vcb.jld.save(str(p), gm.weights[:2], gm.means, gm.covars)


def test_jld_reader_rejects_garbage(vcb, tmp_path):
    p = tmp_path / "x.jld"
    p.write_bytes(b"not hdf5" * 100)
    with pytest.raises(vcb.jld.JLDFormatError):
        vcb.jld.load(str(p))


def test_fixture_model_properties(fixture_model):
    w, mu, sg = fixture_model
    assert abs(w.sum() - 1) < 1e-12 and mu.shape == (80, 32) and sg.shape == (80, 80, 32)
    conds = [np.linalg.cond(sg[:40, :40, m]) for m in range(32)]
    assert max(conds) > 1e6           # the conditioning that rules out single-pass TF32


def test_diffgmm_is_host_side_and_matches_the_oracle(vcb, oracle):
    """vcb_diffgmm (src/diffgmm.jl:9-25) is a pure parameter transform: it runs without a GPU and is
    bit-identical to the oracle's restatement; also no-GPU entry points of the GV family fail loudly."""
    gm = vcb.synth.random_joint_gmm(17, 5, 12)
    w, mo, so = vcb.diffgmm((gm.weights, gm.means, gm.covars))
    om, os_ = oracle.diffgmm(gm.means, gm.covars)
    assert np.array_equal(mo, om) and np.array_equal(so, os_) and w is gm.weights
    d = vcb.diffgmm(gm)                                   # JointGMM named tuple in, same type out
    assert type(d) is type(gm) and np.array_equal(d.means, om)
    with pytest.raises(vcb.ArgumentError):
        vcb.diffgmm((gm.weights, gm.means[:5], gm.covars[:5, :5]))     # odd joint dimension


def test_gv_entry_points_need_a_gpu(vcb):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    with pytest.raises(vcb.CudaError):
        vcb.fvpostf(vcb.VarianceScaling(np.ones(3)), np.zeros((3, 8)))
