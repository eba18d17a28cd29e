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
