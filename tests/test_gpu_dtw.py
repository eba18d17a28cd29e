"""GPU parity: DTW paths must be bit-exact with the oracle (reference src/dtw.jl)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=[0, 1], ids=["stream", "barrier"])
def dtw_kernel(request, vcb):
    """Every test runs twice: with the default kernel choice (the persistent warp-pipeline kernel where it
    applies: fstep 0, bstep 1 / 2, D = 24 / 40, templates of up to 672 / 416 frames) and with the per-column
    barrier kernel everywhere (variant 1).  Both must be bit-exact with the oracle."""
    vcb.set_kernel_variant(request.param)
    yield request.param
    vcb.set_kernel_variant(0)


def test_reference_known_answers(vcb):
    for c in json.load(open(os.path.join(GOLDEN, "dtw_reference_tests.json"))):
        d = vcb.DTWs.DTW(bstep=c["bstep"], fstep=c["fstep"])
        p = vcb.DTWs.fit(d, np.array(c["template"], float).T, np.array(c["sequence"], float).T)
        assert p.tolist() == c["expected"]


def test_golden_random_pairs(vcb):
    z = np.load(os.path.join(GOLDEN, "dtw_random.npz"))
    for fs, bs in [(0, 1), (0, 2), (1, 2), (0, 5), (3, 20)]:       # 2-, 4- and 8-bit back-pointer builds
        paths, fc = vcb.DTWs.fit_batch(vcb.DTWs.DTW(fstep=fs, bstep=bs), z["tmpl"], z["tmpl_off"], z["seq"], z["seq_off"])
        assert np.array_equal(paths, z[f"paths_f{fs}_b{bs}"]), (fs, bs)
        assert np.array_equal(fc, z[f"cost_f{fs}_b{bs}"]), (fs, bs)


@pytest.mark.parametrize("fstep,bstep", [(0, 2), (0, 1)])
def test_c3_shaped_pairs_bit_exact(vcb, oracle, fstep, bstep):
    tm, to, sq, so = vcb.synth.dtw_pairs(24, 24, (550, 650), 1003)
    paths, fc = vcb.DTWs.fit_batch(vcb.DTWs.DTW(fstep=fstep, bstep=bstep), tm, to, sq, so)
    ref, rfc = oracle.dtw_fit_batch(tm, to, sq, so, fstep, bstep, nthreads=oracle.max_threads())
    assert np.array_equal(paths, ref)
    assert np.array_equal(fc, rfc)


def test_ragged_and_tiny(vcb, oracle):
    rng = np.random.default_rng(4)
    S = [1, 2, 31, 32, 33, 97, 1000, 1024]
    T = [1, 5, 16, 17, 15, 200, 37, 3]
    to = np.concatenate([[0], np.cumsum(S)]); so = np.concatenate([[0], np.cumsum(T)])
    tm = rng.standard_normal((3, to[-1])); sq = rng.standard_normal((3, so[-1]))
    for fs, bs in [(0, 1), (0, 2), (2, 2)]:
        paths, fc = vcb.DTWs.fit_batch(vcb.DTWs.DTW(fstep=fs, bstep=bs), tm, to, sq, so)
        ref, rfc = oracle.dtw_fit_batch(tm, to, sq, so, fs, bs)
        assert np.array_equal(paths, ref) and np.array_equal(fc, rfc)


def test_ragged_and_tiny_compile_time_dimension(vcb, oracle):
    """The same edge cases at D = 24 / 40, which take the compile-time-dimension kernels (the warp-pipeline
    kernel hands tiles of 8 columns from warp to warp: sequences shorter than a tile, partial last tiles,
    templates of less than a warp, exact multiples of 32 states and of 8 / 16 / 32 columns)."""
    rng = np.random.default_rng(5)
    S = [1, 2, 3, 31, 32, 33, 64, 65, 97, 672, 673, 1000, 1024]
    T = [1, 9, 8, 7, 16, 31, 32, 33, 64, 100, 15, 37, 3]
    to = np.concatenate([[0], np.cumsum(S)]); so = np.concatenate([[0], np.cumsum(T)])
    for D in (24, 40):
        tm = rng.standard_normal((D, to[-1])); sq = rng.standard_normal((D, so[-1]))
        for fs, bs in [(0, 1), (0, 2)]:
            paths, fc = vcb.DTWs.fit_batch(vcb.DTWs.DTW(fstep=fs, bstep=bs), tm, to, sq, so)
            ref, rfc = oracle.dtw_fit_batch(tm, to, sq, so, fs, bs)
            assert np.array_equal(paths, ref) and np.array_equal(fc, rfc), (D, fs, bs)
        # templates of up to 672 / 97 frames only: the longest template decides the CTA width of the stream kernel
        # (D = 24: 21 + 1 warps; a few warps); in the full batch above it shares the launch with the barrier kernel
        for smax in (672, 97):
          sel = [i for i in range(len(S)) if S[i] <= smax]
          to2 = np.concatenate([[0], np.cumsum([S[i] for i in sel])]); so2 = np.concatenate([[0], np.cumsum([T[i] for i in sel])])
          tm2 = np.concatenate([tm[:, to[i]:to[i + 1]] for i in sel], 1); sq2 = np.concatenate([sq[:, so[i]:so[i + 1]] for i in sel], 1)
          for bs in (1, 2):
            paths, fc = vcb.DTWs.fit_batch(vcb.DTWs.DTW(fstep=0, bstep=bs), tm2, to2, sq2, so2)
            ref, rfc = oracle.dtw_fit_batch(tm2, to2, sq2, so2, 0, bs)
            assert np.array_equal(paths, ref) and np.array_equal(fc, rfc), (D, smax, bs)


@pytest.mark.parametrize("bstep", [2, 1])
def test_many_pairs_per_persistent_cta(vcb, oracle, bstep):
    """More pairs than SMs, template lengths from 1 to 672 and sequence lengths from 1 to 150 in random order: every
    persistent CTA of the stream kernel walks several pairs whose numbers of active warps differ (a warp's tile counters
    towards its neighbours, the ring stages and the service warp's double buffer all carry over from pair to pair)."""
    rng = np.random.default_rng(31 + bstep)
    n = 700
    S = rng.integers(1, 673, size=n); T = rng.integers(1, 151, size=n)
    S[:8] = [672, 1, 641, 32, 640, 33, 31, 671]; T[:8] = [150, 1, 8, 9, 7, 16, 149, 1]
    to = np.concatenate([[0], np.cumsum(S)]).astype(np.int64); so = np.concatenate([[0], np.cumsum(T)]).astype(np.int64)
    tm = np.asfortranarray(rng.standard_normal((24, to[-1]))); sq = np.asfortranarray(rng.standard_normal((24, so[-1])))
    paths, fc = vcb.DTWs.fit_batch(vcb.DTWs.DTW(fstep=0, bstep=bstep), tm, to, sq, so)
    ref, rfc = oracle.dtw_fit_batch(tm, to, sq, so, 0, bstep, nthreads=oracle.max_threads())
    assert np.array_equal(paths, ref) and np.array_equal(fc, rfc)


@pytest.mark.parametrize("S,T", [(1025, 40), (1500, 300), (2500, 120), (4097, 33), (8192, 17)])
def test_long_templates_bit_exact(vcb, oracle, S, T):
    """Templates beyond 1024 frames (5.1 s at the reference's 5 ms shift): the reference has no
    length limit (src/dtw.jl:93-98); several states per thread keep one CTA per pair."""
    tm, to, sq, so = vcb.synth.dtw_pairs(2, 24, (S, S), 1300 + S)
    so = np.array([0, T, 2 * T], dtype=np.int64)
    sq = np.asfortranarray(np.concatenate([sq[:, :T], sq[:, S:S + T]], axis=1))
    for fs, bs, D in [(0, 2, 24), (0, 1, 24), (1, 2, 24), (0, 2, 7)]:
        a, b = np.asfortranarray(tm[:D]), np.asfortranarray(sq[:D])
        paths, fc = vcb.DTWs.fit_batch(vcb.DTWs.DTW(fstep=fs, bstep=bs), a, to, b, so)
        ref, rfc = oracle.dtw_fit_batch(a, to, b, so, fs, bs, nthreads=2)
        assert np.array_equal(paths, ref) and np.array_equal(fc, rfc), (fs, bs, D)


def test_mixed_long_and_short_batch(vcb, oracle):
    tm, to, sq, so = vcb.synth.dtw_pairs(3, 24, (1200, 1700), 1400)
    tm2, to2, sq2, so2 = vcb.synth.dtw_pairs(3, 24, (40, 90), 1401)
    T = np.concatenate([tm, tm2], 1); Q = np.concatenate([sq, sq2], 1)
    TO = np.concatenate([to, to[-1] + to2[1:]]); SO = np.concatenate([so, so[-1] + so2[1:]])
    paths, fc = vcb.DTWs.fit_batch(vcb.DTWs.DTW(fstep=0, bstep=2), T, TO, Q, SO)
    ref, rfc = oracle.dtw_fit_batch(T, TO, Q, SO, 0, 2, nthreads=oracle.max_threads())
    assert np.array_equal(paths, ref) and np.array_equal(fc, rfc)
    with pytest.raises(vcb.VCBError):                      # one CTA holds at most 8192 states
        vcb.DTWs.fit(vcb.DTWs.DTW(), np.zeros((2, 8193)), np.zeros((2, 4)))


def test_ties_follow_reference_order(vcb, oracle):
    # integer-valued features produce many exact ties: candidate order and strict '<' matter
    rng = np.random.default_rng(9)
    tm = rng.integers(0, 3, size=(2, 80)).astype(float); sq = rng.integers(0, 3, size=(2, 90)).astype(float)
    for fs, bs in [(0, 1), (0, 2), (1, 3)]:
        p = vcb.DTWs.fit(vcb.DTWs.DTW(fstep=fs, bstep=bs), tm, sq)
        assert np.array_equal(p, oracle.DTW(fstep=fs, bstep=bs).fit(tm, sq))


def test_online_update(vcb, oracle):
    rng = np.random.default_rng(12)
    tm, sq = rng.standard_normal((6, 40)), rng.standard_normal((6, 25))
    d = vcb.DTWs.DTW(fstep=0, bstep=2)
    vcb.DTWs.set_template(d, tm)
    for t in range(25):
        vcb.DTWs.update(d, sq[:, t])
    o = oracle.DTW(fstep=0, bstep=2); ref = o.fit(tm, sq)
    c, b = o.tables()
    assert np.array_equal(d.costtable, c) and np.array_equal(d.backpointer, b)
    assert np.array_equal(vcb.DTWs.backward(d), ref)


def test_fit_with_tables_matches_the_reference_state(vcb, oracle):
    """fit!(d, template, sequence) leaves d.costtable / d.backpointer behind in the reference
    (src/dtw.jl:122-127); fit(..., tables=True) reproduces them bit for bit."""
    rng = np.random.default_rng(21)
    tm, sq = rng.standard_normal((5, 33)), rng.standard_normal((5, 40))
    d = vcb.DTWs.DTW(fstep=0, bstep=2)
    p = vcb.DTWs.fit(d, tm, sq, tables=True)
    o = oracle.DTW(fstep=0, bstep=2)
    ref = o.fit(tm, sq)
    c, b = o.tables()
    assert np.array_equal(p, ref) and np.array_equal(d.costtable, c) and np.array_equal(d.backpointer, b)
    assert np.array_equal(vcb.DTWs.fit(vcb.DTWs.DTW(fstep=0, bstep=2), tm, sq), ref)


def test_device_entry_point(vcb, oracle):
    import torch
    tm, to, sq, so = vcb.synth.dtw_pairs(5, 24, (100, 140), 77)
    dt = torch.from_numpy(np.ascontiguousarray(tm.T)).cuda(); ds = torch.from_numpy(np.ascontiguousarray(sq.T)).cuda()
    paths, fc = vcb.DTWs.fit_batch(vcb.DTWs.DTW(fstep=0, bstep=2), dt, to, ds, so)
    torch.cuda.synchronize()
    ref, rfc = oracle.dtw_fit_batch(tm, to, sq, so, 0, 2)
    assert np.array_equal(paths.cpu().numpy(), ref) and np.array_equal(fc.cpu().numpy(), rfc)


def test_full_c3_properties(vcb):
    """BASELINE C3 at full size (1000 pairs): properties that need no oracle run."""
    tm, to, sq, so = vcb.synth.config_c3(1000)
    paths, fc = vcb.DTWs.fit_batch(vcb.DTWs.DTW(fstep=0, bstep=2), tm, to, sq, so)
    for p in range(0, 1000, 37):
        path = paths[so[p]:so[p + 1]]
        S = to[p + 1] - to[p]
        step = np.diff(path)
        assert path.min() >= 1 and path.max() <= S and step.min() >= 0 and step.max() <= 2
        # the accumulated cost along the returned path equals the reported final cost
        t_ = tm[:, to[p]:to[p + 1]]; s_ = sq[:, so[p]:so[p + 1]]
        # first frame: best predecessor in the initial column (cost[:,1] = 1:S); afterwards the
        # predecessor is the previous path state
        prev_candidates = np.arange(1, S + 1, dtype=float)
        total = None
        for t in range(len(path)):
            i = path[t]
            oc = 0.0
            for k in range(24):
                dlt = s_[k, t] - t_[k, i - 1]
                oc = oc + dlt * dlt
            if t == 0:
                best = prev_candidates[i - 1] + oc + 1.0
                for j in range(max(1, i - 2), i + 1):
                    tr = 0.0 if i == j + 1 else (1.0 if i == j else 2.0)
                    best = min(best, prev_candidates[j - 1] + oc + tr)
                total = best
            else:
                j = path[t - 1]
                tr = 0.0 if i == j + 1 else (1.0 if i == j else 2.0)
                total = total + oc + tr
        assert total == fc[p]
