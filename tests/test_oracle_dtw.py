"""Oracle vs the reference's own DTW known-answer tests (test/dtw.jl) + structural properties."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN


def _cases():
    return json.load(open(os.path.join(GOLDEN, "dtw_reference_tests.json")))


@pytest.mark.parametrize("idx", [0, 1])
def test_reference_known_answers(oracle, idx):
    c = _cases()[idx]
    d = oracle.DTW(bstep=c["bstep"], fstep=c["fstep"])
    tm = np.array(c["template"], dtype=float).T
    sq = np.array(c["sequence"], dtype=float).T
    assert d.fit(tm, sq).tolist() == c["expected"]           # test/dtw.jl:17-18, 29-30


def test_online_update_equals_fit(oracle):
    rng = np.random.default_rng(5)
    tm, sq = rng.standard_normal((7, 33)), rng.standard_normal((7, 41))
    d = oracle.DTW(fstep=1, bstep=2)
    path = d.fit(tm, sq)
    c1, b1 = d.tables()
    d2 = oracle.DTW(fstep=1, bstep=2)
    d2.set_template(tm)                                      # src/dtw.jl:53-56
    for t in range(sq.shape[1]):
        d2.update(sq[:, t])                                  # src/dtw.jl:61-90
    c2, b2 = d2.tables()
    assert np.array_equal(c1, c2) and np.array_equal(b1, b2)
    assert np.array_equal(path, d2.backward())
    assert c1.shape == (33, 42) and np.array_equal(c1[:, 0], np.arange(1, 34))   # src/dtw.jl:46-50


@pytest.mark.parametrize("fstep,bstep", [(0, 1), (0, 2), (2, 3)])
def test_path_properties(oracle, fstep, bstep):
    rng = np.random.default_rng(11)
    tm, sq = rng.standard_normal((4, 50)), rng.standard_normal((4, 64))
    p = oracle.DTW(fstep=fstep, bstep=bstep).fit(tm, sq)
    assert p.shape == (64,) and p.min() >= 1 and p.max() <= 50
    step = np.diff(p)
    assert step.max() <= bstep and step.min() >= -fstep     # window of src/dtw.jl:113
    if fstep == 0:
        assert (step >= 0).all()


def test_literal_python_restatement(oracle):
    """Independent, literal Python transcription of src/dtw.jl:93-145 agrees bit-for-bit."""
    rng = np.random.default_rng(3)
    S, T, D, bstep, fstep = 9, 13, 3, 2, 0
    tm, sq = rng.standard_normal((D, S)), rng.standard_normal((D, T))
    cost = np.zeros((S, T + 1)); bp = np.ones((S, T + 1), dtype=int)
    cost[:, 0] = np.arange(1, S + 1); bp[:, 0] = np.arange(1, S + 1)

    def trans(i, j):
        return 0.0 if j == i + 1 else (1.0 if i == j else 2.0)

    for t in range(1, T + 1):
        v = sq[:, t - 1]
        for i in range(1, S + 1):
            oc = 0.0
            for k in range(D):
                oc = oc + (v[k] - tm[k, i - 1]) * (v[k] - tm[k, i - 1])
            mi, mc = i, cost[i - 1, t - 1] + oc + trans(i, i)
            for j in range(i - bstep, i + fstep + 1):
                if j < 1 or j > S:
                    continue
                c = cost[j - 1, t - 1] + oc + trans(j, i)
                if c < mc:
                    mc, mi = c, j
            cost[i - 1, t] = mc; bp[i - 1, t] = mi
    path = np.zeros(T, dtype=int)
    path[-1] = int(np.argmin(cost[:, T])) + 1
    for i in range(T, 1, -1):
        path[i - 2] = bp[path[i - 1] - 1, i]
    d = oracle.DTW(fstep=fstep, bstep=bstep)
    assert np.array_equal(d.fit(tm, sq), path)
    c, b = d.tables()
    assert np.array_equal(c, cost) and np.array_equal(b, bp)


def test_golden_random_pairs(oracle):
    z = np.load(os.path.join(GOLDEN, "dtw_random.npz"))
    for fs, bs in [(0, 1), (0, 2), (1, 2), (0, 5), (3, 20)]:
        paths, fc = oracle.dtw_fit_batch(z["tmpl"], z["tmpl_off"], z["seq"], z["seq_off"], fs, bs, nthreads=2)
        assert np.array_equal(paths, z[f"paths_f{fs}_b{bs}"])
        assert np.array_equal(fc, z[f"cost_f{fs}_b{bs}"])


def test_align_and_push_delta(oracle):
    rng = np.random.default_rng(8)
    src = rng.standard_normal((3, 30))
    tgt = np.repeat(src[:, ::2], 1, axis=1) + 0.01 * rng.standard_normal((3, 15))   # forces 2-steps
    s, newtgt, path = oracle.align(src, tgt)
    assert newtgt.shape == src.shape
    hit = np.zeros(31, dtype=bool); hit[path] = True
    for i in range(path[0], path[-1] + 1):
        if not hit[i]:                                       # src/align.jl:25-32
            assert np.allclose(newtgt[:, i - 1], 0.5 * (newtgt[:, i - 2] + newtgt[:, i]))
    for t in range(15):
        if t == 14 or path[t + 1] != path[t]:
            assert np.array_equal(newtgt[:, path[t] - 1], tgt[:, t])
    pd = oracle.push_delta(src)                              # src/datasets.jl:6-13
    assert np.array_equal(pd[:3], src) and np.array_equal(pd[3:, 0], src[:, 0]) and np.array_equal(pd[3:, -1], src[:, -1])
    assert np.allclose(pd[3:, 1:-1], -0.5 * src[:, :-2] + 0.5 * src[:, 2:])
