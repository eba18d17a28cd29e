"""Generates the committed fixtures in tests/golden/ (run in the build container, where the
reference checkout /root/reference exists):

    python tests/golden/make_golden.py

  gmm32_order40_diff.npz   the reference's real model test/models/clb_to_slt_gmm32_order40_diff.jld
                           (read with voiceconversion.jl_b200/jld.py; values unchanged)
  dtw_reference_tests.json the two known-answer vectors of the reference's test/dtw.jl:7-31
  fbf_c0.npz               BASELINE config C0 stand-in: 96 frames drawn from the real model's source
                           marginal (seed 1000) + the oracle's vc() output
  traj_small.npz           small trajectory case (synthetic model from seed) + oracle vc() output
  dtw_random.npz           random DTW pairs + oracle paths for several (fstep, bstep) windows

Outputs of the oracle are stored so the GPU box (which has no /root/reference) and later rounds
can detect any drift of either side.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import vcb200  # noqa: E402
from oracle import oracle as O  # noqa: E402

REF = "/root/reference"


def main():
    O.build()
    d = vcb200.jld.load(os.path.join(REF, "test/models/clb_to_slt_gmm32_order40_diff.jld"))
    assert d["diff"] is True and d["n_components"] == 32
    np.savez_compressed(os.path.join(HERE, "gmm32_order40_diff.npz"), weights=d["weights"], means=d["means"],
                        covars=d["covars"], diff=np.array(d["diff"]))

    # reference test/dtw.jl:7-31 (values typed from the test file; matrices are D x frames)
    dtw_tests = [
        {"source": "test/dtw.jl:7-19", "bstep": 1, "fstep": 0,
         "template": [[1, 2, 3], [1, 2, 4], [1, 8, 5], [10, 3, 6]],
         "sequence": [[1, 2, 3], [1, 2, 4], [1, 2, 5], [1, 8, 5], [10, 3, 6]],
         "expected": [1, 2, 2, 3, 4]},
        {"source": "test/dtw.jl:21-31", "bstep": 1, "fstep": 0,
         "template": [[0], [1], [2], [3], [4], [5]],
         "sequence": [[0], [0], [1], [2], [3], [4], [4], [5]],
         "expected": [1, 1, 2, 3, 4, 5, 5, 6]},
    ]
    json.dump(dtw_tests, open(os.path.join(HERE, "dtw_reference_tests.json"), "w"), indent=1)

    # C0 stand-in
    gm = vcb200.synth.JointGMM(d["weights"], d["means"], d["covars"])
    fm = vcb200.synth.fbf_feature_matrix(gm, 96, 1000)
    g = O.GMMMap(*gm)
    out = g.vc(fm)
    post = np.stack([g.predict_proba(fm[1:, t]) for t in range(fm.shape[1])], axis=1)
    np.savez_compressed(os.path.join(HERE, "fbf_c0.npz"), fm=fm, out=out, post=post)

    # trajectory
    gm2 = vcb200.synth.random_joint_gmm(77, 8, 24)     # Ds = 6
    fm2, off2 = vcb200.synth.trajectory_utterances(gm2, 3, (25, 60), 78)
    g2 = O.GMMMap(*gm2)
    out2 = O.vc_traj_batch(g2, 40, fm2, off2)
    np.savez_compressed(os.path.join(HERE, "traj_small.npz"), seed=np.array([77, 8, 24]), fm=fm2, offsets=off2,
                        limit=np.array(40), out=out2)

    # DTW
    tm, to, sq, so = vcb200.synth.dtw_pairs(6, 5, (40, 70), 2024, noise=0.3)
    store = {"tmpl": tm, "tmpl_off": to, "seq": sq, "seq_off": so}
    for fs, bs in [(0, 1), (0, 2), (1, 2), (0, 5), (3, 20)]:
        paths, fc = O.dtw_fit_batch(tm, to, sq, so, fs, bs)
        store[f"paths_f{fs}_b{bs}"] = paths
        store[f"cost_f{fs}_b{bs}"] = fc
    np.savez_compressed(os.path.join(HERE, "dtw_random.npz"), **store)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
