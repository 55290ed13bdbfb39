"""Generates tests/golden/pipeline_small.npz.

The reference (ThomasZiegler/RS-aware-differential-SfM) cannot be compiled or run in this
environment (Eigen / Ceres / OpenCV absent, example data absent) and ships no golden vectors, so
these fixtures are produced by the CPU oracle (oracle/rsdsfm_oracle.c) on a small seeded synthetic
pair.  They are REGRESSION vectors: they pin the oracle (and through the GPU parity tests the CUDA
path) against accidental change; they are not outputs of the reference binary.

  python tests/golden/make_golden.py        # rewrites pipeline_small.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def compute(O, synth):
    import helpers
    rows, cols = 48, 64
    K4 = (80.0, 79.0, 32.0, 24.0)
    out = {}
    for tag, k, cacc, seed in (("cv", 0.0, False, 21), ("ca", 0.5, True, 22)):
        c = helpers.make_case(O, synth, rows, cols, K4, k=k, const_acc=cacc, H=6, tol=0.01, seed=seed, noise=0.1, outliers=0.05)
        R = c["ransac"]
        res = O.refine_rectify(c["flow"], c["inliers3"], c["alpha_in"], c["alpha_k_in"], c["m"], R["v"], R["w"], R["k"], cacc, False,
                               c["P"]["image"], K4, c["gamma"])
        out[tag + "_hyps"] = R["hyps"]
        out[tag + "_counts"] = R["counts"].astype(np.int32)
        out[tag + "_best"] = np.array([R["best_idx"]], dtype=np.int32)
        out[tag + "_mask"] = R["mask"].astype(np.uint8)
        out[tag + "_inv_depth"] = R["inv_depth"]
        out[tag + "_motion"] = np.concatenate([res["v"], res["w"], [res["k"]]])
        out[tag + "_z"] = res["z"]
        out[tag + "_iterations"] = np.array([res["summary"]["iterations"]], dtype=np.int32)
        out[tag + "_rectified"] = res["rectified"]
    return out


if __name__ == "__main__":
    import importlib
    from oracle import pyoracle as O
    synth = importlib.import_module("rs-aware-differential-sfm_b200.synth")
    data = compute(O, synth)
    np.savez_compressed(os.path.join(HERE, "pipeline_small.npz"), **data)
    print("wrote", os.path.join(HERE, "pipeline_small.npz"), {k: v.shape for k, v in data.items()})
