"""Run-to-run determinism probe: the same refine / refine_rectify call repeated must give bitwise
identical results (fixed-order reductions).  python tools/determinism_check.py"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
import __graft_entry__ as ge
ge.build()
capi = importlib.import_module("rs-aware-differential-sfm_b200.capi")
synth = importlib.import_module("rs-aware-differential-sfm_b200.synth")
from oracle import pyoracle as O
ctx = capi.Context(0)
for const_acc in (False, True):
    for p in range(5):
        c = helpers.make_case(O, synth, 120, 160, helpers.small_K(8), k=0.5 if const_acc else 0.0, const_acc=const_acc,
                              H=8, seed=20 + p, sample_seed=40 + p, outliers=0.05 + 0.03 * p)
        R = c["ransac"]
        outs = []
        for rep in range(6):
            r = ctx.refine(c["flow"][:2 * c["m"]], c["inliers3"], c["alpha_in"], c["alpha_k_in"], c["m"], R["v"], R["w"], R["k"], const_acc)
            outs.append(dict(v=r[0], w=r[1], k=r[2], z=r[3], summary=r[4]))
        base = outs[0]
        same = [bool(np.array_equal(o["v"], base["v"]) and np.array_equal(o["w"], base["w"]) and np.array_equal(o["z"], base["z"])) for o in outs]
        print("const_acc", const_acc, "pair", p, "m", c["m"], "iters", [o["summary"]["iterations"] for o in outs],
              "cost", ["%.17g" % o["summary"]["final_cost"] for o in outs[:3]], "identical", same, flush=True)
