"""Per-stage GPU times of the whole a2..a15 path at 1080p (device-resident inputs).
python tools/stage_times.py [H]"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
ge.build()
capi = importlib.import_module("rs-aware-differential-sfm_b200.capi")
synth = importlib.import_module("rs-aware-differential-sfm_b200.synth")
H = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ROWS, COLS = 1080, 1920
ctx = capi.Context(0)
dev = torch.device("cuda", 0)
for const_acc in (False, True):
    P = synth.make_pair(ROWS, COLS, "galaxy_stabil", gamma=0.95, v=(0.30, 0.05, 0.02), w=(0.002, -0.004, 0.0087),
                        k=0.5 if const_acc else 0.0, seed=1000, noise_sigma_px=0.3, outlier_frac=0.05)
    fi = torch.from_numpy(P["flow_img"]).to(dev); img = torch.from_numpy(P["image"]).to(dev)

    def t(fn, reps=5):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            r = fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3, r

    ms_flat, (n, coord, flow, cpx, fpx, pidx) = t(lambda: ctx.flatten(fi, P["K4"], P["gamma"]))
    coord, flow, cpx, fpx = coord[:2 * n], flow[:2 * n], cpx[:2 * n], fpx[:2 * n]
    ms_alpha, (alpha, alpha_k) = t(lambda: ctx.alpha(fpx, cpx, n, ROWS, P["gamma"]))
    samples = synth.sample_list(n, H, seed=1100)
    ms_ransac, R = t(lambda: ctx.ransac(coord, flow, alpha, alpha_k, n, const_acc, samples, 0.05), reps=3)
    ms_gather, (inl, a_in, ak_in, ix, m) = t(lambda: ctx.gather_inliers(coord, alpha, alpha_k, n, R["mask"], R["inv_depth"]))
    ms_rr, rr = t(lambda: ctx.refine_rectify(flow.contiguous(), inl.contiguous(), a_in.contiguous(), ak_in.contiguous(), m, R["v"], R["w"], R["k"],
                                             const_acc, False, img, P["K4"], P["gamma"]))
    ms_pipe, pp = t(lambda: ctx.pipeline_pair(fi, img, P["K4"], P["gamma"], 0.05, const_acc, samples=samples), reps=3)
    print("const_acc=%d H=%d n=%d m=%d | flatten %.3f alpha %.3f ransac %.3f (%.3f/hyp) gather %.3f refine+rectify %.3f (%d it) | pipeline_pair %.3f ms"
          % (const_acc, H, n, m, ms_flat, ms_alpha, ms_ransac, ms_ransac / H, ms_gather, ms_rr, rr["summary"]["iterations"], ms_pipe), flush=True)
