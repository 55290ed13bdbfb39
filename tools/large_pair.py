"""One 3840x2160 pair on ONE GPU (BASELINE.json config 4 without the row split): refine + rectify time and
LM iteration time, to judge what a row split could buy.  python tools/large_pair.py"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
ge.build()
capi = importlib.import_module("rs-aware-differential-sfm_b200.capi")
synth = importlib.import_module("rs-aware-differential-sfm_b200.synth")
ROWS, COLS = 2160, 3840
K4 = tuple(2.0 * np.array(synth.INTRINSICS["galaxy_stabil"]))
ctx = capi.Context(0)
dev = torch.device("cuda", 0)
for const_acc in (False, True):
    P = synth.make_pair(ROWS, COLS, K4, gamma=0.95, v=(0.30, 0.05, 0.02), w=(0.002, -0.004, 0.0087), k=0.5 if const_acc else 0.0,
                        seed=1000, noise_sigma_px=0.3, outlier_frac=0.05)
    fi = torch.from_numpy(P["flow_img"]).to(dev); img = torch.from_numpy(P["image"]).to(dev)
    n, coord, flow, cpx, fpx, pidx = ctx.flatten(fi, P["K4"], P["gamma"])
    coord, flow, cpx, fpx = coord[:2 * n], flow[:2 * n], cpx[:2 * n], fpx[:2 * n]
    alpha, alpha_k = ctx.alpha(fpx, cpx, n, ROWS, P["gamma"])
    R = ctx.ransac(coord, flow, alpha, alpha_k, n, const_acc, synth.sample_list(n, 16, seed=1100), 0.05)
    inl, a_in, ak_in, ix, m = ctx.gather_inliers(coord, alpha, alpha_k, n, R["mask"], R["inv_depth"])
    args = (flow.contiguous(), inl.contiguous(), a_in.contiguous(), ak_in.contiguous(), m, R["v"], R["w"], R["k"], const_acc, False, img,
            P["K4"], P["gamma"])
    ctx.refine_rectify(*args); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        r = ctx.refine_rectify(*args)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 3 * 1e3
    s = r["summary"]
    print("4K const_acc=%d m=%d: refine+rectify %.3f ms, LM %.3f ms / %d iterations = %.1f us per iteration; |v|-direction err %.2e, w err %.2e"
          % (const_acc, m, ms, s["device_ms"], s["iterations"], 1e3 * s["device_ms"] / max(s["iterations"], 1),
             np.linalg.norm(r["v"] / np.linalg.norm(r["v"]) - P["v"] / np.linalg.norm(P["v"])), np.abs(r["w"] - P["w"]).max()), flush=True)
