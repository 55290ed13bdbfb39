"""One rsdsfm_ransac call on the 1080p bench pair (for launch lists / profiles of the RANSAC stage)."""
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
dev = torch.device("cuda", 0)
ctx = capi.Context(0)
P = synth.make_pair_device(torch, dev, 1080, 1920, "galaxy_stabil", gamma=0.95, k=0.5, seed=1000, noise_sigma_px=0.3, outlier_frac=0.05)
n, coord, flow, cpx, fpx, pidx = ctx.flatten(P["flow_img"], P["K4"], P["gamma"])
coord, flow, cpx, fpx = coord[:2 * n], flow[:2 * n], cpx[:2 * n], fpx[:2 * n]
alpha, alpha_k = ctx.alpha(fpx, cpx, n, 1080, P["gamma"])
smp = synth.sample_list(n, H, seed=1100)
for rep in range(3):
    torch.cuda.synchronize(); t = time.perf_counter(); l0 = ctx.launch_count()
    R = ctx.ransac(coord, flow, alpha, alpha_k, n, True, smp, 0.05)
    torch.cuda.synchronize()
    print("ransac H=%d: %.3f ms, %d launches, best %d, counts %s" % (H, 1e3 * (time.perf_counter() - t), ctx.launch_count() - l0, R["best_idx"], list(R["counts"][:8])))
