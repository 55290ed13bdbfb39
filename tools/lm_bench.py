"""LM solve micro-benchmark on the bench pair (1080p, seed 1000, const-acc unless --const-vel): per-iteration
time, phase breakdown, and the refined motion (to compare kernel variants: RSDSFM_LM_VARIANT=1|2).
python tools/lm_bench.py [--reps N] [--const-vel] [--rows R --cols C]"""
import argparse, importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--const-vel", action="store_true")
ap.add_argument("--rows", type=int, default=1080)
ap.add_argument("--cols", type=int, default=1920)
ap.add_argument("--H", type=int, default=16)
args = ap.parse_args()
import torch
import __graft_entry__ as ge
ge.build()
capi = importlib.import_module("rs-aware-differential-sfm_b200.capi")
synth = importlib.import_module("rs-aware-differential-sfm_b200.synth")
ca = not args.const_vel
scale = args.cols / 1920.0
K4 = tuple(scale * np.array(synth.INTRINSICS["galaxy_stabil"]))
ctx = capi.Context(0)
dev = torch.device("cuda", 0)
P = synth.make_pair(args.rows, args.cols, K4, gamma=0.95, v=(0.30, 0.05, 0.02), w=(0.002, -0.004, 0.0087), k=0.5 if ca else 0.0,
                    seed=1000, noise_sigma_px=0.3, outlier_frac=0.05)
n, coord, flow, cpx, fpx, pidx = ctx.flatten(torch.from_numpy(P["flow_img"]).to(dev), P["K4"], P["gamma"])
coord, flow, cpx, fpx = coord[:2 * n], flow[:2 * n], cpx[:2 * n], fpx[:2 * n]
alpha, alpha_k = ctx.alpha(fpx, cpx, n, args.rows, P["gamma"])
R = ctx.ransac(coord, flow, alpha, alpha_k, n, ca, synth.sample_list(n, args.H, seed=1100), 0.05)
inl, a_in, ak_in, ix, m = ctx.gather_inliers(coord, alpha, alpha_k, n, R["mask"], R["inv_depth"])
flow = flow.contiguous(); inl = inl.contiguous(); a_in = a_in.contiguous(); ak_in = ak_in.contiguous()
run = lambda: ctx.refine(flow, inl, a_in, ak_in, m, R["v"], R["w"], R["k"], ca)
for _ in range(3):
    r = run()
ctx.profile_enable(True)
for _ in range(args.reps):
    r = run()
prof = ctx.profile_read()
ctx.profile_enable(False)
v, w, k, z, S = r
nb = max(prof["pass_b_launches"], 1); nk = max(prof["kernel_launches"], 1)
out = dict(variant=os.environ.get("RSDSFM_LM_VARIANT", "default"), m=m, iterations=S["iterations"], termination=S["termination"],
           reason=S["reason"], final_cost=S["final_cost"], v=list(v), w=list(w), k=k, z_sum=float(z.sum().item()),
           kernel_ms=prof["kernel_ms"] / nk, us_per_iteration=1e3 * prof["kernel_ms"] / nk / max(S["iterations"], 1),
           iter_phase_us=1e3 * prof["pass_b_ms"] / nb, loop_us=1e3 * prof["b_loop_ms"] / nb, reduce_us=1e3 * prof["b_reduce_ms"] / nb,
           ctl_us=1e3 * prof["b_ctl_ms"] / nb, logic_us=1e3 * prof["b_logic_ms"] / nb,
           init_phase_us=1e3 * prof["pass_a_ms"] / max(prof["pass_a_launches"], 1), init_loop_us=1e3 * prof["a_loop_ms"] / max(prof["pass_a_launches"], 1))
out["phases_per_solve"] = nb / nk
out["outside_phases_us"] = 1e3 * (prof["kernel_ms"] - prof["pass_b_ms"] - prof["pass_a_ms"]) / nk
print(json.dumps(out))
