"""Row split of ONE large frame pair over the GPUs of a node (BASELINE.json config 4): nonLinearRefinement on a
synthetic 3840x2160 pair, the consensus set split into contiguous shares, one share per GPU; per LM iteration the
GPUs exchange one row of sums through peer memory from inside their persistent kernels (include/rsdsfm.h, "row split").

  python tools/row_split_bench.py                                  # 1 GPU (the reference point)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         tools/row_split_bench.py [--rows 2160 --cols 3840 --reps 5] # N = 2, 4, 8

Every rank synthesises the SAME pair (seeded) and runs the upstream stages on all of it (outside the timed region);
rank r then refines residual blocks [r m/N, (r+1) m/N).  Rank 0 prints one JSON line: ms per LM iteration (kernel
device timers and CUDA events, max over ranks), iterations, and the refined motion next to the single-GPU solve of
the same pair that every rank runs first (parity: same iteration count, motion to rounding)."""
import argparse, importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=2160)
ap.add_argument("--cols", type=int, default=3840)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--H", type=int, default=8)
ap.add_argument("--const-vel", action="store_true")
args = ap.parse_args()
import torch
import torch.distributed as dist
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import __graft_entry__ as ge
if rank == 0:
    ge.build()
if world > 1:
    dist.barrier()
capi = importlib.import_module("rs-aware-differential-sfm_b200.capi")
synth = importlib.import_module("rs-aware-differential-sfm_b200.synth")
ca = not args.const_vel
dev = torch.device("cuda", local)
stream = torch.cuda.current_stream()
ctx = capi.Context(local, stream=stream.cuda_stream)
scale = args.cols / 1920.0
K4 = tuple(scale * np.array(synth.INTRINSICS["galaxy_stabil"]))
P = synth.make_pair_device(torch, dev, args.rows, args.cols, K4, gamma=0.95, v=(0.30, 0.05, 0.02), w=(0.002, -0.004, 0.0087),
                           k=0.5 if ca else 0.0, seed=4000, noise_sigma_px=0.3, outlier_frac=0.05)
n, coord, flow, cpx, fpx, pidx = ctx.flatten(P["flow_img"], P["K4"], P["gamma"])
coord, flow, cpx, fpx = coord[:2 * n], flow[:2 * n], cpx[:2 * n], fpx[:2 * n]
alpha, alpha_k = ctx.alpha(fpx, cpx, n, args.rows, P["gamma"])
R = ctx.ransac(coord, flow, alpha, alpha_k, n, ca, synth.sample_list(n, args.H, seed=4100), 0.05)
inl, a_in, ak_in, ix, m = ctx.gather_inliers(coord, alpha, alpha_k, n, R["mask"], R["inv_depth"])
flow = flow.contiguous(); inl = inl.contiguous(); a_in = a_in.contiguous(); ak_in = ak_in.contiguous()
del P, coord, cpx, fpx, alpha, alpha_k, pidx
torch.cuda.synchronize()


def timed(fn, reps):
    fn()                                                   # warm-up (buffers, instruction caches)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        out = fn()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, out


# ---- the whole pair on this GPU alone
ms1, (v1, w1, k1, z1, S1) = timed(lambda: ctx.refine(flow, inl, a_in, ak_in, m, R["v"], R["w"], R["k"], ca), max(args.reps // 2, 1))
line = dict(rows=args.rows, cols=args.cols, residual_blocks=int(m), n_gpus=world, model="const-acc" if ca else "const-vel",
            single_gpu=dict(ms_per_solve=ms1, iterations=S1["iterations"], ms_per_lm_iteration=S1["device_ms"] / max(S1["iterations"], 1),
                            termination=S1["termination"], reason=S1["reason"]))
if world > 1:
    # ---- the group: exchange the mailbox handles, connect, split
    h = ctx.peer_export()
    hs = [None] * world
    dist.all_gather_object(hs, h)
    ctx.peer_connect(hs, rank)
    dist.barrier()
    per = (-(-m // world) + 255) // 256 * 256               # shares are whole tiles of 256 residual blocks
    lo = min(rank * per, m); hi = min(lo + per, m)
    fs, ins, as_, aks = flow[2 * lo:2 * hi].contiguous(), inl[3 * lo:3 * hi].contiguous(), a_in[lo:hi].contiguous(), ak_in[lo:hi].contiguous()
    msN, (vN, wN, kN, zN, SN) = timed(lambda: ctx.refine(fs, ins, as_, aks, hi - lo, R["v"], R["w"], R["k"], ca), args.reps)
    dev_ms = torch.tensor([SN["device_ms"]], dtype=torch.float64, device=dev)
    dist.all_reduce(dev_ms, op=dist.ReduceOp.MAX)
    zerr = float((zN - z1[lo:hi]).abs().div(z1[lo:hi].abs() + 1e-300).max().item()) if hi > lo else 0.0
    zt = torch.tensor([zerr], dtype=torch.float64, device=dev)
    dist.all_reduce(zt, op=dist.ReduceOp.MAX)
    rel = lambda a, b: float(np.max(np.abs(np.asarray(a) - np.asarray(b)) / (np.abs(np.asarray(b)) + 1e-300)))
    line["row_split"] = dict(ms_per_solve=msN, iterations=SN["iterations"], ms_per_lm_iteration=float(dev_ms.item()) / max(SN["iterations"], 1),
                             termination=SN["termination"], reason=SN["reason"], residual_blocks_per_gpu=int(per),
                             speedup_vs_single_gpu=ms1 / msN if msN > 0 else None,
                             parity=dict(iterations_equal=bool(SN["iterations"] == S1["iterations"]),
                                         motion_max_rel_diff=max(rel(vN, v1), rel(wN, w1), abs(kN - k1) / (abs(k1) + 1e-300)),
                                         depth_max_rel_diff=float(zt.item())))
    ctx.peer_disconnect()
if rank == 0:
    print(json.dumps(line))
ctx.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
