#!/usr/bin/env python
"""bench.py -- frame-pairs/s of the refine+rectify hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA path through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  # reference arm: the CPU restatement
                                                            # (oracle/) on all host threads

A "step" = one pass of the hot path over one 1920x1080 frame pair: nonLinearRefinement ->
sign fix -> depth raster -> setPose -> backProject -> interpolateCrackyImage (main.cc:457-523).
`value` = pairs/s with every input already resident in HBM (CUDA events, max over ranks);
`e2e`   = pairs/s through the same C-ABI call with HOST (pinned) buffers, host<->device copies
          inside the timed region.
Multi-GPU: frame pairs are independent, so ranks shard them with no data-path collective
(weak scaling: every rank processes `steps` pairs); one final max-reduce of the elapsed time.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "rs-aware-differential-sfm_b200"

ROWS, COLS = 1080, 1920
WORKLOAD = ("synthetic analytic RS flow 1920x1080 (galaxy_stabil K, gamma 0.95), piecewise-planar depth "
            "(64 Voronoi planes, Z in [2,30]), constant-acceleration trajectory k=0.5, sigma 0.3 px noise + 5% "
            "outliers, refine (const-acc, 7 motion parameters + one inverse depth per inlier) started from the "
            "RANSAC winner (H=16 hypotheses, tol 0.05), then per-scanline GS rectification + crack fill")
ALGO_BYTES_PASS_A = 24.0   # read flow 16 B + inverse depth 8 B per residual block
ALGO_BYTES_PASS_B = 32.0   # read flow 16 B + inverse depth 8 B, write candidate inverse depth 8 B
N_PAIRS = 3                # distinct pairs cycled through: >= 3 x ~170 MB of inputs, larger than the 126 MB L2
CONST_ACC = True           # headline workload: constant-acceleration trajectory (k estimated); --const-vel: k = 0 fixed


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while a timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = str(index)
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-i", self.index, "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def gen_pair(synth, seed):
    return synth.make_pair(ROWS, COLS, "galaxy_stabil", gamma=0.95, v=(0.30, 0.05, 0.02), w=(0.002, -0.004, 0.0087),
                           k=0.5 if CONST_ACC else 0.0, seed=seed, noise_sigma_px=0.3, outlier_frac=0.05)


def prepare_pair_gpu(ctx, synth, torch, seed, H=16, tol=0.05):
    """Upstream stages (flatten, alpha, RANSAC, consensus gather) on the GPU, outside any timed
    region; returns device-resident inputs of the refine+rectify step and pinned host copies."""
    P = gen_pair(synth, seed)
    dev = torch.device("cuda", torch.cuda.current_device())
    flow_img = torch.from_numpy(P["flow_img"]).to(dev)
    n, coord, flow, cpx, fpx, pidx = ctx.flatten(flow_img, P["K4"], P["gamma"])
    coord, flow, cpx, fpx = coord[:2 * n], flow[:2 * n], cpx[:2 * n], fpx[:2 * n]
    alpha, alpha_k = ctx.alpha(fpx, cpx, n, ROWS, P["gamma"])
    samples = synth.sample_list(n, H, seed=seed + 100)
    R = ctx.ransac(coord, flow, alpha, alpha_k, n, CONST_ACC, samples, tol)
    inl, a_in, ak_in, ix, m = ctx.gather_inliers(coord, alpha, alpha_k, n, R["mask"], R["inv_depth"])
    image = torch.from_numpy(P["image"]).to(dev)
    d = dict(flow=flow.contiguous(), inliers3=inl.contiguous(), alpha=a_in.contiguous(), alpha_k=ak_in.contiguous(),
             image=image, m=m, n=n, v=R["v"], w=R["w"], k=R["k"], K4=P["K4"], gamma=P["gamma"])
    d["out"] = (torch.empty(m, dtype=torch.float64, device=dev), torch.empty(ROWS * COLS, dtype=torch.float64, device=dev),
                torch.empty_like(image))
    keep = []

    def pin(t):
        t = t.cpu().pin_memory()
        keep.append(t)
        return t.numpy()

    # (write-combined pinned inputs -- capi.HostBuffer(write_combined=True) -- were measured: no difference)
    h = dict(flow=pin(d["flow"][:2 * m]), inliers3=pin(d["inliers3"]), alpha=pin(d["alpha"]), alpha_k=pin(d["alpha_k"]),
             image=pin(image))
    h["out"] = (pin(d["out"][0]), pin(d["out"][1]), pin(d["out"][2]))
    h["_keep"] = keep
    d["host"] = h
    return d


def prepare_pair_cpu(O, synth, seed, H=16, tol=0.05):
    """Same upstream stages on the CPU oracle (reference arm: no GPU code anywhere on its path)."""
    P = gen_pair(synth, seed)
    n, coord, flow, cpx, fpx = O.flatten(P["flow_img"], P["K4"], P["gamma"])
    alpha = O.get_alpha(fpx, n, ROWS, P["gamma"])
    alpha_k = O.get_alpha_k(cpx, fpx, n, ROWS, P["gamma"])
    samples = synth.sample_list(n, H, seed=seed + 100)
    R = O.ransac(coord[:2 * n], flow[:2 * n], alpha, alpha_k, n, CONST_ACC, tol, samples=samples)
    inl, a_in, ak_in = O.gather_inliers(coord, alpha, alpha_k, n, R["mask"], R["inv_depth"])
    m = len(a_in)
    h = dict(flow=flow[:2 * m].copy(), inliers3=inl, alpha=a_in, alpha_k=ak_in, image=P["image"])
    return dict(host=h, m=m, n=n, v=R["v"], w=R["w"], k=R["k"], K4=P["K4"], gamma=P["gamma"])


def step_device(ctx, capi, p):
    return ctx.refine_rectify(p["flow"], p["inliers3"], p["alpha"], p["alpha_k"], p["m"], p["v"], p["w"], p["k"], CONST_ACC, False,
                              p["image"], p["K4"], p["gamma"], layout=capi.DEPTH_ROWMAJOR, out=p["out"])


def step_host(ctx, capi, p):
    h = p["host"]
    return ctx.refine_rectify(h["flow"], h["inliers3"], h["alpha"], h["alpha_k"], p["m"], p["v"], p["w"], p["k"], CONST_ACC, False,
                              h["image"], p["K4"], p["gamma"], layout=capi.DEPTH_ROWMAJOR, out=h["out"])


def bind_to_gpu_numa_node(index):
    """Pin this rank's host threads to the CPUs next to its GPU (NVML's ideal affinity) before any
    pinned buffer is allocated: first-touch then puts the staging pages on the GPU's own NUMA node,
    which is what keeps the host<->device copies of 8 ranks from crossing the socket interconnect."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:                                                 # CUDA ordinal -> NVML handle by PCI address
            import torch
            p = torch.cuda.get_device_properties(index)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(("%08x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)).encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception as e:                                   # affinity is an optimisation, never a requirement
        print("bench: CPU affinity not set (%s)" % e, file=sys.stderr)


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    bind_to_gpu_numa_node(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    capi = importlib.import_module(PKG + ".capi")
    synth = importlib.import_module(PKG + ".synth")
    stream = torch.cuda.current_stream()
    ctx = capi.Context(local, stream=stream.cuda_stream)

    # weak scaling: every rank gets the SAME three pairs (same seeds), so that the per-GPU work -- in
    # particular the data-dependent number of LM iterations -- does not change with the rank count
    pairs = [prepare_pair_gpu(ctx, synth, torch, 1000 + i) for i in range(N_PAIRS)]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(ctx, capi, pairs[i % N_PAIRS])
        barrier()
        l0 = ctx.launch_count()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        its = 0
        e0.record(stream)
        for i in range(steps):
            r = fn(ctx, capi, pairs[(warmup + i) % N_PAIRS])
            its += r["summary"]["iterations"]
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, ctx.launch_count() - l0, its

    def seq_entries(i0, count, host):
        out = []
        for i in range(count):
            p = pairs[(i0 + i) % N_PAIRS]
            src = p["host"] if host else p
            out.append(dict(flow=src["flow"], inliers3=src["inliers3"], alpha=src["alpha"], alpha_k=src["alpha_k"], image=src["image"],
                            m=p["m"], v=p["v"], w=p["w"], k=p["k"], out=src["out"]))
        return out

    def timed_sequence(host, steps, warmup):
        """K steps = one rsdsfm_refine_rectify_sequence call over K frame pairs (upload i+1 | compute i |
        download i-1 on three streams; with device buffers only the result collection is deferred)."""
        p0 = pairs[0]
        run = lambda ent: ctx.refine_rectify_sequence(ent, CONST_ACC, False, p0["K4"], p0["gamma"], layout=capi.DEPTH_ROWMAJOR)
        run(seq_entries(0, warmup, host))
        ent = seq_entries(warmup, steps, host)
        barrier()
        l0 = ctx.launch_count()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        res = run(ent)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, ctx.launch_count() - l0, sum(r["summary"]["iterations"] for r in res)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches, its = timed_sequence(False, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    lm_ms = 0.0
    for i in range(min(args.steps, N_PAIRS)):          # ms per LM iteration: the solve's own CUDA-event time
        r = step_device(ctx, capi, pairs[i])
        lm_ms += r["summary"]["device_ms"] / max(r["summary"]["iterations"], 1)
    lm_ms /= min(args.steps, N_PAIRS)
    ms_e2e, _, _ = timed_sequence(True, args.steps, max(3, min(args.warmup, 3)))
    # the same step through one synchronous rsdsfm_refine_rectify call per pair (no overlap between pairs)
    # ... which is also where the dominant kernel is timed ALONE on the whole GPU for the roofline (in a
    # sequence two solves share the SMs, so a kernel's own duration says little about the machine)
    ctx.profile_enable(True)
    ms_single, _, _ = timed(step_device, args.steps, 3)
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    ms_single_host, _, _ = timed(step_host, args.steps, 3)

    p0 = pairs[0]
    h2d = sum(p0["host"][k].nbytes for k in ("flow", "inliers3", "alpha", "alpha_k", "image"))
    d2h = sum(a.nbytes for a in p0["host"]["out"])
    peak, peak_src = load_peaks()
    # Dominant kernel = k_lm_persistent (one launch = one whole LM solve).  Algorithmic bytes per
    # launch: 24 B per residual block for the initial evaluation phase + 56 B per
    # LM iteration (candidate step 32 B + evaluation 24 B; here fused into one sweep);
    # duration = CUDA events around the launch on the launching stream, measured live.
    n_launch = max(prof["kernel_launches"], 1)
    algo_bytes = (ALGO_BYTES_PASS_A * prof["pass_a_blocks"] + (ALGO_BYTES_PASS_A + ALGO_BYTES_PASS_B) * prof["pass_b_blocks"]) / n_launch
    k_t = prof["kernel_ms"] / n_launch * 1e-3
    achieved = algo_bytes / k_t / 1e9 if k_t > 0 else 0.0
    f_t = prof["pass_b_ms"] / max(prof["pass_b_launches"], 1) * 1e-3
    f_blocks = prof["pass_b_blocks"] / max(prof["pass_b_launches"], 1)
    f_loop = prof["b_loop_ms"] / max(prof["pass_b_launches"], 1) * 1e-3
    traffic = None
    tp = os.path.join(ROOT, "profiles", "lm_kernel_traffic.json")     # dram bytes per launch from the committed ncu capture
    if os.path.exists(tp) and CONST_ACC:
        with open(tp) as f:
            tj = json.load(f)
        # the captured launch ran `lm_iterations_in_captured_launch` iterations; scale to this run's average
        it_cap = float(tj.get("lm_iterations_in_captured_launch", 0)) or None
        it_now = prof["pass_b_launches"] / n_launch
        traffic = tj.get("dram_bytes_per_launch")
        if traffic is not None and it_cap:
            traffic = traffic * (it_now + 0.45) / (it_cap + 0.45)   # initial evaluation streams 24/56 of an iteration
    # FP64 side of the roofline: ~218 double-precision instructions per residual block per fused
    # iteration (SASS count), 64 FP64 lanes/SM -> 148 * 64 * 2 * 1.965 GHz = 37.2 TFLOP/s
    fp64_instr = 218.0
    line = None
    if rank == 0:
        line = {
            "metric": "frame-pairs/sec (refine+rectify, 1080p)", "value": world * args.steps / (ms * 1e-3), "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rows": ROWS, "cols": COLS, "inliers_per_pair": int(np.mean([p["m"] for p in pairs])),
                       "lm_iterations_per_pair": its / args.steps, "sharding": "independent frame pairs per rank, no collective",
                       "l2": "inputs larger than L2: %d distinct pairs cycled (~%.0f MB device inputs each)" % (N_PAIRS, h2d / 1e6)},
            "ms_per_lm_iteration": lm_ms,
            "lm_phase_breakdown_us": {
                "initial_evaluation": {k[2:-3] + "_us": 1e3 * prof[k] / max(prof["pass_a_launches"], 1)
                                       for k in ("a_loop_ms", "a_reduce_ms", "a_ctl_ms", "a_logic_ms")},
                "lm_iteration": {k[2:-3] + "_us": 1e3 * prof[k] / max(prof["pass_b_launches"], 1)
                                 for k in ("b_loop_ms", "b_reduce_ms", "b_ctl_ms", "b_logic_ms")},
                "kernel_ms_per_solve": prof["kernel_ms"] / max(prof["kernel_launches"], 1)},
            "clocks": clocks,
            "e2e": {"value": world * args.steps / (ms_e2e * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e / args.steps,
                    "api": "rsdsfm_refine_rectify_sequence, pinned host buffers: upload of pair i+1 and download of pair i-1 "
                           "overlap the compute of pair i (PCIe-bound: one compute lane)",
                    "single_call_ms_per_step": ms_single_host / args.steps},
            "api": "rsdsfm_refine_rectify_sequence over `steps` pairs, device buffers (two solves share the SMs: even / odd pairs)",
            "single_call_ms_per_step": ms_single / args.steps,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm",
                         "kernel": "k_lm_persistent<%d> (persistent LM solve: fused candidate-step + residual/Jacobian/Schur "
                                   "evaluation sweep per iteration, FP64 grid reduction, on-device controller)" % (7 if CONST_ACC else 6),
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes,
                         "algorithmic_bytes_per_block": {"initial_evaluation": ALGO_BYTES_PASS_A,
                                                         "lm_iteration": ALGO_BYTES_PASS_A + ALGO_BYTES_PASS_B},
                         "avg_launch_us": k_t * 1e6, "launches": prof["kernel_launches"],
                         "lm_iteration_phase": {"avg_us": f_t * 1e6, "pixel_loop_us": f_loop * 1e6,
                                                "achieved_GBps": ((ALGO_BYTES_PASS_A + ALGO_BYTES_PASS_B) * f_blocks / f_t / 1e9) if f_t > 0 else 0.0,
                                                "streamed_bytes_per_block": 64,
                                                "streamed_GBps_in_loop": (64.0 * f_blocks / f_loop / 1e9) if f_loop > 0 else 0.0},
                         "fp64": {"instr_per_block_per_iteration": fp64_instr, "peak_tflops_nominal": 37.2,
                                  "achieved_tflops_in_loop": (2.0 * fp64_instr * f_blocks / f_loop / 1e12) if f_loop > 0 else 0.0}},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(pairs[0], threads=1, budget_s=args.cpu_budget)
        emit(line)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(p, threads, budget_s, min_reps=1):
    """The CPU restatement (oracle/) timed on this box's host cores on the SAME pair: the whole
    refine+rectify step.  threads > 1: that many pairs concurrently (ctypes releases the GIL)."""
    from oracle import pyoracle as O
    O.build()
    h = p["host"]
    args = (h["flow"], h["inliers3"], h["alpha"], h["alpha_k"], p["m"], p["v"], p["w"], p["k"], CONST_ACC, False, h["image"],
            p["K4"], p["gamma"])
    done = [0]
    t0 = time.perf_counter()

    def work():
        while True:
            O.refine_rectify(*args)
            done[0] += 1
            if time.perf_counter() - t0 > budget_s and done[0] >= min_reps * threads:
                break

    ths = [threading.Thread(target=work) for _ in range(threads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    return {"value": done[0] / dt, "unit": "pairs/s", "cores": threads, "kind": "port",
            "sample": "%d full 1080p refine+rectify steps of the bench pair (m=%d) on %d host thread(s), %.1f s"
                      % (done[0], p["m"], threads, dt)}


def run_reference(args):
    """Reference arm: the reference's own CPU algorithm for the path (its sources cannot be
    compiled here -- Eigen/Ceres/OpenCV absent -- so the oracle port stands in), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as O
    O.build()
    synth = importlib.import_module(PKG + ".synth")
    p = prepare_pair_cpu(O, synth, 1000)               # same seeds => same pair as rank 0 of our arm
    threads = os.cpu_count() or 1
    total = args.steps + args.warmup
    budget = min(240.0 / max(total, 1), 20.0)
    vals = []
    for i in range(total):
        cb = cpu_baseline(p, threads, budget_s=budget)
        if i >= args.warmup:
            vals.append(cb)
    v = float(np.mean([c["value"] for c in vals]))
    line = {"impl": "reference", "metric": "frame-pairs/sec (refine+rectify, 1080p)", "value": v, "unit": "pairs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / v if v > 0 else None,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rows": ROWS, "cols": COLS, "inliers_per_pair": p["m"]},
            "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": vals[-1]["sample"]},
            "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_JSON_FD = None


def claim_stdout():
    """stdout carries exactly one JSON line: everything else that writes to fd 1 (NCCL's version
    banner, library chatter) is sent to stderr; emit() writes to the saved descriptor."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--const-vel", action="store_true",
                    help="secondary workload: constant-velocity trajectory and model (the reference's default mode)")
    args = ap.parse_args()
    if args.const_vel:
        global CONST_ACC, WORKLOAD
        CONST_ACC = False
        WORKLOAD = WORKLOAD.replace("constant-acceleration trajectory k=0.5", "constant-velocity trajectory k=0").replace(
            "refine (const-acc, 7 motion parameters", "refine (const-vel, 6 motion parameters")
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
