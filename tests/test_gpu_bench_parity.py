"""Parity of the CUDA path with the CPU oracle on the configurations bench.py times and DESIGN.md
quotes: the exact bench pairs (1080p, seeds 1000.., k = 0.5,
constant-acceleration model, H = 16) through (i) one synchronous rsdsfm_refine_rectify call and
(ii) the multi-lane device sequence; and one 3840x2160 pair (BASELINE config 4).

Tolerances are BASELINE.json's: motion 1e-6 relative, depth 1e-4 relative at the median and 1e-3 at
p99, rectified 8-bit image within 1 grey level on >= 99.9 % of the pixels; LM iteration counts and
termination reasons equal."""
import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def _check(got, ref, what):
    for key in ("v", "w"):
        scale = np.abs(ref[key]).max()
        err = np.abs(np.asarray(got[key]) - ref[key]).max() / scale
        assert err < 1e-6, "%s %s: rel err %.3e" % (what, key, err)
    assert abs(got["k"] - ref["k"]) <= 1e-6 * max(abs(ref["k"]), 1e-3), "%s k: %r vs %r" % (what, got["k"], ref["k"])
    for key in ("iterations", "termination", "reason", "num_successful", "num_unsuccessful"):
        assert got["summary"][key] == ref["summary"][key], "%s summary.%s: %r vs %r" % (what, key, got["summary"][key], ref["summary"][key])
    rel = helpers.rel_err(got["z"], ref["z"])
    assert np.median(rel) < 1e-4 and np.percentile(rel, 99) < 1e-3, "%s depth: median %.2e p99 %.2e" % (what, np.median(rel), np.percentile(rel, 99))
    diff = np.abs(np.asarray(got["rectified"]).astype(np.int32) - ref["rectified"].astype(np.int32)).max(axis=2)
    frac = float((diff <= 1).mean())
    assert frac >= 0.999, "%s rectified: %.5f within 1 grey level" % (what, frac)


def _prepare(ctx, synth, torch, rows, cols, intr, seed, k, const_acc, H=16, tol=0.05):
    """bench.py's prepare_pair_gpu: upstream stages on the GPU, device-resident step inputs."""
    P = synth.make_pair(rows, cols, intr, gamma=0.95, v=(0.30, 0.05, 0.02), w=(0.002, -0.004, 0.0087), k=k, seed=seed,
                        noise_sigma_px=0.3, outlier_frac=0.05)
    dev = torch.device("cuda", 0)
    n, coord, flow, cpx, fpx, pidx = ctx.flatten(torch.from_numpy(P["flow_img"]).to(dev), P["K4"], P["gamma"])
    coord, flow, cpx, fpx = coord[:2 * n], flow[:2 * n], cpx[:2 * n], fpx[:2 * n]
    alpha, alpha_k = ctx.alpha(fpx, cpx, n, rows, P["gamma"])
    R = ctx.ransac(coord, flow, alpha, alpha_k, n, const_acc, synth.sample_list(n, H, seed=seed + 100), tol)
    inl, a_in, ak_in, ix, m = ctx.gather_inliers(coord, alpha, alpha_k, n, R["mask"], R["inv_depth"])
    d = dict(flow=flow.contiguous(), inliers3=inl.contiguous(), alpha=a_in.contiguous(), alpha_k=ak_in.contiguous(),
             image=torch.from_numpy(P["image"]).to(dev), m=m, v=R["v"], w=R["w"], k=R["k"], K4=P["K4"], gamma=P["gamma"])
    return d


def _oracle_step(oracle, d, const_acc):
    h = {k: d[k].cpu().numpy() for k in ("flow", "inliers3", "alpha", "alpha_k", "image")}
    return oracle.refine_rectify(h["flow"][:2 * d["m"]], h["inliers3"], h["alpha"], h["alpha_k"], d["m"], d["v"], d["w"], d["k"],
                                 const_acc, False, h["image"], d["K4"], d["gamma"])


def _host(r):
    return {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in r.items()}


@pytest.fixture(scope="module")
def bench_pairs(ctx, synth, oracle):
    import torch
    pairs = [_prepare(ctx, synth, torch, 1080, 1920, "galaxy_stabil", 1000 + i, 0.5, True) for i in range(3)]
    refs = [_oracle_step(oracle, d, True) for d in pairs]
    return pairs, refs


def test_bench_pair_single_call_matches_oracle(ctx, capi, bench_pairs):
    """bench.py's pair 0 (seed 1000, k = 0.5, const-acc NF = 7, H = 16, 1080p) through rsdsfm_refine_rectify."""
    pairs, refs = bench_pairs
    d = pairs[0]
    got = _host(ctx.refine_rectify(d["flow"], d["inliers3"], d["alpha"], d["alpha_k"], d["m"], d["v"], d["w"], d["k"], True, False,
                                   d["image"], d["K4"], d["gamma"]))
    _check(got, refs[0], "single call")
    assert d["m"] > 1900000 and refs[0]["summary"]["iterations"] >= 10     # the workload bench.py describes


def test_bench_pairs_two_lane_device_sequence_matches_oracle(ctx, capi, bench_pairs):
    """The `value` path of bench.py: rsdsfm_refine_rectify_sequence with device buffers (several compute lanes, each LM
    solve on a fraction of the SMs) over the three bench pairs, twice over -- every pair against the oracle, and
    against the single call (the solver's sums do not depend on the grid)."""
    pairs, refs = bench_pairs
    ent = [dict(flow=d["flow"], inliers3=d["inliers3"], alpha=d["alpha"], alpha_k=d["alpha_k"], image=d["image"], m=d["m"],
                v=d["v"], w=d["w"], k=d["k"]) for d in pairs + pairs]
    res = ctx.refine_rectify_sequence(ent, True, False, pairs[0]["K4"], pairs[0]["gamma"])
    assert len(res) == 6
    for i, r in enumerate(res):
        assert r["status"] == 0
        _check(_host(r), refs[i % 3], "sequence pair %d" % i)
    # the two occurrences of a pair ran on different lanes
    for i in range(3):
        assert np.array_equal(res[i]["rectified"].cpu().numpy(), res[i + 3]["rectified"].cpu().numpy())
    d = pairs[0]
    one = ctx.refine_rectify(d["flow"], d["inliers3"], d["alpha"], d["alpha_k"], d["m"], d["v"], d["w"], d["k"], True, False,
                             d["image"], d["K4"], d["gamma"])
    assert np.array_equal(one["v"], res[0]["v"]) and np.array_equal(one["w"], res[0]["w"]) and one["k"] == res[0]["k"]
    assert np.array_equal(one["z"].cpu().numpy(), res[0]["z"].cpu().numpy())
    assert np.array_equal(one["rectified"].cpu().numpy(), res[0]["rectified"].cpu().numpy())


def test_4k_pair_matches_oracle(ctx, capi, synth, oracle):
    """BASELINE config 4's pair on one GPU: 3840 x 2160, K x 2, same seeds (8 294 400 residual blocks);
    constant-velocity model to bound the oracle's CPU time (~20 s)."""
    import torch
    K4 = tuple(2.0 * np.array(synth.INTRINSICS["galaxy_stabil"]))
    d = _prepare(ctx, synth, torch, 2160, 3840, K4, 1000, 0.0, False, H=4)
    ref = _oracle_step(oracle, d, False)
    got = _host(ctx.refine_rectify(d["flow"], d["inliers3"], d["alpha"], d["alpha_k"], d["m"], d["v"], d["w"], d["k"], False, False,
                                   d["image"], d["K4"], d["gamma"]))
    _check(got, ref, "4K")
    assert d["m"] > 7500000
