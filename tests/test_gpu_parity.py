"""GPU parity tests: the CUDA path (through the C ABI, librsdsfm.so) against the CPU oracle on the
same seeded inputs.  Tolerances are BASELINE.json's: RANSAC inlier counts/sets bit-exact for the
same hypothesis list; refined motion 1e-6 relative; depth 1e-4 relative at the median and 1e-3
at p99; rectified 8-bit image within 1 grey level on >= 99.9 % of pixels (integer stages are
checked bit-exact)."""
import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu

MOTION_RTOL = 1e-6
DEPTH_MED = 1e-4
DEPTH_P99 = 1e-3


def _motion_close(got, ref, what):
    got = np.asarray(got, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    scale = np.abs(ref).max()
    err = np.abs(got - ref).max() / scale
    assert err < MOTION_RTOL, "%s: rel err %.3e" % (what, err)


def _depth_close(got, ref):
    rel = helpers.rel_err(got, ref)
    assert np.median(rel) < DEPTH_MED, "depth median rel err %.3e" % np.median(rel)
    assert np.percentile(rel, 99) < DEPTH_P99, "depth p99 rel err %.3e" % np.percentile(rel, 99)


@pytest.fixture(scope="module")
def case_cv(oracle, synth):
    return helpers.make_case(oracle, synth, 270, 480, helpers.small_K(4), k=0.0, const_acc=False, H=12)


@pytest.fixture(scope="module")
def case_ca(oracle, synth):
    return helpers.make_case(oracle, synth, 270, 480, helpers.small_K(4), k=0.5, const_acc=True, H=12, seed=5)


# ---------------------------------------------------------------------------- preprocessing (bit-exact)
@pytest.mark.parametrize("zero_frac", [0.0, 0.35])
def test_flatten_alpha_bit_exact(ctx, oracle, synth, zero_frac):
    rows, cols = 200, 333
    K4 = helpers.small_K(5)
    P = synth.make_pair(rows, cols, K4, seed=7, noise_sigma_px=0.3, zero_flow_frac=zero_frac)
    n_o, coord_o, flow_o, cpx_o, fpx_o = oracle.flatten(P["flow_img"], K4, P["gamma"])
    n, coord, flow, cpx, fpx, pidx = ctx.flatten(P["flow_img"], K4, P["gamma"])
    assert n == n_o
    if zero_frac > 0:
        assert n < rows * cols
    for a, b in ((coord, coord_o), (flow, flow_o), (cpx, cpx_o), (fpx, fpx_o)):
        assert np.array_equal(a, b)               # including the padded tail (Q3)
    assert np.array_equal(pidx[:n], (cpx[:2 * n:2] * rows + cpx[1:2 * n:2]).astype(np.int32))
    a_o = oracle.get_alpha(fpx_o, n, rows, P["gamma"])
    ak_o = oracle.get_alpha_k(cpx_o, fpx_o, n, rows, P["gamma"])
    a, ak = ctx.alpha(fpx[:2 * n], cpx[:2 * n], n, rows, P["gamma"])
    assert np.array_equal(a, a_o) and np.array_equal(ak, ak_o)


# ---------------------------------------------------------------------------- RANSAC (bit-exact sets)
@pytest.mark.parametrize("which", ["cv", "ca"])
def test_ransac_inlier_sets_bit_exact(ctx, oracle, case_cv, case_ca, which):
    c = case_cv if which == "cv" else case_ca
    R = c["ransac"]
    got = ctx.ransac_score(c["coord"], c["flow"], c["alpha"], c["alpha_k"], c["n"], R["hyps"], c["tol"])
    assert np.array_equal(got["counts"], R["counts"]), (got["counts"], R["counts"])
    assert got["best_idx"] == R["best_idx"]
    assert np.array_equal(got["mask"], R["mask"])
    assert np.array_equal(got["inv_depth"], R["inv_depth"])          # same LM trajectory, same IEEE ops
    np.testing.assert_allclose(got["sumerr"], R["sumerr"], rtol=1e-12)


def test_ransac_from_samples_and_gather(ctx, oracle, case_cv):
    c = case_cv
    R = c["ransac"]
    got = ctx.ransac(c["coord"], c["flow"], c["alpha"], c["alpha_k"], c["n"], False, c["samples"], c["tol"])
    np.testing.assert_allclose(got["hyps"], R["hyps"], rtol=0, atol=1e-12)
    assert got["best_idx"] == R["best_idx"]
    assert np.array_equal(got["counts"], R["counts"])
    inl, a, ak, ix, m = ctx.gather_inliers(c["coord"], c["alpha"], c["alpha_k"], c["n"], got["mask"], got["inv_depth"])
    assert m == c["m"]
    assert np.array_equal(ix, np.nonzero(R["mask"])[0].astype(np.int32))
    if np.array_equal(got["hyps"], R["hyps"]):
        assert np.array_equal(inl, c["inliers3"]) and np.array_equal(a, c["alpha_in"]) and np.array_equal(ak, c["alpha_k_in"])


def test_ransac_tie_in_inlier_count_is_decided_like_the_reference(ctx, oracle, case_cv):
    """minimal.cc:278: equal inlier counts are decided by `<` on the error sums, which the reference accumulates in
    index order.  Copies of one hypothesis with the translation scaled by 1 +- a few ulps have the same consensus set
    (the depth-only LM absorbs the scale) and error sums that differ in the last bits: the winner must be the one the
    sequential sums pick -- the GPU recomputes the contenders' sums in index order for exactly this case."""
    c = case_cv
    base = c["ransac"]["hyps"][c["ransac"]["best_idx"]].copy()
    hyps = np.stack([base.copy() for _ in range(6)])
    for j, s in enumerate([1.0, 1.0 + 2.3e-16, 1.0 - 1.2e-16, 1.0 + 4.5e-16, 1.0, 1.0 - 3.4e-16]):
        hyps[j, 3:6] = base[3:6] * s
    ref = oracle.ransac(c["coord"], c["flow"], c["alpha"], c["alpha_k"], c["n"], False, c["tol"], hyps=hyps)
    got = ctx.ransac_score(c["coord"], c["flow"], c["alpha"], c["alpha_k"], c["n"], hyps, c["tol"])
    assert np.array_equal(got["counts"], ref["counts"])
    assert len(set(int(x) for x in ref["counts"])) < len(ref["counts"]), "the case must contain ties"
    assert got["best_idx"] == ref["best_idx"]
    assert np.array_equal(got["mask"], ref["mask"])
    # the reported sum of the winner, when it won a tie, is the index-order sum: bit-equal to the reference's
    if list(ref["counts"]).count(ref["counts"][ref["best_idx"]]) > 1 and ref["best_idx"] > 0:
        assert got["sumerr"][got["best_idx"]] == ref["sumerr"][ref["best_idx"]]


def test_ransac_degenerate_hypotheses(ctx, oracle, case_cv):
    c = case_cv
    hyps = c["ransac"]["hyps"].copy()
    hyps[1, 6] = np.inf          # k = inf when no real eigenvalue (minimal.cc:75-80)
    hyps[2, 0] = np.nan          # acos domain error (minimal.cc:128)
    hyps[3, 3:6] = 0.0           # pure rotation: every depth column vanishes
    ref = oracle.ransac(c["coord"], c["flow"], c["alpha"], c["alpha_k"], c["n"], False, c["tol"], hyps=hyps)
    got = ctx.ransac_score(c["coord"], c["flow"], c["alpha"], c["alpha_k"], c["n"], hyps, c["tol"])
    assert np.array_equal(got["counts"], ref["counts"])
    assert got["best_idx"] == ref["best_idx"]
    assert np.array_equal(got["mask"], ref["mask"])


def test_estimate_inverse_depths(ctx, oracle, case_cv):
    c = case_cv
    R = c["ransac"]
    ref, sref = oracle.estimate_inverse_depths(c["coord"], c["flow"], c["n"], R["v"], R["w"], R["k"], c["alpha"], c["alpha_k"])
    got, sgot = ctx.estimate_inverse_depths(c["coord"], c["flow"], c["n"], R["v"], R["w"], R["k"], c["alpha"], c["alpha_k"])
    assert sgot["iterations"] == sref["iterations"] and sgot["termination"] == sref["termination"]
    _depth_close(got, ref)


def test_zero_start_gradient_converges_at_iteration_zero(ctx, oracle, case_cv):
    """Ceres' IterationZero ends with step_is_successful = true, so the gradient tolerance is tested before
    any step: with v = 0 the depth column e = -beta A v vanishes, the gradient is exactly 0 and the solve
    stops with CONVERGENCE / gradient tolerance, 0 iterations, depths at their start value 1."""
    c = case_cv
    z3 = np.zeros(3)
    ref, sref = oracle.estimate_inverse_depths(c["coord"], c["flow"], c["n"], z3, c["ransac"]["w"], 0.0, c["alpha"], c["alpha_k"])
    got, sgot = ctx.estimate_inverse_depths(c["coord"], c["flow"], c["n"], z3, c["ransac"]["w"], 0.0, c["alpha"], c["alpha_k"])
    assert sref["termination"] == 0 and sref["reason"] == 3 and sref["iterations"] == 0 and sref["num_successful"] == 1
    for key in ("termination", "reason", "iterations", "num_successful", "num_unsuccessful"):
        assert sgot[key] == sref[key], key
    assert np.array_equal(np.asarray(got), np.ones(c["n"])) and np.array_equal(ref, np.ones(c["n"]))


# ---------------------------------------------------------------------------- refinement
@pytest.mark.parametrize("which,pairing", [("cv", "reference"), ("cv", "fixed"), ("ca", "reference"), ("ca", "fixed")])
def test_refine_matches_oracle(ctx, oracle, case_cv, case_ca, which, pairing):
    import torch
    c = case_cv if which == "cv" else case_ca
    R = c["ransac"]
    fidx = np.nonzero(R["mask"])[0].astype(np.int32) if pairing == "fixed" else None
    v_o, w_o, k_o, z_o, s_o = oracle.nonlinear_refinement(c["flow"], c["inliers3"], c["alpha_in"], c["alpha_k_in"], c["m"],
                                                          R["v"], R["w"], R["k"], c["const_acc"], flow_index=fidx)
    if fidx is None:
        v, w, k, z, s = ctx.refine(c["flow"], c["inliers3"], c["alpha_in"], c["alpha_k_in"], c["m"], R["v"], R["w"], R["k"],
                                   c["const_acc"])
    else:   # repaired pairing goes through device buffers
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        v, w, k, z, s = ctx.refine(dev(c["flow"]), dev(c["inliers3"]), dev(c["alpha_in"]), dev(c["alpha_k_in"]), c["m"],
                                   R["v"], R["w"], R["k"], c["const_acc"], flow_index=dev(fidx))
        z = z.cpu().numpy()
    assert s["termination"] == s_o["termination"] and s["reason"] == s_o["reason"], (s, s_o)
    assert s["iterations"] == s_o["iterations"], (s, s_o)
    assert s["num_successful"] == s_o["num_successful"]
    np.testing.assert_allclose(s["initial_cost"], s_o["initial_cost"], rtol=1e-10)
    np.testing.assert_allclose(s["final_cost"], s_o["final_cost"], rtol=1e-9)
    _motion_close(v, v_o, "v")
    _motion_close(w, w_o, "w")
    if c["const_acc"]:
        assert abs(k - k_o) <= MOTION_RTOL * max(abs(k_o), 1.0)
    else:
        assert k == k_o
    _depth_close(z, z_o)


@pytest.mark.parametrize("const_acc", [False, True])
def test_refine_focus_of_expansion_inside_image(ctx, oracle, synth, const_acc):
    """Forward motion: the focus of expansion lies inside the image, so a few pixels have a
    vanishing depth column and hit Ceres' min_lm_diagonal clamp (the solver's exception list)."""
    c = helpers.make_case(oracle, synth, 240, 320, (1500.0, 1500.0, 160.0, 120.0), k=0.4 if const_acc else 0.0,
                          const_acc=const_acc, H=10, seed=9, v=(0.02, 0.01, 0.30), w=(0.001, 0.002, -0.003), noise=0.05,
                          outliers=0.0, tol=0.05)
    R = c["ransac"]
    assert c["m"] == c["n"]
    # make sure the clamp is really exercised by this case: |e| s_e < 1e-3 for some inlier
    x = c["inliers3"][0::3]; y = c["inliers3"][1::3]
    beta = (2 / (2 + R["k"])) * (c["alpha_in"] + R["k"] * c["alpha_k_in"])
    e = np.hypot(beta * (R["v"][0] - x * R["v"][2]), beta * (R["v"][1] - y * R["v"][2]))
    assert ((e / (1 + e)) ** 2 < 1e-6).sum() > 0
    v_o, w_o, k_o, z_o, s_o = oracle.nonlinear_refinement(c["flow"], c["inliers3"], c["alpha_in"], c["alpha_k_in"], c["m"],
                                                          R["v"], R["w"], R["k"], const_acc)
    v, w, k, z, s = ctx.refine(c["flow"], c["inliers3"], c["alpha_in"], c["alpha_k_in"], c["m"], R["v"], R["w"], R["k"], const_acc)
    assert s["termination"] == s_o["termination"] and s["iterations"] == s_o["iterations"], (s, s_o)
    _motion_close(v, v_o, "v")
    _motion_close(w, w_o, "w")
    assert abs(k - k_o) <= MOTION_RTOL * max(abs(k_o), 1.0)
    _depth_close(z, z_o)


def test_refine_nonfinite_input_fails_like_ceres(ctx, oracle, case_cv):
    c = case_cv
    R = c["ransac"]
    inl = c["inliers3"].copy()
    inl[3 * 5 + 2] = 0.0        # z = 0 -> d = inf: Ceres refuses non-finite parameters, nothing changes
    v_o, w_o, k_o, z_o, s_o = oracle.nonlinear_refinement(c["flow"], inl, c["alpha_in"], c["alpha_k_in"], c["m"], R["v"],
                                                          R["w"], R["k"], False)
    v, w, k, z, s = ctx.refine(c["flow"], inl, c["alpha_in"], c["alpha_k_in"], c["m"], R["v"], R["w"], R["k"], False)
    assert s_o["termination"] == 2 and s["termination"] == 2
    assert np.array_equal(v, R["v"]) and np.array_equal(w, R["w"])
    assert np.array_equal(z, z_o)


# ---------------------------------------------------------------------------- glue + rectification (bit-exact)
def test_glue_pose_splat_cracks_bit_exact(ctx, oracle, case_cv):
    c = case_cv
    R = c["ransac"]
    rows, cols, K4 = c["rows"], c["cols"], c["K4"]
    v_o, w_o, k_o, z_o, _ = oracle.nonlinear_refinement(c["flow"], c["inliers3"], c["alpha_in"], c["alpha_k_in"], c["m"],
                                                        R["v"], R["w"], R["k"], False)
    inl = c["inliers3"].copy()
    inl[2::3] = z_o
    for flip in (1.0, -1.0):           # exercise the mean-depth sign fix (main.cc:475-478)
        inl_f = inl.copy(); inl_f[2::3] *= flip
        vin = v_o * flip
        i_o, vv_o, dm_o, img_o, _ = oracle.depth_glue(inl_f, c["m"], vin, K4, rows, cols, want_img=True)
        i_g, vv_g, dm_g, img_g = ctx.depth_glue(inl_f, c["m"], vin, K4, rows, cols, want_img=True)
        assert np.array_equal(i_g, i_o) and np.array_equal(vv_g, vv_o)
        assert np.array_equal(dm_g, dm_o)
        assert np.array_equal(img_g.reshape(rows, cols), img_o)
    Rm_o, t_o = oracle.set_relative_pose(vv_o, w_o, k_o, c["gamma"], rows)
    Rm_g, t_g = ctx.set_relative_pose(vv_o, w_o, k_o, c["gamma"], rows)
    assert np.array_equal(Rm_g, Rm_o) and np.array_equal(t_g, t_o)
    img = c["P"]["image"]
    for gs_mode in (False, True):
        gs_o, c3_o = oracle.back_project(img, dm_o, K4, Rm_o, t_o, gs_mode=gs_mode, want_coords=True)
        gs_g, c3_g = ctx.backproject(img, dm_o, K4, Rm_o, t_o, gs_mode=gs_mode, want_coords=True)
        assert np.array_equal(gs_g, gs_o)
        assert np.array_equal(c3_g, c3_o, equal_nan=True)
        assert np.array_equal(ctx.fill_cracks(gs_g, 1), oracle.interpolate_cracky_image(gs_o, 1))
    # row-major depth layout gives the same image
    dm_rm = dm_o.reshape(cols, rows).T.copy().reshape(-1)
    gs_rm, _ = ctx.backproject(img, dm_rm, K4, Rm_o, t_o, layout=1)
    gs_o, _ = oracle.back_project(img, dm_o, K4, Rm_o, t_o)
    assert np.array_equal(gs_rm, gs_o)


def test_fill_cracks_random_images(ctx, oracle):
    rng = np.random.default_rng(11)
    for shape in ((3, 3), (17, 31), (64, 50)):
        img = rng.integers(0, 40, size=shape + (3,), dtype=np.uint8)      # many pixels near the blackness threshold
        for off in (1, 2):
            if min(shape) <= 2 * off:
                continue
            assert np.array_equal(ctx.fill_cracks(img, off), oracle.interpolate_cracky_image(img, off))


# ---------------------------------------------------------------------------- fused driver
@pytest.mark.parametrize("which", ["cv", "ca"])
def test_refine_rectify_pair(ctx, oracle, case_cv, case_ca, which):
    c = case_cv if which == "cv" else case_ca
    R = c["ransac"]
    args = (c["flow"], c["inliers3"], c["alpha_in"], c["alpha_k_in"], c["m"], R["v"], R["w"], R["k"], c["const_acc"], False,
            c["P"]["image"], c["K4"], c["gamma"])
    ref = oracle.refine_rectify(*args)
    got = ctx.refine_rectify(*args)
    _motion_close(got["v"], ref["v"], "v")
    _motion_close(got["w"], ref["w"], "w")
    _depth_close(got["z"], ref["z"])
    nz = ref["depth_map"] != 0
    assert np.array_equal(got["depth_map"] != 0, nz)
    _depth_close(got["depth_map"][nz], ref["depth_map"][nz])
    diff = np.abs(got["rectified"].astype(np.int32) - ref["rectified"].astype(np.int32)).max(axis=2)
    assert (diff <= 1).mean() >= 0.999, "rectified image: %.5f of pixels within 1 grey level" % (diff <= 1).mean()


# ---------------------------------------------------------------------------- C++ host shim (reference class surface)
def test_cpp_host_shim_single_run(capi):
    """Builds rs-aware-differential-sfm_b200/host/example_single_run.cc (the glue of main.cc:398-523
    written against Camera / minimal::ransac / nonLinearRefinement / interpolateCrackyImage) with g++,
    links it with librsdsfm.so and runs it: exact constant-velocity data must give back w and the
    direction of v."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    host = os.path.join(root, "rs-aware-differential-sfm_b200", "host")
    exe = os.path.join(host, "example_single_run")
    cmd = ["g++", "-std=c++17", "-O2", "-I", os.path.join(root, "include"), "-I", host, os.path.join(host, "example_single_run.cc"),
           "-L", os.path.dirname(capi.LIB_PATH), "-lrsdsfm", "-Wl,-rpath," + os.path.dirname(capi.LIB_PATH), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=dict(os.environ, RSDSFM_RANSAC_SEED="7"))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rectified image" in r.stdout


def test_cpp_sweep_driver_evaluate_velocities(capi, tmp_path):
    """host/example_sweep.cc: the body of the reference's sweep driver (main.cc:245, :271-287) -- TrueValues(w, v),
    error_measure::evaluateVelocities with the reference's argument list, VelocityErrors read member by member
    -- compiled against the shim and run on the GPU: motion recovered on exact data, reprojection error
    wired (errorMeasure.cpp:229), depth PNGs and point clouds written per evaluation."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    host = os.path.join(root, "rs-aware-differential-sfm_b200", "host")
    exe = os.path.join(host, "example_sweep")
    cmd = ["g++", "-std=c++17", "-O2", "-I", os.path.join(root, "include"), "-I", host, os.path.join(host, "example_sweep.cc"),
           "-L", os.path.dirname(capi.LIB_PATH), "-lrsdsfm", "-Wl,-rpath," + os.path.dirname(capi.LIB_PATH), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=300, env=dict(os.environ, RSDSFM_RANSAC_SEED="7"))
    assert r.returncode == 0, r.stdout + r.stderr
    errs = open(os.path.join(str(tmp_path), "errors.csv")).read().strip().split(",")
    assert errs[0] == "synthetic_pair" and len(errs) == 4 and all(np.isfinite(float(x)) for x in errs[1:])
    # two evaluations, three values each: "a b c,a b c,"
    assert len(open(os.path.join(str(tmp_path), "w.csv")).read().strip().strip(",").split(",")) == 2


# ---------------------------------------------------------------------------- sequence / whole-pipeline drivers
def _seq_pairs(oracle, synth, n_pairs, const_acc):
    cases = []
    for p in range(n_pairs):
        rows, cols = (120, 160) if p % 2 == 0 else (120, 160)
        c = helpers.make_case(oracle, synth, rows, cols, helpers.small_K(8), k=0.5 if const_acc else 0.0, const_acc=const_acc,
                              H=8, seed=20 + p, sample_seed=40 + p, outliers=0.05 + 0.03 * p)
        cases.append(c)
    return cases


@pytest.mark.parametrize("mem", ["host", "device"])
@pytest.mark.parametrize("const_acc", [False, True])
def test_refine_rectify_sequence_equals_single_calls(ctx, oracle, synth, mem, const_acc):
    """The software-pipelined sequence entry point against n single calls (different m per pair, 5 pairs so
    both I/O slots / compute lanes are reused)."""
    import torch
    cases = _seq_pairs(oracle, synth, 5, const_acc)
    dev = torch.device("cuda", 0)
    to = (lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)) if mem == "device" else (lambda a: np.ascontiguousarray(a))
    pairs, single = [], []
    for c in cases:
        R = c["ransac"]
        d = dict(flow=to(c["flow"][:2 * c["m"]]), inliers3=to(c["inliers3"]), alpha=to(c["alpha_in"]), alpha_k=to(c["alpha_k_in"]),
                 image=to(c["P"]["image"]), m=c["m"], v=R["v"], w=R["w"], k=R["k"])
        pairs.append(d)
        single.append(ctx.refine_rectify(d["flow"], d["inliers3"], d["alpha"], d["alpha_k"], d["m"], d["v"], d["w"], d["k"], const_acc,
                                         False, d["image"], c["K4"], c["gamma"]))
    seq = ctx.refine_rectify_sequence(pairs, const_acc, False, cases[0]["K4"], cases[0]["gamma"])
    assert len(seq) == len(single)
    host = (lambda a: a.cpu().numpy()) if mem == "device" else (lambda a: a)
    # a sequence runs several pairs at once, each LM solve on a fraction of the SMs; the solver's sums are taken
    # over fixed strips of residual blocks whatever the grid (lm_kernel.cuh kStrips): identical to the bit
    for s, r in zip(seq, single):
        assert s["status"] == 0
        assert np.array_equal(s["v"], r["v"]) and np.array_equal(s["w"], r["w"]) and s["k"] == r["k"]
        assert s["summary"]["iterations"] == r["summary"]["iterations"]
        assert s["summary"]["final_cost"] == r["summary"]["final_cost"]
        for key in ("z", "depth_map", "rectified"):
            assert np.array_equal(host(s[key]), host(r[key])), key
    # ... the sequence is reproducible, and a one-pair sequence IS the single call
    again = ctx.refine_rectify_sequence(pairs, const_acc, False, cases[0]["K4"], cases[0]["gamma"])
    for s, t in zip(seq, again):
        assert np.array_equal(s["v"], t["v"]) and np.array_equal(s["w"], t["w"]) and s["k"] == t["k"]
        for key in ("z", "depth_map", "rectified"):
            assert np.array_equal(host(s[key]), host(t[key])), key
    one = ctx.refine_rectify_sequence(pairs[:1], const_acc, False, cases[0]["K4"], cases[0]["gamma"])[0]
    assert np.array_equal(one["v"], single[0]["v"]) and np.array_equal(host(one["z"]), host(single[0]["z"]))
    assert np.array_equal(host(one["rectified"]), host(single[0]["rectified"]))


@pytest.mark.parametrize("mem", ["host", "device"])
def test_long_sequence_on_four_lanes_equals_single_calls(ctx, oracle, synth, mem):
    """14 pairs: the sequence driver runs four LM solves side by side (37 CTAs each, here 2-3 strips per CTA) on
    10 / 14 lanes and hands every pair to the lane that is free first -- each pair bit-identical to its single call."""
    import torch
    cases = _seq_pairs(oracle, synth, 7, True)
    dev = torch.device("cuda", 0)
    to = (lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)) if mem == "device" else (lambda a: np.ascontiguousarray(a))
    host = (lambda a: a.cpu().numpy()) if mem == "device" else (lambda a: a)
    pairs, single = [], []
    for c in cases + cases:
        R = c["ransac"]
        d = dict(flow=to(c["flow"][:2 * c["m"]]), inliers3=to(c["inliers3"]), alpha=to(c["alpha_in"]), alpha_k=to(c["alpha_k_in"]),
                 image=to(c["P"]["image"]), m=c["m"], v=R["v"], w=R["w"], k=R["k"])
        pairs.append(d)
    for d, c in zip(pairs[:7], cases):
        single.append(ctx.refine_rectify(d["flow"], d["inliers3"], d["alpha"], d["alpha_k"], d["m"], d["v"], d["w"], d["k"], True,
                                         False, d["image"], c["K4"], c["gamma"]))
    seq = ctx.refine_rectify_sequence(pairs, True, False, cases[0]["K4"], cases[0]["gamma"])
    assert len(seq) == 14
    for i, s_ in enumerate(seq):
        r = single[i % 7]
        assert s_["status"] == 0 and s_["summary"]["iterations"] == r["summary"]["iterations"]
        assert np.array_equal(s_["v"], r["v"]) and np.array_equal(s_["w"], r["w"]) and s_["k"] == r["k"]
        for key in ("z", "depth_map", "rectified"):
            assert np.array_equal(host(s_[key]), host(r[key])), (i, key)


def test_lm_solve_does_not_depend_on_the_grid(capi):
    """The persistent LM kernel on 148, 74, 37 and 5 CTAs (RSDSFM_LM_GRID, read once per process: one process each):
    the refined motion, the cost and the depths come out identical to the bit."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for grid in ("148", "74", "37", "5"):
        r = subprocess.run([sys.executable, os.path.join(root, "tools", "lm_bench.py"), "--reps", "1", "--rows", "270", "--cols", "480"],
                           capture_output=True, text=True, timeout=300, env=dict(os.environ, RSDSFM_LM_GRID=grid))
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append(json.loads(r.stdout.strip().splitlines()[-1]))
    for o in outs[1:]:
        for key in ("iterations", "termination", "reason", "final_cost", "v", "w", "k", "z_sum"):
            assert o[key] == outs[0][key], (key, o[key], outs[0][key])
    assert outs[0]["iterations"] >= 5


def test_refine_rectify_sequence_bad_pair_is_reported(ctx, oracle, synth):
    cases = _seq_pairs(oracle, synth, 3, False)
    pairs = []
    for c in cases:
        R = c["ransac"]
        pairs.append(dict(flow=c["flow"][:2 * c["m"]], inliers3=c["inliers3"], alpha=c["alpha_in"], alpha_k=c["alpha_k_in"],
                          image=c["P"]["image"], m=c["m"], v=R["v"], w=R["w"], k=R["k"]))
    pairs[1]["m"] = 0
    capi = __import__("importlib").import_module("rs-aware-differential-sfm_b200.capi")
    with pytest.raises(capi.RsdsfmError):
        ctx.refine_rectify_sequence(pairs, False, False, cases[0]["K4"], cases[0]["gamma"])
    # the context stays usable and the valid pairs still run
    ok = ctx.refine_rectify_sequence([pairs[0], pairs[2]], False, False, cases[0]["K4"], cases[0]["gamma"])
    assert [r["status"] for r in ok] == [0, 0]
    # the same with device buffers
    import torch
    dpairs = [{k: (torch.from_numpy(np.ascontiguousarray(v)).cuda() if isinstance(v, np.ndarray) and v.size > 7 else v) for k, v in p.items()}
              for p in pairs]
    with pytest.raises(capi.RsdsfmError):
        ctx.refine_rectify_sequence(dpairs, False, False, cases[0]["K4"], cases[0]["gamma"])
    okd = ctx.refine_rectify_sequence([dpairs[0], dpairs[2], dpairs[0]], False, False, cases[0]["K4"], cases[0]["gamma"])
    assert [r["status"] for r in okd] == [0, 0, 0]
    assert np.array_equal(okd[0]["rectified"].cpu().numpy(), okd[2]["rectified"].cpu().numpy())


@pytest.mark.parametrize("mem", ["host", "device"])
@pytest.mark.parametrize("which,repair", [("cv", False), ("ca", False), ("ca", True)])
def test_pipeline_pair_equals_stagewise_calls(ctx, oracle, case_cv, case_ca, which, repair, mem):
    """rsdsfm_pipeline_pair (flatten .. crack fill in one call) against the oracle's RANSAC winner and
    the per-stage entry points on the same sample list."""
    import torch
    c = case_cv if which == "cv" else case_ca
    dev = torch.device("cuda", 0)
    to = (lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)) if mem == "device" else (lambda a: np.ascontiguousarray(a))
    host = (lambda a: a.cpu().numpy()) if mem == "device" else (lambda a: a)
    got = ctx.pipeline_pair(to(c["P"]["flow_img"]), to(c["P"]["image"]), c["K4"], c["gamma"], c["tol"], c["const_acc"],
                            samples=c["samples"], repair_pairing=repair)
    R = c["ransac"]
    assert got["status"] == 0 and got["n"] == c["n"] and got["m"] == c["m"] and got["best_idx"] == R["best_idx"]
    assert np.array_equal(got["ransac_v"], R["v"]) and np.array_equal(got["ransac_w"], R["w"]) and got["ransac_k"] == R["k"]
    # stage-wise on the device: same kernels, so identical results
    fi = torch.from_numpy(c["P"]["flow_img"]).to(dev)
    n, coord, flow, cpx, fpx, pidx = ctx.flatten(fi, c["K4"], c["gamma"])
    alpha, alpha_k = ctx.alpha(fpx[:2 * n], cpx[:2 * n], n, c["rows"], c["gamma"])
    Rg = ctx.ransac(coord[:2 * n], flow[:2 * n], alpha, alpha_k, n, c["const_acc"], c["samples"], c["tol"])
    inl, a_in, ak_in, ix, m = ctx.gather_inliers(coord[:2 * n], alpha, alpha_k, n, Rg["mask"], Rg["inv_depth"])
    if repair:
        ref_r = ctx.refine(flow[:2 * n].contiguous(), inl.contiguous(), a_in.contiguous(), ak_in.contiguous(), m, Rg["v"], Rg["w"], Rg["k"],
                           c["const_acc"], flow_index=ix[:m].contiguous())
        _motion_close(got["w"], ref_r[1], "w")
        assert abs(got["k"] - ref_r[2]) <= 1e-9 * max(1.0, abs(ref_r[2]))
    else:
        ref = ctx.refine_rectify(flow[:2 * n].contiguous(), inl.contiguous(), a_in.contiguous(), ak_in.contiguous(), m, Rg["v"], Rg["w"],
                                 Rg["k"], c["const_acc"], False, torch.from_numpy(c["P"]["image"]).to(dev), c["K4"], c["gamma"])
        assert np.array_equal(got["v"], ref["v"]) and np.array_equal(got["w"], ref["w"]) and got["k"] == ref["k"]
        assert np.array_equal(host(got["depth_map"]), ref["depth_map"].cpu().numpy())
        assert np.array_equal(host(got["rectified"]), ref["rectified"].cpu().numpy())


def test_pipeline_emulate_padding_follows_main_cc(ctx, oracle, synth):
    """main.cc:398-447 never truncates the flattened arrays: the rows*cols - n tail columns
    (coord = (1,1), flow = 0, alpha = 1) are sampled, scored and can become inliers.  With emulate_padding
    the one-call driver works on that point set; compared with the oracle stages run on the padded arrays."""
    rows, cols = 96, 128
    K4 = helpers.small_K(12)
    gamma = 0.95
    P = synth.make_pair(rows, cols, K4, gamma=gamma, seed=31, k=0.0, noise_sigma_px=0.1, outlier_frac=0.05, zero_flow_frac=0.3)
    tot = rows * cols
    n, coord, flow, cpx, fpx = oracle.flatten(P["flow_img"], K4, gamma)
    assert n < tot
    alpha = oracle.get_alpha(fpx, tot, rows, gamma)
    alpha_k = oracle.get_alpha_k(cpx, fpx, tot, rows, gamma)
    samples = synth.sample_list(tot, 8, seed=5)
    samples[3, :4] = [tot - 1, tot - 2, n, n + 1]            # a trial that draws phantom points
    R = oracle.ransac(coord, flow, alpha, alpha_k, tot, False, 0.01, samples=samples)
    inl, a_in, ak_in = oracle.gather_inliers(coord, alpha, alpha_k, tot, R["mask"], R["inv_depth"])
    m = len(a_in)
    ref = oracle.refine_rectify(flow, inl, a_in, ak_in, m, R["v"], R["w"], R["k"], False, False, P["image"], K4, gamma)
    got = ctx.pipeline_pair(P["flow_img"], P["image"], K4, gamma, 0.01, False, samples=samples, emulate_padding=True)
    assert got["n"] == n and got["m"] == m and got["best_idx"] == R["best_idx"]
    assert np.array_equal(got["ransac_v"], R["v"]) and np.array_equal(got["ransac_w"], R["w"])
    _motion_close(got["w"], ref["w"], "w")
    _motion_close(got["v"], ref["v"], "v")
    assert got["summary"]["iterations"] == ref["summary"]["iterations"]
    diff = np.abs(got["rectified"].astype(np.int32) - ref["rectified"].astype(np.int32)).max(axis=2)
    assert (diff <= 1).mean() >= 0.999
    # and without the flag the point set is the n kept vectors
    plain = ctx.pipeline_pair(P["flow_img"], P["image"], K4, gamma, 0.01, False, samples=synth.sample_list(n, 8, seed=5))
    assert plain["n"] == n and plain["m"] <= n


def test_pipeline_no_refinement_and_gs_mode(ctx, oracle, case_cv):
    """use_refinement = 0 rectifies with the RANSAC winner (main.cc:455); gs_mode sets alpha = 1."""
    c = case_cv
    R = c["ransac"]
    got = ctx.pipeline_pair(c["P"]["flow_img"], c["P"]["image"], c["K4"], c["gamma"], c["tol"], False, samples=c["samples"],
                            use_refinement=False)
    assert got["m"] == c["m"]
    inl = c["inliers3"].copy()
    _, v_fix, dm, _, _ = oracle.depth_glue(inl, c["m"], R["v"].copy(), c["K4"], c["rows"], c["cols"])
    assert np.array_equal(got["depth_map"], dm)
    assert np.array_equal(got["v"], v_fix) and np.array_equal(got["w"], R["w"])
    # global-shutter mode: alpha = 1 everywhere -> same consensus as the oracle run with alpha = 1
    n = c["n"]
    ones = np.ones(n)
    Rgs = oracle.ransac(c["coord"], c["flow"], ones, c["alpha_k"], n, False, c["tol"], samples=c["samples"])
    got = ctx.pipeline_pair(c["P"]["flow_img"], c["P"]["image"], c["K4"], c["gamma"], c["tol"], False, samples=c["samples"],
                            gs_mode=True, use_refinement=False)
    assert got["best_idx"] == Rgs["best_idx"] and got["m"] == int(Rgs["mask"].sum())


def test_pipeline_sequence_and_reference_draws(ctx, oracle, synth):
    """Sequence driver == pair driver per entry; `draws` are mapped to samples like minimal.cc:226-244."""
    cases = _seq_pairs(oracle, synth, 3, False)
    rng = np.random.RandomState(7)
    pairs = []
    for c in cases:
        draws = rng.randint(0, 2 ** 31 - 1, size=(c["samples"].shape[0], 9)).astype(np.uint32)
        pairs.append(dict(flow_img=c["P"]["flow_img"], image=c["P"]["image"], draws=draws))
    seq = ctx.pipeline_sequence(pairs, cases[0]["K4"], cases[0]["gamma"], cases[0]["tol"], False)
    for c, p, s in zip(cases, pairs, seq):
        # the reference's draw: persistent index vector, swap with the last live entry
        idx = np.arange(c["n"])
        samples = []
        for t in range(p["draws"].shape[0]):
            nt = c["n"]
            for j in range(9):
                r = int(p["draws"][t, j]) % nt
                idx[nt - 1], idx[r] = idx[r], idx[nt - 1]
                samples.append(idx[nt - 1])
                nt -= 1
        samples = np.array(samples, dtype=np.int32).reshape(-1, 9)
        one = ctx.pipeline_pair(c["P"]["flow_img"], c["P"]["image"], c["K4"], c["gamma"], c["tol"], False, samples=samples)
        assert s["status"] == 0 and s["n"] == one["n"] and s["m"] == one["m"] and s["best_idx"] == one["best_idx"]
        assert np.array_equal(s["v"], one["v"]) and np.array_equal(s["w"], one["w"])
        assert np.array_equal(s["depth_map"], one["depth_map"]) and np.array_equal(s["rectified"], one["rectified"])


@pytest.mark.parametrize("const_acc", [False, True])
def test_refine_is_bit_reproducible(ctx, oracle, synth, const_acc):
    """Fixed-order reductions: repeated solves of the same problem agree to the last bit, including
    problems with listed (clamped-diagonal) pixels whose list slots are handed out by atomics."""
    for p in (3, 0):
        c = helpers.make_case(oracle, synth, 120, 160, helpers.small_K(8), k=0.5 if const_acc else 0.0, const_acc=const_acc,
                              H=8, seed=20 + p, sample_seed=40 + p, outliers=0.05 + 0.03 * p)
        R = c["ransac"]
        runs = [ctx.refine(c["flow"][:2 * c["m"]], c["inliers3"], c["alpha_in"], c["alpha_k_in"], c["m"], R["v"], R["w"], R["k"], const_acc)
                for _ in range(6)]
        for r in runs[1:]:
            assert np.array_equal(r[0], runs[0][0]) and np.array_equal(r[1], runs[0][1]) and r[2] == runs[0][2]
            assert np.array_equal(r[3], runs[0][3])
            assert r[4]["final_cost"] == runs[0][4]["final_cost"] and r[4]["iterations"] == runs[0][4]["iterations"]


# ---------------------------------------------------------------------------- reprojection error metric
def _reproj_case(oracle, synth, rows=120, cols=160):
    K4 = helpers.small_K(8)
    P = synth.make_pair(rows, cols, K4, gamma=0.95, seed=31, k=0.3)
    a = 0.02
    Rg = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])
    G = synth.ground_truth(P, world_R=Rg, world_t=(0.1, -0.05, 0.2), void_frac=0.01)
    depth_cm = (G["depth"] * 1.3).flatten(order="F")                      # estimate with the usual scale ambiguity
    R, t = oracle.set_relative_pose(P["v"] * 1.3, P["w"], P["k"], P["gamma"], rows)
    _, coords = oracle.back_project(P["image"], depth_cm, K4, R, t, want_coords=True)
    rng = np.random.default_rng(5)
    bad = rng.random((rows, cols)) < 0.01                                 # gross outliers and non-finite estimates
    coords[bad] *= -40.0
    coords[rng.random((rows, cols)) < 0.002] = np.nan
    return P, G, K4, depth_cm, coords


@pytest.mark.parametrize("mem", ["host", "device"])
def test_reprojection_error_matches_oracle(ctx, oracle, synth, capi, mem):
    """Camera::meanReprojectionError / createErrorImage / getGroundtruthDepthMap / relocatePose."""
    import torch
    P, G, K4, depth_cm, coords = _reproj_case(oracle, synth)
    ref = oracle.mean_reprojection_error(coords, *G["unproj"], G["R_gt"], G["t_gt"], depth_cm, K4, max_norm=2.0, want_image=True)
    unproj = [u.flatten(order="F") for u in G["unproj"]]
    to = (lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()) if mem == "device" else (lambda a: a)
    host = (lambda a: a.cpu().numpy()) if mem == "device" else (lambda a: a)
    got = ctx.reprojection_error(to(coords), [to(u) for u in unproj], G["R_gt"], G["t_gt"], to(depth_cm), K4, max_norm=2.0,
                                 want_image=True, want_gt_depth=True)
    assert got["num_outliers"] == ref["num_outliers"] and got["points_used"] == ref["points_used"]
    assert abs(got["mean_scale"] - ref["mean_scale"]) <= 1e-12 * abs(ref["mean_scale"])
    assert abs(got["mean_error"] - ref["mean_error"]) <= 1e-10 * abs(ref["mean_error"])
    # 8-bit error image: identical up to pixels sitting on a rounding boundary of the (summation-order dependent) mean scale
    assert (host(got["error_image"]) != ref["error_image"]).mean() < 1e-3
    gd = oracle.groundtruth_depth_map(*G["unproj"], G["R_gt"], G["t_gt"])
    assert np.array_equal(host(got["gt_depth_map"]).reshape(gd.shape[1], gd.shape[0]).T, gd)
    Rr, tr = capi.relocate_pose(G["R_gt"], G["t_gt"])
    Ro, to_ = oracle.relocate_pose(G["R_gt"], G["t_gt"])
    assert np.array_equal(Rr, Ro) and np.array_equal(tr, to_)


def test_reprojection_error_of_the_pipeline_output(ctx, oracle, synth):
    """End of the reference's evaluation loop: refine + rectify, then the reprojection error of the
    back-projected 3D points against the ground truth -- small for a noise-free pair."""
    K4 = helpers.small_K(8)
    c = helpers.make_case(oracle, synth, 120, 160, K4, k=0.0, const_acc=False, H=8, seed=33, noise=0.0, outliers=0.0)
    R = c["ransac"]
    out = ctx.refine_rectify(c["flow"], c["inliers3"], c["alpha_in"], c["alpha_k_in"], c["m"], R["v"], R["w"], R["k"], False, False,
                             c["P"]["image"], c["K4"], c["gamma"])
    Rrel, trel = ctx.set_relative_pose(out["v"], out["w"], out["k"], c["gamma"], c["rows"])
    _, coords = ctx.backproject(c["P"]["image"], out["depth_map"], c["K4"], Rrel, trel, want_coords=True)
    G = synth.ground_truth(c["P"])
    got = ctx.reprojection_error(coords, [u.flatten(order="F") for u in G["unproj"]], G["R_gt"], G["t_gt"], out["depth_map"], c["K4"])
    mean_depth = float(np.mean(G["depth"]))
    assert got["points_used"] > 0.95 * c["rows"] * c["cols"]
    assert got["mean_error"] < 0.02 * mean_depth, got


# ---------------------------------------------------------------------------- ground-truth flow
@pytest.mark.parametrize("mem", ["host", "device"])
def test_true_flow_bit_exact(ctx, oracle, synth, mem):
    """Camera::calculateTrueFlow: every scanline pose of frame 2 tried per pixel, first minimum wins."""
    import torch
    rows, cols = 96, 128
    K4 = helpers.small_K(12)
    P = synth.make_pair(rows, cols, K4, gamma=0.95, seed=41, k=0.4)
    a = -0.015
    Rg = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    G1 = synth.ground_truth(P, world_R=Rg, world_t=(0.05, 0.02, -0.1), void_frac=0.02)
    G2 = synth.ground_truth(P, world_R=Rg, world_t=(0.05, 0.02, -0.1), frame=2)
    ref = oracle.true_flow(*G1["unproj"], G2["R_gt"], G2["t_gt"], K4)
    unproj = [u.flatten(order="F") for u in G1["unproj"]]
    if mem == "device":
        unproj = [torch.from_numpy(u).cuda() for u in unproj]
    got = ctx.true_flow(unproj, G2["R_gt"], G2["t_gt"], K4, rows, cols)
    got = got.cpu().numpy() if mem == "device" else got
    assert np.array_equal(got, ref)
    void = (G1["unproj"][0] == 0) & (G1["unproj"][1] == 0) & (G1["unproj"][2] == 0)
    assert void.any() and np.all(got[void] == 0)
    # sanity: close to the analytic differential flow the pair was generated with (first-order model)
    solid = ~void
    err = np.abs(got[solid] - P["flow_img"][solid])
    assert np.median(err) < 0.15 * np.median(np.abs(P["flow_img"][solid])) + 0.05


@pytest.mark.parametrize("rows,cols", [(77, 123), (120, 160)])
def test_fused_driver_equals_stagewise_rectification(ctx, oracle, synth, rows, cols):
    """rsdsfm_refine_rectify takes shortcuts the per-stage entry points do not (depth statistics from
    the solve's epilogue, gather + crack fill through a shared-memory tile): the rectified frame must
    be bit-identical to setPose -> backProject -> interpolateCrackyImage run stage by stage, also
    for image sizes that are not multiples of the tile."""
    c = helpers.make_case(oracle, synth, rows, cols, helpers.small_K(10), k=0.0, const_acc=False, H=8, seed=51, outliers=0.1)
    R = c["ransac"]
    got = ctx.refine_rectify(c["flow"], c["inliers3"], c["alpha_in"], c["alpha_k_in"], c["m"], R["v"], R["w"], R["k"], False, False,
                             c["P"]["image"], c["K4"], c["gamma"])
    Rrel, trel = ctx.set_relative_pose(got["v"], got["w"], got["k"], c["gamma"], rows)
    gs, _ = ctx.backproject(c["P"]["image"], got["depth_map"], c["K4"], Rrel, trel)
    assert np.array_equal(ctx.fill_cracks(gs, 1), got["rectified"])
    # and the depth raster equals the stage-wise glue on the refined depths
    inl = c["inliers3"].copy(); inl[2::3] = got["z"] if inl.ndim == 1 else inl[2::3]
    ref = oracle.refine_rectify(c["flow"], c["inliers3"], c["alpha_in"], c["alpha_k_in"], c["m"], R["v"], R["w"], R["k"], False, False,
                                c["P"]["image"], c["K4"], c["gamma"])
    assert np.array_equal(got["depth_map"] != 0, ref["depth_map"] != 0)


# ---------------------------------------------------------------------------- BASELINE.json configs 1 and 3 as parity cases
def test_config1_castle_substitute_full_pipeline(ctx, oracle, synth):
    """BASELINE config 1 (SURVEY 8d substitute for the absent castle tarball): 600 x 600, fx = fy = 800,
    cx = cy = 300, gamma = 0.8, 35 % background with flow exactly 0 (dropped by the 1e-10 test), 5 RANSAC
    trials, tol 0.05, constant-velocity refinement, rectification -- the whole a2..a15 chain in one call
    against the CPU restatement."""
    rows = cols = 600
    K4 = np.array([800.0, 800.0, 300.0, 300.0])
    v = 0.03 * 100.0 * np.array([1.0, 1.0, 0.0]) / np.sqrt(2.0)
    P = synth.make_pair(rows, cols, tuple(K4), gamma=0.8, v=tuple(v), w=(0.0, 0.0, 8.7266e-3), k=0.0, seed=61,
                        noise_sigma_px=0.05, outlier_frac=0.02, zero_flow_frac=0.35, z_range=(60.0, 140.0))
    n, coord, flow, cpx, fpx = oracle.flatten(P["flow_img"], K4, 0.8)
    assert 0.5 * rows * cols < n < 0.9 * rows * cols          # the background blobs are really dropped
    alpha = oracle.get_alpha(fpx, n, rows, 0.8)
    alpha_k = oracle.get_alpha_k(cpx, fpx, n, rows, 0.8)
    samples = synth.sample_list(n, 5, seed=62)
    R = oracle.ransac(coord[:2 * n], flow[:2 * n], alpha, alpha_k, n, False, 0.05, samples=samples)
    inl, a_in, ak_in = oracle.gather_inliers(coord, alpha, alpha_k, n, R["mask"], R["inv_depth"])
    m = len(a_in)
    ref = oracle.refine_rectify(flow[:2 * n], inl, a_in, ak_in, m, R["v"], R["w"], R["k"], False, False, P["image"], K4, 0.8)
    got = ctx.pipeline_pair(P["flow_img"], P["image"], K4, 0.8, 0.05, False, samples=samples)
    assert got["n"] == n and got["m"] == m and got["best_idx"] == R["best_idx"]          # RANSAC: bit-exact
    assert np.array_equal(got["ransac_v"], R["v"]) and np.array_equal(got["ransac_w"], R["w"])
    _motion_close(got["v"], ref["v"], "v")
    _motion_close(got["w"], ref["w"], "w")
    assert got["summary"]["iterations"] == ref["summary"]["iterations"]
    nz = ref["depth_map"] != 0
    assert np.array_equal(got["depth_map"] != 0, nz)
    _depth_close(got["depth_map"][nz], ref["depth_map"][nz])
    diff = np.abs(got["rectified"].astype(np.int32) - ref["rectified"].astype(np.int32)).max(axis=2)
    assert (diff <= 1).mean() >= 0.999


def test_config3_realworld_substitute(ctx, oracle, synth):
    """BASELINE config 3 substitute: `galaxy` intrinsics, flow quantised to float32 like
    DeepFlow's output, a smooth low-frequency flow error and outlier blobs.  Inlier sets must be
    bit-exact for the same sample list; the refinement runs with the reference's flow pairing (m < n)."""
    rows, cols = 270, 480
    K4 = np.array(synth.INTRINSICS["galaxy"]) / 4.0
    P = synth.make_pair(rows, cols, tuple(K4), gamma=0.95, seed=71, k=0.0, noise_sigma_px=0.0, outlier_frac=0.0, flow_f32=False)
    yy, xx = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    smooth = 0.5 * np.stack([np.sin(2 * np.pi * xx / 64.0 + 0.3) * np.cos(2 * np.pi * yy / 64.0),
                             np.cos(2 * np.pi * xx / 64.0) * np.sin(2 * np.pi * yy / 64.0 - 0.2)], axis=-1)
    flow_img = P["flow_img"] + smooth
    blobs = (np.sin(xx * 0.11 + 1.0) * np.cos(yy * 0.13 - 0.5)) > 0.97                 # ~3 % outlier blobs
    rng = np.random.default_rng(72)
    flow_img[blobs] = rng.uniform(-15, 15, size=(int(blobs.sum()), 2))
    flow_img = flow_img.astype(np.float32).astype(np.float64)
    n, coord, flow, cpx, fpx = oracle.flatten(flow_img, K4, 0.95)
    alpha = oracle.get_alpha(fpx, n, rows, 0.95)
    alpha_k = oracle.get_alpha_k(cpx, fpx, n, rows, 0.95)
    samples = synth.sample_list(n, 12, seed=73)
    R = oracle.ransac(coord[:2 * n], flow[:2 * n], alpha, alpha_k, n, False, 0.002, samples=samples)
    G = ctx.ransac(coord[:2 * n], flow[:2 * n], alpha, alpha_k, n, False, samples, 0.002)
    assert np.array_equal(G["counts"], R["counts"]) and G["best_idx"] == R["best_idx"]
    assert np.array_equal(G["mask"], R["mask"]) and np.array_equal(G["inv_depth"], R["inv_depth"])
    m = int(R["mask"].sum())
    assert 0 < m < n                                                                   # Q1 pairing is active
    inl, a_in, ak_in = oracle.gather_inliers(coord, alpha, alpha_k, n, R["mask"], R["inv_depth"])
    ref = oracle.nonlinear_refinement(flow[:2 * n], inl, a_in, ak_in, m, R["v"], R["w"], R["k"], False)
    got = ctx.refine(flow[:2 * m], inl, a_in, ak_in, m, R["v"], R["w"], R["k"], False)
    _motion_close(got[0], ref[0], "v")
    _motion_close(got[1], ref[1], "w")
    _depth_close(got[3], ref[3])


# ---------------------------------------------------------------------------- BASELINE config 2 at full size (1920 x 1080)
@pytest.fixture(scope="module")
def full_hd(ctx, synth):
    """The bench workload's pair, prepared on the GPU (flatten, alpha, RANSAC with 4 trials, gather)."""
    rows, cols = 1080, 1920
    P = synth.make_pair(rows, cols, "galaxy_stabil", gamma=0.95, seed=1000, k=0.0, noise_sigma_px=0.3, outlier_frac=0.05)
    n, coord, flow, cpx, fpx, pidx = ctx.flatten(P["flow_img"], P["K4"], P["gamma"])
    coord, flow, cpx, fpx = coord[:2 * n], flow[:2 * n], cpx[:2 * n], fpx[:2 * n]
    alpha, alpha_k = ctx.alpha(fpx, cpx, n, rows, P["gamma"])
    samples = synth.sample_list(n, 4, seed=1100)
    R = ctx.ransac(coord, flow, alpha, alpha_k, n, False, samples, 0.05)
    inl, a_in, ak_in, ix, m = ctx.gather_inliers(coord, alpha, alpha_k, n, R["mask"], R["inv_depth"])
    return dict(P=P, n=n, coord=coord, flow=flow, alpha=alpha, alpha_k=alpha_k, R=R, inl=inl[:3 * m], a_in=a_in[:m], ak_in=ak_in[:m], m=m,
                rows=rows, cols=cols)


def test_full_hd_refine_rectify_against_oracle_and_reproducible(ctx, oracle, full_hd):
    """One 1080p pair (2 073 600 residual blocks) against the CPU restatement (a few seconds of oracle
    time), plus bitwise reproducibility of a second run."""
    c = full_hd
    R, P = c["R"], c["P"]
    args = (c["flow"], c["inl"], c["a_in"], c["ak_in"], c["m"], R["v"], R["w"], R["k"], False, False, P["image"], P["K4"], P["gamma"])
    got = ctx.refine_rectify(*args)
    again = ctx.refine_rectify(*args)
    for key in ("v", "w", "z", "depth_map", "rectified"):
        assert np.array_equal(got[key], again[key]), key
    ref = oracle.refine_rectify(*args)
    _motion_close(got["v"], ref["v"], "v")
    _motion_close(got["w"], ref["w"], "w")
    assert got["summary"]["iterations"] == ref["summary"]["iterations"]
    _depth_close(got["z"], ref["z"])
    diff = np.abs(got["rectified"].astype(np.int32) - ref["rectified"].astype(np.int32)).max(axis=2)
    assert (diff <= 1).mean() >= 0.999


def test_full_hd_depth_lm_equals_closed_form_and_scoring_is_consistent(ctx, full_hd):
    """Size-independent properties at 1080p: (i) the depth-only LM (a8) converges to the per-point
    closed-form least-squares depth; (ii) the consensus mask of the RANSAC winner is exactly the set
    err < tol recomputed in numpy from the winner's depths with the reference's operation order."""
    c = full_hd
    R = c["R"]
    n = c["n"]
    q = c["coord"].reshape(-1, 2); u = c["flow"].reshape(-1, 2)
    x, y = q[:, 0], q[:, 1]
    v, w, k = R["v"], R["w"], R["k"]
    beta = (c["alpha"][:n] + k * c["alpha_k"][:n]) * (2.0 / (2.0 + k))
    av0 = 1.0 * v[0] + 0.0 * v[1] + (-x) * v[2]
    av1 = 0.0 * v[0] + 1.0 * v[1] + (-y) * v[2]
    bw0 = (-x * y) * w[0] + (1 + x * x) * w[1] + (-y) * w[2]
    bw1 = (-(1 + y * y)) * w[0] + (x * y) * w[1] + x * w[2]
    # (ii) scoring loop, minimal.cc:255-275
    d = R["inv_depth"][:n]
    ue0 = beta * (av0 * d + bw0); ue1 = beta * (av1 * d + bw1)
    dx = ue0 - u[:, 0]; dy = ue1 - u[:, 1]
    err = np.sqrt(dx * dx + dy * dy)
    assert np.array_equal(err < 0.05, R["mask"][:n].astype(bool))
    assert int((err < 0.05).sum()) == int(R["counts"][R["best_idx"]])
    # (i) closed form: d* = e.(u - beta B w) / (e.e), e = beta A v
    e0, e1 = beta * av0, beta * av1
    dstar = (e0 * (u[:, 0] - beta * bw0) + e1 * (u[:, 1] - beta * bw1)) / (e0 * e0 + e1 * e1)
    got = ctx.estimate_inverse_depths(c["coord"], c["flow"], n, v, w, k, c["alpha"][:n], c["alpha_k"][:n])
    dd = got[0] if isinstance(got, tuple) else got["inv_depth"]
    rel = np.abs(dd[:n] - dstar) / np.maximum(np.abs(dstar), 1e-12)
    assert np.median(rel) < 1e-6 and np.percentile(rel, 99) < 1e-3


def test_full_hd_rectification_is_identity_for_a_static_camera(ctx, full_hd):
    """v = w = 0: every scanline pose is the identity, so backProject maps each pixel onto itself and
    the crack fill has nothing to do except at the source's own dark / void pixels."""
    c = full_hd
    P = c["P"]
    rows, cols = c["rows"], c["cols"]
    Rm, tm = ctx.set_relative_pose(np.zeros(3), np.zeros(3), 0.0, P["gamma"], rows)
    depth = np.full(rows * cols, 5.0)
    K4 = np.array([1800.0, 1800.0, 960.0, 540.0])          # fx == fy: spaceToPlane scales y with f_x (Q12)
    gs, _ = ctx.backproject(P["image"], depth, K4, Rm, tm)
    skipped = np.all(P["image"] == 1, axis=2)
    assert np.array_equal(gs[~skipped], P["image"][~skipped]) and np.all(gs[skipped] == 0)


@pytest.mark.parametrize("mem", ["host", "device"])
def test_clamped_pixel_list_overflow_is_retried(capi, ctx, oracle, synth, monkeypatch, mem):
    """More clamped pixels than the list holds: the kernel flags the overflow, the library enlarges the
    list and repeats the solve -- in the synchronous call and in the middle of a pipelined sequence.
    RSDSFM_EXC_CAP shrinks the initial capacity so that a handful of clamped pixels is enough."""
    foe = helpers.make_case(oracle, synth, 240, 320, (1500.0, 1500.0, 160.0, 120.0), k=0.4, const_acc=True, H=10, seed=9,
                            v=(0.02, 0.01, 0.30), w=(0.001, 0.002, -0.003), noise=0.05, outliers=0.0, tol=0.05)
    plain = helpers.make_case(oracle, synth, 240, 320, helpers.small_K(6), k=0.4, const_acc=True, H=8, seed=12)

    import torch
    to = (lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()) if mem == "device" else (lambda a: np.ascontiguousarray(a))
    host = (lambda a: a.cpu().numpy()) if mem == "device" else (lambda a: a)

    def pair(c):
        R = c["ransac"]
        return dict(flow=to(c["flow"][:2 * c["m"]]), inliers3=to(c["inliers3"]), alpha=to(c["alpha_in"]), alpha_k=to(c["alpha_k_in"]),
                    image=to(c["P"]["image"]), m=c["m"], v=R["v"], w=R["w"], k=R["k"])

    def run(context):
        R = foe["ransac"]
        single = context.refine(foe["flow"], foe["inliers3"], foe["alpha_in"], foe["alpha_k_in"], foe["m"], R["v"], R["w"], R["k"], True)
        # same K for the whole sequence (the FOE pair's): only the solve matters here
        seq = context.refine_rectify_sequence([pair(plain), pair(foe), pair(plain), pair(foe)], True, False, foe["K4"], foe["gamma"])
        return single, seq

    roomy = capi.Context(0)
    want_single, want_seq = run(roomy)
    launches_roomy = roomy.launch_count()
    roomy.close()
    monkeypatch.setenv("RSDSFM_EXC_CAP", "2")
    small = capi.Context(0)
    got_single, got_seq = run(small)
    launches_small = small.launch_count()
    small.close()
    assert launches_small > launches_roomy          # the overflowing solves really ran twice
    assert np.array_equal(got_single[0], want_single[0]) and np.array_equal(got_single[1], want_single[1]) and got_single[2] == want_single[2]
    assert np.array_equal(got_single[3], want_single[3]) and got_single[4]["iterations"] == want_single[4]["iterations"]
    for g, w in zip(got_seq, want_seq):
        assert g["status"] == 0
        assert np.array_equal(g["v"], w["v"]) and np.array_equal(g["w"], w["w"]) and g["k"] == w["k"]
        assert np.array_equal(host(g["z"]), host(w["z"])) and np.array_equal(host(g["rectified"]), host(w["rectified"]))


# ---------------------------------------------------------------------------- committed golden vectors
@pytest.mark.parametrize("tag,k,cacc,seed", [("cv", 0.0, False, 21), ("ca", 0.5, True, 22)])
def test_cuda_path_against_committed_golden_vectors(ctx, synth, tag, k, cacc, seed):
    """tests/golden/pipeline_small.npz (tests/golden/make_golden.py): the CUDA path alone, from the seeded
    flow image to the rectified frame, against the committed vectors -- no oracle call in this test."""
    import os
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "pipeline_small.npz"))
    rows, cols = 48, 64
    K4 = np.array([80.0, 79.0, 32.0, 24.0])
    P = synth.make_pair(rows, cols, tuple(K4), gamma=0.95, seed=seed, k=k, noise_sigma_px=0.1, outlier_frac=0.05)
    n, coord, flow, cpx, fpx, pidx = ctx.flatten(P["flow_img"], K4, 0.95)
    coord, flow, cpx, fpx = coord[:2 * n], flow[:2 * n], cpx[:2 * n], fpx[:2 * n]
    alpha, alpha_k = ctx.alpha(fpx, cpx, n, rows, 0.95)
    R = ctx.ransac_score(coord, flow, alpha, alpha_k, n, G[tag + "_hyps"], 0.01)
    assert np.array_equal(R["counts"], G[tag + "_counts"]) and R["best_idx"] == int(G[tag + "_best"][0])
    assert np.array_equal(R["mask"][:n], G[tag + "_mask"][:n]) and np.array_equal(R["inv_depth"][:n], G[tag + "_inv_depth"][:n])
    inl, a_in, ak_in, ix, m = ctx.gather_inliers(coord, alpha, alpha_k, n, R["mask"], R["inv_depth"])
    h = G[tag + "_hyps"][R["best_idx"]]                       # (w, v, k) of the winner
    got = ctx.refine_rectify(flow, inl[:3 * m], a_in[:m], ak_in[:m], m, h[3:6], h[0:3], float(h[6]), cacc, False, P["image"], K4, 0.95)
    want = G[tag + "_motion"]
    _motion_close(got["v"], want[0:3], "v")
    _motion_close(got["w"], want[3:6], "w")
    assert abs(got["k"] - want[6]) <= MOTION_RTOL * max(1.0, abs(want[6]))
    assert got["summary"]["iterations"] == int(G[tag + "_iterations"][0])
    _depth_close(got["z"], G[tag + "_z"])
    diff = np.abs(got["rectified"].astype(np.int32) - G[tag + "_rectified"].astype(np.int32)).max(axis=2)
    assert (diff <= 1).mean() >= 0.999


def test_negative_mean_depth_is_sign_fixed_in_the_fused_driver(ctx, oracle, case_cv):
    """The 9-point solver's null-vector sign is arbitrary, so (v, z) may come out negated; main.cc:466-478
    flips both when the mean depth is negative.  The fused driver takes the depth statistics from the
    solve's epilogue: force the negative branch and compare with the oracle and the stage-wise glue."""
    c = case_cv
    R = c["ransac"]
    inl = c["inliers3"].copy()
    inl[2::3] *= -1.0
    args = (c["flow"], inl, c["alpha_in"], c["alpha_k_in"], c["m"], -R["v"], R["w"], R["k"], False, False, c["P"]["image"], c["K4"], c["gamma"])
    ref = oracle.refine_rectify(*args)
    got = ctx.refine_rectify(*args)
    assert np.mean(got["z"]) > 0 and np.mean(ref["z"]) > 0
    _motion_close(got["v"], ref["v"], "v")
    _motion_close(got["w"], ref["w"], "w")
    _depth_close(got["z"], ref["z"])
    nz = ref["depth_map"] != 0
    assert np.array_equal(got["depth_map"] != 0, nz)
    _depth_close(got["depth_map"][nz], ref["depth_map"][nz])
    # and it is the same raster the stage-wise glue produces from the refined depths
    inl2 = inl.copy(); inl2[2::3] = got["z"]
    g = ctx.depth_glue(inl2, c["m"], got["v"], c["K4"], c["rows"], c["cols"])
    dm = g[2] if isinstance(g, tuple) else g["depth_map"]
    assert np.array_equal(np.asarray(dm).reshape(-1), got["depth_map"].reshape(-1))
    diff = np.abs(got["rectified"].astype(np.int32) - ref["rectified"].astype(np.int32)).max(axis=2)
    assert (diff <= 1).mean() >= 0.999


def test_sharded_sequence_driver(ctx, oracle, synth, pkg):
    """sequence.run_shard_sequence: a rank's block of pairs through the pipelined entry point in batches;
    two 'ranks' together produce the records of the whole sequence."""
    import importlib
    seq = importlib.import_module("rs-aware-differential-sfm_b200.sequence")
    cases = _seq_pairs(oracle, synth, 5, False)

    def make_pair(i):
        c = cases[i]
        R = c["ransac"]
        return dict(flow=c["flow"][:2 * c["m"]], inliers3=c["inliers3"], alpha=c["alpha_in"], alpha_k=c["alpha_k_in"],
                    image=c["P"]["image"], m=c["m"], v=R["v"], w=R["w"], k=R["k"])

    whole = ctx.refine_rectify_sequence([make_pair(i) for i in range(5)], False, False, cases[0]["K4"], cases[0]["gamma"])
    full = np.zeros((5, seq.RECORD))
    for rank in range(2):
        lo, rec, res = seq.run_shard_sequence(ctx, make_pair, 5, rank, 2, False, False, cases[0]["K4"], cases[0]["gamma"], batch=2)
        full[lo:lo + rec.shape[0]] = rec
        for j, r in enumerate(res):
            assert np.array_equal(r["rectified"], whole[lo + j]["rectified"])
    for i in range(5):
        assert np.array_equal(full[i, 0:3], whole[i]["v"]) and np.array_equal(full[i, 3:6], whole[i]["w"])
        assert full[i, 7] == whole[i]["summary"]["iterations"]


def test_host_buffers_from_the_library(capi, ctx, oracle, case_cv):
    """rsdsfm_host_alloc: page-locked (optionally write-combined) host arrays are ordinary inputs / outputs of
    the RSDSFM_HOST paths."""
    c = case_cv
    R = c["ransac"]
    bufs = []

    def pinned(a, wc):
        hb = capi.HostBuffer(a.shape, a.dtype, write_combined=wc)
        hb.array[...] = a
        bufs.append(hb)
        return hb.array

    want = ctx.refine_rectify(c["flow"], c["inliers3"], c["alpha_in"], c["alpha_k_in"], c["m"], R["v"], R["w"], R["k"], False, False,
                              c["P"]["image"], c["K4"], c["gamma"])
    out = (pinned(np.zeros(c["m"]), False), pinned(np.zeros(c["rows"] * c["cols"]), False), pinned(np.zeros_like(c["P"]["image"]), False))
    got = ctx.refine_rectify(pinned(c["flow"][:2 * c["m"]], True), pinned(c["inliers3"], True), pinned(c["alpha_in"], True),
                             pinned(c["alpha_k_in"], True), c["m"], R["v"], R["w"], R["k"], False, False, pinned(c["P"]["image"], True),
                             c["K4"], c["gamma"], out=out)
    assert np.array_equal(got["v"], want["v"]) and np.array_equal(got["z"], want["z"])
    assert np.array_equal(got["rectified"], want["rectified"]) and np.array_equal(got["depth_map"], want["depth_map"])
    for hb in bufs:
        hb.free()


# ---------------------------------------------------------------------------- compact host interface
@pytest.mark.parametrize("mem", ["host", "device"])
@pytest.mark.parametrize("f32", [False, True])
def test_compact_sequence_equals_expanded_calls(ctx, oracle, synth, mem, f32):
    """rsdsfm_refine_rectify_compact_sequence takes the flow field and rsdsfm_ransac's outputs (mask, inverse depths
    over the flattened points) and rebuilds coordinates, alpha factors, pairing and start depths on the device:
    the solver sees bit-identical inputs, so single-lane results are bit-identical to rsdsfm_refine_rectify on the
    expanded arrays (m < n: the reference's flow pairing is active; zero-flow pixels are dropped)."""
    import torch
    rows, cols = 120, 160
    K4 = helpers.small_K(8)
    dev = torch.device("cuda", 0)
    pairs, refs = [], []
    for i in range(3):
        P = synth.make_pair(rows, cols, K4, gamma=0.95, seed=(60, 61, 66)[i], k=0.5, noise_sigma_px=0.1, outlier_frac=0.05 + 0.02 * i,
                            zero_flow_frac=0.2, flow_f32=f32)
        n, coord, flow, cpx, fpx = oracle.flatten(P["flow_img"], K4, 0.95)
        alpha = oracle.get_alpha(fpx, n, rows, 0.95)
        alpha_k = oracle.get_alpha_k(cpx, fpx, n, rows, 0.95)
        R = ctx.ransac(coord[:2 * n], flow[:2 * n], alpha, alpha_k, n, True, synth.sample_list(n, 6, seed=80 + i), 0.01)
        inl, a_in, ak_in, ix, m = ctx.gather_inliers(coord[:2 * n], alpha, alpha_k, n, R["mask"], R["inv_depth"])
        assert 0 < m < n < rows * cols
        refs.append(ctx.refine_rectify(flow[:2 * n], inl, a_in, ak_in, m, R["v"], R["w"], R["k"], True, False, P["image"], K4, 0.95))
        fi = P["flow_img"].astype(np.float32) if f32 else P["flow_img"]
        e = dict(flow_img=fi, image=P["image"], mask=R["mask"], inv_depth=R["inv_depth"], n=n, m=m, v=R["v"], w=R["w"], k=R["k"])
        if mem == "device":
            e = {k: (torch.from_numpy(np.ascontiguousarray(v)).to(dev) if isinstance(v, np.ndarray) and v.size > 7 else v) for k, v in e.items()}
        pairs.append(e)
    host = (lambda a: a.cpu().numpy()) if mem == "device" else (lambda a: a)
    # one at a time (synchronous single-lane calls): bit-identical
    for e, ref in zip(pairs, refs):
        got = ctx.refine_rectify_compact_sequence([e], True, False, K4, 0.95)[0]
        assert got["status"] == 0 and got["summary"]["iterations"] == ref["summary"]["iterations"]
        assert np.array_equal(got["v"], ref["v"]) and np.array_equal(got["w"], ref["w"]) and got["k"] == ref["k"]
        assert np.array_equal(host(got["z"]), ref["z"])
        assert np.array_equal(host(got["depth_map"]), ref["depth_map"]) and np.array_equal(host(got["rectified"]), ref["rectified"])
    # pipelined over the compute lanes (bit-identical whatever the lane's share of the SMs), depth map not requested
    res = ctx.refine_rectify_compact_sequence(pairs, True, False, K4, 0.95, want_depth_map=False)
    for got, ref in zip(res, refs):
        assert got["status"] == 0 and got["depth_map"] is None
        assert got["summary"]["iterations"] == ref["summary"]["iterations"]
        assert np.array_equal(got["v"], ref["v"]) and np.array_equal(got["w"], ref["w"]) and got["k"] == ref["k"]
        assert np.array_equal(host(got["rectified"]), ref["rectified"]) and np.array_equal(host(got["z"]), ref["z"])


def test_compact_sequence_rejects_wrong_counts(capi, ctx, oracle, case_cv):
    """n / m are checked against the flow field and the mask on the device."""
    c = case_cv
    R = c["ransac"]
    e = dict(flow_img=c["P"]["flow_img"], image=c["P"]["image"], mask=R["mask"], inv_depth=R["inv_depth"], n=c["n"], m=c["m"] - 1,
             v=R["v"], w=R["w"], k=R["k"])
    with pytest.raises(capi.RsdsfmError):
        ctx.refine_rectify_compact_sequence([e], False, False, c["K4"], c["gamma"])
    e["m"] = c["m"]
    ok = ctx.refine_rectify_compact_sequence([e], False, False, c["K4"], c["gamma"])[0]
    ref = ctx.refine_rectify(c["flow"], c["inliers3"], c["alpha_in"], c["alpha_k_in"], c["m"], R["v"], R["w"], R["k"], False, False,
                             c["P"]["image"], c["K4"], c["gamma"])
    assert ok["status"] == 0 and np.array_equal(ok["rectified"], ref["rectified"]) and np.array_equal(ok["z"], ref["z"])
