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
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rectified image" in r.stdout
