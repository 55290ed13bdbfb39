"""ctypes binding of the CPU oracle (oracle/rsdsfm_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "librsdsfm_oracle.so")


def build(force=False):
    """Compile the C restatement with gcc (recipe: oracle/Makefile)."""
    src = [os.path.join(_HERE, f) for f in ("rsdsfm_oracle.c", "orc_linalg.h", "Makefile")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


class LmSummary(C.Structure):
    _fields_ = [
        ("termination", C.c_int), ("reason", C.c_int), ("iterations", C.c_int),
        ("num_successful", C.c_int), ("num_unsuccessful", C.c_int),
        ("initial_cost", C.c_double), ("final_cost", C.c_double), ("final_radius", C.c_double),
        ("final_gradient_max_norm", C.c_double),
        ("trace_cost", C.c_double * 64), ("trace_radius", C.c_double * 64),
        ("trace_rho", C.c_double * 64), ("trace_accepted", C.c_int * 64),
    ]

    def as_dict(self):
        n = min(self.iterations + 1, 64)
        return dict(termination=self.termination, reason=self.reason, iterations=self.iterations,
                    num_successful=self.num_successful, num_unsuccessful=self.num_unsuccessful,
                    initial_cost=self.initial_cost, final_cost=self.final_cost,
                    final_radius=self.final_radius, gmax=self.final_gradient_max_norm,
                    trace_cost=list(self.trace_cost)[:n], trace_radius=list(self.trace_radius)[:n],
                    trace_rho=list(self.trace_rho)[:n], trace_accepted=list(self.trace_accepted)[:n])


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_u8p = C.POINTER(C.c_uint8)
_fp = C.POINTER(C.c_float)


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        assert _lib.orc_sizeof_summary() == C.sizeof(LmSummary)
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def flatten(flow_img, K4, gamma, thr=1e-10):
    """main.cc:398-432.  Returns (n, coord, flow, coord_px, flow_px) with padded length rows*cols."""
    flow_img = _f64(flow_img)
    rows, cols = flow_img.shape[:2]
    tot = rows * cols
    coord, flow, cpx, fpx = (np.empty(2 * tot) for _ in range(4))
    L = lib()
    L.orc_flatten.restype = C.c_int
    n = L.orc_flatten(_d(flow_img), rows, cols, C.c_double(K4[0]), C.c_double(K4[1]), C.c_double(K4[2]),
                      C.c_double(K4[3]), C.c_double(gamma), C.c_double(thr), _d(coord), _d(flow), _d(cpx), _d(fpx))
    return n, coord, flow, cpx, fpx


def get_alpha(flow_px, n, h, gamma):
    flow_px = _f64(flow_px)
    out = np.empty(n)
    lib().orc_get_alpha(_d(flow_px), n, C.c_double(h), C.c_double(gamma), _d(out))
    return out


def get_alpha_k(q_px, flow_px, n, h, gamma):
    q_px = _f64(q_px); flow_px = _f64(flow_px)
    out = np.empty(n)
    lib().orc_get_alpha_k(_d(q_px), _d(flow_px), n, C.c_double(h), C.c_double(gamma), _d(out))
    return out


def calculate_velocities(q9, u9, alpha9, alpha_k9, use_alpha_k, svd_sign=1, evec_flip=0):
    """minimal.cc:36-177.  q9,u9: (9,2).  Returns (w, v, k)."""
    q9 = _f64(q9).reshape(-1); u9 = _f64(u9).reshape(-1)
    a = _f64(alpha9); ak = _f64(alpha_k9)
    out = np.empty(7)
    lib().orc_calculate_velocities_ex(_d(q9), _d(u9), _d(a), _d(ak), int(use_alpha_k), int(svd_sign),
                                      int(evec_flip), _d(out))
    return out[0:3].copy(), out[3:6].copy(), float(out[6])


def estimate_inverse_depths(coord, flow, n, v, w, k, alpha, alpha_k):
    coord = _f64(coord); flow = _f64(flow); alpha = _f64(alpha); alpha_k = _f64(alpha_k)
    v = _f64(v); w = _f64(w)
    out = np.empty(n)
    S = LmSummary()
    L = lib()
    L.orc_estimate_inverse_depths.restype = C.c_int
    L.orc_estimate_inverse_depths(_d(coord), _d(flow), n, _d(v), _d(w), C.c_double(k), _d(alpha), _d(alpha_k),
                                  _d(out), C.byref(S))
    return out, S.as_dict()


def ransac(q, u, alpha, alpha_k, n, use_alpha_k, tol, hyps=None, samples=None):
    """minimal.cc:209-306 with an injected hypothesis (H,7: w,v,k) or sample (H,9) list."""
    q = _f64(q); u = _f64(u); alpha = _f64(alpha); alpha_k = _f64(alpha_k)
    L = lib()
    if hyps is not None:
        hyps = _f64(hyps); H = hyps.shape[0]; mode = 0
        samples_c = None
        hyps_c = _d(hyps)
    else:
        samples = np.ascontiguousarray(samples, dtype=np.int32); H = samples.shape[0]; mode = 1
        samples_c = samples.ctypes.data_as(_ip)
        hyps_c = None
    counts = np.zeros(H, dtype=np.int32)
    sumerr = np.zeros(H)
    best_idx = C.c_int(-1)
    best7 = np.zeros(7)
    mask = np.zeros(max(n, 1), dtype=np.uint8)
    invd = np.zeros(max(n, 1))
    hyps_out = np.zeros((H, 7))
    L.orc_ransac.restype = C.c_int
    nin = L.orc_ransac(_d(q), _d(u), _d(alpha), _d(alpha_k), n, int(use_alpha_k), mode, hyps_c, samples_c, H,
                       C.c_double(tol), counts.ctypes.data_as(_ip), _d(sumerr), C.byref(best_idx), _d(best7),
                       mask.ctypes.data_as(_u8p), _d(invd), _d(hyps_out))
    return dict(num_inliers=nin, counts=counts, sumerr=sumerr, best_idx=best_idx.value, w=best7[0:3].copy(),
                v=best7[3:6].copy(), k=float(best7[6]), mask=mask[:n], inv_depth=invd[:n], hyps=hyps_out)


def gather_inliers(q, alpha, alpha_k, n, mask, inv_depth):
    q = _f64(q); alpha = _f64(alpha); alpha_k = _f64(alpha_k); inv_depth = _f64(inv_depth)
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    m = int(mask[:n].sum())
    inl = np.zeros(3 * max(m, 1)); a = np.zeros(max(m, 1)); ak = np.zeros(max(m, 1))
    L = lib()
    L.orc_gather_inliers.restype = C.c_int
    m2 = L.orc_gather_inliers(_d(q), _d(alpha), _d(alpha_k), n, mask.ctypes.data_as(_u8p), _d(inv_depth),
                              _d(inl), _d(a), _d(ak))
    assert m2 == m
    return inl[:3 * m], a[:m], ak[:m]


def nonlinear_refinement(flow, inliers3, alpha, alpha_k, m, v, w, k, const_acc, flow_index=None):
    """nonlinearRefinement.cc:183-252.  Returns (v, w, k, z[m], summary)."""
    flow = _f64(flow); inliers3 = _f64(inliers3); alpha = _f64(alpha); alpha_k = _f64(alpha_k)
    v = _f64(v).copy(); w = _f64(w).copy(); kk = C.c_double(k)
    z = np.empty(max(m, 1))
    S = LmSummary()
    fi = None
    if flow_index is not None:
        flow_index = np.ascontiguousarray(flow_index, dtype=np.int32)
        fi = flow_index.ctypes.data_as(_ip)
    L = lib()
    L.orc_nonlinear_refinement.restype = C.c_int
    L.orc_nonlinear_refinement(_d(flow), _d(inliers3), _d(alpha), _d(alpha_k), m, _d(v), _d(w), C.byref(kk),
                               int(const_acc), fi, _d(z), C.byref(S))
    return v, w, kk.value, z[:m], S.as_dict()


def depth_glue(inliers3, m, v, K4, rows, cols, z_min_init=np.inf, want_img=False):
    """main.cc:466-509.  Returns (inliers3', v', depth_map[col-major rows x cols], depth_img, zmean)."""
    inl = _f64(inliers3).copy(); v = _f64(v).copy()
    dm = np.empty(rows * cols)
    img = np.empty(rows * cols, dtype=np.uint8) if want_img else None
    zm = C.c_double(0)
    lib().orc_depth_glue(_d(inl), m, _d(v), C.c_double(K4[0]), C.c_double(K4[1]), C.c_double(K4[2]),
                         C.c_double(K4[3]), rows, cols, C.c_double(z_min_init), _d(dm),
                         img.ctypes.data_as(_u8p) if want_img else None, C.byref(zm))
    return inl, v, dm, (img.reshape(rows, cols) if want_img else None), zm.value


def set_relative_pose(v, w, k, gamma, rows):
    v = _f64(v); w = _f64(w)
    R = np.empty(rows * 9); t = np.empty(rows * 3)
    lib().orc_set_relative_pose(_d(v), _d(w), C.c_double(k), C.c_double(gamma), rows, _d(R), _d(t))
    return R.reshape(rows, 3, 3), t.reshape(rows, 3)


def back_project(image, depth_map_colmajor, K4, R, t, gs_mode=False, want_coords=False):
    image = np.ascontiguousarray(image, dtype=np.uint8)
    rows, cols = image.shape[:2]
    dm = _f64(depth_map_colmajor).reshape(-1)
    K4 = _f64(K4); R = _f64(R).reshape(-1); t = _f64(t).reshape(-1)
    out = np.empty_like(image)
    coords = np.empty((rows, cols, 3), dtype=np.float32) if want_coords else None
    lib().orc_back_project(image.ctypes.data_as(_u8p), _d(dm), rows, cols, _d(K4), _d(R), _d(t), int(gs_mode),
                           out.ctypes.data_as(_u8p), coords.ctypes.data_as(_fp) if want_coords else None)
    return out, coords


def interpolate_cracky_image(image, offset=1):
    image = np.ascontiguousarray(image, dtype=np.uint8)
    rows, cols = image.shape[:2]
    out = np.empty_like(image)
    lib().orc_interpolate_cracky_image(image.ctypes.data_as(_u8p), rows, cols, C.c_uint(offset),
                                       out.ctypes.data_as(_u8p))
    return out


def refine_rectify(flow, inliers3, alpha, alpha_k, m, v, w, k, const_acc, gs_mode, image, K4, gamma):
    """The timed region 'refine + rectify' of one pair (main.cc:457-523) on the CPU."""
    flow = _f64(flow); inliers3 = _f64(inliers3); alpha = _f64(alpha); alpha_k = _f64(alpha_k)
    image = np.ascontiguousarray(image, dtype=np.uint8)
    rows, cols = image.shape[:2]
    v = _f64(v).copy(); w = _f64(w).copy(); kk = C.c_double(k); K4 = _f64(K4)
    z = np.empty(max(m, 1)); dm = np.empty(rows * cols); out = np.empty_like(image)
    S = LmSummary()
    L = lib()
    L.orc_refine_rectify.restype = C.c_int
    L.orc_refine_rectify(_d(flow), _d(inliers3), _d(alpha), _d(alpha_k), m, _d(v), _d(w), C.byref(kk),
                         int(const_acc), int(gs_mode), image.ctypes.data_as(_u8p), rows, cols, _d(K4),
                         C.c_double(gamma), _d(z), _d(dm), out.ctypes.data_as(_u8p), C.byref(S))
    return dict(v=v, w=w, k=kk.value, z=z[:m], depth_map=dm, rectified=out, summary=S.as_dict())


# ---- ground-truth depth map, relocatePose, meanReprojectionError / createErrorImage
def groundtruth_depth_map(ux, uy, uz, R_gt, t_gt):
    """RsFrame::getGroundtruthDepthMap.  ux,uy,uz: (rows, cols) arrays; returns (rows, cols)."""
    rows, cols = ux.shape
    f = lambda a: _f64(np.asarray(a, dtype=np.float64).flatten(order="F"))
    R = _f64(R_gt).reshape(-1); t = _f64(t_gt).reshape(-1)
    out = np.empty(rows * cols)
    a, b, c = f(ux), f(uy), f(uz)
    lib().orc_groundtruth_depth_map(_d(a), _d(b), _d(c), _d(R), _d(t), rows, cols, _d(out))
    return out.reshape(cols, rows).T.copy()


def relocate_pose(R_gt, t_gt):
    R = _f64(R_gt).reshape(-1).copy(); t = _f64(t_gt).reshape(-1).copy()
    rows = t.size // 3
    lib().orc_relocate_pose(_d(R), _d(t), rows)
    return R.reshape(rows, 3, 3), t.reshape(rows, 3)


def mean_reprojection_error(coords3d, ux, uy, uz, R_gt, t_gt, depth_est_colmajor, K4, max_norm=1.0, want_image=False):
    """Camera::meanReprojectionError (+ createErrorImage).  Returns dict(mean_error, mean_scale,
    num_outliers, points_used, error_image)."""
    rows, cols = ux.shape
    f = lambda a: _f64(np.asarray(a, dtype=np.float64).flatten(order="F"))
    co = np.ascontiguousarray(coords3d, dtype=np.float32)
    R = _f64(R_gt).reshape(-1); t = _f64(t_gt).reshape(-1); K4 = _f64(K4)
    de = _f64(depth_est_colmajor).reshape(-1)
    a, b, c = f(ux), f(uy), f(uz)
    ms = C.c_double(0); no = C.c_int(0); pu = C.c_int(0)
    img = np.zeros((rows, cols), dtype=np.uint8) if want_image else None
    fn = lib().orc_mean_reprojection_error
    fn.restype = C.c_double
    err = fn(co.ctypes.data_as(_fp), _d(a), _d(b), _d(c), _d(R), _d(t), _d(de), rows, cols, _d(K4), C.c_double(max_norm),
             C.byref(ms), C.byref(no), C.byref(pu), img.ctypes.data_as(_u8p) if want_image else None)
    return dict(mean_error=float(err), mean_scale=ms.value, num_outliers=no.value, points_used=pu.value, error_image=img)


# ---- ground-truth flow between two RS frames
def true_flow(ux, uy, uz, R2, t2, K4):
    """Camera::calculateTrueFlow.  ux,uy,uz: (rows, cols) unprojection maps of frame 1; R2,t2: scanline
    poses of frame 2.  Returns (rows, cols, 2)."""
    rows, cols = ux.shape
    f = lambda a: _f64(np.asarray(a, dtype=np.float64).flatten(order="F"))
    R = _f64(R2).reshape(-1); t = _f64(t2).reshape(-1); K4 = _f64(K4)
    out = np.empty(rows * cols * 2)
    a, b, c = f(ux), f(uy), f(uz)
    lib().orc_true_flow(_d(a), _d(b), _d(c), _d(R), _d(t), rows, cols, _d(K4), _d(out))
    return out.reshape(rows, cols, 2)
