"""ctypes binding of librsdsfm.so (include/rsdsfm.h) for the tests and bench.py.

Array arguments may be numpy arrays (RSDSFM_HOST: staged by the library) or torch CUDA tensors
(RSDSFM_DEVICE: used in place).  There is no fallback: if the shared library is missing or no
B200 is visible, construction of a Context raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RSDSFM_LIB") or os.path.join(_HERE, "librsdsfm.so")   # RSDSFM_LIB: an experimental build (tools/)

HOST, DEVICE = 0, 1
DEPTH_COLMAJOR, DEPTH_ROWMAJOR = 0, 1

EXPORTS = [
    "rsdsfm_version", "rsdsfm_lm_default_options", "rsdsfm_create", "rsdsfm_destroy", "rsdsfm_last_error",
    "rsdsfm_synchronize", "rsdsfm_launch_count", "rsdsfm_profile_enable", "rsdsfm_profile_read", "rsdsfm_profile_detail", "rsdsfm_flatten", "rsdsfm_alpha", "rsdsfm_solve9",
    "rsdsfm_ransac_score", "rsdsfm_ransac", "rsdsfm_gather_inliers", "rsdsfm_estimate_inverse_depths",
    "rsdsfm_refine", "rsdsfm_depth_glue", "rsdsfm_set_relative_pose", "rsdsfm_backproject", "rsdsfm_fill_cracks",
    "rsdsfm_refine_rectify", "rsdsfm_refine_rectify_sequence", "rsdsfm_refine_rectify_compact_sequence", "rsdsfm_pipeline_pair", "rsdsfm_pipeline_sequence",
    "rsdsfm_relocate_pose", "rsdsfm_reprojection_error", "rsdsfm_true_flow", "rsdsfm_host_alloc", "rsdsfm_host_free",
    "rsdsfm_peer_export", "rsdsfm_peer_connect", "rsdsfm_peer_connect_local", "rsdsfm_peer_disconnect",
]


class LmOptions(C.Structure):
    _fields_ = [("max_num_iterations", C.c_int), ("function_tolerance", C.c_double),
                ("gradient_tolerance", C.c_double), ("parameter_tolerance", C.c_double),
                ("initial_trust_region_radius", C.c_double), ("max_trust_region_radius", C.c_double),
                ("min_trust_region_radius", C.c_double), ("min_relative_decrease", C.c_double),
                ("min_lm_diagonal", C.c_double), ("max_lm_diagonal", C.c_double),
                ("max_num_consecutive_invalid_steps", C.c_int)]


class LmSummary(C.Structure):
    _fields_ = [("termination", C.c_int), ("reason", C.c_int), ("iterations", C.c_int),
                ("num_successful", C.c_int), ("num_unsuccessful", C.c_int), ("initial_cost", C.c_double),
                ("final_cost", C.c_double), ("final_radius", C.c_double),
                ("final_gradient_max_norm", C.c_double), ("device_ms", C.c_double)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class PairIO(C.Structure):
    """rsdsfm_pair_io (include/rsdsfm.h)."""
    _fields_ = [("flow", C.c_void_p), ("inliers3", C.c_void_p), ("alpha", C.c_void_p), ("alpha_k", C.c_void_p),
                ("image", C.c_void_p), ("m", C.c_int), ("status", C.c_int),
                ("v", C.c_double * 3), ("w", C.c_double * 3), ("k", C.c_double),
                ("z_out", C.c_void_p), ("depth_map", C.c_void_p), ("rectified", C.c_void_p),
                ("summary", LmSummary)]


class CompactPairIO(C.Structure):
    """rsdsfm_compact_pair_io (include/rsdsfm.h)."""
    _fields_ = [("flow_img", C.c_void_p), ("image", C.c_void_p), ("mask", C.c_void_p), ("inv_depth", C.c_void_p),
                ("n", C.c_int), ("m", C.c_int), ("status", C.c_int),
                ("v", C.c_double * 3), ("w", C.c_double * 3), ("k", C.c_double),
                ("z_out", C.c_void_p), ("depth_map", C.c_void_p), ("rectified", C.c_void_p),
                ("summary", LmSummary)]


class PipelineParams(C.Structure):
    """rsdsfm_pipeline_params (include/rsdsfm.h)."""
    _fields_ = [("rows", C.c_int), ("cols", C.c_int), ("K4", C.c_double * 4), ("gamma", C.c_double),
                ("flow_threshold", C.c_double), ("ransac_tolerance", C.c_double), ("num_hypotheses", C.c_int),
                ("const_acceleration", C.c_int), ("gs_mode", C.c_int), ("use_refinement", C.c_int),
                ("repair_pairing", C.c_int), ("layout", C.c_int), ("emulate_padding", C.c_int)]


class PipelineIO(C.Structure):
    """rsdsfm_pipeline_io (include/rsdsfm.h)."""
    _fields_ = [("flow_img", C.c_void_p), ("image", C.c_void_p), ("samples", C.c_void_p), ("draws", C.c_void_p),
                ("depth_map", C.c_void_p), ("rectified", C.c_void_p),
                ("status", C.c_int), ("n", C.c_int), ("m", C.c_int), ("best_idx", C.c_int),
                ("ransac_motion", C.c_double * 7), ("v", C.c_double * 3), ("w", C.c_double * 3), ("k", C.c_double),
                ("summary", LmSummary)]


_lib = None


def load():
    """dlopen the product library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("librsdsfm.so is missing: run __graft_entry__.build() (nvcc, sm_100a); "
                               "there is no CPU fallback")
        _lib = C.CDLL(LIB_PATH)
        _lib.rsdsfm_last_error.restype = C.c_char_p
        _lib.rsdsfm_last_error.argtypes = [C.c_void_p]
        _lib.rsdsfm_launch_count.restype = C.c_longlong
        _lib.rsdsfm_launch_count.argtypes = [C.c_void_p]
        _lib.rsdsfm_destroy.argtypes = [C.c_void_p]
        _lib.rsdsfm_destroy.restype = None
    return _lib


def _is_torch(a):
    return a is not None and type(a).__module__.startswith("torch")


def _ptr(a):
    if a is None:
        return None
    if _is_torch(a):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)


def _mem(*arrays):
    kinds = {(_is_torch(a)) for a in arrays if a is not None}
    if len(kinds) > 1:
        raise ValueError("mixing host (numpy) and device (torch) arrays in one call")
    return DEVICE if (kinds and kinds.pop()) else HOST


def _f64(a):
    if _is_torch(a):
        assert a.is_cuda and a.is_contiguous() and str(a.dtype) == "torch.float64"
        return a
    return np.ascontiguousarray(a, dtype=np.float64)


def _small(a, n):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))
    assert a.size == n
    return a


class RsdsfmError(RuntimeError):
    pass


class Context:
    """One rsdsfm_ctx (one GPU, one stream)."""

    def __init__(self, device=0, stream=None):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.rsdsfm_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != 0:
            raise RsdsfmError("rsdsfm_create failed (%d): %s" % (rc, self.lib.rsdsfm_last_error(None).decode()))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.rsdsfm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise RsdsfmError("rsdsfm error %d: %s" % (rc, self.lib.rsdsfm_last_error(self.h).decode()))

    def synchronize(self):
        self._ck(self.lib.rsdsfm_synchronize(self.h))

    def launch_count(self):
        return int(self.lib.rsdsfm_launch_count(self.h))

    def profile_enable(self, on=True):
        self._ck(self.lib.rsdsfm_profile_enable(self.h, int(on)))

    def profile_read(self):
        out = np.zeros(8)
        self._ck(self.lib.rsdsfm_profile_read(self.h, _ptr(out)))
        det = np.zeros(8)
        self._ck(self.lib.rsdsfm_profile_detail(self.h, _ptr(det)))
        return dict(pass_a_ms=out[0], pass_a_launches=int(out[1]), pass_a_blocks=out[2],
                    pass_b_ms=out[3], pass_b_launches=int(out[4]), pass_b_blocks=out[5],
                    kernel_ms=out[6], kernel_launches=int(out[7]),
                    a_loop_ms=det[0], a_reduce_ms=det[1], a_ctl_ms=det[2],
                    b_loop_ms=det[3], b_reduce_ms=det[4], b_ctl_ms=det[5], a_logic_ms=det[6], b_logic_ms=det[7])

    # ---- a2
    def flatten(self, flow_img, K4, gamma, thr=1e-10, out=None):
        """numpy in -> numpy out: (n, coord, flow, coord_px, flow_px, pixel_index)."""
        flow_img = _f64(flow_img)
        rows, cols = int(flow_img.shape[0]), int(flow_img.shape[1])
        tot = rows * cols
        K4 = _small(K4, 4)
        n = C.c_int(0)
        if _is_torch(flow_img):
            import torch
            mk = lambda k, dt: torch.empty(k, dtype=dt, device=flow_img.device)
            coord, flow, cpx, fpx = (mk(2 * tot, torch.float64) for _ in range(4))
            pidx = mk(tot, torch.int32)
        else:
            coord, flow, cpx, fpx = (np.empty(2 * tot) for _ in range(4))
            pidx = np.empty(tot, dtype=np.int32)
        self._ck(self.lib.rsdsfm_flatten(self.h, _mem(flow_img), _ptr(flow_img), rows, cols, _ptr(K4),
                                         C.c_double(gamma), C.c_double(thr), _ptr(coord), _ptr(flow), _ptr(cpx),
                                         _ptr(fpx), _ptr(pidx), C.byref(n)))
        return n.value, coord, flow, cpx, fpx, pidx

    # ---- a3 / a4
    def alpha(self, flow_px, q_px, n, h, gamma):
        flow_px = _f64(flow_px); q_px = _f64(q_px)
        if _is_torch(flow_px):
            import torch
            a = torch.empty(n, dtype=torch.float64, device=flow_px.device); ak = torch.empty_like(a)
        else:
            a = np.empty(n); ak = np.empty(n)
        self._ck(self.lib.rsdsfm_alpha(self.h, _mem(flow_px, q_px), _ptr(flow_px), _ptr(q_px), int(n), C.c_double(h),
                                       C.c_double(gamma), _ptr(a), _ptr(ak)))
        return a, ak

    # ---- a6 / a8
    def ransac_score(self, q, u, alpha, alpha_k, n, hyps, tol):
        q = _f64(q); u = _f64(u); alpha = _f64(alpha); alpha_k = _f64(alpha_k)
        hyps = np.ascontiguousarray(hyps, dtype=np.float64).reshape(-1, 7)
        H = hyps.shape[0]
        counts = np.zeros(H, dtype=np.int32); sumerr = np.zeros(H); best = C.c_int(-1)
        if _is_torch(q):
            import torch
            mask = torch.zeros(max(n, 1), dtype=torch.uint8, device=q.device)
            invd = torch.zeros(max(n, 1), dtype=torch.float64, device=q.device)
        else:
            mask = np.zeros(max(n, 1), dtype=np.uint8); invd = np.zeros(max(n, 1))
        self._ck(self.lib.rsdsfm_ransac_score(self.h, _mem(q, u, alpha, alpha_k), _ptr(q), _ptr(u), _ptr(alpha),
                                              _ptr(alpha_k), int(n), _ptr(hyps), H, C.c_double(tol), _ptr(counts),
                                              _ptr(sumerr), C.byref(best), _ptr(mask), _ptr(invd)))
        return dict(counts=counts, sumerr=sumerr, best_idx=best.value, mask=mask[:n], inv_depth=invd[:n])

    def ransac(self, q, u, alpha, alpha_k, n, use_alpha_k, samples, tol):
        q = _f64(q); u = _f64(u); alpha = _f64(alpha); alpha_k = _f64(alpha_k)
        samples = np.ascontiguousarray(samples, dtype=np.int32).reshape(-1, 9)
        H = samples.shape[0]
        counts = np.zeros(H, dtype=np.int32); sumerr = np.zeros(H); best = C.c_int(-1)
        best7 = np.zeros(7); hyps = np.zeros((H, 7))
        if _is_torch(q):
            import torch
            mask = torch.zeros(n, dtype=torch.uint8, device=q.device)
            invd = torch.zeros(n, dtype=torch.float64, device=q.device)
        else:
            mask = np.zeros(n, dtype=np.uint8); invd = np.zeros(n)
        self._ck(self.lib.rsdsfm_ransac(self.h, _mem(q, u, alpha, alpha_k), _ptr(q), _ptr(u), _ptr(alpha), _ptr(alpha_k),
                                        int(n), int(use_alpha_k), _ptr(samples), H, C.c_double(tol), _ptr(counts),
                                        _ptr(sumerr), C.byref(best), _ptr(best7), _ptr(mask), _ptr(invd), _ptr(hyps)))
        return dict(counts=counts, sumerr=sumerr, best_idx=best.value, w=best7[0:3].copy(), v=best7[3:6].copy(),
                    k=float(best7[6]), mask=mask, inv_depth=invd, hyps=hyps)

    def gather_inliers(self, q, alpha, alpha_k, n, mask, inv_depth):
        q = _f64(q); alpha = _f64(alpha); alpha_k = _f64(alpha_k); inv_depth = _f64(inv_depth)
        if _is_torch(q):
            import torch
            mk = lambda k, dt: torch.empty(max(k, 1), dtype=dt, device=q.device)
            inl = mk(3 * n, torch.float64); a = mk(n, torch.float64); ak = mk(n, torch.float64); ix = mk(n, torch.int32)
        else:
            mask = np.ascontiguousarray(mask, dtype=np.uint8)
            inl = np.empty(3 * max(n, 1)); a = np.empty(max(n, 1)); ak = np.empty(max(n, 1))
            ix = np.empty(max(n, 1), dtype=np.int32)
        m = C.c_int(0)
        self._ck(self.lib.rsdsfm_gather_inliers(self.h, _mem(q, alpha, alpha_k, mask, inv_depth), _ptr(q), _ptr(alpha),
                                                _ptr(alpha_k), int(n), _ptr(mask), _ptr(inv_depth), _ptr(inl), _ptr(a),
                                                _ptr(ak), _ptr(ix), C.byref(m)))
        m = m.value
        return inl[:3 * m], a[:m], ak[:m], ix[:m], m

    def estimate_inverse_depths(self, coord, flow, n, v, w, k, alpha, alpha_k):
        coord = _f64(coord); flow = _f64(flow); alpha = _f64(alpha); alpha_k = _f64(alpha_k)
        v = _small(v, 3); w = _small(w, 3)
        if _is_torch(coord):
            import torch
            out = torch.empty(max(n, 1), dtype=torch.float64, device=coord.device)
        else:
            out = np.empty(max(n, 1))
        S = LmSummary()
        self._ck(self.lib.rsdsfm_estimate_inverse_depths(self.h, _mem(coord, flow, alpha, alpha_k), _ptr(coord), _ptr(flow),
                                                         int(n), _ptr(v), _ptr(w), C.c_double(k), _ptr(alpha),
                                                         _ptr(alpha_k), _ptr(out), C.byref(S)))
        return out[:n], S.as_dict()

    # ---- a9
    def refine(self, flow, inliers3, alpha, alpha_k, m, v, w, k, const_acc, flow_index=None, opts=None):
        flow = _f64(flow); inliers3 = _f64(inliers3); alpha = _f64(alpha); alpha_k = _f64(alpha_k)
        v = _small(v, 3).copy(); w = _small(w, 3).copy(); kk = C.c_double(k)
        if _is_torch(flow):
            import torch
            z = torch.empty(max(m, 1), dtype=torch.float64, device=flow.device)
        else:
            z = np.empty(max(m, 1))
            if flow_index is not None:
                flow_index = np.ascontiguousarray(flow_index, dtype=np.int32)
        S = LmSummary()
        self._ck(self.lib.rsdsfm_refine(self.h, _mem(flow, inliers3, alpha, alpha_k, flow_index), _ptr(flow),
                                        _ptr(inliers3), _ptr(alpha), _ptr(alpha_k), int(m), _ptr(v), _ptr(w),
                                        C.byref(kk), int(const_acc), _ptr(flow_index),
                                        C.byref(opts) if opts is not None else None, _ptr(z), C.byref(S)))
        return v, w, kk.value, z[:m], S.as_dict()

    # ---- a10
    def depth_glue(self, inliers3, m, v, K4, rows, cols, z_min_init=float("inf"), layout=DEPTH_COLMAJOR, want_img=False):
        v = _small(v, 3).copy(); K4 = _small(K4, 4)
        if _is_torch(inliers3):
            import torch
            inl = inliers3.clone()
            dm = torch.empty(rows * cols, dtype=torch.float64, device=inl.device)
            img = torch.empty(rows * cols, dtype=torch.uint8, device=inl.device) if want_img else None
        else:
            inl = _f64(inliers3).copy()
            dm = np.empty(rows * cols)
            img = np.empty(rows * cols, dtype=np.uint8) if want_img else None
        self._ck(self.lib.rsdsfm_depth_glue(self.h, _mem(inl), _ptr(inl), int(m), _ptr(v), _ptr(K4), rows, cols,
                                            C.c_double(z_min_init), int(layout), _ptr(dm), _ptr(img)))
        return inl, v, dm, img

    # ---- a12
    def set_relative_pose(self, v, w, k, gamma, rows):
        v = _small(v, 3); w = _small(w, 3)
        R = np.empty(rows * 9); t = np.empty(rows * 3)
        self._ck(self.lib.rsdsfm_set_relative_pose(_ptr(v), _ptr(w), C.c_double(k), C.c_double(gamma), rows, _ptr(R), _ptr(t)))
        return R.reshape(rows, 3, 3), t.reshape(rows, 3)

    # ---- a13 / a14
    def backproject(self, image, depth, K4, R, t, gs_mode=False, layout=DEPTH_COLMAJOR, want_coords=False):
        K4 = _small(K4, 4)
        R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1); t = np.ascontiguousarray(t, dtype=np.float64).reshape(-1)
        depth = _f64(depth)
        rows, cols = int(image.shape[0]), int(image.shape[1])
        if _is_torch(image):
            import torch
            out = torch.empty_like(image)
            coords = torch.empty((rows, cols, 3), dtype=torch.float32, device=image.device) if want_coords else None
        else:
            image = np.ascontiguousarray(image, dtype=np.uint8)
            out = np.empty_like(image)
            coords = np.empty((rows, cols, 3), dtype=np.float32) if want_coords else None
        self._ck(self.lib.rsdsfm_backproject(self.h, _mem(image, depth), _ptr(image), _ptr(depth), int(layout), rows, cols,
                                             _ptr(K4), _ptr(R), _ptr(t), int(gs_mode), _ptr(out), _ptr(coords)))
        return out, coords

    # ---- a15
    def fill_cracks(self, image, offset=1):
        rows, cols = int(image.shape[0]), int(image.shape[1])
        if _is_torch(image):
            import torch
            out = torch.empty_like(image)
        else:
            image = np.ascontiguousarray(image, dtype=np.uint8)
            out = np.empty_like(image)
        self._ck(self.lib.rsdsfm_fill_cracks(self.h, _mem(image), _ptr(image), rows, cols, C.c_uint(offset), _ptr(out)))
        return out

    # ---- fused driver
    def refine_rectify(self, flow, inliers3, alpha, alpha_k, m, v, w, k, const_acc, gs_mode, image, K4, gamma,
                       layout=DEPTH_COLMAJOR, out=None):
        """Returns dict(v, w, k, z, depth_map, rectified, summary).  `out` may carry preallocated
        (z, depth_map, rectified) buffers (pinned numpy or torch CUDA)."""
        flow = _f64(flow); inliers3 = _f64(inliers3); alpha = _f64(alpha); alpha_k = _f64(alpha_k)
        v = _small(v, 3).copy(); w = _small(w, 3).copy(); kk = C.c_double(k); K4 = _small(K4, 4)
        rows, cols = int(image.shape[0]), int(image.shape[1])
        if out is not None:
            z, dm, rect = out
        elif _is_torch(flow):
            import torch
            z = torch.empty(m, dtype=torch.float64, device=flow.device)
            dm = torch.empty(rows * cols, dtype=torch.float64, device=flow.device)
            rect = torch.empty_like(image)
        else:
            image = np.ascontiguousarray(image, dtype=np.uint8)
            z = np.empty(m); dm = np.empty(rows * cols); rect = np.empty_like(image)
        S = LmSummary()
        self._ck(self.lib.rsdsfm_refine_rectify(self.h, _mem(flow, inliers3, alpha, alpha_k, image), _ptr(flow),
                                                _ptr(inliers3), _ptr(alpha), _ptr(alpha_k), int(m), _ptr(v), _ptr(w),
                                                C.byref(kk), int(const_acc), int(gs_mode), _ptr(image), rows, cols,
                                                _ptr(K4), C.c_double(gamma), int(layout), _ptr(z), _ptr(dm), _ptr(rect),
                                                C.byref(S)))
        return dict(v=v, w=w, k=kk.value, z=z, depth_map=dm, rectified=rect, summary=S.as_dict())

    def refine_rectify_sequence(self, pairs, const_acc, gs_mode, K4, gamma, layout=DEPTH_COLMAJOR):
        """rsdsfm_refine_rectify_sequence.  `pairs`: list of dicts with flow, inliers3, alpha, alpha_k,
        image, m, v, w, k and optionally out=(z, depth_map, rectified) (pinned numpy or torch CUDA; all
        pairs host or all device).  Returns one dict per pair like refine_rectify (+ status)."""
        n = len(pairs)
        arr = (PairIO * max(n, 1))()
        keep, outs, mems = [], [], set()
        rows = cols = 0
        for i, p in enumerate(pairs):
            flow = _f64(p["flow"]); inl = _f64(p["inliers3"]); a = _f64(p["alpha"]); ak = _f64(p["alpha_k"])
            image = p["image"]
            m = int(p["m"])
            rows, cols = int(image.shape[0]), int(image.shape[1])
            if p.get("out") is not None:
                z, dm, rect = p["out"]
            elif _is_torch(flow):
                import torch
                z = torch.empty(m, dtype=torch.float64, device=flow.device)
                dm = torch.empty(rows * cols, dtype=torch.float64, device=flow.device)
                rect = torch.empty_like(image)
            else:
                image = np.ascontiguousarray(image, dtype=np.uint8)
                z = np.empty(m); dm = np.empty(rows * cols); rect = np.empty_like(image)
            mems.add(_mem(flow, inl, a, ak, image, z, dm, rect))
            keep.append((flow, inl, a, ak, image))
            outs.append((z, dm, rect))
            e = arr[i]
            e.flow, e.inliers3, e.alpha, e.alpha_k, e.image = (_ptr(x).value for x in (flow, inl, a, ak, image))
            e.m = m
            e.v[:] = list(_small(p["v"], 3)); e.w[:] = list(_small(p["w"], 3)); e.k = float(p["k"])
            e.z_out, e.depth_map, e.rectified = _ptr(z).value, _ptr(dm).value, _ptr(rect).value
        if len(mems) > 1:
            raise ValueError("mixing host and device pairs in one sequence")
        K4 = _small(K4, 4)
        rc = self.lib.rsdsfm_refine_rectify_sequence(self.h, mems.pop() if mems else HOST, n, arr, int(const_acc), int(gs_mode),
                                                     rows, cols, _ptr(K4), C.c_double(gamma), int(layout))
        res = []
        for i in range(n):
            e = arr[i]
            z, dm, rect = outs[i]
            res.append(dict(v=np.array(e.v[:]), w=np.array(e.w[:]), k=float(e.k), z=z, depth_map=dm, rectified=rect,
                            summary=e.summary.as_dict(), status=int(e.status)))
        self._ck(rc)
        return res

    def refine_rectify_compact_sequence(self, pairs, const_acc, gs_mode, K4, gamma, layout=DEPTH_COLMAJOR, thr=1e-10,
                                        want_depth_map=True):
        """rsdsfm_refine_rectify_compact_sequence.  `pairs`: list of dicts with flow_img (rows x cols x 2, float64 or
        float32 -- all pairs alike), image, mask (n, uint8), inv_depth (n), n, m, v, w, k and optionally
        out=(z, depth_map or None, rectified).  numpy (pinned) = host, torch CUDA = device."""
        n = len(pairs)
        arr = (CompactPairIO * max(n, 1))()
        keep, outs, mems = [], [], set()
        rows = cols = 0
        f32 = None
        for i, p in enumerate(pairs):
            fi, image, mask, invd = p["flow_img"], p["image"], p["mask"], _f64(p["inv_depth"])
            is32 = str(fi.dtype).endswith("float32")
            f32 = is32 if f32 is None else f32
            assert is32 == f32, "all pairs must store the flow alike"
            rows, cols = int(image.shape[0]), int(image.shape[1])
            m = int(p["m"])
            if _is_torch(fi):
                import torch
                assert fi.is_cuda and fi.is_contiguous() and mask.is_contiguous() and str(mask.dtype) == "torch.uint8"
            else:
                fi = np.ascontiguousarray(fi); image = np.ascontiguousarray(image, dtype=np.uint8)
                mask = np.ascontiguousarray(mask, dtype=np.uint8)
            if p.get("out") is not None:
                z, dm, rect = p["out"]
            elif _is_torch(fi):
                import torch
                z = torch.empty(max(m, 1), dtype=torch.float64, device=fi.device)
                dm = torch.empty(rows * cols, dtype=torch.float64, device=fi.device) if want_depth_map else None
                rect = torch.empty_like(image)
            else:
                z = np.empty(max(m, 1)); dm = np.empty(rows * cols) if want_depth_map else None; rect = np.empty_like(image)
            mems.add(_mem(fi, image, mask, invd, z, dm, rect))
            keep.append((fi, image, mask, invd))
            outs.append((z, dm, rect))
            e = arr[i]
            e.flow_img, e.image, e.mask, e.inv_depth = (_ptr(x).value for x in (fi, image, mask, invd))
            e.n, e.m = int(p["n"]), m
            e.v[:] = list(_small(p["v"], 3)); e.w[:] = list(_small(p["w"], 3)); e.k = float(p["k"])
            e.z_out, e.rectified = _ptr(z).value, _ptr(rect).value
            e.depth_map = _ptr(dm).value if dm is not None else None
        if len(mems) > 1:
            raise ValueError("mixing host and device pairs in one sequence")
        K4 = _small(K4, 4)
        rc = self.lib.rsdsfm_refine_rectify_compact_sequence(self.h, mems.pop() if mems else HOST, n, arr, int(bool(f32)), C.c_double(thr),
                                                             int(const_acc), int(gs_mode), rows, cols, _ptr(K4), C.c_double(gamma), int(layout))
        res = []
        for i in range(n):
            e = arr[i]
            z, dm, rect = outs[i]
            res.append(dict(v=np.array(e.v[:]), w=np.array(e.w[:]), k=float(e.k), z=z[:e.m], depth_map=dm, rectified=rect,
                            summary=e.summary.as_dict(), status=int(e.status)))
        self._ck(rc)
        return res

    # ---- row split of one solve over several GPUs
    def peer_export(self):
        """rsdsfm_peer_export: the 64-byte CUDA IPC handle of this context's mailbox (bytes)."""
        buf = C.create_string_buffer(64)
        self._ck(self.lib.rsdsfm_peer_export(self.h, buf))
        return buf.raw

    def peer_connect(self, handles, my_index):
        """rsdsfm_peer_connect: `handles` = the members' exported handles in group order (list of 64-byte strings)."""
        blob = b"".join(handles)
        assert len(blob) == 64 * len(handles)
        self._ck(self.lib.rsdsfm_peer_connect(self.h, len(handles), int(my_index), blob))

    def peer_disconnect(self):
        self._ck(self.lib.rsdsfm_peer_disconnect(self.h))

    @staticmethod
    def peer_connect_local(contexts):
        """rsdsfm_peer_connect_local: contexts of THIS process (one per GPU) form a group, in list order."""
        arr = (C.c_void_p * len(contexts))(*[c.h for c in contexts])
        rc = contexts[0].lib.rsdsfm_peer_connect_local(arr, len(contexts))
        contexts[0]._ck(rc)

    # ---- SURVEY 8(f)-1
    def reprojection_error(self, coords3d, unproj, R_gt, t_gt, depth_est, K4, max_norm=1.0, layout=DEPTH_COLMAJOR,
                           want_image=False, want_gt_depth=False):
        """rsdsfm_reprojection_error.  coords3d: rows x cols x 3 float32; unproj: (x, y, z) maps of rows*cols doubles
        in `layout`; depth_est likewise.  numpy = host, torch CUDA = device.  Returns dict(mean_error, mean_scale,
        num_outliers, points_used, error_image, gt_depth_map)."""
        rows, cols = int(coords3d.shape[0]), int(coords3d.shape[1])
        ux, uy, uz = (_f64(a) for a in unproj)
        depth_est = _f64(depth_est)
        R = np.ascontiguousarray(np.asarray(R_gt, dtype=np.float64).reshape(-1)); t = np.ascontiguousarray(np.asarray(t_gt, dtype=np.float64).reshape(-1))
        assert R.size == 9 * rows and t.size == 3 * rows
        K4 = _small(K4, 4)
        if _is_torch(coords3d):
            import torch
            assert coords3d.is_cuda and coords3d.is_contiguous() and str(coords3d.dtype) == "torch.float32"
            img = torch.empty((rows, cols), dtype=torch.uint8, device=coords3d.device) if want_image else None
            gd = torch.empty(rows * cols, dtype=torch.float64, device=coords3d.device) if want_gt_depth else None
        else:
            coords3d = np.ascontiguousarray(coords3d, dtype=np.float32)
            img = np.empty((rows, cols), dtype=np.uint8) if want_image else None
            gd = np.empty(rows * cols) if want_gt_depth else None
        me = C.c_double(0); ms = C.c_double(0); no = C.c_int(0); pu = C.c_int(0)
        self._ck(self.lib.rsdsfm_reprojection_error(self.h, _mem(coords3d, ux, uy, uz, depth_est, img, gd), _ptr(coords3d), _ptr(ux), _ptr(uy),
                                                    _ptr(uz), _ptr(R), _ptr(t), _ptr(depth_est), int(layout), rows, cols, _ptr(K4),
                                                    C.c_double(max_norm), C.byref(me), C.byref(ms), C.byref(no), C.byref(pu),
                                                    _ptr(img), _ptr(gd)))
        return dict(mean_error=me.value, mean_scale=ms.value, num_outliers=no.value, points_used=pu.value, error_image=img,
                    gt_depth_map=gd)

    # ---- SURVEY 8(f)-2
    def true_flow(self, unproj, R2, t2, K4, rows, cols, layout=DEPTH_COLMAJOR):
        """rsdsfm_true_flow.  unproj: (x, y, z) maps of frame 1 (rows*cols doubles in `layout`); R2, t2: scanline
        poses of frame 2.  Returns the (rows, cols, 2) flow image (numpy or torch like the inputs)."""
        ux, uy, uz = (_f64(a) for a in unproj)
        R = np.ascontiguousarray(np.asarray(R2, dtype=np.float64).reshape(-1)); t = np.ascontiguousarray(np.asarray(t2, dtype=np.float64).reshape(-1))
        assert R.size == 9 * rows and t.size == 3 * rows
        K4 = _small(K4, 4)
        if _is_torch(ux):
            import torch
            flow = torch.empty((rows, cols, 2), dtype=torch.float64, device=ux.device)
        else:
            flow = np.empty((rows, cols, 2))
        self._ck(self.lib.rsdsfm_true_flow(self.h, _mem(ux, uy, uz), _ptr(ux), _ptr(uy), _ptr(uz), _ptr(R), _ptr(t), int(layout),
                                           int(rows), int(cols), _ptr(K4), _ptr(flow)))
        return flow

    # ---- a2..a15 in one call
    @staticmethod
    def _pipeline_params(rows, cols, K4, gamma, H, tol, const_acc, gs_mode, use_refinement, repair_pairing, layout, thr,
                         emulate_padding=False):
        P = PipelineParams()
        P.rows, P.cols = int(rows), int(cols)
        P.K4[:] = list(_small(K4, 4))
        P.gamma, P.flow_threshold, P.ransac_tolerance = float(gamma), float(thr), float(tol)
        P.num_hypotheses = int(H)
        P.const_acceleration, P.gs_mode, P.use_refinement = int(const_acc), int(gs_mode), int(use_refinement)
        P.repair_pairing, P.layout = int(repair_pairing), int(layout)
        P.emulate_padding = int(emulate_padding)
        return P

    def pipeline_sequence(self, pairs, K4, gamma, tol, const_acc, gs_mode=False, use_refinement=True, repair_pairing=False,
                          layout=DEPTH_COLMAJOR, thr=1e-10, emulate_padding=False):
        """rsdsfm_pipeline_sequence (len(pairs) == 1: rsdsfm_pipeline_pair).  `pairs`: list of dicts with
        flow_img (rows x cols x 2 f64), image (rows x cols x 3 u8), and samples (H x 9 int32) or draws
        (H x 9 uint32); optionally out=(depth_map, rectified).  Returns one dict per pair."""
        n = len(pairs)
        arr = (PipelineIO * max(n, 1))()
        keep, outs, mems = [], [], set()
        rows = cols = H = 0
        for i, p in enumerate(pairs):
            fi = _f64(p["flow_img"]); image = p["image"]
            rows, cols = int(image.shape[0]), int(image.shape[1])
            smp = drw = None
            if p.get("samples") is not None:
                smp = np.ascontiguousarray(p["samples"], dtype=np.int32).reshape(-1, 9); H = smp.shape[0]
            else:
                drw = np.ascontiguousarray(p["draws"], dtype=np.uint32).reshape(-1, 9); H = drw.shape[0]
            if p.get("out") is not None:
                dm, rect = p["out"]
            elif _is_torch(fi):
                import torch
                dm = torch.empty(rows * cols, dtype=torch.float64, device=fi.device)
                rect = torch.empty_like(image)
            else:
                image = np.ascontiguousarray(image, dtype=np.uint8)
                dm = np.empty(rows * cols); rect = np.empty_like(image)
            mems.add(_mem(fi, image, dm, rect))
            keep.append((fi, image, smp, drw))
            outs.append((dm, rect))
            e = arr[i]
            e.flow_img, e.image = _ptr(fi).value, _ptr(image).value
            e.samples = _ptr(smp).value if smp is not None else None
            e.draws = _ptr(drw).value if drw is not None else None
            e.depth_map, e.rectified = _ptr(dm).value, _ptr(rect).value
        if len(mems) > 1:
            raise ValueError("mixing host and device pairs in one sequence")
        mem = mems.pop() if mems else HOST
        P = self._pipeline_params(rows, cols, K4, gamma, H, tol, const_acc, gs_mode, use_refinement, repair_pairing, layout, thr,
                                  emulate_padding)
        if n == 1:
            rc = self.lib.rsdsfm_pipeline_pair(self.h, mem, C.byref(P), C.byref(arr[0]))
        else:
            rc = self.lib.rsdsfm_pipeline_sequence(self.h, mem, C.byref(P), n, arr)
        res = []
        for i in range(n):
            e = arr[i]
            dm, rect = outs[i]
            rm = np.array(e.ransac_motion[:])
            res.append(dict(status=int(e.status), n=int(e.n), m=int(e.m), best_idx=int(e.best_idx), ransac_v=rm[0:3], ransac_w=rm[3:6],
                            ransac_k=float(rm[6]), v=np.array(e.v[:]), w=np.array(e.w[:]), k=float(e.k), depth_map=dm,
                            rectified=rect, summary=e.summary.as_dict()))
        self._ck(rc)
        return res

    def pipeline_pair(self, flow_img, image, K4, gamma, tol, const_acc, samples=None, draws=None, **kw):
        return self.pipeline_sequence([dict(flow_img=flow_img, image=image, samples=samples, draws=draws, out=kw.pop("out", None))],
                                      K4, gamma, tol, const_acc, **kw)[0]


class HostBuffer:
    """A page-locked host array from rsdsfm_host_alloc (optionally write-combined), viewed as numpy."""

    def __init__(self, shape, dtype, write_combined=False):
        self.lib = load()
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        rc = self.lib.rsdsfm_host_alloc(C.c_size_t(max(n, 1)), int(write_combined), C.byref(p))
        if rc != 0:
            raise RsdsfmError("rsdsfm_host_alloc failed: %d" % rc)
        self.ptr = p
        buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if getattr(self, "ptr", None):
            self.array = None
            self.lib.rsdsfm_host_free(self.ptr)
            self.ptr = None


def relocate_pose(R_gt, t_gt):
    """RsFrame::relocatePose (host).  Returns (R, t) with shapes (rows, 3, 3), (rows, 3)."""
    lib = load()
    R = np.ascontiguousarray(np.asarray(R_gt, dtype=np.float64).reshape(-1)); t = np.ascontiguousarray(np.asarray(t_gt, dtype=np.float64).reshape(-1))
    rows = t.size // 3
    Ro = np.empty_like(R); to = np.empty_like(t)
    rc = lib.rsdsfm_relocate_pose(_ptr(R), _ptr(t), rows, _ptr(Ro), _ptr(to))
    if rc != 0:
        raise RsdsfmError("rsdsfm_relocate_pose failed: %d" % rc)
    return Ro.reshape(rows, 3, 3), to.reshape(rows, 3)


def solve9(q9, u9, alpha9, alpha_k9, use_alpha_k):
    """minimal::calculateVelocities (host).  Returns (w, v, k)."""
    lib = load()
    q9 = _small(q9, 18); u9 = _small(u9, 18); a = _small(alpha9, 9); ak = _small(alpha_k9, 9)
    out = np.empty(7)
    rc = lib.rsdsfm_solve9(_ptr(q9), _ptr(u9), _ptr(a), _ptr(ak), int(use_alpha_k), _ptr(out))
    if rc != 0:
        raise RsdsfmError("rsdsfm_solve9 failed: %d" % rc)
    return out[0:3].copy(), out[3:6].copy(), float(out[6])
