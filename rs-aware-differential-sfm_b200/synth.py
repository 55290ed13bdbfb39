"""Analytic synthetic rolling-shutter flow (SURVEY.md section 8(d) workloads).

The reference's example data (examples/*.tar.gz) and its MATLAB generator's renderer are absent,
so the benchmark / parity inputs are generated analytically from the same motion model the
reference inverts (matlab_synthetic_data/calculate_camera_trajectory.m:15-26, report eq. 6-7):

    u = beta(k, alpha, alpha_k) * (A v d + B w)          (normalised, gamma-scaled flow)

with alpha / alpha_k depending on the pixel-unit vertical flow itself (minimal.cc:179-197); the
implicit equation is solved exactly (closed form for k = 0, fixed point for k != 0).  Host-side
numpy only; this is input generation, not part of the measured path.
"""
import numpy as np

# Camera::setIntrinsics hard-coded calibrations (camera.cc:179-206): fx, fy, cx, cy
INTRINSICS = {
    "iphone": (1505.1283359786307, 1513.7789208311444, 657.81734686405991, 349.91807538147589),
    "galaxy_stabil": (1803.29785922382, 1799.35406531529, 945.304708272490, 544.684292978344),
    "galaxy": (1492.41306997746, 1491.09286590722, 949.571146410704, 554.675409391795),
    "galaxy_old": (3154.53208221173, 3152.28696217577, 1969.87107268891, 1521.27056048818),
    "galaxy_vga": (484.450845764569, 485.345469134313, 313.442094604855, 241.383116350144),
}


def piecewise_planar_inverse_depth(rows, cols, K4, rng, ncells=64, z_range=(2.0, 30.0)):
    """Voronoi cells, each a plane in inverse depth d = a x + b y + c, Z in z_range."""
    fx, fy, cx, cy = K4
    sx = rng.uniform(0, cols, ncells)
    sy = rng.uniform(0, rows, ncells)
    dmin, dmax = 1.0 / z_range[1], 1.0 / z_range[0]
    c0 = rng.uniform(dmin * 1.5, dmax * 0.8, ncells)
    a = rng.uniform(-0.15, 0.15, ncells) * c0
    b = rng.uniform(-0.15, 0.15, ncells) * c0
    jj, ii = np.meshgrid(np.arange(rows, dtype=np.float64), np.arange(cols, dtype=np.float64), indexing="ij")
    cell = np.zeros((rows, cols), dtype=np.int32)
    best = np.full((rows, cols), np.inf)
    for c in range(ncells):
        dist = (ii - sx[c]) ** 2 + (jj - sy[c]) ** 2
        m = dist < best
        best[m] = dist[m]
        cell[m] = c
    x = (ii - cx) / fx
    y = (jj - cy) / fy
    xs = (sx - cx) / fx
    ys = (sy - cy) / fy
    d = a[cell] * (x - xs[cell]) + b[cell] * (y - ys[cell]) + c0[cell]
    return np.clip(d, dmin, dmax)


def rs_flow(rows, cols, K4, gamma, v, w, k, inv_depth):
    """Pixel-unit flow image (rows, cols, 2) solving the RS differential model exactly."""
    fx, fy, cx, cy = K4
    v = np.asarray(v, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    jj, ii = np.meshgrid(np.arange(rows, dtype=np.float64), np.arange(cols, dtype=np.float64), indexing="ij")
    x = (ii - cx) / fx
    y = (jj - cy) / fy
    d = inv_depth
    gx = (v[0] - x * v[2]) * d + (-x * y * w[0] + (1 + x * x) * w[1] - y * w[2])
    gy = (v[1] - y * v[2]) * d + (-(1 + y * y) * w[0] + x * y * w[1] + x * w[2])
    h = float(rows)
    if k == 0.0:
        uy = gy / (1.0 - gy * fy / h)
        beta = 1.0 + uy * fy / h
    else:
        uy = gy.copy()
        for _ in range(60):
            dy = uy * fy / gamma
            alpha = 1.0 + gamma * dy / h
            p1 = gamma * jj / h
            p2 = 1.0 + gamma * (jj + dy) / h
            alpha_k = 0.5 * (p2 * p2 - p1 * p1)
            beta = (2.0 / (2.0 + k)) * (alpha + k * alpha_k)
            uy = beta * gy
    ux = beta * gx
    flow = np.empty((rows, cols, 2), dtype=np.float64)
    flow[..., 0] = ux * fx / gamma
    flow[..., 1] = uy * fy / gamma
    return flow


def rs_image(rows, cols, rng, void_frac=0.0005, dark_frac=0.01):
    """Synthetic BGR rolling-shutter frame: smooth texture + noise, a few renderer-void pixels
    BGR(1,1,1) (skipped by RsFrame::backProject, rsframe.cc:815) and a few dark (<= 15 norm)
    pixels that exercise the crack-fill blackness test (camera.cc:694-709)."""
    jj, ii = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    img = np.empty((rows, cols, 3), dtype=np.uint8)
    img[..., 0] = (96 + 80 * np.sin(ii * 0.031) * np.cos(jj * 0.017)).astype(np.uint8)
    img[..., 1] = (128 + 90 * np.sin((ii + jj) * 0.011)).astype(np.uint8)
    img[..., 2] = ((ii * 3 + jj * 5) % 200 + 30).astype(np.uint8)
    img = (img.astype(np.int32) + rng.integers(-8, 9, size=img.shape)).clip(16, 255).astype(np.uint8)
    nd = int(dark_frac * rows * cols)
    if nd:
        img[rng.integers(0, rows, nd), rng.integers(0, cols, nd)] = rng.integers(0, 12, size=(nd, 3), dtype=np.uint8)
    nv = int(void_frac * rows * cols)
    if nv:
        img[rng.integers(0, rows, nv), rng.integers(0, cols, nv)] = 1
    return img


def make_pair(rows=1080, cols=1920, intrinsics="galaxy_stabil", gamma=0.95, v=(0.30, 0.05, 0.02),
              w=(0.002, -0.004, 0.0087), k=0.0, seed=1, noise_sigma_px=0.0, outlier_frac=0.0,
              zero_flow_frac=0.0, ncells=64, z_range=(2.0, 30.0), flow_f32=False):
    """One synthetic RS frame pair: dict(flow_img, image, inv_depth, K4, gamma, v, w, k)."""
    K4 = INTRINSICS[intrinsics] if isinstance(intrinsics, str) else tuple(intrinsics)
    rng = np.random.default_rng(seed)
    d = piecewise_planar_inverse_depth(rows, cols, K4, rng, ncells, z_range)
    flow = rs_flow(rows, cols, K4, gamma, v, w, k, d)
    rng2 = np.random.default_rng(seed + 1)
    if noise_sigma_px > 0:
        flow += rng2.normal(0.0, noise_sigma_px, size=flow.shape)
    if outlier_frac > 0:
        mask = rng2.random((rows, cols)) < outlier_frac
        flow[mask] = rng2.uniform(-20, 20, size=(int(mask.sum()), 2))
    if zero_flow_frac > 0:
        # background blobs with flow exactly 0 (dropped by the 1e-10 test, camera.cc:232-234)
        jj, ii = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
        blob = (np.sin(ii * 0.02 + seed) * np.cos(jj * 0.027 - seed)) > np.cos(np.pi * zero_flow_frac)
        flow[blob] = 0.0
    if flow_f32:
        flow = flow.astype(np.float32).astype(np.float64)   # DeepFlow output is float32 (camera.cc:262-274)
    img = rs_image(rows, cols, np.random.default_rng(seed + 2))
    return dict(flow_img=flow, image=img, inv_depth=d, K4=np.array(K4), gamma=float(gamma),
                v=np.array(v, dtype=np.float64), w=np.array(w, dtype=np.float64), k=float(k),
                rows=rows, cols=cols)


def ground_truth(P, world_R=None, world_t=None, void_frac=0.0, seed=11, frame=1):
    """Ground-truth fixtures of the RS frame of a synthetic pair, in the layout the reference loads from
    disk (rsframe.cc:222-378): per-pixel world points `unproj` = (X, Y, Z) maps (rows x cols each) and
    per-scanline camera-from-world poses R_gt (rows x 3 x 3), t_gt (rows x 3).  The scanline poses follow
    the small-motion model of RsFrame::setRelativePose with the true motion; (world_R, world_t) places the
    world frame away from scanline 0 so that relocatePose has something to do.  void_frac: fraction of
    pixels without a world point (all-zero entries, as for background pixels of the renderer).
    frame=2: the poses of the second frame's scanlines (read-out starts one frame period later); its
    `unproj`/`depth` entries still describe frame 1 and are not meaningful."""
    rows, cols = P["rows"], P["cols"]
    fx, fy, cx, cy = [float(a) for a in P["K4"]]
    v, w, k, gamma = P["v"], P["w"], P["k"], P["gamma"]
    i = np.arange(rows, dtype=np.float64)
    tau = (frame - 1) + gamma * i / rows                      # time in frame periods since the first scanline
    beta = (tau + 0.5 * k * tau * tau) * (2.0 / (2.0 + k))
    skew = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], dtype=np.float64)
    R = np.eye(3)[None] + beta[:, None, None] * skew[None]
    t = beta[:, None] * np.asarray(v, dtype=np.float64)[None]
    if world_R is not None:
        wt = np.zeros(3) if world_t is None else np.asarray(world_t, dtype=np.float64)
        t = np.einsum("rij,j->ri", R, wt) + t
        R = R @ np.asarray(world_R, dtype=np.float64)
    Z = 1.0 / P["inv_depth"]
    xn = (np.arange(cols) - cx) / fx
    yn = (np.arange(rows) - cy) / fy
    Xc = np.stack([Z * xn[None, :], Z * yn[:, None], Z], axis=-1)
    Xw = np.einsum("rij,rcj->rci", np.linalg.inv(R), Xc - t[:, None, :])
    if void_frac > 0:
        hole = np.random.default_rng(seed).random((rows, cols)) < void_frac
        Xw[hole] = 0.0
    return dict(unproj=(np.ascontiguousarray(Xw[..., 0]), np.ascontiguousarray(Xw[..., 1]), np.ascontiguousarray(Xw[..., 2])),
                R_gt=R, t_gt=t, depth=Z)


def sample_list(n, H, seed=3):
    """H x 9 distinct point indices (replaces srand(time)/rand of minimal.cc:230-244)."""
    rng = np.random.default_rng(seed)
    out = np.empty((H, 9), dtype=np.int32)
    for h in range(H):
        out[h] = rng.choice(n, size=9, replace=False)
    return out


def make_pair_device(torch, device, rows=1080, cols=1920, intrinsics="galaxy_stabil", gamma=0.95, v=(0.30, 0.05, 0.02),
                     w=(0.002, -0.004, 0.0087), k=0.0, seed=1, noise_sigma_px=0.0, outlier_frac=0.0, ncells=64,
                     z_range=(2.0, 30.0)):
    """make_pair with the per-pixel work done by torch on `device` (milliseconds instead of seconds at 1080p),
    for benchmarks that need many DISTINCT pairs.  Same scene model and the same exact solution of the RS
    differential equations; the random streams differ from make_pair's (torch generators), so a pair is
    reproducible from its seed on a given device type but is not the numpy pair of the same seed.
    Returns dict(flow_img (rows, cols, 2) float64 tensor, image (rows, cols, 3) uint8 tensor, K4, gamma, v, w, k)."""
    K4 = INTRINSICS[intrinsics] if isinstance(intrinsics, str) else tuple(intrinsics)
    fx, fy, cx, cy = K4
    rng = np.random.default_rng(seed)
    sx = rng.uniform(0, cols, ncells); sy = rng.uniform(0, rows, ncells)
    dmin, dmax = 1.0 / z_range[1], 1.0 / z_range[0]
    c0 = rng.uniform(dmin * 1.5, dmax * 0.8, ncells)
    a = rng.uniform(-0.15, 0.15, ncells) * c0
    b = rng.uniform(-0.15, 0.15, ncells) * c0
    f64 = dict(dtype=torch.float64, device=device)
    T = lambda arr: torch.as_tensor(np.asarray(arr, dtype=np.float64), **f64)
    jj, ii = torch.meshgrid(torch.arange(rows, **f64), torch.arange(cols, **f64), indexing="ij")
    # nearest Voronoi site, in chunks of sites (rows x cols x 8 doubles at a time)
    best = torch.full((rows, cols), float("inf"), **f64)
    cell = torch.zeros((rows, cols), dtype=torch.int64, device=device)
    sxt, syt = T(sx), T(sy)
    for c in range(0, ncells, 8):
        dist = (ii[..., None] - sxt[c:c + 8]) ** 2 + (jj[..., None] - syt[c:c + 8]) ** 2
        dm, am = dist.min(dim=2)
        upd = dm < best
        best = torch.where(upd, dm, best); cell = torch.where(upd, am + c, cell)
    x = (ii - cx) / fx; y = (jj - cy) / fy
    xs, ys = (sxt - cx) / fx, (syt - cy) / fy
    d = (T(a)[cell] * (x - xs[cell]) + T(b)[cell] * (y - ys[cell]) + T(c0)[cell]).clamp(dmin, dmax)
    v = np.asarray(v, dtype=np.float64); w = np.asarray(w, dtype=np.float64)
    gx = (v[0] - x * v[2]) * d + (-x * y * w[0] + (1 + x * x) * w[1] - y * w[2])
    gy = (v[1] - y * v[2]) * d + (-(1 + y * y) * w[0] + x * y * w[1] + x * w[2])
    h = float(rows)
    if k == 0.0:
        uy = gy / (1.0 - gy * fy / h)
        beta = 1.0 + uy * fy / h
    else:
        uy = gy.clone()
        for _ in range(60):
            dy = uy * fy / gamma
            alpha = 1.0 + gamma * dy / h
            p1 = gamma * jj / h
            p2 = 1.0 + gamma * (jj + dy) / h
            beta = (2.0 / (2.0 + k)) * (alpha + k * 0.5 * (p2 * p2 - p1 * p1))
            uy = beta * gy
    flow = torch.stack((beta * gx * fx / gamma, uy * fy / gamma), dim=2)
    g = torch.Generator(device=device); g.manual_seed(int(seed) + 1)
    if noise_sigma_px > 0:
        flow = flow + noise_sigma_px * torch.randn(flow.shape, generator=g, **f64)
    if outlier_frac > 0:
        mask = torch.rand((rows, cols), generator=g, device=device) < outlier_frac
        rnd = torch.rand((rows, cols, 2), generator=g, **f64) * 40.0 - 20.0
        flow = torch.where(mask[..., None], rnd, flow)
    i32 = dict(dtype=torch.int32, device=device)
    ji, iic = torch.meshgrid(torch.arange(rows, **i32), torch.arange(cols, **i32), indexing="ij")
    img = torch.stack(((96 + 80 * torch.sin(iic * 0.031) * torch.cos(ji * 0.017)).to(torch.int32),
                       (128 + 90 * torch.sin((iic + ji) * 0.011)).to(torch.int32),
                       ((iic * 3 + ji * 5) % 200 + 30)), dim=2)
    img = (img + torch.randint(-8, 9, img.shape, generator=g, **i32)).clamp(16, 255).to(torch.uint8)
    nd = int(0.01 * rows * cols); nv = int(0.0005 * rows * cols)
    pick = lambda n: (torch.randint(0, rows, (n,), generator=g, device=device), torch.randint(0, cols, (n,), generator=g, device=device))
    r_, c_ = pick(nd); img[r_, c_] = torch.randint(0, 12, (nd, 3), generator=g, device=device).to(torch.uint8)
    r_, c_ = pick(nv); img[r_, c_] = 1
    return dict(flow_img=flow.contiguous(), image=img.contiguous(), K4=np.array(K4), gamma=float(gamma), v=v, w=w, k=float(k),
                rows=rows, cols=cols)
