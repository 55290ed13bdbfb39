"""Shared case builders for the parity tests (oracle side only; no GPU)."""
import numpy as np


def small_K(scale):
    from importlib import import_module
    synth = import_module("rs-aware-differential-sfm_b200.synth")
    return tuple(np.array(synth.INTRINSICS["galaxy_stabil"]) / scale)


def make_case(O, synth, rows, cols, K4, k=0.0, const_acc=False, seed=1, noise=0.1, outliers=0.05, zero_frac=0.0,
              H=12, tol=0.01, gamma=0.95, flow_f32=False, sample_seed=3, v=(0.30, 0.05, 0.02),
              w=(0.002, -0.004, 0.0087)):
    """Synthetic pair -> flatten -> alpha -> RANSAC (oracle) -> consensus set.  Everything a
    refinement / rectification parity test needs, computed by the CPU oracle."""
    P = synth.make_pair(rows, cols, K4, gamma=gamma, seed=seed, k=k, noise_sigma_px=noise, outlier_frac=outliers,
                        zero_flow_frac=zero_frac, flow_f32=flow_f32, v=v, w=w)
    n, coord, flow, cpx, fpx = O.flatten(P["flow_img"], K4, gamma)
    alpha = O.get_alpha(fpx, n, rows, gamma)
    alpha_k = O.get_alpha_k(cpx, fpx, n, rows, gamma)
    samples = synth.sample_list(n, H, seed=sample_seed)
    R = O.ransac(coord[:2 * n], flow[:2 * n], alpha, alpha_k, n, const_acc, tol, samples=samples)
    inl, a_in, ak_in = O.gather_inliers(coord, alpha, alpha_k, n, R["mask"], R["inv_depth"])
    return dict(P=P, n=n, coord=coord[:2 * n].copy(), flow=flow[:2 * n].copy(), coord_px=cpx[:2 * n].copy(),
                flow_px=fpx[:2 * n].copy(), alpha=alpha, alpha_k=alpha_k, samples=samples, ransac=R, inliers3=inl,
                alpha_in=a_in, alpha_k_in=ak_in, m=len(a_in), K4=np.array(K4), gamma=gamma, rows=rows, cols=cols,
                const_acc=const_acc, tol=tol)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
