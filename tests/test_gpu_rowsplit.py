"""Row split of one LM solve over two GPUs (include/rsdsfm.h "row split"; BASELINE config 4): the members hold
contiguous shares of the residual blocks and exchange one row of sums per LM phase through peer memory.  Needs two
B200s in one box (skipped otherwise): two contexts of this process, one host thread each (ctypes releases the GIL)."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _two_gpus():
    try:
        import torch
        return torch.cuda.is_available() and torch.cuda.device_count() >= 2
    except Exception:
        return False


@pytest.mark.skipif(not _two_gpus(), reason="needs two GPUs")
@pytest.mark.parametrize("const_acc,foe", [(True, False), (False, False), (True, True)])
def test_row_split_equals_single_gpu(capi, oracle, synth, const_acc, foe):
    """Same termination, same iteration count, motion and depths to rounding (the sums are grouped differently);
    foe: the focus of expansion lies inside the image, so both members list clamped pixels (their sums cross the
    GPUs too)."""
    import helpers
    rows, cols = 135, 240
    K4 = helpers.small_K(8)
    v = (0.02, 0.01, 0.30) if foe else (0.30, 0.05, 0.02)
    P = synth.make_pair(rows, cols, K4, gamma=0.95, v=v, w=(0.002, -0.004, 0.0087), k=0.5 if const_acc else 0.0, seed=77,
                        noise_sigma_px=0.1, outlier_frac=0.03)
    n, coord, flow, cpx, fpx = oracle.flatten(P["flow_img"], K4, 0.95)
    alpha = oracle.get_alpha(fpx, n, rows, 0.95)
    alpha_k = oracle.get_alpha_k(cpx, fpx, n, rows, 0.95)
    c0 = capi.Context(0)
    c1 = capi.Context(1)
    R = c0.ransac(coord[:2 * n], flow[:2 * n], alpha, alpha_k, n, const_acc, synth.sample_list(n, 6, seed=5), 0.02)
    inl, a_in, ak_in, ix, m = c0.gather_inliers(coord[:2 * n], alpha, alpha_k, n, R["mask"], R["inv_depth"])
    fl = flow[:2 * m]
    v1, w1, k1, z1, S1 = c0.refine(fl, inl, a_in, ak_in, m, R["v"], R["w"], R["k"], const_acc)
    capi.Context.peer_connect_local([c0, c1])
    cut = (m // 2 + 255) // 256 * 256
    shares = [(0, cut), (cut, m)]
    out = [None, None]

    def work(g, c):
        lo, hi = shares[g]
        out[g] = c.refine(fl[2 * lo:2 * hi], inl[3 * lo:3 * hi], a_in[lo:hi], ak_in[lo:hi], hi - lo, R["v"], R["w"], R["k"], const_acc)

    for rep in range(2):                      # twice: the mailbox slots and tags are reused from solve to solve
        th = [threading.Thread(target=work, args=(g, c)) for g, c in enumerate((c0, c1))]
        for t in th:
            t.start()
        for t in th:
            t.join()
        for g in range(2):
            vN, wN, kN, zN, SN = out[g]
            lo, hi = shares[g]
            assert (SN["termination"], SN["reason"], SN["iterations"]) == (S1["termination"], S1["reason"], S1["iterations"])
            np.testing.assert_allclose(vN, v1, rtol=1e-9, atol=1e-13)
            np.testing.assert_allclose(wN, w1, rtol=1e-9, atol=1e-13)
            assert abs(kN - k1) <= 1e-9 * max(abs(k1), 1e-3)
            np.testing.assert_allclose(zN, z1[lo:hi], rtol=1e-7, atol=0)
        assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1]) and out[0][2] == out[1][2]
    c0.peer_disconnect(); c1.peer_disconnect()
    v2, w2, k2, z2, S2 = c0.refine(fl, inl, a_in, ak_in, m, R["v"], R["w"], R["k"], const_acc)      # back to one GPU
    assert np.array_equal(v2, v1) and np.array_equal(z2, z1)
    c0.close(); c1.close()
