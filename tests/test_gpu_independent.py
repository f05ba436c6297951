"""The CUDA path against the SECOND writing of the reference path -- the numpy / FLANN restatement of
tests/test_surface_independent.py, test_register_independent.py and test_fuser_independent.py -- with the C++ oracle not in the
loop at all: polar image -> cfear_filter -> cfear_compensate -> cfear_surface_points -> cfear_register through the C ABI on one
side, orc-free numpy on the other (the filter rows themselves are pinned to the reference source in test_ref_pin.py).
Same bars as everywhere: counts exact, statistics to 1e-9, the same outer / inner iteration counts, poses within
1e-4 m / 1e-5 rad (they agree to ~1e-9).
"""
import numpy as np
import pytest

from cfear_radarodometry_code_public_b200 import capi, synth
import helpers
from test_register_independent import register_py
from test_surface_independent import cells_np, compensate_np

pytest.importorskip("cv2")
pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("kernel_form")]
POS_TOL, ROT_TOL = 1e-4, 1e-5


@pytest.mark.parametrize("cost,wopt,reg,seed,K", [("P2D", 4, 0.1, 31, 3), ("P2L", 0, 1.0, 32, 2), ("P2P", 2, 1.0, 33, 1)])
def test_cuda_path_matches_the_numpy_flann_restatement(cost, wopt, reg, seed, K):
    radius = 3.0
    imgs, tp = helpers.scan_images(seed, K)
    c = capi.Context(max_batch=K + 1, max_cellsets=K + 1, max_keyframes=K, radius=radius, cost=cost, loss="Huber", loss_limit=0.1,
                     weight_opt=wopt, regularization=reg, cov_scale=1.0, weight_intensity=1)
    mot = synth.se2_mul(synth.se2_inv(tp[K - 1]), tp[K]) if K >= 1 else np.zeros(3)
    f = c.filter(imgs)
    sets = []
    for i in range(K + 1):
        cl = f["clouds"][i]
        if i == K:                                                        # the current scan is motion-compensated
            mine = compensate_np(cl, mot)
            cl = c.compensate(cl, mot)
            ulp = np.abs(mine[:, :2].view(np.int32).astype(np.int64) - cl[:, :2].view(np.int32).astype(np.int64))
            assert ulp.max() <= 1 and (ulp > 0).mean() < 0.01
        n = c.surface_points(cl, i)
        got = c.cells_download(i)
        exp = cells_np(cl, radius, True)
        assert n == exp["mean"].shape[0] > 100
        assert np.array_equal(got["nsamples"], exp["nsamples"])
        np.testing.assert_allclose(got["mean"], exp["mean"], rtol=0, atol=1e-9)
        np.testing.assert_allclose(got["cov"], exp["cov"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(got["normal"], exp["normal"], atol=1e-7)
        np.testing.assert_allclose(got["planarity"], exp["planarity"], rtol=1e-8)
        sets.append(exp)
    P = tp[:K + 1].copy(); P[K] = tp[K] + np.array([0.6, -0.4, 0.02])
    gp, gcov, gst = c.register(np.arange(K + 1, dtype=np.int32), P)
    ok, x, itr, inner, nres, fc, cov = register_py(sets, P, cost, wopt, reg=reg)
    c.close()
    assert ok and gst["success"] == 1
    assert (gst["outer_iterations"], gst["inner_iterations"], gst["num_residuals"]) == (itr, inner, nres)
    d = gp[K] - x
    assert np.hypot(d[0], d[1]) < POS_TOL and abs(d[2]) < ROT_TOL
    assert np.hypot(d[0], d[1]) < 1e-8                                     # in fact
    np.testing.assert_allclose(gst["final_cost"], fc, rtol=1e-8)
    np.testing.assert_allclose(gcov, cov, rtol=1e-5, atol=1e-12)
