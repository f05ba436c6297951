"""GPU parity against the committed golden fixtures (tests/golden/*.npz, frozen oracle outputs): the CUDA path through the
C ABI reproduces them without the oracle in the loop."""
import os

import numpy as np
import pytest

from cfear_radarodometry_code_public_b200 import capi, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_gpu_reproduces_golden_v1():
    g = np.load(os.path.join(GOLD, "cfear_golden_v1.npz"))
    img = synth.make_problem_images(int(g["seed"]), 1)[0]
    c = capi.Context(max_batch=2, max_cellsets=4, max_keyframes=1, cost="P2L", radius=3.5)
    idx, cnt = c.kstrongest(img[1][None])
    assert np.array_equal(idx[0], g["kidx"]) and np.array_equal(cnt[0], g["kcnt"])            # bit-exact index sets
    out = c.filter(img[:2])
    assert np.array_equal(out["clouds"][1], g["cloud"])                                        # bit-exact fp32 cloud
    n0 = c.surface_points(out["clouds"][0], 0)
    n1 = c.surface_points(out["clouds"][1], 1)
    cells = c.cells_download(1)
    assert n1 == g["mean"].shape[0] and n0 > 0
    assert np.array_equal(cells["nsamples"], g["nsamples"])
    np.testing.assert_allclose(cells["mean"], g["mean"], atol=1e-9)
    np.testing.assert_allclose(cells["normal"], g["normal"], atol=1e-7)
    p, _, st, _ = c.register_batch(np.array([[0, 1]], np.int32), g["poses_in"][None])
    assert st["outer_iterations"][0] == int(g["outer"]) and st["inner_iterations"][0] == int(g["inner"])
    d = p[0, 1] - g["poses_out"][1]
    assert np.hypot(d[0], d[1]) < 1e-4 and abs(d[2]) < 1e-5
    c.close()


def test_gpu_reproduces_golden_v2():
    g = np.load(os.path.join(GOLD, "cfear_golden_v2.npz"))
    K = 2
    img = synth.make_problem_images(int(g["seed"]), K)[0]
    c = capi.Context(max_batch=32, max_cellsets=K + 1, max_keyframes=K, cost="P2D", loss="Huber", weight_opt=4, regularization=0.1,
                     radius=3.0)
    out = c.filter(img[:K + 1])
    for i in range(K + 1):
        c.surface_points(out["clouds"][i], i)
    slots = np.arange(K + 1, dtype=np.int32)[None]
    p, cov, st, _ = c.register_batch(slots, g["poses_in"][None])
    assert st["outer_iterations"][0] == int(g["outer"]) and st["inner_iterations"][0] == int(g["inner"])
    assert st["num_residuals"][0] == int(g["num_residuals"])
    d = p[0, K] - g["poses_out"][K]
    assert np.hypot(d[0], d[1]) < 1e-4 and abs(d[2]) < 1e-5
    np.testing.assert_allclose(st["final_cost"][0], float(g["final_cost"]), rtol=1e-6)
    np.testing.assert_allclose(np.asarray(cov).reshape(-1, 6, 6)[0], g["cov"], rtol=1e-5, atol=1e-12)
    # GetCost on the golden pose samples (x, y, yaw offsets in g["samples"][:, :3], cost in [:, 3])
    S = g["samples"]
    poses = np.repeat(g["poses_out"][None], S.shape[0], 0)
    poses[:, K, :] += S[:, :3]
    cost, nres, ok = c.get_cost_batch(np.repeat(slots, S.shape[0], 0), poses)
    assert ok.all()
    np.testing.assert_allclose(cost, S[:, 3], rtol=1e-8)
    c.close()
