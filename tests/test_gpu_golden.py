"""GPU parity against the committed golden fixtures (tests/golden/*.npz, frozen oracle outputs): the CUDA path through the
C ABI reproduces them without the oracle in the loop."""
import os

import numpy as np
import pytest

from cfear_radarodometry_code_public_b200 import capi, synth

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("kernel_form")]
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_gpu_reproduces_golden_v1():
    g = np.load(os.path.join(GOLD, "cfear_golden_v1.npz"))
    img = synth.make_problem_images(int(g["seed"]), 1)[0]
    assert np.array_equal(img[1, ::8, ::8], g["img_sub"]), "the synthetic generator drifted (numpy RNG / rendering): regenerate the goldens"
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
    assert np.array_equal(img[K, ::8, ::8], g["img_sub"]), "the synthetic generator drifted (numpy RNG / rendering): regenerate the goldens"
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
    # covariance by sampling (odometrykeyframefuser.cpp:261-380) from the GPU's costs: quadric fit -> 2 H^-1 * scaler
    x, y, z = S[:, 0], S[:, 1], S[:, 2]
    A = np.stack([x * x, y * y, z * z, x * y, y * z, z * x, x, y, z, np.ones_like(x)], 1)
    q = np.linalg.lstsq(A, cost, rcond=None)[0]
    H = np.array([[2 * q[0], q[3], q[5]], [q[3], 2 * q[1], q[4]], [q[5], q[4], 2 * q[2]]])
    c3 = 2.0 * np.linalg.inv(H) * (st["final_cost"][0] / (st["num_residuals"][0] - 3)) * 4.0
    gs = g["sampled_cov"]
    np.testing.assert_allclose(c3[:2, :2], gs[:2, :2], rtol=1e-4)
    np.testing.assert_allclose([c3[2, 2], c3[0, 2], c3[1, 2]], [gs[5, 5], gs[0, 5], gs[1, 5]], rtol=1e-4, atol=1e-12)
    c.close()


def test_gpu_reproduces_golden_v2_sequence_replay():
    """The frozen OdometryKeyframeFuser replay (8 scans, P2L, window 3) through the device-side fuser (cfear_seq_*)."""
    g = np.load(os.path.join(GOLD, "cfear_golden_v2.npz"))
    seq, _ = synth.make_sequence(int(g["seq_seed"]), 8)
    assert np.array_equal(seq[-1, ::8, ::8], g["seq_sub"]), "the synthetic generator drifted (numpy RNG / rendering): regenerate the goldens"
    c = capi.Context(max_batch=1, max_cellsets=5, max_keyframes=4, cost="P2L", weight_opt=0, radius=3.5, weight_intensity=1)
    S = capi.Sequences(c, 1, 8, submap_scan_size=3)
    for t in range(8):
        S.step(seq[t][None])
    poses, kf, _ = S.read(0, 8)
    assert np.array_equal(kf[0], g["seq_keyframe"])
    d = poses[0] - g["seq_poses"]
    assert np.hypot(d[:, 0], d[:, 1]).max() < 1e-4 and np.abs(d[:, 2]).max() < 1e-5
    S.close(); c.close()
