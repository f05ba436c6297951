"""CPU-side tests (no GPU): the oracle against independent definitions and the committed golden vectors, the C-ABI
library's exported symbols, host-side argument handling, and the synthetic generator's determinism."""
import ctypes
import os
import re

import numpy as np
import pytest

import helpers
from cfear_radarodometry_code_public_b200 import capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


# ---- oracle vs independent definitions ----------------------------------------------------------------------
def _kstrongest_def(img, zmin, k):
    """the k lexicographically largest (intensity, range) pairs with intensity >= z_min, ascending."""
    A, R = img.shape
    idx = np.full((A, k), -1, np.int32); cnt = np.zeros(A, np.int32)
    for a in range(A):
        cand = [(int(v), r) for r, v in enumerate(img[a]) if v >= zmin]
        keep = sorted(cand)[-k:] if cand else []
        cnt[a] = len(keep)
        idx[a, :len(keep)] = [r for _, r in keep]
    return idx, cnt


@pytest.mark.parametrize("k,zmin", [(12, 60), (1, 60), (40, 0), (5, 255)])
def test_oracle_kstrongest_vs_definition(orc, k, zmin):
    img = helpers.adversarial_image(3)[:64, :600].copy()
    oi, oc = orc.kstrongest(img, zmin, k)
    di, dc = _kstrongest_def(img, zmin, k)
    assert np.array_equal(oc, dc) and np.array_equal(oi, di)


def test_oracle_cloud_definition(orc):
    img = helpers.adversarial_image(3)
    oi, oc = orc.kstrongest(img, 60, 12)
    cl = orc.cloud(img, oi, oc)
    A = img.shape[0]
    res = float(np.float32(0.0438)); mrb = int(np.ceil(float(np.float32(2.5)) / res))
    pts = []
    for a in range(A):
        th = ((a + 1) / A) * 2.0 * np.pi
        for j in range(oc[a]):
            r = int(oi[a, j])
            if r > mrb:
                rho = res / 2.0 + res * r
                pts.append((np.float32(rho * np.cos(th)), np.float32(rho * np.sin(th)), 0.0, float(img[a, r])))
    ref = np.array(pts, np.float32)
    assert cl.shape == ref.shape and mrb == 58
    assert np.array_equal(cl, ref)


def test_oracle_cells_vs_numpy_and_ckdtree(orc):
    from scipy.spatial import cKDTree
    im, _ = helpers.scan_images(3, 0)
    cl, sp = helpers.oracle_cells(orc, im[0], radius=3.5)
    cx, cy, ci, vid, dims = orc.voxel_centroids(cl, 3.5)
    tree = cKDTree(cl[:, :2].astype(np.float64))
    nc = 0
    for c in range(cx.shape[0]):
        q = np.array([cx[c], cy[c]], np.float32)
        d2 = ((q[None, :] - cl[:, :2]) ** 2).astype(np.float32)
        nb = np.nonzero((d2[:, 0] + d2[:, 1]).astype(np.float32) < np.float32(3.5 * 3.5))[0]   # fp32 L2_Simple
        assert set(nb) == set(i for i in tree.query_ball_point(q.astype(np.float64), 3.5 + 1e-3) if i in set(nb))
        if nb.size < 6:
            continue
        w = np.maximum(cl[nb, 3].astype(np.float64) - 60.0, 0.0)
        if w.sum() == 0:
            continue
        wn = w / w.sum()
        mu = (wn[:, None] * cl[nb, :2]).sum(0)
        xc = cl[nb, :2] - mu
        cov = xc.T @ (wn[:, None] * xc)
        lam, vec = np.linalg.eigh(cov)
        cond = abs(lam[1] / lam[0])
        if not (cond <= 1e4 and lam[0] * lam[1] > 1e-5 and lam[0] > 0 and lam[1] > 0):
            continue
        assert sp["nsamples"][nc] == nb.size
        np.testing.assert_allclose(sp["mean"][nc], mu, atol=1e-10)
        np.testing.assert_allclose(sp["cov"][nc], cov, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(sp["lambdas"][nc], lam, rtol=1e-8)
        n = vec[:, 0] if vec[:, 0] @ (-mu) >= 0 else -vec[:, 0]
        np.testing.assert_allclose(sp["normal"][nc], n, atol=1e-7)
        np.testing.assert_allclose(sp["planarity"][nc], np.log(1 + cond / 2), rtol=1e-8)
        nc += 1
    assert nc == sp["mean"].shape[0] > 100


def test_oracle_nearest_vs_ckdtree(orc):
    from scipy.spatial import cKDTree
    rng = np.random.Generator(np.random.PCG64(2))
    means = rng.uniform(-100, 100, (700, 2))
    q = rng.uniform(-110, 110, (3000, 2))
    got = orc.nearest(means, q, 2.0)
    m32 = means.astype(np.float32).astype(np.float64); q32 = q.astype(np.float32).astype(np.float64)
    d, i = cKDTree(m32).query(q32, k=1)
    exp = np.where(d * d < 4.0 - 1e-4, i, -1)
    sure = np.abs(d * d - 4.0) > 1e-4                      # away from the fp32 acceptance boundary
    assert np.array_equal(got[sure], exp[sure])


def _residuals(rng, n, cost):
    res = np.zeros((n, 8))
    res[:, 0:2] = rng.uniform(-50, 50, (n, 2)); res[:, 2:4] = res[:, 0:2] + rng.normal(0, 0.3, (n, 2))
    th = rng.uniform(0, 2 * np.pi, n)
    if cost == "P2L":
        res[:, 4], res[:, 5] = np.cos(th), np.sin(th)
    elif cost == "P2D":
        res[:, 4] = rng.uniform(0.5, 3, n); res[:, 5] = rng.normal(0, 0.5, n); res[:, 6] = rng.uniform(0.5, 3, n)
    res[:, 7] = rng.uniform(0.5, 3, n)
    return res


@pytest.mark.parametrize("cost", ["P2L", "P2D", "P2P"])
@pytest.mark.parametrize("loss", ["None", "Huber", "Cauchy", "SoftLOne", "Combined", "Tukey"])
def test_oracle_cost_gradient_and_gauss_newton(orc, cost, loss):
    """J^T r and J^T J of the oracle against numerical differentiation of an independent numpy cost."""
    rng = np.random.Generator(np.random.PCG64(5))
    res = _residuals(rng, 200, cost)
    cfg = orc.reg_cfg(cost=cost, loss=loss, loss_limit=0.1)
    a = 0.1

    def rho(s):
        if loss == "None":
            return s
        if loss == "Huber":
            return np.where(s > a * a, 2 * a * np.sqrt(s) - a * a, s)
        if loss == "Cauchy":
            return a * a * np.log1p(s / (a * a))
        if loss == "SoftLOne":
            return 2 * a * a * (np.sqrt(1 + s / (a * a)) - 1)
        if loss == "Combined":
            g = np.log1p(s)
            return np.where(g > 1, 2 * np.sqrt(g) - 1, g)
        return np.where(s <= a * a, a * a / 3 * (1 - (1 - s / (a * a)) ** 3), a * a / 3)

    def cost_np(x):
        c, s_ = np.cos(x[2]), np.sin(x[2])
        ex = c * res[:, 0] - s_ * res[:, 1] + x[0] - res[:, 2]
        ey = s_ * res[:, 0] + c * res[:, 1] + x[1] - res[:, 3]
        if cost == "P2L":
            s = (ex * res[:, 4] + ey * res[:, 5]) ** 2
        elif cost == "P2D":
            s = (res[:, 4] * ex) ** 2 + (res[:, 5] * ex + res[:, 6] * ey) ** 2
        else:
            s = ex ** 2 + ey ** 2
        return 0.5 * (res[:, 7] * rho(s)).sum()

    x = np.array([0.05, -0.02, 0.01])
    c, H, g = orc.eval_cost(cfg, res, x)
    np.testing.assert_allclose(c, cost_np(x), rtol=1e-12)
    num = np.array([(cost_np(x + h) - cost_np(x - h)) / 2e-6 for h in np.eye(3) * 1e-6])
    np.testing.assert_allclose(g, num, rtol=2e-4, atol=1e-6)


def test_oracle_lm_reaches_scipy_optimum(orc):
    """The restated Ceres loop must land on the same Huber optimum as scipy.optimize.least_squares."""
    from scipy.optimize import least_squares
    im, tp = helpers.scan_images(3, 1)
    sets = [helpers.oracle_cells(orc, im[i], radius=3.0)[1] for i in range(2)]
    P = tp[:2].copy(); P[1] = tp[1] + [0.15, -0.1, 0.01]
    cfg = orc.reg_cfg(cost="P2L", loss="Huber", max_outer=1, max_inner=50)
    ok, op, _, st, assoc = orc.register(sets, P, cfg, want_assoc=True)
    assert ok
    j = np.nonzero(assoc[0] >= 0)[0]; m = assoc[0][j]
    p = sets[1]["mean"][j]; q = sets[0]["mean"][m]; n = sets[0]["normal"][m]

    def fun(x):
        c, s = np.cos(x[2]), np.sin(x[2])
        ex = c * p[:, 0] - s * p[:, 1] + x[0] - q[:, 0]; ey = s * p[:, 0] + c * p[:, 1] + x[1] - q[:, 1]
        return ex * n[:, 0] + ey * n[:, 1]
    sol = least_squares(fun, P[1], loss="huber", f_scale=0.1, xtol=1e-14, ftol=1e-14, gtol=1e-14)
    assert np.hypot(*(op[1, :2] - sol.x[:2])) < 2e-4 and abs(op[1, 2] - sol.x[2]) < 2e-5


def test_oracle_soft_prior_reaches_scipy_optimum(orc):
    """Register(..., soft_constraints=true): the prior block alpha * L * (guess - x) (no loss) joins the Huber-robustified
    P2L blocks; scipy minimises the same objective when the robust blocks are replaced by their Huber square roots."""
    from scipy.optimize import least_squares
    im, tp = helpers.scan_images(3, 1)
    sets = [helpers.oracle_cells(orc, im[i], radius=3.0)[1] for i in range(2)]
    P = tp[:2].copy(); P[1] = tp[1] + [0.15, -0.1, 0.01]
    c6 = np.eye(6); c6[0, 0] = 4.0; c6[1, 1] = 9.0; c6[0, 1] = c6[1, 0] = 1.0; c6[5, 5] = 0.04; c6[0, 5] = c6[5, 0] = 0.1
    L = orc.prior_sqrt_information(c6)
    c3 = c6[np.ix_([0, 1, 5], [0, 1, 5])]
    np.testing.assert_allclose(L @ L.T, np.linalg.inv(c3), rtol=1e-12)
    cfg = orc.reg_cfg(cost="P2L", loss="Huber", max_outer=1, max_inner=50)
    ok, op, _, st, assoc = orc.register(sets, P, cfg, prior_sqrt_info=L)
    ok0, op0, _, st0, _ = orc.register(sets, P, cfg, want_assoc=True)
    assert ok and st.num_residuals == st0.num_residuals + 3 and st.num_blocks == st0.num_blocks + 1
    j = np.nonzero(assoc[0] >= 0)[0]; m = assoc[0][j]
    p = sets[1]["mean"][j]; q = sets[0]["mean"][m]; n = sets[0]["normal"][m]
    alpha = np.sqrt(sets[1]["mean"].shape[0])

    def fun(x):
        c, s = np.cos(x[2]), np.sin(x[2])
        ex = c * p[:, 0] - s * p[:, 1] + x[0] - q[:, 0]; ey = s * p[:, 0] + c * p[:, 1] + x[1] - q[:, 1]
        r = ex * n[:, 0] + ey * n[:, 1]
        a = np.abs(r)
        hub = np.where(a <= 0.1, r, np.sign(r) * np.sqrt(np.maximum(2 * 0.1 * a - 0.01, 0)))    # sqrt(rho(r^2)) with sign
        return np.concatenate([hub, alpha * (L @ (P[1] - x))])
    sol = least_squares(fun, P[1], xtol=1e-14, ftol=1e-14, gtol=1e-14)
    assert np.hypot(*(op[1, :2] - sol.x[:2])) < 2e-4 and abs(op[1, 2] - sol.x[2]) < 2e-5
    assert np.hypot(*(op[1, :2] - op0[1, :2])) > 1e-3                    # and the prior did pull the solution toward the guess


# ---- golden vectors ------------------------------------------------------------------------------------------
def test_golden_vectors(orc):
    g = np.load(os.path.join(GOLD, "cfear_golden_v1.npz"))
    img = synth.make_problem_images(int(g["seed"]), 1)[0]
    assert np.array_equal(img[1, ::8, ::8], g["img_sub"])                      # generator is reproducible
    idx, cnt = orc.kstrongest(img[1], 60, 12)
    assert np.array_equal(idx, g["kidx"]) and np.array_equal(cnt, g["kcnt"])
    cl = orc.cloud(img[1], idx, cnt)
    assert np.array_equal(cl, g["cloud"])
    sp = orc.surface_points(cl, 3.5, True)
    assert np.array_equal(sp["nsamples"], g["nsamples"])
    np.testing.assert_allclose(sp["mean"], g["mean"], atol=1e-12)
    np.testing.assert_allclose(sp["normal"], g["normal"], atol=1e-10)
    sp0 = helpers.oracle_cells(orc, img[0])[1]
    ok, op, cov, st, _ = orc.register([sp0, sp], g["poses_in"], orc.reg_cfg(cost="P2L"))
    assert ok and st.outer_iterations == int(g["outer"]) and st.inner_iterations == int(g["inner"])
    np.testing.assert_allclose(op, g["poses_out"], atol=1e-9)


def test_golden_vectors_v2(orc):
    """P2D registration against two keyframes (pose, iteration counts, cost, covariance), GetCost, covariance by sampling
    and an 8-scan fuser replay, frozen in cfear_golden_v2.npz."""
    g = np.load(os.path.join(GOLD, "cfear_golden_v2.npz"))
    K = 2
    img = synth.make_problem_images(int(g["seed"]), K)[0]
    assert np.array_equal(img[K, ::8, ::8], g["img_sub"])
    sets = [helpers.oracle_cells(orc, img[i], radius=3.0)[1] for i in range(K + 1)]
    cfg = orc.reg_cfg(cost="P2D", loss="Huber", weight_opt=4, regularization=0.1)
    ok, op, cov, st, _ = orc.register(sets, g["poses_in"], cfg)
    assert ok and st.outer_iterations == int(g["outer"]) and st.inner_iterations == int(g["inner"])
    assert st.num_residuals == int(g["num_residuals"])
    np.testing.assert_allclose(op, g["poses_out"], atol=1e-9)
    np.testing.assert_allclose(st.final_cost, float(g["final_cost"]), rtol=1e-9)
    np.testing.assert_allclose(cov, g["cov"], rtol=1e-7, atol=1e-14)
    gok, gcost, gnres = orc.get_cost(sets, op, cfg)
    assert gok and gnres == int(g["get_cost_nres"])
    np.testing.assert_allclose(gcost, float(g["get_cost"]), rtol=1e-9)
    sok, scov, S = orc.sampled_covariance(sets, op, cfg, st.final_cost, st.num_residuals)
    assert sok
    np.testing.assert_allclose(S, g["samples"], rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(scov, g["sampled_cov"], rtol=1e-5, atol=1e-14)
    seq = synth.make_sequence(int(g["seq_seed"]), 8)[0]
    assert np.array_equal(seq[-1, ::8, ::8], g["seq_sub"])
    rep = orc.odometry_sequence(seq, orc.reg_cfg(cost="P2L", weight_opt=0), radius=3.5, weight_intensity=True, submap_scan_size=3)
    assert np.array_equal(rep["keyframe"], g["seq_keyframe"])
    np.testing.assert_allclose(rep["poses"], g["seq_poses"], atol=1e-9)


# ---- boundary: header <-> library <-> binding -------------------------------------------------------------------
def test_capi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "cfear_b200.h")).read()
    declared = set(re.findall(r"\b(cfear_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    lib = ctypes.CDLL(capi.LIB_PATH)
    for s in declared:
        assert hasattr(lib, s), s


def test_config_struct_layout_matches_header():
    cfg = capi.default_config()
    assert (cfg.azimuths, cfg.range_bins, cfg.k_strongest) == (400, 3360, 12)
    assert abs(cfg.z_min - 60) < 1e-6 and abs(cfg.range_res - 0.0438) < 1e-6 and abs(cfg.radius - 3.5) < 1e-6
    assert (cfg.cost, cfg.loss, cfg.max_outer, cfg.min_outer, cfg.max_inner) == (1, 1, 8, 3, 20)
    assert cfg.reg_radius == 2.0 and cfg.loss_limit == 0.1 and cfg.regularization == 1.0
    assert ctypes.sizeof(capi.RegStats) == capi.STATS_DTYPE.itemsize == 48
    assert capi.CELL_DTYPE.itemsize == 88


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.CfearError, match="no CUDA device|CUDA"):
        capi.Context()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "cfear_radarodometry_code_public_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle", src, re.M), f
                assert "cfear_oracle" not in src and "orc_" not in src, f


def test_synth_is_seeded():
    a = synth.make_problem_images(11, 1)
    b = synth.make_problem_images(11, 1)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert a[0].shape == (2, 400, 3360) and a[0].dtype == np.uint8


def test_png_ingestion_conventions(tmp_path):
    import cv2
    from cfear_radarodometry_code_public_b200 import io as cio
    rng = np.random.Generator(np.random.PCG64(1))
    polar = rng.integers(0, 256, (400, 3768), dtype=np.uint8)
    ts = (1547120000000000 + np.arange(400) * 625).astype(np.int64)
    sweep = ((np.arange(400) * 14) % 5600).astype(np.uint16)
    raw = np.concatenate([ts.view(np.uint8).reshape(400, 8), sweep.view(np.uint8).reshape(400, 2),
                          np.full((400, 1), 255, np.uint8), polar], 1)
    p = str(tmp_path / "oxford.png"); cv2.imwrite(p, raw)
    got, gts, az, valid = cio.load_oxford_png(p)
    assert np.array_equal(got, polar) and np.array_equal(gts, ts) and valid.all()
    np.testing.assert_allclose(az, sweep * 2 * np.pi / 5600)
    ra = rng.integers(0, 256, (3360, 400), dtype=np.uint8)          # MulRan layout: range x azimuth
    p2 = str(tmp_path / "mulran.png"); cv2.imwrite(p2, ra)
    rot = cio.load_range_azimuth_png(p2)
    assert rot.shape == (400, 3360) and rot.flags["C_CONTIGUOUS"]
    assert np.array_equal(rot, cv2.rotate(ra, cv2.ROTATE_90_COUNTERCLOCKWISE))
    with pytest.raises(FileNotFoundError):
        cio.load_oxford_png(str(tmp_path / "missing.png"))


def test_trajectory_writers_match_reference_formats(tmp_path):
    """est (KITTI, fixed 6 decimals), tum (sec.nsec9 x y z [fixed 4] qx qy qz qw [defaultfloat, precision 4]) and
    cov (36 numbers, stream precision 6) rows as EvalTrajectory::Write / WriteTUM / WriteCov produce them."""
    import subprocess
    exe = str(tmp_path / "writer_test")
    pkg = os.path.join(ROOT, "cfear_radarodometry_code_public_b200")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "writer_test.cpp"),
                           "-o", exe, "-L" + pkg, "-lcfear_b200", "-Wl,-rpath," + pkg])
    est, tum, cov = (str(tmp_path / n) for n in ("est.txt", "tum.txt", "cov.txt"))
    subprocess.check_call([exe, est, tum, cov])
    rows = open(est).read().strip().split("\n")
    c, s = np.cos(0.3), np.sin(0.3)
    assert rows[0] == "%.6f %.6f %.6f %.6f %.6f %.6f %.6f %.6f %.6f %.6f %.6f %.6f" % (c, -s, 0, 1.23456789, s, c, 0, -2.5, 0, 0, 1, 0)
    t = open(tum).read().strip().split("\n")
    assert t[0] == "1547120000.000000625 1.2346 -2.5000 0.0000 0 0 %.4g %.4g" % (np.sin(0.15), np.cos(0.15))
    # yaw = -2.9: trace <= 0 -> Eigen's largest-diagonal branch: qz > 0, qw < 0
    f = t[1].split(" ")
    assert f[0] == "1547120001.250000000" and f[1:4] == ["100.0000", "0.0000", "0.0000"]
    assert float(f[6]) > 0 and float(f[7]) < 0
    np.testing.assert_allclose([float(f[6]), float(f[7])], [-np.sin(-1.45), -np.cos(-1.45)], rtol=2e-4)
    cr = open(cov).read().strip().split("\n")[0].split(" ")
    assert cr[0] == "1547120000.000000625" and len(cr) == 37
    assert cr[1] == "0.0123457" and cr[6] == "-1.5e-07" and cr[36] == "0.0001" and cr[2] == "0"


def test_mirror_dense_helpers_match_numpy(tmp_path):
    """The Householder least-squares fit, the symmetric 3x3 eigenvalues and the 3x3 inverse behind the mirror's
    approximateCovarianceBySampling (the reference uses Eigen's bdcSvd / SelfAdjointEigenSolver / inverse there),
    on the reference's own sampling grid (+-0.2 m, +-0.00218 rad, 3 samples per axis: badly scaled columns)."""
    import struct
    import subprocess
    exe = str(tmp_path / "linalg_test")
    pkg = os.path.join(ROOT, "cfear_radarodometry_code_public_b200")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "linalg_test.cpp"),
                           "-o", exe, "-L" + pkg, "-lcfear_b200", "-Wl,-rpath," + pkg])
    rng = np.random.default_rng(5)
    xs = np.linspace(-0.2, 0.2, 3); ts = np.linspace(-0.00218125, 0.00218125, 3)
    g = np.array([(x, y, t) for t in ts for x in xs for y in xs])
    x, y, z = g.T
    A = np.stack([x * x, y * y, z * z, x * y, y * z, z * x, x, y, z, np.ones_like(x)], 1)
    qtrue = np.array([40.0, 55.0, 9e4, -6.0, 30.0, -80.0, 0.3, -0.2, 12.0, 17.5])
    c = A @ qtrue + 1e-9 * rng.standard_normal(27)
    H = np.array([[2 * qtrue[0], qtrue[3], qtrue[5]], [qtrue[3], 2 * qtrue[1], qtrue[4]], [qtrue[5], qtrue[4], 2 * qtrue[2]]])
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<ii", 27, 10)); f.write(A.tobytes()); f.write(c.tobytes()); f.write(H.tobytes())
    subprocess.check_call([exe, fin, fout])
    raw = open(fout, "rb").read()
    assert struct.unpack_from("<i", raw, 0)[0] == 1
    q = np.frombuffer(raw, np.float64, 10, 4); ev = np.frombuffer(raw, np.float64, 3, 84); Hi = np.frombuffer(raw, np.float64, 9, 108).reshape(3, 3)
    qn = np.linalg.lstsq(A, c, rcond=None)[0]
    np.testing.assert_allclose(q, qn, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(np.sort(ev), np.linalg.eigvalsh(H), rtol=1e-12)
    np.testing.assert_allclose(Hi, np.linalg.inv(H), rtol=1e-11, atol=1e-18)


def test_oracle_get_cost_and_sampled_covariance(orc):
    """Oracle restatements of GetCost / approximateCovarianceBySampling: the cost at the registered pose is the
    solver's final cost up to the last re-association, the sampled cost surface is convex around it and the fitted
    covariance is symmetric positive definite in (x, y, yaw)."""
    import helpers
    K = 2
    im, tp = helpers.scan_images(31, K)
    sets = [helpers.oracle_cells(orc, im[i], radius=3.0)[1] for i in range(K + 1)]
    P = tp.copy(); P[K] = tp[K - 1]
    cfg = orc.reg_cfg(cost="P2D", loss="Huber", weight_opt=4, regularization=0.1)
    ok, p, _, st, _ = orc.register(sets, P, cfg)
    assert ok
    gok, cost, nres = orc.get_cost(sets, p, cfg)
    assert gok and nres > 100
    np.testing.assert_allclose(cost, st.final_cost, rtol=0.05)         # same pose, fresh association
    sok, cov6, S = orc.sampled_covariance(sets, p, cfg, st.final_cost, st.num_residuals)
    assert sok and S.shape == (27, 4) and S[:, 3].min() >= cost * 0.999
    c3 = cov6[np.ix_([0, 1, 5], [0, 1, 5])]
    np.testing.assert_allclose(c3, c3.T, rtol=1e-9)
    assert np.all(np.linalg.eigvalsh(c3) > 0) and cov6[2, 2] == 1 and cov6[0, 2] == 0
    # degenerate: a source that sees nothing of the keyframes
    far = p.copy(); far[K, :2] += 1e4
    assert orc.get_cost(sets, far, cfg)[0] is False


def test_frame_file_layout(tmp_path):
    """io.write_frames: the raw frame container examples/offline_odometry.cpp reads (header, per-frame stamp + image)."""
    import struct
    from cfear_radarodometry_code_public_b200 import io as cio
    imgs = (np.arange(3 * 4 * 16, dtype=np.uint32) % 251).astype(np.uint8).reshape(3, 4, 16)
    path = str(tmp_path / "s.cfrs")
    cio.write_frames(path, imgs, stamps_ns=[5, 250_000_005, 500_000_005])
    raw = open(path, "rb").read()
    assert raw[:4] == b"CFRS" and struct.unpack_from("<3i", raw, 4) == (3, 4, 16)
    assert len(raw) == 16 + 3 * (8 + 64)
    for i, t in enumerate([5, 250_000_005, 500_000_005]):
        off = 16 + i * 72
        assert struct.unpack_from("<Q", raw, off)[0] == t
        assert raw[off + 8: off + 72] == imgs[i].tobytes()


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): exactly one JSON line on stdout with
    the contract's keys, a cpu_baseline describing the run and an e2e object repeating the line's value."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--nprob", "4"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e", "impl"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "scans/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
