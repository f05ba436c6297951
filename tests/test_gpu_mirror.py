"""The C++ host mirror (include/cfear_b200.hpp: radarDriver / MapPointNormal / n_scan_normal_reg over the C ABI)
driven like offline_odometry drives the reference classes, checked against the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

import helpers
from cfear_radarodometry_code_public_b200.synth import se2_inv, se2_mul

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("kernel_form")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cfear_radarodometry_code_public_b200")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cpp") / "mirror_test")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "mirror_test.cpp"), "-o", out, "-L" + PKG, "-lcfear_b200",
                           "-Wl,-rpath," + PKG])
    return out


@pytest.mark.parametrize("cost,loss,wopt,K,flags", [("P2L", "Huber", 0, 1, 0), ("P2D", "Huber", 4, 4, 2), ("P2P", "Cauchy", 4, 3, 0),
                                                      ("P2D", "Huber", 4, 3, 3), ("P2L", "Huber", 1, 2, 1)])
def test_mirror_matches_oracle(exe, orc, tmp_path, cost, loss, wopt, K, flags):
    """flags: 1 = Register(..., soft_constraints = true) with a non-trivial guess covariance, 2 = the public
    scan_associations_ / weight_associations_ tables are filled and checked."""
    im, tp = helpers.scan_images(31, K)
    radius, reg = 3.5, 0.1
    P = tp.copy(); P[K] = tp[K - 1]
    mot = se2_mul(se2_inv(tp[K - 2]), tp[K - 1]) if K >= 2 else np.array([2.5, 0.0, 0.02])
    ci, li = orc.COST[cost], orc.LOSS[loss]
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<iiifiiidi", K + 1, im.shape[1], im.shape[2], radius, ci, li, wopt, reg, flags))
        f.write(im[:K + 1].tobytes()); f.write(P.astype(np.float64).tobytes()); f.write(mot.astype(np.float64).tobytes())
    subprocess.check_call([exe, fin, fout])
    raw = open(fout, "rb").read()
    ok, itr, nres, ncells, npts = struct.unpack_from("<5i", raw, 0)
    pose = np.frombuffer(raw, np.float64, 3, 20)
    score = np.frombuffer(raw, np.float64, 1, 44)[0]
    cov = np.frombuffer(raw, np.float64, 36, 52).reshape(6, 6)
    near = struct.unpack_from("<i", raw, 52 + 288)[0]
    n_assoc = struct.unpack_from("<i", raw, 52 + 288 + 4)[0]
    sums = np.frombuffer(raw, np.float64, 4, 52 + 288 + 8)
    stamps = np.frombuffer(raw, np.float64, 2, 52 + 288 + 8 + 32)
    # oracle, same call sequence
    sets = []
    for i in range(K + 1):
        cl, s = helpers.oracle_cells(orc, im[i], radius=radius, mot=(mot if i == K else None))
        sets.append(s)
    L = None
    if flags & 1:
        c6 = np.eye(6); c6[0, 0] = 0.04; c6[1, 1] = 0.09; c6[0, 1] = c6[1, 0] = 0.01; c6[5, 5] = 0.0004; c6[0, 5] = c6[5, 0] = 0.001
        L = orc.prior_sqrt_information(c6)
    o_ok, op, ocov, ost, oassoc, osim = orc.register(sets, P, orc.reg_cfg(cost=cost, loss=loss, weight_opt=wopt, regularization=reg),
                                                     prior_sqrt_info=L, want_sim=True)
    assert bool(ok) == o_ok and itr == ost.outer_iterations and nres == ost.num_residuals
    if flags & 2:
        tar, src = np.nonzero(oassoc >= 0)
        assert n_assoc == tar.size > 100
        assert sums[0] == oassoc[oassoc >= 0].sum() and sums[1] == src.sum()
        np.testing.assert_allclose(sums[2], osim[oassoc >= 0].sum(), rtol=1e-12)
        nsrc = sets[-1]["nsamples"][src].astype(float); ntar = np.array([sets[t]["nsamples"][m] for t, m in zip(tar, oassoc[oassoc >= 0])], float)
        psrc = sets[-1]["planarity"][src]; ptar = np.array([sets[t]["planarity"][m] for t, m in zip(tar, oassoc[oassoc >= 0])])
        sim_n, sim_p = 2 * np.minimum(nsrc, ntar) / (nsrc + ntar), 2 * np.minimum(psrc, ptar) / (psrc + ptar)
        w = {0: np.ones_like(sim_n), 1: sim_n, 2: osim[oassoc >= 0], 3: sim_p, 4: sim_n + osim[oassoc >= 0] + sim_p}[wopt]
        np.testing.assert_allclose(sums[3], w.sum(), rtol=1e-12)
    else:
        assert n_assoc == 0
    m0 = sets[-1]["mean"][0]
    a = np.arctan2(m0[1], m0[0]); dd = (a if a > 1e-5 else 2 * np.pi + a) / (2 * np.pi)
    np.testing.assert_allclose(stamps, [dd - 0.5, -(dd - 0.5)], atol=1e-12)
    assert ncells == sets[-1]["mean"].shape[0] and npts == cl.shape[0]
    d = pose - op[K]
    assert np.hypot(d[0], d[1]) < 1e-4 and abs(d[2]) < 1e-5
    np.testing.assert_allclose(score, ost.score, rtol=1e-6)
    np.testing.assert_allclose(cov, ocov, rtol=1e-5, atol=1e-12)
    assert near == orc.nearest(sets[-1]["mean"], np.array([[10.0, 0.0]]), 50.0)[0]


@pytest.fixture(scope="module")
def fuser_exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cpp") / "fuser_test")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "fuser_test.cpp"), "-o", out, "-L" + PKG, "-lcfear_b200",
                           "-Wl,-rpath," + PKG])
    return out


@pytest.mark.parametrize("cost,wopt,submap,wint,res", [("P2L", 0, 3, True, 3.5), ("P2P", 4, 4, True, 3.0), ("P2D", 4, 2, False, 3.0)])
def test_fuser_sequence_replay_matches_oracle(fuser_exe, orc, tmp_path, cost, wopt, submap, wint, res):
    """OdometryKeyframeFuser::processFrame semantics over a 14-scan synthetic sequence: same keyframe decisions, poses
    within 1e-4 m / 1e-5 rad of the oracle replay at every scan, KITTI rows in the reference's fixed 6-decimal format."""
    from cfear_radarodometry_code_public_b200 import synth
    imgs, truth = synth.make_sequence(7, 14)
    reg = 0.1 if cost == "P2D" else 0.0
    fin, fout, fest = str(tmp_path / "in.bin"), str(tmp_path / "out.bin"), str(tmp_path / "est.txt")
    with open(fin, "wb") as f:
        f.write(struct.pack("<6if4sd", imgs.shape[0], imgs.shape[1], imgs.shape[2], submap, wopt, int(wint), res,
                            cost.encode() + b"\0", reg))
        f.write(imgs.tobytes())
    subprocess.check_call([fuser_exe, fin, fout, fest])
    rec = np.frombuffer(open(fout, "rb").read(), dtype=np.dtype([("p", np.float64, 3), ("up", np.int32)]))
    ref = orc.odometry_sequence(imgs, orc.reg_cfg(cost=cost, weight_opt=wopt, regularization=reg), radius=res,
                                weight_intensity=wint, submap_scan_size=submap)
    assert np.array_equal(rec["up"], ref["keyframe"])
    d = rec["p"] - ref["poses"]
    assert np.hypot(d[:, 0], d[:, 1]).max() < 1e-4 and np.abs(d[:, 2]).max() < 1e-5
    assert np.hypot(*(rec["p"][-1, :2] - truth[-1, :2])) < 1.0          # tracks the simulated motion
    rows = open(fest).read().strip().split("\n")
    assert len(rows) == imgs.shape[0]
    x, y, yaw = rec["p"][-1]
    c, s = np.cos(yaw), np.sin(yaw)
    exp = "%.6f %.6f %.6f %.6f %.6f %.6f %.6f %.6f %.6f %.6f %.6f %.6f" % (c, -s, 0, x, s, c, 0, y, 0, 0, 1, 0)
    assert rows[-1] == exp


@pytest.fixture(scope="module")
def covsample_exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cpp") / "covsample_test")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "covsample_test.cpp"), "-o", out, "-L" + PKG, "-lcfear_b200",
                           "-Wl,-rpath," + PKG])
    return out


@pytest.mark.parametrize("cost,wopt,K", [("P2D", 4, 3), ("P2L", 0, 2)])
def test_covariance_by_sampling_matches_oracle(covsample_exe, orc, tmp_path, cost, wopt, K):
    """approximateCovarianceBySampling (27 GetCost samples in one launch + quadric fit) of the mirror vs the oracle
    restatement (one GetCost per sample, numpy lstsq)."""
    im, tp = helpers.scan_images(33, K)
    radius, reg = 3.0, 0.1
    P = tp.copy(); P[K] = tp[K - 1]
    mot = se2_mul(se2_inv(tp[K - 2]), tp[K - 1])
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<iiifiiid", K + 1, im.shape[1], im.shape[2], radius, orc.COST[cost], orc.LOSS["Huber"], wopt, reg))
        f.write(im[:K + 1].tobytes()); f.write(P.astype(np.float64).tobytes()); f.write(mot.astype(np.float64).tobytes())
    subprocess.check_call([covsample_exe, fin, fout])
    raw = open(fout, "rb").read()
    reg_ok, cov_ok, nres = struct.unpack_from("<3i", raw, 0)
    final_cost = np.frombuffer(raw, np.float64, 1, 12)[0]
    pose = np.frombuffer(raw, np.float64, 3, 20)
    cov = np.frombuffer(raw, np.float64, 36, 44).reshape(6, 6)
    cost_here = np.frombuffer(raw, np.float64, 1, 44 + 288)[0]
    sets = [helpers.oracle_cells(orc, im[i], radius=radius, mot=(mot if i == K else None))[1] for i in range(K + 1)]
    cfg = orc.reg_cfg(cost=cost, loss="Huber", weight_opt=wopt, regularization=reg)
    o_ok, op, _, ost, _ = orc.register(sets, P, cfg)
    assert bool(reg_ok) == o_ok and nres == ost.num_residuals
    np.testing.assert_allclose(final_cost, ost.final_cost, rtol=1e-6)
    g_ok, g_cost, g_nres = orc.get_cost(sets, op, cfg)
    np.testing.assert_allclose(cost_here, g_cost, rtol=1e-6)
    s_ok, ocov, _ = orc.sampled_covariance(sets, op, cfg, ost.final_cost, ost.num_residuals)
    assert bool(cov_ok) == s_ok
    if s_ok:
        np.testing.assert_allclose(cov, ocov, rtol=2e-3, atol=1e-12)


def test_offline_odometry_example_matches_oracle_replay(orc, tmp_path):
    """examples/offline_odometry.cpp -- the reference's radarReader loop and command-line options over the mirror -- on a
    synthetic 10-frame file: est/01.txt (KITTI rows) equals the oracle's sequential replay, TUM / cov files are written."""
    from cfear_radarodometry_code_public_b200 import io as cio, synth
    exe = str(tmp_path / "offline_odometry")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "offline_odometry.cpp"), "-o", exe, "-L" + PKG, "-lcfear_b200",
                           "-Wl,-rpath," + PKG])
    imgs, _ = synth.make_sequence(9, 10)
    frames = str(tmp_path / "seq.cfrs")
    cio.write_frames(frames, imgs)
    out = subprocess.check_output([exe, "--frames", frames, "--est_directory", str(tmp_path), "--cost_type", "P2L", "--res", "3.5",
                                   "--submap_scan_size", "3", "--z-min", "65", "--weight_option", "0", "--tum", "true", "--cov=true",
                                   "--sequence", "2019-01-10-12-32-52-radar-oxford-10k"]).decode()
    assert "Frame: 9" in out and "Trajectory saved to" in out
    rows = np.loadtxt(str(tmp_path / "01.txt"))
    assert rows.shape == (10, 12)
    ref = orc.odometry_sequence(imgs, orc.reg_cfg(cost="P2L", weight_opt=0, regularization=1.0), z_min=65, radius=3.5,
                                weight_intensity=True, submap_scan_size=3)
    x, y, yaw = rows[:, 3], rows[:, 7], np.arctan2(rows[:, 4], rows[:, 0])
    assert np.abs(x - ref["poses"][:, 0]).max() < 1e-4 + 1e-6 and np.abs(y - ref["poses"][:, 1]).max() < 1e-4 + 1e-6
    assert np.abs(yaw - ref["poses"][:, 2]).max() < 1e-5 + 2e-6
    assert len(open(str(tmp_path / "01_tum.txt")).read().strip().split("\n")) == 10
    assert len(open(str(tmp_path / "01_cov.txt")).read().strip().split("\n")[0].split(" ")) == 37


def test_oxford_png_directory_to_trajectory_diff(orc, tmp_path):
    """BASELINE configs[3] dry run on synthetic data in the real format: a directory of Oxford Radar RobotCar PNGs (11
    metadata bytes + 3768 range bins per azimuth) -> io.load_oxford_png -> frame file -> examples/offline_odometry (which
    shapes the device context from the frame header) -> est/01.txt -> tools/traj_diff.py against the trajectory of the
    oracle's sequential replay written in the same KITTI format.  The day the dataset is available the same three commands
    produce the diff against the reference's est/01.txt."""
    import sys
    import cv2
    from cfear_radarodometry_code_public_b200 import io as cio, synth
    n, R = 8, 3768
    imgs, _ = synth.make_sequence(12, n, R=R)
    radar_dir = tmp_path / "radar"; radar_dir.mkdir()
    t0 = 1547131046353776
    for i in range(n):
        raw = np.zeros((400, 11 + R), np.uint8)
        raw[:, :8] = (np.int64(t0 * 1000) + np.arange(400, dtype=np.int64) * 625000 + i * 250000000).view(np.uint8).reshape(400, 8)
        raw[:, 8:10] = (np.arange(400, dtype=np.uint16) * 14).view(np.uint8).reshape(400, 2)
        raw[:, 10] = 255
        raw[:, 11:] = imgs[i]
        assert cv2.imwrite(str(radar_dir / f"{t0 + i * 250000}.png"), raw)
    files = sorted(os.listdir(str(radar_dir)))
    loaded = [cio.load_oxford_png(str(radar_dir / f)) for f in files]
    frames = str(tmp_path / "oxford.cfrs")
    cio.write_frames(frames, np.stack([l[0] for l in loaded]), stamps_ns=np.array([l[1][0] for l in loaded], dtype=np.uint64))
    exe = str(tmp_path / "offline_odometry")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "offline_odometry.cpp"), "-o", exe, "-L" + PKG, "-lcfear_b200",
                           "-Wl,-rpath," + PKG])
    est_dir = tmp_path / "est"; est_dir.mkdir()
    # the CFEAR-3 style options of launch/oxford_demo (k = 12 here keeps the test light), window of 4 keyframes
    subprocess.check_output([exe, "--frames", frames, "--est_directory", str(est_dir), "--cost_type", "P2D", "--res", "3.0",
                             "--submap_scan_size", "4", "--z-min", "60", "--weight_option", "4", "--regularization", "0.1",
                             "--k_strongest", "12", "--sequence", "2019-01-10-12-32-52-radar-oxford-10k"])
    ref = orc.odometry_sequence(imgs, orc.reg_cfg(cost="P2D", weight_opt=4, regularization=0.1), z_min=60, radius=3.0,
                                weight_intensity=True, submap_scan_size=4)
    ref_file = str(tmp_path / "ref_01.txt")
    with open(ref_file, "w") as f:
        for x, y, t in ref["poses"]:
            c, s = np.cos(t), np.sin(t)
            f.write(" ".join("%.6f" % v for v in (c, -s, 0, x, s, c, 0, y, 0, 0, 1, 0)) + "\n")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "traj_diff.py"), str(est_dir / "01.txt"), ref_file,
                          "--lengths", "5,10", "--json", "--tol-pos", "1e-4", "--tol-rot", "1e-5"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    import json
    d = json.loads(out.stdout)
    assert d["poses"] == n and d["kitti_drift"]["segments"] > 0 and d["kitti_drift"]["trans_percent"] < 1e-3
    assert np.hypot(*ref["poses"][-1, :2]) > 10.0            # the vehicle actually moved (2.5 m per scan)
