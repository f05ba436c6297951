"""tools/traj_diff.py: the SE(2) / KITTI-drift comparer of two est/NN.txt files (BASELINE configs[3])."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import traj_diff  # noqa: E402


def _write(path, poses):
    """KITTI rows like EvalTrajectory::Write (eval_trajectory.cpp:169-183): fixed, 6 decimals."""
    with open(path, "w") as f:
        for x, y, t in poses:
            c, s = np.cos(t), np.sin(t)
            f.write(" ".join("%.6f" % v for v in (c, -s, 0, x, s, c, 0, y, 0, 0, 1, 0)) + "\n")


def _arc(n, v=2.5, w=0.01):
    p = np.zeros((n, 3))
    for i in range(1, n):
        p[i, 2] = p[i - 1, 2] + w
        p[i, 0] = p[i - 1, 0] + v * np.cos(p[i, 2]); p[i, 1] = p[i - 1, 1] + v * np.sin(p[i, 2])
    return p


def test_identical_files_have_zero_error(tmp_path):
    p = _arc(500)
    a, b = str(tmp_path / "a.txt"), str(tmp_path / "b.txt")
    _write(a, p); _write(b, p)
    d = traj_diff.diff(traj_diff.load_kitti(a), traj_diff.load_kitti(b))
    assert d["poses"] == 500 and d["absolute"]["pos_m"]["max"] == 0.0 and d["end_point"]["yaw_rad"] == 0.0
    assert d["kitti_drift"]["trans_percent"] == 0.0 and d["kitti_drift"]["segments"] > 0
    assert subprocess.call([sys.executable, os.path.join(ROOT, "tools", "traj_diff.py"), a, b, "--tol-pos", "1e-6", "--tol-rot", "1e-6"],
                           stdout=subprocess.DEVNULL) == 0


def test_scale_error_gives_that_drift_percentage(tmp_path):
    ref = np.zeros((2000, 3)); ref[:, 0] = np.arange(2000) * 1.0          # 2 km straight line
    est = ref.copy(); est[:, 0] *= 1.01                                     # 1 % too long
    d = traj_diff.diff(est, ref)
    assert abs(d["kitti_drift"]["trans_percent"] - 1.0) < 1e-9 and d["kitti_drift"]["rot_deg_per_m"] == 0.0
    assert abs(d["end_point"]["pos_m"] - 19.99) < 1e-9
    assert set(d["kitti_drift"]["per_length"]) == {str(L) for L in range(100, 900, 100)}
    # a constant yaw-rate bias: rotational drift = bias per metre
    est = ref.copy(); est[:, 2] = np.arange(2000) * 1e-4
    d = traj_diff.diff(est, ref)
    assert abs(d["kitti_drift"]["rot_deg_per_m"] - np.degrees(1e-4)) < 1e-9


def test_per_pose_and_per_step_errors_and_cli(tmp_path):
    ref = _arc(60)
    est = ref.copy(); est[30:, 0] += 0.05; est[45, 2] += 0.002
    a, b = str(tmp_path / "est.txt"), str(tmp_path / "ref.txt")
    _write(a, est); _write(b, ref)
    d = traj_diff.diff(traj_diff.load_kitti(a), traj_diff.load_kitti(b), lengths=[20, 50])
    assert abs(d["absolute"]["pos_m"]["max"] - 0.05) < 2e-6 and abs(d["absolute"]["yaw_rad"]["max"] - 0.002) < 2e-6
    assert abs(d["relative_per_step"]["pos_m"]["max"] - 0.05) < 1e-3       # the jump shows up in one inter-scan motion
    assert d["kitti_drift"]["segments"] > 0 and d["kitti_drift"]["trans_percent"] > 0
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "traj_diff.py"), a, b, "--lengths", "20,50", "--json",
                          "--tol-pos", "0.01"], capture_output=True, text=True)
    assert out.returncode == 1 and '"kitti_drift"' in out.stdout              # tolerance exceeded -> exit status 1
    short = traj_diff.diff(est[:5], ref[:5])
    assert short["kitti_drift"]["segments"] == 0 and short["kitti_drift"]["trans_percent"] is None
