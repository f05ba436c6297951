"""OdometryKeyframeFuser::processFrame written a second time in numpy from the reference source
(odometrykeyframefuser.cpp:143-259: compensation with the previous motion, constant-velocity guess, Register against the
keyframe window, sanity check :76-94, keyframe rule :62-73, AddToReference :470-476) on top of the independent stages of
test_surface_independent.py and test_register_independent.py, against the oracle's sequence replay.  From the filtered cloud
onwards nothing of the oracle is used; the filter rows are pinned to the reference source (test_ref_pin.py).  CPU only.
"""
import numpy as np
import pytest

from cfear_radarodometry_code_public_b200 import synth
from test_register_independent import register_py
from test_surface_independent import cells_np, compensate_np

pytest.importorskip("cv2")


def _mul(a, b):
    c, s = np.cos(a[2]), np.sin(a[2])
    return np.array([a[0] + c * b[0] - s * b[1], a[1] + s * b[0] + c * b[1], a[2] + b[2]])


def _inv(a):
    c, s = np.cos(a[2]), np.sin(a[2])
    return np.array([-(c * a[0] + s * a[1]), -(-s * a[0] + c * a[1]), -a[2]])


def _wrap(a):
    return (a + np.pi) % (2 * np.pi) - np.pi


def fuser_py(orc, imgs, cost, wopt, reg, res=3.0, submap=3, min_dist=1.5, min_rot_deg=5.0):
    T_prev, Tmot = np.zeros(3), np.zeros(3)
    keyframes, poses, kf_flag = [], [], []
    for img in imgs:
        idx, cnt = orc.kstrongest(img, 60, 12)
        cloud = orc.cloud(img, idx, cnt)
        TprevMot = Tmot.copy()
        cloud = compensate_np(cloud, TprevMot)                                       # :147-150
        cells = cells_np(cloud, res, True)                                           # :161
        Tguess = _mul(T_prev, TprevMot)                                              # :165
        if not keyframes:                                                            # :171-177
            keyframes.append((np.zeros(3), cells))
            poses.append(np.zeros(3)); kf_flag.append(1)
            continue
        sets = [k[1] for k in keyframes] + [cells]
        P = np.array([k[0] for k in keyframes] + [Tguess])
        _ok, x, *_ = register_py(sets, P, cost, wopt, reg=reg)                       # :186 (the return value is shadowed)
        Tcur = x
        Tmot_cur = _mul(_inv(T_prev), Tcur)
        dt = 0.25                                                                    # :76-94
        vel = np.hypot(*Tmot_cur[:2]) / dt
        acc = np.hypot(*(Tmot_cur[:2] - Tmot[:2])) / (dt * dt)
        if acc > 200 or vel > 200:
            Tcur = Tguess
        Tmot = _mul(_inv(T_prev), Tcur)                                              # :199
        diff = _mul(_inv(keyframes[-1][0]), Tcur)                                    # :231
        fuse = np.hypot(*diff[:2]) > min_dist or abs(_wrap(diff[2])) > min_rot_deg * np.pi / 180.0   # :62-73
        if fuse:                                                                     # :238-252
            keyframes.append((Tcur.copy(), cells))
            if len(keyframes) > submap:                                              # :470-476
                keyframes.pop(0)
        T_prev = Tcur.copy()                                                         # :258
        poses.append(Tcur.copy()); kf_flag.append(int(fuse))
    return np.array(poses), np.array(kf_flag)


@pytest.mark.parametrize("cost,wopt,reg,seed,min_dist", [("P2L", 0, 1.0, 21, 1.5), ("P2D", 4, 0.1, 22, 1.5), ("P2D", 4, 0.1, 23, 4.0),
                                                         ("P2L", 2, 1.0, 24, 6.0)])
def test_oracle_sequence_replay_matches_an_independent_numpy_fuser(orc, cost, wopt, reg, seed, min_dist):
    n = 8
    imgs, truth = synth.make_sequence(seed, n)
    cfg = orc.reg_cfg(cost=cost, loss="Huber", loss_limit=0.1, weight_opt=wopt, regularization=reg, cov_scale=1.0)
    ref = orc.odometry_sequence(imgs, cfg, radius=3.0, weight_intensity=True, compensate=True, submap_scan_size=3,
                                min_keyframe_dist=min_dist)
    poses, kf = fuser_py(orc, imgs, cost, wopt, reg, res=3.0, submap=3, min_dist=min_dist)
    # 2.5 m between scans: every scan is a keyframe at the default 1.5 m, every second / third one at 4 m / 6 m
    assert np.array_equal(kf[1:], ref["keyframe"][1:]) and kf.sum() >= 3 and (min_dist == 1.5 or kf.sum() < n)
    d = poses - ref["poses"]
    d[:, 2] = _wrap(d[:, 2])
    # the compensated clouds differ by an ulp on a few coordinates (numpy's sin / cos vs glibc's): poses agree to ~1e-8
    assert np.hypot(d[:, 0], d[:, 1]).max() < 1e-6 and np.abs(d[:, 2]).max() < 1e-7, (np.hypot(d[:, 0], d[:, 1]).max(), np.abs(d[:, 2]).max())
    end = np.hypot(*(poses[-1, :2] - truth[-1, :2]))
    assert end < 0.5 and np.hypot(*poses[-1, :2]) > 10.0                             # and it is odometry: ~17 m travelled, end point within 0.5 m
