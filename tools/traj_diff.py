#!/usr/bin/env python
"""traj_diff.py -- compare two trajectory files in the format EvalTrajectory::Write produces.

The reference stores its estimate as KITTI odometry rows, one per scan: the upper 3x4 block of the pose, row-major, 12
numbers, fixed 6 decimals (src/cfear_radarodometry/eval_trajectory.cpp:169-183; file name <est_directory>/<NN>.txt from the
sequence name, :74-143).  BASELINE.json configs[3] is defined as the difference between such a file produced by this
repository (examples/offline_odometry, or replay.py) and the reference's own est/01.txt for Oxford 2019-01-10-12-32-52.

Reported (planar: x, y, yaw are taken from the 3x4 block; the pose path is SE(2)):
  * absolute per-pose error   |t_est - t_ref| and wrapped yaw difference: max / mean / rmse, and the end-point error
  * relative per-step error   of the inter-scan motions  T_i^-1 T_{i+1}  (what the registration actually estimates)
  * KITTI drift               translational [%] and rotational [deg/m] error averaged over all sub-trajectories of the
                              given lengths (default 100..800 m in steps of 100 m, every 10th frame as a start), i.e. the
                              number the reference quotes as "drift" (launch/oxford_demo:32,62)

usage: traj_diff.py EST.txt REF.txt [--lengths 100,200,...] [--step 10] [--json] [--tol-pos M --tol-rot RAD]
Exit status 1 if --tol-pos / --tol-rot are given and the absolute per-pose error exceeds them.
"""
from __future__ import annotations

import argparse
import json
import sys

import numpy as np


def load_kitti(path: str) -> np.ndarray:
    """-> [n, 3] (x, y, yaw) from 12-column KITTI rows."""
    rows = np.loadtxt(path, ndmin=2)
    if rows.shape[1] != 12:
        raise ValueError(f"{path}: expected 12 columns per row (3x4 pose, row-major), found {rows.shape[1]}")
    return np.stack([rows[:, 3], rows[:, 7], np.arctan2(rows[:, 4], rows[:, 0])], 1)


def wrap(a):
    return (a + np.pi) % (2.0 * np.pi) - np.pi


def se2_inv_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """a^-1 * b for arrays of (x, y, yaw)."""
    c, s = np.cos(a[:, 2]), np.sin(a[:, 2])
    dx, dy = b[:, 0] - a[:, 0], b[:, 1] - a[:, 1]
    return np.stack([c * dx + s * dy, -s * dx + c * dy, wrap(b[:, 2] - a[:, 2])], 1)


def _stats(v):
    v = np.asarray(v, dtype=np.float64)
    if v.size == 0:
        return {"max": None, "mean": None, "rmse": None}
    return {"max": float(v.max()), "mean": float(v.mean()), "rmse": float(np.sqrt((v * v).mean()))}


def kitti_drift(est: np.ndarray, ref: np.ndarray, lengths, step: int = 10):
    """KITTI odometry devkit metric (evaluate_odometry.cpp: trajectoryDistances / calcSequenceErrors) on planar poses."""
    seg = np.hypot(np.diff(ref[:, 0]), np.diff(ref[:, 1]))
    dist = np.concatenate([[0.0], np.cumsum(seg)])
    n = ref.shape[0]
    t_err, r_err, per_len = [], [], {}
    for L in lengths:
        te, re = [], []
        for first in range(0, n, step):
            last = int(np.searchsorted(dist, dist[first] + L, side="left"))
            if last >= n:
                break
            d_ref = se2_inv_mul(ref[first:first + 1], ref[last:last + 1])
            d_est = se2_inv_mul(est[first:first + 1], est[last:last + 1])
            e = se2_inv_mul(d_est, d_ref)[0]                  # pose error  (d_est)^-1 d_ref
            te.append(np.hypot(e[0], e[1]) / L)
            re.append(abs(e[2]) / L)
        if te:
            per_len[str(L)] = {"n": len(te), "trans_percent": 100.0 * float(np.mean(te)), "rot_deg_per_m": float(np.degrees(np.mean(re)))}
            t_err += te; r_err += re
    if not t_err:
        return {"trans_percent": None, "rot_deg_per_m": None, "segments": 0, "per_length": {}, "path_length_m": float(dist[-1])}
    return {"trans_percent": 100.0 * float(np.mean(t_err)), "rot_deg_per_m": float(np.degrees(np.mean(r_err))),
            "segments": len(t_err), "per_length": per_len, "path_length_m": float(dist[-1])}


def diff(est: np.ndarray, ref: np.ndarray, lengths=None, step: int = 10) -> dict:
    if est.shape != ref.shape:
        raise ValueError(f"trajectories differ in length: {est.shape[0]} vs {ref.shape[0]} poses")
    lengths = list(lengths) if lengths is not None else list(range(100, 900, 100))
    dpos = np.hypot(est[:, 0] - ref[:, 0], est[:, 1] - ref[:, 1])
    drot = np.abs(wrap(est[:, 2] - ref[:, 2]))
    out = {"poses": int(est.shape[0]),
           "absolute": {"pos_m": _stats(dpos), "yaw_rad": _stats(drot)},
           "end_point": {"pos_m": float(dpos[-1]), "yaw_rad": float(drot[-1])}}
    if est.shape[0] > 1:
        me, mr = se2_inv_mul(est[:-1], est[1:]), se2_inv_mul(ref[:-1], ref[1:])
        out["relative_per_step"] = {"pos_m": _stats(np.hypot(me[:, 0] - mr[:, 0], me[:, 1] - mr[:, 1])),
                                    "yaw_rad": _stats(np.abs(wrap(me[:, 2] - mr[:, 2])))}
    out["kitti_drift"] = kitti_drift(est, ref, lengths, step)
    return out


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("est"); ap.add_argument("ref")
    ap.add_argument("--lengths", default=None, help="comma separated sub-trajectory lengths in metres (default 100..800)")
    ap.add_argument("--step", type=int, default=10)
    ap.add_argument("--json", action="store_true")
    ap.add_argument("--tol-pos", type=float, default=None)
    ap.add_argument("--tol-rot", type=float, default=None)
    a = ap.parse_args(argv)
    lengths = [float(x) for x in a.lengths.split(",")] if a.lengths else None
    d = diff(load_kitti(a.est), load_kitti(a.ref), lengths, a.step)
    if a.json:
        print(json.dumps(d))
    else:
        ab, ep, kd = d["absolute"], d["end_point"], d["kitti_drift"]
        print(f"poses {d['poses']}   path length {kd['path_length_m']:.1f} m")
        print(f"absolute  pos max {ab['pos_m']['max']:.3e} m  rmse {ab['pos_m']['rmse']:.3e} m | yaw max {ab['yaw_rad']['max']:.3e} rad")
        print(f"end point pos {ep['pos_m']:.3e} m  yaw {ep['yaw_rad']:.3e} rad")
        if "relative_per_step" in d:
            r = d["relative_per_step"]
            print(f"per step  pos max {r['pos_m']['max']:.3e} m  rmse {r['pos_m']['rmse']:.3e} m | yaw max {r['yaw_rad']['max']:.3e} rad")
        if kd["segments"]:
            print(f"KITTI drift  {kd['trans_percent']:.4f} %   {kd['rot_deg_per_m']:.6f} deg/m   ({kd['segments']} segments)")
        else:
            print("KITTI drift  n/a (trajectory shorter than the shortest segment length)")
    bad = (a.tol_pos is not None and d["absolute"]["pos_m"]["max"] > a.tol_pos) or \
          (a.tol_rot is not None and d["absolute"]["yaw_rad"]["max"] > a.tol_rot)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
