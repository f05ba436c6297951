"""ctypes binding of the CPU oracle (oracle/cfear_oracle.cc).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(cfear_radarodometry_code_public_b200) must never import this module.

Parity unpinned: see the header of cfear_oracle.cc.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libcfear_oracle.so")
_lib = None

COST = {"P2P": 0, "P2L": 1, "P2D": 2}
LOSS = {"None": 0, "Huber": 1, "Cauchy": 2, "SoftLOne": 3, "Combined": 4, "Tukey": 5}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "cfear_oracle.cc")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libcfear_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_eval_cost.restype = C.c_double
    return _lib


def _p(a, t=None):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


class RegStats(C.Structure):
    _fields_ = [("success", C.c_int32), ("outer_iterations", C.c_int32), ("inner_iterations", C.c_int32),
                ("num_residuals", C.c_int32), ("num_blocks", C.c_int32), ("usable", C.c_int32),
                ("final_cost", C.c_double), ("score", C.c_double), ("pose_written", C.c_int32), ("reserved", C.c_int32)]


def kstrongest(img: np.ndarray, z_min: int, k: int):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    A, R = img.shape
    idx = np.full((A, k), -1, np.int32)
    cnt = np.zeros(A, np.int32)
    rc = lib().orc_kstrongest(_p(img), A, R, int(z_min), int(k), _p(idx), _p(cnt))
    assert rc == 0
    return idx, cnt


def peaks(img, idx, cnt):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    A, R = img.shape
    k = idx.shape[1]
    pidx = np.full((A, k), -1, np.int32)
    pcnt = np.zeros(A, np.int32)
    lib().orc_peaks(_p(img), A, R, k, _p(np.ascontiguousarray(idx)), _p(np.ascontiguousarray(cnt)), _p(pidx), _p(pcnt))
    return pidx, pcnt


def cloud(img, idx, cnt, min_distance=2.5, range_res=0.0438):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    A, R = img.shape
    k = idx.shape[1]
    out = np.zeros((A * k, 4), np.float32)
    n = lib().orc_cloud(_p(img), A, R, k, _p(np.ascontiguousarray(idx)), _p(np.ascontiguousarray(cnt)),
                        C.c_float(min_distance), C.c_float(range_res), _p(out))
    return out[:n].copy()


def compensate(xyzi, mot, ccw=False):
    out = np.ascontiguousarray(xyzi, dtype=np.float32).copy()
    m = np.ascontiguousarray(mot, dtype=np.float64)
    lib().orc_compensate(_p(out), out.shape[0], _p(m), int(bool(ccw)))
    return out


def voxel_centroids(xyzi, radius, downsample_factor=1.0):
    xyzi = np.ascontiguousarray(xyzi, dtype=np.float32)
    n = xyzi.shape[0]
    cx = np.zeros(n, np.float32); cy = np.zeros(n, np.float32); ci = np.zeros(n, np.float32)
    vid = np.zeros(n, np.int32); dims = np.zeros(4, np.int32)
    nv = lib().orc_voxel_centroids(_p(xyzi), n, C.c_float(radius), C.c_double(downsample_factor),
                                   _p(cx), _p(cy), _p(ci), _p(vid), _p(dims))
    return cx[:nv], cy[:nv], ci[:nv], vid[:nv], dims


def surface_points(xyzi, radius, weight_intensity=True, downsample_factor=1.0, origin=(0.0, 0.0)):
    """Returns dict of SoA arrays (mean[n,2], normal[n,2], cov[n,2,2], planarity, nsamples, avg_intensity, lambdas)."""
    xyzi = np.ascontiguousarray(xyzi, dtype=np.float32)
    n = xyzi.shape[0]
    m = max(n, 1)
    mean = np.zeros((m, 2)); normal = np.zeros((m, 2)); cov = np.zeros((m, 2, 2))
    plan = np.zeros(m); ns = np.zeros(m, np.int32); avg = np.zeros(m); lam = np.zeros((m, 2))
    nc = lib().orc_surface_points(_p(xyzi), n, C.c_float(radius), C.c_double(downsample_factor), int(bool(weight_intensity)),
                                  C.c_double(origin[0]), C.c_double(origin[1]),
                                  _p(mean), _p(normal), _p(cov), _p(plan), _p(ns), _p(avg), _p(lam))
    return dict(mean=mean[:nc].copy(), normal=normal[:nc].copy(), cov=cov[:nc].copy(), planarity=plan[:nc].copy(),
                nsamples=ns[:nc].copy(), avg_intensity=avg[:nc].copy(), lambdas=lam[:nc].copy())


def nearest(means, queries, radius):
    means = np.ascontiguousarray(means, dtype=np.float64)
    queries = np.ascontiguousarray(queries, dtype=np.float64)
    out = np.zeros(queries.shape[0], np.int32)
    lib().orc_nearest(_p(means), means.shape[0], _p(queries), queries.shape[0], C.c_double(radius), _p(out))
    return out


def set_nn_tie_largest(on: bool):
    """Test knob: NN ties between cells with identical fp32 means go to the largest index instead of the smallest."""
    lib().orc_set_nn_tie_largest(int(bool(on)))


def reg_cfg(cost="P2L", loss="Huber", loss_limit=0.1, weight_opt=0, cov_scale=1.0, regularization=1.0,
            radius=2.0, max_outer=8, min_outer=3, max_inner=20, solver_mode=0, gn_iters=10):
    ci = np.array([COST[cost] if isinstance(cost, str) else cost, LOSS[loss] if isinstance(loss, str) else loss,
                   weight_opt, max_outer, min_outer, max_inner, solver_mode, gn_iters], np.int32)
    cd = np.array([loss_limit, cov_scale, regularization, radius], np.float64)
    return ci, cd


def concat_cellsets(sets):
    offs = np.zeros(len(sets) + 1, np.int32)
    for i, s in enumerate(sets):
        offs[i + 1] = offs[i] + s["mean"].shape[0]
    cat = lambda k, dt: np.ascontiguousarray(np.concatenate([np.asarray(s[k]).reshape(s["mean"].shape[0], -1) for s in sets], 0), dtype=dt)
    return (offs, cat("mean", np.float64), cat("normal", np.float64), cat("cov", np.float64),
            cat("planarity", np.float64).ravel(), cat("nsamples", np.int32).ravel())


def register(sets, poses, cfg, want_assoc=False, prior_sqrt_info=None, want_sim=False):
    """sets: list of K+1 cell dicts (last = current scan); poses: (K+1,3) (x,y,yaw), last = guess.
    prior_sqrt_info: 3x3 lower-triangular L of Register(..., soft_constraints=true) (n_scan_normal.cpp:373-377) or None.
    Returns (success, poses_out, cov6x6, RegStats, assoc or None) (+ the direction-similarity table when want_sim)."""
    ci, cd = cfg
    if prior_sqrt_info is not None or want_sim:
        return _register_ex(sets, poses, cfg, prior_sqrt_info, want_sim)
    offs, mean, normal, cov, plan, ns = concat_cellsets(sets)
    p = np.ascontiguousarray(poses, dtype=np.float64).copy()
    cov36 = np.zeros(36)
    st = RegStats()
    n_src = sets[-1]["mean"].shape[0]
    assoc = np.full((len(sets) - 1, n_src), -1, np.int32) if want_assoc else None
    ok = lib().orc_register(_p(ci), _p(cd), len(sets), _p(offs), _p(mean), _p(normal), _p(cov), _p(plan), _p(ns),
                            _p(p), _p(cov36), C.byref(st), _p(assoc))
    return bool(ok), p, cov36.reshape(6, 6), st, assoc


def _register_ex(sets, poses, cfg, prior_sqrt_info, want_sim):
    ci, cd = cfg
    offs, mean, normal, cov, plan, ns = concat_cellsets(sets)
    p = np.ascontiguousarray(poses, dtype=np.float64).copy()
    cov36 = np.zeros(36)
    st = RegStats()
    n_src = sets[-1]["mean"].shape[0]
    assoc = np.full((len(sets) - 1, n_src), -1, np.int32)
    sim = np.zeros((len(sets) - 1, n_src))
    L = None if prior_sqrt_info is None else np.ascontiguousarray(prior_sqrt_info, dtype=np.float64).reshape(9)
    ok = lib().orc_register_ex(_p(ci), _p(cd), len(sets), _p(offs), _p(mean), _p(normal), _p(cov), _p(plan), _p(ns),
                               _p(p), _p(cov36), C.byref(st), _p(assoc), _p(sim), _p(L))
    if want_sim:
        return bool(ok), p, cov36.reshape(6, 6), st, assoc, sim
    return bool(ok), p, cov36.reshape(6, 6), st, assoc


def prior_sqrt_information(cov6):
    """Cov6to3(cov).inverse().llt().matrixL() (n_scan_normal.cpp:374, registration.cpp:123-129)."""
    c = np.asarray(cov6, dtype=np.float64).reshape(6, 6)
    c3 = c[np.ix_([0, 1, 5], [0, 1, 5])]
    return np.linalg.cholesky(np.linalg.inv(c3))


def get_cost(sets, poses, cfg):
    """n_scan_normal_reg::GetCost at fixed poses (K+1, 3).  Returns (ok, cost, num_residuals)."""
    ci, cd = cfg
    offs, mean, normal, cov, plan, ns = concat_cellsets(sets)
    p = np.ascontiguousarray(poses, dtype=np.float64)
    cost = C.c_double(0.0); nres = C.c_int32(0)
    ok = lib().orc_get_cost(_p(ci), _p(cd), len(sets), _p(offs), _p(mean), _p(normal), _p(cov), _p(plan), _p(ns),
                            _p(p), C.byref(cost), C.byref(nres))
    return bool(ok), float(cost.value), int(nres.value)


def sampled_covariance(sets, poses, cfg, final_cost, num_residuals, xy_range=0.4, yaw_range=0.0043625, samples_per_axis=3,
                       covariance_scaler=4.0):
    """OdometryKeyframeFuser::approximateCovarianceBySampling (odometrykeyframefuser.cpp:261-380): sample GetCost on a
    samples_per_axis^3 grid around poses[-1] (yaw-major, then x, then y), least-squares quadric fit (the reference's
    bdcSvd solve == numpy lstsq), Hessian convexity check, cov = 2 H^-1 * cov_scale * covariance_scaler with
    cov_scale = final_cost / (num_residuals - 3) of the preceding Register (GetCovarianceScaler, n_scan_normal.cpp:435-441).
    Returns (success, cov6x6, samples [n^3, 4] = x, y, yaw, cost)."""
    poses = np.asarray(poses, dtype=np.float64)
    xs = np.linspace(-xy_range * 0.5, xy_range * 0.5, samples_per_axis)
    ts = np.linspace(-yaw_range * 0.5, yaw_range * 0.5, samples_per_axis)
    rows = []
    for t in ts:
        for x in xs:
            for y in xs:
                p = poses.copy()
                p[-1] = [poses[-1, 0] + x, poses[-1, 1] + y, poses[-1, 2] + t]
                _, c, _ = get_cost(sets, p, cfg)
                rows.append((x, y, t, c))
    S = np.array(rows)
    x, y, z, c = S.T
    A = np.stack([x * x, y * y, z * z, x * y, y * z, z * x, x, y, z, np.ones_like(x)], 1)
    q = np.linalg.lstsq(A, c, rcond=None)[0]
    H = np.array([[2 * q[0], q[3], q[5]], [q[3], 2 * q[1], q[4]], [q[5], q[4], 2 * q[2]]])
    cov6 = np.eye(6)
    if np.any(np.linalg.eigvalsh(H) <= 0.0) or num_residuals - 3 == 0:
        return False, cov6, S
    c3 = 2.0 * np.linalg.inv(H) * (final_cost / (num_residuals - 3)) * covariance_scaler
    cov6[0:2, 0:2] = c3[0:2, 0:2]; cov6[5, 5] = c3[2, 2]
    cov6[0, 5] = c3[0, 2]; cov6[1, 5] = c3[1, 2]; cov6[5, 0] = c3[2, 0]; cov6[5, 1] = c3[2, 1]
    return True, cov6, S


def eval_cost(cfg, res8, x):
    ci, cd = cfg
    res8 = np.ascontiguousarray(res8, dtype=np.float64)
    H = np.zeros(6); g = np.zeros(3)
    c = lib().orc_eval_cost(_p(ci), _p(cd), res8.shape[0], _p(res8), _p(np.ascontiguousarray(x, dtype=np.float64)), _p(H), _p(g))
    return c, H, g


def pipeline_batch(polar, mot, kf_sets, kf_ids, poses, cfg, *, k=12, z_min=60, min_distance=2.5, range_res=0.0438,
                   radius=3.5, weight_intensity=True, compensate=True, ccw=False, nthreads=1):
    """Whole per-scan path for nprob independent scans.  polar (nprob,A,R) u8; kf_sets: list of cell dicts;
    kf_ids (nprob,K) into kf_sets; poses (nprob,K+1,3)."""
    polar = np.ascontiguousarray(polar, dtype=np.uint8)
    nprob, A, R = polar.shape
    kf_ids = np.ascontiguousarray(kf_ids, dtype=np.int32)
    K = kf_ids.shape[1]
    ci, cd = cfg
    pipe_i = np.array([A, R, k, int(z_min), int(weight_intensity), int(compensate), int(ccw), K], np.int32)
    pipe_f = np.array([min_distance, range_res, radius], np.float32)
    offs, mean, normal, cov, plan, ns = concat_cellsets(kf_sets)
    p = np.ascontiguousarray(poses, dtype=np.float64).copy()
    cov36 = np.zeros((nprob, 36))
    stats = (RegStats * nprob)()
    ncells = np.zeros(nprob, np.int32); npts = np.zeros(nprob, np.int32)
    stage_ms = np.zeros(3)
    m = np.ascontiguousarray(mot, dtype=np.float64)
    lib().orc_pipeline_batch(int(nthreads), nprob, _p(pipe_i), _p(pipe_f), _p(ci), _p(cd), _p(polar), _p(m),
                             _p(kf_ids), _p(offs), _p(mean), _p(normal), _p(cov), _p(plan), _p(ns),
                             _p(p), _p(cov36), C.byref(stats), _p(ncells), _p(npts), _p(stage_ms))
    return dict(poses=p, cov=cov36.reshape(nprob, 6, 6), stats=list(stats), ncells=ncells, npts=npts, stage_ms=stage_ms)


def odometry_sequence(polar, cfg, *, k=12, z_min=60, min_distance=2.5, range_res=0.0438, radius=3.5, weight_intensity=True,
                      compensate=True, ccw=False, submap_scan_size=3, use_guess=True, min_keyframe_dist=1.5,
                      min_keyframe_rot_deg=5.0):
    """OdometryKeyframeFuser replay of one sequence of polar images (nscans,A,R).  Returns dict(poses [n,3], keyframe, stats, ncells)."""
    polar = np.ascontiguousarray(polar, dtype=np.uint8)
    n, A, R = polar.shape
    ci, cd = cfg
    pipe_i = np.array([A, R, k, int(z_min), int(weight_intensity), int(compensate), int(ccw), submap_scan_size, int(use_guess)], np.int32)
    pipe_f = np.array([min_distance, range_res, radius], np.float32)
    kf_d = np.array([min_keyframe_dist, min_keyframe_rot_deg], np.float64)
    poses = np.zeros((n, 3)); kf = np.zeros(n, np.int32); ncells = np.zeros(n, np.int32)
    stats = (RegStats * n)()
    lib().orc_odometry_sequence(n, _p(pipe_i), _p(pipe_f), _p(kf_d), _p(ci), _p(cd), _p(polar), _p(poses), _p(kf), C.byref(stats), _p(ncells))
    return dict(poses=poses, keyframe=kf, stats=list(stats), ncells=ncells)


def cfar(img, window_size=10, false_alarm_rate=0.01, nb_guard_cells=20, range_res=0.0438, z_min=60.0, min_distance=2.5,
         max_distance=400.0):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    A, R = img.shape
    out = np.zeros((A * R, 4), np.float32)
    n = lib().orc_cfar(_p(img), A, R, int(window_size), C.c_double(false_alarm_rate), int(nb_guard_cells), C.c_float(range_res),
                       C.c_float(z_min), C.c_float(min_distance), C.c_double(max_distance), _p(out), out.shape[0])
    return out[:n].copy()
