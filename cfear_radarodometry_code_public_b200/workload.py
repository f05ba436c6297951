"""Seeded synthetic workloads of BASELINE.json's configs (inputs only; no CFEAR arithmetic here).

config 3 ("CFEAR-3-like batch"): nprob independent (scan, K keyframes) problems.  Problem b renders K+1 polar
images of world `seed0+b` along the synthetic trajectory (synth.trajectory); images 0..K-1 are the keyframes
(known poses), image K is the current scan; the registration guess is the previous pose (constant-position prior,
~2.5 m / ~1.4 deg off) and the compensation motion is the previous inter-scan motion.
"""
from __future__ import annotations

import multiprocessing as mp
import os

import numpy as np

from . import synth

CFEAR3 = dict(k_strongest=12, z_min=60.0, range_res=0.0438, min_distance=2.5, radius=3.0, weight_intensity=1,
              compensate=1, cost="P2D", loss="Huber", loss_limit=0.1, weight_opt=4, regularization=0.1, cov_scale=1.0,
              solver_mode="ceres_lm")


def _one(args):
    seed, K = args
    imgs, poses = synth.make_problem_images(seed, K)
    return imgs, poses


def _cuda_live() -> bool:
    import sys
    t = sys.modules.get("torch")
    try:
        if t is not None and t.cuda.is_initialized():
            return True
    except Exception:
        pass
    from . import capi
    return capi._lib is not None          # the C-ABI library (and with it a CUDA context) has been loaded


def make_batch(nprob: int, K: int = 4, seed0: int = 0, workers: int | None = None):
    """Returns dict(kf_polar [nprob,K,A,R] u8, polar [nprob,A,R] u8, poses [nprob,K+1,3] (last = guess),
    truth [nprob,3], mot [nprob,3])."""
    if workers is None:
        workers = min(nprob, max(1, (os.cpu_count() or 1)))
    jobs = [(seed0 + b, K) for b in range(nprob)]
    if workers > 1 and _cuda_live():
        # forking a process that already holds CUDA / its helper threads is unsafe: use threads (numpy releases the GIL)
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(workers) as ex:
            res = list(ex.map(_one, jobs))
    elif workers > 1:
        with mp.get_context("fork").Pool(workers) as pool:
            res = pool.map(_one, jobs, chunksize=max(1, nprob // (4 * workers)))
    else:
        res = [_one(j) for j in jobs]
    A, R = res[0][0].shape[1:]
    kf_polar = np.empty((nprob, K, A, R), np.uint8)
    polar = np.empty((nprob, A, R), np.uint8)
    poses = np.empty((nprob, K + 1, 3))
    truth = np.empty((nprob, 3))
    mot = np.zeros((nprob, 3))
    for b, (imgs, tp) in enumerate(res):
        kf_polar[b] = imgs[:K]
        polar[b] = imgs[K]
        poses[b] = tp
        truth[b] = tp[K]
        poses[b, K] = tp[K - 1]
        if K >= 2:
            mot[b] = synth.se2_mul(synth.se2_inv(tp[K - 2]), tp[K - 1])
        else:
            mot[b] = synth.se2_mul(synth.se2_inv(tp[K - 1]), tp[K])
    return dict(kf_polar=kf_polar, polar=polar, poses=poses, truth=truth, mot=mot)


# ---- config 2 ("registration micro"): two ~3000-cell sets of one world, scan observed from a pose offset ------------
def _sample_cells(world: np.ndarray, n: int, rng, pose, max_range: float = 140.0, sigma: float = 0.05):
    """n oriented surface points of `world` as seen from `pose` (x, y, yaw), in the sensor frame: dict in the layout
    capi.Context.cells_upload / the oracle take.  Input generator only: the statistics are drawn, not computed."""
    p0, p1 = world[:, 0:2], world[:, 2:4]
    mid = 0.5 * (p0 + p1)
    near = np.hypot(mid[:, 0] - pose[0], mid[:, 1] - pose[1]) < max_range
    p0, p1 = p0[near], p1[near]
    ln = np.hypot(*(p1 - p0).T)
    seg = rng.choice(p0.shape[0], size=n, p=ln / ln.sum())
    u = rng.uniform(0, 1, n)[:, None]
    pts = p0[seg] * (1 - u) + p1[seg] * u
    tang = (p1[seg] - p0[seg]) / ln[seg, None]
    c, s = np.cos(pose[2]), np.sin(pose[2])
    Rinv = np.array([[c, s], [-s, c]])
    mean = (pts - np.asarray(pose[:2])) @ Rinv.T + rng.normal(0, sigma, (n, 2))
    t_s = tang @ Rinv.T
    ang = rng.normal(0, 0.03, n)                                           # ~2 deg of normal noise
    nrm = np.stack([-t_s[:, 1] * np.cos(ang) - t_s[:, 0] * np.sin(ang), t_s[:, 0] * np.cos(ang) - t_s[:, 1] * np.sin(ang)], 1)
    flip = (nrm * (-mean)).sum(1) < 0                                      # toward the sensor, like cell::ComputeNormal
    nrm[flip] *= -1
    lmin, lmax = rng.uniform(0.01, 0.05, n), rng.uniform(0.3, 1.5, n)
    tx, ty = -nrm[:, 1], nrm[:, 0]
    cov = np.empty((n, 2, 2))
    cov[:, 0, 0] = lmin * nrm[:, 0] ** 2 + lmax * tx ** 2
    cov[:, 0, 1] = cov[:, 1, 0] = lmin * nrm[:, 0] * nrm[:, 1] + lmax * tx * ty
    cov[:, 1, 1] = lmin * nrm[:, 1] ** 2 + lmax * ty ** 2
    return dict(mean=mean, normal=nrm, cov=cov, planarity=np.log(1.0 + (lmax / lmin) / 2.0),
                nsamples=rng.integers(6, 40, n).astype(np.int32), avg_intensity=rng.uniform(5, 90, n))


def make_cellset_pair(n_cells: int = 3000, seed: int = 0, delta=(0.5, 0.2, np.deg2rad(2.0))):
    """BASELINE configs[1]: (keyframe set at the identity, scan set observed from `delta`), ~n_cells cells each, plus the
    pose table [[0,0,0],[0,0,0]] (identity guess).  The registration should recover `delta`."""
    world = synth.make_world(seed)
    rng = np.random.Generator(np.random.PCG64(seed + 31337))
    kf = _sample_cells(world, n_cells, rng, (0.0, 0.0, 0.0))
    cur = _sample_cells(world, n_cells, rng, delta)
    return [kf, cur], np.zeros((2, 3)), np.asarray(delta, dtype=np.float64)
