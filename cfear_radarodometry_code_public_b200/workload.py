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
