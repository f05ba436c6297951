#!/usr/bin/env python
"""Writes tests/golden/cfear_golden_v1.npz and cfear_golden_v2.npz: frozen outputs of the CPU oracle on seeded synthetic problems
(v1: filter / cloud / surface points / P2L registration; v2: P2D registration against two keyframes with covariance, GetCost,
covariance by sampling, and a short OdometryKeyframeFuser replay).

The reference ships no tests or golden vectors and cannot be built here (ROS/PCL/Ceres absent), so these pin the
ORACLE (parity unpinned with respect to the reference itself; see oracle/cfear_oracle.cc).  Regenerate only when the
oracle's contract changes:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as orc  # noqa: E402
from cfear_radarodometry_code_public_b200 import synth  # noqa: E402

seed = 42
img, tp = synth.make_problem_images(seed, 1)
idx, cnt = orc.kstrongest(img[1], 60, 12)
cl = orc.cloud(img[1], idx, cnt)
sp = orc.surface_points(cl, 3.5, True)
i0, c0 = orc.kstrongest(img[0], 60, 12)
sp0 = orc.surface_points(orc.cloud(img[0], i0, c0), 3.5, True)
P = tp.copy(); P[1] = tp[0]
ok, op, cov, st, _ = orc.register([sp0, sp], P, orc.reg_cfg(cost="P2L"))
assert ok
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cfear_golden_v1.npz"), seed=seed, img_sub=img[1, ::8, ::8],
                    kidx=idx, kcnt=cnt, cloud=cl, nsamples=sp["nsamples"], mean=sp["mean"], normal=sp["normal"],
                    poses_in=P, poses_out=op, outer=st.outer_iterations, inner=st.inner_iterations)
print("wrote golden: cells", sp["mean"].shape[0], "pose", op[1], "outer", st.outer_iterations, "inner", st.inner_iterations)


# ---- v2: the rows added after v1 (P2D + covariance, GetCost, covariance by sampling, sequence replay) ----------------
K = 2
img2, tp2 = synth.make_problem_images(43, K)
sets = []
for i in range(K + 1):
    ii, cc = orc.kstrongest(img2[i], 60, 12)
    sets.append(orc.surface_points(orc.cloud(img2[i], ii, cc), 3.0, True))
P2 = tp2.copy(); P2[K] = tp2[K - 1]
cfg2 = orc.reg_cfg(cost="P2D", loss="Huber", weight_opt=4, regularization=0.1)
ok2, op2, cov2, st2, _ = orc.register(sets, P2, cfg2)
assert ok2
gok, gcost, gnres = orc.get_cost(sets, op2, cfg2)
sok, scov, S = orc.sampled_covariance(sets, op2, cfg2, st2.final_cost, st2.num_residuals)
assert gok and sok
seq, _ = synth.make_sequence(11, 8)
rep = orc.odometry_sequence(seq, orc.reg_cfg(cost="P2L", weight_opt=0), radius=3.5, weight_intensity=True, submap_scan_size=3)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cfear_golden_v2.npz"), seed=43, seq_seed=11,
                    img_sub=img2[K, ::8, ::8], poses_in=P2, poses_out=op2, cov=cov2, outer=st2.outer_iterations,
                    inner=st2.inner_iterations, final_cost=st2.final_cost, num_residuals=st2.num_residuals,
                    get_cost=gcost, get_cost_nres=gnres, sampled_cov=scov, samples=S,
                    seq_sub=seq[-1, ::8, ::8], seq_poses=rep["poses"], seq_keyframe=rep["keyframe"])
print("wrote golden v2: pose", op2[K], "outer", st2.outer_iterations, "cost", st2.final_cost, "get_cost", gcost,
      "sampled cov diag", scov[0, 0], scov[1, 1], scov[5, 5], "seq end", rep["poses"][-1])
