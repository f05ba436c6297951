#!/usr/bin/env python
"""Writes tests/golden/cfear_golden_v1.npz: frozen outputs of the CPU oracle on one seeded synthetic problem.

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
