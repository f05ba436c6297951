"""Repeat BASELINE configs[1] (3000-cell P2L gn_fixed registration with association tables) and report every run whose
result differs from the first one (debugging aid for a run-to-run difference seen once in the GPU suite)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cfear_radarodometry_code_public_b200 import capi, workload

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
solver = sys.argv[2] if len(sys.argv) > 2 else "gn_fixed"
sets, P, delta = workload.make_cellset_pair(3000, seed=0)
first = None
# something else on the device first, like the test that precedes it in the suite
b = workload.make_batch(8, 4, seed0=0, workers=4)
for it in range(n):
    if it % 4 == 0:
        cx = capi.Context(device=0, max_batch=8, max_cellsets=8 * 6, max_keyframes=4, **workload.CFEAR3)
        kf = np.arange(32, dtype=np.int32).reshape(8, 4)
        for i in range(4):
            cx.scans_to_cells_batch(b["kf_polar"][:, i], None, kf[:, i])
        cx.odometry_step_batch(b["polar"], b["mot"], kf, (32 + np.arange(8)).astype(np.int32), b["poses"])
        cx.close()
    c = capi.Context(max_batch=2, max_cellsets=4, max_keyframes=1, cost="P2L", loss="Huber", weight_opt=0,
                     solver_mode=solver, gn_iters=10, regularization=1.0)
    c.cells_upload(0, sets[0]); c.cells_upload(1, sets[1])
    gp, gcov, gst, gassoc = c.register_batch(np.array([[0, 1]], np.int32), P[None], want_assoc=True)
    cur = (gp.copy(), {k: gst[k][0] for k in gst.dtype.names}, gassoc.copy())
    if first is None:
        first = cur
        print("run 0:", cur[1], gp[0, 1])
    else:
        same = np.array_equal(cur[0], first[0]) and cur[1] == first[1] and np.array_equal(cur[2], first[2])
        if not same:
            print("run %d DIFFERS:" % it, cur[1], gp[0, 1], "assoc diffs", int((cur[2] != first[2]).sum()))
    c.close()
print("done", n)
