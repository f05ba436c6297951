"""A/B harness for kernel variants: per-stage CUDA-event times of the configs[2] step for one build of the library.

  python profiles/ab_stage.py --make /tmp/b.npz                      # generate the 256-problem batch once
  python profiles/ab_stage.py --lib path/to/libcfear_X.so --batch /tmp/b.npz [--ref /tmp/poses_ref.npy] [--prof]

Prints one line: K1 / K3 / K5 ms per step (device-resident, stage events on the library's stream), iteration statistics
and, with --ref, the largest pose / iteration-count difference against the poses another build saved.  --prof reads
the clock64 counters a -DCFEAR_K5_PROFILE build leaves in the covariance output.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cfear_radarodometry_code_public_b200 import capi, workload  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--make")
ap.add_argument("--lib")
ap.add_argument("--batch")
ap.add_argument("--ref")
ap.add_argument("--save")
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--nprob", type=int, default=256)
ap.add_argument("--prof", action="store_true")
args = ap.parse_args()
K = 4
if args.make:
    b = workload.make_batch(args.nprob, K, seed0=0)
    np.savez(args.make, **b)
    sys.exit(0)

capi.LIB_PATH = os.path.abspath(args.lib)
b = dict(np.load(args.batch))
nprob = b["polar"].shape[0]
ctx = capi.Context(device=0, max_batch=nprob, max_cellsets=nprob * (K + 1), max_keyframes=K, **workload.CFEAR3)
kf = np.arange(nprob * K, dtype=np.int32).reshape(nprob, K)
cur = (nprob * K + np.arange(nprob)).astype(np.int32)
for i in range(K):
    ctx.scans_to_cells_batch(b["kf_polar"][:, i], None, kf[:, i])


def dev(a):
    a = np.ascontiguousarray(a)
    p = ctx.dev_alloc(a.nbytes)
    ctx.h2d(p, a)
    return p


d_polar, d_mot, d_kf, d_cur = dev(b["polar"]), dev(b["mot"]), dev(kf), dev(cur)
d_poses = dev(b["poses"])
d_cov = dev(np.zeros((nprob, 36)))
d_stats = dev(np.zeros(nprob, capi.STATS_DTYPE))


def step():
    ctx.h2d(d_poses, b["poses"])
    ctx.odometry_step_batch_dev(nprob, d_polar, d_mot, d_kf, K, d_cur, d_poses, d_cov, d_stats)


for _ in range(5):
    step()
ctx.sync()
ctx.stage_timing(True)
for _ in range(args.steps):
    step()
ctx.sync()
n, ms = ctx.stage_timing(False)
poses = np.zeros((nprob, K + 1, 3)); st = np.zeros(nprob, capi.STATS_DTYPE); cov = np.zeros((nprob, 36))
ctx.d2h(poses, d_poses); ctx.d2h(st, d_stats); ctx.d2h(cov, d_cov)
line = "%-28s K1 %.4f  K3 %.4f  K5 %.4f ms/step (n=%d) | outer %.2f inner %.2f blocks %.0f success %.3f" % (
    os.path.basename(args.lib), ms[0] / n, ms[1] / n, ms[2] / n, n, st["outer_iterations"].mean(),
    st["inner_iterations"].mean(), st["num_blocks"].mean(), st["success"].mean())
if args.ref and os.path.exists(args.ref):
    r = np.load(args.ref)
    d = poses[:, K] - r["poses"][:, K]
    line += " | vs ref: dpos %.2e drot %.2e outer_diff %d inner_diff %d" % (
        np.hypot(d[:, 0], d[:, 1]).max(), np.abs(d[:, 2]).max(),
        int((st["outer_iterations"] != r["outer"]).sum()), int((st["inner_iterations"] != r["inner"]).sum()))
if args.save:
    np.savez(args.save, poses=poses, outer=st["outer_iterations"], inner=st["inner_iterations"])
print(line, flush=True)
if args.prof:
    assoc, solve, stage, total = cov[:, 13], cov[:, 14], cov[:, 16], cov[:, 15]
    ev = st["inner_iterations"] + st["outer_iterations"]      # ~ evaluations per problem (one per LM iteration + the initial ones)
    print("  clock64 per problem: total %.0f (max %.0f)  association %.0f (%.0f / outer)  solve %.0f (%.0f / eval)  grid staging %.0f  rest %.0f"
          % (total.mean(), total.max(), assoc.mean(), (assoc / st["outer_iterations"]).mean(), solve.mean(),
             (solve / ev).mean(), stage.mean(), (total - assoc - solve - stage).mean()), flush=True)
    nev = cov[:, 11]
    print("  warp 0 per evaluation (%.1f evals): sincos+publish %.0f  own share %.0f  wait+barrier %.0f  -> scalar logic between evaluations %.0f"
          % (nev.mean(), (cov[:, 8] / nev).mean(), (cov[:, 9] / nev).mean(), (cov[:, 10] / nev).mean(),
             ((solve - cov[:, 8] - cov[:, 9] - cov[:, 10]) / nev).mean()), flush=True)
    no = st["outer_iterations"]
    print("  association per outer iteration: phase 1 %.0f (+wait %.0f)  scan+list %.0f  phase 2 %.0f (+wait %.0f)"
          % ((cov[:, 17] / no).mean(), (cov[:, 18] / no).mean(), (cov[:, 19] / no).mean(), (cov[:, 20] / no).mean(),
             (cov[:, 22] / no).mean()), flush=True)
if args.prof:
    npair = np.maximum(cov[:, 26], 1)
    print("  phase 1 per pair (thread 0, %.1f pairs per problem): index+transform %.0f  nn_query %.0f  gate+store %.0f"
          % (npair.mean(), (cov[:, 23] / npair).mean(), (cov[:, 24] / npair).mean(), (cov[:, 25] / npair).mean()), flush=True)
ctx.close()
