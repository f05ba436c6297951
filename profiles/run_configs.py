#!/usr/bin/env python
"""Numbers for the BASELINE.json configurations that bench.py does not headline (bench.py = configs[2]).

  configs[0]  single synthetic 400x3360 scan, k=12 filter + cloud + surface points (r=3.5) on the CPU oracle: ms per stage
  configs[1]  scan-to-1-keyframe P2L registration on ~3000-cell sets: `gn_fixed` with 10 iterations (and the ceres_lm loop),
              a batch of independent problems on one GPU: us per problem, pose vs the oracle in the same mode
  configs[3]  needs the Oxford dataset (not in the image): prints "not run"; the same pipeline runs on synthetic Oxford-format
              PNGs in tests/test_gpu_mirror.py::test_oxford_png_directory_to_trajectory_diff
  configs[4]  is replay.py (run under torch.distributed.run for 2 / 4 / 8 GPUs)

One JSON line per configuration on stdout.   python profiles/run_configs.py [--nprob 256]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as orc  # noqa: E402
from cfear_radarodometry_code_public_b200 import synth, workload  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nprob", type=int, default=256)
ap.add_argument("--no-gpu", action="store_true")
a = ap.parse_args()
orc.build()

# ---- configs[0] ------------------------------------------------------------------------------------------------------
img = synth.make_problem_images(0, 0)[0][0]
ts = {"filter": [], "cloud": [], "surface_points": []}
for _ in range(30):
    t0 = time.perf_counter(); idx, cnt = orc.kstrongest(img, 60, 12)
    t1 = time.perf_counter(); cl = orc.cloud(img, idx, cnt)
    t2 = time.perf_counter(); sp = orc.surface_points(cl, 3.5, True)
    t3 = time.perf_counter()
    ts["filter"].append(t1 - t0); ts["cloud"].append(t2 - t1); ts["surface_points"].append(t3 - t2)
print(json.dumps({"config": "configs[0]: single synthetic 400x3360 scan (seed 0), k=12, z_min=60, r=3.5, CPU oracle, one thread",
                  "ms_median": {k: 1e3 * float(np.median(v[5:])) for k, v in ts.items()}, "n_pts": int(cl.shape[0]),
                  "n_cells": int(sp["mean"].shape[0]), "host_cores_available": os.cpu_count()}), flush=True)

# ---- configs[1] ------------------------------------------------------------------------------------------------------
if not a.no_gpu:
    from cfear_radarodometry_code_public_b200 import capi
    nprob = a.nprob
    nseed = 8                                           # distinct worlds, cycled over the batch
    pairs = [workload.make_cellset_pair(3000, seed=s) for s in range(nseed)]
    for solver in ("gn_fixed", "ceres_lm"):
        c = capi.Context(max_batch=nprob, max_cellsets=2 * nseed, max_keyframes=1, cost="P2L", loss="Huber", weight_opt=0,
                         solver_mode=solver, gn_iters=10)
        for s, (sets, _, _) in enumerate(pairs):
            c.cells_upload(2 * s, sets[0]); c.cells_upload(2 * s + 1, sets[1])
        slots = np.array([[2 * (b % nseed), 2 * (b % nseed) + 1] for b in range(nprob)], np.int32)
        P = np.zeros((nprob, 2, 3))
        for _ in range(3):
            gp, gcov, gst, _ = c.register_batch(slots, P)
        reps, t0 = 20, time.perf_counter()
        for _ in range(reps):
            gp, gcov, gst, _ = c.register_batch(slots, P)
        el = time.perf_counter() - t0
        t1 = time.perf_counter()
        refs = [orc.register(pairs[s][0], np.zeros((2, 3)), orc.reg_cfg(cost="P2L", loss="Huber", weight_opt=0, solver_mode=capi.SOLVER[solver], gn_iters=10))
                for s in range(nseed)]
        cpu_el = (time.perf_counter() - t1) / nseed
        dpos = max(float(np.hypot(*(gp[b, 1, :2] - refs[b % nseed][1][1, :2]))) for b in range(nprob))
        drot = max(float(abs(gp[b, 1, 2] - refs[b % nseed][1][1, 2])) for b in range(nprob))
        terr = max(float(np.hypot(*(gp[b, 1, :2] - pairs[b % nseed][2][:2]))) for b in range(nprob))
        print(json.dumps({"config": f"configs[1]: scan-to-1-keyframe P2L, Huber 0.1, ~3000-cell sets, offset (0.5 m, 0.2 m, 2 deg), identity guess, {solver}"
                                    + (" with 10 iterations" if solver == "gn_fixed" else ""),
                          "problems_per_launch": nprob, "us_per_problem_gpu": 1e6 * el / (reps * nprob), "ms_per_launch": 1e3 * el / reps,
                          "api": "cfear_register_batch (host poses in / out inside the timed call)",
                          "us_per_problem_cpu_oracle_1_thread": 1e6 * cpu_el,
                          "max_pos_diff_vs_oracle_m": dpos, "max_rot_diff_vs_oracle_rad": drot, "max_pos_err_vs_true_offset_m": terr,
                          "outer_iterations": [int(gst["outer_iterations"][b]) for b in range(nseed)],
                          "outer_iterations_oracle": [int(r[3].outer_iterations) for r in refs],
                          "inner_iterations": [int(gst["inner_iterations"][b]) for b in range(nseed)],
                          "inner_iterations_oracle": [int(r[3].inner_iterations) for r in refs],
                          "residuals_mean": float(gst["num_residuals"].mean())}), flush=True)
        c.close()

# ---- the reference's own execution model: ONE sequence, one scan at a time, through the drop-in classes -----------------
if not a.no_gpu:
    import re
    import subprocess
    import tempfile
    from cfear_radarodometry_code_public_b200 import io as cio
    n = 60
    imgs, _ = synth.make_sequence(5, n)
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "offline_odometry")
        pkg = os.path.join(ROOT, "cfear_radarodometry_code_public_b200")
        subprocess.check_call(["g++", "-std=c++14", "-O2", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "offline_odometry.cpp"),
                               "-o", exe, "-L" + pkg, "-lcfear_b200", "-Wl,-rpath," + pkg])
        frames = os.path.join(td, "seq.cfrs")
        cio.write_frames(frames, imgs)
        out = subprocess.check_output([exe, "--frames", frames, "--est_directory", td, "--cost_type", "P2D", "--res", "3.0", "--submap_scan_size", "4",
                                       "--z-min", "60", "--weight_option", "4", "--regularization", "0.1"]).decode()
        dur = np.array([float(x) for x in re.findall(r"dur: ([0-9.eE+-]+)", out)])
        rows = np.loadtxt(os.path.join(td, "01.txt"))
    t0 = time.perf_counter()
    ref = orc.odometry_sequence(imgs, orc.reg_cfg(cost="P2D", weight_opt=4, regularization=0.1), z_min=60, radius=3.0, weight_intensity=True, submap_scan_size=4)
    cpu_ms = 1e3 * (time.perf_counter() - t0) / n
    d = np.hypot(rows[:, 3] - ref["poses"][:, 0], rows[:, 7] - ref["poses"][:, 1])
    print(json.dumps({"config": "single sequence, one scan at a time (the reference's execution model): examples/offline_odometry over the C++ mirror "
                                "(radarDriver::CallbackOffline -> OdometryKeyframeFuser::pointcloudCallback per frame, every call synchronous), "
                                f"{n} synthetic frames, P2D, window 4",
                      "ms_per_frame_gpu_median": 1e3 * float(np.median(dur[5:])), "frames_per_s_gpu": 1.0 / float(np.median(dur[5:])),
                      "ms_per_frame_cpu_oracle_1_thread": cpu_ms, "frames_per_s_cpu_oracle_1_thread": 1e3 / cpu_ms,
                      "max_pos_diff_vs_oracle_replay_m": float(d.max()), "sensor_rate_hz": 4}), flush=True)

print(json.dumps({"config": "configs[3]: Oxford 2019-01-10-12-32-52 full sequence replay, trajectory diff vs the reference's est/01.txt",
                  "status": "not run: neither the dataset nor the reference's est/01.txt exist in this image (no network)",
                  "how": "python -c 'from cfear_radarodometry_code_public_b200 import io ...load_oxford_png / write_frames' -> examples/offline_odometry "
                         "--frames seq.cfrs --est_directory est ... -> python tools/traj_diff.py est/01.txt <reference est/01.txt>",
                  "dry_run": "tests/test_gpu_mirror.py::test_oxford_png_directory_to_trajectory_diff (synthetic Oxford-format PNGs)"}), flush=True)
