#!/usr/bin/env python
"""bench.py -- radar scans/sec of the CFEAR per-scan hot path on B200 (BASELINE.json metric).

Workload (config.workload): BASELINE configs[2] -- "CFEAR-3-like batch": 256 independent synthetic 400x3360 polar
scans per GPU, each: k=12 k-strongest filter -> cloud -> Compensate -> oriented surface points (r=3.0) ->
registration against 4 resident keyframe cell sets (P2D, Huber 0.1, regularization 0.1, weight option 4,
Ceres-style LM loop).  One "step" = one pass of that path over the batch.

  value      device-timed: polar images already resident in HBM (344 MB per step per GPU > 126 MB L2)
  e2e        the same through cfear_odometry_step_batch_submit/_wait with pinned HOST buffers (two steps in flight):
             H2D of the images + D2H of the poses / covariances / stats of every step inside the timed region
  roofline   dominant kernel's algorithmic bytes / its CUDA-event duration vs the measured HBM copy bandwidth
  cpu_baseline / --impl reference: the CPU oracle port (oracle/cfear_oracle.cc; the reference itself needs
             ROS+PCL+Ceres and cannot be built here) on the host cores, bounded sample of the same workload

Launch:  python bench.py --gpus N --steps K --warmup W      (N>1: under torch.distributed.run, one rank per GPU)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NPROB, K = 256, 4
A, R, KS = 400, 3360, 12
METRIC = "radar scans/sec (400x3360 polar, k=12, 4 keyframes)"
WORKLOAD = ("configs[2]: CFEAR-3-like batch, 256 independent synthetic 400x3360 scans per GPU, k=12 filter + "
            "surface points r=3.0 + scan-to-4-keyframes P2D (Huber 0.1, reg 0.1, weight_opt 4), ceres_lm solver")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "50", "-i", str(dev)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self, t0, t1):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        rows, inreg = [], []
        import datetime
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                sm, smax = float(parts[1]), float(parts[2])
            except Exception:
                continue
            row = (ts, sm, smax, parts[4:8])
            rows.append(row)
            if t0 - 0.05 <= ts <= t1 + 0.05:
                inreg.append(row)
        use = inreg if inreg else rows
        if use:
            out["sm_mhz"] = float(np.median([r[1] for r in use]))
            out["sm_max_mhz"] = float(max(r[2] for r in use))
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            out["reasons"] = sorted({n for r in use for n, v in zip(names, r[3]) if v.lower().startswith("active")})
            out["samples"] = len(use)
        try:
            os.unlink(self.f.name)
        except Exception:
            pass
        return out


def oracle_cfg(orc):
    return orc.reg_cfg(cost="P2D", loss="Huber", loss_limit=0.1, weight_opt=4, regularization=0.1, cov_scale=1.0)


def cpu_leg(batch, nsample, min_seconds, steps=None, warmup=0):
    """Times the CPU oracle port on the first `nsample` problems with all host threads.
    Returns (scans_per_s, threads, seconds, reps, poses of the sample)."""
    import oracle as orc
    orc.build()
    threads = os.cpu_count() or 1
    cfg = oracle_cfg(orc)
    sl = slice(0, nsample)
    # keyframe cell sets (untimed set-up, like the resident keyframes of the GPU arm)
    kf_sets, kf_ids = [], np.zeros((nsample, K), np.int32)
    for b in range(nsample):
        for i in range(K):
            idx, cnt = orc.kstrongest(batch["kf_polar"][b, i], 60, KS)
            cl = orc.cloud(batch["kf_polar"][b, i], idx, cnt)
            kf_ids[b, i] = len(kf_sets)
            kf_sets.append(orc.surface_points(cl, 3.0, True))
    run = lambda: orc.pipeline_batch(batch["polar"][sl], batch["mot"][sl], kf_sets, kf_ids, batch["poses"][sl], cfg,
                                     k=KS, z_min=60, radius=3.0, weight_intensity=True, compensate=True, nthreads=threads)
    for _ in range(max(warmup, 1)):
        out = run()
    reps, t0 = 0, time.perf_counter()
    while True:
        out = run()
        reps += 1
        el = time.perf_counter() - t0
        if steps is not None:
            if reps >= steps:
                break
        elif el >= min_seconds:
            break
    return nsample * reps / el, threads, el, reps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nprob", type=int, default=NPROB)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    nprob = args.nprob
    warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    from cfear_radarodometry_code_public_b200 import workload
    config = {"workload": WORKLOAD, "scans_per_gpu_per_step": nprob, "keyframes": K, "azimuths": A, "range_bins": R,
              "k_strongest": KS, "l2": "inputs larger than L2: %.0f MB of polar images per step per GPU" % (nprob * A * R / 1e6),
              "parallelism": "one independent batch per GPU, no data-path collective; one NCCL all_gather of the poses"}

    if args.impl == "reference":
        if rank != 0:
            return
        nsample = min(64, nprob)
        batch = workload.make_batch(nsample, K, seed0=0)
        sps, threads, el, reps, _ = cpu_leg(batch, nsample, 0.0, steps=max(args.steps, 1), warmup=args.warmup)
        sample = f"{nsample} scans of the workload per step (first {nsample} problems, seeds 0..{nsample - 1}), {threads} host threads"
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": sps, "unit": "scans/s", "n_gpus": args.gpus,
                          "steps": reps, "warmup": args.warmup, "ms_per_step": 1e3 * el / reps, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": sps, "unit": "scans/s", "cores": threads, "kind": "port", "sample": sample},
                          "e2e": {"value": sps, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ---- inputs (numpy, before CUDA is touched: the generator forks workers) ----
    batch = workload.make_batch(nprob, K, seed0=rank * nprob)

    import torch
    import torch.distributed as dist
    from cfear_radarodometry_code_public_b200 import capi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout while it initialises (NCCL_DEBUG=VERSION in some environments); stdout
        # must carry the one JSON line only, so fd 1 points at stderr until the communicator exists
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    ctx = capi.Context(device=local, max_batch=nprob, max_cellsets=nprob * (K + 1), max_keyframes=K, **workload.CFEAR3)
    ext = torch.cuda.ExternalStream(ctx.stream_ptr, device=dev)

    # resident keyframe cell sets, built by the GPU path from the keyframe images (untimed set-up)
    kf_slots = np.arange(nprob * K, dtype=np.int32).reshape(nprob, K)
    cur_slots = (nprob * K + np.arange(nprob)).astype(np.int32)
    for i in range(K):
        ctx.scans_to_cells_batch(batch["kf_polar"][:, i], None, kf_slots[:, i])

    # ---- device-resident arm ----
    t_polar = torch.from_numpy(batch["polar"]).to(dev)
    t_mot = torch.from_numpy(batch["mot"]).to(dev)
    t_kf = torch.from_numpy(kf_slots).to(dev)
    t_cur = torch.from_numpy(cur_slots).to(dev)
    t_poses0 = torch.from_numpy(batch["poses"]).to(dev)
    t_poses = t_poses0.clone()
    t_cov = torch.zeros(nprob, 36, dtype=torch.float64, device=dev)
    t_stats = torch.zeros(nprob, capi.STATS_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    t_gather = torch.zeros(world * nprob, K + 1, 3, dtype=torch.float64, device=dev) if world > 1 else None
    torch.cuda.synchronize()

    def step_dev():
        with torch.cuda.stream(ext):
            t_poses.copy_(t_poses0, non_blocking=True)          # Register() works in/out on Tsrc: restore the guess
        ctx.odometry_step_batch_dev(nprob, t_polar.data_ptr(), t_mot.data_ptr(), t_kf.data_ptr(), K, t_cur.data_ptr(),
                                    t_poses.data_ptr(), t_cov.data_ptr(), t_stats.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step_dev()
    ctx.sync()
    npts, ncells = ctx.last_counts(cur_slots)
    _, kf_ncells = None, np.array([ctx.cells_count(s) for s in kf_slots[: min(nprob, 32)].ravel()])
    sampler = ClockSampler(local)
    time.sleep(0.3)
    ctx.stage_timing(True)
    l0 = ctx.launches
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.time()
    e0.record(ext)
    for _ in range(args.steps):
        step_dev()
    if world > 1:
        with torch.cuda.stream(ext):
            dist.all_gather_into_tensor(t_gather.view(-1), t_poses.view(-1))   # the path's only collective
    e1.record(ext)
    barrier()
    tw1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = ctx.launches - l0
    nst, stage_ms = ctx.stage_timing(False)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * nprob * args.steps / (ms * 1e-3)
    stats = np.frombuffer(t_stats.cpu().numpy().tobytes(), dtype=capi.STATS_DTYPE)
    poses_dev = t_poses.cpu().numpy()

    # ---- roofline of the dominant kernel (algorithmic bytes: SURVEY.md 8(d), per-kernel split in DESIGN.md) ----
    hbm_peak, peak_src = peaks()
    n_pts, n_cells = float(npts.mean()), float(ncells.mean())
    n_kf_cells = float(kf_ncells.mean())
    alg = {"k1_kstrongest": A * R + A * (KS * 4 + 4) + A * (KS * 16 + 4),
           "k3_surface_points": A * KS * 16 + A * 4 + n_cells * 80 + n_cells * 12,
           "k5_register": (K * n_kf_cells + n_cells) * 80 + K * n_kf_cells * 12 + (K + 1) * 24 + 36 * 8 + 40}
    names = ["k1_kstrongest", "k3_surface_points", "k5_register"]
    per_launch_ms = [m / max(nst, 1) for m in stage_ms]
    dom = int(np.argmax(per_launch_ms))
    ach = alg[names[dom]] * nprob / (per_launch_ms[dom] * 1e-3) / 1e9 if per_launch_ms[dom] > 0 else 0.0
    b_scan = A * R + 2 * n_pts * 16 + n_cells * 80 + (K * n_kf_cells + n_cells) * 80 + 24
    roof = {"bound": "hbm", "kernel": names[dom], "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
            "frac": ach / hbm_peak, "traffic": None, "peak_source": peak_src,
            "algorithmic_bytes_per_scan": alg[names[dom]],
            "stage_ms_per_step": dict(zip(names, per_launch_ms)),
            "stage_frac_of_hbm_peak": {n: (alg[n] * nprob / (t * 1e-3) / 1e9 / hbm_peak if t > 0 else None)
                                       for n, t in zip(names, per_launch_ms)},
            "whole_path": {"bytes_per_scan": b_scan, "achieved": b_scan * value / world / 1e9,
                           "frac": b_scan * value / world / 1e9 / hbm_peak}}
    tr = os.path.join(ROOT, "profiles", "traffic.json")        # dram bytes per launch from the committed ncu capture
    if os.path.exists(tr):
        try:
            roof["traffic"] = json.load(open(tr)).get(names[dom])
        except Exception:
            pass

    # ---- end-to-end arm: host buffers through the public C-ABI call ----
    e2e = None
    if not args.no_e2e:
        numa = capi.bind_to_device_numa(local) if world > 1 else {"numa_node": None, "reason": "single GPU"}
        h_polar = capi.pinned_array(batch["polar"].shape, np.uint8); h_polar[...] = batch["polar"]
        h_mot = capi.pinned_array(batch["mot"].shape, np.float64); h_mot[...] = batch["mot"]

        def out_set():
            return dict(poses=capi.pinned_array((nprob, K + 1, 3), np.float64), cov=capi.pinned_array((nprob, 36), np.float64),
                        stats=capi.pinned_array((nprob,), capi.STATS_DTYPE), npts=capi.pinned_array((nprob,), np.int32))
        # Two steps in flight: step i+1 is submitted (its host->device copies start) before step i is waited for, so the
        # PCIe link does not idle during the registration tail of step i.  Every step still moves all of its inputs from
        # pinned host memory and all of its results back; each step's results are complete at its wait.
        h_outs = [out_set(), out_set()]

        def run_e2e(nsteps):
            pending = None
            for i in range(nsteps):
                o = h_outs[i & 1]
                np.copyto(o["poses"], batch["poses"])               # Register() works in/out on Tsrc: restore the guess
                t = ctx.odometry_step_batch_submit(h_polar, h_mot, kf_slots, cur_slots, o)
                if pending is not None:
                    ctx.odometry_step_batch_wait(pending)
                pending = t
            ctx.odometry_step_batch_wait(pending)
            return h_outs[(nsteps - 1) & 1]

        run_e2e(warmup)
        barrier()
        t0 = time.perf_counter()
        h_out = run_e2e(args.steps)
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([el], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            el = float(t.item())
        h2d = h_polar.nbytes + batch["mot"].nbytes + kf_slots.nbytes + cur_slots.nbytes + batch["poses"].nbytes
        d2h = h_out["poses"].nbytes + h_out["cov"].nbytes + h_out["stats"].nbytes + h_out["npts"].nbytes
        e2e = {"value": world * nprob * args.steps / el, "unit": "scans/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * el / args.steps, "host_numa_binding": numa,
               "api": "cfear_odometry_step_batch_submit / _wait, two steps in flight, pinned host buffers"}
        assert np.allclose(h_out["poses"], poses_dev, atol=1e-12), "e2e and device-resident arms disagree"

    clocks = sampler.stop(tw0, time.time())      # samples taken during the device-resident and end-to-end timed regions

    # ---- CPU baseline (rank 0, N=1 only): the oracle port on the host cores, bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        nsample = min(64, nprob)
        sps, threads, el, reps, out = cpu_leg(batch, nsample, args.cpu_seconds)
        d = poses_dev[:nsample, K] - out["poses"][:, K]
        cpu = {"value": sps, "unit": "scans/s", "cores": threads, "kind": "port",
               "sample": f"first {nsample} scans of the workload x {reps} passes ({el:.1f} s), {threads} host threads; "
                         "oracle/cfear_oracle.cc (the ROS/PCL/Ceres reference cannot be built here)",
               "parity_vs_gpu": {"max_pos_err_m": float(np.hypot(d[:, 0], d[:, 1]).max()), "max_rot_err_rad": float(np.abs(d[:, 2]).max())}}

    if rank == 0:
        err = poses_dev[:, K] - batch["truth"]
        line = {"metric": METRIC, "value": value, "unit": "scans/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config, "roofline": roof, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": int(launches), "clocks": clocks,
                "workload_stats": {"n_pts_mean": n_pts, "n_cells_mean": n_cells, "kf_cells_mean": n_kf_cells,
                                   "outer_iterations_mean": float(stats["outer_iterations"].mean()),
                                   "inner_iterations_mean": float(stats["inner_iterations"].mean()),
                                   "residuals_mean": float(stats["num_residuals"].mean()),
                                   "success_frac": float(stats["success"].mean()),
                                   "median_pos_err_vs_truth_m": float(np.median(np.hypot(err[:, 0], err[:, 1])))}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
