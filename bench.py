#!/usr/bin/env python
"""bench.py -- radar scans/sec of the CFEAR per-scan hot path on B200 (BASELINE.json metric).

Workload (config.workload): BASELINE configs[2] -- "CFEAR-3-like batch": 256 independent synthetic 400x3360 polar
scans per GPU, each: k=12 k-strongest filter -> cloud -> Compensate -> oriented surface points (r=3.0) ->
registration against 4 resident keyframe cell sets (P2D, Huber 0.1, regularization 0.1, weight option 4,
Ceres-style LM loop).  One "step" = one pass of that path over the batch.

  value      device-timed: polar images already resident in HBM (344 MB per step per GPU > 126 MB L2); consecutive steps
             are submitted through cfear_odometry_step_batch_dev_submit (--inflight steps on the library's own streams,
             each step's K1 -> K3 -> K5 chain intact, results double-buffered) -- `--serial` times the stream-ordered call
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


def cpu_leg(batch, nsample, min_seconds, steps=None, warmup=0, threads=None):
    """Times the CPU oracle port on the first `nsample` problems with `threads` host threads (default: all).
    Returns (scans_per_s, threads, seconds, reps, output of the last pass, per-scan stage ms [filter, build_normals, register])."""
    import oracle as orc
    orc.build()
    threads = threads or (os.cpu_count() or 1)
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
    stage = np.zeros(3)
    while True:
        out = run()
        stage += out["stage_ms"]
        reps += 1
        el = time.perf_counter() - t0
        if steps is not None:
            if reps >= steps:
                break
        elif el >= min_seconds:
            break
    return nsample * reps / el, threads, el, reps, out, (stage / (reps * nsample)).tolist()


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
    ap.add_argument("--lib", default=None, help="experiment builds: path of another libcfear_b200.so (profiles/ab)")
    ap.add_argument("--batch-cache", default=None, help="npz file caching the generated workload between runs")
    ap.add_argument("--save-poses", default=None, help="experiment builds: save poses / iteration counts (npz)")
    ap.add_argument("--ref-poses", default=None, help="experiment builds: compare with the npz another build saved")
    ap.add_argument("--inflight", type=int, default=5, help="result / slot sets the overlapped steps rotate through")
    ap.add_argument("--serial", action="store_true", help="device-resident arm through the stream-ordered cfear_odometry_step_batch_dev")
    ap.add_argument("--min-seconds", type=float, default=0.25, help="the K-step timed region is repeated until it lasts this long")
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
        # one step = the same nprob scans one b200 step processes (seeds 0..nprob-1), all host threads
        batch = workload.make_batch(nprob, K, seed0=0)
        sps, threads, el, reps, _, stage = cpu_leg(batch, nprob, 0.0, steps=max(args.steps, 1), warmup=args.warmup)
        sample = (f"{nprob} scans per step = the whole per-GPU batch of the workload (seeds 0..{nprob - 1}), {threads} host threads; "
                  "oracle/cfear_oracle.cc, g++ -O3 -ffp-contract=off (the reference's own flags class, CMakeLists.txt:32-33: -O3, no -march); "
                  "the ROS/PCL/Ceres reference itself cannot be built here")
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": sps, "unit": "scans/s", "n_gpus": args.gpus,
                          "steps": reps, "warmup": args.warmup, "ms_per_step": 1e3 * el / reps, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": sps, "unit": "scans/s", "cores": threads, "kind": "port", "sample": sample,
                                           "stage_ms_per_scan_per_thread": dict(zip(["filter", "build_normals", "register"], stage))},
                          "e2e": {"value": sps, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ---- inputs (numpy, before CUDA is touched: the generator forks workers) ----
    cache = args.batch_cache and f"{args.batch_cache}.{rank}.{nprob}.npz"
    if cache and os.path.exists(cache):
        batch = dict(np.load(cache))
    else:
        batch = workload.make_batch(nprob, K, seed0=rank * nprob)
        if cache:
            np.savez(cache, **batch)

    import torch
    import torch.distributed as dist
    from cfear_radarodometry_code_public_b200 import capi
    if args.lib:
        capi.LIB_PATH = os.path.abspath(args.lib)
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

    NSETS = max(1, args.inflight)  # result / current-slot sets the overlapped steps rotate through
    ctx = capi.Context(device=local, max_batch=nprob, max_cellsets=nprob * (K + NSETS), max_keyframes=K, steps_in_flight=NSETS,
                       **workload.CFEAR3)
    ext = torch.cuda.ExternalStream(ctx.stream_ptr, device=dev)

    # resident keyframe cell sets, built by the GPU path from the keyframe images (untimed set-up)
    kf_slots = np.arange(nprob * K, dtype=np.int32).reshape(nprob, K)
    cur_sets = [(nprob * (K + j) + np.arange(nprob)).astype(np.int32) for j in range(NSETS)]
    cur_slots = cur_sets[0]
    for i in range(K):
        ctx.scans_to_cells_batch(batch["kf_polar"][:, i], None, kf_slots[:, i])

    # ---- device-resident arm ----
    t_polar = torch.from_numpy(batch["polar"]).to(dev)
    t_mot = torch.from_numpy(batch["mot"]).to(dev)
    t_kf = torch.from_numpy(kf_slots).to(dev)
    t_cur = [torch.from_numpy(cs).to(dev) for cs in cur_sets]
    t_poses0 = torch.from_numpy(batch["poses"]).to(dev)
    t_poses = [t_poses0.clone() for _ in range(NSETS)]
    t_cov = [torch.zeros(nprob, 36, dtype=torch.float64, device=dev) for _ in range(NSETS)]
    t_stats = [torch.zeros(nprob, capi.STATS_DTYPE.itemsize, dtype=torch.uint8, device=dev) for _ in range(NSETS)]
    t_gather = torch.zeros(world * nprob, K + 1, 3, dtype=torch.float64, device=dev) if world > 1 else None
    torch.cuda.synchronize()
    tickets = [None] * NSETS
    nstep = [0]

    def step_dev():
        j = nstep[0] % NSETS
        nstep[0] += 1
        if tickets[j] is not None:
            ctx.stream_wait_ticket(tickets[j])                  # the restore below overwrites that step's in/out pose table
            tickets[j] = None
        with torch.cuda.stream(ext):
            t_poses[j].copy_(t_poses0, non_blocking=True)       # Register() works in/out on Tsrc: restore the guess
        a = (nprob, t_polar.data_ptr(), t_mot.data_ptr(), t_kf.data_ptr(), K, t_cur[j].data_ptr(),
             t_poses[j].data_ptr(), t_cov[j].data_ptr(), t_stats[j].data_ptr())
        if args.serial:
            ctx.odometry_step_batch_dev(*a)
        else:
            tickets[j] = ctx.odometry_step_batch_dev_submit(*a)
        return j

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step_dev()
    ctx.sync()
    npts, ncells = ctx.last_counts(cur_sets[(nstep[0] - 1) % NSETS])
    _, kf_ncells = None, np.array([ctx.cells_count(s) for s in kf_slots[: min(nprob, 32)].ravel()])
    # the K-step region is repeated until the timed region lasts --min-seconds (a 20-step region is ~10 ms: too few clock samples)
    est = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    est[0].record(ext)
    for _ in range(4):
        step_dev()
    ctx.join()
    est[1].record(ext)
    torch.cuda.synchronize()
    ms_est = est[0].elapsed_time(est[1]) / 4
    regions = max(1, int(np.ceil(args.min_seconds * 1e3 / max(ms_est * args.steps, 1e-6))))
    regions = min(regions, max(1, 4096 // max(args.steps, 1)))          # stage-event capacity of the library
    total_steps = regions * args.steps
    sampler = ClockSampler(local)
    time.sleep(0.3)
    ctx.stage_timing(True)
    l0 = ctx.launches
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.time()
    e0.record(ext)
    for _ in range(total_steps):
        jlast = step_dev()
    ctx.join()                                                     # the context stream waits for the steps in flight
    if world > 1:
        with torch.cuda.stream(ext):
            dist.all_gather_into_tensor(t_gather.view(-1), t_poses[jlast].view(-1))   # the path's only collective
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
    value = world * nprob * total_steps / (ms * 1e-3)
    stats = np.frombuffer(t_stats[jlast].cpu().numpy().tobytes(), dtype=capi.STATS_DTYPE)
    poses_dev = t_poses[jlast].cpu().numpy()
    for j in range(NSETS):                                          # every set holds the same problems: identical results
        assert np.array_equal(t_poses[j].cpu().numpy(), poses_dev), "overlapped steps disagree between result sets"
    # the same kernels timed one at a time (stream-ordered call): what each costs when it has the GPU to itself
    ctx.stage_timing(True)
    for _ in range(100):
        with torch.cuda.stream(ext):
            t_poses[0].copy_(t_poses0, non_blocking=True)
        ctx.odometry_step_batch_dev(nprob, t_polar.data_ptr(), t_mot.data_ptr(), t_kf.data_ptr(), K, t_cur[0].data_ptr(),
                                    t_poses[0].data_ptr(), t_cov[0].data_ptr(), t_stats[0].data_ptr())
    ctx.sync()
    nser, stage_ser = ctx.stage_timing(False)
    # same poses up to rounding: the stream-ordered arm launches K5 in the form that suits a launch with the GPU to itself
    # (192 threads x 2 per SM at 256 problems, 384 x 1 for small --nprob), whose sums are formed in another order
    dser = float(np.abs(t_poses[0].cpu().numpy() - poses_dev).max())
    assert dser < 1e-11, "overlapped and stream-ordered steps disagree (%g)" % dser
    stats_ser = np.frombuffer(t_stats[0].cpu().numpy().tobytes(), dtype=capi.STATS_DTYPE)
    assert np.array_equal(stats_ser["outer_iterations"], stats["outer_iterations"]) and \
        np.array_equal(stats_ser["inner_iterations"], stats["inner_iterations"]), "iteration counts differ between the arms"

    # ---- roofline of the dominant kernel (algorithmic bytes: SURVEY.md 8(d), per-kernel split in DESIGN.md) ----
    hbm_peak, peak_src = peaks()
    n_pts, n_cells = float(npts.mean()), float(ncells.mean())
    n_kf_cells = float(kf_ncells.mean())
    alg = {"k1_kstrongest": A * R + A * (KS * 4 + 4) + A * (KS * 16 + 4),
           "k3_surface_points": A * KS * 16 + A * 4 + n_cells * 80 + n_cells * 12,
           "k5_register": (K * n_kf_cells + n_cells) * 80 + K * n_kf_cells * 12 + (K + 1) * 24 + 36 * 8 + 40}
    names = ["k1_kstrongest", "k3_surface_points", "k5_register"]
    per_launch_ms = [m / max(nst, 1) for m in stage_ms]
    serial_ms = [m / max(nser, 1) for m in stage_ser]
    # The roofline is quoted for the dominant kernel running ALONE (stream-ordered steps timed right after the overlapped
    # region, same inputs, CUDA events on its stream): with several steps in flight a kernel shares the SMs with the
    # neighbouring steps' kernels and its start-to-end time says how long it was resident, not how fast it runs.
    dom = int(np.argmax(serial_ms))
    ach = alg[names[dom]] * nprob / (serial_ms[dom] * 1e-3) / 1e9 if serial_ms[dom] > 0 else 0.0
    b_scan = A * R + 2 * n_pts * 16 + n_cells * 80 + (K * n_kf_cells + n_cells) * 80 + 24
    roof = {"bound": "hbm", "kernel": names[dom], "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
            "frac": ach / hbm_peak, "traffic": None, "peak_source": peak_src,
            "algorithmic_bytes_per_scan": alg[names[dom]],
            "duration_ms": serial_ms[dom],
            "duration_note": f"{names[dom]} alone: average over {nser} stream-ordered steps run right after the timed region (CUDA events on its stream); "
                             "a launch with the GPU to itself takes K5 in its 192-thread x 2-per-SM form at this batch size, the overlapped steps of the "
                             "timed region use 128 x 3 (alone: 0.38 ms; DESIGN.md section 3)",
            "stage_ms_alone": dict(zip(names, serial_ms)),
            "stage_frac_of_hbm_peak_alone": {n: (alg[n] * nprob / (t * 1e-3) / 1e9 / hbm_peak if t > 0 else None)
                                             for n, t in zip(names, serial_ms)},
            "stage_ms_in_flight": dict(zip(names, per_launch_ms)),
            "stage_ms_in_flight_note": ("CUDA events around each kernel on the stream it is launched on, averaged over the timed region; "
                                        + ("stream-ordered steps" if args.serial else
                                           "several steps are in flight, so a kernel shares the SMs with the neighbouring steps' kernels and "
                                           "the durations overlap (their sum exceeds ms_per_step)")),
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
                t = ctx.odometry_step_batch_submit(h_polar, h_mot, kf_slots, cur_slots, o)   # (host path: stream-ordered chunks, one slot set)
                if pending is not None:
                    ctx.odometry_step_batch_wait(pending)
                pending = t
            ctx.odometry_step_batch_wait(pending)
            return h_outs[(nsteps - 1) & 1]

        run_e2e(warmup)
        e2e_steps = max(args.steps, 20)                  # >= 0.1 s of timed region at ~6 ms per step
        barrier()
        t0 = time.perf_counter()
        h_out = run_e2e(e2e_steps)
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([el], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            el = float(t.item())
        # the ceiling of this arm: nothing but the images' host->device copies (same pinned buffer, same bytes per step,
        # every rank at once) -- what the PCIe link / the host's memory system give this many GPUs copying concurrently
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ctx.h2d(t_polar.data_ptr(), h_polar)
        torch.cuda.synchronize()
        el_copy = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([el_copy], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            el_copy = float(t.item())
        h2d = h_polar.nbytes + batch["mot"].nbytes + kf_slots.nbytes + cur_slots.nbytes + batch["poses"].nbytes
        d2h = h_out["poses"].nbytes + h_out["cov"].nbytes + h_out["stats"].nbytes + h_out["npts"].nbytes
        e2e = {"value": world * nprob * e2e_steps / el, "unit": "scans/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * el / e2e_steps, "steps": e2e_steps, "host_numa_binding": numa,
               "api": "cfear_odometry_step_batch_submit / _wait, two steps in flight, pinned host buffers",
               "h2d_only": {"ms_per_step": 1e3 * el_copy / e2e_steps, "gbytes_per_s_per_gpu": h_polar.nbytes * e2e_steps / el_copy / 1e9,
                            "note": "the same image bytes copied host->device and nothing else, all ranks concurrently (max over ranks)"},
               "frac_of_h2d_only": el_copy / el}
        assert np.allclose(h_out["poses"], poses_dev, atol=1e-12), "e2e and device-resident arms disagree"

    clocks = sampler.stop(tw0, time.time())      # samples taken during the device-resident and end-to-end timed regions

    # ---- CPU baseline (rank 0, N=1 only): the oracle port on the host cores, bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        nsample = min(64, nprob)
        sps, threads, el, reps, out, stage_n = cpu_leg(batch, nsample, args.cpu_seconds * 0.6)
        sps1, _, el1, reps1, _, stage_1 = cpu_leg(batch, nsample, args.cpu_seconds * 0.4, threads=1)
        d = poses_dev[:nsample, K] - out["poses"][:, K]
        par = {"max_pos_err_m": float(np.hypot(d[:, 0], d[:, 1]).max()), "max_rot_err_rad": float(np.abs(d[:, 2]).max()),
               "outer_iterations_equal": bool(np.array_equal(stats["outer_iterations"][:nsample], [s.outer_iterations for s in out["stats"]])),
               "inner_iterations_equal": bool(np.array_equal(stats["inner_iterations"][:nsample], [s.inner_iterations for s in out["stats"]])),
               "num_residuals_equal": bool(np.array_equal(stats["num_residuals"][:nsample], [s.num_residuals for s in out["stats"]])),
               "npts_equal": bool(np.array_equal(npts[:nsample], out["npts"])), "ncells_equal": bool(np.array_equal(ncells[:nsample], out["ncells"]))}
        cpu = {"value": sps, "unit": "scans/s", "cores": threads, "kind": "port",
               "sample": f"first {nsample} scans of the workload x {reps} passes ({el:.1f} s), {threads} host threads; "
                         "oracle/cfear_oracle.cc built -O3 -ffp-contract=off (the reference's flags class, CMakeLists.txt:32-33); "
                         "the ROS/PCL/Ceres reference cannot be built here",
               "single_thread": {"value": sps1, "unit": "scans/s", "cores": 1,
                                 "sample": f"same {nsample} scans x {reps1} passes ({el1:.1f} s), one scan at a time (the reference's execution model)",
                                 "stage_ms_per_scan": dict(zip(["filter", "build_normals", "register"], stage_1))},
               "stage_ms_per_scan_per_thread": dict(zip(["filter", "build_normals", "register"], stage_n)),
               "parity_vs_gpu": par}
        # north_star tolerance: 1e-4 m / 1e-5 rad after the same iteration count -- a fast wrong answer is not a result
        if not (par["max_pos_err_m"] < 1e-4 and par["max_rot_err_rad"] < 1e-5 and par["outer_iterations_equal"]
                and par["inner_iterations_equal"] and par["num_residuals_equal"] and par["npts_equal"] and par["ncells_equal"]):
            raise SystemExit("bench.py: the CUDA path disagrees with the CPU oracle on the bench workload: " + json.dumps(par))

    ab = None
    if rank == 0 and args.save_poses:
        np.savez(args.save_poses, poses=poses_dev, outer=stats["outer_iterations"], inner=stats["inner_iterations"])
    if rank == 0 and args.ref_poses and os.path.exists(args.ref_poses):
        r = np.load(args.ref_poses)
        dd = poses_dev[:, K] - r["poses"][:, K]
        ab = {"dpos": float(np.hypot(dd[:, 0], dd[:, 1]).max()), "drot": float(np.abs(dd[:, 2]).max()),
              "outer_diff": int((stats["outer_iterations"] != r["outer"]).sum()), "inner_diff": int((stats["inner_iterations"] != r["inner"]).sum())}
    if rank == 0:
        err = poses_dev[:, K] - batch["truth"]
        config["steps_in_flight"] = 1 if args.serial else NSETS
        config["api"] = "cfear_odometry_step_batch_dev" if args.serial else f"cfear_odometry_step_batch_dev_submit ({NSETS} steps in flight)"
        line = {"metric": METRIC, "value": value, "unit": "scans/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
                "timed_regions": regions, "steps_timed": total_steps, "ms_per_step": ms / total_steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config, "roofline": roof, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": int(launches), "clocks": clocks, **({"ab_vs_ref": ab} if ab else {}),
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
