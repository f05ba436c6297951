#!/usr/bin/env python
"""replay.py -- BASELINE configs[4]: batched odometry replay of synthetic Oxford-shaped (sub)sequences.

Each GPU (rank) replays `--nseq` independent sequences in lock-step for `--steps` scans through cfear_seq_* (the whole
OdometryKeyframeFuser loop on the device: compensate -> k-strongest -> surface points -> registration against the
sliding keyframe window -> keyframe bookkeeping), then the per-rank pose tables are gathered with one NCCL all_gather.
Prints one JSON line (rank 0).  Secondary measurement; the headline metric lives in bench.py.

  python replay.py --nseq 32 --steps 40                      # 1 GPU
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 replay.py --nseq 32 --steps 40
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def _seq(args):
    from cfear_radarodometry_code_public_b200 import synth
    seed, n = args
    return synth.make_sequence(seed, n)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nseq", type=int, default=32)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--submap", type=int, default=4)
    ap.add_argument("--check", type=int, default=1, help="sequences per rank verified against the CPU oracle replay")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    with mp.get_context("fork").Pool(min(args.nseq, os.cpu_count() or 1)) as pool:
        res = pool.map(_seq, [(1000 * rank + b, args.steps) for b in range(args.nseq)])
    imgs = np.stack([r[0] for r in res])                     # [nseq, steps, A, R]
    truth = np.stack([r[1] for r in res])

    import torch
    import torch.distributed as dist
    from cfear_radarodometry_code_public_b200 import capi, shard, workload
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        sys.stdout.flush()                         # NCCL's version banner goes to stdout during init: keep stdout to the JSON line
        saved_stdout = os.dup(1); os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier(); torch.cuda.synchronize()
        finally:
            sys.stdout.flush(); os.dup2(saved_stdout, 1); os.close(saved_stdout)
    K = args.submap
    ctx = capi.Context(device=local, max_batch=args.nseq, max_cellsets=args.nseq * (K + 1), max_keyframes=K, **workload.CFEAR3)
    ext = torch.cuda.ExternalStream(ctx.stream_ptr, device=dev)
    by_step = np.ascontiguousarray(imgs.transpose(1, 0, 2, 3))          # [steps, nseq, A, R]
    t_imgs = torch.from_numpy(by_step).to(dev)
    out = {}
    for mode in ("resident", "host"):
        S = capi.Sequences(ctx, args.nseq, args.steps, submap_scan_size=K)
        h = capi.pinned_array(by_step.shape, np.uint8); h[...] = by_step
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for t in range(args.steps):
            if mode == "resident":
                S.step_dev(t_imgs[t].data_ptr())
            else:
                S.step(h[t])
        ctx.join()                                     # the context stream waits for the steps on the library's streams
        e1.record(ext)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        out[mode] = world * args.nseq * args.steps / (ms * 1e-3)
        poses, kf, st = S.read(0, args.steps)
        S.close()
    tables = shard.gather_pose_tables(poses.reshape(-1, 3), device=dev) if world > 1 else [poses.reshape(-1, 3)]
    parity = None
    if args.check:
        import oracle as orc
        orc.build()
        mx = 0.0
        for b in range(min(args.check, args.nseq)):
            ref = orc.odometry_sequence(imgs[b], orc.reg_cfg(cost="P2D", weight_opt=4, regularization=0.1), radius=3.0,
                                        weight_intensity=True, submap_scan_size=K)
            d = poses[b] - ref["poses"]
            mx = max(mx, float(np.hypot(d[:, 0], d[:, 1]).max()))
            assert np.array_equal(kf[b], ref["keyframe"])
        parity = mx
    if rank == 0:
        end_err = np.hypot(*(poses[:, -1, :2] - truth[:, -1, :2]).T)
        print(json.dumps({"workload": "configs[4]: lock-step replay of independent synthetic sequences, CFEAR-3-like parameters (P2D), k=12, r=3.0, window %d" % K,
                          "n_gpus": world, "sequences_per_gpu": args.nseq, "scans_per_sequence": args.steps,
                          "scans_per_s_device_resident": out["resident"], "scans_per_s_host_images": out["host"],
                          "gathered_pose_rows": int(sum(t.shape[0] for t in tables)),
                          "max_pos_err_vs_oracle_replay_m": parity, "keyframes_per_sequence_mean": float(kf.sum(1).mean()),
                          "median_end_point_err_vs_sim_truth_m": float(np.median(end_err))}))
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
