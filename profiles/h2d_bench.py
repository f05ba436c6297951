"""Concurrent host->device copy ceiling of the box: every rank copies a pinned buffer of one bench step's image bytes
(256 x 400 x 3360 = 344 MB) to its GPU, all ranks at once, for a fixed number of repetitions.

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/h2d_bench.py [--streams S] [--numa]

Prints (rank 0) GB/s per GPU (min / mean over ranks) and the aggregate.  --streams S splits every copy over S CUDA streams;
--numa pins the process to the CPUs of the GPU's NUMA node before the buffer is allocated (first touch)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cfear_radarodometry_code_public_b200 import capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=1)
ap.add_argument("--numa", action="store_true")
ap.add_argument("--reps", type=int, default=40)
ap.add_argument("--mb", type=float, default=344.064)
a = ap.parse_args()
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
numa = capi.bind_to_device_numa(local) if a.numa else None
n = int(a.mb * 1e6)
h = torch.from_numpy(capi.pinned_array((n,), np.uint8)); h.fill_(7)
d = torch.empty(n, dtype=torch.uint8, device=dev)
streams = [torch.cuda.Stream(dev) for _ in range(a.streams)]
chunks = [(i * n // a.streams, (i + 1) * n // a.streams) for i in range(a.streams)]


def run(reps):
    for _ in range(reps):
        for s, (lo, hi) in zip(streams, chunks):
            with torch.cuda.stream(s):
                d[lo:hi].copy_(h[lo:hi], non_blocking=True)
    torch.cuda.synchronize()


run(3)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
run(a.reps)
el = time.perf_counter() - t0
gbs = torch.tensor([n * a.reps / el / 1e9], dtype=torch.float64, device=dev)
if world > 1:
    allg = [torch.zeros_like(gbs) for _ in range(world)]
    dist.all_gather(allg, gbs)
    v = [float(x.item()) for x in allg]
else:
    v = [float(gbs.item())]
if rank == 0:
    print(json.dumps({"n_gpus": world, "streams": a.streams, "numa": numa, "mb_per_copy": a.mb, "gbs_per_gpu_min": min(v), "gbs_per_gpu_mean": float(np.mean(v)),
                      "gbs_aggregate": float(np.sum(v)), "is_pinned": bool(h.is_pinned())}))
if world > 1:
    dist.destroy_process_group()
