"""Multi-GPU plumbing: the path shards over independent scans / (sub)sequences with no data-path collective
(SURVEY 8e -- the reference's own parallel model is independent offline_odometry processes,
launch/oxford/eval/utils/worker:86-87).  The only exchange is the final gather of the pose tables.

torch.distributed is used for the rendezvous and the gather (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of [0, n_total) owned by `rank`; block sizes differ by at most one."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_pose_tables(local, group=None, device=None):
    """all_gather of per-rank pose tables [n_local, 3] (x, y, yaw) -> list of numpy tables, one per rank.
    Ragged sizes are handled by padding to the max length.  `local` may be a numpy array or a torch tensor
    (a CUDA tensor keeps the gather on NCCL)."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(local, dtype=torch.float64)
    if device is not None:
        t = t.to(device)
    world = dist.get_world_size(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    nmax = int(max(s.item() for s in sizes))
    pad = torch.zeros((nmax, 3), dtype=torch.float64, device=t.device)
    pad[: t.shape[0]] = t.reshape(-1, 3)
    out = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return [o[: int(s.item())].cpu().numpy() for o, s in zip(out, sizes)]


def se2_mul(a, b):
    ca, sa = np.cos(a[2]), np.sin(a[2])
    return np.array([a[0] + ca * b[0] - sa * b[1], a[1] + sa * b[0] + ca * b[1], a[2] + b[2]])


def chain_subsequences(tables, seams=None):
    """Compose per-rank subsequence trajectories (each starting at identity) into one trajectory: a host-side
    prefix product over <= world SE(2) elements.  seams[i] (optional) is the relative pose between the last pose of
    block i and the first pose of block i+1 (identity if None)."""
    out, base = [], np.zeros(3)
    for i, tab in enumerate(tables):
        tab = np.asarray(tab, dtype=np.float64).reshape(-1, 3)
        blk = np.array([se2_mul(base, p) for p in tab]) if len(tab) else np.zeros((0, 3))
        out.append(blk)
        if len(blk):
            base = blk[-1]
            if seams is not None and i < len(seams):
                base = se2_mul(base, np.asarray(seams[i], dtype=np.float64))
    return np.concatenate(out, 0) if out else np.zeros((0, 3))
