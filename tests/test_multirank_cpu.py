"""world_size-2 gloo tests of the N>1 host logic (sharding, pose gather, subsequence chaining)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cfear_radarodometry_code_public_b200 import shard


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.shard_range(n_total, rank, world)
    # each rank "replays" its own scans: pose of global scan g is a fixed function of g
    local = np.stack([[0.5 * g, -0.1 * g, 0.01 * g] for g in range(lo, hi)]) if hi > lo else np.zeros((0, 3))
    tables = shard.gather_pose_tables(local)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)          # the bench's max-over-ranks timing reduction
    if rank == 0:
        q.put((tables, float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    for n in (0, 1, 7, 256, 1001):
        for w in (1, 2, 3, 8):
            r = [shard.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_gloo_pose_gather_world2():
    world, n_total = 2, 7                              # ragged: 4 + 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in ps:
        p.start()
    tables, tmax = q.get(timeout=120)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 2.0
    assert [t.shape[0] for t in tables] == [4, 3]
    allp = np.concatenate(tables, 0)
    exp = np.stack([[0.5 * g, -0.1 * g, 0.01 * g] for g in range(n_total)])
    np.testing.assert_allclose(allp, exp)


def test_chain_subsequences():
    rng = np.random.Generator(np.random.PCG64(0))
    steps = np.concatenate([rng.normal(2.5, 0.1, (20, 1)), rng.normal(0, 0.05, (20, 1)), rng.normal(0, 0.02, (20, 1))], 1)
    full = [np.zeros(3)]
    for s in steps:
        full.append(shard.se2_mul(full[-1], s))
    full = np.array(full)                              # 21 poses
    # two blocks, each restarted at identity; seam = the step between them
    def local(block):
        inv0 = block[0]
        c, s = np.cos(-inv0[2]), np.sin(-inv0[2])
        out = []
        for p in block:
            d = p[:2] - inv0[:2]
            out.append([c * d[0] - s * d[1], s * d[0] + c * d[1], p[2] - inv0[2]])
        return np.array(out)
    a, b = local(full[:11]), local(full[11:])
    got = shard.chain_subsequences([a, b], seams=[steps[10]])
    np.testing.assert_allclose(got, full, atol=1e-12)
