"""Multi-rank host logic on CPU: world size 2, gloo backend (no GPU). Covers the subject / pair-block sharding and the
variable-size all-gather that newmsm_b200.group_cost and bench.py rely on for N > 1."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        import torch
        import torch.distributed as dist
        from newmsm_b200.group_cost import Collective, shard_range
        dist.init_process_group("gloo", rank=rank, world_size=world)
        coll = Collective(dist)
        ok = coll.rank == rank and coll.world == world
        for n_total in (7, 8, 1, 2):          # uneven and even splits, and fewer items than... ranks + 0-sized shards
            b, e = shard_range(n_total, rank, world)
            # "fields" of subject s = a deterministic function of s, shape [n_local, L=2, N_t=3, D=2]
            local = torch.stack([torch.full((2, 3, 2), float(s)) + torch.arange(12.0).reshape(2, 3, 2) for s in range(b, e)]) if e > b \
                else torch.empty((0, 2, 3, 2))
            full = coll.all_gather_blocks(local.to(torch.float64), n_total)
            expect = torch.stack([torch.full((2, 3, 2), float(s)) + torch.arange(12.0).reshape(2, 3, 2) for s in range(n_total)]).to(torch.float64)
            ok = ok and full.shape == expect.shape and bool(torch.equal(full, expect))
        # pair-cost blocks: every rank evaluates its block of a fake cost function, gathered = the serial result
        P = 1001
        b, e = shard_range(P, rank, world)
        blk = torch.tensor([[p * 4 + c for c in range(4)] for p in range(b, e)], dtype=torch.float64)
        allc = coll.all_gather_blocks(blk, P)
        ok = ok and bool(torch.equal(allc, torch.arange(4.0 * P, dtype=torch.float64).reshape(P, 4)))
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, ok, ""))
    except Exception as ex:   # pragma: no cover
        q.put((rank, False, repr(ex)))


def test_shard_ranges_cover_everything():
    sys.path.insert(0, ROOT)
    from newmsm_b200.group_cost import shard_counts, shard_range
    for n in (0, 1, 5, 64, 1001):
        for world in (1, 2, 3, 8):
            rs = [shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            assert sum(shard_counts(n, world)) == n and max(shard_counts(n, world)) - min(shard_counts(n, world)) <= 1


def test_all_gather_blocks_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
