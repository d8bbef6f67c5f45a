"""Multi-GPU sharding logic, exercised on CPU with the gloo backend (world_size 2 and 3)."""
import os
import socket

import numpy as np
import pytest


def test_shard_bounds_cover_and_match_reference_split():
    from hpgeom_b200.sharding import shard_bounds
    for n in (0, 1, 7, 8, 1000, 8_000_000_001):
        for w in (1, 2, 3, 4, 8):
            bounds = [shard_bounds(n, w, r) for r in range(w)]
            assert bounds[0][0] == 0 and bounds[-1][1] == n
            for (a0, a1), (b0, b1) in zip(bounds, bounds[1:]):
                assert a1 == b0 and a1 >= a0
            sizes = [b - a for a, b in bounds]
            assert max(sizes) - min(sizes) <= 1
            # reference rule (hpgeom/hpgeom.c:272-279): the first n % w workers get the extra element
            assert sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _worker(rank, world, port, n, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hpgeom_b200.sharding import gather, shard_bounds
    full = torch.arange(n, dtype=torch.int64) * 3 + 1
    lo, hi = shard_bounds(n, world, rank)
    local = full[lo:hi] * 2  # stands in for "this rank's kernel output", resident per rank
    got = gather(local)
    rows = gather(torch.stack([local, local + 1], dim=1))  # (n_local, 2) row outputs shard on dim 0
    ok = bool(torch.equal(got, full * 2)) and bool(torch.equal(rows[:, 1], full * 2 + 1))
    q.put((rank, ok, hi - lo))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 1001), (3, 10), (2, 1), (2, 1000), (3, 0)])
def test_gather_over_gloo(world, n):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res)
    assert sum(sz for _, _, sz in res) == n
