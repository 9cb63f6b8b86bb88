"""world_size-2 gloo tests (CPU) of the N>1 host logic: contiguous batch sharding with no data-path
collective, then the final result gather — including a ragged batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deeppowers_b200.sharding import gather_results, shard_range


def test_shard_range_partitions_exactly():
    for batch in (0, 1, 7, 8, 4096, 65536 + 3):
        for world in (1, 2, 3, 8):
            spans = [shard_range(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, batch, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import Oracle
        o = Oracle(10, 2)
        s = o.keygen_secret(1)
        evk = o.keygen_relin(2, 65537, s)
        lo, hi = shard_range(batch, rank, world)
        # every rank regenerates only ITS slice of the global synthetic stream (counter-based generator)
        a = o.fill_uniform(7, 2 * (hi - lo), first_poly=2 * lo).reshape(hi - lo, 2, 2, o.N)
        b = o.fill_uniform(8, 2 * (hi - lo), first_poly=2 * lo).reshape(hi - lo, 2, 2, o.N)
        local = o.ct_mul_relin(a, b, evk) if hi > lo else np.zeros((0, 2, 2, o.N), dtype=np.uint64)
        full = gather_results(torch.from_numpy(local.view(np.int64)), batch, dst=0)
        if rank == 0:
            fa = o.fill_uniform(7, 2 * batch).reshape(batch, 2, 2, o.N)
            fb = o.fill_uniform(8, 2 * batch).reshape(batch, 2, 2, o.N)
            ref = o.ct_mul_relin(fa, fb, evk)
            q.put(bool(np.array_equal(full.numpy().view(np.uint64), ref)))
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [4, 5])
def test_two_rank_shard_then_gather_equals_single_rank(batch, oracle_mod):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
