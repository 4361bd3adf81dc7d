"""Host-side logic of the N>1 path on CPU: world_size-2 gloo group, sharding of independent image
pairs and max-over-ranks timing (the op itself has no collective: it is per-sample)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cerberusnet_b200.parallel import aggregate_throughput, max_over_ranks, shard_range, sum_over_ranks


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 8, 33):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, e = shard_range(9, world, rank)
        ms = 10.0 + 5.0 * rank  # rank 1 is slower
        res = (sum_over_ranks(e - b), max_over_ranks(ms), aggregate_throughput(e - b, ms))
        dist.barrier()
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_aggregation():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in (0, 1):
        total, ms, thr = out[r]
        assert total == 9 and ms == 15.0
        assert abs(thr - 9 / 0.015) < 1e-6


def test_single_process_is_identity():
    assert max_over_ranks(3.5) == 3.5 and sum_over_ranks(2.0) == 2.0
    assert aggregate_throughput(10, 100.0) == 100.0
