"""The N > 1 path on CPU: world_size-2 gloo processes exercise the batch partition, the
max-over-ranks timing reduction and the final gather (the only collectives the workload has)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bionc_b200.utils.sharding import gather_states, max_over_ranks, shard_range


def test_shard_range_partitions_the_batch():
    for total in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(total, rank, world)
        # stand-in for "integrate my shard": every system is tagged with its global index
        local = torch.arange(lo, hi, dtype=torch.float64)[:, None] * torch.tensor([[1.0, 10.0, 100.0]], dtype=torch.float64)
        full = gather_states(local, total)
        times = max_over_ranks([1.0 + rank, 5.0 - rank])
        q.put((rank, full.shape, bool(torch.equal(full[:, 0], torch.arange(total, dtype=torch.float64))), times))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gather_and_timing_reduction():
    world, total = 2, 11  # ragged: 6 + 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, shape, ordered, times in results:
        assert tuple(shape) == (total, 3) and ordered
        assert times == [2.0, 5.0]
