"""N > 1 plumbing on CPU: two gloo ranks own disjoint shards and combine their timings the way bench.py does."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from metasnv_b200.sharding import device_for_split, lpt_bins, reduce_over_ranks, splits_of_rank


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = splits_of_rank(5, rank, world)
    times, work = reduce_over_ranks(dist, "cpu", [10.0 + rank, 3.0 - rank], [len(mine), 100 * (rank + 1)])
    owners = [None] * world
    dist.all_gather_object(owners, mine)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, times, work, owners))


def test_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, times, work, owners in res:
        assert times == [11.0, 3.0]                      # max over ranks
        assert work == [5.0, 300.0]                      # sum over ranks: every shard owned exactly once
        assert sorted(owners[0] + owners[1]) == [0, 1, 2, 3, 4] and not set(owners[0]) & set(owners[1])


def test_split_to_device_rule():
    assert device_for_split("out/snpCaller/indiv_called.best_split_11", 8) == 3
    assert device_for_split("indiv_called", 8) == 0
    assert [device_for_split("x.best_split_%d" % k, 4) for k in range(6)] == [0, 1, 2, 3, 0, 1]


def test_lpt_bins_matches_createoptimumsplit_rule():
    # heaviest first into the lightest bin (createOptimumSplit.py:56-60)
    assert lpt_bins([5, 4, 3, 3, 3], 2) in ([0, 1, 1, 0, 1], [0, 1, 1, 0, 0])
    b = lpt_bins([1.0] * 8, 8)
    assert sorted(b) == list(range(8))
