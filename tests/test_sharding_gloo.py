"""The N>1 host logic on CPU: world_size-2 gloo processes exercise the shard ranges, the
max-over-ranks timing reduction and the result gather used by bench.py / multi-GPU callers."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from cse305_parallel_sequence_alignment_b200 import sharding
from cse305_parallel_sequence_alignment_b200.capi import ITEM_DTYPE


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            r = [sharding.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def test_shard_by_cells_balances_ragged_batches():
    rng = np.random.default_rng(0)
    la = rng.integers(100, 5000, size=4000)
    lb = la + rng.integers(0, 300, size=4000)
    parts = sharding.shard_by_cells(la, lb, 8)
    assert parts[0][0] == 0 and parts[-1][1] == 4000
    cells = (la.astype(np.float64) * lb)
    loads = [cells[lo:hi].sum() for lo, hi in parts]
    assert max(loads) / (cells.sum() / 8) < 1.02


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_pairs = 1001
    lo, hi = sharding.shard_range(n_pairs, rank, world)
    items = np.zeros(hi - lo, dtype=ITEM_DTYPE)
    items["score"] = np.arange(lo, hi)            # stand-in for this rank's results
    items["end_i"] = rank
    counts = [b - a for a, b in (sharding.shard_range(n_pairs, r, world) for r in range(world))]
    slowest = sharding.max_over_ranks(10.0 + rank)
    total = sharding.sum_over_ranks(float(hi - lo))
    allitems = sharding.gather_items(items, counts)
    if rank == 0:
        q.put((slowest, total, allitems["score"].tolist() == list(range(n_pairs)),
               int(allitems["end_i"][-1]), len(allitems)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_reduce_and_gather():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    slowest, total, ordered, last_rank, n = q.get(timeout=120)
    [p.join(timeout=120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert slowest == 11.0 and total == 1001.0 and ordered and last_rank == 1 and n == 1001
