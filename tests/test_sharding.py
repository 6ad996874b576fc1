"""Host logic of the sharded certification loop on the CPU: draw partitioning and the vote-count
all-reduce over a world_size-2 gloo group."""

import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from audiopure_b200.certified_robust import shard_range, torch_counts_allreduce


def test_shard_range_partitions_every_draw_once():
    for n in (0, 1, 7, 100, 10000, 10001):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b and c <= d
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    assert shard_range(10000, 0, 8) == (0, 1250)  # 1250 draws per GPU at 8 GPUs


def _worker(rank, world, port, n, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # a deterministic per-draw "prediction": every draw index votes for class (7*i) % 10
    lo, hi = shard_range(n, rank, world)
    counts = torch.zeros(10, dtype=torch.int64)
    for i in range(lo, hi):
        counts[(7 * i) % 10] += 1
    torch_counts_allreduce(counts)
    ret[rank] = counts.tolist()
    dist.destroy_process_group()


def test_counts_allreduce_gloo_world2():
    n, world = 1001, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29653, n, ret), nprocs=world, join=True)
    want = torch.zeros(10, dtype=torch.int64)
    for i in range(n):
        want[(7 * i) % 10] += 1
    assert ret[0] == ret[1] == want.tolist()  # integer sums: order-independent, bit-exact
