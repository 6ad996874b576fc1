"""Host logic of the sharded certification loop on the CPU: draw partitioning and the vote-count
all-reduce over a world_size-2 gloo group."""

import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from audiopure_b200.certified_robust import flat_to_clip_draw, shard_range, torch_counts_allreduce, work_batches


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


def test_work_batches_cover_every_clip_draw_pair_once_with_full_batches():
    """The certify work list (clips x (n_0 + n) draws, clip-major) over 1..8 ranks: every (clip, draw) pair lands in
    exactly one batch of one rank, every batch but the globally last one is full, the batches are THE SAME at every
    world size (an item keeps its batch and its row), and the n_0 / n split is by draw index."""
    for clips, n_0, n, bs in ((1, 100, 10000, 64), (4, 100, 10000, 64), (3, 5, 21, 8), (2, 32, 128, 64), (5, 1, 1, 7)):
        per_clip = n_0 + n
        single = list(work_batches(clips, per_clip, 0, 1, bs))
        assert all(b == bs for _, b in single[:-1]) and 0 < single[-1][1] <= bs
        for world in (1, 2, 3, 8):
            seen = set()
            merged = []
            for r in range(world):
                batches = list(work_batches(clips, per_clip, r, world, bs))
                merged += batches
                for s, b in batches:
                    for flat in range(s, s + b):
                        item = flat_to_clip_draw(flat, per_clip)
                        assert item not in seen
                        seen.add(item)
            assert merged == single                       # same batches, same order, whatever the world size
            counts = [len(list(work_batches(clips, per_clip, r, world, bs))) for r in range(world)]
            assert max(counts) - min(counts) <= 1         # ranks differ by at most one batch
            assert seen == {(c, d) for c in range(clips) for d in range(per_clip)}
            assert sum(1 for c, d in seen if d < n_0) == clips * n_0
    # 8 GPUs, bench.py's certify leg (4 clips x 10100 draws = 632 batches of 64): 79 batches per rank, all full but the
    # last one of the last rank -- instead of the 12-13-draw slivers a per-clip, per-pass split of n_0 = 100 gives
    assert [len(list(work_batches(4, 10100, r, 8, 64))) for r in range(8)] == [79] * 8


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
