"""CPU: host logic of the data-parallel path (trial sharding, gradient-bucket plan, mean all-reduce) with the gloo
backend at world size 2."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multi_modal_foundation_model_b200.parallel import all_reduce_mean, plan_buckets, shard_range


def test_shard_range_partitions_every_trial_once():
    for n in (1, 7, 16, 256, 257):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))
            sizes = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_plan_buckets_cover_buffer_in_order():
    marks = [(10, 100), (20, 150), (30, 400), (40, 420), (50, 1000)]
    b = plan_buckets(marks, 1000, target_elems=200)
    assert b[0][1] == 0 and b[-1][2] == 1000 and b[-1][0] == 50
    for (c0, lo0, hi0), (c1, lo1, hi1) in zip(b, b[1:]):
        assert hi0 == lo1 and c0 <= c1
    assert all(hi - lo >= 200 for _, lo, hi in b[:-1])
    # a mark's range is only reduced once the schedule has passed it
    for ci, lo, hi in b:
        assert any(mc <= ci and off >= hi for mc, off in marks)
    assert plan_buckets([(5, 50)], 50, 10) == [(5, 0, 50)]


def _worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # each rank holds the gradient of its own shard; DDP semantics = arithmetic mean of the rank gradients
        g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
        for ci, lo, hi in plan_buckets([(1, 300), (2, 700), (3, 1000)], 1000, 250):
            all_reduce_mean(g[lo:hi])
        assert torch.allclose(g, torch.arange(1000, dtype=torch.float32) * 1.5)
        lo, hi = shard_range(10, rank, world)
        cnt = torch.tensor([hi - lo], dtype=torch.float32)
        dist.all_reduce(cnt)
        assert cnt.item() == 10
    finally:
        dist.destroy_process_group()


def test_gloo_world2_mean_allreduce_over_buckets():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port), nprocs=2, join=True)
