"""N > 1 host logic on CPU: world_size-2 gloo process group (no GPU needed)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from multimodal_registration_b200 import sharding  # noqa: E402


def test_shard_items_partition():
    for n in (0, 1, 7, 64):
        for world in (1, 2, 3, 8):
            parts = [sharding.shard_items(n, r, world) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(n))               # every item exactly once
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert sharding.shard_items(64, 3, 8) == list(range(3, 64, 8))          # config 4: 8 subjects per GPU
    with pytest.raises(ValueError):
        sharding.shard_items(4, 2, 2)
    assert sharding.per_device_batch(8, 4) == 2
    with pytest.raises(ValueError):
        sharding.per_device_batch(3, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        items = sharding.shard_items(7, rank, world)
        # each rank "processes" its items; scalars (count, sum of ids, a timing) are gathered
        g = sharding.gather_scalars([len(items), float(sum(items)), 10.0 + rank])
        grad = torch.full((5,), float(rank + 1))
        sharding.allreduce_mean_(grad)
        t = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)                            # max-over-ranks timing rule
        q.put((rank, g.tolist(), grad.tolist(), t.item()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_gather_and_allreduce():
    world, port = 2, 29611
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, g, grad, tmax in res:
        assert g == [[4.0, 12.0, 10.0], [3.0, 9.0, 11.0]]                   # items 0,2,4,6 | 1,3,5
        assert grad == [1.5] * 5                                            # mean of 1 and 2
        assert tmax == 11.0
