"""world_size-2 gloo test of the only multi-rank logic on the path: env sharding (disjoint,
seed-offset shards) and the episode-statistics all-reduce."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dexterity_b200 import sharding
    lo, hi = sharding.shard_range(131072, rank, world)
    stats = torch.tensor([10.0 * (rank + 1), 100.0, 4.0 + rank, 1.0, 2.0, 0.0, 0.0, 0.0], dtype=torch.float64)
    red = sharding.reduce_episode_stats(stats)
    out[rank] = (lo, hi, sharding.rank_seed(12345, rank), red)
    dist.destroy_process_group()


def test_sharding_and_stats_reduce():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, 29517, out), nprocs=world, join=True)
        r0, r1 = out[0], out[1]
    assert (r0[0], r0[1], r1[0], r1[1]) == (0, 65536, 65536, 131072)
    assert r0[2] != r1[2]
    for r in (r0, r1):
        assert r[3]["episodes"] == 9.0 and abs(r[3]["mean_return"] - 30.0 / 9.0) < 1e-12
        assert r[3]["successes"] == 2.0 and r[3]["falls"] == 4.0


def test_single_process_passthrough():
    sys.path.insert(0, ROOT)
    from dexterity_b200 import sharding
    s = sharding.reduce_episode_stats(torch.tensor([5.0, 10.0, 2.0, 1, 0, 1, 0, 0], dtype=torch.float64))
    assert s["mean_length"] == 5.0 and sharding.shard_range(10, 0, 1) == (0, 10)
