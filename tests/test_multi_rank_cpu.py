"""Host-side logic of the N > 1 path on CPU: world_size-2 gloo group, conformer sharding and max-over-ranks timing."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import bench
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = bench.shard_conformers(64, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    slow = bench.max_over_ranks(10.0 + 5.0 * rank, dist, torch.device("cpu"))
    dist.barrier()
    if rank == 0:
        out.put((gathered, slow))
    dist.destroy_process_group()


def test_sharding_and_timing_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, slow = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert sorted(gathered[0] + gathered[1]) == list(range(64))          # every conformer exactly once
    assert len(gathered[0]) == len(gathered[1]) == 32 and not set(gathered[0]) & set(gathered[1])
    assert slow == 15.0                                                   # the slowest rank defines the step time


def test_single_rank_sharding_is_identity():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.shard_conformers(5, 0, 1) == [0, 1, 2, 3, 4]
    assert bench.max_over_ranks(3.5, None, torch.device("cpu")) == 3.5
