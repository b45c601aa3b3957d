"""world_size-2 test of the multi-GPU plumbing on CPU with the gloo backend: env sharding + the statistics
all-reduce (the path's only collective)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fleetrl_b200.dist import all_reduce_stats, shard_range, world as w
    assert w() == (rank, world, rank)
    lo, hi = shard_range(total, rank, world)
    # each rank's "partial statistics": [number of envs, sum of global env ids, ...]
    ids = torch.arange(lo, hi, dtype=torch.float64)
    stats = torch.zeros(10, dtype=torch.float64)
    stats[0] = hi - lo; stats[1] = ids.sum(); stats[2] = float(rank + 1)
    all_reduce_stats(stats)
    q.put((rank, lo, hi, stats.tolist()))
    dist.destroy_process_group()


def test_shard_and_allreduce_two_ranks():
    world, total = 2, 1001
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, s0), (r1, lo1, hi1, s1) = out
    assert (lo0, hi0, lo1, hi1) == (0, 501, 501, 1001)
    assert s0 == s1                                   # every rank holds the reduced vector
    assert s0[0] == total and s0[1] == total * (total - 1) / 2 and s0[2] == 3.0
