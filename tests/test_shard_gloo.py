"""The N > 1 host logic on CPU: two gloo ranks shard candidates, reduce timings and gather the best."""
import os
import socket

import numpy as np


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from topay_b200 import shard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.shard_range(9, rank, world)
    rng = np.random.default_rng(5)
    status = (rng.random(9) < 0.7).astype(np.int32)
    duration, cost = rng.uniform(5, 9, 9), rng.uniform(100, 900, 9)
    best = shard.gather_best(dist, status[lo:hi], duration[lo:hi], cost[lo:hi], index_offset=lo)
    tmax = shard.reduce_max(dist, 1.0 + rank)
    none = shard.gather_best(dist, np.zeros(3, np.int32), np.ones(3), np.ones(3), index_offset=3 * rank)
    q.put((rank, (lo, hi), best, tmax, none))
    dist.destroy_process_group()


def test_two_rank_shard_and_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
    assert out[0][1] == (0, 5) and out[1][1] == (5, 9)
    rng = np.random.default_rng(5)
    status = (rng.random(9) < 0.7).astype(np.int32)
    duration, cost = rng.uniform(5, 9, 9), rng.uniform(100, 900, 9)
    ok = np.nonzero(status == 1)[0]
    want = (int(ok[np.argmin(duration[ok])]), int(ok[np.argmin(cost[ok])]))
    assert out[0][2] == want and out[1][2] == want          # every rank learns the same winner
    assert out[0][3] == 2.0 and out[1][3] == 2.0            # max over ranks
    assert out[0][4] == (-1, -1)


def test_shard_helpers():
    from topay_b200 import shard
    assert [shard.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert shard.round_robin(7, 1, 3) == [1, 4]
    assert shard.gather_best(None, [0, 1, 1], [3.0, 2.0, 2.0], [5.0, 9.0, 1.0]) == (1, 2)
    assert shard.reduce_max(None, 3.5) == 3.5
