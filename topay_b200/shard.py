"""Multi-GPU sharding of the path: one process per GPU, independent candidates / scenarios per
rank, no data-path collective. The only exchanges are (a) the timing reduction of
the benchmark and (b) the tiny best-cost gather of a plan whose candidates were split over ranks
(24 B per rank; the reference's selection rule, planner.cpp:999-1010). Both work on any
torch.distributed backend (nccl on the GPU box, gloo in the CPU tests)."""
import os


def env_rank():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_range(n_items, rank, world):
    """Contiguous balanced shard [lo, hi) of n_items for `rank` (first n_items % world ranks get one more)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def round_robin(n_items, rank, world):
    """Static round-robin sharding of scenarios (BASELINE configs[4])."""
    return list(range(rank, n_items, world))


def reduce_max(dist, value, device="cpu"):
    """max over ranks of a python float; identity when dist is None (single process)."""
    if dist is None:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_best(dist, status, duration, cost, index_offset=0, device="cpu"):
    """Best successful candidate across ranks. Each rank passes its local arrays; returns
    (best_by_duration, best_by_cost) as global indices (-1 when no rank succeeded). Ties keep the
    lowest global index, which is the reference's first-success-wins scan order."""
    import numpy as np
    status, duration, cost = (np.asarray(a) for a in (status, duration, cost))
    ok = np.nonzero(status == 1)[0]
    loc = [-1.0, float("inf"), -1.0, float("inf")]
    if len(ok):
        bd = ok[np.argmin(duration[ok])]
        bc = ok[np.argmin(cost[ok])]
        loc = [float(bd + index_offset), float(duration[bd]), float(bc + index_offset), float(cost[bc])]
    rows = [loc]
    if dist is not None:
        import torch
        t = torch.tensor(loc, dtype=torch.float64, device=device)
        out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(out, t)
        rows = [o.tolist() for o in out]
    best_d = min((r for r in rows if r[0] >= 0), key=lambda r: (r[1], r[0]), default=None)
    best_c = min((r for r in rows if r[2] >= 0), key=lambda r: (r[3], r[2]), default=None)
    return (int(best_d[0]) if best_d else -1, int(best_c[2]) if best_c else -1)
