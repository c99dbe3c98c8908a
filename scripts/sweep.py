"""BASELINE configs[4]: a sweep of independent table / cuboid scenarios, sharded round-robin over the ranks
(one process per GPU, no data-path collective). Every scenario runs the whole device pipeline the planner
would drive: rasterise the scene's point cloud -> rebuild the field (4 x 2-D + 3-D ESDF) -> solve its
candidates in one batch -> success gate (checkFeasible / printConstraintsSituations) -> shortest successful
trajectory. Candidates are synthetic S-curves at the reference's scale (the randomised CPU front-ends are out
of scope). Prints one JSON line on rank 0.

  python scripts/sweep.py --scenarios 64
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/sweep.py --scenarios 512
"""
import argparse
import json
import os
import sys
import threading
import time

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenarios", type=int, default=64)
    ap.add_argument("--candidates", type=int, default=8, help="candidates per scenario (reference: <= 8)")
    ap.add_argument("--in-flight", type=int, default=8, help="scenarios in flight per GPU (own field + solver each)")
    args = ap.parse_args()
    import torch
    import topay_b200 as tp
    from topay_b200 import scenes, shard
    rank, world, local = shard.env_rank()
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    mine = shard.round_robin(args.scenarios, rank, world)
    P = max(1, min(args.in_flight, len(mine)))
    opt, rp = tp.opt_params_default(), tp.robot_params_default()
    slots = []
    for _ in range(P):
        gm = tp.GridMap(tp.grid_desc(), device=local)
        gm.set_keep_sqdist(False)
        slots.append((gm, tp.MomaTrajOpt(gm, max_cand=args.candidates, max_pieces=16, opt_param=opt, robot=rp)))
    # scene generation is host work of the simulator, not of the path: done up front
    clouds = {s: (scenes.tables_scene(s)[0] if s % 2 == 0 else scenes.cuboids_scene(s)[0]) for s in mine}
    cands = {s: scenes.short_candidates(args.candidates, 100000 + s) for s in mine}
    out, lock = {}, threading.Lock()

    def worker(slot):
        gm, solver = slots[slot]
        for s in mine[slot::P]:
            t0 = time.perf_counter()
            gm.regenerateMap(clouds[s])                       # clear + rasterise + updateESDF
            paths, bv, ba = cands[s]
            res = solver.optimizeTrajBatch(paths, bv, ba)
            arrs, winner = solver.checkFeasibleBatch()
            with lock:
                out[s] = (winner, int(res["status"].sum()), int(arrs["feasible_print"].sum()),
                          float(res["duration"][winner]) if winner >= 0 else float("nan"),
                          (time.perf_counter() - t0) * 1e3)

    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    th = [threading.Thread(target=worker, args=(i,)) for i in range(P)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    torch.cuda.synchronize()
    dt = shard.reduce_max(dist, time.perf_counter() - t0, device=f"cuda:{local}")
    n_win = sum(1 for v in out.values() if v[0] >= 0)
    tot = torch.tensor([float(len(out)), float(n_win)], dtype=torch.float64, device=f"cuda:{local}")
    if dist is not None:
        dist.all_reduce(tot)
    if rank == 0:
        lat = np.array([v[4] for v in out.values()])
        print(json.dumps({"workload": "scenario sweep (BASELINE configs[4]): rasterise + field rebuild + batched solve + "
                                      "success gate + selection per scenario, round-robin over ranks",
                          "scenarios": int(tot[0].item()), "candidates_per_scenario": args.candidates, "n_gpus": world,
                          "in_flight_per_gpu": P, "scenarios_per_s": tot[0].item() / dt,
                          "trajectories_per_s": tot[0].item() * args.candidates / dt,
                          "scenarios_with_a_feasible_winner": int(tot[1].item()),
                          "rank0_scenario_latency_ms_p50": float(np.median(lat)), "seconds": dt}), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
