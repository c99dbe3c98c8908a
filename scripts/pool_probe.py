"""Dev probe: continuous batching. (1) identical results with few and many slots, (2) throughput of P plans of 256
candidates through a slot pool."""
import argparse, sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import topay_b200 as tp
from topay_b200 import scenes

ap = argparse.ArgumentParser()
ap.add_argument("--plans", type=int, default=8)
ap.add_argument("--slots", type=int, nargs="+", default=[256, 512, 1024])
ap.add_argument("--cand", type=int, default=256)
ap.add_argument("--skip-small", action="store_true")
ap.add_argument("--no-timed", action="store_true")
args = ap.parse_args()

pts, _ = scenes.cuboids_scene(42)
gm = tp.GridMap(tp.grid_desc(), device=0)
gm.regenerateMap(pts)

if not args.skip_small:
    opt, rp = tp.opt_params_default(), tp.robot_params_default()
    paths, bv, ba = scenes.short_candidates(32, 11)
    res = {}
    for ns in (32, 8, 3):
        s = tp.MomaTrajOpt(gm, max_cand=32, max_pieces=16, opt_param=opt, robot=rp, n_slots=ns)
        t0 = time.perf_counter()
        r = s.optimizeTrajBatch(paths, bv, ba)
        dt = time.perf_counter() - t0
        st = s.stats()
        res[ns] = r
        print(f"small: slots {ns:3d}: {dt*1e3:7.1f} ms ticks {st['ticks']} slot_ticks {st['slot_ticks']} evals {int(r['evals'].sum())} "
              f"status {int(r['status'].sum())}/32 util {r['evals'].sum()/max(st['slot_ticks'],1):.2f}")
        s.close()
    for ns in (8, 3):
        same = all(np.array_equal(res[32][k], res[ns][k]) for k in ("status", "cost", "T", "coeff", "evals", "x"))
        print(f"small: slots {ns} identical to slots 32: {same}")

opt, rp = tp.opt_params_default(), tp.robot_params_default()
opt.int_K, opt.min_piece_num, opt.sample_interval = 32, 64, 1e9
P, C = args.plans, args.cand
batches = [scenes.synthetic_batch(C, 1234 + 1000 * p) for p in range(P)]
paths = [q for b in batches for q in b[0]]
bv = np.concatenate([b[1] for b in batches]); ba = np.concatenate([b[2] for b in batches])
for ns in args.slots:
    s = tp.MomaTrajOpt(gm, max_cand=P * C, max_pieces=64, opt_param=opt, robot=rp, n_slots=ns)
    s.upload(paths, bv, ba)
    s.run()     # warm-up (graphs)
    t0 = time.perf_counter()
    s.run()
    dt = time.perf_counter() - t0
    st = s.stats()
    r = s.download()
    print(f"pool: {P} plans x {C}, slots {ns:5d}: {dt:6.3f} s  {P*C/dt:7.1f} traj/s  ticks {st['ticks']} ms/tick {st['ms_total']/st['ticks']:.3f} "
          f"util {r['evals'].sum()/max(st['slot_ticks'],1):.2f} evals/cand {r['evals'].mean():.0f} ok {int(r['status'].sum())}", flush=True)
    s.close()

if args.no_timed:
    sys.exit(0)
# per-kernel device time, plain launches with events (timed mode), the last slot count
ns = args.slots[-1]
s = tp.MomaTrajOpt(gm, max_cand=P * C, max_pieces=64, opt_param=opt, robot=rp, n_slots=ns)
s.upload(paths, bv, ba)
s.set_timed(True)
s.run()
st = s.stats()
tk = st["ticks"]
print(f"timed: slots {ns}: ms/tick total {st['ms_total']/tk:.3f} | integrate {st['ms_integrate']/tk:.3f} penalty {st['ms_eval']/tk:.3f} "
      f"chain {st['ms_chain']/tk:.3f} adj {st['ms_adj']/tk:.3f} lbfgs {st['ms_lbfgs']/tk:.3f} gen {st['ms_gen']/tk:.3f} | "
      f"avg live {st['slot_ticks']/tk:.0f} hist GB/s {st['hist_bytes']/max(st['ms_lbfgs'],1e-9)/1e6:.0f}")
