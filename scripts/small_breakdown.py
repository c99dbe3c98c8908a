"""Dev probe: per-kernel device time of a reference-scale plan (8 candidates, int_K 12, 3-10 pieces), timed mode
(plain launches, events around every kernel), and the graph-replay time of the same plan next to it."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import numpy as np
import topay_b200 as tp
from topay_b200 import scenes
pts, _ = scenes.cuboids_scene(42)
gm = tp.GridMap(tp.grid_desc()); gm.regenerateMap(pts)
opt, rp = tp.opt_params_default(), tp.robot_params_default()
s = tp.MomaTrajOpt(gm, max_cand=8, max_pieces=16, opt_param=opt, robot=rp)
for plan in range(3):
    paths, bv, ba = scenes.short_candidates(8, 5000 + plan)
    s.set_timed(False)
    s.optimizeTrajBatch(paths, bv, ba)
    t0 = time.perf_counter(); r = s.optimizeTrajBatch(paths, bv, ba); dt = (time.perf_counter() - t0) * 1e3
    st = s.stats(); tk = st["ticks"]
    print(f"plan {plan}: graph {dt:.1f} ms, device {st['ms_total']:.1f} ms, ticks {tk}, {1e3*st['ms_total']/tk:.1f} us/tick, max evals {r['evals'].max()} pieces {r['piece_num']}")
    s.set_timed(True)
    s.optimizeTrajBatch(paths, bv, ba)
    st = s.stats(); tk = st["ticks"]
    print(f"   timed: {1e3*st['ms_total']/tk:.1f} us/tick | integrate {1e3*st['ms_integrate']/tk:.1f} penalty {1e3*st['ms_eval']/tk:.1f} "
          f"chain {1e3*st['ms_chain']/tk:.1f} adj {1e3*st['ms_adj']/tk:.1f} lbfgs {1e3*st['ms_lbfgs']/tk:.1f} gen {1e3*st['ms_gen']/tk:.1f}")
