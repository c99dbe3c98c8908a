"""Dev probe: single-plan latency at the reference's default scale (<= 8 candidates, K = 12, N ~ 3-10)."""
import sys, os, time, ctypes
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import numpy as np
import topay_b200 as tp
from topay_b200 import scenes
C = int(sys.argv[1]) if len(sys.argv) > 1 else 8
pts, _ = scenes.cuboids_scene(42)
gm = tp.GridMap(tp.grid_desc()); gm.regenerateMap(pts)
opt, rp = tp.opt_params_default(), tp.robot_params_default()
solver = tp.MomaTrajOpt(gm, max_cand=8, max_pieces=16, opt_param=opt, robot=rp)
lat = []
solver._l.topay_solver_phase_clocks(solver.h, 1, None)
tk = 0
for plan in range(30):
    paths, bv, ba = scenes.short_candidates(C, 1000 + plan)
    t = time.perf_counter(); r = solver.optimizeTrajBatch(paths, bv, ba); lat.append((time.perf_counter() - t) * 1e3)
    tk += solver.stats()["ticks"]
clk = (ctypes.c_longlong * 16)(); solver._l.topay_solver_phase_clocks(solver.h, 1, clk)
names = ["adj:load+terms", "adj:misc+jerk", "adj:wk", "adj:sweeps", "adj:grad", "adv:ls+hist", "adv:loop1", "adv:loop2", "adv:rest", "gen:fill", "gen:lu", "gen:sweeps", "gen:store"]
print("k_cand phases (cand 0) us/tick:", {n: round(clk[i] / 1965.0 / tk, 1) for i, n in enumerate(names)}, "sum", round(sum(clk[:13]) / 1965.0 / tk, 1))
lat = np.array(lat[3:])
print(f"plans of {C} candidates: p50 {np.median(lat):.1f} ms, p90 {np.percentile(lat, 90):.1f} ms, min {lat.min():.1f}; last: pieces {r['piece_num']}, evals {r['evals']}, ok {r['status'].sum()}, ticks/plan {tk/30:.0f}, ms/tick {np.sum(lat)/ (tk*27/30):.3f}")
