"""Run the same small batch several times on fresh and reused solvers; the results must be bit-identical."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import numpy as np
import topay_b200 as tp
from topay_b200 import scenes
pts, _ = scenes.cuboids_scene(42)
gm = tp.GridMap(tp.grid_desc()); gm.regenerateMap(pts)
opt, rp = tp.opt_params_default(), tp.robot_params_default()
paths, bv, ba = scenes.short_candidates(4, 7)
ref = None
for rep in range(6):
    solver = tp.MomaTrajOpt(gm, max_cand=4, max_pieces=16, opt_param=opt, robot=rp) if rep % 2 == 0 else solver
    r = solver.optimizeTrajBatch(paths, bv, ba)
    print(rep, r["evals"], r["cost"])
    if ref is None:
        ref = r
    else:
        same = np.array_equal(ref["evals"], r["evals"]) and np.array_equal(ref["cost"], r["cost"]) and np.array_equal(ref["coeff"], r["coeff"])
        print("   identical to run 0:", same)
