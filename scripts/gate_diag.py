"""Which constraint fails the success gate on the golden solve cases (device vs oracle trajectories)."""
import os, sys
import numpy as np
sys.path.insert(0, "tests")
import oracle_lib as O
import topay_b200 as tp
from topay_b200 import scenes
G = "tests/golden"
pts, boxes = scenes.cuboids_scene(42)
gm = tp.GridMap(tp.grid_desc()); gm.regenerateMap(pts)
z = np.load(os.path.join(G, "solve_cases.npz")); n = int(z["n"])
paths = [z[f"s{c}_path"] for c in range(n)]
solver = tp.MomaTrajOpt(gm, max_cand=n, max_pieces=16)
res = solver.optimizeTrajBatch(paths, np.zeros((n, 10, 2)), np.zeros((n, 10, 2)))
arrs, best = solver.checkFeasibleBatch()
rp = tp.robot_params_default()
np.set_printoptions(precision=4, suppress=True, linewidth=200)
print("status", res["status"], "best", best, "dur", res["duration"])
for k in ("feasible", "feasible_print", "max_vel", "max_acc", "max_domega", "max_d2omega", "min_dist"):
    print(k, arrs[k])
print("limits v a w dw", rp.max_v, rp.max_a, rp.max_w, rp.max_dw, "chassis r", rp.chassis_colli_radius)
print("max_q/lim", arrs["max_q"] / np.array(rp.joint_pos_limit_max))
print("max_dq/lim", arrs["max_dq"] / np.array(rp.joint_vel_limit))
print("max_d2q/lim", arrs["max_d2q"] / np.array(rp.joint_acc_limit))
print("min_dist_mani", arrs["min_dist_mani"])
