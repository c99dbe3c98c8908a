"""Dev probe: config 3 (256 candidates x 64 pieces x K=32) timing breakdown on one GPU."""
import sys, os, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import numpy as np
import topay_b200 as tp
from topay_b200 import scenes

C = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pts, _ = scenes.cuboids_scene(42)
gm = tp.GridMap(tp.grid_desc()); gm.regenerateMap(pts)
print("field rebuild ms", gm.last_rebuild_ms())
opt, rp = tp.opt_params_default(), tp.robot_params_default()
opt.int_K = 32; opt.min_piece_num = 64; opt.sample_interval = 1e9
paths, bv, ba = scenes.synthetic_batch(C, 1234)
solver = tp.MomaTrajOpt(gm, max_cand=C, max_pieces=64, opt_param=opt, robot=rp)
solver.set_timed(os.environ.get('TIMED', '1') == '1')
t = time.time(); solver.upload(paths, bv, ba); print("upload s", time.time() - t)
import ctypes
solver._l.topay_solver_phase_clocks(solver.h, 1, None)
t = time.time(); solver.run(); w = time.time() - t
clk = (ctypes.c_longlong * 16)(); solver._l.topay_solver_phase_clocks(solver.h, 1, clk)
names = ["adj:load+terms", "adj:misc+jerk", "adj:wk", "adj:sweeps", "adj:grad", "adv:ls+hist", "adv:loop1", "adv:loop2", "adv:rest", "gen:fill", "gen:lu", "gen:sweeps", "gen:store"]
tk = solver.stats()["ticks"]
print("k_cand phases of candidate 0, us per tick @1.965GHz:", {n: round(clk[i] / 1965.0 / tk, 1) for i, n in enumerate(names)}, "sum", round(sum(clk[:13]) / 1965.0 / tk, 1))
r = solver.download(); st = solver.stats()
print("run wall s", w, "stats", st)
print("status ok", int(r["status"].sum()), "/", C, "pieces", set(r["piece_num"].tolist()))
print("evals min/med/max", r["evals"].min(), int(np.median(r["evals"])), r["evals"].max(), "iters med", int(np.median(r["iters"])), "alm", np.bincount(r["alm_rounds"]))
print("codes", np.unique(r["lbfgs_code"], return_counts=True))
print("traj/s", C / (st["ms_total"] / 1e3), "ms/tick", st["ms_total"] / st["ticks"], "ms penalty/launch", st["ms_eval"] / st["eval_launches"])
