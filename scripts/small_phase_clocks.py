"""Dev probe: clock64() deltas of k_cand's phases for candidate 0 of a reference-scale plan (topay_solver_phase_clocks)."""
import os, sys, ctypes as C
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import numpy as np
import topay_b200 as tp
from topay_b200 import scenes, _lib
pts, _ = scenes.cuboids_scene(42)
gm = tp.GridMap(tp.grid_desc()); gm.regenerateMap(pts)
opt, rp = tp.opt_params_default(), tp.robot_params_default()
s = tp.MomaTrajOpt(gm, max_cand=8, max_pieces=16, opt_param=opt, robot=rp)
names = ["0 load/assemble", "1 adj sweeps", "2 grad", "3 ", "4 adj end", "5 ls/state machine + new pair", "6", "7 two-loop",
         "8 advance rest", "9 fill", "10 LU", "11 solve sweeps", "12 store", "13", "14", "15"]
for plan in range(2):
    paths, bv, ba = scenes.short_candidates(8, 5000 + plan)
    s.optimizeTrajBatch(paths, bv, ba)
    out = (C.c_longlong * 16)()
    _lib.check(s._l.topay_solver_phase_clocks(s.h, 1, out), "phase_clocks")
    r = s.optimizeTrajBatch(paths, bv, ba)
    _lib.check(s._l.topay_solver_phase_clocks(s.h, 0, out), "phase_clocks")
    st = s.stats()
    tot = sum(out)
    print(f"plan {plan}: ticks {st['ticks']} evals(c0) {r['evals'][0]} iters(c0) {r['iters'][0]} N {r['piece_num'][0]}  total {tot/1.965e3/max(r['evals'][0],1):.1f} us per evaluation of candidate 0")
    for n, v in zip(names, out):
        if v: print(f"   {n:32s} {v/1.965e3/max(r['evals'][0],1):7.2f} us/eval  {v/tot:6.1%}")
