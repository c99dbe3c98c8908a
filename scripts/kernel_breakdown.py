"""Live per-kernel device time of a solve (timed mode: CUDA events around every launch), bench workload
(256 x 64 x K32) and the reference-scale small plan. Usage: PYTHONPATH=. python scripts/kernel_breakdown.py [n_cand]"""
import sys
import numpy as np
import topay_b200 as tp
from topay_b200 import scenes

n_cand = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pts, _ = scenes.cuboids_scene(42)
gm = tp.GridMap(tp.grid_desc())
gm.regenerateMap(pts)
opt, rp = tp.opt_params_default(), tp.robot_params_default()
opt.int_K = 32
opt.min_piece_num = 64
opt.sample_interval = 1e9
paths, bv, ba = scenes.synthetic_batch(n_cand, 1234)
solver = tp.MomaTrajOpt(gm, max_cand=n_cand, max_pieces=64, opt_param=opt, robot=rp)
solver.upload(paths, bv, ba)
for timed in (False, True):
    solver.set_timed(timed)
    solver.run()
    st = solver.stats()
    print("timed" if timed else "graph", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in st.items()})
t = st["ticks"]
print("per tick us: integrate %.1f penalty %.1f chain %.1f cand %.1f | total %.1f" % (
    1e3 * st["ms_integrate"] / t, 1e3 * st["ms_eval"] / t, 1e3 * st["ms_chain"] / t, 1e3 * st["ms_cand"] / t,
    1e3 * st["ms_total"] / t))
# small plans
o2 = tp.opt_params_default()
s2 = tp.MomaTrajOpt(gm, max_cand=8, max_pieces=16, opt_param=o2, robot=rp)
p8, bv8, ba8 = scenes.short_candidates(8, 5003)
s2.upload(p8, bv8, ba8)
for timed in (False, True):
    s2.set_timed(timed)
    s2.run()
    st = s2.stats()
    print("small", "timed" if timed else "graph", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in st.items()})
t = st["ticks"]
print("small per tick us: integrate %.1f penalty %.1f chain %.1f cand %.1f | total %.1f" % (
    1e3 * st["ms_integrate"] / t, 1e3 * st["ms_eval"] / t, 1e3 * st["ms_chain"] / t, 1e3 * st["ms_cand"] / t,
    1e3 * st["ms_total"] / t))
