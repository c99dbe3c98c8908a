"""Dev probe: where a grouped scenario sweep spends its time (field builds vs the shared solve vs the rest)."""
import sys, os, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import numpy as np
import topay_b200 as tp
from topay_b200 import scenes
G, C = 256, 8
opt, rp = tp.opt_params_default(), tp.robot_params_default()
fields = [tp.GridMap(tp.grid_desc()) for _ in range(G)]
for f in fields: f.set_keep_sqdist(False)
clouds = [(scenes.tables_scene(s)[0] if s % 2 == 0 else scenes.cuboids_scene(s)[0]) for s in range(G)]
cands = [scenes.short_candidates(C, 100000 + s) for s in range(G)]
print("points per cloud", np.mean([len(c) for c in clouds]))
solver = tp.MomaTrajOpt(fields[0], max_cand=G * C, max_pieces=16, opt_param=opt, robot=rp)
for rep in range(2):
    t0 = time.perf_counter()
    for f, c in zip(fields, clouds): f.regenerateMap(c)
    t1 = time.perf_counter()
    paths = [q for p in cands for q in p[0]]
    bv = np.concatenate([np.asarray(p[1]).reshape(len(p[0]), 10, 2) for p in cands]); ba = np.concatenate([np.asarray(p[2]).reshape(len(p[0]), 10, 2) for p in cands])
    solver.upload(paths, bv, ba)
    solver.assign_fields(fields, np.repeat(np.arange(G), C))
    t2 = time.perf_counter()
    solver.run()
    t3 = time.perf_counter()
    st = solver.stats()
    ws = solver.planWinners(cands, use_gate=True, fields=fields)
    t4 = time.perf_counter()
    print(f"rep {rep}: fields {1e3*(t1-t0):.0f} ms ({1e3*(t1-t0)/G:.2f} per scenario) | upload+assign {1e3*(t2-t1):.0f} | run {1e3*(t3-t2):.0f} ms, ticks {st['ticks']}, "
          f"{1e3*st['ms_total']/st['ticks']:.0f} us/tick | whole planWinners {1e3*(t4-t3):.0f} ms")
