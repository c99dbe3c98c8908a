import sys; import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests'))
import numpy as np, oracle_lib as O, topay_b200 as tp
from topay_b200 import scenes
pts,_ = scenes.cuboids_scene(42); desc = tp.grid_desc()
gm = tp.GridMap(desc); gm.regenerateMap(pts)
of = O.Field(desc); of.rasterize(pts); of.rebuild()
opt, rp = tp.opt_params_default(), tp.robot_params_default()
paths, bv, ba = scenes.short_candidates(8, 7)
solver = tp.MomaTrajOpt(gm, max_cand=8, max_pieces=16, opt_param=opt, robot=rp)
solver.set_trace(4000)
res = solver.optimizeTrajBatch(paths, bv, ba)
for c in range(3):
    tg = solver.trace(c)
    r = O.solve_one(opt, rp, of, paths[c], bv[c], ba[c], trace=True); tc = r['trace']
    m = min(len(tg), len(tc))
    rel = np.abs(tg[:m,0]-tc[:m,0])/np.abs(tc[:m,0])
    same_ls = (tg[:m,3]==tc[:m,3]) & (tg[:m,2]==tc[:m,2])
    first_bad = np.argmax(~same_ls) if (~same_ls).any() else m
    print(f"cand {c}: len gpu {len(tg)} cpu {len(tc)}; first k/ls mismatch at entry {first_bad}; rel f err at 0,10,50,100: {[float('%.1e'%rel[i]) for i in (0,10,50,100) if i<m]}")
    print("   gpu:", tg[max(0,first_bad-2):first_bad+3].tolist())
    print("   cpu:", tc[max(0,first_bad-2):first_bad+3].tolist())
