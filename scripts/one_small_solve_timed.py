import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import numpy as np
import topay_b200 as tp
from topay_b200 import scenes
pts, _ = scenes.cuboids_scene(42)
gm = tp.GridMap(tp.grid_desc()); gm.regenerateMap(pts)
paths, bv, ba = scenes.short_candidates(4, 7)
solver = tp.MomaTrajOpt(gm, max_cand=4, max_pieces=16)
solver.set_timed(True)
r = solver.optimizeTrajBatch(paths, bv, ba)
print("timed evals", r["evals"], "cost", r["cost"])
