"""Race probe: the same candidates solved alone, inside a larger batch, and with plain launches instead of
graph replay must give bit-identical per-candidate results."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import numpy as np
import topay_b200 as tp
from topay_b200 import scenes
pts, _ = scenes.cuboids_scene(42)
gm = tp.GridMap(tp.grid_desc()); gm.regenerateMap(pts)
p4, bv4, ba4 = scenes.short_candidates(4, 7)
p8 = p4 + scenes.short_candidates(12, 99)[0]
bv8, ba8 = np.zeros((16, 10, 2)), np.zeros((16, 10, 2))
s4 = tp.MomaTrajOpt(gm, max_cand=4, max_pieces=16)
s16 = tp.MomaTrajOpt(gm, max_cand=16, max_pieces=16)
a = s4.optimizeTrajBatch(p4, bv4, ba4)
b = s16.optimizeTrajBatch(p8, bv8, ba8)
s4.set_timed(True)
c = s4.optimizeTrajBatch(p4, bv4, ba4)
print("alone   ", a["evals"], a["cost"])
print("in 16   ", b["evals"][:4], b["cost"][:4])
print("timed   ", c["evals"], c["cost"])
