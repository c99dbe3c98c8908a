"""Dev probe: bench.py's sweep leg at several group sizes (one GPU)."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import bench
import topay_b200 as tp
from topay_b200 import scenes
for G, W in ((256, 2), (512, 1), (256, 1)):
    r = bench.sweep_probe(tp, scenes, 0, 0, 1, None, per_rank=512, group=G, workers=W)
    print(G, W, round(r["scenarios_per_s"], 1), "scenarios/s", round(r["seconds"], 2), "s", r["scenarios_with_a_feasible_winner"], flush=True)
