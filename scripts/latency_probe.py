"""p50 latency of reference-scale plans (bench.py's latency probe alone)."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import bench
import topay_b200 as tp
from topay_b200 import scenes
pts, _ = scenes.cuboids_scene(42)
gm = tp.GridMap(tp.grid_desc()); gm.regenerateMap(pts)
print(bench.latency_probe(tp, scenes, gm))
