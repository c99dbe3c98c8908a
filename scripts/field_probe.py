"""Dev probe: BASELINE configs[3] — ESDF rebuild + query on 800 x 800 x 80 @ 0.05 m."""
import sys, os, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import numpy as np
import topay_b200 as tp
from topay_b200 import scenes
desc = tp.grid_desc(map_size=(40.0, 40.0, 4.0), resolution=0.05)
gm = tp.GridMap(desc)
pts, _ = scenes.cuboids_scene(7, size_x=40.0, size_y=40.0, scale=2.0)
gm.regenerateMap(pts)
print("occupied fraction", gm.getOccBuffer3d().mean())
keep = os.environ.get("KEEP_SQ", "1") == "1"
gm.set_keep_sqdist(keep)
for _ in range(3):
    gm.updateESDF()
    print("rebuild ms (total, 3-D):", gm.last_rebuild_ms())
tot, t3 = gm.last_rebuild_ms()
vox = 800 * 800 * 80
print(f"3-D: {vox * 9 / (t3 * 1e-3) / 1e9:.1f} GB/s algorithmic (9 B/voxel) = {vox * 9 / (t3 * 1e-3) / 1e9 / 6550.1:.3f} of measured HBM peak")
