"""Device time of ESDFMap::updateESDF3D at the full 803x803x83 ring + query rate."""
import time
import numpy as np
import topay_b200 as tp
from topay_b200.rog import ESDFMap

dev = ESDFMap(tp.rog_desc())
rng = np.random.default_rng(0)
res = dev.resolution
# ~ the cuboid scene density: 60 boxes of 1 x 1 x 1.5 m
pts = []
for _ in range(60):
    c = rng.uniform([-18, -18, 0], [18, 18, 0.2])
    g = np.stack(np.meshgrid(np.arange(0, 1, res), np.arange(0, 1, res), np.arange(0, 1.5, res), indexing="ij"), -1)
    pts.append(c + g.reshape(-1, 3))
pts = np.concatenate(pts)
dev.updateGridCounter(pts, 1, 3)
for odom in [(0, 0, 0), (0, 0, 0), (0.5, 0.3, 0.0), (0.5, 0.3, 0.0)]:
    dev.mapSliding(odom)
    t = time.time()
    dev.updateESDF3D(odom)
    w = time.time() - t
    print("update_esdf odom", odom, "device ms total/3d", dev.last_update_ms(), "wall ms", round(w * 1e3, 2))
q = rng.uniform([-19, -19, 0], [19, 19, 1.9], (2_000_000, 3))
t = time.time()
d, g = dev.getValueGrad(q)
print("2M host queries incl. copies: %.1f ms" % ((time.time() - t) * 1e3), "min d", d.min())
