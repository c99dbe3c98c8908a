"""One 256-candidate plan of the bench workload, solved once — the lightest process that shows the
kernels at their bench geometry (for `ncu --set full`, which replays every profiled launch ~40 times)."""
import sys
import topay_b200 as tp
from topay_b200 import scenes

n_cand = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pts, _ = scenes.cuboids_scene(42)
gm = tp.GridMap(tp.grid_desc())
gm.regenerateMap(pts)
opt, rp = tp.opt_params_default(), tp.robot_params_default()
opt.int_K, opt.min_piece_num, opt.sample_interval = 32, 64, 1e9
paths, bv, ba = scenes.synthetic_batch(n_cand, 1234)
solver = tp.MomaTrajOpt(gm, max_cand=n_cand, max_pieces=64, opt_param=opt, robot=rp)
solver.upload(paths, bv, ba)
solver.set_timed(len(sys.argv) > 2 and sys.argv[2] == "timed")
solver.run()
print(solver.stats())
