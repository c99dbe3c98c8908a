"""Dev probe: the same 8 x 256 candidates through P concurrent slot pools (P solvers on P streams, P host threads)."""
import sys, os, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import topay_b200 as tp
from topay_b200 import scenes

pts, _ = scenes.cuboids_scene(42)
gm = tp.GridMap(tp.grid_desc(), device=0)
gm.regenerateMap(pts)
opt, rp = tp.opt_params_default(), tp.robot_params_default()
opt.int_K, opt.min_piece_num, opt.sample_interval = 32, 64, 1e9
PL, C = 8, 256
batches = [scenes.synthetic_batch(C, 1234 + 1000 * p) for p in range(PL)]
for P, total_slots in ((1, 1024), (2, 1024), (4, 1024), (2, 1536), (2, 2048), (1, 2048)):
    per = PL // P
    solvers = []
    for i in range(P):
        bs = batches[i * per:(i + 1) * per]
        paths = [q for b in bs for q in b[0]]
        bv = np.concatenate([b[1] for b in bs]); ba = np.concatenate([b[2] for b in bs])
        s = tp.MomaTrajOpt(gm, max_cand=per * C, max_pieces=64, opt_param=opt, robot=rp, n_slots=min(total_slots // P, per * C))
        s.upload(paths, bv, ba)
        solvers.append(s)
    def go():
        th = [threading.Thread(target=s.run) for s in solvers]
        t0 = time.perf_counter()
        [t.start() for t in th]; [t.join() for t in th]
        return time.perf_counter() - t0
    go()
    dt = go()
    ok = sum(int(s.download()["status"].sum()) for s in solvers)
    print(f"pools {P} x {total_slots // P} slots: {dt:6.3f} s  {PL * C / dt:7.1f} traj/s ok {ok}", flush=True)
    for s in solvers: s.close()
