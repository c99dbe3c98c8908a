"""Bit-level fingerprint of one stage-1 and one stage-2 evaluation (for A/B runs under tools)."""
import os, sys, hashlib
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import numpy as np
import topay_b200 as tp
from topay_b200 import scenes
pts, _ = scenes.cuboids_scene(42)
gm = tp.GridMap(tp.grid_desc()); gm.regenerateMap(pts)
opt, rp = tp.opt_params_default(), tp.robot_params_default()
paths, bv, ba = scenes.short_candidates(4, 7)
solver = tp.MomaTrajOpt(gm, max_cand=4, max_pieces=16, opt_param=opt, robot=rp)
prep = [tp.prepare_candidate(opt, rp, p, bv[0], ba[0], 16) for p in paths]
rng = np.random.default_rng(0)
xs = [q["x0"] + 0.05 * rng.normal(size=len(q["x0"])) for q in prep]
dump = {}
for stage in (1, 2):
    ev = solver.evaluate(stage, [q["piece_num"] for q in prep], [q["head_pva"] for q in prep],
                         [q["tail_pva"] for q in prep], [q["start_xy"] for q in prep], [q["end_xy"] for q in prep],
                         [q["init_inner_xy"] for q in prep], xs, alm_lambda=np.zeros((4, 2)), alm_rho=np.full((4, 2), 1e4))
    dump[f"g{stage}"] = ev["grad"]
    import ctypes as C
    for which, name in enumerate(("gnode", "gsum", "gdC", "gdT", "tot", "Ixy", "gdev", "dbg")):
        buf = np.zeros(4 * 16 * 13 * 2 * 30)
        n = solver._l.topay_solver_debug_download(solver.h, which, buf.ctypes.data_as(C.POINTER(C.c_double)), len(buf))
        dump[f"{name}{stage}"] = buf[:n].copy()
    dump["pn"] = np.array([q["piece_num"] for q in prep])
    for c in range(4):
        print(stage, c, ev["cost"][c].hex(), hashlib.md5(ev["grad"][c].tobytes()).hexdigest()[:10],
              hashlib.md5(ev["coeff"][c].tobytes()).hexdigest()[:10], " ".join(t.hex()[:22] for t in ev["terms"][c]))

if len(sys.argv) > 1:
    np.savez(os.path.join(R, "gpurun_out", sys.argv[1]), **dump)
