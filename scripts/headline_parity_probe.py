"""Dev probe: the headline configuration (64 pieces x int_K 32) solved on the device and by the oracle (which is
bit-identical to the reference, tests/test_ref_pin.py): statuses, rounds, costs, durations, gate verdicts, and the
oracle's own band under a 1e-15 relative perturbation of the waypoints."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
import topay_b200 as tp
from topay_b200 import scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
pts, _ = scenes.cuboids_scene(42)
desc = tp.grid_desc()
gm = tp.GridMap(desc, device=0); gm.regenerateMap(pts)
of = O.Field(desc); of.rasterize(pts); of.rebuild()
opt, rp = tp.opt_params_default(), tp.robot_params_default()
opt.int_K, opt.min_piece_num, opt.sample_interval = 32, 64, 1e9
paths, bv, ba = scenes.synthetic_batch(256, 1234)
paths, bv, ba = paths[:n], bv[:n], ba[:n]
s = tp.MomaTrajOpt(gm, max_cand=n, max_pieces=64, opt_param=opt, robot=rp)
s.set_trace(20000)
r = s.optimizeTrajBatch(paths, bv, ba)
arrs, best = s.checkFeasibleBatch()
import threading
outs = [None] * n; outs_p = [None] * n
def work(c):
    outs[c] = O.solve_one(opt, rp, of, paths[c], bv[c], ba[c], max_pieces=64, trace=True)
    pp = paths[c].copy(); pp[1:-1] *= (1 + 1e-15)
    outs_p[c] = O.solve_one(opt, rp, of, pp, bv[c], ba[c], max_pieces=64)
t0 = time.time()
th = [threading.Thread(target=work, args=(c,)) for c in range(n)]
[t.start() for t in th]; [t.join() for t in th]
print("oracle time", time.time() - t0)
for c in range(n):
    o, p = outs[c], outs_p[c]
    g = O.check_feasible(of, rp, [(o["T"], o["coeff"], paths[c][0, :3])])
    gp = O.check_feasible(of, rp, [(p["T"], p["coeff"], paths[c][0, :3])])
    tg, tc = s.trace(c), o["trace"]
    m = min(len(tg), len(tc))
    same = 0
    while same < m and np.array_equal(tg[same, 2:], tc[same, 2:]) and abs(tg[same, 0] - tc[same, 0]) <= 1e-9 * abs(tc[same, 0]):
        same += 1
    print(f"cand {c}: GPU status {r['status'][c]} rounds {r['alm_rounds'][c]} evals {r['evals'][c]} cost {r['cost'][c]:.6g} dur {r['duration'][c]:.4f} "
          f"gate {arrs['feasible'][c]}/{arrs['feasible_print'][c]} | oracle status {o['status']} rounds {o['alm_rounds']} evals {o['evals']} "
          f"cost {o['cost']:.6g} dur {o['duration']:.4f} gate {g['feasible'][0]}/{g['feasible_print'][0]} | perturbed oracle status {p['status']} "
          f"rounds {p['alm_rounds']} evals {p['evals']} cost {p['cost']:.6g} dur {p['duration']:.4f} gate {gp['feasible'][0]}/{gp['feasible_print'][0]} "
          f"| trace agrees for {same} accepted iterations (of {len(tg)} / {len(tc)})")
