"""Generates the golden vectors under tests/golden/ from the oracle.

The reference ships no golden vectors, known-answer tests or fixtures for this path and cannot be
built or imported in this environment (C++/ROS/Eigen), so these files pin the ORACLE's outputs
(parity stays "unpinned" with respect to the reference itself): they detect drift
of the restatement and give the GPU tests committed targets that do not depend on rebuilding the
oracle. Schema after the reference's datapoint fields (src/planner/include/utils/data.hpp:24-40):
occ_2d, occ_3d, esdf_2d, esdf_3d, traj.   Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib as O  # noqa: E402
from topay_b200 import scenes  # noqa: E402
from topay_b200._structs import grid_desc  # noqa: E402


def field_fixture():
    rng = np.random.default_rng(2024)
    shape, res = (24, 20, 10), 0.1
    desc = grid_desc(map_size=tuple((s - 0.5) * res for s in shape), resolution=res)
    f = O.Field(desc)
    occ3 = (rng.random(shape) < 0.04).astype(np.int8)
    occ3[3:6, 4:9, 0:5] = 1
    occ2 = occ3[:, :, :2].max(axis=2)
    occ2c = occ3.max(axis=2)
    f.set_occupancy(occ3, occ2, occ2c)
    f.rebuild()
    pos = rng.uniform([-1.3, -1.1, -0.1], [1.3, 1.1, 1.1], (256, 3))
    d3, g3 = f.query3d(pos)
    d2, g2 = f.query2d(pos[:, :2], 0)
    np.savez_compressed(os.path.join(HERE, "field_small.npz"), map_size=np.array(desc.map_size[:]), resolution=res,
                        occ_3d=occ3, occ_2d=occ2, occ_2d_critical=occ2c,
                        sq_pos_3d=f.download_sqdist(3)[0], sq_neg_3d=f.download_sqdist(3)[1],
                        esdf_3d=f.download(3), esdf_2d=f.download(0), esdf_2d_inflate=f.download(1),
                        esdf_2d_critical=f.download(2), query_pos=pos, query_d3=d3, query_g3=g3, query_d2=d2,
                        query_g2=g2)


def eval_fixture():
    pts, _ = scenes.cuboids_scene(42)
    desc = grid_desc()
    f = O.Field(desc)
    f.rasterize(pts)
    f.rebuild()
    rp = O.robot_defaults()
    out = {}
    rng = np.random.default_rng(7)
    case = 0
    for K in (12, 32):
        opt = O.opt_defaults()
        opt.int_K = K
        for min_pieces in (3, 8):
            opt.min_piece_num = min_pieces
            paths, bv, ba = scenes.short_candidates(2, 50 + K + min_pieces)
            for p in paths:
                pr = O.prepare_candidate(opt, rp, p, bv[0], ba[0], 16)
                N = pr["piece_num"]
                x = pr["x0"] + 0.05 * rng.normal(size=len(pr["x0"]))
                lam, rho = rng.normal(size=2) * 20, np.array([1e4, 3e4])
                for stage in (1, 2):
                    c, g, terms, coeff, fxy = O.eval_one(opt, rp, f, stage, N, pr["head_pva"], pr["tail_pva"],
                                                         pr["start_xy"], pr["end_xy"], pr["init_inner_xy"][:N], lam,
                                                         rho, x)
                    k = f"c{case}"
                    out.update({f"{k}_K": K, f"{k}_min_pieces": min_pieces, f"{k}_stage": stage, f"{k}_path": p,
                                f"{k}_x": x, f"{k}_lambda": lam, f"{k}_rho": rho, f"{k}_cost": c, f"{k}_grad": g,
                                f"{k}_terms": terms, f"{k}_coeff": coeff, f"{k}_final_xy": fxy})
                    case += 1
    out["n_cases"] = case
    np.savez_compressed(os.path.join(HERE, "eval_cases.npz"), **out)


def solve_fixture():
    pts, _ = scenes.cuboids_scene(42)
    f = O.Field(grid_desc())
    f.rasterize(pts)
    f.rebuild()
    rp, opt = O.robot_defaults(), O.opt_defaults()
    paths, bv, ba = scenes.short_candidates(4, 7)
    out = {"n": len(paths)}
    def shape_metrics(r, p):
        """length of the base path (pose table, moma_traj_opt.h:38-68) and the clearances checkFeasible
        accumulates (moma_traj_opt.h:993-1006): chassis (flat 2-D map) and the tightest arm sphere margin."""
        tr = (r["T"], r["coeff"], p[0, :3])
        seq = O.traj_car_seq([tr])[0]
        end = O.traj_sample([tr], [[float(np.sum(r["T"]))]])[0][0, 0]
        xy = np.concatenate([seq[:, :2], end[None, :2]])
        length = float(np.linalg.norm(np.diff(xy, axis=0), axis=1).sum())
        fz = O.check_feasible(f, rp, [tr])
        radii = np.zeros((12, 4))
        import ctypes
        n = O.lib().oracle_colli_pts(ctypes.byref(rp), O._p(np.zeros(10)), O._p(radii))
        arm = float((fz["min_dist_mani"][0, :n] - radii[:n, 3]).min())
        return length, float(fz["min_dist"][0]), arm

    for c, p in enumerate(paths):
        r = O.solve_one(opt, rp, f, p, bv[c], ba[c], trace=True)
        # sensitivity band of the reference algorithm itself: the same solve with the interior waypoints
        # perturbed by +-1e-15 relative (see DESIGN.md "final-trajectory tolerance")
        costs, durs = [r["cost"]], [r["duration"]]
        shp = [shape_metrics(r, p)]
        for eps in (1e-15, -1e-15, 2e-15, -2e-15):
            q = p.copy()
            q[1:-1, :2] *= (1 + eps)
            rr = O.solve_one(opt, rp, f, q, bv[c], ba[c])
            if rr["status"] == 1:
                costs.append(rr["cost"])
                durs.append(rr["duration"])
                shp.append(shape_metrics(rr, q))
        shp = np.array(shp)
        out.update({f"s{c}_path": p, f"s{c}_status": r["status"], f"s{c}_cost": r["cost"],
                    f"s{c}_duration": r["duration"], f"s{c}_piece_num": r["piece_num"],
                    f"s{c}_trace": r["trace"][:40], f"s{c}_cost_band": np.array([min(costs), max(costs)]),
                    f"s{c}_duration_band": np.array([min(durs), max(durs)]),
                    f"s{c}_length_band": np.array([shp[:, 0].min(), shp[:, 0].max()]),
                    f"s{c}_clear_base_band": np.array([shp[:, 1].min(), shp[:, 1].max()]),
                    f"s{c}_clear_arm_band": np.array([shp[:, 2].min(), shp[:, 2].max()])})
    np.savez_compressed(os.path.join(HERE, "solve_cases.npz"), **out)


if __name__ == "__main__":
    import sys
    if "solve" not in sys.argv[1:]:
        field_fixture()
        eval_fixture()
    solve_fixture()
    for fn in sorted(os.listdir(HERE)):
        print(fn, os.path.getsize(os.path.join(HERE, fn)))
