"""The oracle against the committed golden vectors (tests/golden/, made by make_golden.py)."""
import os

import numpy as np

from topay_b200._structs import grid_desc

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_field_golden(oracle):
    z = np.load(os.path.join(G, "field_small.npz"))
    f = oracle.Field(grid_desc(map_size=tuple(z["map_size"]), resolution=float(z["resolution"])))
    f.set_occupancy(z["occ_3d"], z["occ_2d"], z["occ_2d_critical"])
    f.rebuild()
    assert np.array_equal(f.download_sqdist(3)[0], z["sq_pos_3d"]) and np.array_equal(f.download_sqdist(3)[1], z["sq_neg_3d"])
    for which, key in ((3, "esdf_3d"), (0, "esdf_2d"), (1, "esdf_2d_inflate"), (2, "esdf_2d_critical")):
        assert np.array_equal(f.download(which), z[key]), key
    d, g = f.query3d(z["query_pos"])
    assert np.array_equal(d, z["query_d3"]) and np.array_equal(g, z["query_g3"])


def test_eval_golden(oracle, small_scene):
    z = np.load(os.path.join(G, "eval_cases.npz"))
    rp = oracle.robot_defaults()
    for c in range(int(z["n_cases"])):
        k = f"c{c}"
        opt = oracle.opt_defaults()
        opt.int_K, opt.min_piece_num = int(z[f"{k}_K"]), int(z[f"{k}_min_pieces"])
        pr = oracle.prepare_candidate(opt, rp, z[f"{k}_path"], np.zeros((10, 2)), np.zeros((10, 2)), 16)
        N = pr["piece_num"]
        f, g, terms, coeff, fxy = oracle.eval_one(opt, rp, small_scene["field"], int(z[f"{k}_stage"]), N,
                                                  pr["head_pva"], pr["tail_pva"], pr["start_xy"], pr["end_xy"],
                                                  pr["init_inner_xy"][:N], z[f"{k}_lambda"], z[f"{k}_rho"], z[f"{k}_x"])
        assert f == float(z[f"{k}_cost"]) and np.array_equal(g, z[f"{k}_grad"]) and np.array_equal(coeff, z[f"{k}_coeff"])


def test_solve_golden(oracle, small_scene):
    z = np.load(os.path.join(G, "solve_cases.npz"))
    rp, opt = oracle.robot_defaults(), oracle.opt_defaults()
    for c in range(int(z["n"])):
        r = oracle.solve_one(opt, rp, small_scene["field"], z[f"s{c}_path"], np.zeros((10, 2)), np.zeros((10, 2)),
                             trace=True)
        assert r["status"] == int(z[f"s{c}_status"]) and r["cost"] == float(z[f"s{c}_cost"])
        assert np.array_equal(r["trace"][:40], z[f"s{c}_trace"])
