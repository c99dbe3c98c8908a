"""The per-thread device functions (topay_b200/csrc/*.cuh, TP_HD) instantiated on the CPU by
tests/host_harness.cpp and walked in the kernels' dataflow, against the oracle. Catches maths
errors in the device code without a GPU; the lane mapping itself is covered by the -m gpu tests."""
import numpy as np
import pytest

import host_harness as HH
from topay_b200 import scenes


@pytest.fixture(scope="module")
def harness(oracle, small_scene):
    return HH.Harness(oracle.opt_defaults(), oracle.robot_defaults(), small_scene["desc"], small_scene["field"])


def test_queries_bit_exact(oracle, small_scene, harness):
    f = small_scene["field"]
    rng = np.random.default_rng(0)
    pos = rng.uniform([-10.5, -10.5, -0.2], [10.5, 10.5, 1.8], (20000, 3))
    d0, g0 = f.query3d(pos)
    d1, g1 = harness.query3d(pos)
    assert np.array_equal(d0, d1) and np.array_equal(g0, g1)
    for w in range(3):
        d0, g0 = f.query2d(pos[:, :2], w)
        d1, g1 = harness.query2d(pos[:, :2], w)
        assert np.array_equal(d0, d1) and np.array_equal(g0, g1)


def test_fk_and_linear_time_adjoint(oracle, harness):
    rp = oracle.robot_defaults()
    rng = np.random.default_rng(1)
    for _ in range(100):
        p10 = np.concatenate([rng.uniform(-5, 5, 2), rng.uniform(-3, 3, 1), scenes.random_joints(rng)])
        assert np.abs(oracle.colli_pts(rp, p10)[:, :3] - harness.fk(p10)).max() < 1e-14
        g = rng.normal(size=(12, 3))
        a, b = oracle.colli_grads(rp, p10, g), harness.fk_adjoint(p10, g)
        assert np.abs(a - b).max() <= 1e-13 * np.abs(a).max()


@pytest.mark.parametrize("K", [12, 32, 5])
def test_penalty_dataflow(oracle, small_scene, K):
    rp, opt = oracle.robot_defaults(), oracle.opt_defaults()
    opt.int_K = K
    f = small_scene["field"]
    H = HH.Harness(opt, rp, small_scene["desc"], f)
    rng = np.random.default_rng(K)
    paths, bv, ba = scenes.short_candidates(3, 11)
    lam, rho = np.array([30.0, -20.0]), np.array([1e4, 2e4])
    for p in paths:
        pr = oracle.prepare_candidate(opt, rp, p, bv[0], ba[0], 64)
        N = pr["piece_num"]
        x = pr["x0"] + 0.05 * rng.normal(size=len(pr["x0"]))
        T = np.array([(0.5 * t + 1) * t + 1 if t > 0 else 1 / ((0.5 * t - 1) * t + 1) for t in x[:N]])
        for stage in (1, 2):
            coeff = oracle.eval_one(opt, rp, f, stage, N, pr["head_pva"], pr["tail_pva"], pr["start_xy"],
                                    pr["end_xy"], pr["init_inner_xy"][:N], lam, rho, x)[3]
            _, gdC, gdT, tm, fx0 = oracle.penalty_only(opt, rp, f, stage, N, coeff, T, pr["start_xy"], pr["end_xy"],
                                                       pr["init_inner_xy"][:N], lam, rho)
            gdC1, gdT1, tm1, fx1 = H.penalty_eval(stage, N, coeff, T, pr["start_xy"], pr["end_xy"],
                                                  pr["init_inner_xy"][:N], lam, rho)
            assert np.abs(gdC - gdC1).max() <= 1e-12 * np.abs(gdC).max()
            assert np.abs(gdT - gdT1).max() <= 1e-12 * np.abs(gdT).max()
            assert np.abs(tm - tm1).max() <= 1e-12 * np.abs(tm).max()
            assert np.abs(fx0 - fx1).max() < 1e-13
