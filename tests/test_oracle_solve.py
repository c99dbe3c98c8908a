"""Pins the solve half of the oracle: banded LU vs a dense solver, FK adjoint vs finite differences,
cost gradient vs finite differences (with the reference's two inexact-adjoint quirks switched off),
L-BFGS on known-answer problems, and the pre-processing invariants."""
import numpy as np
import pytest

from topay_b200 import scenes
from topay_b200._structs import num_vars


def test_banded_lu_matches_dense(oracle):
    rng = np.random.default_rng(0)
    n, p, q = 30, 6, 6
    A = np.zeros((n, n))
    for i in range(n):
        for j in range(max(0, i - p), min(n, i + q + 1)):
            A[i, j] = rng.normal()
        A[i, i] += 8.0
    B = rng.normal(size=(n, 9))
    assert np.allclose(oracle.banded_solve(A, p, q, B), np.linalg.solve(A, B), rtol=1e-11, atol=1e-12)
    assert np.allclose(oracle.banded_solve(A, p, q, B, adjoint=True), np.linalg.solve(A.T, B), rtol=1e-11, atol=1e-12)


def test_minco_interpolates_and_is_c4(oracle):
    rng = np.random.default_rng(1)
    N = 6
    ew = np.array([0.33] + [1.0] * 8)
    head, tail = rng.normal(size=(9, 3)), rng.normal(size=(9, 3))
    inner, T = rng.normal(size=(N - 1, 9)), rng.uniform(0.5, 2.0, N)
    c, jerk, gdC, gdT = oracle.minco_generate(N, ew, head, tail, inner, T)
    c = c.reshape(N, 6, 9)

    def ev(i, t, der):
        k = np.arange(6)
        coef = np.array([np.prod(k_ - np.arange(der)) if k_ >= der else 0.0 for k_ in k])
        return (coef * np.power(t, np.maximum(k - der, 0))) @ c[i]

    for der in range(3):
        assert np.allclose(ev(0, 0.0, der), head[:, der])
        assert np.allclose(ev(N - 1, T[-1], der), tail[:, der])
    for i in range(N - 1):
        assert np.allclose(ev(i, T[i], 0), inner[i]) and np.allclose(ev(i + 1, 0.0, 0), inner[i])
        for der in range(1, 5):
            assert np.allclose(ev(i, T[i], der), ev(i + 1, 0.0, der), atol=1e-8)
    assert jerk > 0
    # dJ/dT by finite differences of the energy at fixed coefficients is gdT; dJ/dc is gdC
    h = 1e-6
    for i in (0, 3):
        Tp, Tm = T.copy(), T.copy()
        Tp[i] += h
        Tm[i] -= h
        # total derivative through c(q,T): use the adjoint map
        jp = oracle.minco_generate(N, ew, head, tail, inner, Tp)[1]
        jm = oracle.minco_generate(N, ew, head, tail, inner, Tm)[1]
        gdP, gdTail, gT = oracle.minco_backprop(N, ew, head, tail, inner, T, gdC, gdT)
        assert abs((jp - jm) / (2 * h) - gT[i]) <= 1e-5 * max(1.0, abs(gT[i]))


def test_fk_adjoint_matches_finite_differences(oracle):
    rp = oracle.robot_defaults()
    rng = np.random.default_rng(2)
    for _ in range(5):
        pos = np.concatenate([rng.uniform(-3, 3, 2), rng.uniform(-3, 3, 1), scenes.random_joints(rng)])
        g = rng.normal(size=(12, 3))
        grad = oracle.colli_grads(rp, pos, g)
        for k in range(10):
            e = np.zeros(10)
            e[k] = 1e-6
            fd = ((oracle.colli_pts(rp, pos + e)[:, :3] - oracle.colli_pts(rp, pos - e)[:, :3]) * g).sum() / 2e-6
            assert abs(fd - grad[k]) < 1e-6 * max(1.0, abs(grad[k]))
    pts = oracle.colli_pts(rp, np.zeros(10))
    assert len(pts) == 12 and np.allclose(pts[:, 0], 0) and np.allclose(pts[:, 1], 0.115)
    assert (np.array(rp.collision_matrix[:]) == -1).sum() == 110          # 55 checked pairs


@pytest.mark.parametrize("stage", [1, 2])
def test_cost_gradient_matches_finite_differences(oracle, small_scene, stage):
    """With exact_chain (debug switch replacing reference quirks 1 and 2 by the exact adjoint) the
    oracle's gradient must be the derivative of its cost; with the quirks it must not be."""
    rp, opt = oracle.robot_defaults(), oracle.opt_defaults()
    f = small_scene["field"]
    paths, bv, ba = scenes.short_candidates(2, 7)
    lam, rho = np.array([3.0, -2.0]), np.array([1e4, 1e4])
    for p in paths:
        pr = oracle.prepare_candidate(opt, rp, p, bv[0], ba[0], 64)
        N, x0 = pr["piece_num"], pr["x0"]
        args = (opt, rp, f, stage, N, pr["head_pva"], pr["tail_pva"], pr["start_xy"], pr["end_xy"],
                pr["init_inner_xy"][:N], lam, rho)
        oracle.lib().oracle_set_exact_chain(1)
        try:
            g = oracle.eval_one(*args, x0)[1]
            rng = np.random.default_rng(0)
            for i in rng.choice(len(x0), 6, replace=False):
                e = np.zeros(len(x0))
                e[i] = 1e-6
                fd = (oracle.eval_one(*args, x0 + e)[0] - oracle.eval_one(*args, x0 - e)[0]) / 2e-6
                assert abs(fd - g[i]) <= 2e-4 * (abs(fd) + abs(g[i]) + 1e-3)
        finally:
            oracle.lib().oracle_set_exact_chain(0)
        gq = oracle.eval_one(*args, x0)[1]
        assert np.abs(gq - g).max() > 1e-6 * np.abs(g).max()     # the reference's gradient is inexact


def test_lbfgs_known_answers(oracle):
    from topay_b200._structs import LbfgsParams
    p = LbfgsParams(mem_size=8, g_epsilon=1e-8, past=0, delta=0.0, max_iterations=0, max_linesearch=64,
                    min_step=1e-20, max_step=1e20, f_dec_coeff=1e-4, s_curv_coeff=0.9, cautious_factor=1e-6,
                    machine_prec=1e-16)
    rng = np.random.default_rng(0)
    a, c = rng.uniform(0.5, 5.0, 20), rng.normal(size=20)
    r, x, f, it, ev = oracle.lbfgs_test_problem(a, c, 0.0, p, np.zeros(20))
    assert r == 0 and np.allclose(x, c, atol=1e-7) and f < 1e-12
    # chained Rosenbrock-like coupling: minimum at x = 1 when a = 0 ... use a = 1, c = 1, b = 10
    r, x, f, it, ev = oracle.lbfgs_test_problem(np.ones(6), np.ones(6), 10.0, p, np.full(6, -0.5))
    assert r == 0 and np.allclose(x, 1.0, atol=1e-6)
    # max_iterations -> LBFGSERR_MAXIMUMITERATION (-1008); past/delta -> LBFGS_STOP (1)
    p.max_iterations = 2
    assert oracle.lbfgs_test_problem(np.ones(6), np.ones(6), 10.0, p, np.full(6, -0.5))[0] == -1008
    p.max_iterations, p.past, p.delta, p.g_epsilon = 0, 2, 1e-1, 0.0
    assert oracle.lbfgs_test_problem(np.ones(6), np.ones(6), 10.0, p, np.full(6, -0.5))[0] == 1
    p.mem_size = 0
    assert oracle.lbfgs_test_problem(a, c, 0.0, p, np.zeros(20))[0] == -1022      # INVALID_MEMSIZE


def test_prepare_candidate_invariants(oracle):
    rp, opt = oracle.robot_defaults(), oracle.opt_defaults()
    paths, bv, ba = scenes.short_candidates(5, 3)
    for p in paths:
        pr = oracle.prepare_candidate(opt, rp, p, bv[0], ba[0], 64)
        N = pr["piece_num"]
        assert N >= opt.min_piece_num and len(pr["x0"]) == num_vars(N)
        assert np.allclose(pr["start_xy"], p[0, :2]) and np.allclose(pr["end_xy"], p[-1, :2])
        assert np.allclose(pr["init_inner_xy"][N - 1], p[-1, :2])
        assert np.all(np.diff(pr["x0"][N + (N - 1):N + (N - 1) + N]) >= -1e-12)       # arc is monotone
    assert oracle.prepare_candidate(opt, rp, paths[0], bv[0], ba[0], 2)["rc"] == -5    # TOO_LARGE
    # a very short path switches stage 1 to the long "past" window (moma_traj_opt.cpp:354-357)
    short = np.array([np.r_[0.0, 0.0, 0.0, np.zeros(7)], np.r_[0.3, 0.0, 0.0, np.zeros(7)]])
    assert oracle.prepare_candidate(opt, rp, short, bv[0], ba[0], 64)["s1_past"] == opt.s1_lbfgs_shot_path_past


def test_solve_reaches_goal(oracle, small_scene):
    rp, opt = oracle.robot_defaults(), oracle.opt_defaults()
    paths, bv, ba = scenes.short_candidates(2, 7)
    for p in paths:
        r = oracle.solve_one(opt, rp, small_scene["field"], p, bv[0], ba[0], trace=True)
        assert r["status"] == 1 and np.linalg.norm(r["final_xy_err"]) < opt.alm_tolerance
        assert abs(r["duration"] - r["T"].sum()) < 1e-9
        tr = r["trace"]
        assert len(tr) == r["iters"] or len(tr) > 0
