"""Pins the oracle restatement to the REFERENCE'S OWN CODE, bit for bit.

oracle/_ref/libtopay_ref.so is the reference's lbfgs.hpp, banded_system.hpp, minco.hpp (MinJerkOpt<9>)
and moma_param.h (getColliPts / getColliGrads) compiled unmodified from /root/reference against the
Eigen / ROS stand-ins of oracle/ref_stubs (oracle/Makefile target _ref, tests/ref_lib.py). The
oracle functions of the same rows must return
IDENTICAL doubles on the same inputs: `np.array_equal`, no tolerance. The one stated limit: the stand-in
sums reductions (dot, norm, sum) sequentially, as the oracle does; real Eigen's packet order may differ
in the last ulp of a reduction.
"""
import numpy as np
import pytest

import ref_lib as R
from topay_b200._structs import LbfgsParams

pytestmark = pytest.mark.skipif(not R.available(), reason="neither /root/reference nor a prebuilt oracle/_ref")


def _lp(**kw):
    d = dict(mem_size=8, g_epsilon=1e-8, past=3, delta=1e-10, max_iterations=0, max_linesearch=64, min_step=1e-20,
             max_step=1e20, f_dec_coeff=1e-4, s_curv_coeff=0.9, cautious_factor=1e-6, machine_prec=1e-16)
    d.update(kw)
    return LbfgsParams(**d)


def test_reference_sources_are_the_reference():
    assert b"/root/reference" in R.lib().ref_sources()


@pytest.mark.parametrize("n,p,q,m", [(6, 6, 6, 9), (18, 6, 6, 9), (60, 6, 6, 9), (384, 6, 6, 9), (40, 2, 3, 1), (25, 4, 1, 5)])
def test_banded_system_bit_exact(oracle, n, p, q, m):
    """a4: factorizeLU / solve / solveAdj (banded_system.hpp:66-145), incl. exact zeros inside the band."""
    rng = np.random.default_rng(n * 31 + p)
    A = np.zeros((n, n))
    for i in range(n):
        for j in range(max(0, i - p), min(n, i + q + 1)):
            A[i, j] = rng.normal() if rng.random() > 0.25 else 0.0   # the `!= 0.0` skips matter
        A[i, i] = 6.0 + rng.random()
    B = rng.normal(size=(n, m))
    for adj in (False, True):
        ref, orc = R.banded_solve(A, p, q, B, adj), oracle.banded_solve(A, p, q, B, adj)
        assert np.array_equal(ref, orc)
        assert np.abs((A.T if adj else A) @ ref - B).max() < 1e-9


@pytest.mark.parametrize("N", [1, 2, 3, 8, 17, 64])
def test_minjerk_generate_and_backprop_bit_exact(oracle, N):
    """a3, a5, a11: MinJerkOpt<9>::generate, getTrajJerkCost, calJerkGradCT, calGradCTtoQT (minco.hpp:824-1069)."""
    rng = np.random.default_rng(100 + N)
    ew = np.array([0.33] + [1.0] * 8)
    head, tail = rng.normal(size=(9, 3)), rng.normal(size=(9, 3))
    inner, T = rng.normal(size=(max(N - 1, 0), 9)), rng.uniform(0.3, 2.5, N)
    ref, orc = R.minco_generate(N, ew, head, tail, inner, T), oracle.minco_generate(N, ew, head, tail, inner, T)
    for a, b, name in zip(ref, orc, ("coeff", "jerk", "gdC_jerk", "gdT_jerk")):
        assert np.array_equal(np.asarray(a), np.asarray(b)), name
    gdC, gdT = rng.normal(size=(6 * N, 9)), rng.normal(size=N)
    ref = R.minco_backprop(N, ew, head, tail, inner, T, gdC, gdT)
    orc = oracle.minco_backprop(N, ew, head, tail, inner, T, gdC, gdT)
    for a, b, name in zip(ref, orc, ("gdP", "gdTail", "gdT")):
        assert np.array_equal(a, b), name


def test_robot_constants_are_the_reference_constructor(oracle):
    """a10 + the constants of moma_param.h:36-144 as its constructor evaluates them."""
    rp, k = oracle.robot_defaults(), R.robot_constants()
    assert np.array_equal(np.array(rp.colli_length), k["colli_length"])
    assert np.array_equal(np.array(rp.colli_points), k["colli_points"])
    assert np.array_equal(np.array(rp.colli_point_radius), k["colli_radius"])
    assert np.array_equal(np.array(rp.joint_pos_limit_max), k["joint_pos_max"])
    assert np.array_equal(np.array(rp.relative_t), k["relative_t"])
    assert np.array_equal(np.array(rp.relative_R).reshape(3, 3), k["relative_R"])
    assert np.array_equal(np.array(rp.collision_matrix).reshape(12, 12), k["collision_matrix"])
    assert [rp.max_v, rp.max_a, rp.max_w, rp.max_dw, rp.chassis_colli_radius, rp.chassis_height] == list(k["limits"])
    # the product's defaults are the same table
    import topay_b200 as tp
    pp = tp.robot_params_default()
    assert bytes(pp) == bytes(rp)


def test_fk_and_adjoint_bit_exact(oracle):
    """a8, a9: getColliPts / getColliGrads (moma_param.h:203-337) on random whole-body states."""
    rng = np.random.default_rng(5)
    rp = oracle.robot_defaults()
    for _ in range(200):
        pos = np.concatenate([rng.uniform(-8, 8, 2), rng.uniform(-np.pi, np.pi, 1), rng.uniform(-3, 3, 7)])
        a, b = R.colli_pts(pos), oracle.colli_pts(rp, pos)
        assert a.shape == (12, 4) and np.array_equal(a, b)
        g = rng.normal(size=(12, 3)) * (rng.random((12, 1)) > 0.3)
        assert np.array_equal(R.colli_grads(pos, g), oracle.colli_grads(rp, pos, g))


@pytest.mark.parametrize("case", ["rosenbrock", "ring_wrap", "stage1_like", "stage2_like", "max_iter", "ls_fail", "g_eps"])
def test_lbfgs_bit_exact(oracle, case):
    """a13, a14: lbfgs_optimize + line_search_lewisoverton (lbfgs.hpp:276-722): same return code, iterate,
    cost and evaluation count on chained-Rosenbrock problems, incl. a history ring that wraps many times
    (mem_size 4), the reference's early-accept (:327-330, past > 0), max_iterations and line-search failure."""
    rng = np.random.default_rng(sum(map(ord, case)))
    n, b = 40, 1.0
    kw = {}
    if case == "ring_wrap":
        n, kw = 120, dict(mem_size=4, past=0, g_epsilon=1e-9)
    elif case == "stage1_like":
        kw = dict(mem_size=256, g_epsilon=0.0, past=2, delta=1e-2, min_step=0.0, max_iterations=8000)
    elif case == "stage2_like":
        n, b, kw = 200, 5.0, dict(mem_size=256, g_epsilon=0.0, past=3, delta=1e-4, min_step=1e-32, max_iterations=8000)
    elif case == "max_iter":
        kw = dict(max_iterations=7, past=0)
    elif case == "ls_fail":
        b, kw = 50.0, dict(max_linesearch=2, past=0)
    elif case == "g_eps":
        kw = dict(g_epsilon=1e-3, past=0)
    p = _lp(**kw)
    a, c, x0 = rng.uniform(0.5, 2.0, n), rng.normal(size=n), rng.normal(size=n)
    if case == "ls_fail":
        x0 = 3.0 * x0
    r_ref, x_ref, f_ref, it_ref, ev_ref, trace = R.lbfgs_test_problem(a, c, b, p, x0, trace_cap=10000)
    r_orc, x_orc, f_orc, it_orc, ev_orc = oracle.lbfgs_test_problem(a, c, b, p, x0)
    assert r_ref == r_orc and ev_ref == ev_orc
    assert np.array_equal(x_ref, x_orc) and f_ref == f_orc
    if case == "ring_wrap":
        assert len(trace) > 6 * p.mem_size      # the ring really wrapped
    if case == "max_iter":
        assert r_ref == -997 or len(trace) == 7   # LBFGSERR_MAXIMUMITERATION
    if case == "ls_fail":
        assert r_ref == -1009 and len(trace) > 5   # LBFGSERR_MAXIMUMLINESEARCH after some accepted iterations
