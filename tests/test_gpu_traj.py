"""GPU parity of the result post-processing and the worker's success gate (SURVEY.md §8 row a16 + "next" row
N1) through the C ABI against oracle/oracle_traj.hpp: pose table, getState/getDState, checkFeasible /
printConstraintsSituations verdicts and metrics on solved trajectories, synthetic limit cases, and the
solver-resident gate + shortest-duration selection. Tolerance 1e-10 relative on the metrics (device sin/cos are
<= 2 ulp from glibc's, FMA contraction); verdicts, sample counts and table lengths exact."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-10


def _close(a, b, tol=TOL):
    a, b = np.asarray(a), np.asarray(b)
    return np.all(np.abs(a - b) <= tol * np.maximum(1.0, np.abs(b)))


@pytest.fixture(scope="module")
def solved(small_scene):
    import topay_b200 as tp
    gm = tp.GridMap(small_scene["desc"], device=0)
    gm.regenerateMap(small_scene["points"])
    z = np.load(os.path.join(G, "solve_cases.npz"))
    n = int(z["n"])
    paths = [z[f"s{c}_path"] for c in range(n)]
    solver = tp.MomaTrajOpt(gm, max_cand=n, max_pieces=16)
    res = solver.optimizeTrajBatch(paths, np.zeros((n, 10, 2)), np.zeros((n, 10, 2)))
    trajs = [solver.getTraj(c) for c in range(n)]
    return dict(gm=gm, solver=solver, res=res, trajs=trajs, n=n)


def test_pose_table_and_state_samplers(solved, oracle):
    for tr in solved["trajs"]:
        seq = tr.car_seq
        exp = oracle.traj_car_seq([tr._tuple()])[0]
        assert seq.shape == exp.shape
        assert np.array_equal(seq[:, 3], exp[:, 3]) and _close(seq[:, :3], exp[:, :3])
        total = tr.getTotalDuration()
        t = np.concatenate([np.linspace(-0.5, total + 0.5, 257), np.cumsum(tr.durations)])
        st, ds = tr.getState(t), tr.getDState(t)
        est, eds = oracle.traj_sample([tr._tuple()], t[None])
        assert _close(st, est[0]) and _close(ds, eds[0])
        # a scalar time keeps the reference's call shape
        assert _close(tr.getState(0.37 * total), est[0][0] * 0 + oracle.traj_sample([tr._tuple()], [[0.37 * total]])[0][0, 0])
        pts = tr.sampleTimePoints(20)
        assert pts.shape == (20, 12) and np.all(np.abs(pts[:, 2]) <= np.pi)


def _assert_same(got, exp):
    assert np.array_equal(got["n_samples"], exp["n_samples"])
    assert np.array_equal(got["feasible"], exp["feasible"])
    assert np.array_equal(got["feasible_print"], exp["feasible_print"])
    for k in ("max_vel", "max_acc", "max_domega", "max_d2omega", "max_q", "max_dq", "max_d2q", "min_dist",
              "min_dist_mani"):
        assert _close(got[k], exp[k]), k


def test_check_feasible_on_solved_trajectories(solved, small_scene, oracle):
    import topay_b200 as tp
    solver, trajs = solved["solver"], solved["trajs"]
    rp = tp.robot_params_default()
    verdict = solver.checkFeasible(trajs)
    exp = oracle.check_feasible(small_scene["field"], rp, [t._tuple() for t in trajs])
    _assert_same(solver.constraints, exp)
    assert np.array_equal(verdict, exp["feasible"].astype(bool))
    assert np.array_equal(solver.printConstraintsSituations(trajs), exp["feasible_print"].astype(bool))
    assert solver.checkFeasible(trajs[0]) == bool(exp["feasible"][0])
    # some of the fixture's solved candidates pass the gate, and the verdicts differ between candidates
    assert exp["feasible_print"].any()


def test_solver_resident_gate_matches_host_batch_path(solved, small_scene, oracle):
    import topay_b200 as tp
    solver, res, n = solved["solver"], solved["res"], solved["n"]
    arrs, best = solver.checkFeasibleBatch()
    exp = oracle.check_feasible(small_scene["field"], tp.robot_params_default(),
                                [t._tuple() for t in solved["trajs"]])
    _assert_same(arrs, exp)
    ok = (res["status"] == 1) & (exp["feasible_print"] == 1)
    assert best == oracle.select_shortest(ok.astype(np.int32), res["duration"])
    assert best == solver._l.topay_select_shortest(
        np.ascontiguousarray(ok, dtype=np.int32).ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_int32)),
        np.ascontiguousarray(res["duration"]).ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_double)), n)


def test_limit_cases_and_signed_extremes(solved, small_scene, oracle):
    """Scaled copies of a solved trajectory: time-compressed (limits violated), mirrored joints (negative
    extremes), one very short and one 120 s long (sample clock, table growth)."""
    import topay_b200 as tp
    from topay_b200.optimizer import MomaTraj
    base = solved["trajs"][0]
    rp = tp.robot_params_default()

    def rescale(tr, s):
        # t -> t / s: coefficient of t^k scales by s^k, durations by 1/s
        N = len(tr.durations)
        c = tr.coeff.reshape(N, 6, 9) * (s ** np.arange(6))[None, :, None]
        return MomaTraj(tr.durations / s, c.reshape(6 * N, 9), tr.start_se2)

    fast, slow = rescale(base, 3.0), rescale(base, base.getTotalDuration() / 120.0)
    mirror = MomaTraj(base.durations, base.coeff * np.array([1, 1] + [-1] * 7)[None, :], base.start_se2)
    tiny = rescale(base, base.getTotalDuration() / 0.004)
    cases = [base, fast, slow, mirror, tiny]
    solver = solved["solver"]
    solver.checkFeasible(cases)
    exp = oracle.check_feasible(small_scene["field"], rp, [t._tuple() for t in cases])
    _assert_same(solver.constraints, exp)
    assert exp["feasible"][1] == 0 and exp["n_samples"][4] == 1 and exp["n_samples"][2] > 11000
    assert np.allclose(exp["max_q"][3], -exp["max_q"][0])


def test_traj_error_paths(solved):
    import ctypes as C
    import topay_b200 as tp
    from topay_b200 import _lib
    from topay_b200._structs import alloc_feasibility, pack_trajs
    from topay_b200.optimizer import MomaTraj
    base = solved["trajs"][0]
    # a field that was never rebuilt
    gm = tp.GridMap(tp.grid_desc(map_size=(4.0, 4.0, 1.0)))
    solver = tp.MomaTrajOpt(gm, max_cand=1, max_pieces=8)
    with pytest.raises(_lib.TopayError) as e:
        solver.checkFeasible(base)
    assert e.value.code == _lib.ERR_NOT_READY
    # piece_num outside 1..max_pieces
    tb, keep = pack_trajs([base._tuple()])
    keep[0][0] = 0
    f, arrs = alloc_feasibility(1)
    rc = _lib.lib().topay_traj_check_feasible(solved["gm"].h, C.byref(tp.robot_params_default()), C.byref(tb), C.byref(f))
    assert rc == _lib.ERR_INVALID_ARG
    # a trajectory longer than the sample clock table (1310 s)
    N = len(base.durations)
    slow = MomaTraj(base.durations * 1e3, base.coeff, base.start_se2)
    with pytest.raises(_lib.TopayError) as e:
        solved["solver"].checkFeasible(slow)
    assert e.value.code == _lib.ERR_TOO_LARGE
    # pose table larger than the caller's buffer
    tb, keep = pack_trajs([base._tuple()])
    out = np.zeros((1, 2, 4))
    ln = np.zeros(1, dtype=np.int32)
    rc = _lib.lib().topay_traj_car_seq(0, C.byref(tb), 2, out.ctypes.data_as(C.POINTER(C.c_double)),
                                       ln.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == _lib.ERR_TOO_LARGE and ln[0] > 2
