"""Known-answer tests that pin oracle/oracle_traj.hpp (MomaTraj pose table, getState, checkFeasible,
selection — SURVEY.md §8 row a16 and "next" row N1): analytic arcs, hand-counted samples, limits."""
import numpy as np
import pytest

import oracle_lib as O
from topay_b200._structs import grid_desc


def poly_piece(theta=(0, 0, 0, 0, 0, 0), arc=(0, 0, 0, 0, 0, 0), joints=None):
    """coeff[6][9]: row k = coefficient of t^k; columns theta, arc, q1..q7."""
    c = np.zeros((6, 9))
    c[:, 0] = theta
    c[:, 1] = arc
    if joints is not None:
        c[:, 2:] = joints
    return c


def circle_traj(v=0.5, w=0.4, T=(2.0, 3.5, 1.7)):
    """theta = w t, s = v t over three pieces (each piece restarts its local time)."""
    durs, cs, t0 = [], [], 0.0
    for d in T:
        cs.append(poly_piece(theta=(w * t0, w, 0, 0, 0, 0), arc=(v * t0, v, 0, 0, 0, 0)))
        durs.append(d)
        t0 += d
    return (np.array(durs), np.concatenate(cs), np.array([0.0, 0.0, 0.0]))


def test_car_seq_follows_the_circle():
    v, w = 0.5, 0.4
    tr = circle_traj(v, w)
    seq = O.traj_car_seq([tr])[0]
    total = tr[0].sum()
    assert seq.shape[0] == 1 + int(np.floor(np.floor(total / 0.025) / 4))
    t = seq[:, 3]
    assert np.allclose(t, 0.1 * np.arange(len(t)), atol=1e-12)
    assert np.allclose(seq[:, 0], v / w * np.sin(w * t), atol=1e-9)
    assert np.allclose(seq[:, 1], v / w * (1 - np.cos(w * t)), atol=1e-9)
    assert np.allclose(seq[:, 2], w * t, atol=1e-12)


def test_get_state_and_dstate():
    v, w = 0.5, 0.4
    joints = np.zeros((6, 7))
    joints[0] = np.linspace(-0.3, 0.3, 7)
    joints[1] = 0.05 * np.arange(7)
    c = poly_piece(theta=(0.2, w, 0, 0, 0, 0), arc=(0, v, 0, 0, 0, 0), joints=joints)
    tr = (np.array([4.0]), c, np.array([1.0, -2.0, 0.2]))
    t = np.array([[0.0, 0.033, 0.1, 1.2345, 3.99, 4.0, 7.0, -1.0]])
    st, ds = O.traj_sample([tr], t)
    tc = np.clip(t[0], 0, 4.0)
    yaw = 0.2 + w * tc
    assert np.allclose(st[0, :, 2], yaw, atol=1e-12)
    assert np.allclose(st[0, :, 0], 1.0 + v / w * (np.sin(yaw) - np.sin(0.2)), atol=1e-9)
    assert np.allclose(st[0, :, 1], -2.0 - v / w * (np.cos(yaw) - np.cos(0.2)), atol=1e-9)
    assert np.allclose(st[0, :, 3:], joints[0] + joints[1] * tc[:, None], atol=1e-12)
    assert np.allclose(ds[0, :, 0], v) and np.allclose(ds[0, :, 1], w) and np.all(ds[0, :, 2] == 0)
    assert np.allclose(ds[0, :, 3:], joints[1])


def test_locate_extends_the_last_piece_and_switches_at_boundaries():
    # two pieces with different slopes; at t == T0 exactly the first piece is still used (t > dur is strict)
    c = np.concatenate([poly_piece(arc=(0, 1.0, 0, 0, 0, 0)), poly_piece(arc=(5.0, 2.0, 0, 0, 0, 0))])
    tr = (np.array([1.0, 1.0]), c, np.zeros(3))
    st, ds = O.traj_sample([tr], np.array([[1.0, 1.0000001, 2.0]]))
    assert ds[0, 0, 0] == 1.0 and ds[0, 1, 0] == 2.0 and ds[0, 2, 0] == 2.0


def _empty_field(occ_pts=None):
    desc = grid_desc(map_size=(10.0, 10.0, 1.6), resolution=0.1)
    f = O.Field(desc)
    f.clear(True)
    if occ_pts is not None:
        f.rasterize(np.asarray(occ_pts, dtype=np.float32))
    f.rebuild()
    return f


def straight(v, T=3.0, x0=-2.0, q=None):
    joints = np.zeros((6, 7))
    if q is not None:
        joints[0] = q
    return (np.array([T]), poly_piece(arc=(0, v, 0, 0, 0, 0), joints=joints), np.array([x0, 0.0, 0.0]))


def test_sample_count_uses_the_accumulated_clock():
    f = _empty_field()
    rp = O.robot_defaults()
    r = O.check_feasible(f, rp, [straight(0.3, T=1.0), straight(0.3, T=0.995), straight(0.3, T=2.5)])
    # t += 0.01 a hundred times gives 1.0000000000000007: exactly 100 samples fall below 1.0
    t, n = 0.0, []
    for T in (1.0, 0.995, 2.5):
        t, k = 0.0, 0
        while t < T:
            k += 1
            t += 0.01
        n.append(k)
    assert r["n_samples"].tolist() == n and n[0] == 100


def test_feasibility_limits():
    rp = O.robot_defaults()
    f = _empty_field()
    ok = straight(0.9 * rp.max_v)
    fast = straight(1.02 * rp.max_v)
    edge = straight(1.005 * rp.max_v)           # inside the 1 % allowance
    qbad = straight(0.3, q=[0, 0, 0, 1.02 * rp.joint_pos_limit_max[3], 0, 0, 0])
    r = O.check_feasible(f, rp, [ok, fast, edge, qbad])
    assert r["feasible"].tolist() == [1, 0, 1, 0]
    assert r["feasible_print"].tolist() == [1, 0, 1, 0]
    assert np.allclose(r["max_vel"], [0.9 * rp.max_v, 1.02 * rp.max_v, 1.005 * rp.max_v, 0.3])
    assert r["max_q"][3, 3] == pytest.approx(1.02 * rp.joint_pos_limit_max[3])
    # a backwards run reports the signed extreme
    back = straight(-0.4)
    r = O.check_feasible(f, rp, [back])
    assert r["max_vel"][0] == pytest.approx(-0.4) and r["feasible"][0] == 1


def test_feasibility_clearances():
    rp = O.robot_defaults()
    # a post next to the path at chassis height: the chassis clearance fails
    post = [[0.0, 0.31, z] for z in np.arange(0.02, 0.14, 0.02)]
    f = _empty_field(post)
    r = O.check_feasible(f, rp, [straight(0.3, T=12.0)])
    assert r["min_dist"][0] < 0.99 * rp.chassis_colli_radius
    assert r["feasible"][0] == 0 and r["feasible_print"][0] == 0
    # an obstacle only the arm touches (above the chassis, at the zero-pose arm): checkFeasible fails,
    # printConstraintsSituations reports it but keeps its verdict (moma_traj_opt.h:1199)
    pts = np.zeros((12, 4))
    n = O.lib().oracle_colli_pts(__import__("ctypes").byref(rp), O._p(np.array([0.0] * 10)), O._p(pts))
    top = pts[n - 1, :3]
    blob = [[top[0] + dx, top[1] + dy, top[2] + dz] for dx in (-0.1, 0, 0.1) for dy in (-0.1, 0, 0.1)
            for dz in (-0.1, 0, 0.1)]
    f = _empty_field(blob)
    r = O.check_feasible(f, rp, [straight(0.0001, T=0.5, x0=0.0)])
    assert r["min_dist_mani"][0].min() < 0.99 * pts[:n, 3].min() + 0.2
    if r["min_dist"][0] >= 0.99 * rp.chassis_colli_radius and (r["min_dist_mani"][0, :n] < 0.99 * pts[:n, 3]).any():
        assert r["feasible"][0] == 0 and r["feasible_print"][0] == 1
    else:
        pytest.fail("scene did not isolate the manipulator clearance: %s %s" % (r["min_dist"], r["min_dist_mani"]))


def test_select_shortest_keeps_the_first_of_equal_durations():
    assert O.select_shortest([0, 1, 1, 1], [1.0, 5.0, 3.0, 3.0]) == 2
    assert O.select_shortest([0, 0], [1.0, 2.0]) == -1
    assert O.select_shortest([1, 0, 1], [2.0, 1.0, 2.0]) == 0
