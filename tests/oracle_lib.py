"""ctypes binding of oracle/liboracle.so — the CPU restatement of the reference hot path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py. The product package (topay_b200/) never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from topay_b200._structs import (GridDesc, RogDesc, alloc_feasibility, pack_trajs, LbfgsParams, OptParams, RobotParams, NTERMS, num_vars)

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "liboracle.so")


def build(force=False):
    srcs = [os.path.join(_ROOT, "oracle", f) for f in
            ("oracle_capi.cpp", "oracle_field.hpp", "oracle_robot.hpp", "oracle_rog.hpp", "oracle_solve.hpp", "oracle_traj.hpp")]
    srcs.append(os.path.join(_ROOT, "include", "topay_b200.h"))
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle"), "liboracle.so"])
    return _SO


class SolveOut(C.Structure):
    _fields_ = [("status", C.c_int32), ("lbfgs_code", C.c_int32), ("piece_num", C.c_int32),
                ("iters", C.c_int32), ("evals", C.c_int32), ("alm_rounds", C.c_int32),
                ("cost", C.c_double), ("duration", C.c_double), ("final_xy_err", C.c_double * 2)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_field_create.restype = C.c_void_p
        _lib.oracle_rog_create.restype = C.c_void_p
    return _lib


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def use_rog(rog_field):
    """Route the solve / evaluation / gate of the oracle to a RogField (None = back to the dense Field)."""
    lib().oracle_use_rog(rog_field.h if rog_field is not None else None)


def robot_defaults():
    rp = RobotParams()
    lib().oracle_robot_params_default(C.byref(rp))
    return rp


def opt_defaults():
    o = OptParams()
    lib().oracle_opt_params_default(C.byref(o))
    return o


def dense_path(raw_xy, step_size, start_yaw, end_yaw, v_max, w_max):
    """graph_search.cpp:119-176: rows (x, y, theta, dt)."""
    raw = _f64(raw_xy)
    cap = 16 * len(raw) + int(np.abs(np.diff(raw, axis=0)).sum() / step_size * 4) + 64
    out = np.zeros((cap, 4))
    n = lib().oracle_dense_path(_p(raw), len(raw), C.c_double(step_size), C.c_double(start_yaw), C.c_double(end_yaw),
                                C.c_double(v_max), C.c_double(w_max), _p(out), cap)
    assert n <= cap
    return out[:n]


def discretize_path(path, pt_num):
    """topo_prm.cpp:472-506."""
    path = _f64(path)
    out = np.zeros((pt_num, 3))
    n = lib().oracle_discretize_path(_p(path), len(path), int(pt_num), _p(out))
    assert n == pt_num
    return out


def path_length(path):
    path = _f64(path)
    lib().oracle_path_length.restype = C.c_double
    return lib().oracle_path_length(_p(path), len(path))


class Field:
    def __init__(self, desc: GridDesc):
        self.h = C.c_void_p(lib().oracle_field_create(C.byref(desc)))
        d = (C.c_int32 * 3)()
        lib().oracle_field_dims(self.h, d)
        self.dims = tuple(d)

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_field_destroy(self.h)
            self.h = None

    def set_occupancy(self, occ3d=None, occ2d=None, occ2d_critical=None):
        a = [None if o is None else np.ascontiguousarray(o, dtype=np.int8) for o in (occ3d, occ2d, occ2d_critical)]
        lib().oracle_field_set_occupancy(self.h, *[_p(x, C.c_int8) for x in a])

    def clear(self, clear_critical=False):
        lib().oracle_field_clear(self.h, int(clear_critical))

    def rasterize(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        lib().oracle_field_rasterize(self.h, _p(xyz, C.c_float), C.c_int64(xyz.shape[0]))

    def rebuild(self):
        lib().oracle_field_rebuild(self.h)

    def same_topo_path(self, p1, p2, thresh, use_critical=False):
        """topo_prm.cpp:424-448."""
        p1, p2 = _f64(p1), _f64(p2)
        return bool(lib().oracle_same_topo_path(self.h, _p(p1), len(p1), _p(p2), len(p2), C.c_double(thresh), int(use_critical)))

    def line_visib(self, p1, p2, thresh, use_critical=False):
        """topo_prm.cpp:278-315 on n segments: (visible, pc); pc = nan where visible."""
        p1, p2 = _f64(p1), _f64(p2)
        n = len(p1)
        vis, pc = np.zeros(n, dtype=np.int8), np.full((n, 3), np.nan)
        lib().oracle_line_visib(self.h, _p(p1), _p(p2), C.c_int64(n), C.c_double(thresh), int(use_critical),
                                _p(vis, C.c_int8), _p(pc))
        return vis.astype(bool), pc

    def query3d(self, pos):
        pos = _f64(pos)
        n = pos.shape[0]
        d, g = np.empty(n), np.empty((n, 3))
        lib().oracle_field_query3d(self.h, _p(pos), C.c_int64(n), _p(d), _p(g))
        return d, g

    def query2d(self, pos, which=0):
        pos = _f64(pos)
        n = pos.shape[0]
        d, g = np.empty(n), np.empty((n, 2))
        lib().oracle_field_query2d(self.h, _p(pos), C.c_int64(n), which, _p(d), _p(g))
        return d, g

    def distance3d(self, pos):
        pos = _f64(pos)
        d = np.empty(pos.shape[0])
        lib().oracle_field_distance3d(self.h, _p(pos), C.c_int64(pos.shape[0]), _p(d))
        return d

    def distance2d(self, pos):
        pos = _f64(pos)
        d = np.empty(pos.shape[0])
        lib().oracle_field_distance2d(self.h, _p(pos), C.c_int64(pos.shape[0]), _p(d))
        return d

    def is_collision(self, pos, threshold):
        pos = _f64(pos)
        out = np.empty(pos.shape[0], dtype=np.int8)
        lib().oracle_field_is_collision(self.h, pos.shape[1], _p(pos), C.c_int64(pos.shape[0]), C.c_double(threshold),
                                        _p(out, C.c_int8))
        return out.astype(bool)

    def dist_coarse2d(self, pos, critical=False):
        pos = _f64(pos)
        out = np.empty(pos.shape[0])
        lib().oracle_field_dist_coarse2d(self.h, _p(pos), C.c_int64(pos.shape[0]), int(critical), _p(out))
        return out

    def dist_coarse2i(self, idx, critical=False):
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        out = np.empty(idx.shape[0])
        lib().oracle_field_dist_coarse2i(self.h, _p(idx, C.c_int32), C.c_int64(idx.shape[0]), int(critical), _p(out))
        return out

    def line_collision2d(self, p1, p2, threshold=0.0):
        p1, p2 = _f64(p1), _f64(p2)
        out = np.empty(p1.shape[0], dtype=np.int8)
        lib().oracle_field_line_collision2d(self.h, _p(p1), _p(p2), C.c_int64(p1.shape[0]), C.c_double(threshold),
                                            _p(out, C.c_int8))
        return out.astype(bool)

    def whole_body_collision(self, rp, states):
        states = _f64(states)
        out = np.empty(states.shape[0], dtype=np.int8)
        lib().oracle_field_whole_body_collision(self.h, C.byref(rp), _p(states), C.c_int64(states.shape[0]),
                                                _p(out, C.c_int8))
        return out

    def _shape(self, which):
        return self.dims if which == 3 else self.dims[:2]

    def download(self, which):
        out = np.empty(self._shape(which))
        lib().oracle_field_download(self.h, which, _p(out))
        return out

    def download_sqdist(self, which):
        a = np.empty(self._shape(which), dtype=np.int32)
        b = np.empty(self._shape(which), dtype=np.int32)
        lib().oracle_field_download_sqdist(self.h, which, _p(a, C.c_int32), _p(b, C.c_int32))
        return a, b

    def download_occupancy(self, which):
        a = np.empty(self._shape(which), dtype=np.int8)
        lib().oracle_field_download_occupancy(self.h, which, _p(a, C.c_int8))
        return a


def colli_pts(rp, pos10):
    pos10 = _f64(pos10)
    out = np.zeros((12, 4))
    n = lib().oracle_colli_pts(C.byref(rp), _p(pos10), _p(out))
    return out[:n]


def colli_grads(rp, pos10, grads):
    pos10, grads = _f64(pos10), _f64(grads)
    out = np.zeros(10)
    lib().oracle_colli_grads(C.byref(rp), _p(pos10), _p(grads), _p(out))
    return out


def minco_generate(N, ew, head, tail, inner, T):
    """inner: (N-1, 9) array (row i = inner point i). Returns coeff (6N,9), jerk, gdC, gdT."""
    ew, head, tail, inner, T = map(_f64, (ew, head, tail, inner, T))
    coeff = np.zeros((6 * N, 9))
    jerk = C.c_double()
    gdC, gdT = np.zeros((6 * N, 9)), np.zeros(N)
    lib().oracle_minco_generate(N, _p(ew), _p(head), _p(tail), _p(inner), _p(T), _p(coeff), C.byref(jerk),
                                _p(gdC), _p(gdT))
    return coeff, jerk.value, gdC, gdT


def minco_backprop(N, ew, head, tail, inner, T, gdC, gdT):
    ew, head, tail, inner, T, gdC = map(_f64, (ew, head, tail, inner, T, gdC))
    gdT = _f64(gdT).copy()
    gdP, gdTail = np.zeros((max(N - 1, 0), 9)), np.zeros((9, 3))
    lib().oracle_minco_backprop(N, _p(ew), _p(head), _p(tail), _p(inner), _p(T), _p(gdC), _p(gdT), _p(gdP),
                                _p(gdTail))
    return gdP, gdTail, gdT


def banded_solve(dense, p, q, B, adjoint=False):
    dense = _f64(dense)
    B = _f64(B).copy()
    n, m = B.shape
    lib().oracle_banded_solve(n, p, q, _p(dense), m, _p(B), int(adjoint))
    return B


def lbfgs_test_problem(a, c, b, params: LbfgsParams, x0):
    a, c = _f64(a), _f64(c)
    x = _f64(x0).copy()
    f, it, ev = C.c_double(), C.c_int(), C.c_int()
    r = lib().oracle_lbfgs_test_problem(len(x), _p(a), _p(c), C.c_double(b), C.byref(params), _p(x), C.byref(f),
                                        C.byref(it), C.byref(ev))
    return r, x, f.value, it.value, ev.value


def prepare_candidate(opt, rp, init_path, bvel, bacc, max_pieces):
    init_path, bvel, bacc = _f64(init_path), _f64(bvel), _f64(bacc)
    N = C.c_int32()
    past = C.c_int32()
    head, tail = np.zeros((9, 3)), np.zeros((9, 3))
    sxy, exy = np.zeros(2), np.zeros(2)
    inner_xy = np.zeros((max_pieces, 2))
    x0 = np.zeros(num_vars(max_pieces))
    rc = lib().oracle_prepare_candidate(C.byref(opt), C.byref(rp), _p(init_path), init_path.shape[0], _p(bvel),
                                        _p(bacc), max_pieces, C.byref(N), _p(head), _p(tail), _p(sxy), _p(exy),
                                        _p(inner_xy), _p(x0), C.byref(past))
    return dict(rc=rc, piece_num=N.value, head_pva=head, tail_pva=tail, start_xy=sxy, end_xy=exy,
                init_inner_xy=inner_xy, x0=x0[:num_vars(N.value)] if rc == 0 else x0, s1_past=past.value)


def eval_one(opt, rp, field: Field, stage, N, head, tail, sxy, exy, inner_xy, lam, rho, x):
    head, tail, sxy, exy, inner_xy, x = map(_f64, (head, tail, sxy, exy, inner_xy, x))
    lam = _f64(lam) if lam is not None else np.zeros(2)
    rho = _f64(rho) if rho is not None else np.ones(2)
    n = num_vars(N)
    cost = C.c_double()
    grad, terms = np.zeros(n), np.zeros(NTERMS)
    coeff, fxy = np.zeros((6 * N, 9)), np.zeros(2)
    lib().oracle_eval(C.byref(opt), C.byref(rp), field.h if field is not None else None, stage, N, _p(head), _p(tail), _p(sxy), _p(exy),
                      _p(inner_xy), _p(lam), _p(rho), _p(x), C.byref(cost), _p(grad), _p(terms), _p(coeff), _p(fxy))
    return cost.value, grad, terms, coeff, fxy


def solve_one(opt, rp, field: Field, init_path, bvel, bacc, max_pieces=64, wall_cap_s=0.0, trace=False):
    init_path, bvel, bacc = _f64(init_path), _f64(bvel), _f64(bacc)
    out = SolveOut()
    T = np.zeros(max_pieces)
    coeff = np.zeros((6 * max_pieces, 9))
    x = np.zeros(num_vars(max_pieces))
    cap = 4 * 200000
    tr = np.zeros(cap) if trace else None
    tl = C.c_int(0)
    lib().oracle_solve(C.byref(opt), C.byref(rp), field.h, _p(init_path), init_path.shape[0], _p(bvel), _p(bacc),
                       C.c_double(wall_cap_s), C.byref(out), _p(T), _p(coeff), _p(x), _p(tr), cap, C.byref(tl))
    N = out.piece_num
    res = dict(status=out.status, lbfgs_code=out.lbfgs_code, piece_num=N, iters=out.iters, evals=out.evals,
               alm_rounds=out.alm_rounds, cost=out.cost, duration=out.duration,
               final_xy_err=np.array(out.final_xy_err[:]), T=T[:N].copy(), coeff=coeff[:6 * N].copy(),
               x=x[:num_vars(N)].copy())
    if trace:
        res["trace"] = tr[:tl.value].reshape(-1, 4).copy()
    return res


def solve_batch(opt, rp, field: Field, paths, bvel, bacc, n_threads, wall_cap_s=0.0):
    """paths: list of (len_i, 10) arrays; bvel/bacc: (n, 10, 2). Thread-per-candidate like planner.cpp:921."""
    n = len(paths)
    plen = np.array([p.shape[0] for p in paths], dtype=np.int32)
    flat = _f64(np.concatenate(paths, axis=0))
    bvel, bacc = _f64(bvel), _f64(bacc)
    outs = (SolveOut * n)()
    lib().oracle_solve_batch(C.byref(opt), C.byref(rp), field.h, n, _p(plen, C.c_int32), _p(flat), _p(bvel),
                             _p(bacc), C.c_double(wall_cap_s), n_threads, outs)
    return [dict(status=o.status, lbfgs_code=o.lbfgs_code, piece_num=o.piece_num, iters=o.iters, evals=o.evals,
                 alm_rounds=o.alm_rounds, cost=o.cost, duration=o.duration) for o in outs]


def penalty_only(opt, rp, field: Field, stage, N, coeff, T, sxy, exy, inner_xy, lam, rho):
    coeff, T, sxy, exy, inner_xy, lam, rho = map(_f64, (coeff, T, sxy, exy, inner_xy, lam, rho))
    cost = C.c_double()
    gdC, gdT, terms, fxy = np.zeros((6 * N, 9)), np.zeros(N), np.zeros(NTERMS), np.zeros(2)
    lib().oracle_penalty_only(C.byref(opt), C.byref(rp), field.h, stage, N, _p(coeff), _p(T), _p(sxy), _p(exy),
                              _p(inner_xy), _p(lam), _p(rho), C.byref(cost), _p(gdC), _p(gdT), _p(terms), _p(fxy))
    return cost.value, gdC, gdT, terms, fxy


class RogField:
    """oracle_rog.hpp (the ROG-Map ESDFMap restatement)."""

    def __init__(self, desc: RogDesc):
        self.h = C.c_void_p(lib().oracle_rog_create(C.byref(desc)))
        half, size, org, hb = ((C.c_int32 * 3)() for _ in range(4))
        res = C.c_double()
        lib().oracle_rog_geometry(self.h, half, size, C.byref(res), org, hb)
        self.half, self.size, self.half_box = tuple(half), tuple(size), tuple(hb)
        self.resolution = res.value

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_rog_destroy(self.h)
            self.h = None

    @property
    def origin_i(self):
        half, size, org, hb = ((C.c_int32 * 3)() for _ in range(4))
        res = C.c_double()
        lib().oracle_rog_geometry(self.h, half, size, C.byref(res), org, hb)
        return tuple(org)

    def slide(self, odom):
        o = _f64(odom)
        lib().oracle_rog_slide(self.h, _p(o))

    def update_counters(self, pos, from_type, to_type):
        pos = _f64(pos)
        a = np.ascontiguousarray(from_type, dtype=np.uint8)
        b = np.ascontiguousarray(to_type, dtype=np.uint8)
        lib().oracle_rog_update_counters(self.h, _p(pos), _p(a, C.c_uint8), _p(b, C.c_uint8), C.c_int64(pos.shape[0]))

    def set_occupied_cnt(self, cnt):
        cnt = np.ascontiguousarray(cnt, dtype=np.int16)
        assert cnt.size == int(np.prod(self.size))
        lib().oracle_rog_set_occupied_cnt(self.h, _p(cnt, C.c_int16))

    def download_counters(self):
        a = np.empty(self.size, dtype=np.int16)
        b = np.empty(self.size, dtype=np.int16)
        lib().oracle_rog_download_counters(self.h, _p(a, C.c_int16), _p(b, C.c_int16))
        return a, b

    def update_esdf(self, odom):
        o = _f64(odom)
        lib().oracle_rog_update_esdf(self.h, _p(o))

    def query(self, kind, pos):
        pos = _f64(pos)
        n = pos.shape[0]
        d, g = np.empty(n), np.zeros((n, 3))
        lib().oracle_rog_query(self.h, kind, _p(pos), C.c_int64(n), _p(d), _p(g))
        return d, g

    def evaluate_edt(self, pos):
        pos = _f64(pos)
        d = np.empty(pos.shape[0])
        lib().oracle_rog_evaluate_edt(self.h, _p(pos), C.c_int64(pos.shape[0]), _p(d))
        return d

    def is_line_free2d(self, start, end, threshold=0.0):
        s, e = _f64(start), _f64(end)
        out = np.empty(s.shape[0], dtype=np.int8)
        lib().oracle_rog_is_line_free2d(self.h, _p(s), _p(e), C.c_int64(s.shape[0]), C.c_double(threshold),
                                        _p(out, C.c_int8))
        return out

    def download(self, which):
        out = np.empty(self.size if which < 2 else self.size[:2])
        lib().oracle_rog_download(self.h, which, _p(out))
        return out


# ---- MomaTraj post-processing + success gate (oracle_traj.hpp) -----------------------------------------
def traj_car_seq(trajs, cap=4096):
    tb, keep = pack_trajs(trajs)
    out = np.zeros((len(trajs), cap, 4))
    ln = np.zeros(len(trajs), dtype=np.int32)
    rc = lib().oracle_traj_car_seq(C.byref(tb), cap, _p(out), _p(ln, C.c_int32))
    assert rc == 0
    return [out[i, :ln[i]].copy() for i in range(len(trajs))]


def traj_sample(trajs, t):
    tb, keep = pack_trajs(trajs)
    t = _f64(t)
    n, m = t.shape
    st, ds = np.zeros((n, m, 10)), np.zeros((n, m, 10))
    lib().oracle_traj_sample(C.byref(tb), _p(t), m, _p(st), _p(ds))
    return st, ds


def check_feasible(field, rp, trajs):
    tb, keep = pack_trajs(trajs)
    f, arrs = alloc_feasibility(len(trajs))
    lib().oracle_check_feasible(field.h if field is not None else None, C.byref(rp), C.byref(tb), C.byref(f))
    return arrs


def select_shortest(succ, dur):
    s = np.ascontiguousarray(succ, dtype=np.int32)
    d = _f64(dur)
    return lib().oracle_select_shortest(_p(s, C.c_int32), _p(d), len(s))


class RogProb:
    """oracle_rog.hpp RogProb (the rog_map::ProbMap restatement) over a RogField."""

    def __init__(self, field: RogField, rog_desc, prob_desc):
        lib().oracle_prob_create.restype = C.c_void_p
        self.field = field
        self.h = C.c_void_p(lib().oracle_prob_create(field.h, C.byref(rog_desc), C.byref(prob_desc)))
        sz = (C.c_int32 * 3)()
        lib().oracle_prob_size(self.h, sz)
        self.size = tuple(sz)

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_prob_destroy(self.h)
            self.h = None

    def update(self, cloud_xyzi, pos):
        c = np.ascontiguousarray(cloud_xyzi, dtype=np.float32).reshape(-1, 4)
        p = _f64(pos)
        lib().oracle_prob_update(self.h, _p(c, C.c_float), C.c_int64(c.shape[0]), _p(p))

    def set_first_frame(self, armed):
        lib().oracle_prob_set_first_frame(self.h, int(armed))

    def download(self):
        occ = np.empty(self.size, dtype=np.float32)
        org = (C.c_int32 * 3)()
        lib().oracle_prob_download(self.h, _p(occ, C.c_float), org)
        return occ, tuple(org)
