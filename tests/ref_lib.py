"""ctypes binding of oracle/_ref/libtopay_ref.so — the REFERENCE'S OWN headers (lbfgs.hpp,
banded_system.hpp, minco.hpp's MinJerkOpt<9>, moma_param.h's FK) compiled unmodified from
/root/reference against the Eigen / ROS stand-ins of oracle/ref_stubs (oracle/Makefile, target _ref).

TEST INFRASTRUCTURE ONLY. Where /root/reference exists (the build container) the library is (re)built
on demand; elsewhere (the GPU box) the prebuilt file that travelled with the snapshot is used, and
`available()` is False when there is neither.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from topay_b200._structs import LbfgsParams

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "_ref", "libtopay_ref.so")
REFERENCE = os.environ.get("TOPAY_REFERENCE", "/root/reference")
_lib = None


def available():
    return os.path.exists(_SO) or os.path.isdir(os.path.join(REFERENCE, "src", "planner"))


def lib():
    global _lib
    if _lib is None:
        if os.path.isdir(os.path.join(REFERENCE, "src", "planner")):
            subprocess.check_call(["make", "-s", "-C", os.path.join(_ROOT, "oracle"), "_ref", f"REFERENCE={REFERENCE}"])
        _lib = C.CDLL(_SO)
        _lib.ref_sources.restype = C.c_char_p
    return _lib


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def lbfgs_test_problem(a, c, b, params: LbfgsParams, x0, trace_cap=0, cancel_after=0):
    a, c = _f64(a), _f64(c)
    x = _f64(x0).copy()
    f, it, ev, tl = C.c_double(), C.c_int(), C.c_int(), C.c_int()
    tr = np.zeros((max(trace_cap, 1), 4))
    r = lib().ref_lbfgs_test_problem(len(x), _p(a), _p(c), C.c_double(b), C.byref(params), _p(x), C.byref(f),
                                     C.byref(it), C.byref(ev), _p(tr), trace_cap, C.byref(tl), cancel_after)
    return r, x, f.value, it.value, ev.value, tr[:tl.value]


def banded_solve(dense, p, q, B, adjoint=False):
    dense = _f64(dense)
    B = _f64(B).copy()
    n, m = B.shape
    lib().ref_banded_solve(n, p, q, _p(dense), m, _p(B), int(adjoint))
    return B


def minco_generate(N, ew, head, tail, inner, T):
    ew, head, tail, inner, T = map(_f64, (ew, head, tail, inner, T))
    coeff = np.zeros((6 * N, 9))
    jerk = C.c_double()
    gdC, gdT = np.zeros((6 * N, 9)), np.zeros(N)
    lib().ref_minco_generate(N, _p(ew), _p(head), _p(tail), _p(inner), _p(T), _p(coeff), C.byref(jerk), _p(gdC), _p(gdT))
    return coeff, jerk.value, gdC, gdT


def minco_backprop(N, ew, head, tail, inner, T, gdC, gdT):
    ew, head, tail, inner, T, gdC = map(_f64, (ew, head, tail, inner, T, gdC))
    gdT = _f64(gdT).copy()
    gdP, gdTail = np.zeros((max(N - 1, 0), 9)), np.zeros((9, 3))
    lib().ref_minco_backprop(N, _p(ew), _p(head), _p(tail), _p(inner), _p(T), _p(gdC), _p(gdT), _p(gdP), _p(gdTail))
    return gdP, gdTail, gdT


def colli_pts(pos10):
    pos10 = _f64(pos10)
    out = np.zeros((12, 4))
    n = lib().ref_colli_pts(_p(pos10), _p(out))
    return out[:n]


def colli_grads(pos10, grads):
    pos10, grads = _f64(pos10), _f64(grads)
    out = np.zeros(10)
    lib().ref_colli_grads(_p(pos10), _p(grads), len(grads), _p(out))
    return out


def robot_constants():
    cl, cp, cr = np.zeros(8), np.zeros(16), np.zeros(16)
    lm = np.zeros(12, dtype=np.int32)
    lim, jp, rt, rR = np.zeros(6), np.zeros(7), np.zeros(3), np.zeros(9)
    cm = np.zeros(144, dtype=np.int32)
    lib().ref_robot_constants(_p(cl), _p(cp), _p(cr), _p(lm, C.c_int), _p(lim), _p(jp), _p(rt), _p(rR), _p(cm, C.c_int))
    return dict(colli_length=cl, colli_points=cp, colli_radius=cr, link_map=lm, limits=lim, joint_pos_max=jp,
                relative_t=rt, relative_R=rR.reshape(3, 3), collision_matrix=cm.reshape(12, 12))


# ---------------------------------------------------------------------------------------------------------------
# The reference's GridMap and MomaTrajOpt themselves (oracle/ref_driver_full.cpp)
class RefSolveOut(C.Structure):
    _fields_ = [("status", C.c_int32), ("piece_num", C.c_int32), ("cost", C.c_double), ("duration", C.c_double),
                ("final_xy_err", C.c_double * 2)]


class GridMap:
    """nmoma_planner::GridMap (src/map), `use_rog: false`."""

    def __init__(self, desc):
        l = lib()
        l.ref_grid_create.restype = C.c_void_p
        self.desc = desc
        self.h = C.c_void_p(l.ref_grid_create(C.byref(desc)))
        d = (C.c_int32 * 3)()
        l.ref_grid_dims(self.h, d)
        self.dims = tuple(d)

    def set_cloud(self, xyz, clear=1):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        lib().ref_grid_set_cloud(self.h, _p(xyz, C.c_float), C.c_int64(len(xyz)), clear)

    def load_map(self, occ2d, occ3d):
        occ2d, occ3d = np.ascontiguousarray(occ2d, dtype=np.int8), np.ascontiguousarray(occ3d, dtype=np.int8)
        lib().ref_grid_load_map(self.h, _p(occ2d, C.c_int8), _p(occ3d, C.c_int8))

    def download(self, which):
        nx, ny, nz = self.dims
        out = np.zeros((nx, ny, nz) if which == 3 else (nx, ny))
        lib().ref_grid_download(self.h, which, _p(out))
        return out

    def download_occupancy(self, which):
        nx, ny, nz = self.dims
        out = np.zeros((nx, ny, nz) if which == 3 else (nx, ny), dtype=np.int8)
        lib().ref_grid_download_occupancy(self.h, which, _p(out, C.c_int8))
        return out

    def query3d(self, pos):
        pos = _f64(pos)
        d, g = np.zeros(len(pos)), np.zeros((len(pos), 3))
        lib().ref_grid_query3d(self.h, _p(pos), C.c_int64(len(pos)), _p(d), _p(g))
        return d, g

    def query2d(self, pos, which=0):
        pos = _f64(pos)
        d, g = np.zeros(len(pos)), np.zeros((len(pos), 2))
        lib().ref_grid_query2d(self.h, _p(pos), C.c_int64(len(pos)), which, _p(d), _p(g))
        return d, g

    def distance3d(self, pos):
        pos = _f64(pos)
        d = np.zeros(len(pos))
        lib().ref_grid_distance3d(self.h, _p(pos), C.c_int64(len(pos)), _p(d))
        return d

    def distance2d(self, pos):
        pos = _f64(pos)
        d = np.zeros(len(pos))
        lib().ref_grid_distance2d(self.h, _p(pos), C.c_int64(len(pos)), _p(d))
        return d

    def whole_body_collision(self, states):
        states = _f64(states)
        out = np.zeros(len(states), dtype=np.int8)
        lib().ref_grid_whole_body_collision(self.h, _p(states), C.c_int64(len(states)), _p(out, C.c_int8))
        return out.astype(bool)


def dense_path(grid, raw_xy, step_size, start_yaw, end_yaw, v_max, w_max):
    """GraphSearch::getDensePath of the compiled reference: rows (x, y, theta, dt)."""
    raw = _f64(raw_xy)
    cap = 16 * len(raw) + int(np.abs(np.diff(raw, axis=0)).sum() / step_size * 4) + 64
    out = np.zeros((cap, 4))
    n = lib().ref_dense_path(grid.h, _p(raw), len(raw), C.c_double(step_size), C.c_double(start_yaw), C.c_double(end_yaw),
                             C.c_double(v_max), C.c_double(w_max), _p(out), cap)
    assert n <= cap
    return out[:n]


def line_visib(grid, p1, p2, thresh, use_critical=False):
    """TopologyPRM::lineVisib of the compiled reference on n segments: (visible, pc); pc = nan where visible."""
    p1, p2 = _f64(p1), _f64(p2)
    n = len(p1)
    vis, pc = np.zeros(n, dtype=np.int8), np.full((n, 3), np.nan)
    lib().ref_line_visib(grid.h, _p(p1), _p(p2), C.c_int64(n), C.c_double(thresh), int(use_critical), _p(vis, C.c_int8), _p(pc))
    return vis.astype(bool), pc


def same_topo_path(grid, p1, p2, thresh, use_critical=False):
    """TopologyPRM::sameTopoPath of the compiled reference."""
    p1, p2 = _f64(p1), _f64(p2)
    return bool(lib().ref_same_topo_path(grid.h, _p(p1), len(p1), _p(p2), len(p2), C.c_double(thresh), int(use_critical)))


def discretize_path(grid, path, pt_num):
    path = _f64(path)
    out = np.zeros((pt_num, 3))
    n = lib().ref_discretize_path(grid.h, _p(path), len(path), int(pt_num), _p(out))
    assert n == pt_num
    return out


def path_length(grid, path):
    path = _f64(path)
    lib().ref_path_length.restype = C.c_double
    return lib().ref_path_length(grid.h, _p(path), len(path))


class MomaTrajOpt:
    """nmoma_planner::MomaTrajOpt (src/planner), parameters taken from a topay_opt_params."""

    def __init__(self, grid: GridMap, opt):
        l = lib()
        l.ref_opt_create.restype = C.c_void_p
        self.grid, self.opt = grid, opt
        self.h = C.c_void_p(l.ref_opt_create(grid.h, C.byref(opt)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_opt_destroy(self.h)
            self.h = None

    def eval_one(self, stage, N, head, tail, sxy, exy, inner_xy, lam, rho, x):
        from topay_b200._structs import NTERMS, num_vars
        head, tail, sxy, exy, inner_xy, x, lam, rho = map(_f64, (head, tail, sxy, exy, inner_xy, x, lam, rho))
        n = num_vars(N)
        cost = C.c_double()
        grad, terms, coeff, fxy = np.zeros(n), np.zeros(NTERMS), np.zeros((6 * N, 9)), np.zeros(2)
        lib().ref_opt_eval(self.h, stage, N, _p(head), _p(tail), _p(sxy), _p(exy), _p(inner_xy), _p(lam), _p(rho),
                           _p(x), C.byref(cost), _p(grad), _p(terms), _p(coeff), _p(fxy))
        return cost.value, grad, terms, coeff, fxy

    def solve_one(self, init_path, bvel, bacc, alm_max_rounds=20, wall_clock=False, max_pieces=64):
        init_path, bvel, bacc = _f64(init_path), _f64(bvel), _f64(bacc)
        out = RefSolveOut()
        T, coeff = np.zeros(max_pieces), np.zeros((6 * max_pieces, 9))
        lib().ref_opt_solve(self.h, _p(init_path), init_path.shape[0], _p(bvel), _p(bacc), alm_max_rounds,
                            int(wall_clock), C.byref(out), _p(T), _p(coeff))
        N = out.piece_num
        return dict(status=out.status, piece_num=N, cost=out.cost, duration=out.duration,
                    final_xy_err=np.array(out.final_xy_err[:]), T=T[:N].copy(), coeff=coeff[:6 * N].copy())

    def gate(self):
        a, b, d = C.c_int32(), C.c_int32(), C.c_double()
        lib().ref_opt_gate(self.h, C.byref(a), C.byref(b), C.byref(d))
        return bool(a.value), bool(b.value), d.value


def solve_batch(grid: GridMap, opt, paths, bvel, bacc, n_threads, alm_max_rounds=20, wall_clock=False):
    n = len(paths)
    plen = np.array([p.shape[0] for p in paths], dtype=np.int32)
    flat = _f64(np.concatenate(paths, axis=0))
    bvel, bacc = _f64(bvel), _f64(bacc)
    outs = (RefSolveOut * n)()
    lib().ref_solve_batch(grid.h, C.byref(opt), n, _p(plen, C.c_int32), _p(flat), _p(bvel), _p(bacc), alm_max_rounds,
                          int(wall_clock), n_threads, outs)
    return [dict(status=o.status, piece_num=o.piece_num, cost=o.cost, duration=o.duration) for o in outs]


class RogESDFMap:
    """rog_map::ESDFMap itself (oracle/ref_driver_rog.cpp); same surface as oracle_lib.RogField."""

    def __init__(self, desc):
        l = lib()
        l.ref_rog_create.restype = C.c_void_p
        self.h = C.c_void_p(l.ref_rog_create(C.byref(desc)))
        g = self._geometry()
        self.half, self.size, self.half_box, self.resolution = g[0], g[1], g[3], g[4]

    def _geometry(self):
        half, size, org, hb = ((C.c_int32 * 3)() for _ in range(4))
        res = C.c_double()
        lib().ref_rog_geometry(self.h, half, size, C.byref(res), org, hb)
        return tuple(half), tuple(size), tuple(org), tuple(hb), res.value

    @property
    def origin_i(self):
        return self._geometry()[2]

    def slide(self, odom):
        o = _f64(odom)
        lib().ref_rog_slide(self.h, _p(o))

    def update_counters(self, pos, from_type, to_type):
        pos = _f64(pos)
        a = np.ascontiguousarray(from_type, dtype=np.uint8)
        b = np.ascontiguousarray(to_type, dtype=np.uint8)
        lib().ref_rog_update_counters(self.h, _p(pos), _p(a, C.c_uint8), _p(b, C.c_uint8), C.c_int64(pos.shape[0]))

    def set_occupied_cnt(self, cnt):
        cnt = np.ascontiguousarray(cnt, dtype=np.int16)
        lib().ref_rog_set_occupied_cnt(self.h, _p(cnt, C.c_int16))

    def download_counters(self):
        a, b = np.empty(self.size, dtype=np.int16), np.empty(self.size, dtype=np.int16)
        lib().ref_rog_download_counters(self.h, _p(a, C.c_int16), _p(b, C.c_int16))
        return a, b

    def update_esdf(self, odom):
        o = _f64(odom)
        lib().ref_rog_update_esdf(self.h, _p(o))

    def query(self, kind, pos):
        pos = _f64(pos)
        n = pos.shape[0]
        d, g = np.empty(n), np.zeros((n, 3))
        lib().ref_rog_query(self.h, kind, _p(pos), C.c_int64(n), _p(d), _p(g))
        return d, g

    def evaluate_edt(self, pos):
        pos = _f64(pos)
        d = np.empty(pos.shape[0])
        lib().ref_rog_evaluate_edt(self.h, _p(pos), C.c_int64(pos.shape[0]), _p(d), None)
        return d

    def is_line_free2d(self, start, end, threshold=0.0):
        s, e = _f64(start), _f64(end)
        out = np.empty(s.shape[0], dtype=np.int8)
        lib().ref_rog_is_line_free2d(self.h, _p(s), _p(e), C.c_int64(s.shape[0]), C.c_double(threshold), _p(out, C.c_int8))
        return out

    def download(self, which):
        out = np.empty(self.size if which < 2 else self.size[:2])
        lib().ref_rog_download(self.h, which, _p(out))
        return out


class RogProbMap:
    """rog_map::ProbMap itself with its own ESDFMap (oracle/ref_driver_rog.cpp). `esdf` exposes the ESDF layer
    through the RogESDFMap surface. The reference clears the first-frame sphere ONCE PER PROCESS (a function-level
    static in updateProbMap): `first_frame_pending()` tells whether that has happened yet in this process."""
    _first_done = False

    def __init__(self, rog_desc, prob_desc):
        l = lib()
        l.ref_prob_create.restype = C.c_void_p
        l.ref_prob_esdf.restype = C.c_void_p
        self.h = C.c_void_p(l.ref_prob_create(C.byref(rog_desc), C.byref(prob_desc)))
        sz = (C.c_int32 * 3)()
        l.ref_prob_size(self.h, sz)
        self.size = tuple(sz)
        self.esdf = RogESDFMap.__new__(RogESDFMap)
        self.esdf.h = C.c_void_p(l.ref_prob_esdf(self.h))
        g = self.esdf._geometry()
        self.esdf.half, self.esdf.size, self.esdf.half_box, self.esdf.resolution = g[0], g[1], g[3], g[4]

    @classmethod
    def first_frame_pending(cls):
        return not cls._first_done

    def update(self, cloud_xyzi, pos):
        c = np.ascontiguousarray(cloud_xyzi, dtype=np.float32).reshape(-1, 4)
        p = _f64(pos)
        lib().ref_prob_update(self.h, _p(c, C.c_float), C.c_int64(c.shape[0]), _p(p))

    def download(self):
        occ = np.empty(self.size, dtype=np.float32)
        org = (C.c_int32 * 3)()
        lib().ref_prob_download(self.h, _p(occ, C.c_float), org)
        return occ, tuple(org)
