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
