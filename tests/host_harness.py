"""Builds tests/host_harness.cpp (CPU instantiation of the TP_HD device functions) into a scratch
library. TEST HARNESS ONLY — see the header of host_harness.cpp."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


def lib():
    global _lib
    if _lib is None:
        out = os.path.join(tempfile.gettempdir(), f"libtopay_hh_{os.getuid()}.so")
        src = os.path.join(_ROOT, "tests", "host_harness.cpp")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", out, src])
        _lib = C.CDLL(out)
    return _lib


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


class Harness:
    def __init__(self, opt, rp, desc, ofield):
        l = lib()
        self.l = l
        self.P = C.create_string_buffer(l.hh_params_size())
        l.hh_make_params(C.byref(opt), C.byref(rp), self.P)
        self.bufs = [np.ascontiguousarray(ofield.download(w)) for w in (3, 0, 1, 2)]
        self.G = C.create_string_buffer(l.hh_grid_size())
        l.hh_make_grid(C.byref(desc), *[_p(b) for b in self.bufs], self.G)

    def query3d(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        d, g = np.empty(len(pos)), np.empty((len(pos), 3))
        self.l.hh_query3d(self.G, _p(pos), C.c_int64(len(pos)), _p(d), _p(g))
        return d, g

    def query2d(self, pos, which=0):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        d, g = np.empty(len(pos)), np.empty((len(pos), 2))
        self.l.hh_query2d(self.G, which, _p(pos), C.c_int64(len(pos)), _p(d), _p(g))
        return d, g

    def line_visib(self, p1, p2, thresh, critical=False):
        p1, p2 = np.ascontiguousarray(p1, dtype=np.float64), np.ascontiguousarray(p2, dtype=np.float64)
        n = len(p1)
        vis, pc = np.zeros(n, dtype=np.int8), np.full((n, 3), np.nan)
        self.l.hh_line_visib(self.G, _p(p1), _p(p2), C.c_int64(n), C.c_double(thresh), int(critical), _p(vis, C.c_int8), _p(pc))
        return vis.astype(bool), pc

    def fk(self, pos10):
        pos10 = np.ascontiguousarray(pos10, dtype=np.float64)
        out = np.zeros((12, 3))
        self.l.hh_fk(self.P, _p(pos10), _p(out))
        return out

    def fk_adjoint(self, pos10, g):
        pos10, g = np.ascontiguousarray(pos10, dtype=np.float64), np.ascontiguousarray(g, dtype=np.float64)
        out = np.zeros(10)
        self.l.hh_fk_adjoint(self.P, _p(pos10), _p(g), _p(out))
        return out

    def penalty_eval(self, stage, N, coeff, T, sxy, exy, inner_xy, lam, rho):
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (coeff, T, sxy, exy, inner_xy, lam, rho)]
        gdC, gdT, terms, fxy = np.zeros((6 * N, 9)), np.zeros(N), np.zeros(13), np.zeros(2)
        self.l.hh_penalty_eval(self.P, self.G, stage, N, *[_p(v) for v in a], _p(gdC), _p(gdT), _p(terms), _p(fxy))
        return gdC, gdT, terms, fxy


class RogHarness:
    """rog_query.cuh on the CPU over the buffers of an oracle RogField."""

    def __init__(self, orog):
        self.l = lib()
        self.bufs = [np.ascontiguousarray(orog.download(w)) for w in (0, 2, 3)]     # dist3, critical, flat
        half = np.array(orog.half, dtype=np.int32)
        size = np.array(orog.size, dtype=np.int32)
        self.G = C.create_string_buffer(self.l.hh_grid_size())
        self.l.hh_make_rog_grid(C.c_double(orog.resolution), _p(half, C.c_int32), _p(size, C.c_int32),
                                *[_p(b) for b in self.bufs], self.G)

    def query(self, kind, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        d, g = np.empty(len(pos)), np.zeros((len(pos), 3))
        self.l.hh_rog_query(self.G, kind, _p(pos), C.c_int64(len(pos)), _p(d), _p(g))
        return d, g

    def line_free(self, s, e, thr):
        s, e = np.ascontiguousarray(s, dtype=np.float64), np.ascontiguousarray(e, dtype=np.float64)
        out = np.empty(len(s), dtype=np.int8)
        self.l.hh_rog_line_free(self.G, _p(s), _p(e), C.c_int64(len(s)), C.c_double(thr), _p(out, C.c_int8))
        return out

    def dispatch(self, pos):
        return field_dispatch(self.l, self.G, pos)


def field_dispatch(l, G, pos):
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    n = len(pos)
    d2, g2, d3, g3, v2, v3 = np.empty(n), np.empty((n, 2)), np.empty(n), np.empty((n, 3)), np.empty(n), np.empty(n)
    l.hh_field_dispatch(G, _p(pos), C.c_int64(n), _p(d2), _p(g2), _p(d3), _p(g3), _p(v2), _p(v3))
    return d2, g2, d3, g3, v2, v3


def field_misc(h: Harness, p2, q2, p3, idx, thr, critical):
    p2, q2, p3 = (np.ascontiguousarray(a, dtype=np.float64) for a in (p2, q2, p3))
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    n = len(p2)
    c2, c3, line = (np.empty(n, dtype=np.int8) for _ in range(3))
    cd, ci = np.empty(n), np.empty(n)
    h.l.hh_field_misc(h.G, _p(p2), _p(q2), _p(p3), _p(idx, C.c_int32), C.c_int64(n), C.c_double(thr), int(critical),
                      _p(c2, C.c_int8), _p(c3, C.c_int8), _p(line, C.c_int8), _p(cd), _p(ci))
    return c2.astype(bool), c3.astype(bool), line.astype(bool), cd, ci


def traj_sample(T, coeff, car_seq, t):
    l = lib()
    T, coeff, car_seq, t = (np.ascontiguousarray(a, dtype=np.float64) for a in (T, coeff, car_seq, t))
    m = len(t)
    st, ds, pva = np.empty((m, 10)), np.empty((m, 10)), np.empty((m, 27))
    l.hh_traj_sample(len(T), _p(T), _p(coeff), _p(car_seq), len(car_seq), _p(t), m, _p(st), _p(ds), _p(pva))
    return st, ds, pva


def edt_scan_line(f, segments=4):
    """One line through the two-scan lower envelopes of k_edt_scan's schedule (edt_line.cuh): sign-packed int32 in, (pos, neg, deepest stack) out."""
    f = np.ascontiguousarray(f, dtype=np.int32)
    n = len(f)
    pos, neg, depth = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32), C.c_int32(0)
    lib().hh_edt_scan_line(_p(f, C.c_int32), n, int(segments), _p(pos, C.c_int32), _p(neg, C.c_int32), C.byref(depth))
    return pos, neg, depth.value


def edt_line(f, chunk=32):
    """One line through the strided EDT pass's schedule (edt_line.cuh): sign-packed int32 in, (pos, neg) out."""
    l = lib()
    f = np.ascontiguousarray(f, dtype=np.int32)
    n = len(f)
    pos, neg = np.full(n, -1, dtype=np.int32), np.full(n, -1, dtype=np.int32)
    l.hh_edt_line(_p(f, C.c_int32), n, int(chunk), _p(pos, C.c_int32), _p(neg, C.c_int32))
    return pos, neg
