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
