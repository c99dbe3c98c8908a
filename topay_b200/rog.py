"""ESDFMap — host-side mirror of the reference's ROG-Map distance field over the C ABI.

Same method names and argument meaning as rog_map::ESDFMap (src/rog_map/include/rog_map/esdf_map.h:34-93)
and the SlidingMap / CounterMap members it inherits (mapSliding, updateGridCounter): vectorised, positions
are (n, 3) arrays, line end points (n, 2). All computation happens in libtopay_b200.so on the GPU.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._structs import ProbDesc, RogDesc, prob_desc, rog_desc

# rog_map::GridType (include/utils/common_lib.hpp:74-81)
UNDEFINED, UNKNOWN, OUT_OF_MAP, OCCUPIED, KNOWN_FREE = 0, 1, 2, 3, 4
Q_EDT, Q_FLAT, Q_CRITICAL, Q_CELL, Q_CELL_FLAT, Q_CELL_CRITICAL = range(6)
BUF_DIST3, BUF_NEG3, BUF_CRITICAL, BUF_FLAT = range(4)


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _pos3(pos):
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    if pos.ndim != 2 or pos.shape[1] != 3:
        raise ValueError("positions must be (n, 3)")
    return pos


class ESDFMap:
    def __init__(self, desc: RogDesc = None, device: int = 0):
        """ESDFMap::initESDFMap (esdf_map.cpp:28-57)."""
        self._l = _lib.lib()
        self.desc = desc if desc is not None else rog_desc()
        self.h = C.c_void_p()
        _lib.check(self._l.topay_rogfield_create(C.byref(self.desc), device, C.byref(self.h)), "topay_rogfield_create")
        g = self._geometry()
        self.half_map_size_i, self.map_size_i, self.resolution, _, self.half_local_update_box_i = g
        self.device = device

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self._l.topay_rogfield_destroy(self.h)
            self.h = C.c_void_p()

    __del__ = close

    def _geometry(self):
        half, size, org, hb = ((C.c_int32 * 3)() for _ in range(4))
        res = C.c_double()
        _lib.check(self._l.topay_rogfield_geometry(self.h, half, size, C.byref(res), org, hb), "geometry")
        return tuple(half), tuple(size), res.value, tuple(org), tuple(hb)

    @property
    def local_map_origin_i(self):
        return self._geometry()[3]

    # ---- SlidingMap / CounterMap -------------------------------------------
    def mapSliding(self, odom):
        o = np.ascontiguousarray(odom, dtype=np.float64)
        _lib.check(self._l.topay_rogfield_slide(self.h, _p(o)), "topay_rogfield_slide")

    def updateGridCounter(self, pos, from_type, to_type):
        pos = _pos3(pos)
        a = np.ascontiguousarray(np.broadcast_to(from_type, pos.shape[:1]), dtype=np.uint8)
        b = np.ascontiguousarray(np.broadcast_to(to_type, pos.shape[:1]), dtype=np.uint8)
        _lib.check(self._l.topay_rogfield_update_counters(self.h, _p(pos), _p(a, C.c_uint8), _p(b, C.c_uint8),
                                                          pos.shape[0]), "topay_rogfield_update_counters")

    def setOccupiedCnt(self, cnt):
        cnt = np.ascontiguousarray(cnt, dtype=np.int16)
        if cnt.size != int(np.prod(self.map_size_i)):
            raise ValueError("occupied_cnt must have map_size_i cells")
        _lib.check(self._l.topay_rogfield_set_occupied_cnt(self.h, _p(cnt, C.c_int16)), "set_occupied_cnt")

    def getCounters(self):
        a = np.empty(self.map_size_i, dtype=np.int16)
        b = np.empty(self.map_size_i, dtype=np.int16)
        _lib.check(self._l.topay_rogfield_download_counters(self.h, _p(a, C.c_int16), _p(b, C.c_int16)), "counters")
        return a, b

    # ---- ESDFMap -----------------------------------------------------------
    def updateESDF3D(self, cur_odom):
        o = np.ascontiguousarray(cur_odom, dtype=np.float64)
        _lib.check(self._l.topay_rogfield_update_esdf(self.h, _p(o)), "topay_rogfield_update_esdf")

    def last_update_ms(self):
        a, b = C.c_float(), C.c_float()
        self._l.topay_rogfield_last_update_ms(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def _query(self, kind, pos, want_grad=True):
        pos = _pos3(pos)
        n = pos.shape[0]
        d = np.empty(n)
        g = np.zeros((n, 3)) if want_grad else None
        _lib.check(self._l.topay_rogfield_query(self.h, kind, _p(pos), n, _p(d), _p(g)), "topay_rogfield_query")
        return (d, g) if want_grad else d

    def evaluateEDT(self, pos):
        return self._query(Q_EDT, pos, want_grad=False)

    def evaluateFirstGrad(self, pos):
        return self._query(Q_EDT, pos)[1]

    def getValueGrad(self, pos):
        return self._query(Q_EDT, pos)

    def getValueGrad2d(self, pos):
        return self._query(Q_FLAT, pos)

    def getCriticalValueGrad(self, pos):
        return self._query(Q_CRITICAL, pos)

    def getDistance(self, pos):
        return self._query(Q_CELL, pos, want_grad=False)

    def getDistance2d(self, pos):
        return self._query(Q_CELL_FLAT, pos, want_grad=False)

    def getCriticalDistance(self, pos):
        return self._query(Q_CELL_CRITICAL, pos, want_grad=False)

    def isLineFree2d(self, start, end, threshold=0.0):
        s = np.ascontiguousarray(start, dtype=np.float64)
        e = np.ascontiguousarray(end, dtype=np.float64)
        out = np.empty(s.shape[0], dtype=np.int8)
        _lib.check(self._l.topay_rogfield_is_line_free2d(self.h, _p(s), _p(e), s.shape[0], C.c_double(threshold),
                                                         _p(out, C.c_int8)), "topay_rogfield_is_line_free2d")
        return out.astype(bool)

    def getBuffer(self, which):
        out = np.empty(self.map_size_i if which <= BUF_NEG3 else self.map_size_i[:2])
        _lib.check(self._l.topay_rogfield_download(self.h, which, _p(out)), "topay_rogfield_download")
        return out


class ProbMap:
    """rog_map::ProbMap (src/rog_map/include/rog_map/prob_map.h) over an ESDFMap: the probabilistic occupancy layer
    that ingests point clouds and drives the ESDF counter map (the inflation map and the frontier counters of the
    reference are not built). `esdf_map_` is the ESDFMap it feeds, as in the reference."""

    def __init__(self, esdf_map: ESDFMap, desc: ProbDesc = None):
        """initProbMap (prob_map.cpp:25-88); the grid geometry (half_map_size_i, resolution, sliding, fixed origin)
        is the ESDFMap descriptor's."""
        self._l = _lib.lib()
        self.esdf_map_ = esdf_map
        self.cfg_ = desc if desc is not None else prob_desc(inflation_resolution=esdf_map.desc.prob_resolution)
        self.h = C.c_void_p()
        _lib.check(self._l.topay_probmap_create(esdf_map.h, C.byref(self.cfg_), C.byref(self.h)), "topay_probmap_create")
        self.map_size_i = tuple(2 * int(h) + 1 for h in esdf_map.desc.half_prob_map_size_i)

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self._l.topay_probmap_destroy(self.h)
            self.h = C.c_void_p()

    __del__ = close

    def updateProbMap(self, cloud_xyzi, pose_pos):
        """ProbMap::updateProbMap(cloud, pose) (prob_map.cpp:302-373), incl. esdf_map_->updateESDF3D(pos).
        cloud_xyzi: (n, 4) x, y, z, intensity."""
        c = np.ascontiguousarray(cloud_xyzi, dtype=np.float32).reshape(-1, 4)
        p = np.ascontiguousarray(pose_pos, dtype=np.float64)
        _lib.check(self._l.topay_probmap_update(self.h, _p(c, C.c_float), C.c_int64(c.shape[0]), _p(p)),
                   "topay_probmap_update")

    def setFirstFrame(self, armed):
        _lib.check(self._l.topay_probmap_set_first_frame(self.h, int(armed)), "topay_probmap_set_first_frame")

    def getOccupancyBuffer(self):
        """occupancy_buffer_ (float log-odds in ring memory) and local_map_origin_i_."""
        occ = np.empty(self.map_size_i, dtype=np.float32)
        org = (C.c_int32 * 3)()
        _lib.check(self._l.topay_probmap_download(self.h, _p(occ, C.c_float), org), "topay_probmap_download")
        return occ, tuple(org)
