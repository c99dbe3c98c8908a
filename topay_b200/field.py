"""GridMap — host-side mirror of the reference's dense distance field over the C ABI.

Same method names and argument meaning as nmoma_planner::GridMap
(src/map/include/map/grid_map.h:140-218): loadMap, regenerateMap-style ingest
(clear + rasterize), updateESDF, getDisWithGradI2d/3d, getDistance2d/3d,
isWholeBodyCollision, getESDFBuffer2d/3d, getOccBuffer2d/3d. Vectorised: positions are
(n, 2|3) arrays. All computation happens in libtopay_b200.so on the GPU.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._structs import MAP2D_CRITICAL, MAP2D_FLAT, MAP2D_INFLATE, MAP3D, GridDesc, RobotParams, grid_desc


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def robot_params_default():
    rp = RobotParams()
    _lib.lib().topay_robot_params_default(C.byref(rp))
    return rp


class GridMap:
    def __init__(self, desc: GridDesc = None, device: int = 0):
        self._l = _lib.lib()
        self.desc = desc if desc is not None else grid_desc()
        self.h = C.c_void_p()
        _lib.check(self._l.topay_field_create(C.byref(self.desc), device, C.byref(self.h)), "topay_field_create")
        d = (C.c_int32 * 3)()
        self._l.topay_field_dims(self.h, d)
        self.voxel_num = tuple(d)
        self.device = device

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self._l.topay_field_destroy(self.h)
            self.h = C.c_void_p()

    __del__ = close

    # ---- ingest -------------------------------------------------------------
    def loadMap(self, occ_2d, occ_3d, occ_2d_critical=None):
        """GridMap::loadMap (grid_map.cpp:800-809) followed by updateESDF."""
        a = [None if o is None else np.ascontiguousarray(o, dtype=np.int8) for o in (occ_3d, occ_2d, occ_2d_critical)]
        _lib.check(self._l.topay_field_set_occupancy(self.h, *[_p(x, C.c_int8) for x in a]), "set_occupancy")
        self.updateESDF()

    def clear(self, clear_critical=False):
        _lib.check(self._l.topay_field_clear(self.h, int(clear_critical)), "topay_field_clear")

    def rasterize(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        _lib.check(self._l.topay_field_rasterize_points(self.h, _p(xyz, C.c_float), xyz.shape[0]), "rasterize")

    def regenerateMap(self, xyz):
        """The ingest half of GridMap::regenerateMap (grid_map.cpp:716-750) for a given cloud."""
        self.clear(False)
        self.rasterize(xyz)
        self.updateESDF()

    def updateESDF(self):
        _lib.check(self._l.topay_field_rebuild(self.h), "topay_field_rebuild")

    def set_keep_sqdist(self, keep):
        _lib.check(self._l.topay_field_set_keep_sqdist(self.h, int(keep)), "set_keep_sqdist")

    def last_rebuild_ms(self):
        a, b = C.c_float(), C.c_float()
        self._l.topay_field_last_rebuild_ms(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    # ---- queries ------------------------------------------------------------
    def getDisWithGradI3d(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        n = pos.shape[0]
        d, g = np.empty(n), np.empty((n, 3))
        _lib.check(self._l.topay_field_query3d(self.h, _p(pos), n, _p(d), _p(g)), "topay_field_query3d")
        return d, g

    def getDisWithGradI2d(self, pos, inflate=False, critical=False):
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 2)
        n = pos.shape[0]
        which = MAP2D_CRITICAL if critical else (MAP2D_INFLATE if inflate else MAP2D_FLAT)
        d, g = np.empty(n), np.empty((n, 2))
        _lib.check(self._l.topay_field_query2d(self.h, _p(pos), n, which, _p(d), _p(g)), "topay_field_query2d")
        return d, g

    def getDistance3d(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        d = np.empty(pos.shape[0])
        _lib.check(self._l.topay_field_distance3d(self.h, _p(pos), pos.shape[0], _p(d)), "topay_field_distance3d")
        return d

    def getDistance2d(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 2)
        d = np.empty(pos.shape[0])
        _lib.check(self._l.topay_field_distance2d(self.h, _p(pos), pos.shape[0], _p(d)), "topay_field_distance2d")
        return d

    def isWholeBodyCollision(self, states, robot: RobotParams = None):
        states = np.ascontiguousarray(states, dtype=np.float64).reshape(-1, 10)
        rp = robot if robot is not None else robot_params_default()
        out = np.empty(states.shape[0], dtype=np.int8)
        _lib.check(self._l.topay_field_whole_body_collision(self.h, C.byref(rp), _p(states), states.shape[0],
                                                            _p(out, C.c_int8)), "whole_body_collision")
        return out.astype(bool)

    def _flags(self, fn, name, *arrays, threshold=0.0):
        n = arrays[0].shape[0]
        out = np.empty(n, dtype=np.int8)
        _lib.check(fn(self.h, *[_p(a) for a in arrays], n, C.c_double(threshold), _p(out, C.c_int8)), name)
        return out.astype(bool)

    def isCollision2d(self, pos, threshold):
        """GridMap::isCollision2d (grid_map.h:511-536) for (n, 2) positions."""
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 2)
        return self._flags(self._l.topay_field_is_collision2d, "is_collision2d", pos, threshold=threshold)

    def isCollision3d(self, pos, threshold):
        """GridMap::isCollision3d (grid_map.h:695-724) for (n, 3) positions."""
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        return self._flags(self._l.topay_field_is_collision3d, "is_collision3d", pos, threshold=threshold)

    def isLineCollisionGrid2d(self, p1, p2, threshold=0.0):
        """GridMap::isLineCollisionGrid2d (grid_map.h:565-611) for (n, 2) end points."""
        p1 = np.ascontiguousarray(p1, dtype=np.float64).reshape(-1, 2)
        p2 = np.ascontiguousarray(p2, dtype=np.float64).reshape(-1, 2)
        return self._flags(self._l.topay_field_is_line_collision_grid2d, "is_line_collision_grid2d", p1, p2,
                           threshold=threshold)

    def getDistCoarse2d(self, pos, critical=False):
        """GridMap::getDistCoarse2d (grid_map.h:887-912)."""
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 2)
        out = np.empty(pos.shape[0])
        _lib.check(self._l.topay_field_dist_coarse2d(self.h, _p(pos), pos.shape[0], int(critical), _p(out)),
                   "dist_coarse2d")
        return out

    def getDistCoarse2i(self, idx, critical=False):
        """GridMap::getDistCoarse2i (grid_map.h:914-940) for (n, 2) cell indices."""
        idx = np.ascontiguousarray(idx, dtype=np.int32).reshape(-1, 2)
        out = np.empty(idx.shape[0])
        _lib.check(self._l.topay_field_dist_coarse2i(self.h, _p(idx, C.c_int32), idx.shape[0], int(critical), _p(out)),
                   "dist_coarse2i")
        return out

    def lineVisib(self, p1, p2, thresh, use_critical=False, pc=None):
        """TopologyPRM::lineVisib (topo_prm.cpp:278-315) for (n, 3) end points on the device: (visible, pc). pc rows of
        visible segments keep what was passed in (nan by default)."""
        p1 = np.ascontiguousarray(p1, dtype=np.float64).reshape(-1, 3)
        p2 = np.ascontiguousarray(p2, dtype=np.float64).reshape(-1, 3)
        n = p1.shape[0]
        vis = np.zeros(n, dtype=np.int8)
        pc = np.full((n, 3), np.nan) if pc is None else np.ascontiguousarray(pc, dtype=np.float64).reshape(n, 3).copy()
        _lib.check(self._l.topay_field_line_visible(self.h, _p(p1), _p(p2), n, C.c_double(thresh), int(use_critical),
                                                    _p(vis, C.c_int8), _p(pc)), "line_visible")
        return vis.astype(bool), pc

    def sameTopoPaths(self, paths, pairs, thresh, use_critical=False):
        """TopologyPRM::sameTopoPath (topo_prm.cpp:424-448) for many pairs at once: `paths` is a list of (n_i, 3)
        polylines, `pairs` an (m, 2) array of path indices; returns m booleans."""
        paths = [np.ascontiguousarray(p, dtype=np.float64).reshape(-1, 3) for p in paths]
        pts = np.ascontiguousarray(np.concatenate(paths))
        off = np.cumsum([0] + [len(p) for p in paths]).astype(np.int32)
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        same = np.zeros(len(pairs), dtype=np.int8)
        _lib.check(self._l.topay_field_same_topo_paths(self.h, _p(pts), _p(off, C.c_int32), len(paths), _p(pairs, C.c_int32),
                                                       len(pairs), C.c_double(thresh), int(use_critical), _p(same, C.c_int8)),
                   "same_topo_paths")
        return same.astype(bool)

    # ---- index helpers (pure host arithmetic, grid_map.h:727-885) ---------------
    @property
    def map_origin(self):
        ms = np.array(self.desc.map_size[:])
        return np.array([-ms[0] / 2.0, -ms[1] / 2.0, 0.0])        # grid_map.cpp:41-47

    def posToIndex3d(self, pos):
        pos = np.asarray(pos, dtype=np.float64)
        return np.floor((pos - self.map_origin) * (1.0 / self.desc.resolution)).astype(np.int64)

    def posToIndex2d(self, pos):
        pos = np.asarray(pos, dtype=np.float64)
        return np.floor((pos - self.map_origin[:2]) * (1.0 / self.desc.resolution)).astype(np.int64)

    def indexToPos3d(self, idx):
        return (np.asarray(idx, dtype=np.float64) + 0.5) * self.desc.resolution + self.map_origin

    def indexToPos2d(self, idx):
        return (np.asarray(idx, dtype=np.float64) + 0.5) * self.desc.resolution + self.map_origin[:2]

    def boundIndex3d(self, idx):
        return np.clip(np.asarray(idx), 0, np.array(self.voxel_num) - 1)

    def boundIndex2d(self, idx):
        return np.clip(np.asarray(idx), 0, np.array(self.voxel_num[:2]) - 1)

    def isInMap3d(self, pos):
        """Position form (1e-4 margin, grid_map.h:818-833)."""
        pos = np.asarray(pos, dtype=np.float64)
        ms = np.array(self.desc.map_size[:])
        lo = self.map_origin + 1e-4
        hi = np.array([ms[0] / 2.0, ms[1] / 2.0, ms[2]]) - 1e-4
        return np.all((pos >= lo) & (pos <= hi), axis=-1)

    def isInMap2d(self, pos):
        pos = np.asarray(pos, dtype=np.float64)
        ms = np.array(self.desc.map_size[:2])
        return np.all((pos >= -ms / 2.0 + 1e-4) & (pos <= ms / 2.0 - 1e-4), axis=-1)

    # ---- buffers ------------------------------------------------------------
    def _shape(self, which):
        return self.voxel_num if which == MAP3D else self.voxel_num[:2]

    def _download(self, which):
        out = np.empty(self._shape(which))
        _lib.check(self._l.topay_field_download(self.h, which, _p(out)), "topay_field_download")
        return out

    def getESDFBuffer3d(self):
        return self._download(MAP3D)

    def getESDFBuffer2d(self):
        return self._download(MAP2D_FLAT)

    def getESDFBuffer2dInflate(self):
        return self._download(MAP2D_INFLATE)

    def getESDFBuffer2dCritical(self):
        return self._download(MAP2D_CRITICAL)

    def getSqDist(self, which):
        a = np.empty(self._shape(which), dtype=np.int32)
        b = np.empty(self._shape(which), dtype=np.int32)
        _lib.check(self._l.topay_field_download_sqdist(self.h, which, _p(a, C.c_int32), _p(b, C.c_int32)), "sqdist")
        return a, b

    def _occ(self, which):
        a = np.empty(self._shape(which), dtype=np.int8)
        _lib.check(self._l.topay_field_download_occupancy(self.h, which, _p(a, C.c_int8)), "occupancy")
        return a

    def getOccBuffer3d(self):
        return self._occ(MAP3D)

    def getOccBuffer2d(self):
        return self._occ(MAP2D_FLAT)

    def getOccBuffer2dCritical(self):
        return self._occ(MAP2D_CRITICAL)
