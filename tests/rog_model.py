"""Independent numpy model of the ROG-Map ring-buffer field, used to pin oracle/oracle_rog.hpp
(tests/test_oracle_rog.py) and, on the GPU box, as a second opinion next to the oracle. Brute-force
distance transforms (O(n^2)), so small maps only. Formulas cite /root/reference/src/rog_map."""
import numpy as np

UNK, OCC, FREE = 1, 3, 4          # rog_map::GridType (include/utils/common_lib.hpp:74-81)
DMAX = np.finfo(np.float64).max


class RogModel:
    def __init__(self, r):
        self.half = np.array(r.half)
        self.size = np.array(r.size)
        self.res = r.resolution
        self.res_inv = 1.0 / r.resolution
        self.half_box = np.array(r.half_box)
        self.origin = np.array(r.origin_i)

    # sliding_map.cpp:175-203 (ORIGIN_AT_CORNER)
    def pos_to_global(self, p):
        return np.floor(np.asarray(p, dtype=np.float64) * self.res_inv).astype(np.int64)

    def global_to_pos(self, g):
        return (np.asarray(g).astype(np.float64) + 0.5) * self.res

    # sliding_map.cpp:205-218: C remainder (sign of the dividend), then one normalisation step
    def global_to_local(self, g):
        g = np.asarray(g, dtype=np.int64)
        l = np.sign(g) * (np.abs(g) % self.size)
        l = np.where(l > self.half, l - self.size, l)
        l = np.where(l < -self.half, l + self.size, l)
        return l

    def hash_local(self, l):
        l = l + self.half
        return (l[..., 0] * self.size[1] + l[..., 1]) * self.size[2] + l[..., 2]

    def hash_global(self, g):
        return self.hash_local(self.global_to_local(g))

    def hash2_global(self, g):
        l = self.global_to_local(g) + self.half
        return l[..., 0] * self.size[1] + l[..., 1]

    def all_global_indices(self):
        ax = [np.arange(self.origin[i] - self.half[i], self.origin[i] + self.half[i] + 1) for i in range(3)]
        return np.stack(np.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3)


def _edt_sq(src):
    """Brute-force squared distance (in cells) from every cell of `src`'s grid to the nearest True cell;
    DMAX when there is none (the reference's sentinel survives the additions)."""
    shape = src.shape
    coords = np.stack(np.meshgrid(*[np.arange(s) for s in shape], indexing="ij"), -1).reshape(-1, len(shape))
    pts = coords[src.reshape(-1)]
    out = np.full(coords.shape[0], DMAX)
    if pts.shape[0]:
        best = np.full(coords.shape[0], np.iinfo(np.int64).max)
        for s in range(0, pts.shape[0], 256):
            d = ((coords[:, None, :] - pts[None, s:s + 256, :]) ** 2).sum(-1).min(1)
            best = np.minimum(best, d)
        out = best.astype(np.float64)
    return out.reshape(shape)


def brute_update(m, occ_cnt, odom, state=None):
    """What ESDFMap::updateESDF3D (esdf_map.cpp:154-500) leaves in its four persistent buffers."""
    S = m.size
    if state is None:
        state = dict(dist3=np.zeros(S), neg3=np.zeros(S), crit=np.zeros(S[:2]), flat=np.zeros(S[:2]))
    state = {k: v.copy() for k, v in state.items()}
    res = m.res
    cur = m.pos_to_global(odom)
    bmin, bmax = m.origin - m.half, m.origin + m.half
    lo = np.maximum(cur - m.half_box, bmin) - bmin
    hi = np.minimum(cur + m.half_box, bmax) - 1 - bmin
    idl = m.global_to_local(bmin) + m.half
    mem_end = S - 1 - idl
    if np.any(hi < lo):
        return state

    def wrap(a):
        q = np.arange(lo[a], hi[a] + 1)
        return np.where(q > mem_end[a], q + idl[a] - S[a], q + idl[a])

    mx, my, mz = wrap(0), wrap(1), wrap(2)
    occ = occ_cnt > 0
    box = occ[np.ix_(mx, my, mz)]
    pos = res * np.sqrt(_edt_sq(box))
    neg = res * np.sqrt(_edt_sq(~box))
    state["dist3"][np.ix_(mx, my, mz)] = pos
    state["neg3"][np.ix_(mx, my, mz)] = neg
    # combine (esdf_map.cpp:305-315): the box coordinates are used as memory coordinates
    sl = tuple(slice(lo[a], hi[a] + 1) for a in range(3))
    d, n = state["dist3"][sl], state["neg3"][sl]
    state["dist3"][sl] = np.where(n > 0.0, d + (-n + res), d)
    # 2-D maps
    lz = int((m.global_to_local(m.pos_to_global([0.0, 0.0, 0.155])) + m.half)[2])
    z_hi = lz if lz < hi[2] else hi[2]
    for key, zmax in (("crit", hi[2]), ("flat", z_hi)):
        nz = max(0, zmax - lo[2] + 1)
        col = box[:, :, :nz].any(axis=2) if nz > 0 else np.zeros(box.shape[:2], bool)
        p2 = res * np.sqrt(_edt_sq(col))
        n2 = res * np.sqrt(_edt_sq(~col))
        negbuf = np.zeros(S[:2])
        state[key][np.ix_(mx, my)] = p2
        negbuf[np.ix_(mx, my)] = n2
        # combine (esdf_map.cpp:391-398): y walks the x range; clipped to the row
        ys = slice(lo[0], min(hi[0], S[1] - 1) + 1)
        xs = slice(lo[0], hi[0] + 1)
        d, n = state[key][xs, ys], negbuf[xs, ys]
        state[key][xs, ys] = np.where(n > 0.0, d - n + res, d)
    return state


def np_query(m, bufs, kind, pos):
    """esdf_map.cpp:78-120 (cell getters) and :903-1097 (interpolated queries), vectorised."""
    pos = np.asarray(pos, dtype=np.float64)
    res, res_inv = m.res, m.res_inv
    d3, _, crit, flat = bufs
    n = pos.shape[0]
    grad = np.zeros((n, 3))
    if kind >= 3:
        g = m.pos_to_global(pos)
        if kind == 3:
            return d3.reshape(-1)[m.hash_global(g)], grad
        b = flat if kind == 4 else crit
        return b.reshape(-1)[m.hash2_global(g)], grad
    if kind == 0:
        pm = pos - 0.5 * res * np.ones(3)
    else:
        pm = pos - 0.5 * res * np.array([1.0, 1.0, 0.0])
    idx = m.pos_to_global(pm)
    f = (pos - m.global_to_pos(idx)) / res

    def tap3(dx, dy, dz):
        c = idx + np.array([dx, dy, dz])
        # the reference converts the tap index to a position and back (getDistance(pts[x][y][z]))
        return d3.reshape(-1)[m.hash_global(m.pos_to_global(m.global_to_pos(c)))]

    if kind == 0:
        d = [[[tap3(x, y, z) for z in range(2)] for y in range(2)] for x in range(2)]
        f0, f1, f2 = f[:, 0], f[:, 1], f[:, 2]
        v00 = (1 - f0) * d[0][0][0] + f0 * d[1][0][0]
        v01 = (1 - f0) * d[0][0][1] + f0 * d[1][0][1]
        v10 = (1 - f0) * d[0][1][0] + f0 * d[1][1][0]
        v11 = (1 - f0) * d[0][1][1] + f0 * d[1][1][1]
        v0 = (1 - f1) * v00 + f1 * v10
        v1 = (1 - f1) * v01 + f1 * v11
        dist = (1 - f2) * v0 + f2 * v1
        grad[:, 2] = (v1 - v0) * res_inv
        grad[:, 1] = ((1 - f2) * (v10 - v00) + f2 * (v11 - v01)) * res_inv
        g0 = (1 - f2) * (1 - f1) * (d[1][0][0] - d[0][0][0])
        g0 = g0 + (1 - f2) * f1 * (d[1][1][0] - d[0][1][0])
        g0 = g0 + f2 * (1 - f1) * (d[1][0][1] - d[0][0][1])
        g0 = g0 + f2 * f1 * (d[1][1][1] - d[0][1][1])
        grad[:, 0] = g0 * res_inv
        return dist, grad
    b = (flat if kind == 1 else crit).reshape(-1)

    def tap2(dx, dy):
        c = idx + np.array([dx, dy, 0])
        return b[m.hash2_global(m.pos_to_global(m.global_to_pos(c)))]

    d = [[tap2(x, y) for y in range(2)] for x in range(2)]
    f0, f1 = f[:, 0], f[:, 1]
    fxy1 = f0 * d[1][0] + (1 - f0) * d[0][0]
    fxy2 = f0 * d[1][1] + (1 - f0) * d[0][1]
    dist = (1 - f1) * fxy1 + f1 * fxy2
    grad[:, 0] = ((1 - f1) * (d[1][0] - d[0][0]) + f1 * (d[1][1] - d[0][1])) * res_inv
    grad[:, 1] = (-fxy1 + fxy2) * res_inv
    return dist, grad
