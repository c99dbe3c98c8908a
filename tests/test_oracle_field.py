"""Pins the field half of the oracle (parity unpinned by the reference: it ships no tests) with
analytic known-answer tests and a brute-force exact EDT on small grids."""
import numpy as np
import pytest

from topay_b200._structs import grid_desc

INT_MAX = np.iinfo(np.int32).max


def brute_sq(occ):
    """Exact squared distance (cells^2) to the nearest True cell, by brute force."""
    src = np.argwhere(occ)
    idx = np.indices(occ.shape).reshape(occ.ndim, -1).T
    if len(src) == 0:
        return np.full(occ.shape, INT_MAX, dtype=np.int64)
    d = ((idx[:, None, :] - src[None, :, :]) ** 2).sum(-1).min(1)
    return d.reshape(occ.shape)


def signed(res, sqp, sqn):
    big = np.sqrt(np.finfo(np.float64).max)
    dp = res * np.where(sqp == INT_MAX, big, np.sqrt(sqp.astype(np.float64)))
    dn = res * np.where(sqn == INT_MAX, big, np.sqrt(sqn.astype(np.float64)))
    return np.where(dn > 0.0, dp + (-dn + res), dp)


def small_field(oracle, shape=(12, 9, 5), res=0.1):
    d = grid_desc(map_size=tuple((s - 0.5) * res for s in shape), resolution=res)   # ceil() -> shape
    f = oracle.Field(d)
    assert f.dims == shape
    return f, d


def test_single_voxel_kat(oracle):
    f, d = small_field(oracle, (11, 11, 7))
    occ = np.zeros(f.dims, np.int8)
    occ[5, 5, 3] = 1
    f.set_occupancy(occ, np.zeros(f.dims[:2], np.int8), np.zeros(f.dims[:2], np.int8))
    f.rebuild()
    sqp, sqn = f.download_sqdist(3)
    i = np.indices(f.dims)
    want = (i[0] - 5) ** 2 + (i[1] - 5) ** 2 + (i[2] - 3) ** 2
    assert np.array_equal(sqp, want)
    assert sqn[5, 5, 3] == 1 and (sqn[occ == 0] == 0).all()
    e = f.download(3)
    assert e[5, 5, 3] == 0.0 + (-0.1 + 0.1)          # inside: pos 0, neg 1 cell -> 0
    assert e[5, 5, 0] == 0.1 * np.sqrt(9.0)


def test_plane_kat(oracle):
    f, d = small_field(oracle, (8, 6, 10))
    occ = np.zeros(f.dims, np.int8)
    occ[:, :, 0] = 1
    f.set_occupancy(occ, None, None)
    f.rebuild()
    sqp, _ = f.download_sqdist(3)
    z = np.indices(f.dims)[2]
    assert np.array_equal(sqp, z ** 2)


@pytest.mark.parametrize("seed,fill", [(0, 0.02), (1, 0.3), (2, 0.9)])
def test_matches_brute_force(oracle, seed, fill):
    f, d = small_field(oracle, (13, 10, 6))
    rng = np.random.default_rng(seed)
    occ3 = (rng.random(f.dims) < fill).astype(np.int8)
    occ2 = (rng.random(f.dims[:2]) < fill).astype(np.int8)
    occ2c = (rng.random(f.dims[:2]) < fill).astype(np.int8)
    f.set_occupancy(occ3, occ2, occ2c)
    f.rebuild()
    sqp, sqn = f.download_sqdist(3)
    assert np.array_equal(sqp, brute_sq(occ3 == 1)) and np.array_equal(sqn, brute_sq(occ3 == 0))
    assert np.array_equal(f.download(3), signed(0.1, sqp, sqn))
    sqp, sqn = f.download_sqdist(0)
    assert np.array_equal(sqp, brute_sq(occ2 == 1)) and np.array_equal(sqn, brute_sq(occ2 == 0))
    flat = f.download(0)
    assert np.array_equal(flat, signed(0.1, sqp, sqn))
    # inflated map: sources are the cells of the flat map below the chassis radius (grid_map.cpp:360)
    src = flat < 0.4
    sqp, sqn = f.download_sqdist(1)
    assert np.array_equal(sqp, brute_sq(src)) and np.array_equal(sqn, brute_sq(~src))


def test_empty_and_full_occupancy(oracle):
    f, d = small_field(oracle, (6, 5, 4))
    f.rebuild()   # nothing occupied: the positive transform has no source -> DBL_MAX sentinel
    sqp, sqn = f.download_sqdist(3)
    assert (sqp == INT_MAX).all() and (sqn == 0).all()
    assert np.all(f.download(3) == 0.1 * np.sqrt(np.finfo(np.float64).max))
    f.set_occupancy(np.ones(f.dims, np.int8), None, None)
    f.rebuild()
    sqp, sqn = f.download_sqdist(3)
    assert (sqp == 0).all() and (sqn == INT_MAX).all()


def test_rasterize_and_critical_accumulates(oracle):
    f, d = small_field(oracle, (10, 10, 4))
    pts = np.array([[0.01, 0.01, 0.05], [0.26, -0.31, 0.2], [5.0, 0.0, 0.0], [0.0, 0.0, 0.39]], np.float32)
    f.rasterize(pts)
    o3, o2, o2c = f.download_occupancy(3), f.download_occupancy(0), f.download_occupancy(2)
    assert o3.sum() == 3 and o2.sum() == 1 and o2c.sum() == 2     # the point at x = 5 is outside
    org = np.array([-d.map_size[0] / 2, -d.map_size[1] / 2, 0.0])
    for q in (0, 1, 3):
        i = np.floor((pts[q].astype(np.float64) - org) * (1.0 / 0.1)).astype(int)
        assert o3[i[0], i[1], i[2]] == 1 and o2c[i[0], i[1]] == 1
    i = np.floor((pts[0].astype(np.float64) - org) * 10.0).astype(int)
    assert o2[i[0], i[1]] == 1                                   # only z < 0.155 reaches the flat map
    f.clear(False)                                               # regenerateMap keeps critical (quirk 6)
    assert f.download_occupancy(3).sum() == 0 and f.download_occupancy(2).sum() == 2
    f.clear(True)
    assert f.download_occupancy(2).sum() == 0


def test_query_interpolation_and_bounds(oracle, small_scene):
    f = small_scene["field"]
    e3 = f.download(3)
    # at a cell centre the trilinear value is the cell value
    idx = np.array([[10, 20, 3], [150, 7, 0], [199, 199, 15]])
    pos = (idx + 0.5) * 0.1 + np.array([-10.0, -10.0, 0.0])
    d, g = f.query3d(pos)
    assert np.allclose(d, e3[idx[:, 0], idx[:, 1], idx[:, 2]], rtol=0, atol=1e-12)
    # outside the map (1e-4 margin): gradient variant returns 0, value variant 1e10
    out = np.array([[10.0, 0.0, 0.5], [0.0, -10.00001, 0.5], [0.0, 0.0, 1.6]])
    d, g = f.query3d(out)
    assert (d == 0).all() and (g == 0).all()
    assert (f.distance3d(out) == 1e10).all()
    # gradient = derivative of the interpolant (finite differences inside one cell)
    rng = np.random.default_rng(3)
    p = (rng.integers(5, 190, (200, 3)) % [200, 200, 14] + 0.3 + 0.4 * rng.random((200, 3))) * 0.1
    p += np.array([-10.0, -10.0, 0.0]) + 0.05
    d0, g0 = f.query3d(p)
    for k in range(3):
        e = np.zeros(3)
        e[k] = 1e-6
        fd = (f.query3d(p + e)[0] - f.query3d(p - e)[0]) / 2e-6
        assert np.abs(fd - g0[:, k]).max() < 1e-6
