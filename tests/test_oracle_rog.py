"""Known-answer tests that pin oracle/oracle_rog.hpp (the ROG-Map ESDFMap restatement, SURVEY.md §8 rows
a23/a24) against independent numpy brute force: no GPU, no reference at run time."""
import numpy as np
import pytest

import oracle_lib as O
from rog_model import RogModel, brute_update, np_query, OCC, FREE, UNK
from topay_b200._structs import rog_desc


def small_desc(**kw):
    d = dict(half_prob_map_size_i=(6, 5, 3), prob_resolution=0.1, esdf_resolution=0.1,
             local_update_box=(4.0, 4.0, 4.0), map_sliding_en=True)
    d.update(kw)
    return rog_desc(**d)


def test_geometry_matches_survey_row_a23():
    r = O.RogField(rog_desc())
    assert r.size == (803, 803, 83)        # (2*(400+1)+1, ., 2*(40+1)+1)
    assert r.half == (401, 401, 41)
    assert r.half_box == (400, 400, 40)    # floor(40/0.05)/2
    r2 = O.RogField(small_desc())
    assert r2.size == (15, 13, 9) and r2.half == (7, 6, 4)


def test_index_roundtrip_and_hash_layout():
    r = O.RogField(small_desc())
    m = RogModel(r)
    rng = np.random.default_rng(0)
    # every global index inside the local map hashes to a distinct ring cell
    for origin in [(0, 0, 0), (3, -5, 2), (-11, 40, -7)]:
        m.origin = np.array(origin)
        g = m.all_global_indices()
        h = m.hash_global(g)
        assert np.unique(h).size == g.shape[0] == np.prod(r.size)
    # ORIGIN_AT_CORNER: floor(pos / res), cell centre (i + 0.5) * res
    p = rng.uniform(-3, 3, (100, 3))
    gi = m.pos_to_global(p)
    c = m.global_to_pos(gi)
    assert np.all(np.abs(p - c) <= 0.05 + 1e-12)


def _fill_random(r, m, rng, frac=0.06):
    g = m.all_global_indices()
    pick = g[rng.random(g.shape[0]) < frac]
    pos = m.global_to_pos(pick)
    r.update_counters(pos, np.full(len(pos), UNK, np.uint8), np.full(len(pos), OCC, np.uint8))
    return pick


@pytest.mark.parametrize("box", [(4.0, 4.0, 4.0), (0.9, 0.7, 0.5)])
def test_update_esdf_no_wrap_matches_brute_force(box):
    r = O.RogField(small_desc(local_update_box=box))
    m = RogModel(r)
    rng = np.random.default_rng(1)
    _fill_random(r, m, rng)
    occ, _ = r.download_counters()
    assert occ.sum() > 10
    odom = (0.12, -0.07, 0.03)
    r.update_esdf(odom)
    exp = brute_update(m, occ, odom, state=None)
    for which, key in enumerate(("dist3", "neg3", "crit", "flat")):
        got = r.download(which)
        assert np.array_equal(got, exp[key]), key
    # inside the box, away from obstacles the value is the Euclidean distance to the nearest occupied
    # cell centre of the box; inside obstacles it is negative
    d3 = r.download(0)
    assert (d3[occ > 0] <= 0).sum() > 0


def test_update_esdf_ring_wrap_and_stale_state():
    """Slide the origin so that id_l != 0 on every axis; obstacles are placed by global position. The
    brute force scatters through the same wrap rule and applies the reference's unwrapped combine."""
    r = O.RogField(small_desc())
    m = RogModel(r)
    rng = np.random.default_rng(2)
    state = None
    for step, odom in enumerate([(0.0, 0.0, 0.0), (0.31, -0.22, 0.13), (0.58, -0.41, 0.27), (-0.2, 0.3, -0.1)]):
        r.slide(odom)
        m.origin = np.array(r.origin_i)
        assert tuple(m.origin) == tuple(np.floor(np.array(odom) / 0.1 + 1e-12).astype(int))
        _fill_random(r, m, rng, frac=0.03)
        occ, _ = r.download_counters()
        r.update_esdf(odom)
        state = brute_update(m, occ, odom, state=state)
        for which, key in enumerate(("dist3", "neg3", "crit", "flat")):
            assert np.array_equal(r.download(which), state[key]), (step, key)


def test_slide_clears_the_slabs_that_left_the_map():
    r = O.RogField(small_desc())
    m = RogModel(r)
    r.set_occupied_cnt(np.ones(r.size, np.int16))
    g_before = m.all_global_indices()
    r.slide((0.25, 0.0, -0.1))     # +2 cells in x, -1 in z
    m2 = RogModel(r)
    m2.origin = np.array(r.origin_i)
    assert tuple(m2.origin) == (2, 0, -1)
    occ, unk = r.download_counters()
    g_after = m2.all_global_indices()
    before = {tuple(x) for x in g_before}
    keep = np.array([tuple(x) in before for x in g_after])
    h = m2.hash_global(g_after)
    flat = occ.reshape(-1)
    assert np.all(flat[h[keep]] == 1)          # cells still inside keep their counters
    assert np.all(flat[h[~keep]] == 0)         # cells that entered the map start empty
    assert np.all(unk.reshape(-1)[h[~keep]] == 1)
    # a jump of more than the map size resets everything
    r.slide((10.0, 0.0, 0.0))
    occ, _ = r.download_counters()
    assert occ.sum() == 0


def test_counter_update_rules():
    r = O.RogField(small_desc())
    p = np.array([[0.05, 0.05, 0.05]] * 3)
    r.update_counters(p[:1], [UNK], [OCC])
    occ, unk = r.download_counters()
    assert occ.sum() == 1 and unk.sum() == np.prod(r.size) - 1
    r.update_counters(p[:1], [OCC], [FREE])
    occ, unk = r.download_counters()
    assert occ.sum() == 0 and unk.sum() == np.prod(r.size) - 1
    r.update_counters(p[:1], [FREE], [UNK])
    occ, unk = r.download_counters()
    assert unk.sum() == np.prod(r.size)


def test_queries_match_numpy_model_including_wrapped_positions():
    r = O.RogField(small_desc())
    m = RogModel(r)
    rng = np.random.default_rng(3)
    r.slide((0.31, -0.22, 0.13))
    m.origin = np.array(r.origin_i)
    _fill_random(r, m, rng)
    r.update_esdf((0.31, -0.22, 0.13))
    bufs = [r.download(w) for w in range(4)]
    pos = np.concatenate([rng.uniform(-0.6, 0.9, (300, 3)), rng.uniform(-5, 5, (100, 3))])  # the second part wraps
    for kind in range(6):
        d, g = r.query(kind, pos)
        ed, eg = np_query(m, bufs, kind, pos)
        assert np.array_equal(d, ed), kind
        if kind < 3:
            assert np.array_equal(g, eg), kind
    # evaluateEDT alone equals the value part of getValueGrad
    assert np.array_equal(r.evaluate_edt(pos), r.query(0, pos)[0])
    # at a cell centre the interpolation returns the cell value
    gi = m.pos_to_global(pos[:50])
    c = m.global_to_pos(gi)
    d, _ = r.query(0, c)
    assert np.allclose(d, r.query(3, c)[0], rtol=0, atol=1e-12)


def test_is_line_free2d():
    r = O.RogField(small_desc())
    m = RogModel(r)
    # a wall at x = 0.25 (cell 2) spanning all y, at ground level
    g = m.all_global_indices()
    wall = g[(g[:, 0] == 2) & (g[:, 2] == 0)]
    pos = m.global_to_pos(wall)
    r.update_counters(pos, np.full(len(pos), UNK, np.uint8), np.full(len(pos), OCC, np.uint8))
    r.update_esdf((0.0, 0.0, 0.0))
    s = np.array([[-0.45, 0.0], [-0.45, 0.0], [-0.45, 0.02], [0.05, 0.05], [-0.45, -0.3]])
    e = np.array([[0.55, 0.1], [0.05, 0.31], [-0.45, 0.03], [0.15, 0.05], [0.14, 0.31]])
    out = r.is_line_free2d(s, e, threshold=0.05)
    #   crosses the wall | stays left of it (clearance > 0.05) | same cell | ends next to the wall: the
    #   end cell itself is never tested and the start cell is 0.2 away | last tested cell 0.1 from the wall
    assert out.tolist() == [0, 1, 1, 1, 1]
    out = r.is_line_free2d(s, e, threshold=0.15)
    assert out.tolist() == [0, 1, 1, 1, 0]      # (1, 2), one cell from the wall, is crossed before the end cell
    out = r.is_line_free2d(s, e, threshold=0.25)
    assert out.tolist() == [0, 0, 1, 0, 0]      # cells two away from the wall (0.2) now count
