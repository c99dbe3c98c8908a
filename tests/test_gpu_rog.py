"""GPU parity of the ROG-Map ring-buffer field through the C ABI against
oracle/oracle_rog.hpp: counters and the four persistent distance buffers bit-exact over a sequence of slides
and updates (ring wrap on every axis, stale-cell quirks included), queries within 1e-12 (FMA contraction
only), nearest-cell getters and isLineFree2d exact; at the full 803x803x83 size, properties that do not need
the CPU transform (sign structure, 1-Lipschitz, exact distance to an isolated obstacle)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

UNK, OCC, FREE = 1, 3, 4


def _pair(**kw):
    import oracle_lib as O
    import topay_b200 as tp
    from topay_b200.rog import ESDFMap
    d = dict(half_prob_map_size_i=(30, 22, 9), prob_resolution=0.1, esdf_resolution=0.1,
             local_update_box=(5.0, 3.0, 1.5), map_sliding_en=True)
    d.update(kw)
    desc = tp.rog_desc(**d)
    return ESDFMap(desc), O.RogField(desc)


def _random_points(rng, origin, n, spread):
    return origin + rng.uniform(-1, 1, (n, 3)) * np.array(spread)


def _compare(dev, orc, tag):
    oc, un = dev.getCounters()
    eo, eu = orc.download_counters()
    assert np.array_equal(oc, eo) and np.array_equal(un, eu), (tag, "counters")
    for which, key in enumerate(("dist3", "neg3", "critical", "flat")):
        assert np.array_equal(dev.getBuffer(which), orc.download(which)), (tag, key)


def test_geometry_and_empty_update():
    dev, orc = _pair()
    assert dev.map_size_i == orc.size and dev.half_map_size_i == orc.half
    assert dev.half_local_update_box_i == orc.half_box and dev.resolution == orc.resolution
    dev.updateESDF3D((0.0, 0.0, 0.0))
    orc.update_esdf((0.0, 0.0, 0.0))
    _compare(dev, orc, "empty")          # no obstacle: res * sqrt(DBL_MAX) everywhere in the box


def test_sequence_of_slides_and_updates_bit_exact():
    dev, orc = _pair()
    rng = np.random.default_rng(11)
    path = [(0.0, 0.0, 0.0), (0.43, -0.31, 0.12), (1.27, -0.9, 0.33), (1.31, 0.55, -0.21), (-0.8, 0.2, 0.05)]
    for step, odom in enumerate(path):
        dev.mapSliding(odom)
        orc.slide(odom)
        assert dev.local_map_origin_i == orc.origin_i
        pts = _random_points(rng, np.array(odom), 400, (2.8, 2.0, 0.8))
        dev.updateGridCounter(pts, UNK, OCC)
        orc.update_counters(pts, np.full(len(pts), UNK), np.full(len(pts), OCC))
        if step == 2:      # some cells are seen free again
            dev.updateGridCounter(pts[:150], OCC, FREE)
            orc.update_counters(pts[:150], np.full(150, OCC), np.full(150, FREE))
        dev.updateESDF3D(odom)
        orc.update_esdf(odom)
        _compare(dev, orc, step)


def test_update_box_clipped_by_the_map_and_full_box():
    for box in [(100.0, 100.0, 100.0), (0.45, 9.0, 0.3)]:
        dev, orc = _pair(local_update_box=box)
        rng = np.random.default_rng(5)
        odom = (0.7, -0.4, 0.2)
        for m in (dev.mapSliding, orc.slide):
            m(odom)
        pts = _random_points(rng, np.array(odom), 600, (3.0, 2.2, 0.9))
        dev.updateGridCounter(pts, UNK, OCC)
        orc.update_counters(pts, np.full(len(pts), UNK), np.full(len(pts), OCC))
        dev.updateESDF3D((1.9, -1.0, 0.5))          # box centred off the origin: clipped on one side
        orc.update_esdf((1.9, -1.0, 0.5))
        _compare(dev, orc, box)


def test_queries_and_line_checks_against_oracle():
    dev, orc = _pair()
    rng = np.random.default_rng(21)
    odom = (0.43, -0.31, 0.12)
    dev.mapSliding(odom)
    orc.slide(odom)
    pts = _random_points(rng, np.array(odom), 500, (2.8, 2.0, 0.8))
    dev.updateGridCounter(pts, UNK, OCC)
    orc.update_counters(pts, np.full(len(pts), UNK), np.full(len(pts), OCC))
    dev.updateESDF3D(odom)
    orc.update_esdf(odom)
    pos = np.concatenate([_random_points(rng, np.array(odom), 4000, (2.0, 1.3, 0.6)),
                          rng.uniform(-40, 40, (500, 3))])          # the second part wraps round the ring
    for kind in range(6):
        d, g = dev._query(kind, pos)
        ed, eg = orc.query(kind, pos)
        if kind >= 3:
            assert np.array_equal(d, ed), kind
        else:
            assert np.abs(d - ed).max() <= 1e-12 * max(1.0, np.abs(ed).max()), kind
            assert np.abs(g - eg).max() <= 1e-11 * max(1.0, np.abs(eg).max()), kind
    assert np.array_equal(dev.evaluateEDT(pos), dev.getValueGrad(pos)[0])
    s = _random_points(rng, np.array(odom), 3000, (2.0, 1.3, 0.0))[:, :2]
    e = _random_points(rng, np.array(odom), 3000, (2.0, 1.3, 0.0))[:, :2]
    e[:20] = s[:20]                                                   # same-cell rays
    for thr in (0.0, 0.15, 0.4):
        got = dev.isLineFree2d(s, e, thr)
        exp = orc.is_line_free2d(s, e, thr).astype(bool)
        assert np.array_equal(got, exp), thr
    assert 0 < dev.isLineFree2d(s, e, 0.15).sum() < len(s)


def test_full_size_ring_properties():
    """803 x 803 x 83. The CPU transform of 53.5 M cells takes minutes, so the full size is
    checked through properties: exact distances to isolated obstacles, sign structure, 1-Lipschitz."""
    import topay_b200 as tp
    from topay_b200.rog import ESDFMap
    dev = ESDFMap(tp.rog_desc())
    assert dev.map_size_i == (803, 803, 83)
    res = dev.resolution
    rng = np.random.default_rng(3)
    obs = np.floor(rng.uniform([-15, -15, 0.0], [15, 15, 1.5], (40, 3)) / res)        # cell indices
    centres = (obs + 0.5) * res
    dev.updateGridCounter(centres, UNK, OCC)
    dev.updateESDF3D((0.0, 0.0, 0.0))
    ms_total, ms_3d = dev.last_update_ms()
    assert ms_total > 0 and ms_3d > 0
    # nearest-cell value at random cells = res * distance (in cells) to the nearest obstacle cell
    q = np.floor(rng.uniform([-19, -19, -1.9], [19, 19, 1.9], (20000, 3)) / res)
    d = dev.getDistance((q + 0.5) * res)
    sq = ((q[:, None, :] - obs[None, :, :]) ** 2).sum(-1).min(1)
    exp = res * np.sqrt(sq)
    inside = sq == 0
    assert np.array_equal(d[~inside], exp[~inside])
    assert np.all(d[inside] == 0.0)          # an isolated occupied cell: pos 0, neg = res -> 0 - res + res
    # obstacles read 0 at their own cell and the field is 1-Lipschitz between neighbouring cells
    assert np.all(dev.getDistance(centres) == 0.0)
    d1 = dev.getDistance((q + 0.5) * res + np.array([res, 0, 0]))
    ok = np.abs(q[:, 0] + 1) < 399
    assert np.all(np.abs(d1 - d)[ok] <= res + 1e-12)
    # critical map: column distance ignores z
    d2 = dev.getCriticalDistance((q + 0.5) * res)
    sq2 = ((q[:, None, :2] - obs[None, :, :2]) ** 2).sum(-1).min(1)
    assert np.array_equal(d2[sq2 > 0], (res * np.sqrt(sq2))[sq2 > 0])


def test_solver_on_the_rog_field_matches_the_oracle(oracle):
    """GridMap's use_rog branches inside the solve (grid_map.h:364-392, 443-461): one cost / gradient
    evaluation per stage within 1e-9 of the oracle reading the same ring, then a batched solve whose
    success gate (solver-resident) agrees with the oracle's gate on the same trajectories."""
    import topay_b200 as tp
    from topay_b200 import scenes
    from topay_b200.rog import ESDFMap
    desc = tp.rog_desc(half_prob_map_size_i=(100, 100, 18), prob_resolution=0.1, esdf_resolution=0.1,
                       local_update_box=(19.0, 19.0, 3.0), map_sliding_en=False)
    dev, orc = ESDFMap(desc), oracle.RogField(desc)
    pts, _ = scenes.cuboids_scene(42)
    pts = pts[(np.abs(pts[:, 0]) < 9.4) & (np.abs(pts[:, 1]) < 9.4)]
    # one counter increment per occupied cell
    cells = np.unique(np.floor(pts.astype(np.float64) / 0.1).astype(np.int64), axis=0)
    centres = (cells + 0.5) * 0.1
    dev.updateGridCounter(centres, UNK, OCC)
    orc.update_counters(centres, np.full(len(centres), UNK), np.full(len(centres), OCC))
    dev.updateESDF3D((0.0, 0.0, 0.0))
    orc.update_esdf((0.0, 0.0, 0.0))
    assert np.array_equal(dev.getBuffer(0), orc.download(0))
    opt, rp = tp.opt_params_default(), tp.robot_params_default()
    n = 4
    paths, bv, ba = scenes.short_candidates(n, 77, span=6.0)
    solver = tp.MomaTrajOpt(dev, max_cand=n, max_pieces=16, opt_param=opt, robot=rp)
    assert solver.use_rog
    oracle.use_rog(orc)
    try:
        prep = [tp.prepare_candidate(opt, rp, paths[c], bv[c], ba[c], 16) for c in range(n)]
        for stage in (1, 2):
            args = dict(piece_num=[p["piece_num"] for p in prep], head_pva=np.stack([p["head_pva"] for p in prep]),
                        tail_pva=np.stack([p["tail_pva"] for p in prep]), start_xy=np.stack([p["start_xy"] for p in prep]),
                        end_xy=np.stack([p["end_xy"] for p in prep]), init_inner_xy=[p["init_inner_xy"] for p in prep],
                        x=[p["x0"] for p in prep])
            got = solver.evaluate(stage, **args)
            for c in range(n):
                p = prep[c]
                cost, grad, _, _, _ = oracle.eval_one(opt, rp, None, stage, p["piece_num"], p["head_pva"], p["tail_pva"],
                                                      p["start_xy"], p["end_xy"], p["init_inner_xy"], None, None, p["x0"])
                nv = tp.num_vars(p["piece_num"])
                assert abs(got["cost"][c] - cost) <= 1e-9 * abs(cost)
                assert np.abs(got["grad"][c, :nv] - grad[:nv]).max() <= 1e-9 * np.abs(grad).max()
        res = solver.optimizeTrajBatch(paths, bv, ba)
        assert res["status"].sum() >= 1
        arrs, best = solver.checkFeasibleBatch()
        trajs = [solver.getTraj(c)._tuple() for c in range(n)]
        exp = oracle.check_feasible(None, rp, trajs)
        assert np.array_equal(arrs["feasible"], exp["feasible"]) and np.array_equal(arrs["n_samples"], exp["n_samples"])
        assert np.allclose(arrs["min_dist"], exp["min_dist"], rtol=1e-10, atol=1e-12)
        assert np.allclose(arrs["min_dist_mani"], exp["min_dist_mani"], rtol=1e-10, atol=1e-12)
    finally:
        oracle.use_rog(None)


def test_rog_error_paths():
    import ctypes as C
    import topay_b200 as tp
    from topay_b200 import _lib
    from topay_b200.rog import ESDFMap
    with pytest.raises(_lib.TopayError) as e:       # counter resolution finer than the probability map (counter_map.cpp:48)
        ESDFMap(tp.rog_desc(prob_resolution=0.1, esdf_resolution=0.05))
    assert e.value.code == _lib.ERR_INVALID_ARG
    with pytest.raises(_lib.TopayError):            # unk_thresh outside [0, 1] (counter_map.cpp:54)
        ESDFMap(tp.rog_desc(half_prob_map_size_i=(4, 4, 2), unk_thresh=1.5))
    dev = ESDFMap(tp.rog_desc(half_prob_map_size_i=(8, 8, 4), prob_resolution=0.1, esdf_resolution=0.1))
    with pytest.raises(_lib.TopayError):
        dev._query(9, np.zeros((1, 3)))
    # a solver on a ring that was never updated refuses to run
    solver = tp.MomaTrajOpt(dev, max_cand=1, max_pieces=8)
    from topay_b200 import scenes
    paths, bv, ba = scenes.short_candidates(1, 3)
    with pytest.raises(_lib.TopayError) as e:
        solver.optimizeTrajBatch(paths, bv, ba)
    assert e.value.code == _lib.ERR_NOT_READY
    # empty update box (odometry far outside the local map): nothing changes, no error
    before = dev.getBuffer(0)
    dev.updateESDF3D((500.0, 0.0, 0.0))
    assert np.array_equal(before, dev.getBuffer(0))


def _scan_cloud(rng, pos, n, spread, walls):
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = np.minimum(np.abs(walls[0] / np.where(np.abs(d[:, 0]) > 1e-3, d[:, 0], 1e-3)),
                   np.abs(walls[1] / np.where(np.abs(d[:, 1]) > 1e-3, d[:, 1], 1e-3)))
    r = np.minimum(r, rng.uniform(0.2, spread, n))
    pts = pos + d * r[:, None]
    pts[::17] += rng.normal(size=(len(pts[::17]), 3)) * 5.0
    return np.concatenate([pts, rng.uniform(0, 100, (n, 1))], axis=1).astype(np.float32)


@pytest.mark.parametrize("cfg", [
    dict(half=(20, 20, 8), res=0.1, esdf_res=0.1, box=(3.0, 3.0, 1.2), upd=(3.5, 3.5, 1.4), filt=1, batch=1, inten=-1, ray=True, step=1),
    dict(half=(16, 18, 6), res=0.1, esdf_res=0.2, box=(2.5, 2.5, 1.0), upd=(99.0, 99.0, 99.0), filt=3, batch=2, inten=30, ray=True, step=1),
    dict(half=(14, 14, 6), res=0.1, esdf_res=0.1, box=(2.0, 2.0, 1.0), upd=(2.0, 2.0, 1.0), filt=2, batch=1, inten=-1, ray=False, step=2),
    dict(half=(60, 60, 12), res=0.05, esdf_res=0.05, box=(5.0, 5.0, 1.0), upd=(6.0, 6.0, 1.2), filt=1, batch=1, inten=-1, ray=True, step=1),
])
def test_prob_map_against_oracle(oracle, cfg):
    """N3: ProbMap::updateProbMap on the device — point filters, clipping to ceiling / ground / range / update box,
    the 3-D ray walks, batched hit / miss log-odds, type transitions into the ESDF counters, map sliding with
    resetCell, poses outside the map / above the ceiling, the first-frame sphere — occupancy_buffer_ (float), the
    ESDF counters and the ESDF buffers bit-identical to the oracle (= the reference's own ProbMap + ESDFMap,
    tests/test_ref_pin.py::test_prob_map_bit_exact) after every frame."""
    import topay_b200 as tp
    d = tp.rog_desc(half_prob_map_size_i=cfg["half"], prob_resolution=cfg["res"], esdf_resolution=cfg["esdf_res"],
                    local_update_box=cfg["box"], map_sliding_en=True)
    p = tp.prob_desc(ray_range=(0.3, 2.4), virtual_ceil_height=0.62, virtual_ground_height=-0.48,
                     local_update_box=cfg["upd"], map_sliding_thresh=0.25, point_filt_num=cfg["filt"],
                     batch_update_size=cfg["batch"], intensity_thresh=cfg["inten"], raycasting_en=cfg["ray"],
                     inflation_resolution=cfg["res"], inflation_step=cfg["step"])
    esdf = tp.ESDFMap(d)
    dev = tp.ProbMap(esdf, p)
    of = oracle.RogField(d)
    orc = oracle.RogProb(of, d, p)
    rng = np.random.default_rng(cfg["half"][1])
    poses = [(0.0, 0.0, 0.0), (0.05, 0.02, 0.0), (0.31, -0.22, 0.05), (0.33, -0.2, 0.05), (0.9, 0.7, 0.1),
             (0.9, 0.7, 0.9), (9.0, 9.0, 0.0), (9.1, 9.0, 0.0), (8.7, 9.3, -0.1)]
    touched = 0
    for step, pos in enumerate(poses):
        cloud = _scan_cloud(rng, np.array(pos), 3000, 3.0, (1.3, 1.7))
        dev.updateProbMap(cloud, pos)
        orc.update(cloud, pos)
        (a, oa), (b, ob) = dev.getOccupancyBuffer(), orc.download()
        assert oa == ob, step
        assert np.array_equal(a, b), (step, int((a != b).sum()))
        touched += int((b != 0).sum())
        for x, y in zip(esdf.getCounters(), of.download_counters()):
            assert np.array_equal(x, y), step
        if cfg["half"][0] == cfg["half"][1]:
            for which in range(4):
                assert np.array_equal(esdf.getBuffer(which), of.download(which)), (step, which)
    assert touched > 1000
    dev.close()
    esdf.close()
