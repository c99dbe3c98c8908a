"""The C-ABI library loads without a GPU, exports every symbol include/topay_b200.h declares,
agrees with the oracle on the pure-host entry points, and refuses to compute without a device."""
import ctypes as C
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "topay_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(topay_[a-z0-9_]+)\s*\(", src))
    names.discard("topay_num_vars")     # static inline
    return sorted(names)


def test_library_exports_every_declared_symbol():
    from topay_b200 import _lib
    lib = _lib.lib()
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in topay_b200.h but not exported"
    assert set(_lib.PROTOTYPES) == set(syms), set(_lib.PROTOTYPES) ^ set(syms)
    assert b"sm_100a" in lib.topay_version()


def test_struct_sizes_match_the_compiler():
    import subprocess
    import tempfile
    from topay_b200 import _structs as S
    prog = ('#include "topay_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
            "sizeof(topay_robot_params),sizeof(topay_lbfgs_params),sizeof(topay_opt_params),sizeof(topay_grid_desc),"
            "sizeof(topay_problem_batch),sizeof(topay_result_batch),sizeof(topay_solver_stats),"
            "sizeof(topay_rog_desc),sizeof(topay_traj_batch),sizeof(topay_feasibility));return 0;}")
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "a.c"), "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "a"),
                               os.path.join(d, "a.c")])
        sizes = list(map(int, subprocess.check_output([os.path.join(d, "a")]).split()))
    want = [C.sizeof(x) for x in (S.RobotParams, S.LbfgsParams, S.OptParams, S.GridDesc, S.ProblemBatch,
                                  S.ResultBatch, S.SolverStats, S.RogDesc, S.TrajBatch, S.Feasibility)]
    assert sizes == want


def test_host_entry_points_match_the_oracle(oracle):
    import topay_b200 as tp
    from topay_b200 import scenes
    assert bytes(tp.robot_params_default()) == bytes(oracle.robot_defaults())
    assert bytes(tp.opt_params_default()) == bytes(oracle.opt_defaults())
    opt, rp = tp.opt_params_default(), tp.robot_params_default()
    paths, bv, ba = scenes.short_candidates(6, 3)
    paths += scenes.synthetic_batch(2, 5)[0]
    for p in paths:
        a = tp.prepare_candidate(opt, rp, p, bv[0], ba[0], 64)
        b = oracle.prepare_candidate(opt, rp, p, bv[0], ba[0], 64)
        assert a["piece_num"] == b["piece_num"] and a["s1_past"] == b["s1_past"] and a["rc"] == b["rc"] == 0
        for k in ("head_pva", "tail_pva", "start_xy", "end_xy", "init_inner_xy", "x0"):
            assert np.array_equal(a[k], b[k]), k


def test_compute_fails_loudly_without_a_device():
    import pytest
    import topay_b200 as tp
    from topay_b200 import _lib
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.TopayError) as e:
        tp.GridMap(tp.grid_desc())
    assert e.value.code == _lib.ERR_NO_DEVICE
    h = C.c_void_p()
    assert _lib.lib().topay_field_create(C.byref(tp.grid_desc()), 0, C.byref(h)) == _lib.ERR_NO_DEVICE
    # the ROG ring and the trajectory post-processing refuse as well
    with pytest.raises(_lib.TopayError) as e:
        tp.ESDFMap(tp.rog_desc(half_prob_map_size_i=(4, 4, 2)))
    assert e.value.code == _lib.ERR_NO_DEVICE
    import numpy as np
    from topay_b200.optimizer import MomaTraj
    with pytest.raises(_lib.TopayError) as e:
        MomaTraj(np.ones(1), np.zeros((6, 9)), np.zeros(3)).getState(0.5)
    assert e.value.code == _lib.ERR_NO_DEVICE


def test_select_shortest_is_host_only():
    """topay_select_shortest (planner.cpp:999-1010) needs no device: first success, strictly shorter replaces."""
    import numpy as np
    from topay_b200 import _lib
    l = _lib.lib()

    def pick(ok, dur):
        a = np.ascontiguousarray(ok, dtype=np.int32)
        d = np.ascontiguousarray(dur, dtype=np.float64)
        return l.topay_select_shortest(a.ctypes.data_as(C.POINTER(C.c_int32)), d.ctypes.data_as(C.POINTER(C.c_double)), len(a))
    assert pick([0, 1, 1, 1], [1.0, 5.0, 3.0, 3.0]) == 2
    assert pick([0, 0], [1.0, 2.0]) == -1
    assert pick([1, 0, 1], [2.0, 1.0, 2.0]) == 0


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "topay_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle_lib" not in txt and "liboracle" not in txt and "oracle/" not in txt, f


def test_cpp_shims_compile_against_the_header():
    """shim/topay_shim.hpp (GridMap / MomaTrajOpt / MomaTraj / rog_map::ESDFMap look-alikes) and the example
    call site of INTEGRATION.md §3 compile with the host compiler alone."""
    import subprocess
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "shim"), os.path.join(ROOT, "shim", "example_worker.cpp")])


def test_gridmap_index_helpers_match_the_oracle_grid(oracle):
    """posToIndex / indexToPos / boundIndex / isInMap (grid_map.h:727-885) are host arithmetic: no device."""
    import numpy as np
    import topay_b200 as tp
    g = tp.GridMap.__new__(tp.GridMap)           # geometry only, no device object
    g.desc, g.voxel_num = tp.grid_desc(), (200, 200, 16)
    rng = np.random.default_rng(0)
    p = rng.uniform([-10.4, -10.4, -0.2], [10.4, 10.4, 1.8], (2000, 3))
    idx = g.posToIndex3d(p)
    assert np.array_equal(idx, np.floor((p - np.array([-10.0, -10.0, 0.0])) * (1.0 / 0.1)).astype(np.int64))
    inside = g.isInMap3d(p)
    c = g.indexToPos3d(idx[inside])
    assert np.all(np.abs(c - p[inside]) <= 0.05 + 1e-12)
    assert np.array_equal(g.posToIndex2d(p[:, :2]), idx[:, :2])
    b = g.boundIndex3d(idx)
    assert b.min() >= 0 and np.all(b.max(axis=0) <= np.array([199, 199, 15]))
    # the nearest-cell lookups of the oracle use the same indices
    f = oracle.Field(tp.grid_desc())
    f.rebuild()
    assert np.array_equal(g.isInMap2d(p[:, :2]), f.distance2d(p[:, :2]) != 1e10)     # 1e10 <=> outside (grid_map.h:269-273)
    assert np.array_equal(g.isInMap3d(p), f.distance3d(p) != 1e10)


def test_plain_c_caller_links_and_runs(tmp_path):
    """examples/c_abi_demo.c: the header is C99-clean and the library is callable without C++ or Python; on a
    box without a GPU the compute entry points refuse (no CPU path), the host-only ones answer."""
    import subprocess
    from topay_b200 import _lib
    exe = str(tmp_path / "c_abi_demo")
    libdir = os.path.dirname(_lib.SO_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "c_abi_demo.c"), "-L", libdir, "-ltopay_b200",
                           f"-Wl,-rpath,{libdir}", "-lm", "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "select_shortest -> 2" in out.stdout
    assert ("no CUDA device" in out.stdout) or ("status 1" in out.stdout)


def test_dense_path_matches_the_oracle(oracle):
    """N2: topay_dense_path (host arithmetic, no GPU) against the oracle (= the reference's getDensePath bit for bit,
    test_ref_pin.py::test_dense_path_bit_exact): ragged polylines, short and zero-length segments, straight
    continuations, a capacity too small for the path, bad arguments."""
    import topay_b200 as tp
    rng = np.random.default_rng(15)
    for i in range(80):
        k = int(rng.integers(2, 8))
        raw = rng.uniform(-9, 9, (k, 2))
        if i % 4 == 0:
            raw[1] = raw[0] + np.array([0.2, 0.1])
        if i % 6 == 0 and k > 2:
            raw[2] = raw[1] + (raw[1] - raw[0])
        if i % 9 == 0 and k > 2:
            raw[2] = raw[1]
        y0, y1 = rng.uniform(-np.pi, np.pi, 2)
        for step in (1.414, 0.3):
            a = tp.getDensePath(raw, step, y0, y1, 1.0, 1.25)
            b = oracle.dense_path(raw, step, y0, y1, 1.0, 1.25)
            assert a.shape == b.shape and np.array_equal(a, b), (i, step)
    from topay_b200 import _lib
    import ctypes as C
    raw = np.array([[0.0, 0.0], [5.0, 0.0]])
    out = np.zeros((2, 4))
    dp = C.POINTER(C.c_double)
    n = _lib.lib().topay_dense_path(raw.ctypes.data_as(dp), 2, 1.0, 0.0, 0.0, 1.0, 1.0, out.ctypes.data_as(dp), 2)
    assert n == len(tp.getDensePath(raw, 1.0, 0.0, 0.0, 1.0, 1.0)) > 2
    assert _lib.lib().topay_dense_path(raw.ctypes.data_as(dp), 1, 1.0, 0.0, 0.0, 1.0, 1.0, out.ctypes.data_as(dp), 2) < 0
    assert _lib.lib().topay_dense_path(raw.ctypes.data_as(dp), 2, 0.0, 0.0, 0.0, 1.0, 1.0, out.ctypes.data_as(dp), 2) < 0


def test_discretize_path_matches_the_oracle(oracle):
    """N2: topay_path_length / topay_discretize_path (host arithmetic) against the oracle (= TopologyPRM::pathLength /
    discretizePath of the compiled reference, test_ref_pin.py::test_same_topo_path_bit_exact): ragged polylines, a
    repeated waypoint (zero-length segment: the reference's 0 / 0 is reproduced), bad arguments."""
    import topay_b200 as tp
    from topay_b200 import _lib
    import ctypes as C
    rng = np.random.default_rng(31)
    for i in range(200):
        k = int(rng.integers(2, 8))
        p = rng.uniform(-9, 9, (k, 3))
        if i % 5 == 0 and k > 2:
            p[2] = p[1]
        n = int(rng.integers(2, 200))
        assert np.array_equal(tp.discretizePath(p, n), oracle.discretize_path(p, n), equal_nan=True)
        assert tp.pathLength(p) == oracle.path_length(p)
    dp = C.POINTER(C.c_double)
    p, out = np.zeros((2, 3)), np.zeros((4, 3))
    assert _lib.lib().topay_discretize_path(p.ctypes.data_as(dp), 2, 1, out.ctypes.data_as(dp)) < 0
    assert _lib.lib().topay_discretize_path(p.ctypes.data_as(dp), 1, 4, out.ctypes.data_as(dp)) < 0
