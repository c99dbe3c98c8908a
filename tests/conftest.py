import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _product_library():
    """The built library is git-ignored: compile it (nvcc cross-compiles without a GPU) when a fresh checkout
    runs the tests before __graft_entry__.build(). Building is not a fallback — without the library nothing runs."""
    from topay_b200 import _lib
    if not os.path.exists(_lib.SO_PATH):
        _lib.build()


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def small_scene(oracle):
    """Default-size field (200 x 200 x 16 @ 0.1 m, params/grid_map.yaml) on a seeded cuboids scene."""
    from topay_b200 import scenes
    from topay_b200._structs import grid_desc
    pts, boxes = scenes.cuboids_scene(42)
    f = oracle.Field(grid_desc())
    f.rasterize(pts)
    f.rebuild()
    return dict(points=pts, boxes=boxes, field=f, desc=grid_desc())
