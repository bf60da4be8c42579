import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
# repo root (solvcon_b200, oracle), tests/ (helper modules: cases.py, the gloo workers), and the drop-in pair: the
# `_solvcon` shim + the reference's unmodified Python package staged by dropin/stage.py
for p in (os.path.join(ROOT, "dropin", "_stage"), os.path.join(ROOT, "dropin"), HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

REFSUITE = os.path.join(HERE, "_refsuite")
# staged reference test files that never construct a device solver (mesh tables, Gmsh reader, grid, cursors and the
# generic host Solver): they run in the CPU suite too; every other staged test needs the GPU
REFSUITE_HOST_FILES = {"test_mesh.py", "test_mesh_2d.py", "test_gmsh.py", "test_spacetime_grid.py",
                       "test_spacetime_celm_selm.py", "test_spacetime_solution.py"}


def pytest_collection_modifyitems(config, items):
    for item in items:
        path = str(item.fspath)
        if path.startswith(REFSUITE + os.sep) and os.path.basename(path) not in REFSUITE_HOST_FILES:
            item.add_marker(pytest.mark.gpu)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        from solvcon_b200 import _lib
        return _lib.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    """Bind the device library; GPU tests FAIL (not skip) if it cannot be loaded on a GPU box."""
    from solvcon_b200 import _lib
    _lib.load()
    if _lib.device_count() == 0:
        pytest.fail("gpu test selected but libscg_cuda sees no CUDA device (no CPU fallback exists)")
    _lib.init(0)
    return _lib


@pytest.fixture(scope="session")
def bridge():
    from oracle import bridge as b
    if not b.have_port():
        pytest.fail("oracle/liboracle_cese.so missing: run `make -C oracle port`")
    return b


@pytest.fixture(scope="session")
def refbridge(bridge):
    if not bridge.have_ref():
        pytest.skip("oracle/_ref/libsolvcon_ref.so not built (needs /root/reference)")
    return bridge
