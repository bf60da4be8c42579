import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


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
