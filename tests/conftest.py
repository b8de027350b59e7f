import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def zctx():
    """One libzkr context on cuda:0 for the whole GPU test session."""
    import ctypes as C
    from simple_zk_rollups_b200 import _lib
    L = _lib.lib()
    h = C.c_void_p()
    _lib.check(L.zkr_ctx_create(0, C.byref(h)))
    yield h
    L.zkr_ctx_destroy(h)
