import os
import sys

import pytest

# The multi-rank tests emulate several GPUs' ranks on ONE device (tests/test_gpu_sharded.py): a rank's barrier
# kernel spins until its peers arrive, so nothing a peer's host thread does may wait for the whole device.
# Lazy module loading does exactly that on a kernel's first launch, and streams that share a hardware queue
# serialise behind the spinning kernel -- both are artefacts of sharing a device, not of the multi-GPU path.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

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
