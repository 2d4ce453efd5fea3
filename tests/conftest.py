import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def lib():
    from bikg_graph_explainability_public_b200 import _lib

    return _lib.load()


@pytest.fixture
def knobs(lib):
    """Engine options for one test: ``knobs(compact=0, seg=8)`` calls ``xpgnn_set_option`` (the XPGNN_* environment is read
    once at library load, include/xpgnn_b200.h); the previous values come back after the test."""
    import ctypes as C

    from bikg_graph_explainability_public_b200 import _lib

    saved = {}

    def setter(**kv):
        for k, v in kv.items():
            if k not in saved:
                old = C.c_int32(0)
                _lib.check(lib.xpgnn_get_option(k.encode(), C.byref(old)))
                saved[k] = old.value
            _lib.check(lib.xpgnn_set_option(k.encode(), int(v)))

    yield setter
    for k, v in saved.items():
        lib.xpgnn_set_option(k.encode(), v)
