import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def _have_gpu():
    # the C ABI itself is the probe: solver_b200_new() returns NULL without a device (no torch import needed)
    try:
        from russell_b200 import _lib
        lib = _lib.load()
        h = lib.solver_b200_new()
        if h:
            lib.solver_b200_drop(h)
            return True
    except OSError:
        pass
    return False


def pytest_collection_modifyitems(config, items):
    # GPU tests never fake-pass: without a device they are skipped with an explicit reason
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if gpu_items and not _have_gpu():
        skip = pytest.mark.skip(reason="no CUDA device / libsolver_b200.so not loadable")
        for it in gpu_items:
            it.add_marker(skip)
