import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _gpu_reason():
    """None when libg3b.so loads and a B200 context can be created, else why not (asked once per session)."""
    try:
        from g3py_b200 import _cabi
        ctx = _cabi.Context(0)
        ctx.close()
        return None
    except Exception as e:                                   # missing library, no device, wrong architecture
        return "no usable B200 / libg3b.so: %s" % e


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on the GPU box must FAIL loudly if the device path is broken (no silent skip); a plain `pytest`
    # or any other selection on a CPU box skips the GPU tests instead of erroring in their fixtures.
    if "gpu" in (config.getoption("-m") or "") and "not gpu" not in (config.getoption("-m") or ""):
        return
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items:
        return
    why = _gpu_reason()
    if why is None:
        return
    skip = pytest.mark.skip(reason=why)
    for it in gpu_items:
        it.add_marker(skip)
