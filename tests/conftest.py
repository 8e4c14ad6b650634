import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _device_count():
    try:
        from texture_synthesis_b200 import capi
        return capi.device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) on a box without a CUDA device; the product itself never falls back to the CPU."""
    if _device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device: the CUDA path has no CPU fallback")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
