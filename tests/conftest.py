import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line(
        'markers', 'gpu: needs a CUDA device (run on the B200 box)')


def _gpu_count():
    try:
        from myokit_b200 import capi
        return capi.device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    # Tests marked `gpu` need a device: skipped (not errored) where there is none.
    if not any('gpu' in item.keywords for item in items):
        return
    if _gpu_count() > 0:
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)
