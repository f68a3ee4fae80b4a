import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the live reference tree (/root/reference, build container only)")


def pytest_collection_modifyitems(config, items):
    from oracle.ref_harness import reference_available

    if reference_available():
        return
    skip = pytest.mark.skip(reason="reference tree not present")
    for it in items:
        if "reference" in it.keywords:
            it.add_marker(skip)
