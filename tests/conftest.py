import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
collect_ignore_glob = ["ref_layout/*"]      # a stand-in tree with the reference's module names (fcn/test_dataset.py), not tests


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """GPU tests fail loudly (not skip) when selected without a device; they are simply deselected by
    `-m "not gpu"` on the CPU box."""
    return


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture
def knob():
    """Set library knobs (uoc_set_knob) for one test; every knob is back at its default afterwards."""
    from unseenobjectclustering_b200 import _lib
    touched = []

    def setter(name, value):
        touched.append(name)
        _lib.set_knob(name, value)

    yield setter
    for name in touched:
        _lib.set_knob(name, _lib.KNOB_DEFAULTS[name])
