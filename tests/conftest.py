import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def engine_lib():
    """Builds (if stale) and loads libmvtm.so; GPU tests go through it, CPU tests only inspect its symbols."""
    from mvtopicmodel_b200 import build, _lib
    # One launch shape for every parity test: without this the engine samples ring depths 1, 2, 1, 2, ... over a view's first
    # passes and keeps the faster one -- a wall-clock decision that changes how many documents are resident, i.e. the
    # asynchrony of a stochastic trajectory.  Tests of the autotune itself delete the variable; mvtm_config.ring_depth wins over it.
    os.environ.setdefault("MVTM_RING", "1")
    if build.needs_build():
        build.build()
    return _lib.lib()
