import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    import importlib
    return importlib.import_module("rs-aware-differential-sfm_b200")


@pytest.fixture(scope="session")
def capi():
    import importlib
    import __graft_entry__ as ge
    ge.build()
    return importlib.import_module("rs-aware-differential-sfm_b200.capi")


@pytest.fixture(scope="session")
def synth():
    import importlib
    return importlib.import_module("rs-aware-differential-sfm_b200.synth")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def ctx(capi):
    """One GPU context for the whole session; fails loudly (no fallback) when no B200 is there."""
    c = capi.Context(0)
    yield c
    c.close()
