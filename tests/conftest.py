import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module("3dfeaturematcher_b200")


@pytest.fixture(scope="session")
def api():
    return importlib.import_module("3dfeaturematcher_b200.api")


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module("3dfeaturematcher_b200.synth")


@pytest.fixture(scope="session")
def ctx(api):
    """One GPU context for the whole session; fails loudly when the CUDA path is unusable."""
    c = api.Context(0)
    yield c
    c.close()
