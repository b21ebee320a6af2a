import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    from oracle import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref not built (needs /root/reference in the build container)")
    return Reference()


@pytest.fixture(scope="session")
def emu():
    from emu import Emu
    return Emu()


@pytest.fixture(scope="session")
def cuda():
    """The product backend, selected through the reference-shaped API (select_backend)."""
    import portablert_b200 as prt
    assert prt.cuda_backend.is_available(), "no CC 10.x device: -m gpu tests need a B200"
    prt.select_backend(prt.cuda_backend)
    yield prt.cuda_backend
    prt.cuda_backend.shutdown()


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
