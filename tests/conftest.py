import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle.oracle import Ref, REF_SO, build
    if not os.path.exists(REF_SO):
        try:
            build()
        except Exception:
            pass
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/falcon.so not available (reference tree absent)")
    return Ref()


@pytest.fixture(scope="session")
def engine():
    from falcon_b200.binding import Engine
    return Engine(0)


GOLDEN = os.path.join(ROOT, "tests", "golden")
