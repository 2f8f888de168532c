import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "needs_ref: needs oracle/_ref/libref_harness.so (the compiled reference)")


@pytest.fixture(scope="session")
def oracle():
    import reflib
    return reflib.Oracle()


@pytest.fixture(scope="session")
def ref():
    import reflib
    if not reflib.have_ref():
        pytest.skip("compiled reference (oracle/_ref/libref_harness.so) not available")
    return reflib.Ref()
