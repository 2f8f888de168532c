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


# ---- gated GPU modules (tests/test_gpu_z*.py) ---------------------------------------------------
# They hold the tests of code that was written without GPU time left (non-strict xfail).  Once one
# of their tests has failed, the rest of that module is skipped: a systematic problem (or a CUDA
# context left in an error state) then costs one test, not the whole module's time budget.  With
# --runxfail (tools/r2_first_gpu_call.sh) every test runs and reports on its own.
_gated_broken = set()


@pytest.hookimpl(hookwrapper=True)
def pytest_runtest_makereport(item, call):
    outcome = yield
    rep = outcome.get_result()
    name = getattr(item.module, "__name__", "")
    if call.when == "call" and name.startswith("test_gpu_z") and not item.config.getoption("runxfail"):
        if rep.failed or (rep.skipped and hasattr(rep, "wasxfail")):
            _gated_broken.add(name)


def pytest_runtest_setup(item):
    if getattr(item.module, "__name__", "") in _gated_broken:
        pytest.skip("an earlier test of this gated module failed; see its report")
