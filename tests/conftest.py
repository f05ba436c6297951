import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(params=["batch_forms", "wide_forms"])
def kernel_form(request, monkeypatch):
    """K3 / K5 come in two launch shapes: the batch-sized CTAs bench.py's 256-scan steps use, and one wide CTA per SM for
    batches of at most one scan per SM -- which is what nearly every test's small input would select.  GPU modules that
    use this fixture run each test both ways (the library reads CFEAR_K3_WIDE / CFEAR_K5_WIDE at every launch)."""
    v = "1" if request.param == "wide_forms" else "0"
    monkeypatch.setenv("CFEAR_K3_WIDE", v)
    monkeypatch.setenv("CFEAR_K5_WIDE", v)
    return request.param
