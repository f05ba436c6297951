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


@pytest.fixture(params=["batch_forms", "mid_forms", "wide_forms"])
def kernel_form(request, monkeypatch):
    """K3 / K5 come in several launch shapes: the batch-sized CTAs bench.py's overlapped 256-scan steps use (K3 512 threads
    x 2 per SM, K5 128 x 3), K5's 192 x 2 form for a stream-ordered launch of up to two problems per SM, and one wide CTA
    per SM (K3 1024, K5 384) for batches of at most one scan per SM -- which is what nearly every test's small input would
    select.  GPU modules that use this fixture run each test in every shape (the library reads CFEAR_K3_WIDE /
    CFEAR_K5_WIDE / CFEAR_K5_FORM at every launch)."""
    if request.param == "mid_forms":
        monkeypatch.setenv("CFEAR_K5_FORM", "1")
    else:
        v = "1" if request.param == "wide_forms" else "0"
        monkeypatch.setenv("CFEAR_K3_WIDE", v)
        monkeypatch.setenv("CFEAR_K5_WIDE", v)
    return request.param
