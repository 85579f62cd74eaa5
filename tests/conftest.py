import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def ref():
    """The compiled reference (oracle/_ref/libwpt_ref_strict.so); tests needing it skip when it is absent."""
    from oracle import refapi
    if not refapi.available("strict"):
        pytest.skip("oracle/_ref/libwpt_ref_strict.so not built (needs /root/reference at build time)")
    return refapi


@pytest.fixture(scope="session")
def capi():
    from wurblpt_b200 import capi as c
    c.lib()
    return c
