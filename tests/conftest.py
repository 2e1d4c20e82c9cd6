import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line(
        "markers", "gpu: needs a B200 (run with `pytest -m gpu` on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The reference implementation compiled from its own sources
    (oracle/_ref); skipped when it has not been built."""
    import harness
    if not harness.have_oracle():
        pytest.skip("oracle/_ref/libpumas_oracle.so not built "
                    "(make -C oracle; needs /root/reference)")
    return harness.Oracle()


@pytest.fixture(scope="session")
def port():
    import harness
    if not harness.have_port():
        pytest.skip("oracle/_port/libpumas_port.so not built (make -C oracle port)")
    return harness.Port()
