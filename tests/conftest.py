import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA GPU (run on the B200 box)")


@pytest.fixture(scope="session")
def product_lib():
    """Builds (if stale) and loads libmodle_b200.so; nvcc cross-compiles without a GPU."""
    from modle_b200 import build, host

    build.build()
    return host.lib()


@pytest.fixture(scope="session")
def gpu_ctx(product_lib):
    from modle_b200.simulation import Context

    ctx = Context(0, rng_mode=0)
    yield ctx
    ctx.close()
