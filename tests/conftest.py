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


# `-m gpu` tests that have not met a device yet go last and under a timeout (the driver runs the
# GPU suite with `-x`: a surprise in a new test must not hide the parity tests that already
# passed on a B200). Everything written in rounds 1 and 2 has run on a device (GPUTEST_r01, the
# r02 calls listed in profiles/README.md);
# add the file or test name of a NEW gpu test here until it has passed once on a B200.
_NOT_YET_RUN_ON_A_DEVICE = ()


def pytest_collection_modifyitems(config, items):
    def unproven(item):
        return any(tag in item.nodeid for tag in _NOT_YET_RUN_ON_A_DEVICE)

    items.sort(key=unproven)  # stable: keeps file order inside each group
    # A kernel that has never run could also hang; a blocked cudaDeviceSynchronize only yields to
    # pytest-timeout's thread method (stack dump + os._exit), which also tears the context down.
    for item in items:
        if unproven(item) and item.get_closest_marker("gpu") is not None:
            item.add_marker(pytest.mark.timeout(300, method="thread"))
