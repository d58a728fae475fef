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


# `-m gpu` tests written after the round's GPU budget was spent have not met a device yet
# (profiles/README.md, "Not re-measured after the last code changes"). The driver runs the GPU
# suite with `-x`, so they go last: a surprise in one of them must not hide the parity tests
# that have already passed on a B200. Drop a name from this list once it has run on a device.
_NOT_YET_RUN_ON_A_DEVICE = (
    "test_zz_gpu_throughput_mode.py",
    "test_statistical_parity.py",
    "test_internal_state_log_matches_oracle",
    "test_barriers_outside_the_interval_are_dead_but_draw",
    "test_cuda_occupancy_profile_matches_oracle",
)


def pytest_collection_modifyitems(config, items):
    def unproven(item):
        return any(tag in item.nodeid for tag in _NOT_YET_RUN_ON_A_DEVICE)

    items.sort(key=unproven)  # stable: keeps file order inside each group
    # A kernel that has never run could also hang; a blocked cudaDeviceSynchronize only yields to
    # pytest-timeout's thread method (stack dump + os._exit), which also tears the context down.
    for item in items:
        if unproven(item) and item.get_closest_marker("gpu") is not None:
            item.add_marker(pytest.mark.timeout(300, method="thread"))
