"""modle_b200 -- B200-native replacement for libmodle's per-(interval, cell) simulation loop.

The compute path is hand-written CUDA (sm_100a) behind a C ABI (include/modle_b200.h, built into
modle_b200/libmodle_b200.so by modle_b200/build.py). This package is the Python host side above
that ABI: it mirrors the slice of modle::Simulation that drives the hot path
(reference: src/libmodle/cpu/scheduler_simulate.cpp:43-170). There is no CPU fallback.
"""
from . import abi  # noqa: F401
from .host import (band_shape, compute_contacts_per_epoch, compute_num_lefs,  # noqa: F401
                   default_params, interval_hash, lib, make_cell_tasks, transform_params)
from .simulation import Config, Context, GenomicInterval, Simulation  # noqa: F401
