"""The C ABI through a COMPILED consumer (tests/cabi/consumer.cpp includes include/modle_b200.h and
links the library, like a MoDLE binding would): struct layout against the ctypes mirror the other
tests go through (a drift between the header and modle_b200/abi.py fails here), and -- on a GPU --
the same small run through both paths, compared by checksum."""
import ctypes as C
import json
import subprocess

import numpy as np
import pytest

from modle_b200 import abi


@pytest.fixture(scope="module")
def consumer(product_lib):
    import cabi_build

    return cabi_build.build_consumer()


MIRRORS = {
    "modle_b200_sim_params": abi.SimParams, "modle_b200_interval": abi.Interval,
    "modle_b200_barrier": abi.Barrier, "modle_b200_cell_task": abi.CellTask,
    "modle_b200_cell_stats": abi.CellStats, "modle_b200_cell_snapshot": abi.CellSnapshot,
    "modle_b200_pixel": abi.Pixel, "modle_b200_shard": abi.ShardRecord,
}


def test_header_layout_equals_the_ctypes_mirror(consumer):
    out = subprocess.run([consumer, "layout"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    lay = json.loads(out.stdout)
    assert lay["abi_version"] == 1
    from modle_b200 import host

    assert lay["num_phases"] == len(host.PHASE_NAMES)
    for name, mirror in MIRRORS.items():
        assert lay[name]["sizeof"] == C.sizeof(mirror), name
        fields = lay[name]["fields"]
        assert list(fields) == [f for f, _ in mirror._fields_], name  # same names, same order
        for f, _ in mirror._fields_:
            assert fields[f] == getattr(mirror, f).offset, (name, f)
    rec = abi.epoch_record_dtype()
    assert lay["modle_b200_epoch_record"]["sizeof"] == rec.itemsize
    for f, off in lay["modle_b200_epoch_record"]["fields"].items():
        assert rec.fields[f][1] == off, f
    barrier_dt, task_dt, stats_dt = abi.np_dtypes()
    for dt, name in ((barrier_dt, "modle_b200_barrier"), (task_dt, "modle_b200_cell_task"),
                     (stats_dt, "modle_b200_cell_stats"), (abi.pixel_dtype(), "modle_b200_pixel"),
                     (abi.shard_dtype(), "modle_b200_shard")):
        for f, off in lay[name]["fields"].items():
            assert dt.fields[f][1] == off, (name, f)


def _fnv1a(a):
    h = 1469598103934665603
    for b in a.tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def _python_side(cells):
    """The inputs consumer.cpp builds, restated through the ctypes path."""
    from modle_b200 import host

    p = host.default_params()
    p.num_cells = cells
    p.seed = 42
    p.target_contact_density = 0.02
    host.transform_params(p)
    iv = abi.Interval(6_000_000, 0, 6_000_000, host.compute_num_lefs(p, 6_000_000))
    barrier_dt, _, _ = abi.np_dtypes()
    pos = list(range(50021, 6_000_000, 97003))
    bars = np.zeros(len(pos), dtype=barrier_dt)
    for k, x in enumerate(pos):
        bars[k] = (x, host.lib().modle_b200_stp_active_from_occupancy(
            p.barrier_not_occupied_stp, 0.70 + 0.01 * float(k % 25)), p.barrier_not_occupied_stp,
            abi.DIR_REV if k % 2 == 0 else abi.DIR_FWD, 0)
    tasks = host.make_cell_tasks(p, "chrCabi", iv)
    return p, iv, bars, tasks


@pytest.mark.gpu
def test_compiled_consumer_equals_the_ctypes_path_and_the_oracle(consumer, gpu_ctx):
    from modle_b200 import host
    from oracle import pyoracle

    cells = 6
    out = subprocess.run([consumer, "run", str(cells)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    got = json.loads(out.stdout)
    p, iv, bars, tasks = _python_side(cells)
    band, occ, stats, missed = gpu_ctx.simulate_interval(p, iv, bars, tasks)
    nrows, ncols = host.band_shape(p, 6_000_000)
    pixels = gpu_ctx.band_to_pixels(band, nrows, ncols, bin_offset=1000)
    expect = dict(num_lefs=int(iv.num_lefs), num_barriers=len(bars), nrows=nrows, ncols=ncols,
                  band_sum=int(band.astype(np.uint64).sum()), band_hash=_fnv1a(band),
                  occ_hash=_fnv1a(occ), missed=missed, contacts=int(stats["num_contacts"].sum()),
                  epochs=int(stats["num_epochs"].sum()), rng_draws=int(stats["num_rng_draws"].sum()),
                  faults=0, num_pixels=len(pixels), pixel_count_sum=int(pixels["count"].sum()),
                  pixel_hash=_fnv1a(pixels))
    for k, v in expect.items():
        assert got[k] == v, k
    assert got["kernel_launches"] >= 3  # simulate + count + fill
    ora = pyoracle.simulate_interval(p, iv, bars, tasks, nthreads=4)
    assert np.array_equal(ora[0], band) and np.array_equal(ora[1], occ) and ora[3] == missed
