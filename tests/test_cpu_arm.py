"""The CPU arm of bench.py: oracle-side parameter derivation (oracle/pyparams.py) against the
product's host layer, and the one-queue scheduler (oracle_simulate_genome) against the
per-interval oracle call. No GPU."""
import ctypes as C
import itertools
import os
import subprocess
import sys

import numpy as np
import pytest

from modle_b200 import abi, host, workloads
from oracle import pyoracle, pyparams

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _same(p, q):
    return [f for f, _ in p._fields_ if getattr(p, f) != getattr(q, f) and
            not (getattr(p, f) != getattr(p, f) and getattr(q, f) != getattr(q, f))]


def test_default_params_match_the_host_layer(product_lib):
    assert _same(host.default_params(), pyparams.default_params()) == []


@pytest.mark.parametrize("overrides", [
    {}, {"bin_size": 1000}, {"bin_size": 1000, "diagonal_width": 3_000_000},
    {"rev_extrusion_speed": 3000}, {"fwd_extrusion_speed": 0, "rev_extrusion_speed": 8000},
    {"extrusion_barrier_occupancy": 0.9}, {"extrusion_barrier_occupancy": 0.0},
    {"bin_size": 2000, "extrusion_barrier_occupancy": 0.7, "probability_of_extrusion_unit_bypass": 0.3},
    {"contact_sampling_strategy": abi.SAMPLE_LOOP | abi.SAMPLE_NOISIFY},
    {"contact_sampling_strategy": abi.SAMPLE_TAD},
    {"stopping_criterion": abi.STOP_SIMULATION_EPOCHS, "target_simulation_epochs": 50},
    {"burnin_speed_coefficient": 2.5, "max_burnin_epochs": 100},
    {"bin_size": 10_000, "lef_bar_major_collision_pblock": 0.8, "lef_bar_minor_collision_pblock": 0.1},
    {"normalize_probabilities": 0, "bin_size": 1000},
    {"rev_extrusion_speed_std": 250.0, "fwd_extrusion_speed_std": 0.0},
    {"number_of_lefs_per_mbp": 80.0, "probability_of_extrusion_unit_bypass": 0.01},
])
def test_transform_params_matches_the_host_layer(product_lib, overrides):
    p = host.default_params()
    for k, v in overrides.items():
        setattr(p, k, v)
    host.transform_params(p, "rev_extrusion_speed" in overrides, "fwd_extrusion_speed" in overrides,
                          "extrusion_barrier_occupancy" in overrides)
    q = pyparams.make_params(**overrides)
    assert _same(p, q) == []
    for size in (1, 4999, 5000, 64_444_167, 248_956_422):
        assert pyparams.compute_num_lefs(q, size) == host.compute_num_lefs(p, size)
        assert pyparams.band_shape(q, size) == host.band_shape(p, size)


def test_barriers_and_tasks_match_the_host_layer(product_lib):
    for ov in ({}, {"extrusion_barrier_occupancy": 0.85}, {"bin_size": 1000}):
        p = pyparams.make_params(num_cells=7, **ov)
        recs = workloads.synthetic_barrier_records(5_000_000, 60, seed=5) + [(123, "+", 0.0)]
        assert np.array_equal(pyparams.barriers_from_records(recs, p),
                              host.barriers_from_records(recs, p))
        _, jobs = pyoracle.genome_jobs(dict(num_cells=7, **ov), [("chrQ", 5_000_000, 0, 5_000_000, recs),
                                                                 ("chrEmpty", 1_000_000, 0, 1_000_000, [])])
        assert len(jobs) == 1  # intervals without barriers are skipped
        iv, bars, tasks = jobs[0]
        assert np.array_equal(tasks, host.make_cell_tasks(p, "chrQ", iv))


def test_one_queue_scheduler_equals_per_interval_runs():
    ov = dict(num_cells=5, target_contact_density=0.02)
    genome = [("chrA", 4_000_000, 0, 4_000_000, workloads.synthetic_barrier_records(4_000_000, 50, seed=1)),
              ("chrB", 2_500_000, 500_000, 2_500_000,
               [r for r in workloads.synthetic_barrier_records(2_500_000, 30, seed=2) if r[0] >= 500_000]),
              ("chrC", 6_000_000, 0, 6_000_000, workloads.synthetic_barrier_records(6_000_000, 70, seed=3))]
    p, jobs = pyoracle.genome_jobs(ov, genome)
    for nthreads in (1, 3, 8):
        res = pyoracle.simulate_genome(p, jobs, nthreads=nthreads)
        for (iv, bars, tasks), (band, occ, st, missed) in zip(jobs, res):
            b2, o2, s2, m2 = pyoracle.simulate_interval(p, iv, bars, tasks, nthreads=2)
            assert np.array_equal(band, b2) and np.array_equal(occ, o2) and missed == m2
            assert np.array_equal(st, s2)


def test_sample_plan_gives_every_thread_a_queue():
    sys.path.insert(0, ROOT)
    import bench

    ov, genome = workloads.spec("c2", 512)
    for cores in (8, 16, 32, 64):
        cpi = bench.cpu_sample_plan(ov, genome, cores, 12.0)
        assert cpi * len(genome) >= bench.CPU_MIN_CELLS_PER_THREAD * cores
        assert cpi <= 512


def test_reference_arm_does_not_load_the_product_library():
    """bench.py --impl reference builds its inputs with oracle-side code only."""
    code = ("import sys, os; sys.path.insert(0, %r); import bench; from modle_b200 import workloads;"
            "ov, g = workloads.spec('c1', 4); ov['target_contact_density'] = 0.001;"
            "from oracle import pyoracle; p, jobs = pyoracle.genome_jobs(ov, g, 2);"
            "pyoracle.simulate_genome(p, jobs, nthreads=2);"
            "maps = open('/proc/self/maps').read();"
            "assert 'liboracle.so' in maps; assert 'libmodle_b200' not in maps; print('ok')") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr
