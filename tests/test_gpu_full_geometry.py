"""`-m gpu`: BASELINE configs C4 and C5 at their REAL geometry (the kernels, shared-memory
footprints and band sizes the bench runs), a few cells each at a low contact density, bit-exact
against the CPU oracle; and repeated runs of C1 at full size (determinism on the device).

  C4  chr20 size, 80 LEFs/Mbp -> N = 5156 LEFs, 4296 synthetic barriers (1 / 15 kb), bypass 0.01:
      the largest shared-memory footprint (212 KB, k_simulate_cells<1024, 1>), the regime that
      stresses rank_lefs and the collision scans.
  C5  chr2 at 1 kb bins, 3 Mbp diagonal width: 3000 x 242,194 pixels = 2.9 GB band (HBM-bound
      scatter; the band must allocate and come back whole), N = 4844 LEFs, ~1,000 burn-in epochs.
"""
import numpy as np
import pytest

from common import results_equal
from modle_b200 import abi, host, workloads
from oracle import pyoracle

pytestmark = pytest.mark.gpu


def _inputs(name, ncells, **more):
    cfg, genome = getattr(workloads, "config_" + name)(ncells, **more)
    p = cfg.params
    chrom, size, start, end, recs = genome[0]
    bars = host.barriers_from_records(recs, p)
    iv = abi.Interval(size, start, end, host.compute_num_lefs(p, end - start))
    tasks = host.make_cell_tasks(p, chrom, iv)
    return p, iv, bars, tasks


def test_c4_full_geometry_matches_oracle(gpu_ctx):
    p, iv, bars, tasks = _inputs("c4", 3, target_contact_density=0.002)
    assert int(iv.num_lefs) == 5156 and len(bars) == 4296
    threads, per_sm, smem = host.launch_geometry(int(iv.num_lefs), len(bars))
    assert (threads, per_sm) == (1024, 1) and smem > 200_000
    a = pyoracle.simulate_interval(p, iv, bars, tasks, nthreads=3)
    b = gpu_ctx.simulate_interval(p, iv, bars, tasks)
    assert b[2]["device_fault"].max() == 0
    assert results_equal(a, b) == []
    assert (b[2]["num_contacts"] == tasks["num_target_contacts"]).all()


def test_c4_full_geometry_per_epoch_state(gpu_ctx):
    p, iv, bars, tasks = _inputs("c4", 1, target_contact_density=0.002)
    for epochs in (2, 250):
        p.debug_max_epochs = epochs
        a = pyoracle.snapshot_cell(p, iv, bars, tasks[0:1])
        b = gpu_ctx.snapshot_cell(p, iv, bars, tasks[0:1])
        for k in a:
            if isinstance(a[k], np.ndarray):
                assert np.array_equal(a[k], b[k]), (epochs, k)
            else:
                assert a[k] == b[k], (epochs, k)


def test_c5_full_geometry_matches_oracle(gpu_ctx):
    p, iv, bars, tasks = _inputs("c5", 2, target_contact_density=1e-4)
    nrows, ncols = host.band_shape(p, int(iv.end - iv.start))
    assert (nrows, ncols) == (3000, 242_194) and int(iv.num_lefs) == 4844
    assert (nrows * ncols + 1) * 4 > 2.9e9  # the 2.9 GB band
    a = pyoracle.simulate_interval(p, iv, bars, tasks, nthreads=2)
    b = gpu_ctx.simulate_interval(p, iv, bars, tasks)
    assert b[2]["device_fault"].max() == 0
    assert results_equal(a, b) == []
    assert int(b[0].astype(np.uint64).sum()) + b[3] == int(tasks["num_target_contacts"].sum())
    # the renormalised probabilities of the 1 kb configuration are in effect
    assert abs(p.probability_of_extrusion_unit_bypass - 0.02) < 1e-12
    assert (b[2]["num_burnin_epochs"] > 900).all()  # activation alone takes 937 epochs


def test_c1_repeated_runs_are_identical(gpu_ctx):
    """Ten runs of the full C1 workload (512 cells): per-cell epochs / burn-in epochs / contacts /
    raw draw counts and the band checksum must not move (a missing CTA barrier shows up as a
    cell whose burn-in length differs in some runs: round 1's flake hit 1-2 cells in ~35 % of
    runs)."""
    p, iv, bars, tasks = _inputs("c1", 512)
    ref = None
    for rep in range(10):
        band, occ, stats, missed = gpu_ctx.simulate_interval(p, iv, bars, tasks)
        assert stats["device_fault"].max() == 0
        key = (stats["num_epochs"].copy(), stats["num_burnin_epochs"].copy(),
               stats["num_contacts"].copy(), stats["num_rng_draws"].copy(),
               int((band.astype(np.uint64) * (np.arange(band.size, dtype=np.uint64) % 65521 + 1)).sum()),
               int(occ.sum()), missed)
        if ref is None:
            ref = key
            continue
        for x, y in zip(ref, key):
            assert np.array_equal(x, y), f"run {rep} differs from run 0"


def test_c1_repeated_runs_are_identical_throughput_mode(product_lib):
    from modle_b200.simulation import Context

    p, iv, bars, tasks = _inputs("c1", 512)
    ctx = Context(0, rng_mode=1)
    try:
        ref = None
        for rep in range(5):
            band, occ, stats, missed = ctx.simulate_interval(p, iv, bars, tasks)
            assert stats["device_fault"].max() == 0
            key = (stats["num_epochs"].copy(), stats["num_burnin_epochs"].copy(),
                   stats["num_contacts"].copy(), int(band.astype(np.uint64).sum()),
                   int((band.astype(np.uint64) * (np.arange(band.size, dtype=np.uint64) % 65521 + 1)).sum()),
                   int(occ.sum()), missed)
            if ref is None:
                ref = key
                continue
            for x, y in zip(ref, key):
                assert np.array_equal(x, y), f"run {rep} differs from run 0"
    finally:
        ctx.close()
