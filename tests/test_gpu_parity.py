"""`-m gpu`: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded
inputs (bit-exact: every quantity is integer), plus size-independent properties at the full
BASELINE C1 size."""
import ctypes as C

import numpy as np
import pytest

from common import make_case, results_equal
from oracle import pyoracle

pytestmark = pytest.mark.gpu

CASES = {
    "defaults_small": dict(size=3_000_000, ncells=8, target_contact_density=0.01),
    "many_small_cells": dict(size=1_500_000, ncells=700, nbar=25, target_contact_density=0.05),
    "more_barriers": dict(size=8_000_000, ncells=6, nbar=150, target_contact_density=0.02),
    "no_bypass": dict(size=5_000_000, ncells=5, nbar=90, target_contact_density=0.02,
                      probability_of_extrusion_unit_bypass=0.0),
    "always_bypass": dict(size=5_000_000, ncells=3, nbar=90, target_contact_density=0.02,
                          probability_of_extrusion_unit_bypass=1.0),
    "fractional_pblock": dict(size=5_000_000, ncells=5, nbar=90, target_contact_density=0.02,
                              lef_bar_major_collision_pblock=0.8,
                              lef_bar_minor_collision_pblock=0.1),
    "c4_high_collision": dict(size=12_000_000, ncells=4, nbar=800, target_contact_density=0.01,
                              number_of_lefs_per_mbp=80,
                              probability_of_extrusion_unit_bypass=0.01),
    "epochs_criterion": dict(size=5_000_000, ncells=5, nbar=90, stopping_criterion=1,
                             target_simulation_epochs=50),
    "skip_burnin": dict(size=5_000_000, ncells=5, nbar=90, skip_burnin=1,
                        target_contact_density=0.02),
    "loop_only_no_noise": dict(size=4_000_000, ncells=4, nbar=60, target_contact_density=0.01,
                               contact_sampling_strategy=4),
    "tad_only": dict(size=4_000_000, ncells=4, nbar=60, target_contact_density=0.01,
                     contact_sampling_strategy=3),
    "no_1d_track": dict(size=4_000_000, ncells=4, nbar=60, target_contact_density=0.01,
                        track_1d_lef_position=0),
    "sub_interval": dict(size=9_000_000, start=2_000_000, end=6_500_000, ncells=4, nbar=70,
                         target_contact_density=0.02),
    "narrow_band_missed_updates": dict(size=4_000_000, ncells=4, nbar=20, diagonal_width=20_000,
                                       target_contact_density=0.5),
    "constant_speed": dict(size=4_000_000, ncells=4, nbar=60, target_contact_density=0.01,
                           rev_extrusion_speed_std=0.0, fwd_extrusion_speed_std=0.0),
    "max_burnin_forced": dict(size=4_000_000, ncells=4, nbar=60, target_contact_density=0.01,
                              max_burnin_epochs=150),
    "tiny_interval": dict(size=120_000, ncells=5, nbar=3, target_contact_density=0.002),
    "no_barriers": dict(size=2_000_000, ncells=3, nbar=0, target_contact_density=0.01),
    "c5_1kb_bins": dict(size=3_000_000, ncells=2, nbar=60, bin_size=1000,
                        target_contact_density=0.005),
    "c1_chr20_shape": dict(size=64_444_167, ncells=6, nbar=1132, target_contact_density=0.002,
                           name="chr20"),
    # two CTAs per SM (k_simulate_cells<512, 2>); chr8 is the smallest shape that needs <1024, 1>
    "mid_chr13_shape": dict(size=114_364_328, ncells=3, nbar=943, target_contact_density=0.0008,
                            name="chr13"),
    "large_chr8_shape": dict(size=145_138_636, ncells=3, nbar=1772,
                                 target_contact_density=0.0006, name="chr8"),
    "c3_chr1_shape": dict(size=248_956_422, ncells=3, nbar=3518, target_contact_density=0.0004,
                          name="chr1"),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_matches_oracle(gpu_ctx, name):
    p, iv, bars, tasks = make_case(**CASES[name])
    a = pyoracle.simulate_interval(p, iv, bars, tasks, nthreads=8)
    b = gpu_ctx.simulate_interval(p, iv, bars, tasks)
    assert b[2]["device_fault"].max() == 0
    assert results_equal(a, b) == []


@pytest.mark.parametrize("epochs", [1, 3, 60, 400])
def test_cuda_per_epoch_state_matches_oracle(gpu_ctx, epochs):
    """LEF positions, binding epochs, both rank permutations, barrier states and the number of
    raw RNG draws after a fixed number of epochs."""
    p, iv, bars, tasks = make_case(size=20_000_000, ncells=1, nbar=350)
    p.debug_max_epochs = epochs
    a = pyoracle.snapshot_cell(p, iv, bars, tasks[0:1])
    b = gpu_ctx.snapshot_cell(p, iv, bars, tasks[0:1])
    for k in a:
        if isinstance(a[k], np.ndarray):
            assert np.array_equal(a[k], b[k]), k
        else:
            assert a[k] == b[k], k


def test_c1_full_size_properties(gpu_ctx):
    """BASELINE config C1 (chr20 shape, 512 cells, density 1): properties that do not need the
    oracle at full size, plus an oracle check of a sample of cells."""
    p, iv, bars, tasks = make_case(size=64_444_167, ncells=512, nbar=1132, name="chr20")
    band, occ, stats, missed = gpu_ctx.simulate_interval(p, iv, bars, tasks)
    assert stats["device_fault"].max() == 0
    # every cell hits its contact target exactly; the band holds all of them
    assert np.array_equal(stats["num_contacts"], tasks["num_target_contacts"])
    assert int(band.astype(np.uint64).sum()) + missed == int(stats["num_contacts"].sum())
    assert int(stats["num_contacts"].sum()) == 600 * 12889
    assert band[-1] == 0  # the reference's +1 slack element is never written
    # 1D occupancy: two increments per successful event, never more than 2 per sampling event
    assert int(occ.sum()) % 2 == 0 and int(occ.sum()) <= 2 * int(stats["num_contacts"].sum()) + 2 * 206 * 512
    assert (stats["num_epochs"] > stats["num_burnin_epochs"] - 5).all()
    # determinism: a second run is identical
    band2, occ2, stats2, missed2 = gpu_ctx.simulate_interval(p, iv, bars, tasks)
    assert results_equal((band, occ, stats, missed), (band2, occ2, stats2, missed2)) == []
    # independence of the cell batch: cells 100..103 alone give the same per-cell stats
    sub = gpu_ctx.simulate_interval(p, iv, bars, tasks[100:104])
    for f in ("num_contacts", "num_epochs", "num_burnin_epochs", "num_rng_draws"):
        assert np.array_equal(sub[2][f], stats[f][100:104])
    ora = pyoracle.simulate_interval(p, iv, bars, tasks[100:104], nthreads=4)
    assert results_equal(ora, sub) == []


def test_results_accumulate_into_caller_buffers(gpu_ctx):
    p, iv, bars, tasks = make_case(size=3_000_000, ncells=4, target_contact_density=0.01)
    band, occ, stats, missed = gpu_ctx.simulate_interval(p, iv, bars, tasks)
    band2 = band.copy()
    occ2 = occ.copy()
    gpu_ctx.simulate_interval(p, iv, bars, tasks, band=band2, occ1d=occ2)
    assert np.array_equal(band2, 2 * band) and np.array_equal(occ2, 2 * occ)


def test_overwrite_entry_point_ignores_what_the_buffers_held(gpu_ctx):
    """modle_b200_simulate_interval_overwrite: same results as the accumulating call into zeroed
    buffers, whatever the caller's buffers held before."""
    from modle_b200 import abi, host

    p, iv, bars, tasks = make_case(size=3_000_000, ncells=4, target_contact_density=0.01)
    nrows, ncols = host.band_shape(p, 3_000_000)
    band0 = np.zeros(nrows * ncols + 1, dtype=np.uint32)
    occ0 = np.zeros(ncols, dtype=np.uint64)
    ref = gpu_ctx.simulate_interval(p, iv, bars, tasks, band=band0, occ1d=occ0)  # accumulating call
    band = np.full(nrows * ncols + 1, 0xDEADBEEF, dtype=np.uint32)
    occ = np.full(ncols, 12345678901234, dtype=np.uint64)
    _, _, stats_dt = abi.np_dtypes()
    stats = np.zeros(len(tasks), dtype=stats_dt)
    missed = C.c_uint64(777)
    host.check(host.lib().modle_b200_simulate_interval_overwrite(
        gpu_ctx.handle, C.byref(p), C.byref(iv), bars.ctypes.data, len(bars), tasks.ctypes.data,
        len(tasks), band.ctypes.data, occ.ctypes.data, stats.ctypes.data, C.byref(missed)))
    assert results_equal(ref, (band, occ, stats, int(missed.value))) == []
    assert results_equal(ref, gpu_ctx.simulate_interval(p, iv, bars, tasks)) == []  # fresh buffers


def test_register_contacts_kernel(gpu_ctx):
    import torch

    rng = np.random.default_rng(3)
    nrows, ncols = 50, 4000
    n = 1_000_003
    b1 = rng.integers(0, ncols, n, dtype=np.uint32)
    b2 = np.clip(b1.astype(np.int64) + rng.integers(-70, 70, n), 0, ncols - 1).astype(np.uint32)
    i = np.abs(b1.astype(np.int64) - b2.astype(np.int64))
    j = np.maximum(b1, b2).astype(np.int64)
    ok = i < nrows
    expect = np.bincount(j[ok] * nrows + i[ok], minlength=nrows * ncols + 1).astype(np.uint32)
    d1 = torch.from_numpy(b1.view(np.int32)).cuda()
    d2 = torch.from_numpy(b2.view(np.int32)).cuda()
    band = torch.zeros(nrows * ncols + 1, dtype=torch.int32, device="cuda")
    missed = torch.zeros(1, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    gpu_ctx.register_contacts_device(d1.data_ptr(), d2.data_ptr(), n, nrows, ncols,
                                     band.data_ptr(), missed.data_ptr())
    gpu_ctx.synchronize()
    assert np.array_equal(band.cpu().numpy().view(np.uint32), expect)
    assert int(missed.item()) == int((~ok).sum())


def test_unsupported_inputs_fail_loudly(gpu_ctx):
    from modle_b200.host import ModleB200Error

    p, iv, bars, tasks = make_case(size=3_000_000, ncells=2)
    bad = bars.copy()
    bad["pos"][0] = iv.end + 5
    with pytest.raises(ModleB200Error):
        gpu_ctx.simulate_interval(p, iv, bad, tasks)
    unsorted = bars[::-1].copy()
    with pytest.raises(ModleB200Error):
        gpu_ctx.simulate_interval(p, iv, unsorted, tasks)
    # an interval whose per-cell state cannot live in shared memory is refused, not degraded
    p2, iv2, bars2, tasks2 = make_case(size=248_956_422, ncells=1, nbar=100,
                                       number_of_lefs_per_mbp=100)
    with pytest.raises(ModleB200Error) as e:
        gpu_ctx.simulate_interval(p2, iv2, bars2, tasks2)
    assert e.value.code in (-4,)


def _small_genome():
    from modle_b200 import workloads

    sizes = [("chrA", 9_000_000, 160), ("chrB", 4_000_000, 70), ("chrC", 2_500_000, 0),
             ("chrD", 30_000_000, 500), ("chrE", 1_000_000, 12)]
    return [(n, s, 0, s, workloads.synthetic_barrier_records(s, nb, seed=77 + k))
            for k, (n, s, nb) in enumerate(sizes)]


def _oracle_genome(cfg, genome):
    from modle_b200 import abi, host

    out = {}
    p = cfg.params
    for name, size, start, end, recs in genome:
        bars = host.barriers_from_records(recs, p)
        if len(bars) == 0:
            continue
        iv = abi.Interval(size, start, end, host.compute_num_lefs(p, end - start))
        tasks = host.make_cell_tasks(p, name, iv)
        out[name] = pyoracle.simulate_interval(p, iv, bars, tasks, nthreads=8)
    return out


def test_run_simulate_with_worker_threads_matches_oracle(product_lib):
    """The public host path: Simulation.run_simulate with 3 worker contexts (overlapping
    launches on one GPU) gives, per interval, exactly the oracle's matrices; intervals without
    barriers are skipped like the reference does."""
    from modle_b200.simulation import Config, Simulation

    cfg = Config(num_cells=12, target_contact_density=0.01).transform()
    genome = _small_genome()
    sim = Simulation(cfg, genome)
    try:
        for _ in range(2):  # the second call reuses the contexts and replaces the results
            sim.run_simulate(num_workers=3)
    finally:
        sim.close()
    ora = _oracle_genome(cfg, genome)
    for iv in sim.intervals:
        if iv.chrom_name not in ora:
            assert iv.contacts is None
            continue
        band, occ, stats, missed = ora[iv.chrom_name]
        assert np.array_equal(iv.contacts, band), iv.chrom_name
        assert np.array_equal(iv.lef_1d_occupancy, occ), iv.chrom_name
        assert iv.missed_updates == missed
        assert np.array_equal(iv.stats["num_rng_draws"], stats["num_rng_draws"])


def test_device_engine_sharded_runner_matches_oracle(product_lib):
    """modle_b200.distributed on one GPU: an interval split into three cell ranges that run on
    different streams and add into one device band."""
    from modle_b200 import distributed
    from modle_b200.simulation import Config, Simulation

    cfg = Config(num_cells=12, target_contact_density=0.01).transform()
    genome = _small_genome()
    sim = Simulation(cfg, genome)
    S = distributed.Shard
    shards = [S(0, 0, 5, 0, 5.0), S(0, 5, 6, 0, 1.0), S(0, 6, 12, 0, 6.0), S(1, 0, 12, 0, 4.0),
              S(3, 0, 12, 0, 30.0), S(4, 0, 12, 0, 1.0)]
    eng = distributed.DeviceEngine(0, num_streams=3)
    try:
        out = distributed.run_sharded(eng, cfg.params, sim.intervals, 0, 1, None, shards=shards)
        ora = _oracle_genome(cfg, genome)
        for idx, o in out.items():
            band, occ, stats, missed = ora[sim.intervals[idx].chrom_name]
            assert np.array_equal(o["band"].cpu().numpy().view(np.uint32), band)
            assert np.array_equal(o["occ1d"].cpu().numpy().view(np.uint64)[:len(occ)], occ)
            assert int(o["missed"].item()) == missed
            got = np.concatenate(o["stats"])
            assert int(got["num_contacts"].sum()) == int(stats["num_contacts"].sum())
            assert int(got["device_fault"].max()) == 0
    finally:
        eng.close()


def test_phase_cycles_are_reported(gpu_ctx):
    p, iv, bars, tasks = make_case(size=3_000_000, ncells=4, target_contact_density=0.01)
    gpu_ctx.phase_cycles(reset=True)
    gpu_ctx.simulate_interval(p, iv, bars, tasks)
    ph = gpu_ctx.phase_cycles(reset=True)
    assert ph["total"] > 0 and ph["secondary"] > 0 and ph["moves_generate"] > 0
    main = sum(v for k, v in ph.items() if "." not in k and k not in ("total", "rng_refill(nested)"))
    assert 0.5 * ph["total"] < main <= 1.01 * ph["total"]
    assert all(v == 0 for v in gpu_ctx.phase_cycles().values())


def test_register_contacts_binned_path(gpu_ctx):
    """Bands larger than the L2 go through the binned path (count, scatter by 32 MB tile,
    replay); the result is the same histogram."""
    import torch

    rng = np.random.default_rng(5)
    nrows, ncols = 3000, 20_000  # 240 MB band
    n = 60_000_000 // 4 + 77     # dense enough for the binned path (>= a quarter of the pixels)
    b2 = rng.integers(0, ncols, n, dtype=np.int64)
    d = rng.integers(0, nrows + 40, n, dtype=np.int64)  # some pairs fall outside the band
    b1 = np.clip(b2 - d, 0, None)
    swap = rng.random(n) < 0.5
    x1 = np.where(swap, b2, b1).astype(np.uint32)
    x2 = np.where(swap, b1, b2).astype(np.uint32)
    i = np.abs(x1.astype(np.int64) - x2.astype(np.int64))
    j = np.maximum(x1, x2).astype(np.int64)
    ok = i < nrows
    expect = np.bincount(j[ok] * nrows + i[ok], minlength=nrows * ncols + 1).astype(np.uint32)
    d1 = torch.from_numpy(x1.view(np.int32)).cuda()
    d2 = torch.from_numpy(x2.view(np.int32)).cuda()
    band = torch.zeros(nrows * ncols + 1, dtype=torch.int32, device="cuda")
    missed = torch.zeros(1, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    launches0 = gpu_ctx.kernel_launches()
    for _ in range(2):  # twice: the scratch is reused, counts must double
        gpu_ctx.register_contacts_device(d1.data_ptr(), d2.data_ptr(), n, nrows, ncols,
                                         band.data_ptr(), missed.data_ptr())
    gpu_ctx.synchronize()
    torch.cuda.synchronize()
    assert gpu_ctx.kernel_launches() - launches0 == 8  # 4 kernels per binned call
    assert np.array_equal(band.cpu().numpy().view(np.uint32), 2 * expect)
    assert int(missed.item()) == 2 * int((~ok).sum())


@pytest.mark.parametrize("name", ["defaults_small", "c4_high_collision", "fractional_pblock",
                                  "skip_burnin", "mid_chr13_shape"])
def test_internal_state_log_matches_oracle(gpu_ctx, name):
    """the per-epoch quantities of Simulation::dump_stats
    (simulation.cpp:995-1056), record by record."""
    p, iv, bars, tasks = make_case(**CASES[name])
    cap = 700
    a = pyoracle.simulate_interval(p, iv, bars, tasks, nthreads=8, log_capacity_per_cell=cap)
    b = gpu_ctx.simulate_interval(p, iv, bars, tasks, log_capacity_per_cell=cap)
    assert results_equal(a[:4], b[:4]) == []
    assert np.array_equal(a[4], b[4])
    for c in range(len(tasks)):
        ne = min(cap, int(b[2]["num_epochs"][c]))
        rec = b[4][c]
        assert np.array_equal(rec["epoch"][:ne], np.arange(ne))
        assert not rec[ne:].view(np.uint8).any()  # epochs beyond the cell's last stay zero
        assert (rec["lefs_stalled_both"] <= np.minimum(rec["lefs_stalled_rev"],
                                                       rec["lefs_stalled_fwd"])).all()
        assert (rec["barriers_occupied"][:ne] <= len(bars)).all()


def test_barriers_outside_the_interval_are_dead_but_draw(gpu_ctx):
    """A barrier whose midpoint falls outside a --genomic-intervals range is kept by the reference
    (genome.cpp:285-294): unreachable, but it takes its draws. GPU == oracle with two of them."""
    p, iv, bars, tasks = make_case(size=9_000_000, start=2_000_000, end=5_000_000, ncells=4,
                                   nbar=40, seed=4, target_contact_density=0.02)
    extra = np.zeros(2, dtype=bars.dtype)
    extra["pos"] = [int(iv.start) - 7, int(iv.end) + 3]
    extra["stp_active"], extra["stp_inactive"] = bars["stp_active"][:2], bars["stp_inactive"][:2]
    extra["blocking_direction"] = [1, 2]
    with_dead = np.concatenate([extra[:1], bars, extra[1:]])
    a = gpu_ctx.simulate_interval(p, iv, with_dead, tasks)
    o = pyoracle.simulate_interval(p, iv, with_dead, tasks, nthreads=4)
    assert results_equal(a, o) == []
    assert results_equal(a, gpu_ctx.simulate_interval(p, iv, bars, tasks)) != []
