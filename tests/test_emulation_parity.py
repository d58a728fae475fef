"""The data-parallel restatement used by the CUDA kernel (modle_b200/csrc/sim_core.hpp), run on
the CPU through tests/emu, must agree bit for bit with the sequential oracle: LEF trajectories,
ranks, barrier states, RNG draw counts, epochs, band matrix, 1D occupancy."""
import numpy as np
import pytest

import emu_lib
from common import make_case, results_equal
from oracle import pyoracle

CASES = {
    "defaults_small": dict(size=3_000_000, ncells=4, target_contact_density=0.01),
    "large_staging": dict(size=3_000_000, ncells=3, target_contact_density=0.01, _staging=2,
                          _vt=96),
    "mid_staging": dict(size=3_000_000, ncells=2, target_contact_density=0.01, _staging=3,
                        _vt=96),
    "wide_staging": dict(size=3_000_000, ncells=2, target_contact_density=0.01, _staging=4),
    "more_barriers": dict(size=8_000_000, ncells=3, nbar=150, target_contact_density=0.02,
                          _vt=128),
    "no_bypass": dict(size=5_000_000, ncells=3, nbar=90, target_contact_density=0.02,
                      probability_of_extrusion_unit_bypass=0.0),
    "fractional_pblock": dict(size=5_000_000, ncells=3, nbar=90, target_contact_density=0.02,
                              lef_bar_major_collision_pblock=0.8,
                              lef_bar_minor_collision_pblock=0.1),
    "lef_density_x4": dict(size=5_000_000, ncells=2, nbar=300, target_contact_density=0.02,
                           number_of_lefs_per_mbp=80, probability_of_extrusion_unit_bypass=0.01),
    "epochs_criterion": dict(size=5_000_000, ncells=3, nbar=90, stopping_criterion=1,
                             target_simulation_epochs=50),
    "skip_burnin": dict(size=5_000_000, ncells=3, nbar=90, skip_burnin=1,
                        target_contact_density=0.02),
    "loop_only_no_noise": dict(size=4_000_000, ncells=2, nbar=60, target_contact_density=0.01,
                               contact_sampling_strategy=4),
    "tad_only": dict(size=4_000_000, ncells=2, nbar=60, target_contact_density=0.01,
                     contact_sampling_strategy=3),
    "no_1d_track": dict(size=4_000_000, ncells=2, nbar=60, target_contact_density=0.01,
                        track_1d_lef_position=0),
    "sub_interval": dict(size=9_000_000, start=2_000_000, end=6_500_000, ncells=2, nbar=70,
                         target_contact_density=0.02),
    "narrow_band_missed_updates": dict(size=4_000_000, ncells=2, nbar=20, diagonal_width=20_000,
                                       target_contact_density=0.5),
    "constant_speed": dict(size=4_000_000, ncells=2, nbar=60, target_contact_density=0.01,
                           rev_extrusion_speed_std=0.0, fwd_extrusion_speed_std=0.0),
    "max_burnin_forced": dict(size=4_000_000, ncells=2, nbar=60, target_contact_density=0.01,
                              max_burnin_epochs=150),
    "tiny_interval": dict(size=120_000, ncells=3, nbar=3, target_contact_density=0.002),
    "no_barriers": dict(size=2_000_000, ncells=2, nbar=0, target_contact_density=0.01),
    # 480 contacts + 480 1D events per epoch against a scratch area that holds 62 events: the
    # pooled sampling pass runs 16 batches per epoch (plus the rounds its off-stride events end)
    "dense_sampling_many_batches": dict(size=3_000_000, ncells=2, nbar=40,
                                        contact_sampling_interval=1000,
                                        target_contact_density=0.2),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_emulation_matches_oracle(name):
    kw = dict(CASES[name])
    vt = kw.pop("_vt", 64)
    staging = kw.pop("_staging", 0)
    p, iv, bars, tasks = make_case(**kw)
    a = pyoracle.simulate_interval(p, iv, bars, tasks, nthreads=4)
    b = emu_lib.simulate_interval(p, iv, bars, tasks, virtual_threads=vt, staging=staging)
    assert b[2]["device_fault"].max() == 0
    assert results_equal(a, b) == []
    assert int(a[0].sum()) + a[3] == int(a[2]["num_contacts"].sum())


def test_1kb_bins_long_sampling_phase():
    p, iv, bars, tasks = make_case(size=3_000_000, ncells=1, nbar=60, bin_size=1000,
                                   target_contact_density=0.005)
    a = pyoracle.simulate_interval(p, iv, bars, tasks)
    b = emu_lib.simulate_interval(p, iv, bars, tasks, virtual_threads=77)
    assert results_equal(a, b) == []


@pytest.mark.parametrize("epochs", [1, 2, 7, 40, 250])
def test_per_epoch_state_matches_oracle(epochs):
    p, iv, bars, tasks = make_case(size=6_000_000, ncells=1, nbar=110)
    p.debug_max_epochs = epochs
    a = pyoracle.snapshot_cell(p, iv, bars, tasks[0:1])
    b = emu_lib.snapshot_cell(p, iv, bars, tasks[0:1], virtual_threads=33)
    for k in a:
        if isinstance(a[k], np.ndarray):
            assert np.array_equal(a[k], b[k]), k
        else:
            assert a[k] == b[k], k


def test_virtual_cta_width_does_not_matter():
    p, iv, bars, tasks = make_case(size=3_000_000, ncells=2, nbar=50, target_contact_density=0.01)
    ref = emu_lib.simulate_interval(p, iv, bars, tasks, virtual_threads=1)
    for vt in (2, 31, 64, 257):
        assert results_equal(ref, emu_lib.simulate_interval(p, iv, bars, tasks,
                                                           virtual_threads=vt)) == []


def _random_state(rng, n, nb, size, speed):
    rev = rng.integers(0, size, n)
    span = rng.integers(0, max(2, size // max(n, 1)) * 2, n)
    fwd = np.minimum(rev + span, size - 1)
    for i in range(n):
        if rng.random() < 0.1 and i > 0:
            j = rng.integers(0, i)
            if rng.random() < 0.5:
                rev[i] = rev[j]
                fwd[i] = max(fwd[i], rev[i])
            else:
                fwd[i] = fwd[j]
                rev[i] = min(rev[i], fwd[i])
        if rng.random() < 0.03:
            rev[i] = 0
        if rng.random() < 0.03:
            fwd[i] = size - 1
    ep = rng.integers(0, 5, n)
    rr, fr = pyoracle.rank_lefs(rev, fwd, ep, np.arange(n), np.arange(n), init_buffers=True)
    rm = np.maximum(0, np.round(rng.normal(speed, speed * 0.3, n))).astype(np.int64)
    fm = np.maximum(0, np.round(rng.normal(speed, speed * 0.3, n))).astype(np.int64)
    bp = np.sort(rng.choice(size, nb, replace=False)) if nb else np.zeros(0, dtype=np.int64)
    return rev, fwd, ep, rr, fr, rm, fm, bp, rng.integers(1, 3, nb), \
        (rng.random(nb) < 0.8).astype(np.uint8)


ALL_STEPS = ["adjust", "clamp", "boundaries", "lef_bar", "primary", "correct_lef_bar",
             "correct_primary", "secondary", "fix_secondary"]


@pytest.mark.parametrize("bypass,pmaj,pmin", [(0.0, 1.0, 0.0), (0.1, 1.0, 0.0), (0.3, 0.7, 0.2),
                                              (1.0, 1.0, 1.0)])
def test_collision_pipeline_fuzz(bypass, pmaj, pmin):
    """Random dense layouts (ties, units at both interval ends, inactive barriers): the kernel's
    scans must reproduce the sequential pipeline including the number of RNG draws."""
    for seed in range(400):
        rng = np.random.default_rng(seed)
        n = int(rng.integers(1, 60))
        nb = int(rng.integers(0, 30))
        size = int(rng.choice([300, 2000, 20000]))
        st = _random_state(rng, n, nb, size, size / 40)
        kw = dict(prob_bypass=bypass, pblock_major=pmaj, pblock_minor=pmin, rng_seed=seed)
        a = pyoracle.collision_steps(ALL_STEPS, 0, size, *st, **kw)
        b = emu_lib.collision_steps(ALL_STEPS, 0, size, *st, virtual_threads=1 + seed % 9, **kw)
        assert b["fault"] == 0
        for k in ("rev", "fwd", "rr", "fr", "rm", "fm", "rc", "fc", "n5", "n3", "ndraws"):
            assert np.array_equal(a[k], b[k]), (seed, k)


def test_rank_lefs_fuzz_with_ties():
    for seed in range(300):
        rng = np.random.default_rng(1000 + seed)
        n = int(rng.integers(2, 80))
        rev = rng.integers(0, 40, n)
        fwd = rev + rng.integers(0, 10, n)
        ep = rng.integers(0, 3, n)
        rr0 = rng.permutation(n)
        fr0 = rng.permutation(n)
        a = pyoracle.rank_lefs(rev, fwd, ep, rr0, fr0)
        b = emu_lib.rank_lefs(rev, fwd, ep, rr0, fr0)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), seed


@pytest.mark.parametrize("n,vt", [(300, 64), (5000, 512), (10000, 512), (10000, 16), (777, 1)])
@pytest.mark.parametrize("seed", [1, 20260117])
def test_move_generation_stream_order(n, vt, seed):
    """draw_normal_moves (parallel over the raw stream, slow ziggurat paths stitched in) against
    the oracle's sequential Normal draws: same moves, same number of raw draws. The large cases
    hold > 100 slow-path draws, some of them swallowed by an earlier slow path."""
    from modle_b200 import host

    state = host.rng_seed(seed)
    z, _, draws = pyoracle.sample("normal", n, state, 4000.0, 200.0)
    expect = np.floor(np.maximum(z, 0.0) + 0.5).astype(np.uint64)  # == std::round for x >= 0 here
    moves, used = emu_lib.sample_moves(state, n, 4000.0, 200.0, virtual_threads=vt)
    assert used == draws
    assert np.array_equal(moves, expect)


@pytest.mark.parametrize("name", ["defaults_small", "lef_density_x4", "fractional_pblock",
                                  "skip_burnin", "large_staging"])
def test_internal_state_log_of_the_kernel_source_matches_oracle(name):
    """log_epoch_state (sim_core.hpp) against the oracle's dump_stats restatement
    (simulation.cpp:995-1056), record by record, for several virtual CTA widths."""
    kw = dict(CASES[name])
    kw.pop("_vt", None)
    staging = kw.pop("_staging", 0)
    p, iv, bars, tasks = make_case(**kw)
    cap = 500
    a = pyoracle.simulate_interval(p, iv, bars, tasks, nthreads=4, log_capacity_per_cell=cap)
    for vt in (32, 96):
        b = emu_lib.simulate_interval(p, iv, bars, tasks, virtual_threads=vt, staging=staging,
                                      log_capacity_per_cell=cap)
        assert results_equal(a[:4], b[:4]) == []
        assert np.array_equal(a[4], b[4]), vt
    ne = int(a[2]["num_epochs"][0])
    assert a[4][0]["num_lefs"][:min(ne, cap)].any()


@pytest.fixture
def thread_order():
    yield emu_lib.set_thread_order
    emu_lib.set_thread_order(0)


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("name", sorted(CASES))
def test_regions_do_not_depend_on_thread_order(name, mode, thread_order):
    """A region (the code between two CTA barriers) must give the same result whatever order its
    threads run in; otherwise the device needs a barrier the emulation's ascending loop hides.
    Every case is replayed with the virtual threads of each region descending and shuffled."""
    kw = dict(CASES[name])
    vt = kw.pop("_vt", 64)
    staging = kw.pop("_staging", 0)
    p, iv, bars, tasks = make_case(**kw)
    a = pyoracle.simulate_interval(p, iv, bars, tasks, nthreads=4, log_capacity_per_cell=300)
    thread_order(mode)
    b = emu_lib.simulate_interval(p, iv, bars, tasks, virtual_threads=vt, staging=staging,
                                  log_capacity_per_cell=300)
    assert b[2]["device_fault"].max() == 0
    assert results_equal(a[:4], b[:4]) == []
    assert np.array_equal(a[4], b[4])


@pytest.mark.parametrize("mode", [1, 2])
def test_reference_goldens_under_other_thread_orders(mode, thread_order):
    import test_reference_goldens as trg

    thread_order(mode)
    for case in trg.STEP_CASES:
        if case["released"]:
            continue
        for vthreads in (3, 8):
            try:
                trg.test_kernel_emulation_reproduces_reference_goldens(case, vthreads)
            except pytest.skip.Exception:
                pass
    for case in trg.RANK_CASES:
        trg.test_rank_lefs_goldens(case)


@pytest.mark.parametrize("mode", [0, 1], ids=["deterministic", "throughput"])
def test_barriers_outside_the_interval_are_dead_but_draw(mode):
    """A barrier record can overlap a --genomic-intervals range while its midpoint falls outside
    it; the reference keeps such a barrier (genome.cpp:285-294 only asserts): no unit reaches it,
    but it takes its draws, so the cell's trajectory differs from the run without it."""
    p, iv, bars, tasks = make_case(size=9_000_000, start=2_000_000, end=5_000_000, ncells=2,
                                   nbar=40, seed=4, target_contact_density=0.02)
    extra = np.zeros(2, dtype=bars.dtype)
    extra["pos"] = [int(iv.start) - 7, int(iv.end) + 3]
    extra["stp_active"], extra["stp_inactive"] = bars["stp_active"][:2], bars["stp_inactive"][:2]
    extra["blocking_direction"] = [1, 2]
    with_dead = np.concatenate([extra[:1], bars, extra[1:]])
    emu_lib.set_rng_mode(mode)
    try:
        a = emu_lib.simulate_interval(p, iv, with_dead, tasks, virtual_threads=48)
        b = emu_lib.simulate_interval(p, iv, bars, tasks, virtual_threads=48)
    finally:
        emu_lib.set_rng_mode(0)
    assert a[2]["device_fault"].max() == 0
    assert np.array_equal(a[2]["num_contacts"], tasks["num_target_contacts"])
    if mode == 0:
        o = pyoracle.simulate_interval(p, iv, with_dead, tasks, nthreads=2)
        assert results_equal(a, o) == []
        assert results_equal(a, b) != []  # the dead barriers shifted the stream
    else:
        # counter-based draws are keyed by barrier index: the leading dead barrier renumbers them
        assert results_equal(a, b) != []



@pytest.mark.parametrize("name", ["defaults_small", "more_barriers", "lef_density_x4", "sub_interval"])
def test_lef_bar_walk_without_the_barrier_lookup_table(name):
    """The LEF-BAR walk finds the barriers below a unit through a position-bucket table when the
    table fits in shared memory and through a per-thread cursor otherwise: both forms against the
    oracle (the default emulation run uses the table)."""
    kw = dict(CASES[name])
    vt = kw.pop("_vt", 64)
    kw.pop("_staging", 0)
    p, iv, bars, tasks = make_case(**kw)
    a = pyoracle.simulate_interval(p, iv, bars, tasks, nthreads=4)
    emu_lib.lib().emu_set_barrier_lut(0)
    try:
        b = emu_lib.simulate_interval(p, iv, bars, tasks, virtual_threads=vt)
    finally:
        emu_lib.lib().emu_set_barrier_lut(1)
    assert results_equal(a, b) == []


def test_division_by_a_per_interval_constant_is_exact():
    """sim_core.hpp div_u64 (bind, contact sampling: draw / bucket of boost's uniform_int, which the
    reference computes with a plain 64-bit division, uniform_int_distribution.hpp) against Python
    integers: the general reciprocal, and range + 1 standing in for the reciprocal of the bucket."""
    import ctypes as C
    import random

    L = emu_lib.lib()
    L.emu_div_u64.restype = C.c_uint64
    L.emu_div_u64.argtypes = [C.c_uint64] * 3
    L.emu_uniform_int_bucket.restype = C.c_uint64
    L.emu_uniform_int_bucket.argtypes = [C.c_uint64]
    rnd = random.Random(7)
    top = (1 << 64) - 1
    xs = [0, 1, 2, top, top - 1, 1 << 63, (1 << 63) - 1, (1 << 32), (1 << 32) - 1] + \
        [rnd.getrandbits(64) for _ in range(200)] + [rnd.getrandbits(rnd.randrange(1, 65)) for _ in range(200)]
    ds = [1, 2, 3, 5, 7, 1 << 31, (1 << 32) - 1, 1 << 32, (1 << 32) + 1, 1 << 63, (1 << 63) + 1, top, top - 1] + \
        [rnd.getrandbits(rnd.randrange(1, 65)) | 1 for _ in range(100)]
    for d in ds:
        for x in xs + [d - 1, d, d + 1 if d < top else d, (top // d) * d, max((top // d) * d - 1, 0)]:
            assert int(L.emu_div_u64(x, d, 0)) == x // d, (x, d)
    # uniform_int over [0, range]: bucket = (2^64 - 1) / (range + 1) (+1 when the remainder is range)
    # (ranges are LEF counts and distances on a chromosome: below 2^32, where range + 1 is within
    # one unit of the exact reciprocal; the kernel keeps positions in 32 bits)
    ranges = [1, 2, 3, 10, 1288, 4978, 248_956_421, (1 << 28) - 1, (1 << 32) - 2] + \
        [rnd.getrandbits(rnd.randrange(1, 33)) + 1 for _ in range(100)]
    for r in ranges:
        bucket = top // (r + 1) + (1 if top % (r + 1) == r else 0)
        assert int(L.emu_uniform_int_bucket(r)) == bucket
        for x in xs + [bucket - 1, bucket, bucket * r, bucket * (r + 1) - 1 if bucket * (r + 1) - 1 <= top else top]:
            assert int(L.emu_div_u64(x, bucket, r + 1)) == x // bucket, (x, r)


@pytest.mark.parametrize("name", ["defaults_small", "skip_burnin", "sub_interval", "tiny_interval"])
def test_bind_redo_path_gives_the_same_cells(name):
    """A uniform_int rejection while binding (one draw in 2^36 on a human chromosome: a few per
    cent of genome-wide runs see one) sends bind_lefs down a sequential redo that no input can
    trigger in a test; with the emulation's switch every epoch takes it. Nothing was rejected, so
    the result must still be the oracle's."""
    kw = dict(CASES[name])
    vt = kw.pop("_vt", 64)
    staging = kw.pop("_staging", 0)
    p, iv, bars, tasks = make_case(**kw)
    a = pyoracle.simulate_interval(p, iv, bars, tasks, nthreads=4)
    emu_lib.set_force_bind_redo(True)
    try:
        b = emu_lib.simulate_interval(p, iv, bars, tasks, virtual_threads=vt, staging=staging)
    finally:
        emu_lib.set_force_bind_redo(False)
    assert results_equal(a, b) == []
