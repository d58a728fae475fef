"""Throughput mode (MODLE_B200_RNG_COUNTER, sim_core.hpp CellSimT<true>) on the CPU, through the
emulation of the kernel source: the counter-based draw function against an independent numpy
restatement and the SplitMix64 test vector, the properties the mode promises (a cell's result is a
pure function of its task -- not of the CTA width or of the order threads run in), the invariants
of the simulation, and the statistical gate of SURVEY 8c(ii) against the CPU oracle (the same
tolerances as tests/test_statistical_parity.py, which were calibrated on the oracle's own
seed-to-seed variability)."""
import ctypes as C
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import emu_lib
from common import make_case, results_equal
from modle_b200 import host
from oracle import pyoracle
from stats_eval import per_diagonal_mean_var, stratum_adjusted_correlation, stripe_pearson

M64 = (1 << 64) - 1
GOLDEN = 0x9E3779B97F4A7C15


def mix64(z):
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def ctr_draw(state, epoch, phase, item, k):
    """Independent restatement of sim_core.hpp raw() in throughput mode."""
    key1 = mix64(state[0] ^ mix64(state[1]))
    key2 = mix64(state[2] ^ mix64(state[3]))
    ctr = (((epoch << 40) & M64) | ((phase & 15) << 36) | ((item & 0x0FFFFFFF) << 8)) + k
    return mix64((((ctr ^ key1) * GOLDEN) + key2) & M64)


@pytest.fixture
def throughput():
    emu_lib.set_rng_mode(1)
    yield
    emu_lib.set_rng_mode(0)
    emu_lib.set_thread_order(0)


def _emu_draw(state, epoch, phase, item, k):
    L = emu_lib.lib()
    L.emu_ctr_draw.restype = C.c_uint64
    L.emu_ctr_draw.argtypes = [C.POINTER(C.c_uint64), C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
    st = (C.c_uint64 * 4)(*state)
    return int(L.emu_ctr_draw(st, epoch, phase, item, k))


def test_mixer_is_splitmix64():
    # SplitMix64 seeded with 0: outputs mix64(golden), mix64(2 * golden), ...
    L = emu_lib.lib()
    L.emu_mix64.restype = C.c_uint64
    L.emu_mix64.argtypes = [C.c_uint64]
    expect = [0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4, 0x06C45D188009454F, 0xF88BB8A8724C81EC]
    for i, e in enumerate(expect):
        x = ((i + 1) * GOLDEN) & M64
        assert int(L.emu_mix64(x)) == e
        assert mix64(x) == e


def test_counter_draws_match_the_independent_restatement():
    rng = np.random.default_rng(5)
    for _ in range(200):
        state = [int(x) for x in rng.integers(0, 1 << 63, 4)]
        epoch, phase = int(rng.integers(0, 5000)), int(rng.integers(0, 13))
        item, k = int(rng.integers(0, 1 << 20)), int(rng.integers(0, 255))
        assert _emu_draw(state, epoch, phase, item, k) == ctr_draw(state, epoch, phase, item, k)


def test_counter_draws_are_uniform_and_uncorrelated_along_every_counter_field():
    state = host.rng_seed(12345)
    n = 20000
    axes = {
        "item": [ctr_draw(state, 7, 7, i, 0) for i in range(n)],
        "epoch": [ctr_draw(state, e, 7, 3, 0) for e in range(n)],
        "draw": [ctr_draw(state, 7, 7, i // 200, i % 200) for i in range(n)],
        "phase x item": [ctr_draw(state, 9, i % 13, i // 13, 0) for i in range(n)],
    }
    for name, xs in axes.items():
        u = np.array(xs, dtype=np.float64) / 2.0 ** 64
        assert abs(u.mean() - 0.5) < 4 / np.sqrt(12 * n), name
        assert abs(u.var() - 1 / 12) < 0.004, name
        for lag in (1, 2, 13, 256):
            r = np.corrcoef(u[:-lag], u[lag:])[0, 1]
            assert abs(r) < 4 / np.sqrt(n), (name, lag, r)
        bits = np.array([[(x >> b) & 1 for b in range(0, 64, 7)] for x in xs[:4000]])
        assert np.all(np.abs(bits.mean(axis=0) - 0.5) < 4 * 0.5 / np.sqrt(4000)), name
    # two cells (states one jump() apart, as the reference seeds them) share nothing
    other = host.rng_jump(list(state))
    a = np.array([ctr_draw(state, 7, 7, i, 0) for i in range(n)], dtype=np.float64)
    b = np.array([ctr_draw(other, 7, 7, i, 0) for i in range(n)], dtype=np.float64)
    assert abs(np.corrcoef(a, b)[0, 1]) < 4 / np.sqrt(n)


def test_moves_follow_the_reference_distribution(throughput):
    # generate_moves_helper (simulation.cpp:272-297): round(max(0, Normal(speed, sd)))
    mv = np.concatenate([
        emu_lib.sample_moves(host.rng_seed(99 + s), 6000, 4000.0, 200.0, virtual_threads=256,
                             staging=2)[0] for s in range(10)]).astype(np.float64)
    assert abs(mv.mean() - 4000.0) < 4 * 200.0 / np.sqrt(len(mv))
    assert abs(mv.std() - 200.0) < 3.0
    # tails beyond 3.4426 sigma come from the ziggurat's slow paths
    z = (mv - 4000.0) / 200.0
    assert 5 <= (np.abs(z) > 3.4426).sum() <= 80      # expected 34.6
    from scipy.stats import kstest
    assert kstest(z + np.random.default_rng(0).uniform(-0.5, 0.5, len(z)) / 200.0, "norm").pvalue > 1e-3


CASES = {
    "defaults": dict(size=3_000_000, ncells=2, nbar=40),
    "frac_pblock_bypass": dict(size=2_000_000, ncells=2, nbar=60, lef_bar_major_collision_pblock=0.7,
                               lef_bar_minor_collision_pblock=0.2,
                               probability_of_extrusion_unit_bypass=0.3),
    "no_bypass": dict(size=2_000_000, ncells=2, nbar=30, probability_of_extrusion_unit_bypass=0.0),
    "always_bypass": dict(size=2_000_000, ncells=2, nbar=30, probability_of_extrusion_unit_bypass=1.0),
    "high_collision": dict(size=1_500_000, ncells=2, nbar=100, number_of_lefs_per_mbp=80.0,
                           probability_of_extrusion_unit_bypass=0.01),
    "epochs_skip_burnin": dict(size=2_000_000, ncells=2, nbar=30, skip_burnin=1,
                               stopping_criterion=1, target_simulation_epochs=150),
    "loop_only": dict(size=2_000_000, ncells=2, nbar=30, tad_to_loop_contact_ratio=0.0),
    "no_barriers": dict(size=1_000_000, ncells=2, nbar=0),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_result_depends_on_the_task_only(throughput, name):
    kw = dict(CASES[name])
    kw.setdefault("target_contact_density", 0.05)  # ~100 sampling epochs after the burn-in
    p, iv, bars, tasks = make_case(seed=11, **kw)
    emu_lib.set_thread_order(0)
    ref = emu_lib.simulate_interval(p, iv, bars, tasks, virtual_threads=64)
    st = ref[2]
    assert st["device_fault"].max() == 0
    assert st["num_rng_draws"].max() == 0
    if p.stopping_criterion == 0:
        assert np.array_equal(st["num_contacts"], tasks["num_target_contacts"])
        assert int(ref[0].sum()) + ref[3] == int(tasks["num_target_contacts"].sum())
    else:
        assert np.all(st["num_epochs"] - st["num_burnin_epochs"] == tasks["num_target_epochs"])
    for order, threads in ((0, 256), (1, 96), (2, 33)):
        emu_lib.set_thread_order(order)
        got = emu_lib.simulate_interval(p, iv, bars, tasks, virtual_threads=threads)
        assert results_equal(ref, got) == [], (order, threads)
    # cells are simulated independently: cell 1 alone equals cell 1 of the batch
    emu_lib.set_thread_order(0)
    one = emu_lib.simulate_interval(p, iv, bars, tasks[1:2], virtual_threads=64)
    assert one[2]["num_epochs"][0] == st["num_epochs"][1]
    assert one[2]["num_lef_updates"][0] == st["num_lef_updates"][1]
    # and it is not the deterministic mode's trajectory
    emu_lib.set_rng_mode(0)
    det = emu_lib.simulate_interval(p, iv, bars, tasks, virtual_threads=64)
    emu_lib.set_rng_mode(1)
    assert results_equal(ref, det) != []


def _emu_parallel(p, iv, bars, tasks, nthreads=8):
    """Cells spread over host threads (the emulation keeps its mode per thread)."""
    def part(chunk):
        emu_lib.set_rng_mode(1)
        return emu_lib.simulate_interval(p, iv, bars, chunk, virtual_threads=128)
    chunks = [c for c in np.array_split(tasks, nthreads) if len(c)]
    with ThreadPoolExecutor(len(chunks)) as ex:
        parts = list(ex.map(part, chunks))
    band = sum(x[0].astype(np.uint64) for x in parts).astype(np.uint32)
    occ = sum(x[1] for x in parts)
    return band, occ, np.concatenate([x[2] for x in parts]), sum(x[3] for x in parts)


REGIMES = {
    "defaults": dict(),
    # BASELINE C4: LEF density x4, dense barriers, bypass ~ 0 (rank + collision handling stressed)
    "high_collision": dict(number_of_lefs_per_mbp=80.0, nbar=1300,
                           probability_of_extrusion_unit_bypass=0.01),
    # every Bernoulli of the collision pipeline fractional; TAD-heavy sampling
    "fractional_everything": dict(lef_bar_major_collision_pblock=0.7,
                                  lef_bar_minor_collision_pblock=0.2,
                                  probability_of_extrusion_unit_bypass=0.3,
                                  tad_to_loop_contact_ratio=12.0),
}


@pytest.mark.parametrize("regime", sorted(REGIMES))
def test_statistically_equivalent_to_the_oracle(throughput, regime):
    """Gate (ii) of SURVEY 8c with the tolerances of tests/test_statistical_parity.py."""
    from scipy.stats import ks_2samp

    kw = dict(size=20_000_000, ncells=128, nbar=350, target_contact_density=1.0, name="chrS")
    kw.update(REGIMES[regime])
    runs = {}
    for seed in (1, 2, 3):
        p, iv, bars, tasks = make_case(seed=7, **kw)
        p.seed = seed
        runs[seed] = (p, iv, bars, host.make_cell_tasks(p, "chrS", iv))
    nrows, ncols = host.band_shape(runs[1][0], 20_000_000)
    thr = _emu_parallel(*runs[1])
    ora2 = pyoracle.simulate_interval(*runs[2], nthreads=8)
    ora3 = pyoracle.simulate_interval(*runs[3], nthreads=8)
    assert thr[2]["device_fault"].max() == 0
    assert int(thr[0].sum()) + thr[3] == int(ora2[0].sum()) + ora2[3]  # (+ out-of-band updates)

    m_t, v_t, tot_t = per_diagonal_mean_var(thr[0], nrows, ncols)
    m_o, v_o, tot_o = per_diagonal_mean_var(ora2[0], nrows, ncols)
    big = (tot_t >= 1e4) & (tot_o >= 1e4)
    assert big.sum() >= 10
    assert np.all(np.abs(m_t[big] / m_o[big] - 1.0) < 0.05)
    assert np.all(np.abs(v_t[big] / v_o[big] - 1.0) < 0.20)

    scc_to = stratum_adjusted_correlation(thr[0], ora2[0], nrows, ncols, max_d=200)
    scc_oo = stratum_adjusted_correlation(ora2[0], ora3[0], nrows, ncols, max_d=200)
    assert scc_to > scc_oo - 0.01, (scc_to, scc_oo)
    if regime == "defaults":  # (the other regimes leave less structure above the sampling noise)
        assert scc_to > 0.4
    # per-stripe Pearson as `modle_tools eval` computes it: medians within 0.005 of the floor
    for direction in ("vertical", "horizontal"):
        r_to = np.nanmedian(stripe_pearson(thr[0], ora2[0], nrows, ncols, direction))
        r_oo = np.nanmedian(stripe_pearson(ora2[0], ora3[0], nrows, ncols, direction))
        assert r_to > r_oo - 0.005, (direction, r_to, r_oo)

    assert ks_2samp(thr[2]["num_burnin_epochs"], ora2[2]["num_burnin_epochs"]).pvalue > 0.001
    assert ks_2samp(thr[2]["num_epochs"], ora2[2]["num_epochs"]).pvalue > 0.001
    # 1D occupancy track: same total mass per registered event, same profile
    assert abs(int(thr[1].sum()) / int(ora2[1].sum()) - 1.0) < 0.01
    r = np.corrcoef(thr[1].astype(np.float64), ora2[1].astype(np.float64))[0, 1]
    r_oo = np.corrcoef(ora3[1].astype(np.float64), ora2[1].astype(np.float64))[0, 1]
    assert r > r_oo - 0.02, (r, r_oo)


def test_stripe_pearson_matches_a_pixel_by_pixel_loop():
    rng = np.random.default_rng(3)
    nrows, ncols = 5, 23
    b1 = rng.integers(0, 50, nrows * ncols + 1).astype(np.uint32)
    b2 = rng.integers(0, 50, nrows * ncols + 1).astype(np.uint32)
    for b in (b1, b2):  # pixels above the first row of the matrix do not exist
        for j in range(ncols):
            for d in range(nrows):
                if d > j:
                    b[j * nrows + d] = 0

    def get(b, i, j):  # ContactMatrixDense::unsafe_get (contact_matrix_dense_impl.hpp)
        if i > j:
            i, j = j, i
        return float(b[j * nrows + (j - i)]) if j - i < nrows and j < ncols else 0.0

    for direction in ("vertical", "horizontal"):
        got = stripe_pearson(b1, b2, nrows, ncols, direction)
        for i in range(ncols):
            if direction == "vertical":
                x = [get(b1, i - d, i) if i >= d else 0.0 for d in range(nrows)]
                y = [get(b2, i - d, i) if i >= d else 0.0 for d in range(nrows)]
            else:
                x = [get(b1, i, i + d) for d in range(nrows)]
                y = [get(b2, i, i + d) for d in range(nrows)]
            if np.std(x) == 0 or np.std(y) == 0:
                assert np.isnan(got[i])
            else:
                assert abs(got[i] - np.corrcoef(x, y)[0, 1]) < 1e-12


def test_internal_state_statistics_match_the_oracle(throughput):
    """The per-epoch quantities of Simulation::dump_stats (collisions by kind, stalled units,
    occupied barriers, mean loop size) are much more sensitive to a mis-wired trial probability
    than the contact matrix is: compare their per-cell averages over the epochs 300-599 between
    the throughput mode and the oracle (independent seeds; every collision Bernoulli fractional),
    against the cell-to-cell spread (4 standard errors)."""
    kw = dict(size=10_000_000, ncells=48, nbar=180, name="chrL", target_contact_density=20.0,
              lef_bar_major_collision_pblock=0.7, lef_bar_minor_collision_pblock=0.2,
              probability_of_extrusion_unit_bypass=0.3)
    cap = 600
    p, iv, bars, _ = make_case(seed=7, **kw)
    p.seed = 1
    t_thr = host.make_cell_tasks(p, "chrL", iv)

    def part(chunk):
        emu_lib.set_rng_mode(1)
        return emu_lib.simulate_interval(p, iv, bars, chunk, virtual_threads=64,
                                         log_capacity_per_cell=cap)[4]
    with ThreadPoolExecutor(8) as ex:
        log_t = np.concatenate(list(ex.map(part, np.array_split(t_thr, 8))))
    p2 = p.copy()
    p2.seed = 2
    log_o = pyoracle.simulate_interval(p2, iv, bars, host.make_cell_tasks(p2, "chrL", iv),
                                       nthreads=8, log_capacity_per_cell=cap)[4]
    assert log_t.shape == log_o.shape == (48, cap)
    fields = ["barriers_occupied", "lefs_stalled_rev", "lefs_stalled_fwd", "lefs_stalled_both",
              "lef_bar_collisions", "lef_lef_primary_collisions", "lef_lef_secondary_collisions",
              "loop_size_sum", "num_lefs"]
    for f in fields:
        a = log_t[f][:, 300:].astype(np.float64).mean(axis=1)  # one average per cell
        b = log_o[f][:, 300:].astype(np.float64).mean(axis=1)
        se = np.sqrt(a.var(ddof=1) / len(a) + b.var(ddof=1) / len(b))
        assert abs(a.mean() - b.mean()) <= 4 * se + 1e-9, (f, a.mean(), b.mean(), se)
        assert b.mean() > 0 or f == "lefs_stalled_both", f


@pytest.mark.parametrize("virtual_threads", [64, 33, 256])
def test_results_are_pinned_by_the_golden_digests(throughput, virtual_threads):
    """The throughput mode's outputs are a pure function of the task; the digests committed in
    tests/golden/throughput_mode_digests.json (written by make_throughput_mode_digests.py) pin
    that function, so a kernel change that alters any throughput-mode result is seen."""
    import json
    import os
    import sys

    gold_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, gold_dir)
    import make_throughput_mode_digests as mk

    with open(os.path.join(gold_dir, "throughput_mode_digests.json")) as f:
        gold = json.load(f)
    assert mk.compute(virtual_threads) == gold


def test_c1_scale_statistical_gate_on_the_emulation(throughput):
    """The throughput mode at the size of BASELINE C1 (512 cells, 7.73 M contacts) through the
    emulation of the kernel source -- what the device produces bit for bit
    (tests/test_zz_gpu_throughput_mode.py) -- against the oracle under three other seeds."""
    from stats_eval import c1_inputs, c1_scale_gate

    runs, (nrows, ncols) = c1_inputs()
    p, iv, bars, tasks = runs[1]

    def chunk(lo):
        emu_lib.set_rng_mode(1)  # (the mode is per thread)
        return emu_lib.simulate_interval(p, iv, bars, tasks[lo:lo + 32], virtual_threads=64)

    with ThreadPoolExecutor(8) as ex:
        parts = list(ex.map(chunk, range(0, 512, 32)))
    band = sum(x[0].astype(np.uint64) for x in parts).astype(np.uint32)
    occ = sum(x[1] for x in parts)
    stats = np.concatenate([x[2] for x in parts])
    assert stats["device_fault"].max() == 0
    assert int(band.astype(np.uint64).sum()) + sum(x[3] for x in parts) == 600 * 12889
    ora = {s: pyoracle.simulate_interval(*runs[s], nthreads=8) for s in (2, 3, 4)}
    c1_scale_gate((band, occ, stats), ora, nrows, ncols)
