"""`-m gpu`: the throughput mode (MODLE_B200_RNG_COUNTER) of the CUDA path, through the C ABI.

Bit-exact checks against the CPU emulation of the same kernel source (the mode's results are a
pure function of the task, so device and emulation must agree on every integer), the statistical
gate of SURVEY 8c(ii) against the CPU oracle (tolerances of tests/test_statistical_parity.py), and
size-independent properties at the BASELINE C1 size. Runs last (file name) -- the deterministic
mode's parity tests come first."""
import numpy as np
import pytest

import emu_lib
from common import make_case, results_equal
from oracle import pyoracle
from stats_eval import per_diagonal_mean_var, stratum_adjusted_correlation, stripe_pearson
from test_gpu_parity import CASES

pytestmark = pytest.mark.gpu

SUBSET = ["defaults_small", "many_small_cells", "no_bypass", "always_bypass", "fractional_pblock",
          "c4_high_collision", "epochs_criterion", "skip_burnin", "tad_only", "no_1d_track",
          "sub_interval", "narrow_band_missed_updates", "constant_speed", "tiny_interval",
          "no_barriers", "c5_1kb_bins", "c1_chr20_shape", "mid_chr13_shape", "c3_chr1_shape"]


@pytest.fixture(scope="module")
def thr_ctx(product_lib):
    from modle_b200.simulation import RNG_COUNTER, Context

    ctx = Context(0, rng_mode=RNG_COUNTER)
    assert ctx.rng_mode == RNG_COUNTER
    yield ctx
    ctx.close()


@pytest.fixture
def emu_throughput():
    emu_lib.set_rng_mode(1)
    yield
    emu_lib.set_rng_mode(0)


@pytest.mark.parametrize("name", SUBSET)
def test_device_equals_emulation(thr_ctx, emu_throughput, name):
    p, iv, bars, tasks = make_case(seed=5, **CASES[name])
    gpu = thr_ctx.simulate_interval(p, iv, bars, tasks)
    assert gpu[2]["device_fault"].max() == 0
    assert gpu[2]["num_rng_draws"].max() == 0
    emu = emu_lib.simulate_interval(p, iv, bars, tasks, virtual_threads=32)  # (any width: same result)
    assert results_equal(gpu, emu) == []


def test_mode_is_per_context_and_switchable(thr_ctx, gpu_ctx):
    from modle_b200.simulation import RNG_COUNTER, RNG_REFERENCE_ORDER

    p, iv, bars, tasks = make_case(seed=5, **CASES["defaults_small"])
    thr = thr_ctx.simulate_interval(p, iv, bars, tasks)
    det = gpu_ctx.simulate_interval(p, iv, bars, tasks)
    assert gpu_ctx.rng_mode == RNG_REFERENCE_ORDER
    assert results_equal(det, pyoracle.simulate_interval(p, iv, bars, tasks)) == []
    assert results_equal(thr, det) != []
    thr_ctx.set_rng_mode(RNG_REFERENCE_ORDER)
    try:
        assert results_equal(thr_ctx.simulate_interval(p, iv, bars, tasks), det) == []
    finally:
        thr_ctx.set_rng_mode(RNG_COUNTER)
    assert results_equal(thr_ctx.simulate_interval(p, iv, bars, tasks), thr) == []


def test_statistically_equivalent_to_the_oracle(thr_ctx):
    from scipy.stats import ks_2samp

    from modle_b200 import host

    kw = dict(size=20_000_000, ncells=128, nbar=350, target_contact_density=1.0, name="chrS")
    runs = {}
    for seed in (1, 2, 3):
        p, iv, bars, tasks = make_case(seed=7, **kw)
        p.seed = seed
        runs[seed] = (p, iv, bars, host.make_cell_tasks(p, "chrS", iv))
    nrows, ncols = host.band_shape(runs[1][0], 20_000_000)
    gpu = thr_ctx.simulate_interval(*runs[1])
    ora2 = pyoracle.simulate_interval(*runs[2], nthreads=8)
    ora3 = pyoracle.simulate_interval(*runs[3], nthreads=8)
    assert gpu[2]["device_fault"].max() == 0
    m_g, v_g, tot_g = per_diagonal_mean_var(gpu[0], nrows, ncols)
    m_o, v_o, tot_o = per_diagonal_mean_var(ora2[0], nrows, ncols)
    big = (tot_g >= 1e4) & (tot_o >= 1e4)
    assert big.sum() >= 10
    assert np.all(np.abs(m_g[big] / m_o[big] - 1.0) < 0.05)
    assert np.all(np.abs(v_g[big] / v_o[big] - 1.0) < 0.20)
    scc_go = stratum_adjusted_correlation(gpu[0], ora2[0], nrows, ncols, max_d=200)
    scc_oo = stratum_adjusted_correlation(ora2[0], ora3[0], nrows, ncols, max_d=200)
    assert scc_go > scc_oo - 0.01, (scc_go, scc_oo)
    assert scc_go > 0.4
    for direction in ("vertical", "horizontal"):  # per-stripe Pearson of `modle_tools eval`
        r_go = np.nanmedian(stripe_pearson(gpu[0], ora2[0], nrows, ncols, direction))
        r_oo = np.nanmedian(stripe_pearson(ora2[0], ora3[0], nrows, ncols, direction))
        assert r_go > r_oo - 0.005, (direction, r_go, r_oo)
    assert ks_2samp(gpu[2]["num_burnin_epochs"], ora2[2]["num_burnin_epochs"]).pvalue > 0.001


def test_c1_size_properties(thr_ctx):
    """BASELINE C1 (chr20 shape, 512 cells): totals, repeatability, independence of batching."""
    from modle_b200 import workloads
    from modle_b200.simulation import RNG_COUNTER, Simulation

    cfg, genome = workloads.config_c1(512)
    sim = Simulation(cfg, genome, rng_mode=RNG_COUNTER)
    iv = sim.intervals[0]
    from modle_b200 import host
    p = cfg.params
    tasks = host.make_cell_tasks(p, iv.chrom_name, iv.abi_interval())
    a = thr_ctx.simulate_interval(p, iv.abi_interval(), iv.barriers, tasks)
    assert a[2]["device_fault"].max() == 0
    assert np.array_equal(a[2]["num_contacts"], tasks["num_target_contacts"])
    assert int(a[0].sum()) + a[3] == int(tasks["num_target_contacts"].sum())
    b = thr_ctx.simulate_interval(p, iv.abi_interval(), iv.barriers, tasks)
    assert results_equal(a, b) == []
    # two halves, added into the same buffers
    band, occ, s1, m1 = thr_ctx.simulate_interval(p, iv.abi_interval(), iv.barriers, tasks[:200])
    band, occ, s2, m2 = thr_ctx.simulate_interval(p, iv.abi_interval(), iv.barriers, tasks[200:],
                                                  band=band, occ1d=occ)
    assert np.array_equal(band, a[0]) and np.array_equal(occ, a[1]) and m1 + m2 == a[3]
    assert np.array_equal(np.concatenate([s1, s2])["num_epochs"], a[2]["num_epochs"])
    sim.close()


def test_c1_scale_statistical_gate(thr_ctx):
    """Throughput mode on the device vs the oracle with other seeds at the size of BASELINE C1;
    tolerances and their calibration: stats_eval.c1_scale_gate."""
    from stats_eval import c1_inputs, c1_scale_gate

    runs, (nrows, ncols) = c1_inputs()
    gpu = thr_ctx.simulate_interval(*runs[1])
    assert gpu[2]["device_fault"].max() == 0
    assert int(gpu[0].astype(np.uint64).sum()) + gpu[3] == 600 * 12889
    ora = {s: pyoracle.simulate_interval(*runs[s], nthreads=16) for s in (2, 3, 4)}
    c1_scale_gate((gpu[0], gpu[1], gpu[2]), ora, nrows, ncols)
