"""Gate (ii) of SURVEY 8c: the CUDA path against the oracle run with DIFFERENT seeds, compared
the way `modle_tools eval` compares matrices -- per-diagonal mean / variance and the
stratum-adjusted correlation -- with two independent oracle seeds as the noise floor, plus the
burn-in length distribution. (With the same seed the two are bit-identical, test_gpu_parity.py;
this test is what a relaxed-draw-order mode would have to pass as well.)

Tolerances, tuned against the oracle's own seed-to-seed variability at this size (128 cells,
2.4 M pixels, 2.4 M contacts; five seed pairs gave: per-diagonal mean within 2.8 %, variance
within 12.7 %, SCC 0.467 - 0.471, KS p >= 0.06):
  * per-diagonal mean: within 5 % for diagonals holding >= 1e4 contacts
  * per-diagonal variance: within 20 % for the same diagonals
  * SCC(gpu, oracle) not more than 0.01 below SCC(oracle, oracle')
  * median per-stripe Pearson (the `modle_tools eval` definition, vertical and horizontal
    stripes; 0.965 +- 0.0003 over oracle seed pairs) not more than 0.005 below the oracle pair's
  * burn-in epochs: two-sample KS test p > 0.001
"""
import numpy as np
import pytest

from common import make_case
from oracle import pyoracle
from stats_eval import per_diagonal_mean_var, stratum_adjusted_correlation, stripe_pearson

pytestmark = pytest.mark.gpu


def test_independent_seeds_are_statistically_equivalent(gpu_ctx):
    from scipy.stats import ks_2samp

    from modle_b200 import host

    kw = dict(size=20_000_000, ncells=128, nbar=350, target_contact_density=1.0, name="chrS")
    runs = {}
    for seed in (1, 2, 3):
        p, iv, bars, tasks = make_case(seed=7, **kw)  # same barriers, different Config::seed
        p.seed = seed
        tasks = host.make_cell_tasks(p, "chrS", iv)
        runs[seed] = (p, iv, bars, tasks)
    nrows, ncols = host.band_shape(runs[1][0], 20_000_000)
    gpu = gpu_ctx.simulate_interval(*runs[1])
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
    assert scc_go > 0.4  # barriers leave a shared structure on top of the sampling noise
    for direction in ("vertical", "horizontal"):  # per-stripe Pearson of `modle_tools eval`
        r_go = np.nanmedian(stripe_pearson(gpu[0], ora2[0], nrows, ncols, direction))
        r_oo = np.nanmedian(stripe_pearson(ora2[0], ora3[0], nrows, ncols, direction))
        assert r_go > r_oo - 0.005, (direction, r_go, r_oo)

    ks = ks_2samp(gpu[2]["num_burnin_epochs"], ora2[2]["num_burnin_epochs"])
    assert ks.pvalue > 0.001, ks
