"""How close does the burn-in decision come to depending on the ORDER of a floating-point sum?

The reference adds the squared deviations behind the loop-size coefficient of variation left to
right (src/stats/descriptive_impl.hpp:64-76, reached from compute_loop_size_stats,
src/libmodle/cpu/simulation.cpp:795-819); the CUDA kernel adds them in a fixed tree order
(sim_core.hpp burnin_step), which moves the value by a few ulp (measured below). evaluate_burnin
(:821-864) only COMPARES window means of that series, so the two orders can disagree only when
two window means are within a few ulp of each other. This test measures, on the oracle, how close
the comparisons actually come on BASELINE-shaped cells: the claim "no burn-in decision is within
reach of the summation order" is a number here, not an argument. (The parity tests then confirm
the consequence: every cell's burn-in length equals the oracle's.)"""
import numpy as np

from common import make_case
from modle_b200 import abi, host, workloads
from oracle import pyoracle


def test_window_mean_comparisons_stay_far_from_ties_on_c1_cells():
    cfg, genome = workloads.config_c1(512)
    name, size, start, end, recs = genome[0]
    p = cfg.params.copy()
    p.target_contact_density = 0.01  # the burn-in is what matters here
    bars = host.barriers_from_records(recs, p)
    iv = abi.Interval(size, start, end, host.compute_num_lefs(p, end - start))
    tasks = host.make_cell_tasks(p, name, iv)[:48]
    pyoracle.burnin_margin(reset=True)
    res = pyoracle.simulate_interval(p, iv, bars, tasks, nthreads=8)
    comparisons, near, min_gap = pyoracle.burnin_margin(reset=True)
    epochs = int(res[2]["num_burnin_epochs"].sum())
    assert epochs > 48 * 200 and comparisons > 1_000_000
    # summation order moves the coefficient of variation by ~n * eps / sqrt(n) relative in the
    # worst case: a handful of ulp for n ~ 1,300 .. 5,000 terms; 64 ulp is a generous bound
    assert near == 0, (comparisons, near, min_gap)
    # and the closest comparison of the whole run is still orders of magnitude away from it
    assert min_gap > 1e-11, min_gap


def test_tree_and_left_to_right_sums_differ_by_a_few_ulp_only():
    """The size of the effect itself: both orders on loop-size-like data."""
    rng = np.random.default_rng(3)
    worst = 0.0
    for n in (1289, 4979):
        for _ in range(20):
            x = rng.exponential(120_000.0, n).round()
            d = (x - x.mean()) ** 2
            left = 0.0
            for v in d:
                left = left + v
            tree = d.copy()
            while len(tree) > 1:  # pairwise
                if len(tree) % 2:
                    tree = np.append(tree, 0.0)
                tree = tree[0::2] + tree[1::2]
            worst = max(worst, abs(left - tree[0]) / left)
    assert worst < 64 * 2.220446049250313e-16, worst
