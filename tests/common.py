"""Shared helpers for the tests: seeded synthetic intervals in the shape of the BASELINE configs."""
import numpy as np

from modle_b200 import abi, host


def make_case(size=3_000_000, ncells=2, nbar=40, seed=1, name="chrT", start=0, end=None, **kw):
    """Returns (params, interval, barriers, tasks) for a synthetic chromosome.

    Barriers follow the generator described in SURVEY 8(d): sorted unique positions, strand
    Bernoulli(0.5), score U[0.6, 1.0].
    """
    p = host.default_params()
    p.num_cells = ncells
    given = set(kw)
    for k, v in kw.items():
        setattr(p, k, v)
    host.transform_params(p, "rev_extrusion_speed" in given, "fwd_extrusion_speed" in given,
                          "extrusion_barrier_occupancy" in given)
    end = size if end is None else end
    rng = np.random.default_rng(seed)
    pos = np.sort(rng.choice(np.arange(start, end), nbar, replace=False)) if nbar else []
    recs = [(int(x), '+' if rng.random() < 0.5 else '-', float(rng.uniform(0.6, 1.0)))
            for x in pos]
    bars = host.barriers_from_records(recs, p)
    iv = abi.Interval(size, start, end, host.compute_num_lefs(p, end - start))
    tasks = host.make_cell_tasks(p, name, iv)
    return p, iv, bars, tasks


def results_equal(a, b):
    """(band, occ1d, stats, missed) tuples; returns a list of the fields that differ."""
    bad = []
    if not np.array_equal(a[0], b[0]):
        bad.append("band")
    if not np.array_equal(a[1], b[1]):
        bad.append("occ1d")
    if a[3] != b[3]:
        bad.append("missed")
    for f in a[2].dtype.names:
        if f == "device_fault":
            continue
        if not np.array_equal(a[2][f], b[2][f]):
            bad.append("stats." + f)
    return bad
