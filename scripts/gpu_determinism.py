"""Repeats the C1 full-size run and reports cells whose per-cell stats differ between repeats,
checking the differing cells against the oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

from common import make_case
from modle_b200.simulation import Context
from oracle import pyoracle


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    warm = sys.argv[2] if len(sys.argv) > 2 else ""
    ctx = Context(0)
    if warm:  # run other configurations first, like the test suite does
        for kw in (dict(size=114_364_328, ncells=3, nbar=943, target_contact_density=0.0008, name="chr13"),
                   dict(size=145_138_636, ncells=3, nbar=1772, target_contact_density=0.0006, name="chr8")):
            p, iv, bars, tasks = make_case(**kw)
            ctx.simulate_interval(p, iv, bars, tasks)
    p, iv, bars, tasks = make_case(size=64_444_167, ncells=512, nbar=1132, name="chr20")
    runs = [ctx.simulate_interval(p, iv, bars, tasks) for _ in range(reps)]
    bad = set()
    for r in runs[1:]:
        for f in ("num_epochs", "num_burnin_epochs", "num_rng_draws", "num_lef_updates"):
            bad |= set(np.nonzero(r[2][f] != runs[0][2][f])[0].tolist())
    print("band equal:", [bool(np.array_equal(r[0], runs[0][0])) for r in runs])
    print("differing cells:", sorted(bad))
    for c in sorted(bad)[:6]:
        o = pyoracle.simulate_interval(p, iv, bars, tasks[c:c + 1], nthreads=1)
        print("cell", c, "oracle epochs/draws", int(o[2]["num_epochs"][0]), int(o[2]["num_rng_draws"][0]),
              "runs:", [(int(r[2]["num_epochs"][c]), int(r[2]["num_rng_draws"][c]),
                         int(r[2]["device_fault"][c])) for r in runs])
    ctx.close()


if __name__ == "__main__":
    main()
