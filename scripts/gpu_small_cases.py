"""Runs a few small intervals through the CUDA library (used under compute-sanitizer)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from common import make_case
from modle_b200.simulation import Context

CASES = {
    "burnin": dict(size=30_000_000, ncells=2, nbar=300, debug_max_epochs=30),
    "sampling": dict(size=30_000_000, ncells=2, nbar=300, skip_burnin=1, target_contact_density=0.002),
    # whole cells: burn-in to its end, then sampling up to the contact target
    "whole_small": dict(size=6_000_000, ncells=2, nbar=100, target_contact_density=0.01),
    "whole_c4": dict(size=4_000_000, ncells=1, nbar=260, target_contact_density=0.01,
                     number_of_lefs_per_mbp=80, probability_of_extrusion_unit_bypass=0.01),
    "whole_pblock": dict(size=5_000_000, ncells=1, nbar=90, target_contact_density=0.01,
                         lef_bar_major_collision_pblock=0.8, lef_bar_minor_collision_pblock=0.1),
    "mid": dict(size=110_000_000, ncells=1, nbar=900, debug_max_epochs=12),
    "large": dict(size=200_000_000, ncells=1, nbar=2500, debug_max_epochs=8),
}


def main():
    names = sys.argv[1].split(",")
    ctx = Context(0)
    for n in names:
        p, iv, bars, tasks = make_case(**CASES[n])
        band, occ, stats, missed = ctx.simulate_interval(p, iv, bars, tasks)
        print(n, "epochs", stats["num_epochs"].tolist(), "contacts", int(band.sum()), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
