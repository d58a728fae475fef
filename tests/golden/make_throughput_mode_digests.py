#!/usr/bin/env python3
"""Writes tests/golden/throughput_mode_digests.json: SHA-256 digests of what the throughput mode
(MODLE_B200_RNG_COUNTER) produces for a set of seeded cases, taken from the CPU emulation of the
kernel source (tests/emu). The throughput mode has no reference counterpart to be bit-exact with
(its gate against the reference is statistical), but its results are a pure function of the task:
these digests pin that function, so that a change of the kernel source that alters any
throughput-mode result -- on purpose or not -- is seen (tests/test_throughput_mode.py, and on the
device tests/test_zz_gpu_throughput_mode.py, which compares device == emulation).

    python tests/golden/make_throughput_mode_digests.py        # rewrite after an INTENDED change
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

CASES = {
    "defaults": dict(size=6_000_000, ncells=3, nbar=100, target_contact_density=0.02),
    "high_collision": dict(size=5_000_000, ncells=2, nbar=300, target_contact_density=0.02,
                           number_of_lefs_per_mbp=80, probability_of_extrusion_unit_bypass=0.01),
    "fractional": dict(size=5_000_000, ncells=2, nbar=90, target_contact_density=0.02,
                       lef_bar_major_collision_pblock=0.8, lef_bar_minor_collision_pblock=0.1,
                       probability_of_extrusion_unit_bypass=0.4),
    "no_bypass": dict(size=5_000_000, ncells=2, nbar=90, target_contact_density=0.02,
                      probability_of_extrusion_unit_bypass=0.0),
    "always_bypass": dict(size=5_000_000, ncells=2, nbar=90, target_contact_density=0.02,
                          probability_of_extrusion_unit_bypass=1.0),
    "one_kb_bins": dict(size=3_000_000, ncells=1, nbar=60, bin_size=1000,
                        target_contact_density=0.005),
    "tad_only_sub_interval": dict(size=9_000_000, start=2_000_000, end=6_500_000, ncells=2, nbar=70,
                                  target_contact_density=0.02, contact_sampling_strategy=3),
    "epochs_criterion_constant_speed": dict(size=5_000_000, ncells=2, nbar=90, stopping_criterion=1,
                                            target_simulation_epochs=60,
                                            rev_extrusion_speed_std=0.0, fwd_extrusion_speed_std=0.0),
    "skip_burnin_tiny": dict(size=400_000, ncells=3, nbar=6, skip_burnin=1,
                             target_contact_density=0.01),
}


def digest(result):
    band, occ, stats, missed = result
    h = hashlib.sha256()
    h.update(band.tobytes())
    h.update(occ.tobytes())
    for f in ("num_contacts", "num_epochs", "num_burnin_epochs", "num_lef_updates"):
        h.update(stats[f].tobytes())
    h.update(str(int(missed)).encode())
    return h.hexdigest()


def compute(virtual_threads=64):
    import emu_lib
    from common import make_case

    emu_lib.set_rng_mode(1)
    try:
        out = {}
        for name, kw in CASES.items():
            p, iv, bars, tasks = make_case(**kw)
            out[name] = digest(emu_lib.simulate_interval(p, iv, bars, tasks,
                                                         virtual_threads=virtual_threads))
        return out
    finally:
        emu_lib.set_rng_mode(0)


if __name__ == "__main__":
    d = compute()
    with open(os.path.join(HERE, "throughput_mode_digests.json"), "w") as f:
        json.dump(d, f, indent=1, sort_keys=True)
    print(json.dumps(d, indent=1))
