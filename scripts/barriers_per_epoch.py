#!/usr/bin/env python3
"""CTA barriers per cell-epoch, counted by the CPU emulation of the kernel source (cta.hpp
emu_barrier_count: explicit barriers + three per block-wide collective). The epoch loop is bound
by barrier-separated short phases, so this is the figure a restructuring of the
phases moves first; it needs no GPU.

    python scripts/barriers_per_epoch.py [--size 64444167 --nbar 1132 --cells 2]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=64_444_167)
    ap.add_argument("--nbar", type=int, default=1132)
    ap.add_argument("--cells", type=int, default=2)
    ap.add_argument("--threads", type=int, default=256)
    args = ap.parse_args()
    import emu_lib
    from common import make_case

    p, iv, bars, tasks = make_case(size=args.size, ncells=args.cells, nbar=args.nbar, seed=5,
                                   target_contact_density=0.002 * 64_444_167 / args.size)
    for mode, name in ((0, "deterministic"), (1, "throughput")):
        emu_lib.set_rng_mode(mode)
        emu_lib.barrier_count()
        emu_lib.phase_barriers()
        r = emu_lib.simulate_interval(p, iv, bars, tasks, virtual_threads=args.threads)
        n = emu_lib.barrier_count()
        st = r[2]
        ep, burn = int(st["num_epochs"].sum()), int(st["num_burnin_epochs"].sum())
        print(f"{name:14s} {int(iv.num_lefs)} LEFs, {len(bars)} barriers: {ep} cell-epochs "
              f"({burn} burn-in), {n} CTA barriers = {n / ep:.1f} per cell-epoch")
        ph = emu_lib.phase_barriers()
        print("    " + "  ".join(f"{k} {v / ep:.1f}" for k, v in ph.items()
                                 if v and k != "total" and "." not in k and "(" not in k))
    emu_lib.set_rng_mode(0)


if __name__ == "__main__":
    main()
