"""Times one interval on the GPU and prints the kernel's per-phase cycle breakdown.

    python scripts/gpu_phases.py [c1|c3|c4|c5] [cells] [repeat]

Uses whatever library MODLE_B200_LIB points to (default: the product build).
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from modle_b200 import abi, host, workloads
from modle_b200.simulation import Context, Simulation


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c1"
    cells = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    repeat = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    if wl == "c1":
        cfg, genome = workloads.config_c1(cells)
    elif wl == "c3":
        cfg, genome = workloads.config_c3(cells)
    elif wl == "c4":
        cfg, genome = workloads.config_c4(cells)
    elif wl == "c5":  # full C5 geometry (2.9 GB band); density via MODLE_B200_C5_DENSITY
        cfg, genome = workloads.config_c5(
            cells, target_contact_density=float(os.environ.get("MODLE_B200_C5_DENSITY", "0.05")))
    else:
        raise SystemExit("unknown workload")
    sim = Simulation(cfg, genome)
    iv = sim.intervals[0]
    p = cfg.params
    tasks = host.make_cell_tasks(p, iv.chrom_name, iv.abi_interval())
    ctx = Context(0)
    best = None
    for r in range(repeat):
        ctx.phase_cycles(reset=True)
        t0 = time.perf_counter()
        band, occ, stats, missed = ctx.simulate_interval(p, iv.abi_interval(), iv.barriers, tasks)
        dt = time.perf_counter() - t0
        ph = ctx.phase_cycles(reset=True)
        if best is None or dt < best[0]:
            best = (dt, ph, stats)
    dt, ph, stats = best
    lu = int(stats["num_lef_updates"].sum())
    ep = int(stats["num_epochs"].sum())
    tot = max(1, ph["total"])
    print(f"[{os.environ.get('MODLE_B200_LIB', 'product')}] {wl} cells={cells} n_lefs={iv.num_lefs} "
          f"n_bar={len(iv.barriers)}: {dt * 1e3:.1f} ms e2e, {lu / dt / 1e6:.1f} M LEF-updates/s, "
          f"{ep} cell-epochs, {tot / ep:.0f} cycles/cell-epoch, faults={int(stats['device_fault'].max())}")
    for k, v in ph.items():
        if k != "total":
            print(f"    {k:20s} {100 * v / tot:6.2f}%  {v / ep:10.0f} cyc/epoch")
    ctx.close()


if __name__ == "__main__":
    main()
