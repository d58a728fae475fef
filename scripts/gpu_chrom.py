"""Times single chromosomes of the C2 genome on the GPU (device time of the whole host-buffer call).

    python scripts/gpu_chrom.py chr8,chr13 [cells] [repeat]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modle_b200 import host, workloads
from modle_b200.simulation import Context, Simulation


def main():
    names = sys.argv[1].split(",")
    cells = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    repeat = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    cfg, genome = workloads.config_c2(cells)
    sim = Simulation(cfg, [g for g in genome if g[0] in names])
    p = cfg.params
    ctx = Context(0)
    for iv in sim.intervals:
        tasks = host.make_cell_tasks(p, iv.chrom_name, iv.abi_interval())
        best = None
        for _ in range(repeat):
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
        print(f"[MID={os.environ.get('MODLE_B200_MID', 'default')}] {iv.chrom_name} n_lefs={iv.num_lefs} "
              f"n_bar={len(iv.barriers)} cells={cells}: {dt * 1e3:.1f} ms, {lu / dt / 1e6:.1f} M LEF-updates/s, "
              f"{ph['total'] / ep:.0f} cycles/cell-epoch, faults={int(stats['device_fault'].max())}",
              flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
