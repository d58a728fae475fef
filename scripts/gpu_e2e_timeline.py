"""Diagnostic: where an end-to-end step (Simulation.run_simulate, host buffers) spends its time.
Prints per-interval submit / return times of the worker threads and the step's wall time."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from modle_b200 import workloads
from modle_b200.simulation import Context, Simulation


def main():
    workers = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    cfg, genome = workloads.config_c2(512)
    sim = Simulation(cfg, genome)
    log = []
    orig = Context.simulate_interval

    def traced(self, params, interval, barriers, tasks, **kw):
        t0 = time.perf_counter()
        out = orig(self, params, interval, barriers, tasks, **kw)
        log.append((t0, time.perf_counter(), int(interval.num_lefs)))
        return out

    Context.simulate_interval = traced
    for step in range(3):
        log.clear()
        t0 = time.perf_counter()
        sim.run_simulate(num_workers=workers)
        dt = time.perf_counter() - t0
        busy = sum(b - a for a, b, _ in log)
        print(f"step {step}: {dt * 1e3:.0f} ms wall, {len(log)} calls, sum of call times {busy * 1e3:.0f} ms "
              f"({busy / dt:.2f} calls in flight on average)")
    for a, b, n in sorted(log):
        print(f"  n_lefs {n:5d}  start {1e3 * (a - t0):7.0f}  end {1e3 * (b - t0):7.0f}  ({1e3 * (b - a):6.0f} ms)")
    sim.close()


if __name__ == "__main__":
    main()
