#!/usr/bin/env python3
"""What the heavy tail of cell durations costs, and what a single device-side queue could
recover: a list-scheduling model fed with REAL per-cell epoch counts (the CPU oracle runs a sample
of cells of every interval of the C2 workload; durations = epochs x measured cycles per
cell-epoch of the interval's kernel configuration, profiles/r01i_stream_sweep.txt).

    python scripts/schedule_model.py [--cells-per-interval 64]

148 SMs; an SM hosts 1 <1024,1> cell, 2 <512,2> cells or 3 <256,3> cells, one class at a time.
Policies: launches one after the other; overlapped on k streams (a stream's next launch opens
when its previous one has drained; older launches get free SMs first); all launches open at once;
one queue per configuration class. Only occupancy is modelled (a cell runs at the same speed
whatever shares its SM), so read the ratios, not the absolute times.
"""
import argparse
import heapq
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

# cycles per cell-epoch: modle_b200.distributed.cell_epoch_cycles (the planner's cost model,
# fitted to profiles/r01i_stream_sweep.txt), scaled by CAL so that the modelled 3-stream
# single-GPU step equals the measured one (4.70 s, profiles/r01i_bench_default_c2.json).
CAL = 4.70 / 4.342
CAP = {"large": 1, "mid": 2, "small": 3}                  # resident cells per SM
CLOCK = 1.965e9
SMS = 148
LIMITS = (75 * 1024, 113 * 1024)  # bytes of shared memory per cell: <256,3> / <512,2> / <1024,1>


def klass(n_lefs, n_bar):
    # sim_types.hpp cell_array_bytes + sizeof(CellShared)
    w = 260 + 7 * n_lefs + max(n_lefs, n_bar) + 64 + 12 * (n_lefs // 32 + 3) + n_bar + n_lefs + 2 \
        + (n_bar + 3) // 4 + 1
    b = (w * 4 + 15) // 16 * 16 + 6264
    return "small" if b <= LIMITS[0] else ("mid" if b <= LIMITS[1] else "large")


def simulate(launches, streams):
    """Event-driven model of the hardware work distributor. launches: list of dicts {k, dur[]} in
    issue order, launch i on stream i % streams (None: all launches open from the start). An SM
    hosts cells of ONE class at a time (the shared-memory carve-out differs) up to CAP[class];
    an open launch hands its next cell to any SM with room; older launches first."""
    n = len(launches)
    nxt = [0] * n                     # next cell of each launch
    running = [0] * n                 # cells of the launch still running
    opened = [streams is None or i < streams for i in range(n)]
    sm_class = [None] * SMS
    sm_used = [0] * SMS
    events = []                       # (time, sm, launch)
    t = 0.0
    busy_area = 0.0
    while True:
        # hand out work
        progress = True
        while progress:
            progress = False
            for i in range(n):
                if not opened[i] or nxt[i] >= len(launches[i]["dur"]):
                    continue
                k = launches[i]["k"]
                for s in range(SMS):
                    if nxt[i] >= len(launches[i]["dur"]):
                        break
                    if (sm_class[s] in (None, k)) and sm_used[s] < CAP[k]:
                        while sm_used[s] < CAP[k] and nxt[i] < len(launches[i]["dur"]):
                            d = launches[i]["dur"][nxt[i]]
                            nxt[i] += 1
                            running[i] += 1
                            sm_class[s] = k
                            sm_used[s] += 1
                            heapq.heappush(events, (t + d, s, i))
                            busy_area += d / CAP[k]
                            progress = True
        if not events:
            break
        t, s, i = heapq.heappop(events)
        sm_used[s] -= 1
        if sm_used[s] == 0:
            sm_class[s] = None
        running[i] -= 1
        if streams is not None and running[i] == 0 and nxt[i] >= len(launches[i]["dur"]):
            j = i + streams           # the stream's next launch may start
            if j < n:
                opened[j] = True
    return t, busy_area / SMS


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells-per-interval", type=int, default=64)
    ap.add_argument("--cells", type=int, default=512)
    ap.add_argument("--cache", default="", help="npz file that keeps the oracle's epoch samples")
    ap.add_argument("--ranks", default="", help="comma list of world sizes: compare shard plans")
    ap.add_argument("--streams", type=int, default=3)
    args = ap.parse_args()
    from modle_b200 import distributed, host, workloads
    from modle_b200.simulation import Simulation
    from oracle import pyoracle

    cfg, genome = workloads.config_c2(args.cells)
    sim = Simulation(cfg, genome)
    p = cfg.params
    rng = np.random.default_rng(1)
    intervals = []
    cache = {}
    if args.cache and os.path.exists(args.cache):
        cache = dict(np.load(args.cache))
    for iv in sim.intervals:
        key = f"{iv.chrom_name}:{args.cells_per_interval}"
        if key not in cache:
            tasks = host.make_cell_tasks(p, iv.chrom_name, iv.abi_interval())[:args.cells_per_interval]
            st = pyoracle.simulate_interval(p, iv.abi_interval(), iv.barriers, tasks,
                                            nthreads=os.cpu_count() or 1, want_occ=False)[2]
            cache[key] = st["num_epochs"].astype(np.float64)
            if args.cache:
                np.savez(args.cache, **cache)
        ep = cache[key]
        k = klass(iv.num_lefs, len(iv.barriers))
        # bootstrap the interval's 512 cells from the sample
        epochs = rng.choice(ep, size=args.cells, replace=True)
        dur = epochs * CAL * (distributed.cell_cost(iv.num_lefs, len(iv.barriers)) * host.launch_geometry(iv.num_lefs, len(iv.barriers))[1]) / CLOCK
        intervals.append(dict(name=iv.chrom_name, k=k, n=iv.num_lefs, nb=len(iv.barriers), dur=dur,
                              mean_epochs=ep.mean(), max_over_mean=ep.max() / ep.mean()))
        print(f"{iv.chrom_name:6s} {k:5s} N={iv.num_lefs:5d} epochs mean {ep.mean():7.1f} "
              f"max/mean {ep.max() / ep.mean():.2f}", flush=True)
    launches = [dict(k=d["k"], dur=list(d["dur"])) for d in sorted(intervals, key=lambda d: -d["n"])]
    t1, ideal = simulate(launches, 1)
    print(f"\nideal (no idle SM)              {ideal:6.3f} s")
    print(f"one launch after the other      {t1:6.3f} s   efficiency {ideal / t1:5.1%}")
    for k in (2, 3, 4, 8):
        tk, _ = simulate(launches, k)
        print(f"{k} streams                       {tk:6.3f} s   efficiency {ideal / tk:5.1%}")
    tall, _ = simulate(launches, None)
    print(f"all 24 launches open at once    {tall:6.3f} s   efficiency {ideal / tall:5.1%}")
    merged = []
    for k in ("large", "mid", "small"):
        dur = [x for d in intervals if d["k"] == k for x in d["dur"]]
        if dur:
            merged.append(dict(k=k, dur=dur))
    tq, _ = simulate(merged, None)
    print(f"one queue per class (3 launches) {tq:5.3f} s   efficiency {ideal / tq:5.1%}")
    if args.ranks:
        compare_plans(intervals, args, ideal)


def rank_makespans(intervals, shards, world, streams):
    """Per-rank makespan of a shard plan: a rank issues its shards heaviest-first on `streams`."""
    out = []
    for r in range(world):
        mine = sorted((s for s in shards if s.rank == r), key=lambda s: (-s.weight, s.interval))
        launches = [dict(k=intervals[s.interval]["k"],
                         dur=list(intervals[s.interval]["dur"][s.cell_lo:s.cell_hi])) for s in mine]
        out.append(simulate(launches, streams)[0] if launches else 0.0)
    return out


def compare_plans(intervals, args, ideal):
    from modle_b200 import distributed

    n_lefs = [d["n"] for d in intervals]
    print("\nshard plans (max over ranks of the modelled per-rank time; reduce not included)")
    for world in [int(x) for x in args.ranks.split(",")]:
        row = [f"{world} ranks: perfect {ideal / world:6.3f} s"]
        cost = [distributed.cell_cost(d["n"], d["nb"]) for d in intervals]
        for name, w, kw in (("whole/LEFs", n_lefs, dict()), ("whole/cost", cost, dict()),
                            ("sliced", cost, dict(slice_all=True))):
            shards = distributed.plan_shards(w, args.cells, world, **kw)
            ts = rank_makespans(intervals, shards, world, args.streams)
            nsplit = sum(1 for _, (_, rk) in distributed.interval_roots(shards).items() if len(rk) > 1)
            row.append(f"{name} {max(ts):6.3f} s ({ideal / world / max(ts):5.1%}, {nsplit} split)")
        print("   ".join(row))


if __name__ == "__main__":
    main()
