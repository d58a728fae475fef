#!/usr/bin/env python3
"""Benchmark of the loop-extrusion hot path (see BASELINE.md / SURVEY.md 8d).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c2|c1|c3|c4|c5]

One "step" = one whole simulation of the workload (default: BASELINE config C2, genome-wide
GRCh38 shape, 24 chromosomes x 512 cells, default parameters, synthetic barriers). Metric:
LEF-updates/s (1 LEF-update = one active LEF carried through one epoch, burn-in included).

  value  device-resident: cell tasks / band matrices live in HBM, kernels only (+ band memset)
  e2e    the reference-facing call with HOST buffers (modle_b200_simulate_interval_overwrite
         through Simulation.run_simulate: H2D of tasks and barriers, D2H of band / 1D occupancy /
         stats inside the timed region); N > 1: device shards + NCCL reduce + D2H on the roots
  extra  (default C2 line only) throughput_mode: the same workload with counter-based draws;
         c3: BASELINE C3 (chr1 x 8192 cells) timed AND verified -- cells split over the ranks, one
         NCCL reduce, checksums against the committed single-GPU golden, sampled cells against the
         CPU oracle; a mismatch ends the bench with a non-zero exit code;
         register: the contact-register kernel in isolation with its rooflines and an independent
         calibration of the atomic ceilings
  roofline, cpu_baseline, clocks, gpu_launches: see DESIGN.md "Measurement"

--rng-mode throughput runs the main workload with counter-based draws ("Throughput
mode"); the default and the headline is the deterministic mode.

N > 1 (launched by torch.distributed.run, one rank per GPU): modle_b200_plan_shards deals whole
chromosomes to ranks heaviest-first, or one cell slice of every chromosome per rank when a slice
still fills a GPU (--plan auto|whole|slices); a chromosome whose cells are split over ranks is
summed with one NCCL reduce. Total work is fixed, so "scaling" is "strong".

--impl reference: the reference's CPU implementation of the path cannot be built offline
(no Boost/xoshiro-cpp/...), so this arm times the oracle restatement
(oracle/liboracle.so) on all host cores over a bounded sample of the same workload, scheduled the
way the reference schedules its run: one queue over all (interval, cell) tasks.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "LEF-updates/sec (GRCh38 genome-wide, default parameters)"
UNIT = "LEF-updates/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def ncu_traffic_per_lef_update(rng_mode=0):
    """DRAM bytes (read + written) per LEF-update of k_simulate_cells from the committed
    `ncu --set full` capture (profiles/ncu_traffic.json; one launch of the chr1 shape). The
    bench scales it by the LEF-updates of an average launch."""
    key = "k_simulate_cells" if rng_mode == 0 else "k_simulate_cells_throughput_mode"
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return float(d[key]["dram_bytes_per_lef_update"]), d[key]["source"]
    except Exception:
        return None, ("no ncu capture of the throughput-mode kernel yet" if rng_mode else None)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                     "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                f = [x.strip() for x in out.stdout.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names)
                   if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


WORKLOAD_DESC = {
    "c1": "C1 chr20 shape, %d cells, 5 kb bins, default parameters",
    "c2": "C2 GRCh38 genome-wide shape (24 chromosomes, 38,815 synthetic barriers), %d cells, "
          "5 kb bins, default parameters",
    "c3": "C3 chr1 shape, %d cells, 5 kb bins, default parameters",
    "c4": "C4 high-collision: chr20 size, 80 LEFs/Mbp, barriers 1/15 kb, bypass 0.01, %d cells",
    "c5": "C5 fine-resolution scatter: chr2 shape, 1 kb bins, 3 Mbp diagonal width, %d cells",
}


def build_workload_desc(name, cells):
    """(Config overrides, genome, description) as plain data (no native code involved)."""
    from modle_b200 import workloads

    overrides, genome = workloads.spec(name, cells or None)
    desc = WORKLOAD_DESC[name]
    chroms = os.environ.get("MODLE_B200_BENCH_CHROMS")  # diagnostics: restrict the genome
    if chroms:
        keep = set(chroms.split(","))
        genome = [g for g in genome if g[0] in keep]
        desc += " [only " + chroms + "]"
    return overrides, genome, desc % int(overrides["num_cells"])


def build_workload(name, cells, **more):
    from modle_b200.simulation import Config

    overrides, genome, desc = build_workload_desc(name, cells)
    overrides.update(more)
    return Config(**overrides).transform(), genome, desc


# ------------------------------------------------------------------------------- CPU (oracle)
# The reference feeds the (interval, cell) tasks of the WHOLE run through one queue that all
# worker threads pop from (scheduler_simulate.cpp:104-160 produce, :190-271 consume), so the CPU
# arm does the same: one queue over every sampled task, `cores` workers, at least
# CPU_MIN_CELLS_PER_THREAD cells per worker (burn-in lengths vary 3x from cell to cell; with one
# cell per thread the wall time would be the slowest cell's). Inputs are built by oracle-side code
# only (oracle/pyparams.py, the oracle's own task fan-out): the product library is not loaded.
CPU_MIN_CELLS_PER_THREAD = 8
CPU_LU_PER_CORE_S = 3.0e6   # what the oracle does per core, to size the sample (a guess is fine)


def cpu_sample_plan(overrides, genome, cores, target_seconds):
    """Cells per interval so that the sample is about `target_seconds` of wall time and every
    worker gets >= CPU_MIN_CELLS_PER_THREAD tasks from the queue."""
    from oracle import pyparams

    p = pyparams.make_params(**overrides)
    n_iv = max(1, len(genome))
    lu_per_cell_all = sum(pyparams.compute_num_lefs(p, e - s) for _, _, s, e, _ in genome) * 545.0
    by_time = target_seconds * cores * CPU_LU_PER_CORE_S / max(1.0, lu_per_cell_all)
    by_queue = CPU_MIN_CELLS_PER_THREAD * cores / n_iv
    return int(max(1, min(int(p.num_cells), math.ceil(max(by_time, by_queue)))))


def oracle_sample_run(overrides, genome, cells_per_interval, cores, jobs=None):
    """Simulates the first `cells_per_interval` cells of every interval with the CPU oracle through
    one task queue. Returns (lef_updates, seconds, params_and_jobs)."""
    from oracle import pyoracle

    if jobs is None:
        jobs = pyoracle.genome_jobs(overrides, genome, cells_per_interval)
    p, j = jobs
    pyoracle.lib()
    t0 = time.perf_counter()
    res = pyoracle.simulate_genome(p, j, nthreads=cores)
    dt = time.perf_counter() - t0
    lu = sum(int(r[2]["num_lef_updates"].sum()) for r in res)
    return lu, dt, jobs


def cpu_arm(workload, cells, cores, target_seconds, warmup, steps):
    """Times the oracle on a bounded sample; returns (value, seconds per step, cpu_baseline dict)."""
    from modle_b200 import workloads

    overrides, genome = workloads.spec(workload, cells or None)
    chroms = os.environ.get("MODLE_B200_BENCH_CHROMS")
    if chroms:
        genome = [g for g in genome if g[0] in set(chroms.split(","))]
    cpi = cpu_sample_plan(overrides, genome, cores, target_seconds)
    jobs = None
    for _ in range(max(1, warmup)):   # first call: page faults, thread start-up, table set-up
        _, _, jobs = oracle_sample_run(overrides, genome, max(1, cpi // 8), cores)
    jobs = None
    lu_tot, t_tot = 0, 0.0
    for _ in range(steps):
        lu, dt, jobs = oracle_sample_run(overrides, genome, cpi, cores, jobs)
        lu_tot += lu
        t_tot += dt
    n_tasks = sum(len(j[2]) for j in jobs[1])
    sample = (f"first {cpi} cell(s) of each of the {len(jobs[1])} intervals per step = {n_tasks} "
              f"(interval, cell) tasks in ONE queue over {cores} worker threads "
              f"({n_tasks / cores:.1f} per thread); {lu_tot // max(1, steps)} LEF-updates, "
              f"{t_tot / max(1, steps):.1f} s per step; warm-up run excluded")
    value = lu_tot / t_tot
    return value, t_tot / max(1, steps), {
        "value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
        "cells_per_thread": n_tasks / cores, "same_config": False,
        "note": "in-repo oracle port (the real modle cannot be built offline); a sub-sample of "
                "the cells of the same workload"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    steps = max(1, args.steps)
    # the whole --steps K --warmup W run should end within a few minutes
    target = max(4.0, min(20.0, 180.0 / (steps + 1)))
    value, s_per_step, cpu = cpu_arm(args.workload, args.cells, cores, target, args.warmup, steps)
    _, _, desc = build_workload_desc(args.workload, args.cells)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * s_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int64/f64",
        "data": "synthetic", "config": {"workload": desc, "sample": cpu["sample"]},
        "cpu_baseline": cpu,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference cannot be built offline (SURVEY 8c); oracle restatement timed instead",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- GPU
class DeviceJob:
    """One workload staged on this rank's GPU: cell tasks and band matrices resident in HBM, the
    shard plan of modle_b200.distributed, and `step()` = one whole simulation of the workload
    (kernels + band memsets + the reduce of every interval whose cells are split over ranks)."""

    def __init__(self, torch, dist, engine, cfg, genome, rank, world, local_rank, rng_mode,
                 slice_all=None, tolerance=1.10, pre_split=1):
        from modle_b200 import abi, distributed, host
        from modle_b200.simulation import Simulation

        self.torch, self.dist, self.engine = torch, dist, engine
        self.rank, self.world = rank, world
        self.p = cfg.params
        self.dev = torch.device("cuda", local_rank)
        self.sim = Simulation(cfg, genome, device=local_rank, rank=rank, world_size=world,
                              rng_mode=rng_mode, slice_all=slice_all)
        _, self.task_dt, self.stats_dt = abi.np_dtypes()
        self.barrier_dt = abi.np_dtypes()[0]
        ivs = self.sim.intervals
        self.shards = distributed.plan_shards(distributed.interval_weights(ivs),
                                              int(self.p.num_cells), world, tolerance=tolerance,
                                              slice_all=slice_all, pre_split=pre_split)
        self.roots = distributed.interval_roots(self.shards)
        self.split = [i for i, (_, ranks) in self.roots.items() if len(ranks) > 1]
        order = sorted(self.shards, key=lambda s: (-s.weight, s.interval, s.cell_lo))
        # reduces are issued in the same order on every rank: that of the first launch of each
        # split interval in the global plan
        self.split.sort(key=lambda i: next(k for k, s in enumerate(order) if s.interval == i))
        self.bufs, self.mine = {}, []
        for s in order:
            if s.rank != rank:
                continue
            iv = ivs[s.interval]
            if s.interval not in self.bufs:
                self.bufs[s.interval] = engine.alloc_outputs(iv.nrows, iv.ncols)
            tasks = host.make_cell_tasks(self.p, iv.chrom_name, iv.abi_interval())[s.cell_lo:s.cell_hi]
            h_tasks = torch.from_numpy(tasks.view(np.uint8).reshape(-1).copy()).pin_memory()
            self.mine.append(dict(
                iv=iv, idx=s.interval, abi_iv=iv.abi_interval(), ntasks=len(tasks),
                cells=(s.cell_lo, s.cell_hi), d_tasks=h_tasks.to(self.dev),
                d_stats=torch.zeros(len(tasks) * self.stats_dt.itemsize, dtype=torch.uint8,
                                    device=self.dev)))
        for idx in self.split:  # a rank without a piece of a split interval contributes zeros
            if idx not in self.bufs:
                self.bufs[idx] = engine.alloc_outputs(ivs[idx].nrows, ivs[idx].ncols)
        self.reduce_bytes = sum(b.numel() * b.element_size() for i in self.split
                                for b in self.bufs[i])

    def parallelism(self):
        return (f"{self.world} rank(s), {len(self.shards)} (interval, cell-range) shards dealt "
                f"heaviest-first, {len(self.split)} interval(s) split over ranks and summed with "
                "one NCCL reduce each")

    def step(self, main_stream, events=None, reduce_events=None):
        torch, engine, ctx = self.torch, self.engine, self.engine.ctx
        for band, occ, missed in self.bufs.values():
            band.zero_()
            occ.zero_()
            missed.zero_()
        done = {}
        for k, e in enumerate(self.mine):
            stream = engine.streams[k % len(engine.streams)]
            stream.wait_stream(main_stream)
            band, occ, missed = self.bufs[e["idx"]]
            if events is not None:
                ev0 = torch.cuda.Event(enable_timing=True)
                ev1 = torch.cuda.Event(enable_timing=True)
                ev0.record(stream)
            ctx.simulate_interval_device(self.p, e["abi_iv"], e["iv"].barriers,
                                         e["d_tasks"].data_ptr(), e["ntasks"], band.data_ptr(),
                                         occ.data_ptr(), e["d_stats"].data_ptr(),
                                         missed.data_ptr(), stream.cuda_stream)
            if events is not None:
                ev1.record(stream)
                events.append((e, ev0, ev1))
            if e["idx"] in self.split:
                ev = torch.cuda.Event()
                ev.record(stream)
                done.setdefault(e["idx"], []).append(ev)
        # the one exchange step: a split interval is summed onto its root as soon as ITS launches
        # are done, on the reduce stream, while the other intervals keep running
        for idx in self.split:
            for ev in done.get(idx, []):
                engine.reduce_stream.wait_event(ev)
            engine.reduce_stream.wait_stream(main_stream)  # (the memsets above)
            with torch.cuda.stream(engine.reduce_stream):
                if reduce_events is not None:
                    r0 = torch.cuda.Event(enable_timing=True)
                    r1 = torch.cuda.Event(enable_timing=True)
                    r0.record(engine.reduce_stream)
                for b in self.bufs[idx]:
                    self.dist.reduce(b, dst=self.roots[idx][0], op=self.dist.ReduceOp.SUM)
                if reduce_events is not None:
                    r1.record(engine.reduce_stream)
                    reduce_events.append((r0, r1))
        for stream in engine.streams + [engine.reduce_stream]:
            main_stream.wait_stream(stream)

    def work(self):
        """(lef_updates, contacts, epochs, faults, algorithmic bytes) of this rank's last step."""
        lu = contacts = epochs = faults = alg = 0
        for e in self.mine:
            st = e["d_stats"].cpu().numpy().view(self.stats_dt)
            lu += int(st["num_lef_updates"].sum())
            contacts += int(st["num_contacts"].sum())
            epochs += int(st["num_epochs"].sum())
            faults += int((st["device_fault"] != 0).sum())
            alg += 32 * int(st["num_lef_updates"].sum()) + \
                2 * len(e["iv"].barriers) * int(st["num_epochs"].sum())
        return lu, contacts, epochs, faults, alg

    def close(self):
        self.sim.close()


def band_checksums(torch, band, occ, missed):
    """Order-independent and position-weighted checksums of a device band (u32 bit pattern in an
    int32 tensor) and its 1D track."""
    b = band.to(torch.int64) & 0xFFFFFFFF
    w = (torch.arange(b.numel(), device=b.device, dtype=torch.int64) % 65521) + 1
    return {"band_sum": int(b.sum().item()), "band_weighted": int((b * w).sum().item()),
            "occ_sum": int(occ.sum().item()),
            "occ_weighted": int((occ * w[:occ.numel()]).sum().item()),
            "missed": int(missed.item())}


C3_GOLDEN = os.path.join(ROOT, "tests", "golden", "c3_chr1_8192cells_checksums.json")
C3_SAMPLED_CELLS = (0, 1, 1023, 1024, 4097, 6143, 8190, 8191)


def extra_c3(torch, dist, engine, rank, world, local_rank, main_stream, sync_all, all_reduce, args):
    """BASELINE config C3 (chr1 x 8192 cells): timed, and PROVEN -- the cells are split over the
    ranks (8 x 1024 at N = 8) and the band is summed with one NCCL reduce onto its root, whose
    checksums must equal the committed single-GPU ones (tests/golden/, written by
    `bench.py --workload c3 --write-c3-golden` on one GPU) at every N; 8 sampled cells are also
    run alone and compared with the CPU oracle bit for bit. Any mismatch ends the bench with a
    non-zero exit code. Reference: the shared-matrix accumulation of
    scheduler_simulate.cpp:129-159 + contact_matrix_dense_safe_impl.hpp:54-68."""
    from modle_b200 import abi, host

    cfg, genome, desc = build_workload("c3", args.c3_cells)
    job = DeviceJob(torch, dist, engine, cfg, genome, rank, world, local_rank, 0)
    ncells = int(cfg.params.num_cells)
    job.step(main_stream)  # warm-up
    sync_all()
    revents = []
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record(main_stream)
    job.step(main_stream, reduce_events=revents)
    t1.record(main_stream)
    sync_all()
    ms = all_reduce(t0.elapsed_time(t1), "max")
    reduce_ms = all_reduce(sum(a.elapsed_time(b) for a, b in revents) if revents else 0.0, "max")
    lu, contacts, epochs, faults, _ = job.work()
    total_lu = all_reduce(float(lu), "sum")
    total_contacts = all_reduce(float(contacts), "sum")
    total_faults = all_reduce(float(faults), "sum")
    out = {"workload": desc, "ms_per_step": ms, "value": total_lu / (ms * 1e-3), "unit": UNIT,
           "lef_updates_per_step": total_lu, "contacts_per_step": total_contacts,
           "parallelism": job.parallelism(), "reduce_ms": reduce_ms,
           "reduce_bytes_per_rank": int(job.reduce_bytes),
           "nvlink_bytes_per_step": int(job.reduce_bytes) * max(0, world - 1),
           "checks": {}}
    errors = []
    if total_faults:
        errors.append(f"{int(total_faults)} cells reported a device fault")
    iv = job.sim.intervals[0]
    p = cfg.params
    all_tasks = host.make_cell_tasks(p, iv.chrom_name, iv.abi_interval())
    if int(total_contacts) != int(all_tasks["num_target_contacts"].sum()):
        errors.append("contacts registered != sum of the cells' targets")
    root = job.roots[0][0]
    if rank == root:
        got = band_checksums(torch, *job.bufs[0])
        out["checks"]["checksums"] = got
        if got["band_sum"] + got["missed"] != int(total_contacts):
            errors.append("band sum + missed updates != contacts registered")
        if args.write_c3_golden:
            if world != 1:
                raise SystemExit("--write-c3-golden needs a single-GPU run")
            with open(C3_GOLDEN, "w") as f:
                json.dump({"workload": desc, "cells": ncells, "written_by": "bench.py --workload "
                           "c3 --write-c3-golden (1 GPU, unsharded)", **got}, f, indent=1)
        if os.path.exists(C3_GOLDEN):
            gold = json.load(open(C3_GOLDEN))
            if int(gold.get("cells", -1)) == ncells:
                bad = [k for k in got if got[k] != gold[k]]
                out["checks"]["equals_single_gpu_golden"] = not bad
                if bad:
                    errors.append("reduced band differs from the single-GPU golden in " + ",".join(bad))
            else:
                out["checks"]["equals_single_gpu_golden"] = None
        else:
            out["checks"]["equals_single_gpu_golden"] = None
    if rank == 0:
        # sampled cells alone, against the oracle (kernel == oracle at the full C3 geometry)
        from oracle import pyoracle

        sel = np.array([c for c in C3_SAMPLED_CELLS if c < ncells])
        tasks = np.ascontiguousarray(all_tasks[sel])
        band, occ, missed = engine.alloc_outputs(iv.nrows, iv.ncols)
        d_tasks = torch.from_numpy(tasks.view(np.uint8).reshape(-1).copy()).to(job.dev)
        d_stats = torch.zeros(len(tasks) * job.stats_dt.itemsize, dtype=torch.uint8, device=job.dev)
        engine.ctx.simulate_interval_device(p, iv.abi_interval(), iv.barriers, d_tasks.data_ptr(),
                                            len(tasks), band.data_ptr(), occ.data_ptr(),
                                            d_stats.data_ptr(), missed.data_ptr(),
                                            main_stream.cuda_stream)
        main_stream.synchronize()
        o_band, o_occ, o_stats, o_missed = pyoracle.simulate_interval(
            p, iv.abi_interval(), iv.barriers, tasks, nthreads=min(len(tasks), os.cpu_count() or 1))
        st = d_stats.cpu().numpy().view(job.stats_dt)
        same = (np.array_equal(band.cpu().numpy().view(np.uint32), o_band) and
                np.array_equal(occ.cpu().numpy().view(np.uint64)[:len(o_occ)], o_occ) and
                int(missed.item()) == o_missed and
                all(np.array_equal(st[f], o_stats[f]) for f in
                    ("num_contacts", "num_epochs", "num_burnin_epochs", "num_lef_updates",
                     "num_rng_draws")))
        out["checks"]["sampled_cells_equal_oracle"] = {"cells": [int(c) for c in sel], "ok": bool(same)}
        if not same:
            errors.append("sampled cells differ from the CPU oracle")
    nerr = all_reduce(float(len(errors)), "sum")
    out["checks"]["errors"] = errors
    job.close()
    return out, int(nerr), errors


def extra_register(torch, ctx, dev, peak):
    """The contact-register kernel in isolation at the C5 geometry (chr2, 1 kb bins: 3000 x
    242,194 pixels = 2.9 GB, density 1 = 726.6 M contacts), loop-like and uniform streams, with
    the figures SURVEY 8(d) defines: algorithmic GB/s (8 B per contact) and, for the path the
    library took (binned: count pass, scatter-by-tile pass, paced replay), the bytes that path
    cannot avoid. Next to it the calibration the survey asks for: plain random red.global.add.u32
    over the same footprint (independent kernel, no stream to read)."""
    nrows, ncols = 3000, 242_194
    npx = nrows * ncols + 1
    n = nrows * ncols
    g = torch.Generator(device=dev)
    g.manual_seed(20260117)
    stream = torch.cuda.Stream(device=dev)
    band = torch.zeros(npx, dtype=torch.int32, device=dev)
    missed = torch.zeros(1, dtype=torch.int64, device=dev)
    res = {"geometry": "C5: 3000 x 242,194 px (2.9 GB band), 726.6 M contacts (density 1)",
           "hbm_peak_GBps": peak, "streams": {}}

    def timed(fn, reps=3):
        ts = []
        with torch.cuda.stream(stream):
            for r in range(reps + 1):
                band.zero_()
                missed.zero_()
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                fn()
                e1.record(stream)
                stream.synchronize()
                if r >= 1:
                    ts.append(e0.elapsed_time(e1))
        return sorted(ts)[len(ts) // 2]

    # bytes the binned path cannot avoid: the (bin1, bin2) stream is read twice (count, scatter),
    # the 4-byte pixel indices are written once and read once, every band sector comes in and
    # goes out once
    compulsory = 2 * 8 * n + 2 * 4 * n + 2 * 4 * npx
    for kind in ("loop", "uniform"):
        b2 = torch.randint(0, ncols, (n,), device=dev, generator=g, dtype=torch.int64)
        if kind == "loop":
            d = torch.empty(n, device=dev, dtype=torch.float32).exponential_(1.0 / 100.0, generator=g)
            d = d.to(torch.int64).clamp_(0, nrows - 1)
        else:
            d = torch.randint(0, nrows, (n,), device=dev, generator=g, dtype=torch.int64)
        b1 = (b2 - d).clamp_(min=0).to(torch.int32).contiguous()
        b2 = b2.to(torch.int32).contiguous()
        del d
        ms = timed(lambda: ctx.register_contacts_device(b1.data_ptr(), b2.data_ptr(), n, nrows,
                                                        ncols, band.data_ptr(), missed.data_ptr(),
                                                        stream.cuda_stream))
        total = int((band.to(torch.int64) & 0xFFFFFFFF).sum().item()) + int(missed.item())
        rate = n / (ms * 1e-3)
        res["streams"][kind] = {
            "ms": ms, "contacts_per_s": rate, "all_contacts_accounted_for": total == n,
            "algorithmic_GBps": 8 * rate / 1e9, "algorithmic_frac_of_hbm": 8 * rate / 1e9 / peak,
            "compulsory_bytes_of_the_binned_path": compulsory,
            "compulsory_GBps": compulsory / (ms * 1e-3) / 1e9,
            "compulsory_frac_of_hbm": compulsory / (ms * 1e-3) / 1e9 / peak}
        del b1, b2
    # calibration: random reductions over the 2.9 GB footprint (HBM regime: every reduction is a
    # 32 B sector in and one out) and over the C1 footprint (30.9 MB: L2 regime)
    ms = timed(lambda: ctx.calibrate_red_device(band.data_ptr(), npx, n, 1, stream.cuda_stream))
    res["calibration_random_red_2.9GB"] = {
        "ms": ms, "reductions_per_s": n / (ms * 1e-3),
        "sector_traffic_GBps": 64 * n / (ms * 1e-3) / 1e9,
        "sector_traffic_frac_of_hbm": 64 * n / (ms * 1e-3) / 1e9 / peak}
    small = 600 * 12_889 + 1
    n2 = 1 << 28
    ms = timed(lambda: ctx.calibrate_red_device(band.data_ptr(), small, n2, 2, stream.cuda_stream))
    res["calibration_random_red_30.9MB_L2"] = {"ms": ms, "reductions_per_s": n2 / (ms * 1e-3)}
    return res


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from modle_b200 import build, distributed

    build.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; modle_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    cfg, genome, desc = build_workload(args.workload, args.cells)
    rng_mode = 1 if args.rng_mode == "throughput" else 0
    engine = distributed.DeviceEngine(local_rank, num_streams=args.streams, rng_mode=rng_mode)
    ctx = engine.ctx
    job = DeviceJob(torch, dist, engine, cfg, genome, rank, world, local_rank, rng_mode,
                    slice_all={"auto": None, "slices": True, "whole": False}[args.plan],
                    tolerance=args.tolerance,
                    pre_split=args.pre_split)
    sim = job.sim
    main_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(main_stream)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def all_reduce(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return float(t.item())

    def all_gather(x):
        if world == 1:
            return [x]
        t = torch.zeros(world, dtype=torch.float64, device=dev)
        t[rank] = x
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    def timed_steps(steps, warmup, events=None):
        for _ in range(warmup):
            job.step(main_stream)
        sync_all()
        t_start = torch.cuda.Event(enable_timing=True)
        t_end = torch.cuda.Event(enable_timing=True)
        sync_all()
        t_start.record(main_stream)
        for _ in range(steps):
            job.step(main_stream, events)
        t_end.record(main_stream)
        sync_all()
        mine_ms = t_start.elapsed_time(t_end)
        return all_reduce(mine_ms, "max"), mine_ms

    # ---- device-resident timing -------------------------------------------------------------
    for _ in range(args.warmup):
        job.step(main_stream)
    sync_all()
    launches0 = ctx.kernel_launches()
    ctx.phase_cycles(reset=True)
    events = []
    with ClockSampler(local_rank) as clocks:
        elapsed_ms, my_ms = timed_steps(args.steps, 0, events)
    launches = ctx.kernel_launches() - launches0
    phases = ctx.phase_cycles(reset=True)

    # work done per step (identical every step: the simulation is deterministic)
    lef_updates, contacts, epochs, faults, alg_bytes = job.work()
    if all_reduce(float(faults), "sum"):
        raise SystemExit(f"bench.py: {faults} cells reported a device fault")
    total_lu = all_reduce(float(lef_updates), "sum")
    total_contacts = all_reduce(float(contacts), "sum")
    value = total_lu * args.steps / (elapsed_ms * 1e-3)
    rank_ms = all_gather(my_ms / args.steps)       # each rank's own device time per step
    rank_lu = all_gather(float(lef_updates))       # ... and its share of the work
    launch_ms = [ev0.elapsed_time(ev1) for _, ev0, ev1 in events]
    per_interval = {}
    for e, ev0, ev1 in events:
        key = f"{e['iv'].chrom_name}[{e['cells'][0]}:{e['cells'][1]}]"
        per_interval.setdefault(key, []).append(ev0.elapsed_time(ev1))
    peak, peak_src = measured_peaks()
    # launches overlap (several streams), so the kernel's rate is taken over the timed region
    achieved = (alg_bytes * args.steps / 1e9) / (my_ms * 1e-3) if my_ms > 0 else 0.0
    dram_per_lu, traffic_src = ncu_traffic_per_lef_update(rng_mode)
    traffic = dram_per_lu * lef_updates / max(1, len(job.mine)) if dram_per_lu else None

    # ---- end to end through the public API (host buffers; copies inside the timed region) ------
    task_dt, stats_dt, barrier_dt = job.task_dt, job.stats_dt, job.barrier_dt
    h2d = sum(e["ntasks"] * task_dt.itemsize + len(e["iv"].barriers) * barrier_dt.itemsize
              for e in job.mine)
    d2h = sum((sim.intervals[i].nrows * sim.intervals[i].ncols + 1) * 4 + sim.intervals[i].ncols * 8
              + 8 for i in job.bufs if job.roots[i][0] == rank) + \
        sum(e["ntasks"] * stats_dt.itemsize for e in job.mine)

    def step_e2e():
        sim.run_simulate(num_workers=args.e2e_workers)

    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(2 if args.steps > 1 else 1):  # warm-up (contexts, staging buffers, page faults)
        step_e2e()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    sync_all()
    e2e_s = all_reduce(time.perf_counter() - t0, "max")
    e2e_value = total_lu * e2e_steps / e2e_s
    sim.close()  # (frees the e2e path's contexts and staging before the extras allocate theirs)

    # ---- extras: the second RNG mode, config C3 with its reduce, the register kernel -----------
    extra, failed = {}, []
    if not args.no_extras and args.workload == "c2" and rng_mode == 0:
        ctx.set_rng_mode(1)
        thr_ms, _ = timed_steps(2, 1)
        ctx.set_rng_mode(0)
        lu_t, _, _, faults_t, _ = job.work()
        thr_lu = all_reduce(float(lu_t), "sum")
        extra["throughput_mode"] = {
            "value": thr_lu * 2 / (thr_ms * 1e-3), "unit": UNIT, "ms_per_step": thr_ms / 2,
            "steps": 2, "warmup": 1, "lef_updates_per_step": thr_lu,
            "rng_mode": "throughput (counter-based draws; statistically equivalent, "
                        "tests/test_zz_gpu_throughput_mode.py holds the gate)"}
        if all_reduce(float(faults_t), "sum"):
            failed.append("throughput mode: device fault")
    if (not args.no_extras and args.workload == "c2") or args.write_c3_golden:
        extra["c3"], nerr, errs = extra_c3(torch, dist, engine, rank, world, local_rank,
                                           main_stream, sync_all, all_reduce, args)
        if nerr:
            failed.append("c3: " + "; ".join(errs) if errs else "c3: check failed on another rank")
    if not args.no_extras and rank == 0 and world == 1 and args.workload == "c2":
        for band, occ, missed in job.bufs.values():  # make room: 4 x 2.9 GB for the C5 replay
            del band, occ, missed
        job.bufs.clear()
        torch.cuda.empty_cache()
        extra["register"] = extra_register(torch, ctx, dev, peak)
        if not all(v["all_contacts_accounted_for"] for v in extra["register"]["streams"].values()):
            failed.append("register: contacts lost")

    # ---- CPU baseline (rank 0, single-GPU runs only) ------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        _, _, cpu = cpu_arm(args.workload, args.cells, os.cpu_count() or 1, 12.0, 1, 1)

    if rank == 0:
        tot_cyc = max(1, phases.get("total", 1))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32/u64 integer state, f64 samplers", "data": "synthetic",
            "config": {"workload": desc, "lef_updates_per_step": total_lu,
                       "contacts_per_step": total_contacts,
                       "l2_policy": "band matrices + RNG staging of a step exceed the 126 MB L2 "
                                    "for the genome-wide workload; bands are re-zeroed every step",
                       "streams": args.streams,
                       "rank_ms_per_step": [round(x, 1) for x in rank_ms],
                       "rank0_launch_ms": {k: round(sum(v) / len(v), 1)
                                           for k, v in per_interval.items()},
                       "rank_lef_updates_per_step": rank_lu,
                       "rng_mode": "deterministic (reference draw order, bit-exact)"
                                   if rng_mode == 0 else
                                   "throughput (counter-based draws, statistically equivalent)",
                       "parallelism": job.parallelism() +
                       ("; extra.c3: " + extra["c3"]["parallelism"] if "c3" in extra else "")},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                    "ms_per_step": 1e3 * e2e_s / e2e_steps,
                    "api": "Simulation.run_simulate -> modle_b200_simulate_interval (host buffers)"
                    if world == 1 else
                    "Simulation.run_simulate -> device shards + NCCL reduce + D2H on roots"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_simulate_cells", "achieved": achieved,
                         "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": traffic,
                         "traffic_source": traffic_src,
                         "avg_launch_ms": sum(launch_ms) / max(1, len(launch_ms)),
                         "algorithmic_bytes_per_step": int(alg_bytes),
                         "note": "cell state is shared-memory resident by design, so this kernel "
                                 "is issue/latency bound, not HBM bound; algorithmic bytes = 32 B "
                                 "per LEF-update + 2 B per barrier-epoch (SURVEY 8d); launches "
                                 "overlap on several streams, so the rate is taken over the "
                                 "whole timed region of rank 0"},
            "phase_share": {k: round(v / tot_cyc, 4) for k, v in phases.items() if k != "total"},
            "cycles_per_cell_epoch": tot_cyc / max(1, epochs * args.steps),
            "clocks": clocks.summary(),
        }
        if extra:
            line["extra"] = extra
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if failed:
            line["failed_checks"] = failed
        print(json.dumps(line), flush=True)
    job.close()
    engine.close()
    if world > 1:
        dist.destroy_process_group()
    if failed:
        raise SystemExit("bench.py: result check failed: " + " | ".join(failed))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--cells", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--rng-mode", default="deterministic", choices=["deterministic", "throughput"],
                    help="deterministic: the reference's draw order, bit-exact (the headline); "
                         "throughput: counter-based draws, statistically equivalent")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip extra.throughput_mode / extra.c3 / extra.register")
    ap.add_argument("--c3-cells", type=int, default=8192, help="cells of the extra C3 run")
    ap.add_argument("--write-c3-golden", action="store_true",
                    help="(1 GPU) run the extra C3 step and write its checksums to "
                         "tests/golden/c3_chr1_8192cells_checksums.json")
    ap.add_argument("--plan", default="auto", choices=["auto", "whole", "slices"],
                    help="whole: whole intervals per rank, cells split only to balance; slices: "
                         "every interval cut into one cell slice per rank + one reduce each; auto "
                         "(default): slices when a slice still holds >= 222 cells (1.5 waves)")
    ap.add_argument("--e2e-workers", type=int, default=6,
                    help="host worker threads (one context each) of the end-to-end run at 1 GPU")
    ap.add_argument("--tolerance", type=float, default=1.10,
                    help="planner: split cells while the heaviest rank exceeds this x the mean")
    ap.add_argument("--pre-split", type=int, default=1,
                    help="planner: cut every interval into this many cell ranges before dealing")
    ap.add_argument("--streams", type=int, default=3,
                    help="concurrent launches per GPU (streams / host worker threads)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
