#!/usr/bin/env python3
"""Benchmark of the loop-extrusion hot path (see BASELINE.md / SURVEY.md 8d).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c2|c1|c3]

One "step" = one whole simulation of the workload (default: BASELINE config C2, genome-wide
GRCh38 shape, 24 chromosomes x 512 cells, default parameters, synthetic barriers). Metric:
LEF-updates/s (1 LEF-update = one active LEF carried through one epoch, burn-in included).

  value  device-resident: cell tasks / band matrices live in HBM, kernels only (+ band memset)
  e2e    the reference-facing call modle_b200_simulate_interval with HOST buffers
         (H2D of tasks and barriers, D2H of band / 1D occupancy / stats inside the timed region)
  roofline, cpu_baseline, clocks, gpu_launches: see DESIGN.md "Measurement"

N > 1 (launched by torch.distributed.run, one rank per GPU): whole chromosomes are dealt to
ranks heaviest-first (no data-path collective is needed for C2; a chromosome whose cells are
split over ranks -- workload c3 -- is summed with one NCCL reduce). Total work is fixed, so
"scaling" is "strong".

--impl reference: the reference's CPU implementation of the path cannot be built offline
(no Boost/xoshiro-cpp/...), so this arm times the oracle restatement
(oracle/liboracle.so) on the host cores over a bounded sample of the same workload.
"""
import argparse
import json
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


def build_workload(name, cells):
    from modle_b200 import workloads

    if name == "c1":
        cfg, genome = workloads.config_c1(cells or 512)
        desc = "C1 chr20 shape, %d cells, 5 kb bins, default parameters"
    elif name == "c3":
        cfg, genome = workloads.config_c3(cells or 8192)
        desc = "C3 chr1 shape, %d cells, 5 kb bins, default parameters"
    else:
        cfg, genome = workloads.config_c2(cells or 512)
        desc = "C2 GRCh38 genome-wide shape (24 chromosomes, 38,815 synthetic barriers), " \
               "%d cells, 5 kb bins, default parameters"
    return cfg, genome, desc % int(cfg.num_cells)


# ------------------------------------------------------------------------------- CPU (oracle)
def oracle_sample_run(cfg, genome, cells_per_interval, nthreads_total):
    """Simulates the first `cells_per_interval` cells of every interval with the CPU oracle,
    intervals in parallel. Returns (lef_updates, seconds, cores_used)."""
    from concurrent.futures import ThreadPoolExecutor

    from modle_b200 import abi, host
    from oracle import pyoracle

    p = cfg.params
    jobs = []
    for name, size, start, end, recs in genome:
        iv = abi.Interval(size, start, end, host.compute_num_lefs(p, end - start))
        bars = host.barriers_from_records(recs, p)
        tasks = host.make_cell_tasks(p, name, iv)[:cells_per_interval]
        jobs.append((iv, bars, tasks))
    per_job = max(1, min(cells_per_interval, nthreads_total // max(1, len(jobs))
                         if len(jobs) < nthreads_total else 1))
    workers = max(1, min(len(jobs), nthreads_total // per_job))
    pyoracle.lib()
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=workers) as ex:
        res = list(ex.map(lambda j: pyoracle.simulate_interval(p, j[0], j[1], j[2],
                                                               nthreads=per_job), jobs))
    dt = time.perf_counter() - t0
    lu = sum(int(r[2]["num_lef_updates"].sum()) for r in res)
    return lu, dt, min(nthreads_total, workers * per_job)


def run_reference(args, rank, world):
    if rank != 0:
        return
    cfg, genome, desc = build_workload(args.workload, args.cells)
    cores = os.cpu_count() or 1
    # bounded sample: every interval, a few cells each (work scales with the core count)
    cpi = max(1, min(int(cfg.num_cells), -(-2 * cores // max(1, len(genome)))))
    sample = f"first {cpi} cell(s) of each of the {len(genome)} intervals of the workload per step"
    for _ in range(args.warmup):
        oracle_sample_run(cfg, genome, cpi, cores)
    lu_tot, t_tot, used = 0, 0.0, cores
    for _ in range(args.steps):
        lu, dt, used = oracle_sample_run(cfg, genome, cpi, cores)
        lu_tot += lu
        t_tot += dt
    value = lu_tot / t_tot
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int64/f64",
        "data": "synthetic", "config": {"workload": desc, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference cannot be built offline (SURVEY 8c); oracle restatement timed instead",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- GPU
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from modle_b200 import abi, build, host
    from modle_b200.simulation import Context, Simulation

    build.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; modle_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    cfg, genome, desc = build_workload(args.workload, args.cells)
    p = cfg.params
    sim = Simulation(cfg, genome, device=local_rank, rank=rank, world_size=world)
    split_cells = len(sim.intervals) < world  # e.g. c3: one chromosome, cells split over ranks
    owner = sim.partition()
    ctx = Context(local_rank)
    barrier_dt, task_dt, stats_dt = abi.np_dtypes()

    # ---- stage my share of the work on the device ------------------------------------------
    mine = []
    for idx, iv in enumerate(sim.intervals):
        if len(iv.barriers) == 0:
            continue
        tasks = host.make_cell_tasks(p, iv.chrom_name, iv.abi_interval())
        if split_cells:
            tasks = tasks[rank::world]
        elif owner[idx] != rank:
            continue
        if len(tasks) == 0:
            continue
        npx = iv.nrows * iv.ncols + 1
        h_tasks = torch.from_numpy(tasks.view(np.uint8).reshape(-1).copy()).pin_memory()
        entry = dict(
            iv=iv, abi_iv=iv.abi_interval(), ntasks=len(tasks), h_tasks_np=tasks,
            h_tasks=h_tasks, d_tasks=h_tasks.to(dev),
            d_band=torch.zeros(npx, dtype=torch.int32, device=dev),
            d_occ=torch.zeros(iv.ncols, dtype=torch.int64, device=dev),
            d_stats=torch.zeros(len(tasks) * stats_dt.itemsize, dtype=torch.uint8, device=dev),
            d_missed=torch.zeros(1, dtype=torch.int64, device=dev))
        mine.append(entry)
    mine.sort(key=lambda e: -e["iv"].num_lefs)
    # a real (non-default) stream: the C ABI treats a NULL stream handle as "the context's own
    # stream", and the timing events must sit on the stream the kernels are launched on
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0

    def step_device(events=None):
        for e in mine:
            e["d_band"].zero_()
            e["d_occ"].zero_()
            e["d_missed"].zero_()
            if events is not None:
                ev0 = torch.cuda.Event(enable_timing=True)
                ev1 = torch.cuda.Event(enable_timing=True)
                ev0.record(stream)
            ctx.simulate_interval_device(p, e["abi_iv"], e["iv"].barriers, e["d_tasks"].data_ptr(),
                                         e["ntasks"], e["d_band"].data_ptr(),
                                         e["d_occ"].data_ptr(), e["d_stats"].data_ptr(),
                                         e["d_missed"].data_ptr(), stream.cuda_stream)
            if events is not None:
                ev1.record(stream)
                events.append((e, ev0, ev1))
        if split_cells and world > 1:
            for e in mine:
                dist.reduce(e["d_band"], dst=0, op=dist.ReduceOp.SUM)  # u32 sum == i32 sum mod 2^32
                dist.reduce(e["d_occ"], dst=0, op=dist.ReduceOp.SUM)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- device-resident timing -------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    sync_all()
    launches0 = ctx.kernel_launches()
    events = []
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        sync_all()
        t_start.record(stream)
        for _ in range(args.steps):
            step_device(events)
        t_end.record(stream)
        sync_all()
    elapsed_ms = max_over_ranks(t_start.elapsed_time(t_end))
    launches = ctx.kernel_launches() - launches0

    # work done per step (identical every step: the simulation is deterministic)
    lef_updates = 0
    alg_bytes = 0
    contacts = 0
    faults = 0
    for e in mine:
        st = e["d_stats"].cpu().numpy().view(stats_dt)
        lef_updates += int(st["num_lef_updates"].sum())
        contacts += int(st["num_contacts"].sum())
        faults += int((st["device_fault"] != 0).sum())
        e["alg_bytes"] = 32 * int(st["num_lef_updates"].sum()) + \
            2 * len(e["iv"].barriers) * int(st["num_epochs"].sum())
        alg_bytes += e["alg_bytes"]
    if faults:
        raise SystemExit(f"bench.py: {faults} cells reported a device fault")
    total_lu = sum_over_ranks(float(lef_updates))
    value = total_lu * args.steps / (elapsed_ms * 1e-3)
    kernel_ms = sum(ev0.elapsed_time(ev1) for _, ev0, ev1 in events)
    peak, peak_src = measured_peaks()
    achieved = (alg_bytes * args.steps / 1e9) / (kernel_ms * 1e-3) if kernel_ms > 0 else 0.0

    # ---- end to end through the host-buffer C ABI call ---------------------------------------
    h2d = sum(e["ntasks"] * task_dt.itemsize + len(e["iv"].barriers) * barrier_dt.itemsize
              for e in mine)
    d2h = sum((e["iv"].nrows * e["iv"].ncols + 1) * 4 + e["iv"].ncols * 8 +
              e["ntasks"] * stats_dt.itemsize + 8 for e in mine)
    host_bands = [np.zeros(e["iv"].nrows * e["iv"].ncols + 1, dtype=np.uint32) for e in mine]
    host_occ = [np.zeros(e["iv"].ncols, dtype=np.uint64) for e in mine]

    def step_e2e():
        for e, hb, ho in zip(mine, host_bands, host_occ):
            hb.fill(0)
            ho.fill(0)
            ctx.simulate_interval(p, e["abi_iv"], e["iv"].barriers, e["h_tasks_np"], band=hb,
                                  occ1d=ho)

    e2e_steps = max(1, min(args.steps, 2))
    step_e2e()  # warm-up (buffers, page faults)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    sync_all()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = total_lu * e2e_steps / e2e_s

    # ---- CPU baseline (rank 0, single-GPU runs only) ------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        cpi = max(1, min(int(cfg.num_cells), -(-2 * cores // max(1, len(genome)))))
        lu, dt, used = oracle_sample_run(cfg, genome, cpi, cores)
        cpu = {"value": lu / dt, "unit": UNIT, "cores": used, "kind": "port",
               "sample": f"first {cpi} cell(s) of each of the {len(genome)} intervals "
                         f"({lu} LEF-updates, {dt:.1f} s)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32/u64 integer state, f64 samplers", "data": "synthetic",
            "config": {"workload": desc, "lef_updates_per_step": total_lu,
                       "contacts_per_step": sum_over_ranks(float(contacts)) if world == 1 else None,
                       "l2_policy": "band matrices (1.48 GB) + RNG staging exceed the 126 MB L2; "
                                    "band is re-zeroed every step",
                       "parallelism": f"{world} rank(s), " +
                       ("cells of one chromosome split over ranks + NCCL reduce" if split_cells
                        else "whole chromosomes dealt heaviest-first")},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                    "ms_per_step": 1e3 * e2e_s / e2e_steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_simulate_cells", "achieved": achieved,
                         "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": None,
                         "kernel_ms_per_step": kernel_ms / args.steps,
                         "algorithmic_bytes_per_step": int(alg_bytes),
                         "note": "state is shared-memory resident by design; algorithmic bytes = "
                                 "32 B per LEF-update + 2 B per barrier-epoch (SURVEY 8d)"},
            "clocks": clocks.summary(),
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3"])
    ap.add_argument("--cells", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
