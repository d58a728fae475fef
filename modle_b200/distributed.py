"""Multi-GPU sharding of the hot path: one process per GPU, (interval, cell-batch) shards.

The reference has one level of parallelism: independent (interval, cell) tasks popped by worker
threads, all adding into the interval's shared contact matrix
(src/libmodle/cpu/scheduler_simulate.cpp:104-160,190-271). Here the same tasks are dealt to the
ranks of a `torch.distributed` job:

  * whole intervals go to ranks heaviest-first (no data-path collective at all);
  * only when that leaves the ranks unbalanced (few or very unequal intervals, e.g. one
    chromosome with 8192 cells) the heaviest pieces are split by cells, and each interval whose
    cells span several ranks has its band matrix (u32) and 1D occupancy (u64) summed onto the
    interval's root rank with ONE reduce each (NCCL over NVLink on GPUs; integer sums commute, so
    the result does not depend on the split).

Every cell keeps the RNG state and contact target it has in the unsharded run (they are
computed for all cells of the interval and then sliced), so outputs are identical for any world
size. The compute engine is pluggable: `DeviceEngine` drives the CUDA library; the CPU tests plug
an engine backed by the kernel emulation to exercise this module over gloo.
"""
from dataclasses import dataclass

import numpy as np

from . import abi, host


@dataclass
class Shard:
    interval: int   # index into the interval list
    cell_lo: int
    cell_hi: int    # cells [cell_lo, cell_hi) of that interval
    rank: int = -1
    weight: float = 0.0


# SM time of one cell-epoch, in cycles: A[threads per CTA] + B[threads] x LEFs, divided by the
# cells an SM hosts at a time. One CTA of 1024 threads: 80 k + 40 per LEF, fitted to
# single-interval runs on a B200 (profiles/r01i_stream_sweep.txt: chr1 279 k, chr2 270 k, chr3
# 238 k cycles per cell-epoch). A LEF costs a narrower CTA proportionally more (each thread owns
# more of them), and the constants then follow from the measured class averages (212 k at 512
# threads / ~2070 LEFs, 217 k at 256 threads / ~1280 LEFs). Only the ratios matter to the planner.
_COST_A = {1024: 80e3, 512: 46e3, 256: 12e3}
_COST_B = {1024: 40.0, 512: 80.0, 256: 160.0}


def cell_epoch_cycles(num_lefs, num_barriers):
    """(modelled SM-clock cycles of one cell-epoch, cells resident per SM) for an interval."""
    threads, per_sm, _ = host.launch_geometry(num_lefs, num_barriers)
    return _COST_A.get(threads, 80e3) + _COST_B.get(threads, 40.0) * num_lefs, per_sm


def cell_cost(num_lefs, num_barriers):
    """Relative SM time of one cell of an interval (epoch counts are the same for every interval
    to within a few percent, so they drop out)."""
    if num_lefs <= 0:
        return 0.0
    cycles, per_sm = cell_epoch_cycles(num_lefs, num_barriers)
    return cycles / per_sm


def interval_weights(intervals):
    """Planner weights of simulation.GenomicInterval-like objects (0 = skipped: no barriers)."""
    return [cell_cost(iv.num_lefs, len(iv.barriers)) if len(iv.barriers) else 0.0
            for iv in intervals]


def _assign(pieces, world):
    """Longest-processing-time-first assignment; returns the per-rank loads."""
    load = [0.0] * world
    for p in sorted(pieces, key=lambda s: (-s.weight, s.interval, s.cell_lo)):
        r = min(range(world), key=lambda k: (load[k], k))
        p.rank = r
        load[r] += p.weight
    return load


def plan_shards(num_lefs, num_cells, world, tolerance=1.10, max_pieces=None, slice_all=False):
    """Deals (interval, cell range) shards to `world` ranks.

    num_lefs[i] is the weight of one cell of interval i (0 = interval skipped, e.g. no
    barriers): its LEF count, or better `interval_weights()` -- the SM time of a cell-epoch, which
    also knows how many cells of that size share an SM; the cost of a shard is weight x cells. Starts from whole intervals and, while the heaviest rank
    carries more than `tolerance` x the mean load, halves the heaviest splittable piece of that
    rank. Deterministic: every rank computes the same plan.

    slice_all: every interval is cut into `world` equal cell ranges instead, one per rank (the
    first range of interval i goes to rank i % world, so the roots -- and with them the reduce
    destinations and the device->host copies -- rotate over the ranks). Every rank then runs the
    same mix of work, which removes both the imbalance between ranks and most of the tail of a
    rank's last launches, at the price of one reduce per interval.
    """
    if slice_all and world > 1:
        out = []
        for i, n in enumerate(num_lefs):
            if n <= 0 or num_cells <= 0:
                continue
            for k in range(world):
                lo, hi = num_cells * k // world, num_cells * (k + 1) // world
                if hi > lo:
                    out.append(Shard(i, lo, hi, (i + k) % world, float(n) * (hi - lo)))
        return out
    pieces = [Shard(i, 0, num_cells, -1, float(n) * num_cells)
              for i, n in enumerate(num_lefs) if n > 0 and num_cells > 0]
    if not pieces:
        return []
    if max_pieces is None:
        max_pieces = len(pieces) + 8 * world
    total = sum(p.weight for p in pieces)
    while True:
        load = _assign(pieces, world)
        worst = max(range(world), key=lambda k: (load[k], -k))
        if world == 1 or load[worst] <= tolerance * total / world or len(pieces) >= max_pieces:
            break
        cand = [p for p in pieces if p.rank == worst and p.cell_hi - p.cell_lo >= 2]
        if not cand:
            break
        p = max(cand, key=lambda s: (s.weight, -s.interval, -s.cell_lo))
        mid = (p.cell_lo + p.cell_hi) // 2
        per_cell = p.weight / (p.cell_hi - p.cell_lo)
        q = Shard(p.interval, mid, p.cell_hi, -1, per_cell * (p.cell_hi - mid))
        p.cell_hi = mid
        p.weight = per_cell * (mid - p.cell_lo)
        pieces.append(q)
    return sorted(pieces, key=lambda s: (s.interval, s.cell_lo))


def interval_roots(shards):
    """{interval: (root rank, [ranks holding a piece])}; root = owner of the first cell range."""
    out = {}
    for s in shards:
        root, ranks = out.get(s.interval, (s.rank, []))
        if s.rank not in ranks:
            ranks.append(s.rank)
        out[s.interval] = (root, ranks)
    return out


class DeviceEngine:
    """Runs shards on one GPU through the device-resident C-ABI call; buffers are torch tensors
    (torch is the allocator / stream / collective plumbing here, nothing more)."""

    def __init__(self, device=0, num_streams=3, rng_mode=0):
        import torch

        from .simulation import Context

        self.torch = torch
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self.ctx = Context(device, rng_mode)
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(max(1, num_streams))]
        self._next = 0

    def close(self):
        self.torch.cuda.synchronize(self.device)
        self.ctx.close()

    def alloc_outputs(self, nrows, ncols):
        t = self.torch
        band = t.zeros(nrows * ncols + 1, dtype=t.int32, device=self.device)   # u32 bit pattern
        occ = t.zeros(max(ncols, 1), dtype=t.int64, device=self.device)        # u64 bit pattern
        missed = t.zeros(1, dtype=t.int64, device=self.device)
        return band, occ, missed

    def run(self, params, abi_interval, barriers, tasks, band, occ, missed):
        """Asynchronous: adds the cells in `tasks` into band / occ / missed; returns the device
        tensor that will hold the per-cell stats and the stream the work was queued on."""
        t = self.torch
        _, _, stats_dt = abi.np_dtypes()
        stream = self.streams[self._next % len(self.streams)]
        self._next += 1
        stream.wait_stream(t.cuda.current_stream(self.device))
        with t.cuda.stream(stream):
            h = t.from_numpy(np.ascontiguousarray(tasks).view(np.uint8).reshape(-1).copy())
            d_tasks = h.pin_memory().to(self.device, non_blocking=True)
            d_stats = t.zeros(len(tasks) * stats_dt.itemsize, dtype=t.uint8, device=self.device)
            self.ctx.simulate_interval_device(
                params, abi_interval, barriers, d_tasks.data_ptr(), len(tasks), band.data_ptr(),
                occ.data_ptr(), d_stats.data_ptr(), missed.data_ptr(), stream.cuda_stream)
            for x in (d_tasks, d_stats, band, occ, missed):
                x.record_stream(stream)
        return d_stats, stream

    def join(self):
        cur = self.torch.cuda.current_stream(self.device)
        for s in self.streams:
            cur.wait_stream(s)


def run_sharded(engine, params, intervals, rank=0, world=1, dist=None, shards=None):
    """Simulates this rank's shards and reduces split intervals onto their roots.

    `intervals`: objects with chrom_name, abi_interval(), barriers, num_lefs, nrows, ncols (see
    simulation.GenomicInterval). Returns {interval index: dict(band, occ1d, missed, stats)} with
    torch tensors on the engine's device; an interval's band/occ1d are complete on its root rank
    only. `dist` is torch.distributed (already initialised) when world > 1.
    """
    if shards is None:
        shards = plan_shards(interval_weights(intervals), int(params.num_cells), world)
    roots = interval_roots(shards)
    _, _, stats_dt = abi.np_dtypes()
    out = {}
    # heaviest intervals first, so that the short ones fill the tail
    mine = sorted((s for s in shards if s.rank == rank), key=lambda s: (-s.weight, s.interval))
    for s in mine:
        iv = intervals[s.interval]
        if s.interval not in out:
            band, occ, missed = engine.alloc_outputs(iv.nrows, iv.ncols)
            out[s.interval] = dict(band=band, occ1d=occ, missed=missed, stats=[], cells=[])
        o = out[s.interval]
        tasks = host.make_cell_tasks(params, iv.chrom_name, iv.abi_interval())[s.cell_lo:s.cell_hi]
        d_stats, _ = engine.run(params, iv.abi_interval(), iv.barriers, tasks, o["band"],
                                o["occ1d"], o["missed"])
        o["stats"].append(d_stats)
        o["cells"].append((s.cell_lo, s.cell_hi))
    engine.join()
    if world > 1:
        # Every rank walks the split intervals in the same order and joins one reduce per buffer
        # on the world group; a rank that holds no piece of the interval contributes zeros (the
        # volume is trivial next to NVLink bandwidth, and no sub-communicators are needed).
        for idx in sorted(roots):
            root, ranks = roots[idx]
            if len(ranks) < 2:
                continue
            iv = intervals[idx]
            o = out.get(idx)
            bufs = (o["band"], o["occ1d"], o["missed"]) if o is not None else \
                engine.alloc_outputs(iv.nrows, iv.ncols)
            for b in bufs:
                dist.reduce(b, dst=root, op=dist.ReduceOp.SUM)
    for idx, o in out.items():
        o["root"] = roots[idx][0]
        o["stats"] = [np.frombuffer(x.cpu().numpy().tobytes(), dtype=stats_dt) for x in o["stats"]]
    return out
