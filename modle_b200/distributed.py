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
import ctypes as C
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


# The planner itself lives in the C ABI (modle_b200_plan_shards / modle_b200_cell_weight,
# csrc/shards.cpp) so that a C++ host and this Python mirror deal the same shards; what follows are
# thin wrappers. Cost model, for the record: the SM time of one cell-epoch is A[threads per CTA] +
# B[threads] x LEFs cycles, divided by the cells an SM hosts at a time. One CTA of 1024 threads:
# 80 k + 40 per LEF, fitted to single-interval runs on a B200 (profiles/r01i_stream_sweep.txt:
# chr1 279 k, chr2 270 k, chr3 238 k cycles per cell-epoch); a LEF costs a narrower CTA
# proportionally more (212 k at 512 threads / ~2070 LEFs, 217 k at 256 threads / ~1280 LEFs).
# Only the ratios matter to the planner.
def cell_cost(num_lefs, num_barriers):
    """Relative SM time of one cell of an interval (epoch counts are the same for every interval
    to within a few percent, so they drop out)."""
    L = host.lib()
    L.modle_b200_cell_weight.restype = C.c_double
    L.modle_b200_cell_weight.argtypes = [C.c_uint64, C.c_uint64]
    return float(L.modle_b200_cell_weight(int(num_lefs), int(num_barriers))) if num_lefs > 0 else 0.0


def interval_weights(intervals):
    """Planner weights of simulation.GenomicInterval-like objects (0 = skipped: no barriers)."""
    return [cell_cost(iv.num_lefs, len(iv.barriers)) if len(iv.barriers) else 0.0
            for iv in intervals]


def plan_shards(num_lefs, num_cells, world, tolerance=1.10, slice_all=False, pre_split=1):
    # slice_all: False (whole intervals), True (one slice per rank) or None (the library decides)
    """Deals (interval, cell range) shards to `world` ranks (modle_b200_plan_shards).

    num_lefs[i] is the weight of one cell of interval i (0 = interval skipped, e.g. no
    barriers): its LEF count, or better `interval_weights()` -- the SM time of a cell-epoch, which
    also knows how many cells of that size share an SM; the cost of a shard is weight x cells.
    Starts from whole intervals and, while the heaviest rank carries more than `tolerance` x the
    mean load, halves the heaviest splittable piece of that rank. Deterministic: every rank
    computes the same plan.

    slice_all: every interval is cut into `world` equal cell ranges instead, one per rank (the
    first range of interval i goes to rank i % world, so the roots -- and with them the reduce
    destinations and the device->host copies -- rotate over the ranks). Every rank then runs the
    same mix of work, which removes both the imbalance between ranks and most of the tail of a
    rank's last launches, at the price of one reduce per interval.
    """
    if pre_split > 1 and world > 1 and slice_all is False and num_cells >= pre_split:
        # every interval is first cut into `pre_split` equal cell ranges, which are then dealt out
        # like intervals of their own (finer grain for the balance, more launches per rank)
        k = int(pre_split)
        bounds = [num_cells * j // k for j in range(k + 1)]
        out = []
        for j in range(k):
            sub = plan_shards(num_lefs, bounds[j + 1] - bounds[j], 1, tolerance)
            out += [Shard(s.interval, s.cell_lo + bounds[j], s.cell_hi + bounds[j], -1, s.weight)
                    for s in sub]
        load = [0.0] * world
        for s in sorted(out, key=lambda s: (-s.weight, s.interval, s.cell_lo)):
            s.rank = min(range(world), key=lambda r: (load[r], r))
            load[s.rank] += s.weight
        out.sort(key=lambda s: (s.interval, s.cell_lo))
        merged = []  # neighbouring ranges of one interval on one rank run as one launch
        for s in out:
            if merged and merged[-1].interval == s.interval and merged[-1].rank == s.rank and \
                    merged[-1].cell_hi == s.cell_lo:
                merged[-1].cell_hi = s.cell_hi
                merged[-1].weight += s.weight
            else:
                merged.append(s)
        return merged
    L = host.lib()
    L.modle_b200_plan_shards.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.c_int, C.c_int,
                                         C.c_double, C.c_void_p, C.c_size_t,
                                         C.POINTER(C.c_size_t)]
    w = np.ascontiguousarray(num_lefs, dtype=np.float64)
    n = C.c_size_t(0)
    args = (w.ctypes.data, len(w), int(num_cells), int(world),
            -1 if slice_all is None else int(bool(slice_all)), float(tolerance))
    host.check(L.modle_b200_plan_shards(*args, None, 0, C.byref(n)))
    out = np.zeros(n.value, dtype=abi.shard_dtype())
    if n.value:
        host.check(L.modle_b200_plan_shards(*args, out.ctypes.data, len(out), C.byref(n)))
    return [Shard(int(s["interval"]), int(s["cell_lo"]), int(s["cell_hi"]), int(s["rank"]),
                  float(s["weight"])) for s in out]


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
    (torch is the allocator / stream / collective plumbing here, nothing more).

    Besides the launch streams it owns a reduce stream and a copy stream: an interval's reduce
    and its device->host copy (into pinned staging that persists across calls) are queued behind
    that interval's own launches only, so they overlap the kernels of the intervals still
    running."""

    def __init__(self, device=0, num_streams=3, rng_mode=0):
        import torch

        from .simulation import Context

        self.torch = torch
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self.ctx = Context(device, rng_mode)
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(max(1, num_streams))]
        self.reduce_stream = torch.cuda.Stream(device=self.device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._next = 0
        self._pinned = {}   # (tag, dtype, numel) -> pinned host tensor, reused across calls

    def close(self):
        self.torch.cuda.synchronize(self.device)
        self._pinned.clear()
        self.ctx.close()

    def alloc_outputs(self, nrows, ncols):
        t = self.torch
        band = t.zeros(nrows * ncols + 1, dtype=t.int32, device=self.device)   # u32 bit pattern
        occ = t.zeros(max(ncols, 1), dtype=t.int64, device=self.device)        # u64 bit pattern
        missed = t.zeros(1, dtype=t.int64, device=self.device)
        return band, occ, missed

    def run(self, params, abi_interval, barriers, tasks, band, occ, missed):
        """Asynchronous: adds the cells in `tasks` into band / occ / missed; returns the device
        tensor that will hold the per-cell stats and a token (CUDA event) that fires when the
        launch is done."""
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
            done = t.cuda.Event()
            done.record(stream)
        return d_stats, done

    def join(self):
        cur = self.torch.cuda.current_stream(self.device)
        for s in self.streams + [self.reduce_stream, self.copy_stream]:
            cur.wait_stream(s)

    # -- hooks of run_sharded ------------------------------------------------------------------
    def reduce(self, bufs, root, tokens, dist):
        """Sums `bufs` onto `root` (one NCCL reduce per buffer) once `tokens` have fired; returns
        the token of the reduce."""
        t = self.torch
        for ev in tokens:
            self.reduce_stream.wait_event(ev)
        with t.cuda.stream(self.reduce_stream):
            for b in bufs:
                dist.reduce(b, dst=root, op=dist.ReduceOp.SUM)
                b.record_stream(self.reduce_stream)
            done = t.cuda.Event()
            done.record(self.reduce_stream)
        return done

    def to_host(self, tag, tensors, tokens):
        """Queues device->host copies of `tensors` into pinned staging behind `tokens`; returns
        the host tensors (valid after finish())."""
        t = self.torch
        for ev in tokens:
            self.copy_stream.wait_event(ev)
        out = []
        with t.cuda.stream(self.copy_stream):
            for k, x in enumerate(tensors):
                key = (tag, k, x.dtype, x.numel())
                h = self._pinned.get(key)
                if h is None:
                    h = t.empty(x.numel(), dtype=x.dtype, pin_memory=True)
                    self._pinned[key] = h
                h.copy_(x, non_blocking=True)
                x.record_stream(self.copy_stream)
                out.append(h)
        return out

    def finish(self):
        self.join()
        self.torch.cuda.current_stream(self.device).synchronize()


class HostEngineHooks:
    """Synchronous versions of the DeviceEngine hooks for engines that compute on the host (the
    CPU tests plug the kernel emulation in through these)."""

    def reduce(self, bufs, root, tokens, dist):
        for b in bufs:
            dist.reduce(b, dst=root, op=dist.ReduceOp.SUM)
        return None

    def to_host(self, tag, tensors, tokens):
        return list(tensors)

    def finish(self):
        pass


def run_sharded(engine, params, intervals, rank=0, world=1, dist=None, shards=None):
    """Simulates this rank's shards, reduces split intervals onto their roots and brings every
    interval this rank is the root of to the host.

    `intervals`: objects with chrom_name, abi_interval(), barriers, num_lefs, nrows, ncols (see
    simulation.GenomicInterval). Returns {interval index: dict(band, occ1d, missed, stats, root,
    host)}: band / occ1d / missed are tensors on the engine's device (complete on the root rank
    only), `host` = (band, occ1d, missed) as host tensors on the root (None elsewhere), `stats`
    numpy records per shard. `dist` is torch.distributed (already initialised) when world > 1.

    Nothing here waits for the whole rank: an interval's reduce is queued behind that interval's
    launches, its device->host copy behind its reduce (or launches), both on their own streams.
    """
    if shards is None:
        shards = plan_shards(interval_weights(intervals), int(params.num_cells), world,
                             slice_all=None)
    roots = interval_roots(shards)
    _, _, stats_dt = abi.np_dtypes()
    out = {}
    # heaviest intervals first, so that the short ones fill the tail
    order = sorted(shards, key=lambda s: (-s.weight, s.interval, s.cell_lo))
    for s in order:
        if s.rank != rank:
            continue
        iv = intervals[s.interval]
        if s.interval not in out:
            band, occ, missed = engine.alloc_outputs(iv.nrows, iv.ncols)
            out[s.interval] = dict(band=band, occ1d=occ, missed=missed, stats=[], cells=[],
                                   tokens=[])
        o = out[s.interval]
        tasks = host.make_cell_tasks(params, iv.chrom_name, iv.abi_interval())[s.cell_lo:s.cell_hi]
        d_stats, token = engine.run(params, iv.abi_interval(), iv.barriers, tasks, o["band"],
                                    o["occ1d"], o["missed"])
        o["stats"].append(d_stats)
        o["cells"].append((s.cell_lo, s.cell_hi))
        if token is not None:
            o["tokens"].append(token)
    # Every rank walks the intervals in the same order (that of their first launch in the global
    # plan) and joins one reduce per buffer of a split interval on the world group; a rank that
    # holds no piece of it contributes zeros (no sub-communicators needed, and the volume is
    # trivial next to NVLink bandwidth).
    seen = []
    for s in order:
        if s.interval not in seen:
            seen.append(s.interval)
    for idx in seen:
        root, ranks = roots[idx]
        o = out.get(idx)
        tokens = list(o["tokens"]) if o is not None else []
        if world > 1 and len(ranks) > 1:
            iv = intervals[idx]
            bufs = (o["band"], o["occ1d"], o["missed"]) if o is not None else \
                engine.alloc_outputs(iv.nrows, iv.ncols)
            tok = engine.reduce(bufs, root, tokens, dist)
            tokens = [tok] if tok is not None else []
        if o is not None:
            o["root"] = root
            o["host"] = engine.to_host(idx, (o["band"], o["occ1d"], o["missed"]), tokens) \
                if root == rank else None
    engine.finish()
    for o in out.values():
        o["stats"] = [np.frombuffer(x.cpu().numpy().tobytes(), dtype=stats_dt) for x in o["stats"]]
        del o["tokens"]
    check_device_faults(out, rank, world, dist, engine)
    return out


def check_device_faults(out, rank=0, world=1, dist=None, engine=None):
    """A cell that reports a device fault stopped early, so its interval's band is incomplete:
    the host-buffer call returns MODLE_B200_ERR_DEVICE_FAULT for that, and so does the sharded
    path. The flag is MAX-all-reduced so that every rank raises (no rank is left waiting in a
    later collective with partial results in hand)."""
    worst, where = 0, None
    for idx, o in out.items():
        for st, (lo, _) in zip(o["stats"], o["cells"]):
            bad = np.nonzero(st["device_fault"])[0]
            if len(bad) and worst == 0:
                worst = int(st["device_fault"][bad[0]])
                where = (idx, lo + int(bad[0]))
    code = worst
    if world > 1:
        import torch

        # (host engines of the CPU tests run over gloo: the flag lives on the CPU there)
        t = torch.tensor([worst], dtype=torch.int64, device=getattr(engine, "device", "cpu"))
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        code = int(t.item())
    if code != 0:
        msg = f"device fault {code}"
        msg += f" in cell {where[1]} of interval {where[0]} on rank {rank}" if where else \
            " reported by another rank"
        raise host.ModleB200Error(abi.ERR_DEVICE_FAULT, msg + "; the results are incomplete")
