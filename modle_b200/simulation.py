"""Host-side mirror of the slice of modle::Simulation that drives the hot path.

Reference (paths relative to the reference checkout):
  Simulation(Config) / run_simulate()          src/libmodle/cpu/simulation.cpp:93-115,
                                               src/libmodle/cpu/scheduler_simulate.cpp:43-170
  Task fan-out (seed, targets, jump per cell)  scheduler_simulate.cpp:104-160
  worker -> simulate_one_cell                  scheduler_simulate.cpp:190-271

Only the boundary is reproduced: the per-cell work itself runs in the CUDA library through the
C ABI (include/modle_b200.h). Names follow the reference (Config fields, GenomicInterval,
run_simulate); there is no CPU fallback.
"""
import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

from . import abi, host


class Config:
    """modle::Config restricted to the fields the path reads (simulation_config.hpp:53-113).

    Attribute access is forwarded to the underlying C struct, so names are the reference's.
    `transform()` applies Cli::transform_args (src/modle/cli.cpp:993-1016).
    """

    def __init__(self, **overrides):
        object.__setattr__(self, "_p", host.default_params())
        object.__setattr__(self, "_given", set())
        for k, v in overrides.items():
            setattr(self, k, v)

    def __getattr__(self, name):
        return getattr(object.__getattribute__(self, "_p"), name)

    def __setattr__(self, name, value):
        p = object.__getattribute__(self, "_p")
        if not hasattr(p, name):
            raise AttributeError(f"modle::Config has no field '{name}' on the simulated path")
        setattr(p, name, value)
        object.__getattribute__(self, "_given").add(name)

    def transform(self):
        g = object.__getattribute__(self, "_given")
        host.transform_params(self._p, "rev_extrusion_speed" in g, "fwd_extrusion_speed" in g,
                              "extrusion_barrier_occupancy" in g)
        return self

    @property
    def params(self):
        return object.__getattribute__(self, "_p")


@dataclass
class GenomicInterval:
    """modle::GenomicInterval (src/libmodle/internal/include/modle/genome.hpp:125-195)."""
    chrom_name: str
    chrom_size: int
    start: int
    end: int
    barriers: np.ndarray = None  # abi barrier dtype, sorted by pos
    num_lefs: int = 0
    nrows: int = 0
    ncols: int = 0
    nrows_lazy: int = 0  # ceil(diagonal_width / bin_size), not clamped to ncols (see npixels)
    contacts: np.ndarray = None          # band, reference layout (nrows*ncols+1 uint32)
    lef_1d_occupancy: np.ndarray = None  # ncols uint64
    missed_updates: int = 0
    stats: np.ndarray = None

    def size(self):
        return self.end - self.start

    def npixels(self):
        """GenomicInterval::npixels (genome_impl.hpp:21,96): ContactMatrixLazy's ncols x nrows,
        whose nrows is NOT clamped to ncols (the contact target of an interval shorter than the
        diagonal width is computed from this product)."""
        return self.ncols * self.nrows_lazy

    def abi_interval(self):
        return abi.Interval(self.chrom_size, self.start, self.end, self.num_lefs)


RNG_REFERENCE_ORDER = 0  # MODLE_B200_RNG_REFERENCE_ORDER
RNG_COUNTER = 1          # MODLE_B200_RNG_COUNTER ("throughput mode")


class Context:
    """RAII wrapper of modle_b200_context (one per GPU)."""

    def __init__(self, device=0, rng_mode=None):
        self._h = C.c_void_p()
        host.check(host.lib().modle_b200_init(C.byref(self._h), int(device)))
        self.device = int(device)
        if rng_mode is None:  # measurement knob for the scripts/ 
            rng_mode = int(os.environ.get("MODLE_B200_RNG_MODE", "0"))
        if rng_mode:
            self.set_rng_mode(rng_mode)

    def set_rng_mode(self, mode):
        """RNG_REFERENCE_ORDER (0, default): deterministic mode, bit-identical to the reference
        algorithm; RNG_COUNTER (1): throughput mode (see modle_b200_set_rng_mode)."""
        host.check(host.lib().modle_b200_set_rng_mode(self._h, int(mode)))

    @property
    def rng_mode(self):
        return int(host.lib().modle_b200_get_rng_mode(self._h))

    def close(self):
        if self._h:
            host.lib().modle_b200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def synchronize(self):
        host.check(host.lib().modle_b200_synchronize(self._h))

    def reserve(self, max_nrows, max_ncols, max_cells):
        """Sizes the buffers of the host-buffer calls once for the largest interval to come."""
        host.check(host.lib().modle_b200_reserve(self._h, int(max_nrows), int(max_ncols),
                                                 int(max_cells)))

    def phase_cycles(self, reset=True):
        """{phase name: SM-clock cycles} summed over the cells simulated since the last reset."""
        out = (C.c_uint64 * len(host.PHASE_NAMES))()
        host.check(host.lib().modle_b200_phase_cycles(self._h, out, len(host.PHASE_NAMES),
                                                      1 if reset else 0))
        return dict(zip(host.PHASE_NAMES, [int(x) for x in out]))

    def kernel_launches(self):
        return int(host.lib().modle_b200_kernel_launches(self._h))

    # -- HOST buffers in, HOST buffers out (the reference-facing call) -----------------------
    def simulate_interval(self, params, interval, barriers, tasks, band=None, occ1d=None,
                          log_capacity_per_cell=0):
        """With log_capacity_per_cell > 0 the internal-state log (Simulation::dump_stats,
        simulation.cpp:995-1056) is returned as a fifth element: records[cell][epoch]."""
        _, _, stats_dt = abi.np_dtypes()
        nrows, ncols = host.band_shape(params, int(interval.end - interval.start))
        # no buffers given: the results are written into fresh ones (overwrite entry point, no
        # zeroing and no add pass); buffers given: the results are ADDED to them
        fresh = band is None and occ1d is None and not log_capacity_per_cell
        if band is None:
            band = (np.empty if fresh else np.zeros)(nrows * ncols + 1, dtype=np.uint32)
        if occ1d is None:
            occ1d = (np.empty if fresh else np.zeros)(ncols, dtype=np.uint64)
        stats = np.zeros(len(tasks), dtype=stats_dt)
        missed = C.c_uint64(0)
        barriers = np.ascontiguousarray(barriers)
        tasks = np.ascontiguousarray(tasks)
        if log_capacity_per_cell:
            log = np.zeros((len(tasks), int(log_capacity_per_cell)), dtype=abi.epoch_record_dtype())
            rc = host.lib().modle_b200_simulate_interval_logged(
                self._h, C.byref(params), C.byref(interval),
                barriers.ctypes.data if len(barriers) else None, len(barriers), tasks.ctypes.data,
                len(tasks), band.ctypes.data, occ1d.ctypes.data, stats.ctypes.data,
                C.byref(missed), log.ctypes.data, int(log_capacity_per_cell))
            host.check(rc)
            return band, occ1d, stats, int(missed.value), log
        fn = host.lib().modle_b200_simulate_interval_overwrite if fresh else \
            host.lib().modle_b200_simulate_interval
        rc = fn(self._h, C.byref(params), C.byref(interval),
                barriers.ctypes.data if len(barriers) else None, len(barriers), tasks.ctypes.data,
                len(tasks), band.ctypes.data, occ1d.ctypes.data, stats.ctypes.data, C.byref(missed))
        host.check(rc)
        return band, occ1d, stats, int(missed.value)

    # -- device-resident buffers (raw device pointers, e.g. torch tensors' data_ptr()) --------
    def simulate_interval_device(self, params, interval, barriers_host, d_tasks, num_cells, d_band,
                                 d_occ1d, d_stats, d_missed, stream=None):
        barriers_host = np.ascontiguousarray(barriers_host)
        rc = host.lib().modle_b200_simulate_interval_device(
            self._h, C.byref(params), C.byref(interval),
            barriers_host.ctypes.data if len(barriers_host) else None, len(barriers_host),
            C.c_void_p(d_tasks), num_cells, C.c_void_p(d_band),
            C.c_void_p(d_occ1d) if d_occ1d else None, C.c_void_p(d_stats) if d_stats else None,
            C.c_void_p(d_missed), C.c_void_p(stream) if stream else None)
        host.check(rc)

    def register_contacts_device(self, d_bin1, d_bin2, n, nrows, ncols, d_band, d_missed,
                                 stream=None):
        rc = host.lib().modle_b200_register_contacts_device(
            self._h, C.c_void_p(d_bin1), C.c_void_p(d_bin2), n, nrows, ncols, C.c_void_p(d_band),
            C.c_void_p(d_missed), C.c_void_p(stream) if stream else None)
        host.check(rc)

    def calibrate_red_device(self, d_band, num_words, num_reductions, seed=0, stream=None):
        """Plain random red.global.add.u32 over d_band[0..num_words): the calibration of the
        contact-register kernel's atomic roofline (the caller times it)."""
        host.check(host.lib().modle_b200_calibrate_red_device(
            self._h, C.c_void_p(d_band), num_words, num_reductions, seed,
            C.c_void_p(stream) if stream else None))

    # -- band -> sorted COO pixels (the .cool writer hand-off) ---------------------------------
    def band_to_pixels(self, band, nrows, ncols, bin_offset=0):
        """append_contact_matrix_to_cooler's pixel loop (contact_matrix_dense_io_impl.hpp:50-71)
        on the GPU: host band in, array of ThinPixel<int32>-layout records out."""
        band = np.ascontiguousarray(band, dtype=np.uint32)
        if band.size < nrows * ncols + 1:
            raise ValueError("band buffer smaller than nrows*ncols+1")
        n = C.c_uint64(0)
        L = host.lib()
        host.check(L.modle_b200_band_to_pixels(self._h, band.ctypes.data, nrows, ncols, bin_offset,
                                               None, 0, C.byref(n)))
        out = np.zeros(int(n.value), dtype=abi.pixel_dtype())
        if n.value:
            host.check(L.modle_b200_band_to_pixels(self._h, band.ctypes.data, nrows, ncols,
                                                   bin_offset, out.ctypes.data, len(out),
                                                   C.byref(n)))
        return out

    def count_pixels_device(self, d_band, nrows, ncols, d_row_offsets, stream=None):
        host.check(host.lib().modle_b200_count_pixels_device(
            self._h, C.c_void_p(d_band), nrows, ncols, C.c_void_p(d_row_offsets),
            C.c_void_p(stream) if stream else None))

    def fill_pixels_device(self, d_band, nrows, ncols, bin_offset, d_row_offsets, d_pixels,
                           capacity, stream=None):
        host.check(host.lib().modle_b200_fill_pixels_device(
            self._h, C.c_void_p(d_band), nrows, ncols, bin_offset, C.c_void_p(d_row_offsets),
            C.c_void_p(d_pixels) if d_pixels else None, capacity,
            C.c_void_p(stream) if stream else None))

    def lef_occupancy_profile(self, occ1d):
        """write_lef_occupancy_to_bwig's normalisation (simulation.cpp:170-197) on the GPU."""
        occ1d = np.ascontiguousarray(occ1d, dtype=np.uint64)
        out = np.zeros(len(occ1d), dtype=np.float32)
        host.check(host.lib().modle_b200_lef_occupancy_profile(
            self._h, occ1d.ctypes.data, len(occ1d), out.ctypes.data))
        return out

    def snapshot_cell(self, params, interval, barriers, task):
        n = int(interval.num_lefs)
        nb = len(barriers)
        arrs = {k: np.zeros(n, dtype=np.uint64)
                for k in ("rev_pos", "fwd_pos", "binding_epoch", "rev_ranks", "fwd_ranks")}
        arrs["barrier_active"] = np.zeros(max(nb, 1), dtype=np.uint8)
        snap = abi.CellSnapshot()
        for k, v in arrs.items():
            ptr_t = C.POINTER(C.c_uint8 if k == "barrier_active" else C.c_uint64)
            setattr(snap, k, v.ctypes.data_as(ptr_t))
        st = abi.CellStats()
        barriers = np.ascontiguousarray(barriers)
        task = np.ascontiguousarray(task)
        rc = host.lib().modle_b200_snapshot_cell(
            self._h, C.byref(params), C.byref(interval),
            barriers.ctypes.data if nb else None, nb, task.ctypes.data, C.byref(snap),
            C.byref(st))
        host.check(rc)
        arrs["barrier_active"] = arrs["barrier_active"][:nb]
        arrs["num_active_lefs"] = int(snap.num_active_lefs)
        arrs["burnin_completed"] = int(snap.burnin_completed)
        arrs["stats"] = {f: int(getattr(st, f)) for f, _ in abi.CellStats._fields_}
        return arrs


@dataclass
class Simulation:
    """modle::Simulation: owns the Config and the genome, `run_simulate()` runs every interval.

    `genome` is a list of (chrom_name, chrom_size, start, end, barrier_records) where
    barrier_records are (pos, strand, score) tuples (what Genome's ctor derives from the
    chrom.sizes file, the optional genomic-intervals BED and the barrier BED,
    src/libmodle/internal/genome.cpp:299-330).

    Work is sharded over `world_size` processes (one per GPU) by whole intervals, heaviest
    first; `rank` processes only its own intervals.
    """
    config: Config
    genome: list
    device: int = 0
    rank: int = 0
    world_size: int = 1
    rng_mode: int = RNG_REFERENCE_ORDER  # RNG_COUNTER selects the throughput mode
    slice_all: object = None  # multi-GPU plan: True = one cell slice of EVERY interval per rank,
    #                           False = whole intervals, None = the library decides
    intervals: list = field(default_factory=list)
    _ctxs: list = field(default_factory=list, repr=False)
    _engine: object = field(default=None, repr=False)

    def __post_init__(self):
        p = self.config.params
        for name, size, start, end, records in self.genome:
            iv = GenomicInterval(name, size, start, end)
            # `records` are the barriers the genome importer ASSIGNED to this interval (BED records
            # that overlap it): one whose midpoint falls outside [start, end) is kept, like the
            # reference keeps it (genome.cpp:285-294) -- unreachable, but it takes its draws --
            # so this mirror and modle_b200_genome_import feed the kernel the same barrier set.
            iv.barriers = host.barriers_from_records(list(records), p)
            iv.num_lefs = host.compute_num_lefs(p, end - start)
            iv.nrows, iv.ncols = host.band_shape(p, end - start)
            iv.nrows_lazy = (int(p.diagonal_width) + int(p.bin_size) - 1) // int(p.bin_size)
            self.intervals.append(iv)

    def run_simulate(self, ctx=None, num_workers=3):
        """Simulation::run_simulate (scheduler_simulate.cpp:43-170) for this process' share.

        world_size == 1: every interval goes through the host-buffer C-ABI call
        (modle_b200_simulate_interval). `num_workers` host threads, each with its own context
        (stream + device buffers), pull intervals heaviest-first from a queue -- the analogue of
        the reference's worker threads -- so the tail of one interval's cells and its
        device->host copy overlap the next interval's kernel.
        world_size > 1: shards come from distributed.plan_shards; bands stay on the device, split
        intervals are reduced onto their root rank (NCCL), and each root copies its intervals out.
        """
        p = self.config.params
        if self.world_size > 1:
            return self._run_simulate_sharded()
        # intervals without barriers are skipped (scheduler_simulate.cpp:111-124)
        todo = sorted((i for i, iv in enumerate(self.intervals) if len(iv.barriers)),
                      key=lambda i: -self.intervals[i].num_lefs)

        def one(c, idx):
            # every call starts from fresh output buffers (a second run_simulate() replaces the
            # results, band / occupancy / missed / stats alike; the sharded path does the same)
            iv = self.intervals[idx]
            tasks = host.make_cell_tasks(p, iv.chrom_name, iv.abi_interval())
            iv.contacts, iv.lef_1d_occupancy, iv.stats, iv.missed_updates = \
                c.simulate_interval(p, iv.abi_interval(), iv.barriers, tasks)

        if ctx is not None or not todo:
            for idx in todo:
                one(ctx, idx)
            return self.intervals
        nw = max(1, min(num_workers, len(todo)))
        while len(self._ctxs) < nw:  # contexts (streams, device buffers) persist across calls
            c = Context(self.device, self.rng_mode)
            big = max((self.intervals[i] for i in todo), key=lambda iv: iv.nrows * iv.ncols)
            c.reserve(big.nrows, big.ncols, int(p.num_cells))  # any worker may get the largest
            self._ctxs.append(c)
        if nw == 1:
            for idx in todo:
                one(self._ctxs[0], idx)
            return self.intervals

        import queue
        import threading

        q = queue.SimpleQueue()
        for idx in todo:
            q.put(idx)
        errors = []

        def worker(c):
            try:
                while not errors:
                    try:
                        idx = q.get_nowait()
                    except queue.Empty:
                        break
                    one(c, idx)
            except Exception as e:  # re-raised on the caller's thread, like ContextManager does
                errors.append(e)

        threads = [threading.Thread(target=worker, args=(self._ctxs[k],)) for k in range(nw)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return self.intervals

    def _run_simulate_sharded(self):
        import torch.distributed as dist

        from . import distributed

        p = self.config.params
        if self._engine is None:
            self._engine = distributed.DeviceEngine(self.device, rng_mode=self.rng_mode)
        shards = distributed.plan_shards(distributed.interval_weights(self.intervals),
                                         int(p.num_cells), self.world_size,
                                         slice_all=self.slice_all)
        out = distributed.run_sharded(self._engine, p, self.intervals, self.rank, self.world_size,
                                      dist, shards=shards)
        for idx, o in out.items():
            iv = self.intervals[idx]
            iv.stats = np.concatenate(o["stats"]) if o["stats"] else None
            if o["root"] != self.rank:
                continue
            band, occ, missed = o["host"]  # pinned staging, copied while other intervals ran
            iv.contacts = band.numpy().view(np.uint32)
            iv.lef_1d_occupancy = occ.numpy().view(np.uint64)[:iv.ncols]
            iv.missed_updates = int(missed[0])
        return self.intervals

    def close(self):
        for c in self._ctxs:
            c.close()
        self._ctxs.clear()
        if self._engine is not None:
            self._engine.close()
            self._engine = None
