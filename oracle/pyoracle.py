"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of the CPU oracle (oracle/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import
this module; the product package modle_b200 never does.
"""
import ctypes as C
import os

import numpy as np

from modle_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


CXXFLAGS = ["-O3", "-march=x86-64-v3", "-ffp-contract=off", "-fno-fast-math", "-std=c++17", "-fPIC",
            "-Wall", "-Wextra", "-pthread", "-shared"]  # keep in step with oracle/Makefile


def build(force=False):
    """Builds liboracle.so when its sources changed (content hash; safe under concurrent callers,
    see modle_b200/buildutil.py). Same compiler flags as oracle/Makefile."""
    from modle_b200 import buildutil

    so = os.path.join(_HERE, "liboracle.so")
    deps = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "oracle_sim.hpp", "oracle_rng.hpp", "ziggurat_tables.inc")]
    deps.append(os.path.join(_HERE, "..", "include", "modle_b200.h"))
    cxx = os.environ.get("CXX", "g++")
    return buildutil.ensure_built(
        so, deps, lambda tmp: [cxx] + CXXFLAGS + ["-o", tmp, os.path.join(_HERE, "oracle_capi.cpp")],
        extra=" ".join(CXXFLAGS), force=force)


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        u64p = C.POINTER(C.c_uint64)
        L.oracle_rng_next.restype = C.c_uint64
        L.oracle_rng_next.argtypes = [u64p]
        L.oracle_rng_seed.argtypes = [C.c_uint64, u64p]
        L.oracle_rng_jump.argtypes = [u64p]
        L.oracle_rng_discard.argtypes = [u64p, C.c_uint64]
        L.oracle_xxh3_64.argtypes = [C.c_char_p, C.c_size_t, C.c_uint64, u64p]
        L.oracle_interval_hash.argtypes = [C.c_char_p, C.c_size_t, C.c_uint64, C.c_uint64,
                                           C.c_uint64, C.c_uint64, u64p]
        L.oracle_compute_num_lefs.restype = C.c_uint64
        L.oracle_compute_num_lefs.argtypes = [C.c_double, C.c_uint64]
        L.oracle_make_cell_tasks.argtypes = [C.POINTER(abi.SimParams), C.c_char_p, C.c_size_t,
                                             C.POINTER(abi.Interval), C.c_void_p]
        L.oracle_sample.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, u64p, C.c_size_t,
                                    C.c_void_p, u64p]
        L.oracle_zig_tables.argtypes = [C.c_void_p] * 4
        L.oracle_rank_lefs.argtypes = [C.c_void_p] * 5 + [C.c_size_t, C.c_int]
        L.oracle_collision_steps.argtypes = (
            [C.c_uint32, C.c_uint64, C.c_uint64, C.c_size_t] + [C.c_void_p] * 9 +
            [C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double,
             C.c_uint64, C.c_void_p])
        L.oracle_simulate_interval.argtypes = [
            C.POINTER(abi.SimParams), C.POINTER(abi.Interval), C.c_void_p, C.c_size_t, C.c_void_p,
            C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, u64p, C.c_int]
        L.oracle_snapshot_cell.argtypes = [
            C.POINTER(abi.SimParams), C.POINTER(abi.Interval), C.c_void_p, C.c_size_t, C.c_void_p,
            C.POINTER(abi.CellSnapshot), C.POINTER(abi.CellStats)]
        _LIB = L
    return _LIB


def _state(st):
    return (C.c_uint64 * 4)(*[int(x) for x in st])


def rng_seed(seed):
    st = (C.c_uint64 * 4)()
    lib().oracle_rng_seed(seed, st)
    return [int(x) for x in st]


def rng_next(st):
    s = _state(st)
    v = lib().oracle_rng_next(s)
    return int(v), [int(x) for x in s]


def rng_jump(st):
    s = _state(st)
    lib().oracle_rng_jump(s)
    return [int(x) for x in s]


def rng_discard(st, n):
    s = _state(st)
    lib().oracle_rng_discard(s, n)
    return [int(x) for x in s]


def xxh3_64(data: bytes, seed=0):
    out = C.c_uint64()
    rc = lib().oracle_xxh3_64(data, len(data), seed, C.byref(out))
    if rc != 0:
        raise ValueError("unsupported input length for the oracle's XXH3 restatement")
    return int(out.value)


def interval_hash(name: str, chrom_size, start, end, seed):
    out = C.c_uint64()
    b = name.encode()
    rc = lib().oracle_interval_hash(b, len(b), chrom_size, start, end, seed, C.byref(out))
    if rc != 0:
        raise ValueError("interval hash failed")
    return int(out.value)


def compute_num_lefs(lefs_per_mbp, size_bp):
    return int(lib().oracle_compute_num_lefs(lefs_per_mbp, size_bp))


def make_cell_tasks(params, name, interval):
    _, task_dt, _ = abi.np_dtypes()
    tasks = np.zeros(int(params.num_cells), dtype=task_dt)
    b = name.encode()
    rc = lib().oracle_make_cell_tasks(C.byref(params), b, len(b), C.byref(interval),
                                      tasks.ctypes.data)
    if rc != 0:
        raise ValueError("oracle_make_cell_tasks failed")
    return tasks


SAMPLE_KINDS = {"bernoulli": 0, "canonical": 1, "uniform01": 2, "uniform_int": 3,
                "unit_normal": 4, "normal": 5, "unit_exponential": 6, "poisson": 7,
                "binomial": 8, "gev": 9, "raw": 10}


def sample(kind, n, state, p0=0.0, p1=0.0, p2=0.0):
    s = _state(state)
    out = np.zeros(n, dtype=np.float64)
    draws = C.c_uint64()
    lib().oracle_sample(SAMPLE_KINDS[kind], p0, p1, p2, s, n, out.ctypes.data, C.byref(draws))
    return out, [int(x) for x in s], int(draws.value)


def zig_tables():
    nx, ny = np.zeros(129), np.zeros(129)
    ex, ey = np.zeros(257), np.zeros(257)
    lib().oracle_zig_tables(nx.ctypes.data, ny.ctypes.data, ex.ctypes.data, ey.ctypes.data)
    return nx, ny, ex, ey


def rank_lefs(rev, fwd, ep, rr, fr, init_buffers=False):
    rev, fwd, ep = (np.ascontiguousarray(a, dtype=np.uint64) for a in (rev, fwd, ep))
    rr = np.array(rr, dtype=np.uint64)
    fr = np.array(fr, dtype=np.uint64)
    lib().oracle_rank_lefs(rev.ctypes.data, fwd.ctypes.data, ep.ctypes.data, rr.ctypes.data,
                           fr.ctypes.data, len(rev), int(init_buffers))
    return rr, fr


STEP = {"adjust": 1, "clamp": 2, "boundaries": 4, "lef_bar": 8, "primary": 16,
        "correct_lef_bar": 32, "correct_primary": 64, "secondary": 128, "fix_secondary": 256}


def collision_steps(steps, start, end, rev, fwd, ep, rr, fr, rm, fm, bar_pos, bar_dir, bar_active,
                    prob_bypass=0.0, pblock_major=1.0, pblock_minor=0.0, rng_seed=0,
                    rc=None, fc=None):
    """Runs the selected reference sub-steps (bitmask or list of STEP names) on a copy of the
    given state; returns a dict with the updated arrays."""
    if not isinstance(steps, int):
        m = 0
        for s in steps:
            m |= STEP[s]
        steps = m
    n = len(rev)
    a = {k: np.array(v, dtype=np.uint64) for k, v in
         dict(rev=rev, fwd=fwd, ep=ep, rr=rr, fr=fr, rm=rm, fm=fm).items()}
    a["rc"] = np.zeros(n, dtype=np.uint32) if rc is None else np.array(rc, dtype=np.uint32)
    a["fc"] = np.zeros(n, dtype=np.uint32) if fc is None else np.array(fc, dtype=np.uint32)
    bp = np.array(bar_pos, dtype=np.uint64)
    bd = np.array(bar_dir, dtype=np.uint8)
    ba = np.array(bar_active, dtype=np.uint8)
    info = np.zeros(3, dtype=np.uint64)
    lib().oracle_collision_steps(
        steps, start, end, n, a["rev"].ctypes.data, a["fwd"].ctypes.data, a["ep"].ctypes.data,
        a["rr"].ctypes.data, a["fr"].ctypes.data, a["rm"].ctypes.data, a["fm"].ctypes.data,
        a["rc"].ctypes.data, a["fc"].ctypes.data, len(bp), bp.ctypes.data, bd.ctypes.data,
        ba.ctypes.data, prob_bypass, pblock_major, pblock_minor, rng_seed, info.ctypes.data)
    a["n5"], a["n3"], a["ndraws"] = (int(x) for x in info)
    return a


def simulate_interval(params, interval, barriers, tasks, nthreads=1, want_occ=True,
                      log_capacity_per_cell=0):
    """CPU counterpart of modle_b200.simulate_interval. Returns (band, occ1d, stats, missed)
    and, with log_capacity_per_cell > 0, the internal-state log records[cell][epoch]."""
    from modle_b200.host import band_shape  # pure-host geometry helper (no GPU code)

    _, _, stats_dt = abi.np_dtypes()
    nrows, ncols = band_shape(params, int(interval.end - interval.start))
    band = np.zeros(nrows * ncols + 1, dtype=np.uint32)
    occ = np.zeros(ncols, dtype=np.uint64)
    stats = np.zeros(len(tasks), dtype=stats_dt)
    missed = C.c_uint64(0)
    barriers = np.ascontiguousarray(barriers)
    tasks = np.ascontiguousarray(tasks)
    if log_capacity_per_cell:
        log = np.zeros((len(tasks), int(log_capacity_per_cell)), dtype=abi.epoch_record_dtype())
        L = lib()
        L.oracle_simulate_interval_logged.argtypes = [
            C.POINTER(abi.SimParams), C.POINTER(abi.Interval), C.c_void_p, C.c_size_t, C.c_void_p,
            C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.c_int,
            C.c_void_p, C.c_size_t]
        rc = L.oracle_simulate_interval_logged(
            C.byref(params), C.byref(interval), barriers.ctypes.data, len(barriers),
            tasks.ctypes.data, len(tasks), band.ctypes.data, occ.ctypes.data if want_occ else None,
            stats.ctypes.data, C.byref(missed), int(nthreads), log.ctypes.data,
            int(log_capacity_per_cell))
        assert rc == 0
        return band, occ, stats, int(missed.value), log
    rc = lib().oracle_simulate_interval(
        C.byref(params), C.byref(interval), barriers.ctypes.data, len(barriers),
        tasks.ctypes.data, len(tasks), band.ctypes.data, occ.ctypes.data if want_occ else None,
        stats.ctypes.data, C.byref(missed), int(nthreads))
    assert rc == 0
    return band, occ, stats, int(missed.value)


def genome_jobs(overrides, genome, cells_per_interval=None):
    """Inputs of a whole run built by oracle-side code only (oracle/pyparams.py + the oracle's own
    task fan-out): (params, [(abi.Interval, barriers, tasks)]). `genome` is the plain list
    modle_b200.workloads.spec() returns; intervals without barriers are skipped like the reference
    does (scheduler_simulate.cpp:111-124)."""
    from . import pyparams

    p = pyparams.make_params(**overrides)
    jobs = []
    for name, size, start, end, recs in genome:
        bars = pyparams.barriers_from_records([r for r in recs], p)
        if len(bars) == 0:
            continue
        iv = abi.Interval(size, start, end, pyparams.compute_num_lefs(p, end - start))
        tasks = make_cell_tasks(p, name, iv)
        if cells_per_interval is not None:
            tasks = tasks[:cells_per_interval]
        jobs.append((iv, bars, np.ascontiguousarray(tasks)))
    return p, jobs


def simulate_genome(params, jobs, nthreads=1):
    """All (interval, cell) tasks of `jobs` through ONE work queue served by `nthreads` workers
    (the reference's scheduling, scheduler_simulate.cpp:104-160,190-271). Returns a list of
    (band, occ1d, stats, missed) per job."""
    from . import pyparams

    _, _, stats_dt = abi.np_dtypes()
    n = len(jobs)
    ivs = (abi.Interval * n)(*[j[0] for j in jobs])
    bands, occs, stats = [], [], []
    for iv, _, tasks in jobs:
        nrows, ncols = pyparams.band_shape(params, int(iv.end - iv.start))
        bands.append(np.zeros(nrows * ncols + 1, dtype=np.uint32))
        occs.append(np.zeros(ncols, dtype=np.uint64))
        stats.append(np.zeros(len(tasks), dtype=stats_dt))
    missed = np.zeros(n, dtype=np.uint64)

    def ptrs(arrs):
        return (C.c_void_p * n)(*[a.ctypes.data for a in arrs])

    def sizes(arrs):
        return (C.c_size_t * n)(*[len(a) for a in arrs])

    L = lib()
    L.oracle_simulate_genome.argtypes = [
        C.POINTER(abi.SimParams), C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    rc = L.oracle_simulate_genome(
        C.byref(params), n, ivs, ptrs([j[1] for j in jobs]), sizes([j[1] for j in jobs]),
        ptrs([j[2] for j in jobs]), sizes([j[2] for j in jobs]), ptrs(bands), ptrs(occs),
        ptrs(stats), missed.ctypes.data, int(nthreads))
    assert rc == 0
    return [(bands[i], occs[i], stats[i], int(missed[i])) for i in range(n)]


def burnin_margin(reset=True):
    """Counters of the oracle's burn-in comparisons since the last reset: (comparisons of two
    window means of the coefficient of variation, those within 64 ulp of a tie, the smallest
    relative gap). See oracle_sim.hpp burnin_margin_note."""
    out = (C.c_uint64 * 3)()
    L = lib()
    L.oracle_burnin_margin.argtypes = [C.POINTER(C.c_uint64), C.c_int]
    L.oracle_burnin_margin.restype = None
    L.oracle_burnin_margin(out, 1 if reset else 0)
    gap = np.array([out[2]], dtype=np.uint64).view(np.float64)[0]
    return int(out[0]), int(out[1]), float(gap)


def band_to_pixels(band, nrows, ncols, bin_offset=0):
    """CPU counterpart of modle_b200_band_to_pixels (the reference's .cool pixel loop)."""
    band = np.ascontiguousarray(band, dtype=np.uint32)
    assert band.size >= nrows * ncols + 1
    L = lib()
    L.oracle_band_to_pixels.restype = C.c_uint64
    L.oracle_band_to_pixels.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64,
                                        C.c_void_p, C.c_uint64]
    n = int(L.oracle_band_to_pixels(band.ctypes.data, nrows, ncols, bin_offset, None, 0))
    out = np.zeros(n, dtype=abi.pixel_dtype())
    if n:
        L.oracle_band_to_pixels(band.ctypes.data, nrows, ncols, bin_offset, out.ctypes.data, n)
    return out


def lef_occupancy_profile(occ1d):
    """write_lef_occupancy_to_bwig (src/libmodle/cpu/simulation.cpp:170-197), numpy restatement:
    float(double(n) / double(max))."""
    occ1d = np.asarray(occ1d, dtype=np.uint64)
    with np.errstate(invalid="ignore", divide="ignore"):
        return (occ1d.astype(np.float64) / np.float64(occ1d.max())).astype(np.float32)


def snapshot_cell(params, interval, barriers, task):
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
    rc = lib().oracle_snapshot_cell(C.byref(params), C.byref(interval), barriers.ctypes.data, nb,
                                    task.ctypes.data, C.byref(snap), C.byref(st))
    assert rc == 0
    arrs["barrier_active"] = arrs["barrier_active"][:nb]
    arrs["num_active_lefs"] = int(snap.num_active_lefs)
    arrs["burnin_completed"] = int(snap.burnin_completed)
    arrs["stats"] = {f: int(getattr(st, f)) for f, _ in abi.CellStats._fields_}
    return arrs
