"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of the CPU emulation of the kernel (tests/emu)."""
import ctypes as C
import os

import numpy as np

from modle_b200 import abi, host

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
_LIB = None


CXXFLAGS = ["-O2", "-march=x86-64-v3", "-ffp-contract=off", "-fno-fast-math", "-std=c++17", "-fPIC",
            "-Wall", "-Wextra", "-Wno-unused-parameter", "-pthread", "-shared"]  # as tests/emu/Makefile


def build(force=False):
    """Builds libemu.so when the kernel sources changed (content hash, modle_b200/buildutil.py)."""
    from modle_b200 import buildutil

    csrc = os.path.join(_HERE, "..", "..", "modle_b200", "csrc")
    deps = [os.path.join(_HERE, "emu_capi.cpp")] + [
        os.path.join(csrc, f) for f in ("sim_core.hpp", "sim_types.hpp", "cta.hpp", "launch_prep.hpp",
                                        "host_rng.hpp", "host.cpp", "status.hpp", "ziggurat_tables.inc")]
    deps.append(os.path.join(_HERE, "..", "..", "include", "modle_b200.h"))
    cxx = os.environ.get("CXX", "g++")
    # MODLE_B200_EMU_DEFINES="A=1,B=2": an experimental copy of the kernel source (the switches of
    # sim_core.hpp, e.g. MODLE_B200_WINDOW_RANK_REPAIR=1) in its own library next to the default one
    defines = [d for d in os.environ.get("MODLE_B200_EMU_DEFINES", "").split(",") if d]
    name = "libemu.so" if not defines else "libemu_variant.so"
    flags = CXXFLAGS + ["-D" + d for d in defines]
    return buildutil.ensure_built(
        os.path.join(_HERE, name), deps,
        lambda tmp: [cxx] + flags + ["-o", tmp, os.path.join(_HERE, "emu_capi.cpp"),
                                     os.path.join(csrc, "host.cpp")],
        extra=" ".join(flags), force=force)


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        u64p = C.POINTER(C.c_uint64)
        L.emu_last_error.restype = C.c_char_p
        L.emu_simulate_interval.argtypes = [
            C.POINTER(abi.SimParams), C.POINTER(abi.Interval), C.c_void_p, C.c_size_t, C.c_void_p,
            C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, u64p, C.c_int, C.c_int]
        L.emu_snapshot_cell.argtypes = [
            C.POINTER(abi.SimParams), C.POINTER(abi.Interval), C.c_void_p, C.c_size_t, C.c_void_p,
            C.POINTER(abi.CellSnapshot), C.POINTER(abi.CellStats), C.c_int, C.c_int]
        L.emu_collision_steps.argtypes = (
            [C.c_uint32, C.c_uint64, C.c_uint64, C.c_size_t] + [C.c_void_p] * 9 +
            [C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double,
             C.c_uint64, C.c_void_p, C.c_int])
        L.emu_rank_lefs.argtypes = [C.c_void_p] * 5 + [C.c_size_t, C.c_int]
        _LIB = L
    return _LIB


def set_thread_order(mode):
    """0: virtual threads of a region run in ascending order, 1: descending, 2: shuffled
    differently for every region (cta.hpp emu_thread_order)."""
    lib().emu_set_thread_order(int(mode))


def set_force_bind_redo(on):
    """bind_lefs takes its sequential redo (the uniform_int rejection path) in every epoch of the
    calling thread's later calls, although nothing was rejected."""
    lib().emu_set_force_bind_redo(1 if on else 0)


def set_rng_mode(mode):
    """0: deterministic mode (reference draw order), 1: throughput mode (counter-based draws);
    see modle_b200_set_rng_mode. Applies to the calling thread's later calls."""
    lib().emu_set_rng_mode(int(mode))


def barrier_count(reset=True):
    """CTA barriers the device build would have executed in this thread's emulation calls."""
    L = lib()
    L.emu_barrier_count_get.restype = C.c_uint64
    return int(L.emu_barrier_count_get(1 if reset else 0))


def phase_barriers(reset=True):
    """{phase name: CTA barriers} of this thread's emulation calls (the emulation's counterpart
    of modle_b200_phase_cycles)."""
    out = (C.c_uint64 * len(host.PHASE_NAMES))()
    lib().emu_phase_barriers_get(out, len(host.PHASE_NAMES), 1 if reset else 0)
    return dict(zip(host.PHASE_NAMES, [int(x) for x in out]))


def _check(rc):
    if rc != 0:
        raise RuntimeError("emu: " + lib().emu_last_error().decode())


def simulate_interval(params, interval, barriers, tasks, virtual_threads=64, staging=0,
                      log_capacity_per_cell=0):
    _, _, stats_dt = abi.np_dtypes()
    nrows, ncols = host.band_shape(params, int(interval.end - interval.start))
    band = np.zeros(nrows * ncols + 1, dtype=np.uint32)
    occ = np.zeros(ncols, dtype=np.uint64)
    stats = np.zeros(len(tasks), dtype=stats_dt)
    missed = C.c_uint64(0)
    barriers = np.ascontiguousarray(barriers)
    tasks = np.ascontiguousarray(tasks)
    if log_capacity_per_cell:
        log = np.zeros((len(tasks), int(log_capacity_per_cell)), dtype=abi.epoch_record_dtype())
        L = lib()
        L.emu_simulate_interval_logged.argtypes = [
            C.POINTER(abi.SimParams), C.POINTER(abi.Interval), C.c_void_p, C.c_size_t, C.c_void_p,
            C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.c_int,
            C.c_int, C.c_void_p, C.c_size_t]
        _check(L.emu_simulate_interval_logged(
            C.byref(params), C.byref(interval), barriers.ctypes.data if len(barriers) else None,
            len(barriers), tasks.ctypes.data, len(tasks), band.ctypes.data, occ.ctypes.data,
            stats.ctypes.data, C.byref(missed), virtual_threads, staging, log.ctypes.data,
            int(log_capacity_per_cell)))
        return band, occ, stats, int(missed.value), log
    _check(lib().emu_simulate_interval(
        C.byref(params), C.byref(interval), barriers.ctypes.data if len(barriers) else None,
        len(barriers), tasks.ctypes.data, len(tasks), band.ctypes.data, occ.ctypes.data,
        stats.ctypes.data, C.byref(missed), virtual_threads, staging))
    return band, occ, stats, int(missed.value)


def snapshot_cell(params, interval, barriers, task, virtual_threads=64, staging=0):
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
    _check(lib().emu_snapshot_cell(C.byref(params), C.byref(interval),
                                   barriers.ctypes.data if nb else None, nb, task.ctypes.data,
                                   C.byref(snap), C.byref(st), virtual_threads, staging))
    arrs["barrier_active"] = arrs["barrier_active"][:nb]
    arrs["num_active_lefs"] = int(snap.num_active_lefs)
    arrs["burnin_completed"] = int(snap.burnin_completed)
    arrs["stats"] = {f: int(getattr(st, f)) for f, _ in abi.CellStats._fields_}
    return arrs


def collision_steps(steps, start, end, rev, fwd, ep, rr, fr, rm, fm, bar_pos, bar_dir, bar_active,
                    prob_bypass=0.0, pblock_major=1.0, pblock_minor=0.0, rng_seed=0,
                    rc=None, fc=None, virtual_threads=7):
    from oracle import pyoracle  # only for the STEP name table

    if not isinstance(steps, int):
        m = 0
        for s in steps:
            m |= pyoracle.STEP[s]
        steps = m
    n = len(rev)
    a = {k: np.array(v, dtype=np.uint64) for k, v in
         dict(rev=rev, fwd=fwd, ep=ep, rr=rr, fr=fr, rm=rm, fm=fm).items()}
    a["rc"] = np.zeros(n, dtype=np.uint32) if rc is None else np.array(rc, dtype=np.uint32)
    a["fc"] = np.zeros(n, dtype=np.uint32) if fc is None else np.array(fc, dtype=np.uint32)
    bp = np.array(bar_pos, dtype=np.uint64)
    bd = np.array(bar_dir, dtype=np.uint8)
    ba = np.array(bar_active, dtype=np.uint8)
    info = np.zeros(4, dtype=np.uint64)
    _check(lib().emu_collision_steps(
        steps, start, end, n, a["rev"].ctypes.data, a["fwd"].ctypes.data, a["ep"].ctypes.data,
        a["rr"].ctypes.data, a["fr"].ctypes.data, a["rm"].ctypes.data, a["fm"].ctypes.data,
        a["rc"].ctypes.data, a["fc"].ctypes.data, len(bp), bp.ctypes.data, bd.ctypes.data,
        ba.ctypes.data, prob_bypass, pblock_major, pblock_minor, rng_seed, info.ctypes.data,
        virtual_threads))
    a["n5"], a["n3"], a["ndraws"], a["fault"] = (int(x) for x in info)
    return a


def rank_lefs(rev, fwd, ep, rr, fr, virtual_threads=7):
    rev, fwd, ep = (np.ascontiguousarray(a, dtype=np.uint64) for a in (rev, fwd, ep))
    rr = np.array(rr, dtype=np.uint64)
    fr = np.array(fr, dtype=np.uint64)
    _check(lib().emu_rank_lefs(rev.ctypes.data, fwd.ctypes.data, ep.ctypes.data, rr.ctypes.data,
                               fr.ctypes.data, len(rev), virtual_threads))
    return rr, fr


def sample_moves(state, n, speed, sd, virtual_threads=64, staging=0):
    """n moves round(max(0, Normal(speed, sd))) through the kernel's draw_normal_moves; returns
    (moves, raw draws consumed)."""
    L = lib()
    L.emu_sample_moves.restype = C.c_longlong
    L.emu_sample_moves.argtypes = [C.POINTER(C.c_uint64), C.c_size_t, C.c_double, C.c_double,
                                   C.c_void_p, C.c_int, C.c_int]
    st = (C.c_uint64 * 4)(*[int(x) for x in state])
    out = np.zeros(n, dtype=np.uint64)
    rc = L.emu_sample_moves(st, n, speed, sd, out.ctypes.data, virtual_threads, staging)
    if rc < 0:
        raise RuntimeError(f"emu_sample_moves failed: {rc} {L.emu_last_error().decode()}")
    return out, int(rc)
