"""ctypes bindings of the host-only half of the C ABI (no GPU needed)."""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# MODLE_B200_LIB: load an experimental build of the same library instead (profiling sessions)
_LIB_PATH = os.environ.get("MODLE_B200_LIB") or os.path.join(_HERE, "libmodle_b200.so")
_LIB = None


class ModleB200Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"modle_b200 error {code}: {message}")
        self.code = code


def lib():
    """Loads libmodle_b200.so; raises loudly when the CUDA extension has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(_LIB_PATH):
            raise ImportError(
                f"{_LIB_PATH} is missing: build it with `python -m modle_b200.build` "
                "(modle_b200 has no CPU fallback)")
        L = C.CDLL(_LIB_PATH)
        u64p = C.POINTER(C.c_uint64)
        P = C.POINTER(abi.SimParams)
        L.modle_b200_abi_version.restype = C.c_int
        L.modle_b200_last_error.restype = C.c_char_p
        L.modle_b200_default_params.argtypes = [P]
        L.modle_b200_default_params.restype = None
        L.modle_b200_transform_params.argtypes = [P, C.c_int, C.c_int, C.c_int]
        L.modle_b200_compute_num_lefs.argtypes = [P, C.c_uint64]
        L.modle_b200_compute_num_lefs.restype = C.c_uint64
        L.modle_b200_compute_contacts_per_epoch.argtypes = [P, C.c_uint64]
        L.modle_b200_compute_contacts_per_epoch.restype = C.c_uint64
        L.modle_b200_band_shape.argtypes = [P, C.c_uint64, u64p, u64p]
        L.modle_b200_band_shape.restype = None
        L.modle_b200_interval_hash.argtypes = [C.c_char_p, C.c_size_t, C.c_uint64, C.c_uint64,
                                               C.c_uint64, C.c_uint64, u64p]
        L.modle_b200_rng_seed.argtypes = [C.c_uint64, u64p]
        L.modle_b200_rng_seed.restype = None
        L.modle_b200_rng_next.argtypes = [u64p]
        L.modle_b200_rng_next.restype = C.c_uint64
        L.modle_b200_rng_jump.argtypes = [u64p]
        L.modle_b200_rng_jump.restype = None
        L.modle_b200_stp_active_from_occupancy.argtypes = [C.c_double, C.c_double]
        L.modle_b200_stp_active_from_occupancy.restype = C.c_double
        L.modle_b200_occupancy_from_stp.argtypes = [C.c_double, C.c_double]
        L.modle_b200_occupancy_from_stp.restype = C.c_double
        L.modle_b200_make_cell_tasks.argtypes = [P, C.c_char_p, C.c_size_t,
                                                 C.POINTER(abi.Interval), C.c_void_p]
        L.modle_b200_init.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        L.modle_b200_destroy.argtypes = [C.c_void_p]
        L.modle_b200_destroy.restype = None
        L.modle_b200_launch_geometry.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_uint32),
                                                 C.POINTER(C.c_uint32), u64p]
        L.modle_b200_set_rng_mode.argtypes = [C.c_void_p, C.c_int]
        L.modle_b200_get_rng_mode.argtypes = [C.c_void_p]
        L.modle_b200_synchronize.argtypes = [C.c_void_p]
        L.modle_b200_phase_cycles.argtypes = [C.c_void_p, u64p, C.c_size_t, C.c_int]
        L.modle_b200_kernel_launches.argtypes = [C.c_void_p]
        L.modle_b200_kernel_launches.restype = C.c_uint64
        L.modle_b200_reserve.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_size_t]
        L.modle_b200_simulate_interval.argtypes = [
            C.c_void_p, P, C.POINTER(abi.Interval), C.c_void_p, C.c_size_t, C.c_void_p,
            C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, u64p]
        L.modle_b200_simulate_interval_overwrite.argtypes = L.modle_b200_simulate_interval.argtypes
        L.modle_b200_simulate_interval_logged.argtypes = [
            C.c_void_p, P, C.POINTER(abi.Interval), C.c_void_p, C.c_size_t, C.c_void_p,
            C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, u64p, C.c_void_p, C.c_size_t]
        L.modle_b200_simulate_interval_device.argtypes = [
            C.c_void_p, P, C.POINTER(abi.Interval), C.c_void_p, C.c_size_t, C.c_void_p,
            C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.modle_b200_snapshot_cell.argtypes = [
            C.c_void_p, P, C.POINTER(abi.Interval), C.c_void_p, C.c_size_t, C.c_void_p,
            C.POINTER(abi.CellSnapshot), C.POINTER(abi.CellStats)]
        L.modle_b200_register_contacts_device.argtypes = [
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_void_p,
            C.c_void_p, C.c_void_p]
        L.modle_b200_calibrate_red_device.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64,
                                                      C.c_uint64, C.c_uint64, C.c_void_p]
        L.modle_b200_count_pixels_device.argtypes = [
            C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        L.modle_b200_fill_pixels_device.argtypes = [
            C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p,
            C.c_uint64, C.c_void_p]
        L.modle_b200_band_to_pixels.argtypes = [
            C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64,
            u64p]
        L.modle_b200_lef_occupancy_profile_device.argtypes = [
            C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        L.modle_b200_lef_occupancy_profile.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t,
                                                       C.c_void_p]
        L.modle_b200_genome_import.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, P, C.c_int,
                                               C.POINTER(C.c_void_p)]
        L.modle_b200_genome_free.argtypes = [C.c_void_p]
        L.modle_b200_genome_free.restype = None
        L.modle_b200_genome_num_chromosomes.argtypes = [C.c_void_p]
        L.modle_b200_genome_num_chromosomes.restype = C.c_size_t
        L.modle_b200_genome_num_intervals.argtypes = [C.c_void_p]
        L.modle_b200_genome_num_intervals.restype = C.c_size_t
        L.modle_b200_genome_num_barriers.argtypes = [C.c_void_p]
        L.modle_b200_genome_num_barriers.restype = C.c_uint64
        L.modle_b200_genome_get_interval.argtypes = [
            C.c_void_p, C.c_size_t, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), u64p, u64p, u64p,
            C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), u64p]
        _LIB = L
    return _LIB


EXPORTED_SYMBOLS = [
    "modle_b200_abi_version", "modle_b200_last_error", "modle_b200_default_params",
    "modle_b200_transform_params", "modle_b200_compute_num_lefs",
    "modle_b200_compute_contacts_per_epoch", "modle_b200_band_shape", "modle_b200_interval_hash",
    "modle_b200_rng_seed", "modle_b200_rng_next", "modle_b200_rng_jump",
    "modle_b200_stp_active_from_occupancy", "modle_b200_occupancy_from_stp",
    "modle_b200_make_cell_tasks", "modle_b200_init", "modle_b200_destroy",
    "modle_b200_set_rng_mode", "modle_b200_get_rng_mode", "modle_b200_launch_geometry",
    "modle_b200_reserve", "modle_b200_simulate_interval", "modle_b200_simulate_interval_overwrite",
    "modle_b200_simulate_interval_logged",
    "modle_b200_simulate_interval_device",
    "modle_b200_synchronize", "modle_b200_snapshot_cell", "modle_b200_register_contacts_device",
    "modle_b200_kernel_launches", "modle_b200_phase_cycles", "modle_b200_calibrate_red_device",
    "modle_b200_cell_weight", "modle_b200_plan_shards", "modle_b200_reduce_band",
    "modle_b200_count_pixels_device", "modle_b200_fill_pixels_device", "modle_b200_band_to_pixels",
    "modle_b200_lef_occupancy_profile_device", "modle_b200_lef_occupancy_profile",
    "modle_b200_genome_import", "modle_b200_genome_free", "modle_b200_genome_num_chromosomes",
    "modle_b200_genome_num_intervals", "modle_b200_genome_num_barriers",
    "modle_b200_genome_get_interval",
]

PHASE_NAMES = ["init", "burnin", "bind", "rank", "contacts", "moves_generate", "moves_adjust",
               "barrier_states", "lef_bar", "primary", "correct_moves", "secondary", "fix_ranks",
               "extrude_release", "rng_refill(nested)", "total",
               "mv.ensure", "mv.scan", "mv.exceptions", "mv.final",
               "sec.compose", "sec.scan", "sec.classify", "sec.draws", "sec.leader", "sec.apply"]


def check(rc):
    if rc != 0:
        raise ModleB200Error(rc, lib().modle_b200_last_error().decode(errors="replace"))


def default_params():
    p = abi.SimParams()
    lib().modle_b200_default_params(C.byref(p))
    return p


def transform_params(p, rev_speed_given=False, fwd_speed_given=False,
                     barrier_occupancy_given=False):
    check(lib().modle_b200_transform_params(C.byref(p), int(rev_speed_given), int(fwd_speed_given),
                                            int(barrier_occupancy_given)))
    return p


def launch_geometry(num_lefs, num_barriers):
    """(threads per CTA, cells resident per SM, shared-memory bytes per cell) for an interval."""
    t, c, b = C.c_uint32(0), C.c_uint32(0), C.c_uint64(0)
    check(lib().modle_b200_launch_geometry(int(num_lefs), int(num_barriers), C.byref(t),
                                           C.byref(c), C.byref(b)))
    return int(t.value), int(c.value), int(b.value)


def compute_num_lefs(p, size_bp):
    return int(lib().modle_b200_compute_num_lefs(C.byref(p), size_bp))


def compute_contacts_per_epoch(p, num_lefs):
    return int(lib().modle_b200_compute_contacts_per_epoch(C.byref(p), num_lefs))


def band_shape(p, size_bp):
    nrows, ncols = C.c_uint64(), C.c_uint64()
    lib().modle_b200_band_shape(C.byref(p), size_bp, C.byref(nrows), C.byref(ncols))
    return int(nrows.value), int(ncols.value)


def interval_hash(name, chrom_size, start, end, seed):
    out = C.c_uint64()
    b = name.encode()
    check(lib().modle_b200_interval_hash(b, len(b), chrom_size, start, end, seed, C.byref(out)))
    return int(out.value)


def rng_seed(seed):
    st = (C.c_uint64 * 4)()
    lib().modle_b200_rng_seed(seed, st)
    return [int(x) for x in st]


def rng_next(state):
    st = (C.c_uint64 * 4)(*state)
    v = lib().modle_b200_rng_next(st)
    return int(v), [int(x) for x in st]


def rng_jump(state):
    st = (C.c_uint64 * 4)(*state)
    lib().modle_b200_rng_jump(st)
    return [int(x) for x in st]


def make_cell_tasks(p, name, interval):
    """Per-interval fan-out of Simulation::run_simulate (scheduler_simulate.cpp:104-160)."""
    _, task_dt, _ = abi.np_dtypes()
    tasks = np.zeros(int(p.num_cells), dtype=task_dt)
    b = name.encode()
    check(lib().modle_b200_make_cell_tasks(C.byref(p), b, len(b), C.byref(interval),
                                           tasks.ctypes.data))
    return tasks


def barriers_from_records(records, p):
    """BED-like records (pos, strand '+'/'-', score) -> barrier array, sorted by position.

    Follows generate_barriers_from_bed_records / compute_barrier_stp
    (src/libmodle/internal/genome.cpp:255-271,423-469) and, when
    override_extrusion_barrier_occupancy is set, simulation.cpp:51-60.
    """
    barrier_dt, _, _ = abi.np_dtypes()
    recs = sorted(records, key=lambda r: r[0])
    out = np.zeros(len(recs), dtype=barrier_dt)
    L = lib()
    for i, (pos, strand, score) in enumerate(recs):
        if p.override_extrusion_barrier_occupancy:
            stp_a, stp_i = p.barrier_occupied_stp, p.barrier_not_occupied_stp
        elif score != 0.0:
            stp_i = p.barrier_not_occupied_stp
            stp_a = L.modle_b200_stp_active_from_occupancy(stp_i, score)
        else:
            stp_a, stp_i = p.barrier_occupied_stp, p.barrier_not_occupied_stp
        out[i] = (pos, stp_a, stp_i, abi.DIR_REV if strand == "+" else abi.DIR_FWD, 0)
    return out


def import_genome(path_to_chrom_sizes, path_to_extr_barriers, p, path_to_genomic_intervals="",
                  interpret_name_field_as_puu=False):
    """Genome::Genome (src/libmodle/internal/genome.cpp:299-330) through the C ABI. Returns a list
    of dicts: chrom_name, chrom_id, chrom_size, start, end, barriers (abi barrier dtype, sorted by
    position), bin_offset. Raises ModleB200Error with the parser's diagnostic on malformed input."""
    barrier_dt, _, _ = abi.np_dtypes()
    L = lib()
    g = C.c_void_p()
    check(L.modle_b200_genome_import(
        os.fsencode(path_to_chrom_sizes), os.fsencode(path_to_extr_barriers),
        os.fsencode(path_to_genomic_intervals) if path_to_genomic_intervals else None,
        C.byref(p), int(interpret_name_field_as_puu), C.byref(g)))
    try:
        out = []
        for i in range(L.modle_b200_genome_num_intervals(g)):
            name, cid, nb = C.c_char_p(), C.c_size_t(), C.c_size_t()
            size, start, end, off = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
            bars = C.c_void_p()
            check(L.modle_b200_genome_get_interval(
                g, i, C.byref(name), C.byref(cid), C.byref(size), C.byref(start), C.byref(end),
                C.byref(bars), C.byref(nb), C.byref(off)))
            arr = np.zeros(nb.value, dtype=barrier_dt)
            if nb.value:
                C.memmove(arr.ctypes.data, bars.value, nb.value * barrier_dt.itemsize)
            out.append(dict(chrom_name=name.value.decode(), chrom_id=int(cid.value),
                            chrom_size=int(size.value), start=int(start.value), end=int(end.value),
                            barriers=arr, bin_offset=int(off.value)))
        assert sum(len(o["barriers"]) for o in out) == L.modle_b200_genome_num_barriers(g)
        return out
    finally:
        L.modle_b200_genome_free(g)
