"""ctypes mirror of include/modle_b200.h (POD structs only; no compute here)."""
import ctypes as C

U64_MAX = (1 << 64) - 1
DIR_REV, DIR_FWD = 1, 2
SAMPLE_NOISIFY, SAMPLE_TAD, SAMPLE_LOOP = 1, 2, 4
STOP_CONTACT_DENSITY, STOP_SIMULATION_EPOCHS = 0, 1

OK = 0
ERR_INVALID_ARGUMENT = -1
ERR_NO_DEVICE = -2
ERR_CUDA = -3
ERR_UNSUPPORTED = -4
ERR_DEVICE_FAULT = -5


class SimParams(C.Structure):
    _fields_ = [
        ("bin_size", C.c_uint64),
        ("diagonal_width", C.c_uint64),
        ("rev_extrusion_speed", C.c_uint64),
        ("fwd_extrusion_speed", C.c_uint64),
        ("rev_extrusion_speed_burnin", C.c_uint64),
        ("fwd_extrusion_speed_burnin", C.c_uint64),
        ("rev_extrusion_speed_std", C.c_double),
        ("fwd_extrusion_speed_std", C.c_double),
        ("prob_of_lef_release", C.c_double),
        ("prob_of_lef_release_burnin", C.c_double),
        ("hard_stall_lef_stability_multiplier", C.c_double),
        ("soft_stall_lef_stability_multiplier", C.c_double),
        ("probability_of_extrusion_unit_bypass", C.c_double),
        ("lef_bar_major_collision_pblock", C.c_double),
        ("lef_bar_minor_collision_pblock", C.c_double),
        ("tad_to_loop_contact_ratio", C.c_double),
        ("genextreme_mu", C.c_double),
        ("genextreme_sigma", C.c_double),
        ("genextreme_xi", C.c_double),
        ("number_of_lefs_per_mbp", C.c_double),
        ("target_contact_density", C.c_double),
        ("target_simulation_epochs", C.c_uint64),
        ("contact_sampling_interval", C.c_uint64),
        ("avg_lef_processivity", C.c_uint64),
        ("probability_normalization_factor", C.c_uint64),
        ("extrusion_barrier_occupancy", C.c_double),
        ("barrier_occupied_stp", C.c_double),
        ("barrier_not_occupied_stp", C.c_double),
        ("burnin_speed_coefficient", C.c_double),
        ("burnin_history_length", C.c_uint64),
        ("burnin_smoothing_window_size", C.c_uint64),
        ("min_burnin_epochs", C.c_uint64),
        ("max_burnin_epochs", C.c_uint64),
        ("burnin_target_epochs_for_lef_activation", C.c_uint64),
        ("num_cells", C.c_uint64),
        ("seed", C.c_uint64),
        ("contact_sampling_strategy", C.c_uint32),
        ("stopping_criterion", C.c_uint32),
        ("track_1d_lef_position", C.c_uint32),
        ("skip_burnin", C.c_uint32),
        ("normalize_probabilities", C.c_uint32),
        ("override_extrusion_barrier_occupancy", C.c_uint32),
        ("debug_max_epochs", C.c_uint64),
    ]

    def copy(self):
        other = SimParams()
        C.memmove(C.byref(other), C.byref(self), C.sizeof(SimParams))
        return other


class Interval(C.Structure):
    _fields_ = [
        ("chrom_size", C.c_uint64),
        ("start", C.c_uint64),
        ("end", C.c_uint64),
        ("num_lefs", C.c_uint64),
    ]


class Barrier(C.Structure):
    _fields_ = [
        ("pos", C.c_uint64),
        ("stp_active", C.c_double),
        ("stp_inactive", C.c_double),
        ("blocking_direction", C.c_uint32),
        ("reserved_", C.c_uint32),
    ]


class CellTask(C.Structure):
    _fields_ = [
        ("cell_id", C.c_uint64),
        ("num_target_epochs", C.c_uint64),
        ("num_target_contacts", C.c_uint64),
        ("rng_state", C.c_uint64 * 4),
    ]


class CellStats(C.Structure):
    _fields_ = [
        ("num_contacts", C.c_uint64),
        ("num_epochs", C.c_uint64),
        ("num_burnin_epochs", C.c_uint64),
        ("num_lef_updates", C.c_uint64),
        ("num_rng_draws", C.c_uint64),
        ("device_fault", C.c_uint64),
    ]


class CellSnapshot(C.Structure):
    _fields_ = [
        ("rev_pos", C.POINTER(C.c_uint64)),
        ("fwd_pos", C.POINTER(C.c_uint64)),
        ("binding_epoch", C.POINTER(C.c_uint64)),
        ("rev_ranks", C.POINTER(C.c_uint64)),
        ("fwd_ranks", C.POINTER(C.c_uint64)),
        ("barrier_active", C.POINTER(C.c_uint8)),
        ("num_active_lefs", C.c_uint64),
        ("burnin_completed", C.c_uint64),
    ]


class Pixel(C.Structure):
    """hictk::ThinPixel<std::int32_t> layout."""
    _fields_ = [
        ("bin1_id", C.c_uint64),
        ("bin2_id", C.c_uint64),
        ("count", C.c_int32),
        ("reserved_", C.c_int32),
    ]


def epoch_record_dtype():
    """modle_b200_epoch_record (the quantities Simulation::dump_stats logs per epoch)."""
    import numpy as np

    dt = np.dtype([("epoch", "<u8"), ("loop_size_sum", "<u8"), ("burnin", "<u4"),
                   ("num_lefs", "<u4"), ("barriers_occupied", "<u4"), ("lefs_stalled_rev", "<u4"),
                   ("lefs_stalled_fwd", "<u4"), ("lefs_stalled_both", "<u4"),
                   ("lef_bar_collisions", "<u4"), ("lef_lef_primary_collisions", "<u4"),
                   ("lef_lef_secondary_collisions", "<u4"), ("reserved_", "<u4")])
    assert dt.itemsize == 56
    return dt


class ShardRecord(C.Structure):
    """modle_b200_shard."""
    _fields_ = [
        ("interval", C.c_uint64),
        ("cell_lo", C.c_uint64),
        ("cell_hi", C.c_uint64),
        ("rank", C.c_int32),
        ("reserved_", C.c_int32),
        ("weight", C.c_double),
    ]


def shard_dtype():
    import numpy as np

    dt = np.dtype([("interval", "<u8"), ("cell_lo", "<u8"), ("cell_hi", "<u8"), ("rank", "<i4"),
                   ("reserved_", "<i4"), ("weight", "<f8")])
    assert dt.itemsize == C.sizeof(ShardRecord) == 40
    return dt


def pixel_dtype():
    import numpy as np

    dt = np.dtype([("bin1_id", "<u8"), ("bin2_id", "<u8"), ("count", "<i4"), ("reserved_", "<i4")])
    assert dt.itemsize == C.sizeof(Pixel) == 24
    return dt


# numpy structured dtypes with the same layout (for zero-copy views of arrays of structs)
def np_dtypes():
    import numpy as np

    barrier = np.dtype(
        [("pos", "<u8"), ("stp_active", "<f8"), ("stp_inactive", "<f8"),
         ("blocking_direction", "<u4"), ("reserved_", "<u4")])
    task = np.dtype(
        [("cell_id", "<u8"), ("num_target_epochs", "<u8"), ("num_target_contacts", "<u8"),
         ("rng_state", "<u8", (4,))])
    stats = np.dtype(
        [("num_contacts", "<u8"), ("num_epochs", "<u8"), ("num_burnin_epochs", "<u8"),
         ("num_lef_updates", "<u8"), ("num_rng_draws", "<u8"), ("device_fault", "<u8")])
    assert barrier.itemsize == C.sizeof(Barrier)
    assert task.itemsize == C.sizeof(CellTask)
    assert stats.itemsize == C.sizeof(CellStats)
    return barrier, task, stats
