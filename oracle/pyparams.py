"""TEST INFRASTRUCTURE ONLY -- the host-side arithmetic of the path restated in pure Python, so that
the CPU arm (bench.py --impl reference / cpu_baseline) and the tests can build their inputs
without loading the product library. Independent of modle_b200/csrc/host.cpp (the tests compare
the two). Only the ctypes struct definitions of modle_b200.abi are shared (layout, no code).

Reference (paths relative to the reference checkout):
  Config{} member initialisers   src/common/include/modle/common/simulation_config.hpp:53-113
  Cli::transform_args            src/modle/cli.cpp:843-863, 886-1016
  compute_num_lefs               src/libmodle/cpu/simulation.cpp:1086-1090
  ContactMatrixDense geometry    src/contact_matrix/contact_matrix_dense_impl.hpp:39-50
  barrier stp from BED scores    src/libmodle/internal/genome.cpp:255-271, simulation.cpp:51-60
"""
import math

import numpy as np

from modle_b200 import abi

U64_MAX = (1 << 64) - 1


def _round_half_away(x):  # std::round
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


def default_params():
    p = abi.SimParams()
    p.bin_size = 5000
    p.diagonal_width = 3_000_000
    p.fwd_extrusion_speed = p.rev_extrusion_speed = p.bin_size * 8 // 10
    p.fwd_extrusion_speed_std = p.rev_extrusion_speed_std = 0.05
    p.number_of_lefs_per_mbp = 20
    p.avg_lef_processivity = 300_000
    p.contact_sampling_interval = 50_000
    p.track_1d_lef_position = 1
    p.extrusion_barrier_occupancy = 0.825
    p.barrier_occupied_stp = 0.0
    p.barrier_not_occupied_stp = 0.70
    p.hard_stall_lef_stability_multiplier = 5.0
    p.soft_stall_lef_stability_multiplier = 1.0
    p.probability_of_extrusion_unit_bypass = 0.1
    p.lef_bar_major_collision_pblock = 1.0
    p.lef_bar_minor_collision_pblock = 0.0
    p.tad_to_loop_contact_ratio = 5.0
    p.genextreme_mu, p.genextreme_sigma, p.genextreme_xi = 0.0, 5000.0, 0.001
    p.target_simulation_epochs = 2000
    p.target_contact_density = 1.0
    p.stopping_criterion = abi.STOP_CONTACT_DENSITY
    p.contact_sampling_strategy = abi.SAMPLE_TAD | abi.SAMPLE_LOOP | abi.SAMPLE_NOISIFY
    p.num_cells = 512
    p.seed = 0
    p.probability_normalization_factor = p.rev_extrusion_speed + p.fwd_extrusion_speed
    p.normalize_probabilities = 1
    p.skip_burnin = 0
    p.burnin_history_length = 100
    p.burnin_smoothing_window_size = 5
    p.min_burnin_epochs = 0
    p.max_burnin_epochs = U64_MAX
    p.burnin_target_epochs_for_lef_activation = 320
    p.burnin_speed_coefficient = 1.0
    p.fwd_extrusion_speed_burnin = p.rev_extrusion_speed_burnin = p.fwd_extrusion_speed
    p.debug_max_epochs = U64_MAX
    return p


def stp_active_from_occupancy(stp_inactive, occupancy):  # cli.cpp:843-852
    if occupancy == 0:
        return 0.0
    to_active = 1.0 - stp_inactive
    to_inactive = (to_active - (occupancy * to_active)) / occupancy
    return min(max(1.0 - to_inactive, 0.0), 1.0)


def occupancy_from_stp(stp_active, stp_inactive):  # cli.cpp:854-863
    if stp_active + stp_inactive == 0:
        return 0.0
    to_active = 1.0 - stp_inactive
    to_inactive = 1.0 - stp_active
    return min(max(to_active / (to_active + to_inactive), 0.0), 1.0)


def transform_params(p, given=()):
    """Cli::transform_args (cli.cpp:993-1016); `given` = names of the options the user passed
    (only rev/fwd_extrusion_speed and extrusion_barrier_occupancy matter)."""
    if "rev_extrusion_speed" not in given:
        p.rev_extrusion_speed = p.bin_size * 8 // 10
    if "fwd_extrusion_speed" not in given:
        p.fwd_extrusion_speed = p.bin_size * 8 // 10
    if 0 < p.fwd_extrusion_speed_std < 1:
        p.fwd_extrusion_speed_std *= float(p.fwd_extrusion_speed)
    if 0 < p.rev_extrusion_speed_std < 1:
        p.rev_extrusion_speed_std *= float(p.rev_extrusion_speed)
    p.rev_extrusion_speed_burnin = _round_half_away(p.burnin_speed_coefficient * float(p.rev_extrusion_speed))
    p.fwd_extrusion_speed_burnin = _round_half_away(p.burnin_speed_coefficient * float(p.fwd_extrusion_speed))
    p.prob_of_lef_release = float(p.rev_extrusion_speed + p.fwd_extrusion_speed) / float(p.avg_lef_processivity)
    p.prob_of_lef_release_burnin = \
        float(p.rev_extrusion_speed_burnin + p.fwd_extrusion_speed_burnin) / float(p.avg_lef_processivity)
    occ_given = "extrusion_barrier_occupancy" in given
    if occ_given:
        p.barrier_occupied_stp = stp_active_from_occupancy(p.barrier_not_occupied_stp,
                                                           p.extrusion_barrier_occupancy)
    else:
        p.extrusion_barrier_occupancy = occupancy_from_stp(p.barrier_occupied_stp,
                                                           p.barrier_not_occupied_stp)
    loop = bool(p.contact_sampling_strategy & abi.SAMPLE_LOOP)
    tad = bool(p.contact_sampling_strategy & abi.SAMPLE_TAD)
    assert loop or tad
    if loop and not tad:
        p.tad_to_loop_contact_ratio = 0.0
    if tad and not loop:
        p.tad_to_loop_contact_ratio = math.inf
    p.burnin_target_epochs_for_lef_activation = min(
        p.max_burnin_epochs,
        5 * p.avg_lef_processivity // (p.rev_extrusion_speed_burnin + p.fwd_extrusion_speed_burnin))
    if p.normalize_probabilities:
        ratio = float(p.rev_extrusion_speed + p.fwd_extrusion_speed) / float(p.probability_normalization_factor)
        if ratio != 1.0:
            def stable_pow(base, e):
                if base == 0.0:
                    return 0.0
                if base == 1.0:
                    return 1.0
                return math.exp(math.log(base) * e)

            p.barrier_not_occupied_stp = stable_pow(p.barrier_not_occupied_stp, ratio)
            p.barrier_occupied_stp = stp_active_from_occupancy(p.barrier_not_occupied_stp,
                                                               p.extrusion_barrier_occupancy)
            bp = p.probability_of_extrusion_unit_bypass
            if bp != 0.0 and bp != 1.0:
                p.probability_of_extrusion_unit_bypass = min(bp * ratio, 1.0)
            p.lef_bar_major_collision_pblock = stable_pow(p.lef_bar_major_collision_pblock, ratio)
            p.lef_bar_minor_collision_pblock = stable_pow(p.lef_bar_minor_collision_pblock, ratio)
    if occ_given:
        p.override_extrusion_barrier_occupancy = 1
    if p.stopping_criterion == abi.STOP_SIMULATION_EPOCHS:
        p.target_contact_density = -1.0
    return p


def make_params(**overrides):
    """Config{} + the given options + transform_args."""
    p = default_params()
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return transform_params(p, set(overrides))


def compute_num_lefs(p, size_bp):  # simulation.cpp:1086-1090
    return max(1, _round_half_away(p.number_of_lefs_per_mbp * (float(size_bp) / 1.0e6)))


def band_shape(p, size_bp):  # contact_matrix_dense_impl.hpp:39-50
    ncols = (size_bp + p.bin_size - 1) // p.bin_size
    nrows = min((p.diagonal_width + p.bin_size - 1) // p.bin_size, ncols)
    return int(nrows), int(ncols)


def barriers_from_records(records, p):
    """(pos, strand, score) records -> barrier array sorted by position (genome.cpp:255-271,
    423-469; simulation.cpp:51-60 for the occupancy override)."""
    barrier_dt, _, _ = abi.np_dtypes()
    recs = sorted(records, key=lambda r: r[0])
    out = np.zeros(len(recs), dtype=barrier_dt)
    for i, (pos, strand, score) in enumerate(recs):
        if p.override_extrusion_barrier_occupancy or score == 0.0:
            stp_a, stp_i = p.barrier_occupied_stp, p.barrier_not_occupied_stp
        else:
            stp_i = p.barrier_not_occupied_stp
            stp_a = stp_active_from_occupancy(stp_i, score)
        out[i] = (pos, stp_a, stp_i, abi.DIR_REV if strand == "+" else abi.DIR_FWD, 0)
    return out
