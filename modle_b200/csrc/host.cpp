// Host half of the C ABI (include/modle_b200.h): parameter derivation, geometry, seeding and the
// per-interval task fan-out. No device code here. Each function names the reference lines whose
// behaviour it reproduces (paths relative to the reference checkout).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <string>

#include "../../include/modle_b200.h"
#include "host_rng.hpp"
#include "status.hpp"

using modle_b200::host::u64;
namespace mh = modle_b200::host;

namespace modle_b200 {
thread_local std::string g_last_error;
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
}  // namespace modle_b200

extern "C" {

int modle_b200_abi_version(void) { return MODLE_B200_ABI_VERSION; }
const char* modle_b200_last_error(void) { return modle_b200::g_last_error.c_str(); }

// Config{} member initialisers, src/common/include/modle/common/simulation_config.hpp:53-113
void modle_b200_default_params(modle_b200_sim_params* p) {
  std::memset(p, 0, sizeof(*p));
  p->bin_size = 5000;
  p->diagonal_width = 3000000;
  p->fwd_extrusion_speed = p->bin_size * 8 / 10;
  p->rev_extrusion_speed = p->fwd_extrusion_speed;
  p->fwd_extrusion_speed_std = 0.05;
  p->rev_extrusion_speed_std = 0.05;
  p->number_of_lefs_per_mbp = 20;
  p->avg_lef_processivity = 300000;
  p->contact_sampling_interval = 50000;
  p->track_1d_lef_position = 1;
  p->extrusion_barrier_occupancy = 0.825;
  p->barrier_occupied_stp = 0.0;
  p->barrier_not_occupied_stp = 0.70;
  p->hard_stall_lef_stability_multiplier = 5.0;
  p->soft_stall_lef_stability_multiplier = 1.0;
  p->probability_of_extrusion_unit_bypass = 0.1;
  p->lef_bar_major_collision_pblock = 1.0;
  p->lef_bar_minor_collision_pblock = 0.0;
  p->tad_to_loop_contact_ratio = 5.0;
  p->genextreme_mu = 0;
  p->genextreme_sigma = 5000;
  p->genextreme_xi = 0.001;
  p->target_simulation_epochs = 2000;
  p->target_contact_density = 1.0;
  p->stopping_criterion = MODLE_B200_STOP_CONTACT_DENSITY;
  p->contact_sampling_strategy =
      MODLE_B200_SAMPLE_TAD | MODLE_B200_SAMPLE_LOOP | MODLE_B200_SAMPLE_NOISIFY;
  p->num_cells = 512;
  p->seed = 0;
  p->probability_normalization_factor = p->rev_extrusion_speed + p->fwd_extrusion_speed;
  p->normalize_probabilities = 1;
  p->skip_burnin = 0;
  p->burnin_history_length = 100;
  p->burnin_smoothing_window_size = 5;
  p->min_burnin_epochs = 0;
  p->max_burnin_epochs = std::numeric_limits<u64>::max();
  p->burnin_target_epochs_for_lef_activation = 320;
  p->burnin_speed_coefficient = 1.0;
  p->fwd_extrusion_speed_burnin = p->fwd_extrusion_speed;
  p->rev_extrusion_speed_burnin = p->fwd_extrusion_speed_burnin;
  p->debug_max_epochs = std::numeric_limits<u64>::max();
}

double modle_b200_stp_active_from_occupancy(double stp_inactive, double occupancy) {
  // extrusion_barriers_impl.hpp:106-116 / cli.cpp:843-852
  if (occupancy == 0) return 0.0;
  const double to_active = 1.0 - stp_inactive;
  const double to_inactive = (to_active - (occupancy * to_active)) / occupancy;
  return std::clamp(1.0 - to_inactive, 0.0, 1.0);
}

double modle_b200_occupancy_from_stp(double stp_active, double stp_inactive) {
  // extrusion_barriers_impl.hpp:118-128 / cli.cpp:854-863
  if (stp_active + stp_inactive == 0) return 0.0;
  const double to_active = 1.0 - stp_inactive;
  const double to_inactive = 1.0 - stp_active;
  return std::clamp(to_active / (to_active + to_inactive), 0.0, 1.0);
}

// Cli::transform_args, src/modle/cli.cpp:886-1016 (paths/IO members excluded)
int modle_b200_transform_params(modle_b200_sim_params* p, int rev_speed_given, int fwd_speed_given,
                                int barrier_occupancy_given) {
  if (!p) return modle_b200::fail(MODLE_B200_ERR_INVALID_ARGUMENT, "params is NULL");
  if (p->bin_size == 0) return modle_b200::fail(MODLE_B200_ERR_INVALID_ARGUMENT, "bin_size == 0");
  // cli_update_extr_speed :886-910
  if (!rev_speed_given) p->rev_extrusion_speed = p->bin_size * 8 / 10;
  if (!fwd_speed_given) p->fwd_extrusion_speed = p->bin_size * 8 / 10;
  if (p->fwd_extrusion_speed_std > 0 && p->fwd_extrusion_speed_std < 1)
    p->fwd_extrusion_speed_std *= static_cast<double>(p->fwd_extrusion_speed);
  if (p->rev_extrusion_speed_std > 0 && p->rev_extrusion_speed_std < 1)
    p->rev_extrusion_speed_std *= static_cast<double>(p->rev_extrusion_speed);
  p->rev_extrusion_speed_burnin = static_cast<u64>(
      std::round(p->burnin_speed_coefficient * static_cast<double>(p->rev_extrusion_speed)));
  p->fwd_extrusion_speed_burnin = static_cast<u64>(
      std::round(p->burnin_speed_coefficient * static_cast<double>(p->fwd_extrusion_speed)));
  // cli_compute_prob_of_lef_release :913-920
  p->prob_of_lef_release =
      static_cast<double>(p->rev_extrusion_speed + p->fwd_extrusion_speed) /
      static_cast<double>(p->avg_lef_processivity);
  p->prob_of_lef_release_burnin =
      static_cast<double>(p->rev_extrusion_speed_burnin + p->fwd_extrusion_speed_burnin) /
      static_cast<double>(p->avg_lef_processivity);
  // cli_update_barrier_stp_and_occupancy :923-937
  if (barrier_occupancy_given) {
    p->barrier_occupied_stp = modle_b200_stp_active_from_occupancy(p->barrier_not_occupied_stp,
                                                                   p->extrusion_barrier_occupancy);
  } else {
    p->extrusion_barrier_occupancy =
        modle_b200_occupancy_from_stp(p->barrier_occupied_stp, p->barrier_not_occupied_stp);
  }
  // cli_update_tad_to_loop_contact_ratio :970-983
  const bool loop = p->contact_sampling_strategy & MODLE_B200_SAMPLE_LOOP;
  const bool tad = p->contact_sampling_strategy & MODLE_B200_SAMPLE_TAD;
  if (!loop && !tad)
    return modle_b200::fail(MODLE_B200_ERR_INVALID_ARGUMENT,
                            "contact_sampling_strategy needs the tad and/or loop bit");
  if (loop && !tad) p->tad_to_loop_contact_ratio = 0;
  if (!loop && tad) p->tad_to_loop_contact_ratio = std::numeric_limits<double>::infinity();
  // cli_update_burnin_params :985-991
  const u64 burnin_speed = p->rev_extrusion_speed_burnin + p->fwd_extrusion_speed_burnin;
  if (burnin_speed == 0)
    return modle_b200::fail(MODLE_B200_ERR_INVALID_ARGUMENT, "burn-in extrusion speed is 0");
  p->burnin_target_epochs_for_lef_activation =
      std::min<u64>(p->max_burnin_epochs, 5 * p->avg_lef_processivity / burnin_speed);
  // cli_normalize_probabilities :939-968
  if (p->normalize_probabilities) {
    const double ratio = static_cast<double>(p->rev_extrusion_speed + p->fwd_extrusion_speed) /
                         static_cast<double>(p->probability_normalization_factor);
    if (ratio != 1.0) {
      auto stable_pow = [](double base, double e) {
        if (base == 0.0) return 0.0;
        if (base == 1.0) return 1.0;
        return std::exp(std::log(base) * e);
      };
      p->barrier_not_occupied_stp = stable_pow(p->barrier_not_occupied_stp, ratio);
      p->barrier_occupied_stp = modle_b200_stp_active_from_occupancy(
          p->barrier_not_occupied_stp, p->extrusion_barrier_occupancy);
      const double bp = p->probability_of_extrusion_unit_bypass;
      if (bp != 0.0 && bp != 1.0)
        p->probability_of_extrusion_unit_bypass = std::min(bp * ratio, 1.0);
      p->lef_bar_major_collision_pblock = stable_pow(p->lef_bar_major_collision_pblock, ratio);
      p->lef_bar_minor_collision_pblock = stable_pow(p->lef_bar_minor_collision_pblock, ratio);
    }
  }
  if (barrier_occupancy_given) p->override_extrusion_barrier_occupancy = 1;  // :1007-1011
  if (p->stopping_criterion == MODLE_B200_STOP_SIMULATION_EPOCHS)
    p->target_contact_density = -1;  // :1013-1015
  return MODLE_B200_OK;
}

// Simulation::compute_num_lefs, simulation.cpp:1086-1090
uint64_t modle_b200_compute_num_lefs(const modle_b200_sim_params* p, uint64_t interval_size_bp) {
  const double mbp = static_cast<double>(interval_size_bp) / 1.0e6;
  return std::max<u64>(1, static_cast<u64>(std::round(p->number_of_lefs_per_mbp * mbp)));
}

// Simulation::compute_contacts_per_epoch, simulation.cpp:1076-1084
uint64_t modle_b200_compute_contacts_per_epoch(const modle_b200_sim_params* p,
                                               uint64_t num_lefs) {
  const double speed = static_cast<double>(p->rev_extrusion_speed + p->fwd_extrusion_speed);
  const double prob = speed / static_cast<double>(p->contact_sampling_interval);
  return static_cast<u64>(std::max(1.0, std::round(static_cast<double>(num_lefs) * prob)));
}

// ContactMatrixDense(length, diagonal_width, bin_size), contact_matrix_dense_impl.hpp:39-50
void modle_b200_band_shape(const modle_b200_sim_params* p, uint64_t interval_size_bp,
                           uint64_t* nrows, uint64_t* ncols) {
  const u64 nc = (interval_size_bp + p->bin_size - 1) / p->bin_size;
  const u64 nr = std::min((p->diagonal_width + p->bin_size - 1) / p->bin_size, nc);
  if (nrows) *nrows = nr;
  if (ncols) *ncols = nc;
}

// GenomicInterval::hash(state, seed), genome.cpp:201-224
int modle_b200_interval_hash(const char* chrom_name, size_t name_len, uint64_t chrom_size,
                             uint64_t start, uint64_t end, uint64_t seed, uint64_t* out) {
  if (!chrom_name || !out)
    return modle_b200::fail(MODLE_B200_ERR_INVALID_ARGUMENT, "NULL argument");
  std::string key(chrom_name, name_len);
  const u64 words[3] = {chrom_size, start, end};
  key.append(reinterpret_cast<const char*>(words), sizeof(words));
  if (!mh::xxh3::hash(reinterpret_cast<const unsigned char*>(key.data()), key.size(), seed, out))
    return modle_b200::fail(MODLE_B200_ERR_UNSUPPORTED,
                            "chromosome names longer than 216 bytes are not supported");
  return MODLE_B200_OK;
}

void modle_b200_rng_seed(uint64_t seed, uint64_t state[4]) {
  const mh::Xoshiro g = mh::Xoshiro::seeded(seed);
  std::memcpy(state, g.s, sizeof(g.s));
}
uint64_t modle_b200_rng_next(uint64_t state[4]) {
  mh::Xoshiro g;
  std::memcpy(g.s, state, sizeof(g.s));
  const u64 r = g.next();
  std::memcpy(state, g.s, sizeof(g.s));
  return r;
}
void modle_b200_rng_jump(uint64_t state[4]) {
  mh::Xoshiro g;
  std::memcpy(g.s, state, sizeof(g.s));
  g.jump();
  std::memcpy(state, g.s, sizeof(g.s));
}

// Simulation::run_simulate, scheduler_simulate.cpp:104-160: one engine per interval seeded with
// the interval hash (:108); per-cell contact targets (:129-141); the task takes a COPY of the
// engine, then the engine jumps (:143-158).
int modle_b200_make_cell_tasks(const modle_b200_sim_params* p, const char* chrom_name,
                               size_t name_len, const modle_b200_interval* interval,
                               modle_b200_cell_task* tasks) {
  if (!p || !interval || !tasks)
    return modle_b200::fail(MODLE_B200_ERR_INVALID_ARGUMENT, "NULL argument");
  if (interval->end <= interval->start || interval->end > interval->chrom_size)
    return modle_b200::fail(MODLE_B200_ERR_INVALID_ARGUMENT, "invalid interval bounds");
  u64 h = 0;
  if (const int rc = modle_b200_interval_hash(chrom_name, name_len, interval->chrom_size,
                                              interval->start, interval->end, p->seed, &h);
      rc != MODLE_B200_OK)
    return rc;
  mh::Xoshiro eng = mh::Xoshiro::seeded(h);
  u64 nrows = 0, ncols = 0;
  modle_b200_band_shape(p, interval->end - interval->start, &nrows, &ncols);
  const bool epochs_mode = p->stopping_criterion == MODLE_B200_STOP_SIMULATION_EPOCHS;
  // GenomicInterval::npixels() is ContactMatrixLazy's ncols * nrows (genome_impl.hpp:21,96), whose
  // nrows is ceil(diagonal_width / bin_size) (genome.cpp:39-41) -- NOT clamped to ncols the way
  // ContactMatrixDense clamps its own (contact_matrix_dense_impl.hpp:39-44): an interval shorter
  // than the diagonal width still gets its contact target from the unclamped product.
  const u64 nrows_lazy = (p->diagonal_width + p->bin_size - 1) / p->bin_size;
  (void)nrows;
  const u64 total = epochs_mode ? 0
                                : static_cast<u64>(std::round(static_cast<double>(nrows_lazy * ncols) *
                                                              p->target_contact_density));
  const u64 per_cell = (total + p->num_cells - 1) / p->num_cells;
  u64 assigned = 0;
  for (u64 cell = 0; cell < p->num_cells; ++cell) {
    modle_b200_cell_task& t = tasks[cell];
    t.cell_id = cell;
    t.num_target_contacts = std::min(per_cell, total - assigned);
    assigned += t.num_target_contacts;
    t.num_target_epochs =
        epochs_mode ? p->target_simulation_epochs : std::numeric_limits<u64>::max();
    std::memcpy(t.rng_state, eng.s, sizeof(eng.s));
    eng.jump();
  }
  return MODLE_B200_OK;
}

}  // extern "C"
