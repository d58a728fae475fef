// A compiled C++ consumer of include/modle_b200.h: what a MoDLE maintainer's binding does, with
// no Python and no ctypes mirror in between. It fills the structs field by field the way
// INTEGRATION.md shows for Simulation::run_simulate (reference seam:
// src/libmodle/cpu/include/modle/simulation.hpp:45-151, Task at :59-69, src/modle/main.cpp:168-171)
// and drives  default/transform params -> make_cell_tasks -> simulate_interval -> band_to_pixels.
//
//   consumer layout            sizeof / offsetof of every struct as JSON (no GPU needed); the
//                              pytest compares it with the ctypes mirror (modle_b200/abi.py)
//   consumer run [cells]       one small interval on GPU 0; prints checksums the pytest compares
//                              with the same run through the ctypes path
//
// TEST INFRASTRUCTURE: built by __graft_entry__.build() / tests/test_cabi_consumer.py.
#include <cinttypes>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "modle_b200.h"

// ---- the layout a binding compiles against (drift between header and mirrors fails HERE) -------
static_assert(sizeof(modle_b200_sim_params) == 320, "modle_b200_sim_params");
static_assert(offsetof(modle_b200_sim_params, rev_extrusion_speed_std) == 48, "");
static_assert(offsetof(modle_b200_sim_params, number_of_lefs_per_mbp) == 152, "");
static_assert(offsetof(modle_b200_sim_params, num_cells) == 272, "");
static_assert(offsetof(modle_b200_sim_params, contact_sampling_strategy) == 288, "");
static_assert(offsetof(modle_b200_sim_params, debug_max_epochs) == 312, "");
static_assert(sizeof(modle_b200_interval) == 32 && offsetof(modle_b200_interval, num_lefs) == 24, "");
static_assert(sizeof(modle_b200_barrier) == 32 && offsetof(modle_b200_barrier, blocking_direction) == 24, "");
static_assert(sizeof(modle_b200_cell_task) == 56 && offsetof(modle_b200_cell_task, rng_state) == 24, "");
static_assert(sizeof(modle_b200_cell_stats) == 48 && offsetof(modle_b200_cell_stats, device_fault) == 40, "");
static_assert(sizeof(modle_b200_cell_snapshot) == 64 && offsetof(modle_b200_cell_snapshot, num_active_lefs) == 48, "");
static_assert(sizeof(modle_b200_epoch_record) == 56 && offsetof(modle_b200_epoch_record, burnin) == 16, "");
static_assert(sizeof(modle_b200_pixel) == 24 && offsetof(modle_b200_pixel, count) == 16, "");
static_assert(sizeof(modle_b200_shard) == 40 && offsetof(modle_b200_shard, weight) == 32, "");
static_assert(MODLE_B200_NUM_PHASES == 26, "");

#define FIELD(S, f) std::printf("%s\"%s\": %zu", first ? "" : ", ", #f, offsetof(S, f)), first = false
#define BEGIN(S) std::printf("%s\"%s\": {\"sizeof\": %zu, \"fields\": {", sfirst ? "" : ", ", #S, sizeof(S)), sfirst = false, first = true
#define END() std::printf("}}")

static int print_layout() {
  bool first = true, sfirst = true;
  std::printf("{");
  BEGIN(modle_b200_sim_params);
  FIELD(modle_b200_sim_params, bin_size); FIELD(modle_b200_sim_params, diagonal_width);
  FIELD(modle_b200_sim_params, rev_extrusion_speed); FIELD(modle_b200_sim_params, fwd_extrusion_speed);
  FIELD(modle_b200_sim_params, rev_extrusion_speed_burnin); FIELD(modle_b200_sim_params, fwd_extrusion_speed_burnin);
  FIELD(modle_b200_sim_params, rev_extrusion_speed_std); FIELD(modle_b200_sim_params, fwd_extrusion_speed_std);
  FIELD(modle_b200_sim_params, prob_of_lef_release); FIELD(modle_b200_sim_params, prob_of_lef_release_burnin);
  FIELD(modle_b200_sim_params, hard_stall_lef_stability_multiplier);
  FIELD(modle_b200_sim_params, soft_stall_lef_stability_multiplier);
  FIELD(modle_b200_sim_params, probability_of_extrusion_unit_bypass);
  FIELD(modle_b200_sim_params, lef_bar_major_collision_pblock);
  FIELD(modle_b200_sim_params, lef_bar_minor_collision_pblock);
  FIELD(modle_b200_sim_params, tad_to_loop_contact_ratio);
  FIELD(modle_b200_sim_params, genextreme_mu); FIELD(modle_b200_sim_params, genextreme_sigma);
  FIELD(modle_b200_sim_params, genextreme_xi); FIELD(modle_b200_sim_params, number_of_lefs_per_mbp);
  FIELD(modle_b200_sim_params, target_contact_density); FIELD(modle_b200_sim_params, target_simulation_epochs);
  FIELD(modle_b200_sim_params, contact_sampling_interval); FIELD(modle_b200_sim_params, avg_lef_processivity);
  FIELD(modle_b200_sim_params, probability_normalization_factor);
  FIELD(modle_b200_sim_params, extrusion_barrier_occupancy); FIELD(modle_b200_sim_params, barrier_occupied_stp);
  FIELD(modle_b200_sim_params, barrier_not_occupied_stp); FIELD(modle_b200_sim_params, burnin_speed_coefficient);
  FIELD(modle_b200_sim_params, burnin_history_length); FIELD(modle_b200_sim_params, burnin_smoothing_window_size);
  FIELD(modle_b200_sim_params, min_burnin_epochs); FIELD(modle_b200_sim_params, max_burnin_epochs);
  FIELD(modle_b200_sim_params, burnin_target_epochs_for_lef_activation);
  FIELD(modle_b200_sim_params, num_cells); FIELD(modle_b200_sim_params, seed);
  FIELD(modle_b200_sim_params, contact_sampling_strategy); FIELD(modle_b200_sim_params, stopping_criterion);
  FIELD(modle_b200_sim_params, track_1d_lef_position); FIELD(modle_b200_sim_params, skip_burnin);
  FIELD(modle_b200_sim_params, normalize_probabilities);
  FIELD(modle_b200_sim_params, override_extrusion_barrier_occupancy);
  FIELD(modle_b200_sim_params, debug_max_epochs);
  END();
  BEGIN(modle_b200_interval);
  FIELD(modle_b200_interval, chrom_size); FIELD(modle_b200_interval, start);
  FIELD(modle_b200_interval, end); FIELD(modle_b200_interval, num_lefs);
  END();
  BEGIN(modle_b200_barrier);
  FIELD(modle_b200_barrier, pos); FIELD(modle_b200_barrier, stp_active);
  FIELD(modle_b200_barrier, stp_inactive); FIELD(modle_b200_barrier, blocking_direction);
  FIELD(modle_b200_barrier, reserved_);
  END();
  BEGIN(modle_b200_cell_task);
  FIELD(modle_b200_cell_task, cell_id); FIELD(modle_b200_cell_task, num_target_epochs);
  FIELD(modle_b200_cell_task, num_target_contacts); FIELD(modle_b200_cell_task, rng_state);
  END();
  BEGIN(modle_b200_cell_stats);
  FIELD(modle_b200_cell_stats, num_contacts); FIELD(modle_b200_cell_stats, num_epochs);
  FIELD(modle_b200_cell_stats, num_burnin_epochs); FIELD(modle_b200_cell_stats, num_lef_updates);
  FIELD(modle_b200_cell_stats, num_rng_draws); FIELD(modle_b200_cell_stats, device_fault);
  END();
  BEGIN(modle_b200_cell_snapshot);
  FIELD(modle_b200_cell_snapshot, rev_pos); FIELD(modle_b200_cell_snapshot, fwd_pos);
  FIELD(modle_b200_cell_snapshot, binding_epoch); FIELD(modle_b200_cell_snapshot, rev_ranks);
  FIELD(modle_b200_cell_snapshot, fwd_ranks); FIELD(modle_b200_cell_snapshot, barrier_active);
  FIELD(modle_b200_cell_snapshot, num_active_lefs); FIELD(modle_b200_cell_snapshot, burnin_completed);
  END();
  BEGIN(modle_b200_epoch_record);
  FIELD(modle_b200_epoch_record, epoch); FIELD(modle_b200_epoch_record, loop_size_sum);
  FIELD(modle_b200_epoch_record, burnin); FIELD(modle_b200_epoch_record, num_lefs);
  FIELD(modle_b200_epoch_record, barriers_occupied); FIELD(modle_b200_epoch_record, lefs_stalled_rev);
  FIELD(modle_b200_epoch_record, lefs_stalled_fwd); FIELD(modle_b200_epoch_record, lefs_stalled_both);
  FIELD(modle_b200_epoch_record, lef_bar_collisions);
  FIELD(modle_b200_epoch_record, lef_lef_primary_collisions);
  FIELD(modle_b200_epoch_record, lef_lef_secondary_collisions); FIELD(modle_b200_epoch_record, reserved_);
  END();
  BEGIN(modle_b200_pixel);
  FIELD(modle_b200_pixel, bin1_id); FIELD(modle_b200_pixel, bin2_id);
  FIELD(modle_b200_pixel, count); FIELD(modle_b200_pixel, reserved_);
  END();
  BEGIN(modle_b200_shard);
  FIELD(modle_b200_shard, interval); FIELD(modle_b200_shard, cell_lo);
  FIELD(modle_b200_shard, cell_hi); FIELD(modle_b200_shard, rank);
  FIELD(modle_b200_shard, reserved_); FIELD(modle_b200_shard, weight);
  END();
  std::printf(", \"abi_version\": %d, \"num_phases\": %d}\n", modle_b200_abi_version(),
              MODLE_B200_NUM_PHASES);
  return 0;
}

static std::uint64_t fnv1a(const void* data, std::size_t nbytes, std::uint64_t h = 1469598103934665603ull) {
  const unsigned char* p = static_cast<const unsigned char*>(data);
  for (std::size_t i = 0; i < nbytes; ++i) h = (h ^ p[i]) * 1099511628211ull;
  return h;
}

#define CHECK(call)                                                                     \
  do {                                                                                  \
    const int rc_ = (call);                                                             \
    if (rc_ != MODLE_B200_OK) {                                                         \
      std::fprintf(stderr, "%s failed: %d: %s\n", #call, rc_, modle_b200_last_error()); \
      return 2;                                                                         \
    }                                                                                   \
  } while (0)

// The inputs are fixed here and restated by tests/test_cabi_consumer.py: chromosome "chrCabi" of
// 6 Mbp, barriers every 97,003 bp from 50,021 with alternating motif strand and occupancies
// 0.70 + 0.01 * (k mod 25), density 0.02, `cells` cells, seed 42.
static int run(std::uint64_t cells) {
  modle_b200_sim_params p;
  modle_b200_default_params(&p);
  p.num_cells = cells;
  p.seed = 42;
  p.target_contact_density = 0.02;
  CHECK(modle_b200_transform_params(&p, 0, 0, 0));

  const char* chrom = "chrCabi";
  modle_b200_interval iv;
  iv.chrom_size = 6000000;
  iv.start = 0;
  iv.end = 6000000;
  iv.num_lefs = modle_b200_compute_num_lefs(&p, iv.end - iv.start);

  std::vector<modle_b200_barrier> bars;
  for (std::uint64_t pos = 50021, k = 0; pos < iv.end; pos += 97003, ++k) {
    modle_b200_barrier b;
    std::memset(&b, 0, sizeof(b));
    b.pos = pos;
    b.stp_inactive = p.barrier_not_occupied_stp;
    b.stp_active = modle_b200_stp_active_from_occupancy(b.stp_inactive, 0.70 + 0.01 * double(k % 25));
    b.blocking_direction = (k % 2 == 0) ? MODLE_B200_DIR_REV : MODLE_B200_DIR_FWD;
    bars.push_back(b);
  }

  std::vector<modle_b200_cell_task> tasks(cells);
  CHECK(modle_b200_make_cell_tasks(&p, chrom, std::strlen(chrom), &iv, tasks.data()));

  std::uint64_t nrows = 0, ncols = 0;
  modle_b200_band_shape(&p, iv.end - iv.start, &nrows, &ncols);
  std::vector<std::uint32_t> band(nrows * ncols + 1, 0);
  std::vector<std::uint64_t> occ(ncols, 0);
  std::vector<modle_b200_cell_stats> stats(cells);
  std::uint64_t missed = 0;

  modle_b200_context* ctx = nullptr;
  CHECK(modle_b200_init(&ctx, 0));
  CHECK(modle_b200_simulate_interval(ctx, &p, &iv, bars.data(), bars.size(), tasks.data(),
                                     tasks.size(), band.data(), occ.data(), stats.data(), &missed));
  std::uint64_t npix = 0;
  CHECK(modle_b200_band_to_pixels(ctx, band.data(), nrows, ncols, 1000, nullptr, 0, &npix));
  std::vector<modle_b200_pixel> pixels(npix);
  CHECK(modle_b200_band_to_pixels(ctx, band.data(), nrows, ncols, 1000, pixels.data(), npix, &npix));
  const std::uint64_t launches = modle_b200_kernel_launches(ctx);
  modle_b200_destroy(ctx);

  std::uint64_t band_sum = 0, contacts = 0, epochs = 0, draws = 0, faults = 0, pix_sum = 0;
  for (std::uint32_t v : band) band_sum += v;
  for (const auto& s : stats) {
    contacts += s.num_contacts;
    epochs += s.num_epochs;
    draws += s.num_rng_draws;
    faults += s.device_fault != 0;
  }
  for (const auto& px : pixels) pix_sum += std::uint64_t(px.count);
  std::printf("{\"num_lefs\": %" PRIu64 ", \"num_barriers\": %zu, \"nrows\": %" PRIu64
              ", \"ncols\": %" PRIu64 ", \"band_sum\": %" PRIu64 ", \"band_hash\": %" PRIu64
              ", \"occ_hash\": %" PRIu64 ", \"missed\": %" PRIu64 ", \"contacts\": %" PRIu64
              ", \"epochs\": %" PRIu64 ", \"rng_draws\": %" PRIu64 ", \"faults\": %" PRIu64
              ", \"num_pixels\": %" PRIu64 ", \"pixel_count_sum\": %" PRIu64
              ", \"pixel_hash\": %" PRIu64 ", \"kernel_launches\": %" PRIu64 "}\n",
              iv.num_lefs, bars.size(), nrows, ncols, band_sum,
              fnv1a(band.data(), band.size() * 4), fnv1a(occ.data(), occ.size() * 8), missed,
              contacts, epochs, draws, faults, npix, pix_sum,
              fnv1a(pixels.data(), pixels.size() * sizeof(modle_b200_pixel)), launches);
  return faults ? 3 : 0;
}

int main(int argc, char** argv) {
  const std::string mode = argc > 1 ? argv[1] : "layout";
  if (mode == "layout") return print_layout();
  if (mode == "run") return run(argc > 2 ? std::strtoull(argv[2], nullptr, 10) : 6);
  std::fprintf(stderr, "usage: consumer layout | run [cells]\n");
  return 1;
}
