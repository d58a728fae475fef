// Host-side preparation of one interval launch: translates the C-ABI structs into the kernel's
// KernelParams / per-interval arrays, builds the ziggurat tables and the RNG jump matrices.
// Pure host code (no CUDA calls) so that the CPU emulation of the kernel can reuse it.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/modle_b200.h"
#include "host_rng.hpp"
#include "sim_types.hpp"

namespace modle_b200 {

// Ziggurat tables of Boost.Random's unit normal (128 layers) and unit exponential (256 layers)
// distributions. boost/random/normal_distribution.hpp and exponential_distribution.hpp hold them
// as 20-digit literals of the exact solution of the ziggurat equations; those headers are not
// available here, so scripts/make_ziggurat_tables.py recomputes the solution with 60 digits and
// writes the same 20-digit literals (the leading entries are pinned by tests/test_oracle_kats.py).
namespace zigdata {
#include "ziggurat_tables.inc"
}  // namespace zigdata
struct ZigguratTables {
  double nx[129], ny[129], ex[257], ey[257];
  ZigguratTables() {
    for (int i = 0; i <= 128; ++i) {
      nx[i] = zigdata::kZigNormalX[i];
      ny[i] = zigdata::kZigNormalY[i];
    }
    for (int i = 0; i <= 256; ++i) {
      ex[i] = zigdata::kZigExpX[i];
      ey[i] = zigdata::kZigExpY[i];
    }
  }
};

// RNG staging configurations: (CTA threads, generator threads G, draws per generator per window l).
struct StagingConfig {
  u32 cta_threads, gen_threads, per_thread, window, jump_slot;
  u32 cells_per_sm;  // launch class: 3 (small), 2 (mid) or 1 (large) CTAs per SM
};
#ifndef MODLE_B200_LARGE_THREADS
#define MODLE_B200_LARGE_THREADS 1024
#endif
// Draws per generator thread per window (l). The window sizes (and with them the jump matrices)
// are fixed per configuration; a smaller l spreads a refill over more generator threads.
inline u32 staging_per_thread() {
  static const u32 l = [] {
    const char* e = std::getenv("MODLE_B200_GEN_PER_THREAD");
    const int v = e ? std::atoi(e) : 128;
    return (v == 32 || v == 64 || v == 128 || v == 256) ? static_cast<u32>(v) : 128u;
  }();
  return l;
}
// one jump table per window size: 32768 -> slot 0, 8192 -> slot 1, 16384 -> slot 2
inline StagingConfig staging_with_window(u32 cta_threads, u32 window) {
  const u32 l = staging_per_thread();
  const u32 jump_slot = window == 32768 ? 0u : (window == 8192 ? 1u : 2u);
  const u32 per_sm = cta_threads == 256 ? 3u : (cta_threads == 512 && window == 16384 ? 2u : 1u);
  return StagingConfig{cta_threads, window / l, l, window, jump_slot, per_sm};
}
inline StagingConfig staging_large() {
  // 16384 draws per window: the rings of all resident CTAs (2 windows each) then stay L2
  // resident -- with 32768 every staged draw was written back to DRAM (19.4 GB instead of 2.4 GB
  // for one chr1 launch, profiles/r01i_*), at the same speed. MODLE_B200_LARGE_WINDOW=32768
  // restores the larger window (needed only above ~7,000 LEFs per cell, which does not fit the
  // shared memory anyway).
  static const u32 w = [] {
    const char* e = std::getenv("MODLE_B200_LARGE_WINDOW");
    const int v = e ? std::atoi(e) : 16384;
    return v == 32768 ? 32768u : 16384u;
  }();
  StagingConfig c = staging_with_window(MODLE_B200_LARGE_THREADS, w);
  c.cells_per_sm = 1;
  return c;
}
inline StagingConfig staging_large_wide() {
  StagingConfig c = staging_with_window(MODLE_B200_LARGE_THREADS, 32768);
  c.cells_per_sm = 1;
  return c;
}
inline StagingConfig staging_small() { return staging_with_window(256, 8192); }
inline StagingConfig staging_mid() { return staging_with_window(512, 16384); }

// Byte-indexed table of T^window (layout: see sim_core.hpp xs_jump): entry (k, v) is the XOR of
// the matrix columns 8k+b over the set bits b of v.
inline void build_jump_table(u32 window, u64* out /*32*256*4*/) {
  const host::StateMatrix m = host::StateMatrix::power(window);
  for (int k = 0; k < 32; ++k) {
    for (int v = 0; v < 256; ++v) {
      u64* e = out + (size_t(k) * 256 + v) * 4;
      e[0] = e[1] = e[2] = e[3] = 0;
      for (int b = 0; b < 8; ++b) {
        if (!((v >> b) & 1)) continue;
        for (int w = 0; w < 4; ++w) e[w] ^= m.col[8 * k + b][w];
      }
    }
  }
}

inline u64 worst_phase_draws(u32 n_lefs) { return 2 * u64(n_lefs) + n_lefs / 4 + 512; }

struct IntervalHostData {
  std::vector<u32> bar_pos, bar_dir_rev;
  std::vector<double> stp_active, stp_inactive, occupancy;
};

// Returns an empty string on success, else the reason the launch is not possible.
inline std::string prepare_interval(const modle_b200_sim_params& p, const modle_b200_interval& iv,
                                    const modle_b200_barrier* bars, size_t nb,
                                    const StagingConfig& sc, KernelParams* kp,
                                    IntervalHostData* hd, size_t smem_limit = 0) {
  if (iv.end <= iv.start) return "empty interval";
  if (iv.end >= 0xFFFFFFF0ull) return "interval end does not fit 32-bit device coordinates";
  if (iv.num_lefs == 0 || iv.num_lefs >= 65535) return "num_lefs must be in [1, 65534]";
  if (nb >= (1u << 24)) return "too many barriers";
  if (p.bin_size == 0 || p.bin_size > 0xFFFFFFFFull) return "invalid bin_size";
  if (p.burnin_history_length > kMaxBurninHistory ||
      p.burnin_smoothing_window_size + 1 >= p.burnin_history_length)
    return "burn-in history/window outside the supported range";
  if (p.burnin_target_epochs_for_lef_activation == 0) return "burnin activation epochs is 0";
  KernelParams k;
  std::memset(&k, 0, sizeof(k));
  k.start = static_cast<u32>(iv.start);
  k.end = static_cast<u32>(iv.end);
  k.n_lefs = static_cast<u32>(iv.num_lefs);
  k.n_bar = static_cast<u32>(nb);
  k.bin_size = static_cast<u32>(p.bin_size);
  u64 nrows = 0, ncols = 0;
  modle_b200_band_shape(&p, iv.end - iv.start, &nrows, &ncols);
  if (nrows * ncols + 1 >= (u64(1) << 32)) return "band matrix too large for 32-bit pixel index";
  k.nrows = static_cast<u32>(nrows);
  k.ncols = static_cast<u32>(ncols);
  k.rev_speed = static_cast<double>(p.rev_extrusion_speed);
  k.fwd_speed = static_cast<double>(p.fwd_extrusion_speed);
  k.rev_speed_burnin = static_cast<double>(p.rev_extrusion_speed_burnin);
  k.fwd_speed_burnin = static_cast<double>(p.fwd_extrusion_speed_burnin);
  k.rev_std = p.rev_extrusion_speed_std;
  k.fwd_std = p.fwd_extrusion_speed_std;
  k.p_release = p.prob_of_lef_release;
  k.p_release_burnin = p.prob_of_lef_release_burnin;
  k.hard_mult = p.hard_stall_lef_stability_multiplier;
  k.soft_mult = p.soft_stall_lef_stability_multiplier;
  k.p_bypass = p.probability_of_extrusion_unit_bypass;
  k.pblock_major = p.lef_bar_major_collision_pblock;
  k.pblock_minor = p.lef_bar_minor_collision_pblock;
  k.tad_to_loop = p.tad_to_loop_contact_ratio;
  k.gev_mu = p.genextreme_mu;
  k.gev_sigma = p.genextreme_sigma;
  k.gev_xi = p.genextreme_xi;
  k.noisify = (p.contact_sampling_strategy & MODLE_B200_SAMPLE_NOISIFY) ? 1 : 0;
  k.track_1d = p.track_1d_lef_position ? 1 : 0;
  k.skip_burnin = p.skip_burnin ? 1 : 0;
  k.stop_on_epochs = p.stopping_criterion == MODLE_B200_STOP_SIMULATION_EPOCHS ? 1 : 0;
  const u64 cpe = modle_b200_compute_contacts_per_epoch(&p, iv.num_lefs);
  if (cpe > 0x7FFFFFFFull) return "contacts per epoch too large";
  k.contacts_per_epoch = static_cast<u32>(cpe);
  k.burnin_history = static_cast<u32>(p.burnin_history_length);
  k.burnin_window = static_cast<u32>(p.burnin_smoothing_window_size);
  k.min_burnin_epochs = p.min_burnin_epochs;
  k.max_burnin_epochs = p.max_burnin_epochs;
  // simulate_one_cell, simulation.cpp:906-908
  k.lef_binding_rate_burnin = static_cast<double>(iv.num_lefs) /
                              static_cast<double>(p.burnin_target_epochs_for_lef_activation);
  k.debug_max_epochs = p.debug_max_epochs;
  {
    // speed + 64 sigma: no draw of the ziggurat sampler gets there in practice; the kernel still
    // checks every move against it and falls back to the true maximum when one does
    const double hi = std::max({k.rev_speed + 64.0 * k.rev_std, k.fwd_speed + 64.0 * k.fwd_std,
                                k.rev_speed_burnin + 64.0 * k.rev_std,
                                k.fwd_speed_burnin + 64.0 * k.fwd_std}) + 2.0;
    if (!(hi < 16000000.0))
      return "extrusion speed too large: moves must stay below 2^24 bp per epoch";
    k.move_bound = static_cast<u32>(hi);
  }
  k.rng_gen_threads = sc.gen_threads;
  k.rng_per_thread = sc.per_thread;
  k.rng_window = sc.window;
  k.rng_jump_slot = sc.jump_slot;
  // Barrier look-up table: the finest bucket (>= 4 kb) whose table still fits next to the cell's
  // arrays under `smem_limit` (what the launch class allows per CTA); dropped when there is no
  // room for buckets of at most 16 average barrier spacings (the walk then keeps its cursor).
  k.lut_shift = 0;
  k.lut_entries = 0;
  if (smem_limit != 0 && nb != 0 && nb < 65535) {
    const size_t base = ((sizeof(CellShared) + 15) / 16) * 16 +
                        cell_array_bytes(static_cast<u32>(iv.num_lefs), static_cast<u32>(nb));
    const u64 span = iv.end - iv.start;
    for (u32 shift = 12; shift <= 24; ++shift) {
      const u64 entries = (span >> shift) + 2;
      if (entries > 60000) continue;
      if (base + ((entries + 1) / 2) * 4 + 32 > smem_limit) continue;
      if ((u64(1) << shift) > 16 * (span / nb + 1)) break;  // too coarse to be worth it
      k.lut_shift = shift;
      k.lut_entries = static_cast<u32>(entries);
      break;
    }
  }
  // every phase must fit one staging window (largest consumer: the normal draws of both move
  // arrays, 2n items + n/4 + 64 slack + 192), and a thread's chunk of them one 32-bit mask
  const u64 worst = worst_phase_draws(static_cast<u32>(iv.num_lefs));
  if (worst > sc.window || nb + 64 > sc.window) return "interval too large for the RNG staging window";

  hd->bar_pos.resize(nb);
  hd->bar_dir_rev.assign((nb + 31) / 32 + 1, 0);
  hd->stp_active.resize(nb);
  hd->stp_inactive.resize(nb);
  hd->occupancy.resize(nb);
  for (size_t i = 0; i < nb; ++i) {
    // A barrier outside [start, end) is legal: the reference keeps a record that overlaps a
    // --genomic-intervals range while its midpoint falls outside it (add_extrusion_barriers only
    // asserts, genome.cpp:285-294); no unit can reach it, but it still takes its draws.
    if (bars[i].pos >= 0xFFFFFFF0ull) return "barrier position does not fit 32-bit coordinates";
    if (i && bars[i].pos < bars[i - 1].pos) return "barriers are not sorted by position";
    if (bars[i].blocking_direction != MODLE_B200_DIR_REV &&
        bars[i].blocking_direction != MODLE_B200_DIR_FWD)
      return "invalid barrier blocking direction";
    hd->bar_pos[i] = static_cast<u32>(bars[i].pos);
    if (bars[i].blocking_direction == MODLE_B200_DIR_REV)
      hd->bar_dir_rev[i >> 5] |= 1u << (i & 31);
    hd->stp_active[i] = bars[i].stp_active;
    hd->stp_inactive[i] = bars[i].stp_inactive;
    // ExtrusionBarriers::occupancy (extrusion_barriers.cpp:140-143)
    hd->occupancy[i] = modle_b200_occupancy_from_stp(bars[i].stp_active, bars[i].stp_inactive);
  }
  *kp = k;
  return std::string();
}

inline StagingConfig pick_staging(u32 n_lefs, u32 n_bar) {
  // small intervals: several CTAs per SM, few generator threads each; large: one fat CTA per SM
  const size_t bytes = cell_array_bytes(n_lefs, n_bar) + sizeof(CellShared);
  const StagingConfig s = staging_small();
  const u64 worst = worst_phase_draws(n_lefs);
  // shared memory per CTA that still lets 3 / 2 CTAs share an SM (228 KB per SM, 1 KB reserved
  // per CTA): 75 KB / 113 KB
  static const int mid_mode = [] {
    const char* e = std::getenv("MODLE_B200_MID");
    return e ? std::atoi(e) : 1;
  }();
  const size_t small_limit = mid_mode ? size_t(75) * 1024 : size_t(100) * 1024;
  if (bytes <= small_limit && worst <= s.window && u64(n_bar) + 64 <= s.window) return s;
  const StagingConfig m = staging_mid();
  if (mid_mode && bytes <= size_t(113) * 1024 && worst <= m.window && u64(n_bar) + 64 <= m.window)
    return m;
  const StagingConfig l = staging_large();
  if (worst <= l.window && u64(n_bar) + 64 <= l.window) return l;
  return staging_large_wide();  // e.g. more than 16k barriers in one interval
}

}  // namespace modle_b200
