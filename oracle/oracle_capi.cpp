// TEST INFRASTRUCTURE ONLY -- C entry points of the CPU oracle (see oracle_rng.hpp header).
// Loaded with ctypes by tests/, __graft_entry__.smoke() and bench.py (cpu_baseline /
// --impl reference). Uses the POD structs of include/modle_b200.h so that the same inputs can be
// handed to the oracle and to the CUDA library.
#include <atomic>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../include/modle_b200.h"
#include "oracle_sim.hpp"

using namespace oracle;

namespace {

Params to_params(const modle_b200_sim_params& a, u64 nlefs) {
  Params p;
  p.bin_size = a.bin_size;
  p.diagonal_width = a.diagonal_width;
  p.rev_speed = static_cast<double>(a.rev_extrusion_speed);
  p.fwd_speed = static_cast<double>(a.fwd_extrusion_speed);
  p.rev_speed_burnin = static_cast<double>(a.rev_extrusion_speed_burnin);
  p.fwd_speed_burnin = static_cast<double>(a.fwd_extrusion_speed_burnin);
  p.rev_speed_std = a.rev_extrusion_speed_std;
  p.fwd_speed_std = a.fwd_extrusion_speed_std;
  p.prob_lef_release = a.prob_of_lef_release;
  p.prob_lef_release_burnin = a.prob_of_lef_release_burnin;
  p.hard_stall_multiplier = a.hard_stall_lef_stability_multiplier;
  p.soft_stall_multiplier = a.soft_stall_lef_stability_multiplier;
  p.prob_bypass = a.probability_of_extrusion_unit_bypass;
  p.pblock_major = a.lef_bar_major_collision_pblock;
  p.pblock_minor = a.lef_bar_minor_collision_pblock;
  p.tad_to_loop_ratio = a.tad_to_loop_contact_ratio;
  p.gev_mu = a.genextreme_mu;
  p.gev_sigma = a.genextreme_sigma;
  p.gev_xi = a.genextreme_xi;
  p.noisify = (a.contact_sampling_strategy & MODLE_B200_SAMPLE_NOISIFY) ? 1 : 0;
  p.track_1d = a.track_1d_lef_position ? 1 : 0;
  p.skip_burnin = a.skip_burnin ? 1 : 0;
  p.stop_on_epochs = a.stopping_criterion == MODLE_B200_STOP_SIMULATION_EPOCHS;
  // Simulation::compute_contacts_per_epoch (simulation.cpp:1076-1084)
  const double speed = static_cast<double>(a.rev_extrusion_speed + a.fwd_extrusion_speed);
  const double prob = speed / static_cast<double>(a.contact_sampling_interval);
  p.contacts_per_epoch =
      static_cast<u64>(std::max(1.0, std::round(static_cast<double>(nlefs) * prob)));
  p.burnin_history = a.burnin_history_length;
  p.burnin_window = a.burnin_smoothing_window_size;
  p.min_burnin_epochs = a.min_burnin_epochs;
  p.max_burnin_epochs = a.max_burnin_epochs;
  p.burnin_activation_epochs = a.burnin_target_epochs_for_lef_activation;
  p.debug_max_epochs = a.debug_max_epochs;
  return p;
}

Barriers to_barriers(const modle_b200_barrier* b, std::size_t nb) {
  Barriers B;
  B.pos.resize(nb);
  B.dir.resize(nb);
  B.stp_active.resize(nb);
  B.stp_inactive.resize(nb);
  B.active.assign(nb, 0);
  for (std::size_t i = 0; i < nb; ++i) {
    B.pos[i] = b[i].pos;
    B.dir[i] = static_cast<std::uint8_t>(b[i].blocking_direction);
    B.stp_active[i] = b[i].stp_active;
    B.stp_inactive[i] = b[i].stp_inactive;
  }
  return B;
}

CellTask to_task(const modle_b200_cell_task& t) {
  CellTask c;
  c.cell_id = t.cell_id;
  c.target_contacts = t.num_target_contacts;
  c.target_epochs = t.num_target_epochs;
  std::memcpy(c.rng_state, t.rng_state, sizeof(c.rng_state));
  return c;
}

void fill_stats(modle_b200_cell_stats& s, const CellResult& r) {
  s.num_contacts = r.num_contacts;
  s.num_epochs = r.epochs;
  s.num_burnin_epochs = r.burnin_epochs;
  s.num_lef_updates = r.lef_updates;
  s.num_rng_draws = r.rng_draws;
  s.device_fault = 0;
}

}  // namespace

extern "C" {

void oracle_rng_seed(u64 seed, u64* state) {
  const Rng g = Rng::from_seed(seed);
  std::memcpy(state, g.s, 32);
}
u64 oracle_rng_next(u64* state) {
  Rng g = Rng::from_state(state);
  const u64 r = g.next();
  std::memcpy(state, g.s, 32);
  return r;
}
void oracle_rng_jump(u64* state) {
  Rng g = Rng::from_state(state);
  g.jump();
  std::memcpy(state, g.s, 32);
}
// advances the state by n plain steps (to cross-check jump-ahead implementations)
void oracle_rng_discard(u64* state, u64 n) {
  Rng g = Rng::from_state(state);
  for (u64 i = 0; i < n; ++i) g.next();
  std::memcpy(state, g.s, 32);
}

int oracle_xxh3_64(const unsigned char* data, std::size_t len, u64 seed, u64* out) {
  return xxh3::hash64(data, len, seed, out) ? 0 : -1;
}

// GenomicInterval::hash (genome.cpp:201-224)
int oracle_interval_hash(const char* name, std::size_t name_len, u64 chrom_size, u64 start, u64 end,
                         u64 seed, u64* out) {
  std::string buf(name, name_len);
  for (u64 v : {chrom_size, start, end}) {
    for (int i = 0; i < 8; ++i) buf.push_back(static_cast<char>((v >> (8 * i)) & 0xFF));
  }
  return oracle_xxh3_64(reinterpret_cast<const unsigned char*>(buf.data()), buf.size(), seed, out);
}

// Simulation::compute_num_lefs (simulation.cpp:1086-1090)
u64 oracle_compute_num_lefs(double lefs_per_mbp, u64 size_bp) {
  const double size_mbp = static_cast<double>(size_bp) / 1.0e6;
  return std::max<u64>(1, static_cast<u64>(std::round(lefs_per_mbp * size_mbp)));
}

// run_simulate's per-interval fan-out (scheduler_simulate.cpp:104-160)
int oracle_make_cell_tasks(const modle_b200_sim_params* p, const char* name, std::size_t name_len,
                           const modle_b200_interval* iv, modle_b200_cell_task* tasks) {
  u64 h = 0;
  if (oracle_interval_hash(name, name_len, iv->chrom_size, iv->start, iv->end, p->seed, &h) != 0)
    return -1;
  Rng g = Rng::from_seed(h);
  const Band band = Band::make(iv->end - iv->start, p->diagonal_width, p->bin_size);
  const bool epochs_mode = p->stopping_criterion == MODLE_B200_STOP_SIMULATION_EPOCHS;
  // interval.npixels() (scheduler_simulate.cpp:129) = ContactMatrixLazy::npixels()
  // (genome_impl.hpp:21,96): ncols * ceil(diagonal_width / bin_size), without the
  // min(nrows, ncols) clamp of ContactMatrixDense
  const u64 npixels_lazy = band.ncols * ((p->diagonal_width + p->bin_size - 1) / p->bin_size);
  const u64 tot = epochs_mode ? 0
                              : static_cast<u64>(std::round(static_cast<double>(npixels_lazy) *
                                                            p->target_contact_density));
  const u64 per_cell = (tot + p->num_cells - 1) / p->num_cells;
  u64 rolling = 0;
  for (u64 c = 0; c < p->num_cells; ++c) {
    const u64 tgt = std::min(per_cell, tot - rolling);
    rolling += tgt;
    tasks[c].cell_id = c;
    tasks[c].num_target_epochs = epochs_mode ? p->target_simulation_epochs : ~u64(0);
    tasks[c].num_target_contacts = tgt;
    std::memcpy(tasks[c].rng_state, g.s, 32);
    g.jump();
  }
  return 0;
}

// kind: 0 bernoulli(p0) 1 canonical 2 uniform01 3 uniform_int[a=p0,b=p1] 4 unit_normal
//       5 normal(p0,p1) 6 unit_exponential 7 poisson(p0) 8 binomial(t=p0,p=p1) 9 gev(p0,p1,p2)
//      10 raw next()
void oracle_sample(int kind, double p0, double p1, double p2, u64* state, std::size_t n,
                   double* out, u64* draws_used) {
  Rng g = Rng::from_state(state);
  for (std::size_t i = 0; i < n; ++i) {
    double v = 0;
    switch (kind) {
      case 0: v = bernoulli(g, p0) ? 1.0 : 0.0; break;
      case 1: v = canonical(g); break;
      case 2: v = uniform01(g); break;
      case 3: v = static_cast<double>(uniform_int(g, static_cast<u64>(p0), static_cast<u64>(p1))); break;
      case 4: v = unit_normal(g); break;
      case 5: v = normal(g, p0, p1); break;
      case 6: v = unit_exponential(g); break;
      case 7: v = static_cast<double>(poisson(g, p0)); break;
      case 8: v = static_cast<double>(binomial(g, static_cast<i64>(p0), p1)); break;
      case 9: v = genextreme(g, p0, p1, p2); break;
      default: v = static_cast<double>(g.next() >> 11); break;
    }
    out[i] = v;
  }
  std::memcpy(state, g.s, 32);
  if (draws_used) *draws_used = g.ndraws;
}

void oracle_zig_tables(double* nx, double* ny, double* ex, double* ey) {
  const ZigTables& t = zig();
  std::memcpy(nx, t.nx, sizeof(t.nx));
  std::memcpy(ny, t.ny, sizeof(t.ny));
  std::memcpy(ex, t.ex, sizeof(t.ex));
  std::memcpy(ey, t.ey, sizeof(t.ey));
}

// The burn-in statistics of compute_loop_size_stats (simulation.cpp:795-819) on given loop sizes:
// out = {mean, standard deviation} (the reference's stats KATs, test/units/stats/descriptive_test.cpp)
void oracle_loop_size_stats(const u64* loop_sizes, std::size_t n, double* out) {
  CellSim::loop_size_mean_sd(n, [&](std::size_t i) { return loop_sizes[i]; }, &out[0], &out[1]);
}

void oracle_rank_lefs(const u64* rev, const u64* fwd, const u64* ep, u64* rr, u64* fr,
                      std::size_t n, int init_buffers) {
  rank_lefs(rev, fwd, ep, rr, fr, n, init_buffers != 0);
}

// Runs a subset of the collision pipeline on caller-provided state (golden-vector tests).
// steps bitmask: 1 adjust_moves, 2 clamp_moves, 4 detect_units_at_interval_boundaries,
// 8 detect_lef_bar, 16 detect_primary, 32 correct_lef_bar, 64 correct_primary,
// 128 process_secondary, 256 fix_secondary. When step 4 is not run, n5 = n3 = 0 (the default
// arguments the reference's test shims use, simulation.hpp:330-400).
void oracle_collision_steps(u32 steps, u64 start, u64 end, std::size_t n, u64* rev, u64* fwd,
                            const u64* ep, u64* rr, u64* fr, u64* rm, u64* fm, u32* rc, u32* fc,
                            std::size_t nb, const u64* bar_pos, const std::uint8_t* bar_dir,
                            const std::uint8_t* bar_active, double prob_bypass,
                            double pblock_major, double pblock_minor, u64 rng_seed,
                            u64* n5_n3_out) {
  Params p;
  p.prob_bypass = prob_bypass;
  p.pblock_major = pblock_major;
  p.pblock_minor = pblock_minor;
  Barriers B;
  B.pos.assign(bar_pos, bar_pos + nb);
  B.dir.assign(bar_dir, bar_dir + nb);
  B.active.assign(bar_active, bar_active + nb);
  B.stp_active.assign(nb, 1.0);
  B.stp_inactive.assign(nb, 0.0);
  Rng g = Rng::from_seed(rng_seed);
  CollisionCtx c;
  c.p = &p;
  c.iv = Interval{start, end};
  c.bars = &B;
  c.rev = rev;
  c.fwd = fwd;
  c.ep = ep;
  c.rr = rr;
  c.fr = fr;
  c.rm = rm;
  c.fm = fm;
  c.rc = rc;
  c.fc = fc;
  c.n = n;
  c.g = &g;
  u64 n5 = 0, n3 = 0;
  if (steps & 1) adjust_moves(c.iv, rev, fwd, ep, rr, fr, rm, fm, n);
  if (steps & 2) clamp_moves(c.iv, rev, fwd, ep, rm, fm, n);
  if (steps & 4) {
    const auto r = detect_units_at_interval_boundaries(c);
    n5 = r.first;
    n3 = r.second;
  }
  if (steps & 8) detect_lef_bar_collisions(c, n5, n3);
  if (steps & 16) detect_primary_lef_lef_collisions(c, n5, n3);
  if (steps & 32) correct_moves_for_lef_bar_collisions(c);
  if (steps & 64) correct_moves_for_primary_lef_lef_collisions(c);
  if (steps & 128) process_secondary_lef_lef_collisions(c, n5, n3);
  if (steps & 256) fix_secondary_lef_lef_collisions(c, n5, n3);
  if (n5_n3_out) {
    n5_n3_out[0] = n5;
    n5_n3_out[1] = n3;
    n5_n3_out[2] = g.ndraws;
  }
}

// CPU counterpart of modle_b200_simulate_interval: `nthreads` worker threads pop cells from a
// shared counter (one cell per task, as the reference's workers do) and add into the shared band
// with atomic increments. Returns 0.
int oracle_simulate_interval_logged(const modle_b200_sim_params* params,
                                    const modle_b200_interval* interval,
                                    const modle_b200_barrier* barriers, std::size_t num_barriers,
                                    const modle_b200_cell_task* tasks, std::size_t num_cells,
                                    u32* band_out, u64* occ1d_out,
                                    modle_b200_cell_stats* stats_out, u64* missed_updates_out,
                                    int nthreads, modle_b200_epoch_record* log_out,
                                    std::size_t log_cap) {
  const Params p = to_params(*params, interval->num_lefs);
  const Barriers B = to_barriers(barriers, num_barriers);
  const Interval iv{interval->start, interval->end};
  ContactSink sink;
  sink.geom = Band::make(iv.end - iv.start, params->diagonal_width, p.bin_size);
  sink.band = band_out;
  sink.occ1d = p.track_1d ? occ1d_out : nullptr;
  u64 missed_local = 0;
  sink.missed = missed_updates_out ? missed_updates_out : &missed_local;
  if (nthreads < 1) nthreads = 1;
  std::atomic<std::size_t> next{0};
  auto worker = [&]() {
    for (;;) {
      const std::size_t c = next.fetch_add(1);
      if (c >= num_cells) return;
      modle_b200_cell_stats st{};
      // tasks without work are skipped (scheduler_simulate.cpp:234-235)
      const bool epochs_mode = p.stop_on_epochs != 0;
      if (epochs_mode || tasks[c].num_target_contacts != 0) {
        CellSim sim(p, iv, B, interval->num_lefs, to_task(tasks[c]), sink);
        if (log_out && log_cap) sim.set_log(log_out + c * log_cap, log_cap);
        fill_stats(st, sim.run());
      }
      if (stats_out) stats_out[c] = st;
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < nthreads; ++t) pool.emplace_back(worker);
  worker();
  for (auto& t : pool) t.join();
  return 0;
}

int oracle_simulate_interval(const modle_b200_sim_params* params,
                             const modle_b200_interval* interval,
                             const modle_b200_barrier* barriers, std::size_t num_barriers,
                             const modle_b200_cell_task* tasks, std::size_t num_cells,
                             u32* band_out, u64* occ1d_out, modle_b200_cell_stats* stats_out,
                             u64* missed_updates_out, int nthreads) {
  return oracle_simulate_interval_logged(params, interval, barriers, num_barriers, tasks,
                                         num_cells, band_out, occ1d_out, stats_out,
                                         missed_updates_out, nthreads, nullptr, 0);
}

// The reference's scheduling for a whole run (scheduler_simulate.cpp:104-160 produce, :190-271
// consume): ONE queue holds the (interval, cell) tasks of every interval in genome order and
// `nthreads` workers pop from it until it is empty, so a worker that finishes a short cell moves
// on to the next task whatever interval it belongs to. All arrays are indexed by interval;
// band / occ1d / stats / missed are caller-owned (zeroed). Used by bench.py's CPU arm.
int oracle_simulate_genome(const modle_b200_sim_params* params, std::size_t num_intervals,
                           const modle_b200_interval* intervals,
                           const modle_b200_barrier* const* barriers,
                           const std::size_t* num_barriers,
                           const modle_b200_cell_task* const* tasks, const std::size_t* num_cells,
                           u32* const* band_out, u64* const* occ1d_out,
                           modle_b200_cell_stats* const* stats_out, u64* missed_updates_out,
                           int nthreads) {
  struct PerInterval {
    Params p;
    Barriers B;
    Interval iv;
    ContactSink sink;
    std::size_t first_task;
  };
  std::vector<PerInterval> ivs(num_intervals);
  std::size_t total = 0;
  for (std::size_t i = 0; i < num_intervals; ++i) {
    PerInterval& x = ivs[i];
    x.p = to_params(*params, intervals[i].num_lefs);
    x.B = to_barriers(barriers[i], num_barriers[i]);
    x.iv = Interval{intervals[i].start, intervals[i].end};
    x.sink.geom = Band::make(x.iv.end - x.iv.start, params->diagonal_width, x.p.bin_size);
    x.sink.band = band_out[i];
    x.sink.occ1d = x.p.track_1d ? occ1d_out[i] : nullptr;
    x.sink.missed = &missed_updates_out[i];
    x.first_task = total;
    total += num_cells[i];
  }
  if (nthreads < 1) nthreads = 1;
  std::atomic<std::size_t> next{0};
  auto worker = [&]() {
    for (;;) {
      const std::size_t t = next.fetch_add(1);
      if (t >= total) return;
      std::size_t i = 0;  // owning interval: last one whose first task is <= t
      {
        std::size_t lo = 0, hi = num_intervals;
        while (hi - lo > 1) {
          const std::size_t mid = (lo + hi) / 2;
          if (ivs[mid].first_task <= t) lo = mid; else hi = mid;
        }
        i = lo;
        while (i + 1 < num_intervals && ivs[i + 1].first_task <= t) ++i;  // (empty intervals)
      }
      const PerInterval& x = ivs[i];
      const std::size_t c = t - x.first_task;
      modle_b200_cell_stats st{};
      const bool epochs_mode = x.p.stop_on_epochs != 0;
      if (epochs_mode || tasks[i][c].num_target_contacts != 0) {
        CellSim sim(x.p, x.iv, x.B, intervals[i].num_lefs, to_task(tasks[i][c]), x.sink);
        fill_stats(st, sim.run());
      }
      if (stats_out && stats_out[i]) stats_out[i][c] = st;
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < nthreads; ++t) pool.emplace_back(worker);
  worker();
  for (auto& t : pool) t.join();
  return 0;
}

// Counters of burnin_margin_note (oracle_sim.hpp): out = {comparisons, within 64 ulp, smallest
// relative gap as double bits}; reset != 0 clears them afterwards.
void oracle_burnin_margin(u64* out, int reset) {
  BurninMargin& m = burnin_margin();
  out[0] = __atomic_load_n(&m.comparisons, __ATOMIC_RELAXED);
  out[1] = __atomic_load_n(&m.within_64_ulp, __ATOMIC_RELAXED);
  out[2] = __atomic_load_n(&m.min_gap_bits, __ATOMIC_RELAXED);
  if (reset) m = BurninMargin{};
}

// Runs one cell for params->debug_max_epochs epochs and dumps its state.
int oracle_snapshot_cell(const modle_b200_sim_params* params, const modle_b200_interval* interval,
                         const modle_b200_barrier* barriers, std::size_t num_barriers,
                         const modle_b200_cell_task* task, modle_b200_cell_snapshot* snap,
                         modle_b200_cell_stats* stats_out) {
  const Params p = to_params(*params, interval->num_lefs);
  const Barriers B = to_barriers(barriers, num_barriers);
  const Interval iv{interval->start, interval->end};
  const Band geom = Band::make(iv.end - iv.start, params->diagonal_width, p.bin_size);
  std::vector<u32> band(geom.npixels() + 1, 0);
  std::vector<u64> occ(geom.ncols, 0);
  u64 missed = 0;
  ContactSink sink;
  sink.geom = geom;
  sink.band = band.data();
  sink.occ1d = occ.data();
  sink.missed = &missed;
  CellSim sim(p, iv, B, interval->num_lefs, to_task(*task), sink);
  const CellResult r = sim.run();
  if (stats_out) fill_stats(*stats_out, r);
  const CellState& s = sim.state();
  for (std::size_t i = 0; i < s.nlefs; ++i) {
    snap->rev_pos[i] = s.rev[i];
    snap->fwd_pos[i] = s.fwd[i];
    snap->binding_epoch[i] = s.epoch_bound[i];
    snap->rev_ranks[i] = s.rev_rank[i];
    snap->fwd_ranks[i] = s.fwd_rank[i];
  }
  for (std::size_t i = 0; i < num_barriers; ++i) snap->barrier_active[i] = sim.barriers().active[i];
  snap->num_active_lefs = s.num_active;
  snap->burnin_completed = s.burnin_completed ? 1 : 0;
  return 0;
}

// Pixel loop of modle::io::internal::append_contact_matrix_to_cooler
// (src/libmodle_io/contact_matrix_dense_io_impl.hpp:50-71) with emplace_pixel's bin offset
// (:30-43) and ContactMatrixDense::unsafe_get (contact_matrix_dense_unsafe_impl.hpp:33-42:
// transpose_coords, "i >= nrows -> 0", encode_idx = col * nrows + row). Returns the number of
// non-zero pixels; writes at most `capacity` of them (pixels may be NULL when capacity is 0).
u64 oracle_band_to_pixels(const u32* band, u64 nrows, u64 ncols, u64 bin_offset,
                          modle_b200_pixel* pixels, u64 capacity) {
  const auto unsafe_get = [&](u64 row, u64 col) -> u32 {
    const u64 i = row > col ? row - col : col - row;  // transpose_coords
    const u64 j = row > col ? row : col;
    if (i >= nrows) return 0;
    return band[j * nrows + i];
  };
  u64 n_out = 0;
  for (u64 i = 0; i < ncols; ++i) {
    for (u64 j = i; j < ncols && j - i < nrows; ++j) {
      if (const u32 n = unsafe_get(i, j); n != 0) {
        if (n_out < capacity)
          pixels[n_out] = modle_b200_pixel{bin_offset + i, bin_offset + j,
                                           static_cast<std::int32_t>(n), 0};
        ++n_out;
      }
    }
  }
  return n_out;
}

// Collision word helpers, for the reference's encoding KATs (collision_encoding_test.cpp:27-176).
// out = {decode_index, decode_event, collision_occurred(), collision_avoided(),
//        collision_occurred(kind), collision_avoided(kind)}; collision_avoided() is
// "!occurred && word != 0" (collision_encoding_impl.hpp:227-230).
u32 oracle_collision_word(u64 idx, u32 event) { return coll_make(idx, event); }
void oracle_collision_query(u32 word, u32 kind, u64* out) {
  out[0] = coll_index(word);
  out[1] = coll_event(word);
  out[2] = coll_occurred(word);
  out[3] = !coll_occurred(word) && word != 0;
  out[4] = coll_occurred(word, kind);
  out[5] = coll_avoided(word, kind);
}

// ContactMatrixDense<u32>::increment semantics (contact_matrix_dense_safe_impl.hpp:54-89) on a
// caller-provided band: returns 1 when the pixel is outside the band (counted as missed).
int oracle_band_increment(u32* band, u64 nrows, u64 ncols, u64 bin_size, u64 b1, u64 b2,
                          u64* missed) {
  ContactSink sink;
  sink.geom.nrows = nrows;
  sink.geom.ncols = ncols;
  (void)bin_size;
  sink.band = band;
  const u64 before = *missed;
  sink.missed = missed;
  sink.increment(b1, b2);
  return *missed != before;
}

}  // extern "C"
