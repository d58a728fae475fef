// TEST INFRASTRUCTURE ONLY -- CPU emulation of the CUDA kernel's per-cell code.
// Compiles modle_b200/csrc/sim_core.hpp with MODLE_B200_EMU (plain g++): every "thread region"
// becomes a loop over virtual thread ids, so the data-parallel restatement of the reference's
// routines can be compared with the sequential oracle on the CPU, for any virtual CTA width.
// It checks the ALGORITHM of the kernel, not the device build; the `-m gpu` tests do that.
// Never linked into the product library.
#define MODLE_B200_EMU 1
#include <memory>
#include <vector>

#include "../../modle_b200/csrc/launch_prep.hpp"
#include "../../modle_b200/csrc/sim_core.hpp"

using namespace modle_b200;
namespace {
std::vector<u64> g_emu_jump[kJumpSlots];
static thread_local int g_emu_barrier_lut = 1;  // emu_set_barrier_lut
}

namespace {

struct EmuCell {
  KernelParams kp;
  IntervalHostData hd;
  IntervalData D;
  ZigguratTables zig;
  std::vector<u64> arrays;  // 8-byte aligned backing store
  std::vector<u64> ring, gstate;
  std::unique_ptr<CellShared> shared;
  CellArrays A;
  

  std::string setup(const modle_b200_sim_params& p, const modle_b200_interval& iv,
                    const modle_b200_barrier* bars, size_t nb, int staging) {
    StagingConfig sc = staging == 1   ? staging_small()
                       : staging == 2 ? staging_large()
                       : staging == 3 ? staging_mid()
                       : staging == 4 ? staging_large_wide()
                                      : pick_staging(static_cast<u32>(iv.num_lefs),
                                                     static_cast<u32>(nb));
    // (with the barrier look-up table unless emu_set_barrier_lut(0): both forms of the LEF-BAR walk)
    const std::string err = prepare_interval(p, iv, bars, nb, sc, &kp, &hd,
                                             g_emu_barrier_lut ? size_t(227) * 1024 : 0);
    if (!err.empty()) return err;
    if (g_emu_jump[sc.jump_slot].empty()) {
      g_emu_jump[sc.jump_slot].resize(kJumpTableWords);
      build_jump_table(sc.window, g_emu_jump[sc.jump_slot].data());
    }
    D.jump_tbl = g_emu_jump[sc.jump_slot].data();
    D.bar_pos = hd.bar_pos.data();
    D.bar_dir_rev = hd.bar_dir_rev.data();
    D.bar_stp_active = hd.stp_active.data();
    D.bar_stp_inactive = hd.stp_inactive.data();
    D.bar_occupancy = hd.occupancy.data();
    D.zig_nx = zig.nx;
    D.zig_ny = zig.ny;
    D.zig_ex = zig.ex;
    D.zig_ey = zig.ey;
    arrays.assign(cell_array_bytes(kp.n_lefs, kp.n_bar, kp.lut_entries) / 8 + 2, 0);
    A = carve_cell_arrays(arrays.data(), kp.n_lefs, kp.n_bar, kp.lut_entries);
    ring.assign(2 * size_t(kp.rng_window), 0);
    gstate.assign(4 * size_t(kp.rng_gen_threads), 0);
    A.rng_ring = ring.data();
    A.rng_state = gstate.data();
    shared = std::make_unique<CellShared>();
    return std::string();
  }
};

CellTaskDev to_task(const modle_b200_cell_task& t) {
  CellTaskDev d;
  d.cell_id = t.cell_id;
  d.target_epochs = t.num_target_epochs;
  d.target_contacts = t.num_target_contacts;
  for (int i = 0; i < 4; ++i) d.rng_state[i] = t.rng_state[i];
  return d;
}

thread_local std::string g_emu_error;
thread_local int g_emu_rng_mode = MODLE_B200_RNG_REFERENCE_ORDER;

// runs one cell in the selected mode (deterministic or throughput, see modle_b200_set_rng_mode)
void run_cell(EmuCell& cell, const Sinks& K, int virtual_threads, const modle_b200_cell_task& t) {
  Cta cta{&cell.shared->scratch, virtual_threads};
  if (g_emu_rng_mode == MODLE_B200_RNG_COUNTER) {
    CellSimThroughput sim{cell.kp, cell.D, cell.A, *cell.shared, K, cta, to_task(t)};
    sim.run();
  } else {
    CellSim sim{cell.kp, cell.D, cell.A, *cell.shared, K, cta, to_task(t)};
    sim.run();
  }
}

}  // namespace

extern "C" {

const char* emu_last_error() { return g_emu_error.c_str(); }

// 0 ascending, 1 descending, 2 shuffled per region (cta.hpp: emu_thread_order)
void emu_set_thread_order(int mode) { emu_thread_order() = mode; }
// bind_lefs takes its sequential redo in every epoch (cta.hpp: emu_force_bind_redo)
void emu_set_force_bind_redo(int on) { emu_force_bind_redo() = on; }
// CTA barriers the device build would have executed in this thread's calls so far (cta.hpp)
u64 emu_barrier_count_get(int reset) {
  const u64 n = emu_barrier_count();
  if (reset) emu_barrier_count() = 0;
  return n;
}
void emu_phase_barriers_get(u64* out, int n, int reset) {
  for (int i = 0; i < n && i < 64; ++i) {
    out[i] = emu_phase_barriers()[i];
    if (reset) emu_phase_barriers()[i] = 0;
  }
}
// MODLE_B200_RNG_* of the calling thread's later emu_simulate_* / emu_snapshot_cell calls
void emu_set_barrier_lut(int on) { g_emu_barrier_lut = on; }
void emu_set_rng_mode(int mode) { g_emu_rng_mode = mode; }

int emu_simulate_interval_logged(const modle_b200_sim_params* params,
                                 const modle_b200_interval* interval,
                                 const modle_b200_barrier* barriers, size_t num_barriers,
                                 const modle_b200_cell_task* tasks, size_t num_cells,
                                 u32* band_out, u64* occ1d_out, modle_b200_cell_stats* stats_out,
                                 u64* missed_updates_out, int virtual_threads, int staging,
                                 modle_b200_epoch_record* log_out, size_t log_cap);

int emu_simulate_interval(const modle_b200_sim_params* params, const modle_b200_interval* interval,
                          const modle_b200_barrier* barriers, size_t num_barriers,
                          const modle_b200_cell_task* tasks, size_t num_cells, u32* band_out,
                          u64* occ1d_out, modle_b200_cell_stats* stats_out,
                          u64* missed_updates_out, int virtual_threads, int staging) {
  return emu_simulate_interval_logged(params, interval, barriers, num_barriers, tasks, num_cells,
                                      band_out, occ1d_out, stats_out, missed_updates_out,
                                      virtual_threads, staging, nullptr, 0);
}

int emu_simulate_interval_logged(const modle_b200_sim_params* params,
                                 const modle_b200_interval* interval,
                                 const modle_b200_barrier* barriers, size_t num_barriers,
                                 const modle_b200_cell_task* tasks, size_t num_cells,
                                 u32* band_out, u64* occ1d_out, modle_b200_cell_stats* stats_out,
                                 u64* missed_updates_out, int virtual_threads, int staging,
                                 modle_b200_epoch_record* log_out, size_t log_cap) {
  EmuCell cell;
  g_emu_error = cell.setup(*params, *interval, barriers, num_barriers, staging);
  if (!g_emu_error.empty()) return -1;
  u64 missed_local = 0;
  Sinks K{band_out, cell.kp.track_1d ? occ1d_out : nullptr,
          missed_updates_out ? missed_updates_out : &missed_local};
  for (size_t c = 0; c < num_cells; ++c) {
    modle_b200_cell_stats st{};
    K.log = (log_out && log_cap) ? log_out + c * log_cap : nullptr;  // as k_simulate_cells does
    K.log_cap = static_cast<u32>(log_cap);
    if (cell.kp.stop_on_epochs || tasks[c].num_target_contacts != 0) {
      run_cell(cell, K, virtual_threads, tasks[c]);
      st.num_contacts = cell.shared->num_contacts;
      st.num_epochs = cell.shared->epoch;
      st.num_burnin_epochs = cell.shared->num_burnin_epochs;
      st.num_lef_updates = cell.shared->lef_updates;
      st.num_rng_draws = cell.shared->rng_pos;
      st.device_fault = cell.shared->fault;
    }
    if (stats_out) stats_out[c] = st;
  }
  return 0;
}

int emu_snapshot_cell(const modle_b200_sim_params* params, const modle_b200_interval* interval,
                      const modle_b200_barrier* barriers, size_t num_barriers,
                      const modle_b200_cell_task* task, modle_b200_cell_snapshot* snap,
                      modle_b200_cell_stats* stats_out, int virtual_threads, int staging) {
  EmuCell cell;
  g_emu_error = cell.setup(*params, *interval, barriers, num_barriers, staging);
  if (!g_emu_error.empty()) return -1;
  std::vector<u32> band(size_t(cell.kp.nrows) * cell.kp.ncols + 1, 0);
  std::vector<u64> occ(cell.kp.ncols, 0);
  u64 missed = 0;
  Sinks K{band.data(), occ.data(), &missed};
  run_cell(cell, K, virtual_threads, *task);
  const CellShared& S = *cell.shared;
  if (stats_out) {
    stats_out->num_contacts = S.num_contacts;
    stats_out->num_epochs = S.epoch;
    stats_out->num_burnin_epochs = S.num_burnin_epochs;
    stats_out->num_lef_updates = S.lef_updates;
    stats_out->num_rng_draws = S.rng_pos;
    stats_out->device_fault = S.fault;
  }
  auto widen = [](u32 v) { return v == kUnbound ? ~u64(0) : u64(v); };
  for (u32 i = 0; i < cell.kp.n_lefs; ++i) {
    snap->rev_pos[i] = widen(cell.A.rev[i]);
    snap->fwd_pos[i] = widen(cell.A.fwd[i]);
    snap->binding_epoch[i] = widen(cell.A.ep[i]);
    snap->rev_ranks[i] = cell.A.rr[i];
    snap->fwd_ranks[i] = cell.A.fr[i];
  }
  for (u32 i = 0; i < cell.kp.n_bar; ++i) snap->barrier_active[i] = cell.A.bar_active[i];
  snap->num_active_lefs = S.num_active;
  snap->burnin_completed = S.burnin_completed;
  return 0;
}

// Collision pipeline on caller-provided state (all LEFs must be bound: the kernel's epoch body
// never sees an unbound active LEF). steps bitmask as in oracle_collision_steps, except that
// adjust (1) and clamp (2) always run together and the two move corrections (32, 64) too.
int emu_collision_steps(u32 steps, u64 start, u64 end, size_t n, u64* rev, u64* fwd, const u64* ep,
                        u64* rr, u64* fr, u64* rm, u64* fm, u32* rc, u32* fc, size_t nb,
                        const u64* bar_pos, const u8* bar_dir, const u8* bar_active,
                        double prob_bypass, double pblock_major, double pblock_minor, u64 rng_seed,
                        u64* info_out, int virtual_threads) {
  modle_b200_sim_params p;
  modle_b200_default_params(&p);
  modle_b200_transform_params(&p, 0, 0, 0);
  p.probability_of_extrusion_unit_bypass = prob_bypass;
  p.lef_bar_major_collision_pblock = pblock_major;
  p.lef_bar_minor_collision_pblock = pblock_minor;
  modle_b200_interval iv{end, start, end, n};
  std::vector<modle_b200_barrier> bars(nb);
  for (size_t i = 0; i < nb; ++i) {
    bars[i].pos = bar_pos[i];
    bars[i].stp_active = 1.0;
    bars[i].stp_inactive = 0.0;
    bars[i].blocking_direction = bar_dir[i];
    bars[i].reserved_ = 0;
  }
  EmuCell cell;
  g_emu_error = cell.setup(p, iv, bars.data(), nb, 1);
  if (!g_emu_error.empty()) return -1;
  u64 missed = 0;
  Sinks K{nullptr, nullptr, &missed};
  Cta cta{&cell.shared->scratch, virtual_threads};
  modle_b200_cell_task t{};
  modle_b200_rng_seed(rng_seed, t.rng_state);
  CellSim sim{cell.kp, cell.D, cell.A, *cell.shared, K, cta, to_task(t)};
  sim.init_cell();
  CellShared& S = *cell.shared;
  S.rng_pos = 0;  // init_cell consumed the barrier-state draws; the goldens start at draw 0
  S.num_active = static_cast<u32>(n);
  S.burnin_completed = 1;
  for (size_t i = 0; i < n; ++i) {
    if (ep[i] == ~u64(0)) {
      g_emu_error = "emu_collision_steps: unbound LEFs are outside the kernel's precondition";
      return -2;
    }
    cell.A.rev[i] = static_cast<u32>(rev[i]);
    cell.A.fwd[i] = static_cast<u32>(fwd[i]);
    cell.A.ep[i] = static_cast<u32>(ep[i]);
    cell.A.rr[i] = static_cast<u16>(rr[i]);
    cell.A.fr[i] = static_cast<u16>(fr[i]);
    cell.A.rm[i] = static_cast<u32>(rm[i]);
    cell.A.fm[i] = static_cast<u32>(fm[i]);
    cell.A.rc[i] = rc[i];
    cell.A.fc[i] = fc[i];
  }
  for (size_t i = 0; i < nb; ++i) cell.A.bar_active[i] = bar_active[i];
  S.n5 = 0;
  S.n3 = 0;
  if (steps & 3) sim.adjust_and_clamp_moves();
  if (steps & 4) sim.detect_boundaries_leader();
  if (steps & 8) sim.detect_lef_bar_collisions();
  if (steps & 16) sim.detect_primary_lef_lef_collisions();
  if (steps & 96) sim.correct_moves();
  if (steps & 128) sim.process_secondary_lef_lef_collisions();
  if (steps & 256) sim.fix_secondary_lef_lef_collisions();
  for (size_t i = 0; i < n; ++i) {
    rev[i] = cell.A.rev[i];
    fwd[i] = cell.A.fwd[i];
    rr[i] = cell.A.rr[i];
    fr[i] = cell.A.fr[i];
    rm[i] = cell.A.rm[i];
    fm[i] = cell.A.fm[i];
    rc[i] = cell.A.rc[i];
    fc[i] = cell.A.fc[i];
  }
  if (info_out) {
    info_out[0] = S.n5;
    info_out[1] = S.n3;
    info_out[2] = S.rng_pos;
    info_out[3] = S.fault;
  }
  return 0;
}

// One burn-in step of the kernel source on LEFs with the given loop sizes (all active and bound):
// out = {mean loop size, coefficient of variation} as pushed into the burn-in history.
int emu_loop_size_stats(const u64* loop_sizes, size_t n, double* out, int virtual_threads) {
  modle_b200_sim_params p;
  modle_b200_default_params(&p);
  modle_b200_transform_params(&p, 0, 0, 0);
  u64 end = 16;
  for (size_t i = 0; i < n; ++i) end = std::max<u64>(end, loop_sizes[i] + 16);
  modle_b200_interval iv{end, 0, end, n};
  EmuCell cell;
  g_emu_error = cell.setup(p, iv, nullptr, 0, 0);
  if (!g_emu_error.empty()) return -1;
  u64 missed = 0;
  Sinks K{nullptr, nullptr, &missed};
  Cta cta{&cell.shared->scratch, virtual_threads};
  modle_b200_cell_task t{};
  CellSim sim{cell.kp, cell.D, cell.A, *cell.shared, K, cta, to_task(t)};
  sim.init_cell();
  CellShared& S = *cell.shared;
  S.num_active = static_cast<u32>(n);
  for (size_t i = 0; i < n; ++i) {
    cell.A.rev[i] = 8;
    cell.A.fwd[i] = static_cast<u32>(8 + loop_sizes[i]);
    cell.A.ep[i] = 0;
  }
  sim.burnin_step();
  if (S.hist_len != 1 || S.fault) return -2;
  out[0] = S.avg_hist[S.hist_head];
  out[1] = S.cv_hist[S.hist_head];
  return 0;
}

int emu_rank_lefs(const u64* rev, const u64* fwd, const u64* ep, u64* rr, u64* fr, size_t n,
                  int virtual_threads) {
  modle_b200_sim_params p;
  modle_b200_default_params(&p);
  modle_b200_transform_params(&p, 0, 0, 0);
  u64 end = 1;
  for (size_t i = 0; i < n; ++i) end = std::max<u64>(end, fwd[i] + 2);
  modle_b200_interval iv{end, 0, end, n};
  EmuCell cell;
  g_emu_error = cell.setup(p, iv, nullptr, 0, 1);
  if (!g_emu_error.empty()) return -1;
  u64 missed = 0;
  Sinks K{nullptr, nullptr, &missed};
  Cta cta{&cell.shared->scratch, virtual_threads};
  modle_b200_cell_task t{};
  CellSim sim{cell.kp, cell.D, cell.A, *cell.shared, K, cta, to_task(t)};
  cell.shared->num_active = static_cast<u32>(n);
  for (size_t i = 0; i < n; ++i) {
    cell.A.rev[i] = static_cast<u32>(rev[i]);
    cell.A.fwd[i] = static_cast<u32>(fwd[i]);
    cell.A.ep[i] = static_cast<u32>(ep[i]);
    cell.A.rr[i] = static_cast<u16>(rr[i]);
    cell.A.fr[i] = static_cast<u16>(fr[i]);
  }
  sim.rank_lefs();
  for (size_t i = 0; i < n; ++i) {
    rr[i] = cell.A.rr[i];
    fr[i] = cell.A.fr[i];
  }
  return 0;
}

// n Normal(speed, sd) moves through the kernel's draw_normal_moves (all LEFs bound), starting at
// draw 0 of PRNG state `state`; returns the number of raw draws consumed.
long long emu_sample_moves(const u64* state, size_t n, double speed, double sd, u64* moves_out,
                           int virtual_threads, int staging) {
  modle_b200_sim_params p;
  modle_b200_default_params(&p);
  modle_b200_transform_params(&p, 0, 0, 0);
  modle_b200_interval iv{1000000, 0, 1000000, n};
  EmuCell cell;
  g_emu_error = cell.setup(p, iv, nullptr, 0, staging);
  if (!g_emu_error.empty()) return -1;
  u64 missed = 0;
  Sinks K{nullptr, nullptr, &missed};
  Cta cta{&cell.shared->scratch, virtual_threads};
  modle_b200_cell_task t{};
  for (int i = 0; i < 4; ++i) t.rng_state[i] = state[i];
  cell.kp.rev_std = sd;
  auto go = [&](auto& sim) {
    sim.init_cell();
    cell.shared->num_active = static_cast<u32>(n);
    for (size_t i = 0; i < n; ++i) {
      cell.A.rev[i] = cell.A.fwd[i] = 500000;
      cell.A.ep[i] = 0;
    }
    sim.draw_normal_moves(static_cast<u32>(n), static_cast<u32>(n), speed, speed);
  };
  if (g_emu_rng_mode == MODLE_B200_RNG_COUNTER) {
    CellSimThroughput sim{cell.kp, cell.D, cell.A, *cell.shared, K, cta, to_task(t)};
    sim.ctr_keys();
    go(sim);
  } else {
    CellSim sim{cell.kp, cell.D, cell.A, *cell.shared, K, cta, to_task(t)};
    go(sim);
  }
  for (size_t i = 0; i < n; ++i) moves_out[i] = cell.A.rm[i];
  if (cell.shared->fault) return -100 - static_cast<long long>(cell.shared->fault);
  return static_cast<long long>(cell.shared->rng_pos);
}

// Throughput mode: the raw draw `k` of (epoch, phase, item) for the cell whose task carries PRNG
// state `state` (sim_core.hpp ctr_pack / raw), and the mixer on its own.
u64 emu_ctr_draw(const u64* state, u64 epoch, u32 phase, u32 item, u32 k) {
  KernelParams kp{};
  IntervalData D{};
  CellArrays A{};
  auto S = std::make_unique<CellShared>();
  Sinks K{};
  Cta cta{&S->scratch, 1};
  CellTaskDev t{};
  for (int i = 0; i < 4; ++i) t.rng_state[i] = state[i];
  CellSimThroughput sim{kp, D, A, *S, K, cta, t};
  sim.ctr_keys();
  return sim.raw(ctr_pack(epoch, phase, item) + k);
}
u64 emu_mix64(u64 x) { return mix64(x); }
// The kernel's division by a per-interval constant (sim_core.hpp div_u64); magic == 0: the
// general reciprocal of inv_u64, else the caller's (uniform_int: range + 1 for the bucket).
u64 emu_div_u64(u64 x, u64 d, u64 magic) {
  return div_u64(x, magic ? InvU64{d, magic} : inv_u64(d));
}
u64 emu_uniform_int_bucket(u64 range) { return uniform_int_bucket(range); }

// The kernel's collision-word helpers (sim_types.hpp), for the reference's encoding KATs.
u32 emu_collision_word(u64 idx, u32 event) { return coll_make(static_cast<u32>(idx), event); }
void emu_collision_query(u32 word, u32 kind, u64* out) {
  out[0] = coll_index(word);
  out[1] = coll_event(word);
  out[2] = coll_occurred(word);
  out[3] = !coll_occurred(word) && word != 0;
  out[4] = coll_is(word, kind);
  out[5] = coll_avoided(word, kind);
}

}  // extern "C"
