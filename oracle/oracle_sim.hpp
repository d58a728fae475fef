// TEST INFRASTRUCTURE ONLY -- CPU oracle for the modle_b200 hot path (see oracle_rng.hpp header
// for the rules and the parity status). Sequential, behavioural restatement of the reference's
// per-(interval, cell) simulation loop. Every function cites the reference lines it follows
// (paths relative to /root/reference). State is kept as structure-of-arrays instead of the
// reference's Lef / Collision objects.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdint>
#include <deque>
#include <limits>
#include <vector>

#include "oracle_rng.hpp"

namespace oracle {

constexpr u64 UNBOUND = ~u64(0);  // extrusion_factors_impl.hpp:96-98,120-124

// Collision word. The reference packs an event byte above a 56-bit index
// (src/libmodle/cpu/include/modle/collision_encoding.hpp:61-65,91-96). Here the same flags sit in
// bits 24..31 of a u32 and the index in bits 0..23.
enum : u32 {
  EV_COLLISION = 0x10,
  EV_CHROM_BOUNDARY = 0x08,
  EV_LEF_BAR = 0x04,
  EV_LEF_LEF_PRIMARY = 0x02,
  EV_LEF_LEF_SECONDARY = 0x01,
};
constexpr u32 coll_make(u64 idx, u32 ev) noexcept { return (ev << 24) | static_cast<u32>(idx); }
constexpr u32 coll_event(u32 c) noexcept { return c >> 24; }
constexpr u32 coll_index(u32 c) noexcept { return c & 0x00FFFFFFu; }
constexpr bool coll_occurred(u32 c) noexcept { return (c >> 24) & EV_COLLISION; }
// collision_encoding_impl.hpp:222-242
constexpr bool coll_occurred(u32 c, u32 kind) noexcept {
  return coll_event(c) == (kind | EV_COLLISION);
}
constexpr bool coll_avoided(u32 c, u32 kind) noexcept {
  return !coll_occurred(c) && coll_event(c) == kind;
}

enum : std::uint8_t { DIR_NONE = 0, DIR_REV = 1, DIR_FWD = 2 };  // common/dna.hpp:77-80

// Transformed simulation parameters: the fields of modle::Config the path reads
// (src/common/include/modle/common/simulation_config.hpp:53-113) after Cli::transform_args.
struct Params {
  u64 bin_size = 5000;
  u64 diagonal_width = 3000000;
  double rev_speed = 4000, fwd_speed = 4000;  // bp/epoch (integers stored as double)
  double rev_speed_burnin = 4000, fwd_speed_burnin = 4000;
  double rev_speed_std = 200, fwd_speed_std = 200;
  double prob_lef_release = 8000.0 / 300000.0;
  double prob_lef_release_burnin = 8000.0 / 300000.0;
  double hard_stall_multiplier = 5.0, soft_stall_multiplier = 1.0;
  double prob_bypass = 0.1;
  double pblock_major = 1.0, pblock_minor = 0.0;
  double tad_to_loop_ratio = 5.0;
  double gev_mu = 0, gev_sigma = 5000, gev_xi = 0.001;
  int noisify = 1;
  int track_1d = 1;
  int skip_burnin = 0;
  int stop_on_epochs = 0;  // 0: contact-density criterion, 1: simulation-epochs criterion
  u64 contacts_per_epoch = 1;
  u64 burnin_history = 100, burnin_window = 5;
  u64 min_burnin_epochs = 0, max_burnin_epochs = ~u64(0);
  u64 burnin_activation_epochs = 187;
  u64 debug_max_epochs = ~u64(0);  // oracle/GPU debug aid: stop after this many epochs
};

struct Barriers {
  std::vector<u64> pos;
  std::vector<std::uint8_t> dir;  // blocking direction (DIR_REV / DIR_FWD)
  std::vector<double> stp_active, stp_inactive;
  std::vector<std::uint8_t> active;  // per-cell mutable state
  std::size_t size() const { return pos.size(); }
};

// extrusion_barriers_impl.hpp:118-128
inline double occupancy_from_stp(double stp_active, double stp_inactive) noexcept {
  if (stp_active + stp_inactive == 0) return 0.0;
  const double tp_i2a = 1.0 - stp_inactive;
  const double tp_a2i = 1.0 - stp_active;
  return std::clamp(tp_i2a / (tp_i2a + tp_a2i), 0.0, 1.0);
}
// extrusion_barriers_impl.hpp:106-116
inline double stp_active_from_occupancy(double stp_inactive, double occupancy) noexcept {
  if (occupancy == 0) return 0.0;
  const double tp_i2a = 1.0 - stp_inactive;
  const double tp_a2i = (tp_i2a - (occupancy * tp_i2a)) / occupancy;
  return std::clamp(1.0 - tp_a2i, 0.0, 1.0);
}

struct Interval {
  u64 start = 0, end = 0;  // [start, end)
};

// Banded matrix geometry: contact_matrix_dense_impl.hpp:39-50, internal_impl.hpp:19-46
struct Band {
  u64 nrows = 0, ncols = 0;
  static Band make(u64 length, u64 diagonal_width, u64 bin_size) {
    Band b;
    b.ncols = (length + bin_size - 1) / bin_size;
    b.nrows = std::min((diagonal_width + bin_size - 1) / bin_size, b.ncols);
    return b;
  }
  u64 npixels() const { return nrows * ncols; }
};

// Per-cell scratch (Simulation::State, simulation.hpp:71-135, simulation.cpp:603-627)
struct CellState {
  std::size_t nlefs = 0;
  std::vector<u64> rev, fwd, epoch_bound;
  std::vector<u64> rev_rank, fwd_rank;
  std::vector<u64> rev_move, fwd_move;
  std::vector<u32> rev_coll, fwd_coll;
  std::vector<std::uint8_t> mask;
  std::deque<double> cv_buff, avg_buff;
  u64 epoch = 0, num_burnin_epochs = 0, num_active = 0, num_contacts = 0;
  bool burnin_completed = false;
  u64 lef_updates = 0;  // Σ_epochs num_active (metric bookkeeping, not in the reference)

  void reset(std::size_t n) {
    nlefs = n;
    rev.assign(n, UNBOUND);
    fwd.assign(n, UNBOUND);
    epoch_bound.assign(n, UNBOUND);
    rev_rank.resize(n);
    fwd_rank.resize(n);
    for (std::size_t i = 0; i < n; ++i) rev_rank[i] = fwd_rank[i] = i;
    rev_move.assign(n, 0);
    fwd_move.assign(n, 0);
    rev_coll.assign(n, 0);
    fwd_coll.assign(n, 0);
    mask.assign(n, 0);
    cv_buff.clear();
    avg_buff.clear();
    epoch = num_burnin_epochs = num_active = num_contacts = lef_updates = 0;
    burnin_completed = false;
  }
  bool bound(std::size_t i) const { return epoch_bound[i] != UNBOUND; }
};

// ------------------------------------------------------------------------------------------
// rank_lefs (simulation.cpp:410-496). The reference sorts with an unstable sort and then fixes
// runs of equal positions with an insertion sort on binding_epoch (ascending for rev, descending
// for fwd). Order among equal (pos, epoch) pairs is implementation-defined there; the oracle
// defines it as "stable w.r.t. the previous rank order".
// ------------------------------------------------------------------------------------------
inline void rank_lefs(const u64* rev, const u64* fwd, const u64* ep, u64* rr, u64* fr,
                      std::size_t n, bool init_buffers = false) {
  if (init_buffers) {
    for (std::size_t i = 0; i < n; ++i) rr[i] = fr[i] = i;
  }
  std::stable_sort(rr, rr + n, [&](u64 a, u64 b) { return rev[a] < rev[b]; });
  std::stable_sort(fr, fr + n, [&](u64 a, u64 b) { return fwd[a] < fwd[b]; });
  for (std::size_t i = 1; i < n; ++i) {
    if (rev[rr[i - 1]] == rev[rr[i]]) {
      const std::size_t b = i - 1;
      for (; i < n; ++i) {
        if (rev[rr[i - 1]] != rev[rr[i]]) break;
      }
      std::stable_sort(rr + b, rr + i, [&](u64 a, u64 c) { return ep[a] < ep[c]; });
    }
  }
  for (std::size_t i = 1; i < n; ++i) {
    if (fwd[fr[i - 1]] == fwd[fr[i]]) {
      const std::size_t b = i - 1;
      for (; i < n; ++i) {
        if (fwd[fr[i - 1]] != fwd[fr[i]]) break;
      }
      std::stable_sort(fr + b, fr + i, [&](u64 a, u64 c) { return ep[c] < ep[a]; });
    }
  }
}

// simulation.cpp:350-407
inline void adjust_moves(const Interval& iv, const u64* rev, const u64* fwd, const u64* ep,
                         const u64* rr, const u64* fr, u64* rm, u64* fm, std::size_t n) {
  auto bound = [&](u64 i) { return ep[i] != UNBOUND; };
  for (std::size_t i = n - 1; i > 0; --i) {
    const u64 i1 = rr[i - 1], i2 = rr[i];
    if (bound(i1) && bound(i2)) {
      if (rev[i1] <= iv.start + rm[i1] || rev[i2] <= iv.start + rm[i2]) continue;
      const u64 p1 = rev[i1] - rm[i1];
      const u64 p2 = rev[i2] - rm[i2];
      if (p2 <= p1) rm[i1] += (p1 - p2) + 1;
    }
  }
  for (std::size_t i = 1; i < n; ++i) {
    const u64 i1 = fr[i - 1], i2 = fr[i];
    if (bound(i1) && bound(i2)) {
      if (fwd[i1] + fm[i1] > iv.end - 1 || fwd[i2] + fm[i2] > iv.end - 1) continue;
      const u64 p1 = fwd[i1] + fm[i1];
      const u64 p2 = fwd[i2] + fm[i2];
      if (p1 >= p2) fm[i2] += (p1 - p2) + 1;
    }
  }
}

// simulation.cpp:332-347
inline void clamp_moves(const Interval& iv, const u64* rev, const u64* fwd, const u64* ep, u64* rm,
                        u64* fm, std::size_t n) {
  for (std::size_t i = 0; i < n; ++i) {
    if (ep[i] == UNBOUND) continue;
    rm[i] = std::min(rm[i], rev[i] - iv.start);
    fm[i] = std::min(fm[i], iv.end - fwd[i] - 1);
  }
}

// simulation.cpp:523-551 -> (rev collision pos, fwd collision pos)
inline std::pair<u64, u64> lef_lef_collision_pos(u64 rev_pos, u64 fwd_pos, u64 rev_move,
                                                 u64 fwd_move) noexcept {
  const u64 relative_speed = rev_move + fwd_move;
  const double time_to_collision =
      static_cast<double>(rev_pos - fwd_pos) / static_cast<double>(relative_speed);
  const u64 collision_pos =
      fwd_pos + static_cast<u64>(std::round(static_cast<double>(fwd_move) * time_to_collision));
  if (collision_pos == fwd_pos) return {collision_pos + 1, collision_pos};
  return {collision_pos, collision_pos - 1};
}

struct CollisionCtx {
  const Params* p;
  Interval iv;
  const Barriers* bars;
  u64 *rev, *fwd;
  const u64* ep;
  u64 *rr, *fr;
  u64 *rm, *fm;
  u32 *rc, *fc;
  std::size_t n;
  Rng* g;
  bool bound(u64 i) const { return ep[i] != UNBOUND; }
  // simulation_impl.hpp:93-101
  bool lef_lef_trial() const {
    return p->prob_bypass == 0.0 || bernoulli(*g, 1.0 - p->prob_bypass);
  }
  bool lef_bar_trial(double pblock) const { return pblock == 1.0 || bernoulli(*g, pblock); }
};

inline u64 sat_dec(u64 x) noexcept { return std::min(x, x - 1); }  // the reference's idiom

// simulation_detect_collisions.cpp:25-120
inline std::pair<u64, u64> detect_units_at_interval_boundaries(CollisionCtx& c) {
  const std::size_t n = c.n;
  u64 n5 = 0, n3 = 0;
  const u64 first_fwd_pos = c.fwd[c.fr[0]];
  u64 last_rev_pos = 0;
  for (std::size_t k = n; k-- > 0;) {
    if (c.bound(c.rr[k])) {
      last_rev_pos = c.rev[c.rr[k]];
      break;
    }
  }
  for (std::size_t i = 0; i < n; ++i) {
    const u64 idx = c.rr[i];
    const u64 pos = c.rev[idx];
    const u64 mv = c.rm[idx];
    if (pos == c.iv.start) {
      ++n5;
      c.rc[idx] = coll_make(5, EV_COLLISION | EV_CHROM_BOUNDARY);
    } else if (pos > first_fwd_pos) {
      break;
    } else if (pos - mv == c.iv.start) {
      c.rc[idx] = coll_make(5, EV_COLLISION | EV_CHROM_BOUNDARY);
      ++n5;
      break;
    }
  }
  for (std::size_t i = n - 1; i > 0; --i) {
    const u64 idx = c.fr[i];
    const u64 pos = c.fwd[idx];
    const u64 mv = c.fm[idx];
    if (!c.bound(idx)) {
      ++n3;
      continue;
    }
    if (pos == c.iv.end - 1) {
      ++n3;
      c.fc[idx] = coll_make(3, EV_COLLISION | EV_CHROM_BOUNDARY);
    } else if (pos < last_rev_pos) {
      break;
    } else if (pos + mv == c.iv.end - 1) {
      c.fc[idx] = coll_make(3, EV_COLLISION | EV_CHROM_BOUNDARY);
      ++n3;
      break;
    }
  }
  return {n5, n3};
}

// simulation_detect_collisions.cpp:123-247
inline void detect_lef_bar_collisions(CollisionCtx& c, u64 n5, u64 n3) {
  const Barriers& B = *c.bars;
  const std::size_t n = c.n;
  if (B.size() == 0) return;  // the reference asserts !barriers.empty() (:219)
  {
    std::size_t j = sat_dec(n5);
    u64 idx = c.rr[j];
    u64 pos = c.rev[idx];
    bool done = false;
    for (std::size_t i = 0; i < B.size() && !done; ++i) {
      if (!B.active[i]) continue;
      const double pblock = B.dir[i] == DIR_REV ? c.p->pblock_major : c.p->pblock_minor;
      while (pos <= B.pos[i]) {
        if (++j == n) {
          done = true;
          break;
        }
        idx = c.rr[j];
        pos = c.rev[idx];
      }
      if (done) break;
      if (c.bound(idx)) {
        const u64 delta = pos - B.pos[i];
        if (delta > 0 && delta <= c.rm[idx] && c.lef_bar_trial(pblock)) {
          c.rc[idx] = coll_make(i, EV_COLLISION | EV_LEF_BAR);
        }
      }
    }
  }
  {
    std::size_t j = n - sat_dec(n3);
    u64 idx = c.fr[--j];
    u64 pos = c.fwd[idx];
    for (std::size_t i = B.size(); i-- > 0;) {
      if (!B.active[i]) continue;
      const double pblock = B.dir[i] == DIR_FWD ? c.p->pblock_major : c.p->pblock_minor;
      while (pos >= B.pos[i]) {
        if (j-- == 0) return;
        idx = c.fr[j];
        pos = c.fwd[idx];
      }
      if (c.bound(idx)) {
        const u64 delta = B.pos[i] - pos;
        if (delta > 0 && delta <= c.fm[idx] && c.lef_bar_trial(pblock)) {
          c.fc[idx] = coll_make(i, EV_COLLISION | EV_LEF_BAR);
        }
      }
    }
  }
}

// simulation_detect_collisions.cpp:250-397
inline void detect_primary_lef_lef_collisions(CollisionCtx& c, u64 n5, u64 n3) {
  const std::size_t n = c.n;
  if (n5 == n || n3 == n) return;
  std::size_t i1 = 0;
  std::size_t j1 = n5;
  const std::size_t i2 = n - sat_dec(n3);
  const std::size_t j2 = n;
  for (;;) {
    u64 rev_idx = c.rr[j1];
    u64 rev_pos = c.rev[rev_idx];
    u64 fwd_idx = c.fr[i1];
    u64 fwd_pos = c.fwd[fwd_idx];
    while (rev_pos <= fwd_pos) {
      if (++j1 == j2) return;
      rev_idx = c.rr[j1];
      rev_pos = c.rev[rev_idx];
    }
    while (fwd_pos < rev_pos) {
      if (++i1 == i2) return;
      fwd_idx = c.fr[i1];
      fwd_pos = c.fwd[fwd_idx];
    }
    fwd_idx = c.fr[sat_dec(i1)];
    fwd_pos = c.fwd[fwd_idx];
    const u64 delta = rev_pos - fwd_pos;
    if (delta > 0 && delta < c.rm[rev_idx] + c.fm[fwd_idx] && c.lef_lef_trial()) {
      const auto [cp_rev, cp_fwd] =
          lef_lef_collision_pos(rev_pos, fwd_pos, c.rm[rev_idx], c.fm[fwd_idx]);
      u32& rcol = c.rc[rev_idx];
      u32& fcol = c.fc[fwd_idx];
      const u32 hit_r = coll_make(fwd_idx, EV_COLLISION | EV_LEF_LEF_PRIMARY);
      const u32 hit_f = coll_make(rev_idx, EV_COLLISION | EV_LEF_LEF_PRIMARY);
      if (!coll_occurred(rcol) && !coll_occurred(fcol)) {
        rcol = hit_r;
        fcol = hit_f;
      } else if (coll_occurred(rcol) && !coll_occurred(fcol)) {
        const u64 barrier_pos = c.bars->pos[coll_index(rcol)];
        if (cp_fwd > barrier_pos) {
          rcol = hit_r;
          fcol = hit_f;
        } else {
          fcol = hit_f;
        }
      } else if (!coll_occurred(rcol) && coll_occurred(fcol)) {
        const u64 barrier_pos = c.bars->pos[coll_index(fcol)];
        rcol = hit_r;
        if (cp_rev < barrier_pos) fcol = hit_f;
      }
    }
  }
}

// simulation_correct_moves.cpp:19-50
inline void correct_moves_for_lef_bar_collisions(CollisionCtx& c) {
  for (std::size_t i = 0; i < c.n; ++i) {
    if (coll_occurred(c.rc[i], EV_LEF_BAR)) {
      c.rm[i] = (c.rev[i] - c.bars->pos[coll_index(c.rc[i])]) - 1;
    }
    if (coll_occurred(c.fc[i], EV_LEF_BAR)) {
      c.fm[i] = (c.bars->pos[coll_index(c.fc[i])] - c.fwd[i]) - 1;
    }
  }
}

// simulation_correct_moves.cpp:53-121
inline void correct_moves_for_primary_lef_lef_collisions(CollisionCtx& c) {
  for (std::size_t k = 0; k < c.n; ++k) {
    const u64 rev_idx = c.rr[k];
    if (coll_occurred(c.rc[rev_idx], EV_LEF_LEF_PRIMARY)) {
      const u64 fwd_idx = coll_index(c.rc[rev_idx]);
      if (coll_occurred(c.fc[fwd_idx], EV_LEF_LEF_PRIMARY)) {
        const auto [p1, p2] =
            lef_lef_collision_pos(c.rev[rev_idx], c.fwd[fwd_idx], c.rm[rev_idx], c.fm[fwd_idx]);
        c.rm[rev_idx] = c.rev[rev_idx] - p1;
        c.fm[fwd_idx] = p2 - c.fwd[fwd_idx];
      } else if (coll_occurred(c.fc[fwd_idx], EV_LEF_BAR)) {
        c.rm[rev_idx] = c.rev[rev_idx] - (c.fwd[fwd_idx] + c.fm[fwd_idx]) - 1;
      }
    }
  }
  for (std::size_t k = 0; k < c.n; ++k) {
    const u64 fwd_idx = c.fr[k];
    if (coll_occurred(c.fc[fwd_idx], EV_LEF_LEF_PRIMARY)) {
      const u64 rev_idx = coll_index(c.fc[fwd_idx]);
      if (coll_occurred(c.rc[rev_idx], EV_LEF_BAR)) {
        c.fm[fwd_idx] = (c.rev[rev_idx] - c.rm[rev_idx]) - c.fwd[fwd_idx] - 1;
      }
    }
  }
}

// simulation_detect_collisions.cpp:400-515
inline void process_secondary_lef_lef_collisions(CollisionCtx& c, u64 n5, u64 n3) {
  const std::size_t n = c.n;
  for (std::size_t i = std::max<u64>(1, n5); i < n; ++i) {
    const u64 idx1 = c.rr[i - 1];
    if (!coll_occurred(c.rc[idx1])) continue;
    const u64 idx2 = c.rr[i];
    if (coll_occurred(c.rc[idx2])) continue;
    const u64 pos1 = c.rev[idx1], pos2 = c.rev[idx2];
    u64& move1 = c.rm[idx1];
    u64& move2 = c.rm[idx2];
    if (pos2 - move2 <= pos1 - move1) {
      if (c.lef_lef_trial()) {
        c.rc[idx2] = coll_make(idx1, EV_COLLISION | EV_LEF_LEF_SECONDARY);
        const u64 move = pos2 - (pos1 - move1);
        move2 = sat_dec(move);
      } else {
        c.rc[idx2] = coll_make(idx1, EV_LEF_LEF_SECONDARY);
      }
    }
  }
  std::size_t i = n - sat_dec(n3) - 1;
  for (; i > 0; --i) {
    const u64 idx2 = c.fr[i];
    if (!coll_occurred(c.fc[idx2])) continue;
    const u64 idx1 = c.fr[i - 1];
    if (coll_occurred(c.fc[idx1])) continue;
    const u64 pos1 = c.fwd[idx1], pos2 = c.fwd[idx2];
    u64& move1 = c.fm[idx1];
    u64& move2 = c.fm[idx2];
    if (pos1 + move1 >= pos2 + move2) {
      if (c.lef_lef_trial()) {
        c.fc[idx1] = coll_make(idx2, EV_COLLISION | EV_LEF_LEF_SECONDARY);
        const u64 move = (pos2 + move2) - pos1;
        move1 = sat_dec(move);
      } else {
        c.fc[idx1] = coll_make(idx2, EV_LEF_LEF_SECONDARY);
      }
    }
  }
}

// simulation_detect_collisions.cpp:517-644
inline void fix_secondary_lef_lef_collisions(CollisionCtx& c, u64 n5, u64 n3) {
  const std::size_t n = c.n;
  const std::size_t num_active_fwd = n - sat_dec(n3);
  for (std::size_t i = std::max<u64>(1, n5); i < n; ++i) {
    const u64 idx2 = c.rr[i];
    if (coll_avoided(c.rc[idx2], EV_LEF_LEF_SECONDARY)) {
      const u64 idx1 = c.rr[i - 1];
      const u64 pos1 = c.rev[idx1] - c.rm[idx1];
      if (c.rev[idx2] > pos1 + 1) {
        c.rm[idx2] = c.rev[idx2] - (pos1 + 1);
      } else {
        c.rm[idx2] = 0;
      }
      c.rc[idx2] = coll_make(idx1, EV_COLLISION | EV_LEF_LEF_SECONDARY);
      const u64 p1 = c.rev[idx1], p2 = c.rev[idx2];
      c.rev[idx1] = std::min(c.fwd[idx1], p2);
      c.rev[idx2] = std::min(c.fwd[idx2], p1);
      std::swap(c.rc[idx1], c.rc[idx2]);
      std::swap(c.rm[idx1], c.rm[idx2]);
      std::swap(c.rr[i - 1], c.rr[i]);
      c.rm[c.rr[i - 1]] = std::min(c.rev[c.rr[i - 1]] - c.iv.start, c.rm[c.rr[i - 1]]);
      c.rm[c.rr[i]] = std::min(c.rev[c.rr[i]] - c.iv.start, c.rm[c.rr[i]]);
    }
  }
  for (std::size_t i = 0; i + 1 < num_active_fwd; ++i) {
    const u64 idx1 = c.fr[i];
    if (coll_avoided(c.fc[idx1], EV_LEF_LEF_SECONDARY)) {
      const u64 idx2 = c.fr[i + 1];
      const u64 pos2 = c.fwd[idx2] + c.fm[idx2];
      if (pos2 > c.fwd[idx1] + 1) {
        c.fm[idx1] = pos2 - (c.fwd[idx1] + 1);
      } else {
        c.fm[idx1] = 0;
      }
      c.fc[idx1] = coll_make(idx2, EV_COLLISION | EV_LEF_LEF_SECONDARY);
      const u64 p1 = c.fwd[idx1], p2 = c.fwd[idx2];
      c.fwd[idx1] = std::max(c.rev[idx1], p2);
      c.fwd[idx2] = std::max(c.rev[idx2], p1);
      std::swap(c.fc[idx1], c.fc[idx2]);
      std::swap(c.fm[idx1], c.fm[idx2]);
      std::swap(c.fr[i], c.fr[i + 1]);
      c.fm[c.fr[i]] = std::min(c.iv.end - 1 - c.fwd[c.fr[i]], c.fm[c.fr[i]]);
      c.fm[c.fr[i + 1]] = std::min(c.iv.end - 1 - c.fwd[c.fr[i + 1]], c.fm[c.fr[i + 1]]);
    }
  }
}

// Simulation::process_collisions (simulation.cpp:763-793); `with_fix` = false reproduces the test
// shim test_process_collisions (simulation.hpp:499-528), which omits fix_secondary.
inline std::pair<u64, u64> process_collisions(CollisionCtx& c, bool with_fix = true) {
  const auto [n5, n3] = detect_units_at_interval_boundaries(c);
  detect_lef_bar_collisions(c, n5, n3);
  detect_primary_lef_lef_collisions(c, n5, n3);
  correct_moves_for_lef_bar_collisions(c);
  correct_moves_for_primary_lef_lef_collisions(c);
  process_secondary_lef_lef_collisions(c, n5, n3);
  if (with_fix) fix_secondary_lef_lef_collisions(c, n5, n3);
  return {n5, n3};
}

// ------------------------------------------------------------------------------------------
// Contact sink: banded u32 matrix + missed updates + 1D occupancy
// (contact_matrix_dense_safe_impl.hpp:54-68; register_contacts.cpp:199-232)
// ------------------------------------------------------------------------------------------
// Measurement aid for ONE question (tests/test_burnin_margin.py): how close do the window-mean
// comparisons of evaluate_burnin come to a tie? The CUDA kernel sums the squared deviations behind
// the coefficient of variation in a tree order, the reference (and this oracle) left to right, so
// a comparison can only come out differently when the two means are within a few ulp of each
// other. Counters are process-wide relaxed atomics: comparisons seen, comparisons within 64 ulp,
// and the smallest relative gap (as the bit pattern of a double).
struct BurninMargin {
  u64 comparisons = 0, within_64_ulp = 0, min_gap_bits = 0x7FF0000000000000ull;
};
inline BurninMargin& burnin_margin() {
  static BurninMargin m;
  return m;
}
inline void burnin_margin_note(double n1, double n2) {
  BurninMargin& m = burnin_margin();
  __atomic_fetch_add(&m.comparisons, u64(1), __ATOMIC_RELAXED);
  const double big = std::fabs(n1) > std::fabs(n2) ? std::fabs(n1) : std::fabs(n2);
  if (!(big > 0.0)) return;
  const double gap = std::fabs(n1 - n2) / big;
  if (gap <= 64 * 2.220446049250313e-16) __atomic_fetch_add(&m.within_64_ulp, u64(1), __ATOMIC_RELAXED);
  u64 bits;
  std::memcpy(&bits, &gap, 8);
  u64 cur = __atomic_load_n(&m.min_gap_bits, __ATOMIC_RELAXED);
  while (bits < cur &&
         !__atomic_compare_exchange_n(&m.min_gap_bits, &cur, bits, true, __ATOMIC_RELAXED,
                                      __ATOMIC_RELAXED)) {
  }
}

struct ContactSink {
  u32* band = nullptr;  // nrows*ncols (+1 slack in the reference layout)
  u64* occ1d = nullptr;
  u64* missed = nullptr;
  Band geom;
  void increment(u64 b1, u64 b2) const {
    const u64 i = b1 > b2 ? b1 - b2 : b2 - b1;
    const u64 j = b1 > b2 ? b1 : b2;
    if (i >= geom.nrows) {
      __atomic_fetch_add(missed, u64(1), __ATOMIC_RELAXED);
      return;
    }
    __atomic_fetch_add(&band[j * geom.nrows + i], u32(1), __ATOMIC_RELAXED);
  }
  void occupancy(u64 b) const {
    if (occ1d) __atomic_fetch_add(&occ1d[b], u64(1), __ATOMIC_RELAXED);
  }
};

struct CellTask {
  u64 cell_id = 0;
  u64 target_contacts = 0;
  u64 target_epochs = 0;
  u64 rng_state[4] = {0, 0, 0, 0};
};

struct CellResult {
  u64 num_contacts = 0, epochs = 0, burnin_epochs = 0, lef_updates = 0, rng_draws = 0;
};

// register_contacts.cpp:23-63
inline bool lef_within_bound(u64 rev, u64 fwd, u64 s, u64 e) noexcept {
  return rev > s && rev < e && fwd > s && fwd < e;
}
inline bool pos_within_bound(double p1, double p2, u64 s, u64 e) noexcept {
  const double sd = static_cast<double>(s), ed = static_cast<double>(e);
  return p1 >= sd && p2 >= sd && p1 < ed && p2 < ed;
}

class CellSim {
 public:
  CellSim(const Params& p, const Interval& iv, const Barriers& shared_bars, std::size_t nlefs,
          const CellTask& task, const ContactSink& sink)
      : p_(p), iv_(iv), bars_(shared_bars), task_(task), sink_(sink) {
    s_.reset(nlefs);
    g_ = Rng::from_state(task.rng_state);
  }

  // internal-state log (Config::log_model_internal_state): records [0, log_cap)
  void set_log(modle_b200_epoch_record* log, std::size_t cap) {
    log_ = log;
    log_cap_ = cap;
  }

  // Simulation::dump_stats (simulation.cpp:995-1056): same quantities, same point of the epoch
  void dump_stats() {
    const CellState& s = s_;
    if (!log_ || s.epoch >= log_cap_) return;
    modle_b200_epoch_record r{};
    r.epoch = s.epoch;
    r.burnin = s.burnin_completed ? 0u : 1u;
    r.num_lefs = static_cast<u32>(s.num_active);
    for (std::size_t i = 0; i < bars_.size(); ++i) r.barriers_occupied += bars_.active[i] != 0;
    double loop_sum = 0.0;  // stats::mean accumulates in double (descriptive_impl.hpp:22-31)
    for (std::size_t i = 0; i < s.num_active; ++i) {
      const bool rv = coll_occurred(s.rev_coll[i]), fw = coll_occurred(s.fwd_coll[i]);
      r.lefs_stalled_rev += rv;
      r.lefs_stalled_fwd += fw;
      r.lefs_stalled_both += rv && fw;
      r.lef_bar_collisions +=
          coll_occurred(s.rev_coll[i], EV_LEF_BAR) + coll_occurred(s.fwd_coll[i], EV_LEF_BAR);
      r.lef_lef_primary_collisions += coll_occurred(s.rev_coll[i], EV_LEF_LEF_PRIMARY) +
                                      coll_occurred(s.fwd_coll[i], EV_LEF_LEF_PRIMARY);
      r.lef_lef_secondary_collisions += coll_occurred(s.rev_coll[i], EV_LEF_LEF_SECONDARY) +
                                        coll_occurred(s.fwd_coll[i], EV_LEF_LEF_SECONDARY);
      loop_sum += static_cast<double>(s.fwd[i] - s.rev[i]);  // Lef::loop_size
    }
    r.loop_size_sum = static_cast<u64>(loop_sum);
    log_[s.epoch] = r;
  }

  CellState& state() { return s_; }
  Rng& rng() { return g_; }
  Barriers& barriers() { return bars_; }

  // Simulation::simulate_one_cell (simulation.cpp:896-986)
  CellResult run() {
    CellState& s = s_;
    const double binding_rate =
        static_cast<double>(s.nlefs) / static_cast<double>(p_.burnin_activation_epochs);
    // ExtrusionBarriers::init_states (extrusion_barriers.cpp:219-230)
    for (std::size_t i = 0; i < bars_.size(); ++i) {
      bars_.active[i] =
          bernoulli(g_, occupancy_from_stp(bars_.stp_active[i], bars_.stp_inactive[i])) ? 1 : 0;
    }
    if (p_.skip_burnin) {
      s.num_active = s.nlefs;
      s.burnin_completed = true;
    }
    auto stop = [&]() {
      if (!p_.stop_on_epochs) return s.num_contacts >= task_.target_contacts;
      return s.epoch - s.num_burnin_epochs >= task_.target_epochs;
    };
    for (; !stop() && s.epoch < p_.debug_max_epochs; ++s.epoch) {
      if (!s.burnin_completed) run_burnin(binding_rate);
      bind_lefs();
      if (s.burnin_completed) {
        sample_and_register_contacts();
        if (task_.target_contacts != 0 && s.num_contacts >= task_.target_contacts) break;
      }
      s.lef_updates += s.num_active;
      generate_moves();
      next_barrier_states();
      std::fill(s.rev_coll.begin(), s.rev_coll.begin() + s.num_active, 0u);
      std::fill(s.fwd_coll.begin(), s.fwd_coll.begin() + s.num_active, 0u);
      CollisionCtx c = ctx();
      process_collisions(c, true);
      extrude();
      dump_stats();
      release_lefs();
    }
    CellResult r;
    r.num_contacts = s.num_contacts;
    r.epochs = s.epoch;
    r.burnin_epochs = s.num_burnin_epochs;
    r.lef_updates = s.lef_updates;
    r.rng_draws = g_.ndraws;
    return r;
  }

  CollisionCtx ctx() {
    CollisionCtx c;
    c.p = &p_;
    c.iv = iv_;
    c.bars = &bars_;
    c.rev = s_.rev.data();
    c.fwd = s_.fwd.data();
    c.ep = s_.epoch_bound.data();
    c.rr = s_.rev_rank.data();
    c.fr = s_.fwd_rank.data();
    c.rm = s_.rev_move.data();
    c.fm = s_.fwd_move.data();
    c.rc = s_.rev_coll.data();
    c.fc = s_.fwd_coll.data();
    c.n = s_.num_active;
    c.g = &g_;
    return c;
  }

  // run_burnin (simulation.cpp:866-894), compute_loop_size_stats (:795-819),
  // evaluate_burnin (:821-864)
  void run_burnin(double binding_rate) {
    CellState& s = s_;
    do {
      ++s.num_burnin_epochs;
      if (s.num_active != s.nlefs) {
        const u64 k = poisson(g_, binding_rate);
        s.num_active = std::min<u64>(s.num_active + k, s.nlefs);
      } else {
        loop_size_stats();
        s.burnin_completed = evaluate_burnin();
        s.burnin_completed = s.burnin_completed && (s.epoch > p_.min_burnin_epochs);
        if (!s.burnin_completed && s.epoch >= p_.max_burnin_epochs) {
          s.burnin_completed = true;
          s.num_active = s.nlefs;
        }
      }
    } while (s.num_active == 0);
  }

  // stats::mean and stats::standard_dev over the loop sizes (src/stats/descriptive_impl.hpp:
  // 22-32, 64-101): left-to-right double accumulation, population variance (divides by n)
  template <class LoopFn>
  static void loop_size_mean_sd(std::size_t n, LoopFn loop, double* mean_out, double* sd_out) {
    double acc = 0.0;
    for (std::size_t i = 0; i < n; ++i) acc = acc + static_cast<double>(loop(i));
    const double mean = acc / static_cast<double>(n);
    double ssd = 0.0;
    for (std::size_t i = 0; i < n; ++i) {
      const double d = static_cast<double>(loop(i)) - mean;
      ssd = ssd + (d * d);
    }
    *mean_out = mean;
    *sd_out = std::sqrt(ssd / static_cast<double>(n));
  }

  void loop_size_stats() {
    CellState& s = s_;
    const std::size_t n = s.num_active;
    if (n == 0) {
      s.cv_buff.clear();
      s.avg_buff.clear();
      return;
    }
    double mean = 0.0, sd = 0.0;
    loop_size_mean_sd(n, [&](std::size_t i) -> u64 { return s.bound(i) ? s.fwd[i] - s.rev[i] : 0; },
                      &mean, &sd);
    if (s.avg_buff.size() == p_.burnin_history) {
      s.avg_buff.pop_front();
      s.cv_buff.pop_front();
    }
    s.avg_buff.push_back(mean);
    s.cv_buff.push_back(sd / s.avg_buff.back());
  }

  bool evaluate_burnin() const {
    const CellState& s = s_;
    const std::size_t cap = p_.burnin_history, w = p_.burnin_window;
    if (s.cv_buff.size() != cap) return false;
    auto stable = [&](const std::deque<double>& b) {
      std::size_t n = 0;
      // windows [j-1, j-1+w) and [j, j+w) for every j with j+w < cap (simulation.cpp:838-844:
      // the loop ends when the exclusive end of the second window reaches end())
      for (std::size_t j = 1; j + w < cap; ++j) {
        double a1 = 0.0, a2 = 0.0;
        for (std::size_t k = 0; k < w; ++k) a1 = a1 + b[j - 1 + k];
        for (std::size_t k = 0; k < w; ++k) a2 = a2 + b[j + k];
        const double n1 = a1 / static_cast<double>(w);
        const double n2 = a2 / static_cast<double>(w);
        n += static_cast<std::size_t>(n1 > n2);
        if (&b == &s.cv_buff) burnin_margin_note(n1, n2);
      }
      const double r = static_cast<double>(n) / static_cast<double>(cap - w - n);
      return r >= 0.95 && r <= 1.05;
    };
    if (!stable(s.cv_buff)) return false;
    return stable(s.avg_buff);
  }

  // select_lefs_to_bind + bind_lefs (simulation_impl.hpp:30-91) + rank_lefs
  void bind_lefs() {
    CellState& s = s_;
    const std::size_t n = s.num_active;
    for (std::size_t i = 0; i < n; ++i) s.mask[i] = !s.bound(i);
    for (std::size_t i = 0; i < n; ++i) {
      if (s.mask[i]) {
        const u64 pos = uniform_int(g_, iv_.start, iv_.end - 1);
        s.rev[i] = s.fwd[i] = pos;
        s.epoch_bound[i] = s.epoch;
      }
    }
    rank_lefs(s.rev.data(), s.fwd.data(), s.epoch_bound.data(), s.rev_rank.data(),
              s.fwd_rank.data(), n);
  }

  // register_contacts.cpp:47-63
  std::pair<double, double> noisy_positions(std::size_t i) {
    auto noise = [&]() {
      return p_.noisify ? genextreme(g_, p_.gev_mu, p_.gev_sigma, p_.gev_xi) : 0.0;
    };
    const double a = static_cast<double>(s_.rev[i]) - noise();
    const double b = static_cast<double>(s_.fwd[i]) + noise();
    return {std::min(a, b), std::max(a, b)};
  }

  // register_contacts.cpp:93-232
  void sample_and_register_contacts() {
    CellState& s = s_;
    u64 nev = p_.contacts_per_epoch;
    if (!p_.stop_on_epochs) nev = std::min(nev, task_.target_contacts - s.num_contacts);
    if (nev == 0) return;
    u64 nloop;
    if (p_.tad_to_loop_ratio == 0) {
      nloop = nev;
    } else if (!std::isfinite(p_.tad_to_loop_ratio)) {
      nloop = 0;
    } else {
      nloop = static_cast<u64>(
          binomial(g_, static_cast<i64>(nev), 1.0 / (p_.tad_to_loop_ratio + 1.0)));
    }
    const u64 ntad = nev - nloop;
    const std::size_t n = s.num_active;
    const u64 sp = iv_.start + 1, ep = iv_.end - 1;
    for (u64 k = nloop; k != 0; --k) {
      const std::size_t i = uniform_int(g_, 0, n - 1);
      if (s.bound(i) && lef_within_bound(s.rev[i], s.fwd[i], sp, ep)) {
        const auto [p1, p2] = noisy_positions(i);
        if (!pos_within_bound(p1, p2, sp, ep)) continue;
        const u64 pos1 = static_cast<u64>(p1) - sp, pos2 = static_cast<u64>(p2) - sp;
        sink_.increment(pos1 / p_.bin_size, pos2 / p_.bin_size);
        ++s.num_contacts;
      }
    }
    for (u64 k = ntad; k != 0; --k) {
      const std::size_t i = uniform_int(g_, 0, n - 1);
      if (s.bound(i) && lef_within_bound(s.rev[i], s.fwd[i], sp, ep)) {
        const auto [p1, p2] = noisy_positions(i);
        if (!pos_within_bound(p1, p2, sp, ep)) continue;
        const u64 p11 = uniform_int(g_, static_cast<u64>(p1), static_cast<u64>(p2));
        const u64 p22 = uniform_int(g_, static_cast<u64>(p1), static_cast<u64>(p2));
        sink_.increment((p11 - sp) / p_.bin_size, (p22 - sp) / p_.bin_size);
        ++s.num_contacts;
      }
    }
    if (p_.track_1d) {
      for (u64 k = nev; k != 0; --k) {
        const std::size_t i = uniform_int(g_, 0, n - 1);
        if (s.bound(i) && lef_within_bound(s.rev[i], s.fwd[i], sp, ep)) {
          const auto [p1, p2] = noisy_positions(i);
          if (!pos_within_bound(p1, p2, sp, ep)) continue;
          sink_.occupancy((static_cast<u64>(p1) - sp) / p_.bin_size);
          sink_.occupancy((static_cast<u64>(p2) - sp) / p_.bin_size);
        }
      }
    }
  }

  // generate_moves (simulation.cpp:272-330)
  void generate_moves() {
    CellState& s = s_;
    const std::size_t n = s.num_active;
    const double rs = s.burnin_completed ? p_.rev_speed : p_.rev_speed_burnin;
    const double fs = s.burnin_completed ? p_.fwd_speed : p_.fwd_speed_burnin;
    auto gen = [&](u64* out, double speed, double sd) {
      const u64 mi = static_cast<u64>(std::round(speed));
      for (std::size_t i = 0; i < n; ++i) {
        if (!s.bound(i)) {
          out[i] = 0;
        } else if (sd == 0.0) {
          out[i] = mi;
        } else {
          out[i] = static_cast<u64>(std::round(std::max(0.0, normal(g_, speed, sd))));
        }
      }
    };
    gen(s.rev_move.data(), rs, p_.rev_speed_std);
    gen(s.fwd_move.data(), fs, p_.fwd_speed_std);
    adjust_moves(iv_, s.rev.data(), s.fwd.data(), s.epoch_bound.data(), s.rev_rank.data(),
                 s.fwd_rank.data(), s.rev_move.data(), s.fwd_move.data(), n);
    clamp_moves(iv_, s.rev.data(), s.fwd.data(), s.epoch_bound.data(), s.rev_move.data(),
                s.fwd_move.data(), n);
  }

  // ExtrusionBarriers::next_state (extrusion_barriers.cpp:145-161)
  void next_barrier_states() {
    for (std::size_t i = 0; i < bars_.size(); ++i) {
      const double u = canonical(g_);
      if (!bars_.active[i] && u > bars_.stp_inactive[i]) {
        bars_.active[i] = 1;
      } else if (bars_.active[i] && u > bars_.stp_active[i]) {
        bars_.active[i] = 0;
      }
    }
  }

  // extrude (simulation.cpp:498-521)
  void extrude() {
    CellState& s = s_;
    for (std::size_t i = 0; i < s.num_active; ++i) {
      if (!s.bound(i)) continue;
      s.rev[i] -= s.rev_move[i];
      s.fwd[i] += s.fwd_move[i];
    }
  }

  // release_lefs (simulation.cpp:553-601)
  void release_lefs() {
    CellState& s = s_;
    const double base = s.burnin_completed ? p_.prob_lef_release : p_.prob_lef_release_burnin;
    for (std::size_t i = 0; i < s.num_active; ++i) {
      if (!s.bound(i)) continue;
      int hard = 0;
      if (coll_occurred(s.rev_coll[i], EV_LEF_BAR) &&
          bars_.dir[coll_index(s.rev_coll[i])] == DIR_REV)
        ++hard;
      if (coll_occurred(s.fwd_coll[i], EV_LEF_BAR) &&
          bars_.dir[coll_index(s.fwd_coll[i])] == DIR_FWD)
        ++hard;
      const double affinity = hard == 0   ? 1.0
                              : hard == 1 ? 1.0 / p_.soft_stall_multiplier
                                          : 1.0 / p_.hard_stall_multiplier;
      if (bernoulli(g_, affinity * base)) {
        s.rev[i] = s.fwd[i] = s.epoch_bound[i] = UNBOUND;
      }
    }
  }

 private:
  Params p_;
  Interval iv_;
  Barriers bars_;
  CellTask task_;
  ContactSink sink_;
  modle_b200_epoch_record* log_ = nullptr;
  std::size_t log_cap_ = 0;
  CellState s_;
  Rng g_;
};

}  // namespace oracle
