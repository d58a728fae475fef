// Plain-data types shared by the host launcher, the CUDA kernel and the CPU emulation of the
// kernel (tests/emu). Positions are held as u32 on the device (largest GRCh38 chromosome is
// 248,956,422 bp); the all-ones value is the reference's "unbound" sentinel
// (src/libmodle/internal/extrusion_factors_impl.hpp:96-98,120-124).
#pragma once
#include "../../include/modle_b200.h"
#include "cta.hpp"

namespace modle_b200 {

constexpr u32 kUnbound = 0xFFFFFFFFu;

// Collision word: event flags of Collision<> (src/libmodle/cpu/include/modle/
// collision_encoding.hpp:91-96) in bits 24..30, index in bits 0..23. Bit 31 is a device-only
// marker used while LEF-BAR hits are being collected.
enum : u32 {
  kEvCollision = 0x10,
  kEvChromBoundary = 0x08,
  kEvLefBar = 0x04,
  kEvPrimary = 0x02,
  kEvSecondary = 0x01,
  kEvTmp = 0x80,
};
MB_FN u32 coll_make(u32 idx, u32 ev) { return (ev << 24) | idx; }
MB_FN u32 coll_event(u32 c) { return c >> 24; }
MB_FN u32 coll_index(u32 c) { return c & 0x00FFFFFFu; }
MB_FN bool coll_occurred(u32 c) { return ((c >> 24) & kEvCollision) != 0; }
MB_FN bool coll_is(u32 c, u32 kind) { return coll_event(c) == (kind | kEvCollision); }
MB_FN bool coll_avoided(u32 c, u32 kind) { return !coll_occurred(c) && coll_event(c) == kind; }

// Kernel-side fault codes (reported through modle_b200_cell_stats::device_fault).
enum : u32 {
  kFaultNone = 0,
  kFaultRngWindow = 1,      // a phase needed more raw draws than the staging window holds
  kFaultUnboundLef = 2,     // an active LEF was unbound inside the epoch body
  kFaultBurninHistory = 3,  // burn-in history longer than the shared-memory ring
  kFaultSerialDraws = 4,    // a serial sampler ran out of its draw budget
  kFaultMoveRange = 5,      // a move of 2^24 bp or more in one epoch (collision words hold 24 bits)
};

constexpr int kMaxBurninHistory = 256;

// Phases of the per-cell loop whose SM-clock cycles the kernel accumulates (cheap: two clock
// reads by one thread per phase). Read back with modle_b200_phase_cycles().
enum : int {
  kPhInit = 0,
  kPhBurnin,
  kPhBind,
  kPhRank,
  kPhContacts,
  kPhMovesGen,
  kPhMovesAdjust,
  kPhBarriers,
  kPhLefBar,
  kPhPrimary,
  kPhCorrect,
  kPhSecondary,
  kPhFix,
  kPhExtrudeRelease,
  kPhRngGenerate,  // nested inside the phases above (time spent refilling the RNG ring)
  kPhTotal,
  // finer split of the two heaviest phases (nested; profiling builds read them the same way)
  kPhMvEnsure,
  kPhMvScan,
  kPhMvExceptions,
  kPhMvFinal,
  kPhSecCompose,
  kPhSecScan,
  kPhSecClassify,
  kPhSecDraws,
  kPhSecLeader,
  kPhSecApply,
  kNumPhases
};

// number of RNG staging configurations (jump tables) a context keeps; see sim_core.hpp
constexpr int kJumpSlots = 3;

// Everything the kernel needs to know about the run; one per launch, in global memory.
struct KernelParams {
  u32 start, end;  // interval [start, end)
  u32 n_lefs, n_bar;
  u32 bin_size;
  u32 nrows, ncols;  // band geometry
  double rev_speed, fwd_speed, rev_speed_burnin, fwd_speed_burnin;
  double rev_std, fwd_std;
  double p_release, p_release_burnin;
  double hard_mult, soft_mult;
  double p_bypass;
  double pblock_major, pblock_minor;
  double tad_to_loop;
  double gev_mu, gev_sigma, gev_xi;
  u32 noisify, track_1d, skip_burnin, stop_on_epochs;
  u32 contacts_per_epoch;
  u32 burnin_history, burnin_window;
  u64 min_burnin_epochs, max_burnin_epochs;
  double lef_binding_rate_burnin;  // n_lefs / burnin_target_epochs_for_lef_activation
  u64 debug_max_epochs;
  u32 move_bound;  // every generated move is <= this unless CellShared::move_bound_hit says so
  // RNG staging configuration
  u32 rng_gen_threads;  // G: threads that own a generator sub-stream
  u32 rng_per_thread;   // l: consecutive draws per generator thread per window
  u32 rng_window;       // W = G * l
  u32 rng_jump_slot;    // which precomputed T^W matrix to use
  // barrier look-up table of the LEF-BAR walk: bar_lut[(pos - start) >> lut_shift] = number of
  // barriers below the bucket's first position; lut_entries == 0: no table (not enough room)
  u32 lut_shift, lut_entries;
};

// Per-interval read-only device arrays.
struct IntervalData {
  const u32* bar_pos;          // n_bar, sorted
  const u32* bar_dir_rev;      // bitmask, bit set = barrier blocks REV-moving units
  const double* bar_stp_active;
  const double* bar_stp_inactive;
  const double* bar_occupancy;  // Bernoulli parameter of init_states
  const double* zig_nx;         // 129: ziggurat x table of the unit normal
  const double* zig_ny;         // 129
  const double* zig_ex;         // 257: unit exponential
  const double* zig_ey;         // 257
  const u64* jump_tbl;          // byte-indexed T^W table of the staging configuration (sim_core)
};

struct CellTaskDev {
  u64 cell_id, target_epochs, target_contacts;
  u64 rng_state[4];
};

struct CellStatsDev {
  u64 num_contacts, num_epochs, num_burnin_epochs, num_lef_updates, num_rng_draws, device_fault;
};

// Output sinks (global memory).
struct Sinks {
  u32* band;    // nrows*ncols+1
  u64* occ1d;   // ncols or null
  u64* missed;  // 1
  modle_b200_epoch_record* log;  // this cell's internal-state log (log_cap records) or null
  u32 log_cap;
};

// Small per-cell state that lives in shared memory next to the arrays.
struct CellShared {
  u64 epoch;
  u64 num_burnin_epochs;
  u64 num_contacts;
  u64 lef_updates;
  u64 rng_pos;        // next raw draw to consume (stream offset)
  u64 rng_generated;  // ring holds offsets [rng_generated - 2W, rng_generated) at most
  u32 num_active;
  u32 burnin_completed;
  u32 fault;
  u32 n5, n3;
  u32 hist_len, hist_head;  // burn-in history ring
  u32 done;
  u32 move_bound_hit;  // a generated move exceeded KernelParams::move_bound
  u32 tmp_u32[8];
  u64 tmp_u64[4];
  u64 phase_cycles[kNumPhases];  // SM clock cycles spent per phase of the epoch loop (thread 0)
  double avg_hist[kMaxBurninHistory];
  double cv_hist[kMaxBurninHistory];
  CtaScratch scratch;
};

// Pointers to the per-cell arrays (shared memory on the device, heap in the emulation).
struct CellArrays {
  u32 *rev, *fwd, *ep;  // n_lefs each
  u16 *rr, *fr;         // rank permutations
  u32 *rm, *fm;         // moves
  u32 *rc, *fc;         // collision words
  u32* scratch;         // max(n_lefs, n_bar) + 64 words
  u32* bits;            // 12 * (n_lefs/32 + 3) words: bitmaps of the secondary-collision pass
  u32* bar_pos;         // n_bar (copy of IntervalData::bar_pos)
  u8* bar_active;       // n_bar bytes (0/1)
  u16* bar_lut;         // KernelParams::lut_entries entries, or null
  double* zig_nx;       // 129 (copy)
  // per-CTA global scratch
  u64* rng_ring;   // 2 * rng_window entries
  u64* rng_state;  // 4 * rng_gen_threads entries, SoA
};

// Layout of the per-cell arrays inside one contiguous, 8-byte aligned buffer.
MB_HD size_t cell_scratch_words(u32 n_lefs, u32 n_bar) {
  return size_t(n_lefs > n_bar ? n_lefs : n_bar) + 64;
}
MB_HD size_t cell_bits_words(u32 n_lefs) { return size_t(12) * (n_lefs / 32 + 3); }
MB_HD size_t cell_array_bytes(u32 n_lefs, u32 n_bar, u32 lut_entries = 0) {
  size_t w = 0;
  w += 260;                                // zig_nx: 129 doubles (+ pad)
  w += size_t(7) * n_lefs;                 // rev, fwd, ep, rm, fm, rc, fc
  w += cell_scratch_words(n_lefs, n_bar);  // scratch
  w += cell_bits_words(n_lefs);            // bits
  w += n_bar;                              // bar_pos
  w += n_lefs + 2;                         // rr, fr (u16 each)
  w += (n_bar + 3) / 4 + 1;                // bar_active (bytes)
  w += (lut_entries + 1) / 2;              // bar_lut (u16 each)
  return ((w * 4 + 15) / 16) * 16;
}
MB_HD CellArrays carve_cell_arrays(void* base, u32 n_lefs, u32 n_bar, u32 lut_entries = 0) {
  CellArrays a;
  u32* p = static_cast<u32*>(base);
  a.zig_nx = reinterpret_cast<double*>(p);
  p += 260;
  a.rev = p;
  p += n_lefs;
  a.fwd = p;
  p += n_lefs;
  a.ep = p;
  p += n_lefs;
  a.rm = p;
  p += n_lefs;
  a.fm = p;
  p += n_lefs;
  a.rc = p;
  p += n_lefs;
  a.fc = p;
  p += n_lefs;
  a.scratch = p;
  p += cell_scratch_words(n_lefs, n_bar);
  a.bits = p;
  p += cell_bits_words(n_lefs);
  a.bar_pos = p;
  p += n_bar;
  a.rr = reinterpret_cast<u16*>(p);
  a.fr = a.rr + n_lefs + (n_lefs & 1);
  p += n_lefs + 2;
  a.bar_active = reinterpret_cast<u8*>(p);
  p += (n_bar + 3) / 4 + 1;
  a.bar_lut = lut_entries ? reinterpret_cast<u16*>(p) : nullptr;
  a.rng_ring = nullptr;
  a.rng_state = nullptr;
  return a;
}

}  // namespace modle_b200
