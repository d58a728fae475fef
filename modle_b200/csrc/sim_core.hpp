// Per-cell loop-extrusion simulation, one CTA per cell, state resident in shared memory.
//
// Data-parallel restatement of Simulation::simulate_one_cell
// (reference: src/libmodle/cpu/simulation.cpp:896-986) and of everything it calls, preserving the
// reference's sequential RNG draw order (one xoshiro256++ stream per cell). Written in the
// bulk-synchronous style of cta.hpp so that the CUDA kernel and the CPU emulation used by the
// tests compile from this one source. Each phase cites the reference routine it replaces.
#pragma once
#include "sim_types.hpp"

namespace modle_b200 {

// ---- jump-ahead: g <- T^W g through byte-indexed tables ------------------------------------------
// tbl[(k * 256 + v) * 4 + w] = word w of T^W applied to the state whose k-th byte is v (all other
// bytes 0); T^W g is the XOR of the 32 entries selected by the bytes of g. One table (256 KB) per
// staging configuration, built on the host (launch_prep.hpp) and kept in global memory.

constexpr size_t kJumpTableWords = size_t(32) * 256 * 4;
#if MB_DEVICE_BUILD
// The RNG ring lives in global memory and is written and read by the same CTA with a CTA
// barrier in between, so plain (L1-cached) loads are coherent; MODLE_B200_RING_LDCG forces the
// L2-only path for comparison.
#ifdef MODLE_B200_RING_LDCG
#define MB_LD_RING_U64(p) __ldcg(p)
#else
#define MB_LD_RING_U64(p) (*(p))
#endif
#define MB_ATOMIC_MIN_U32(ptr, val) atomicMin((ptr), (val))
#else
#define MB_LD_RING_U64(p) (*(p))
#ifdef MODLE_B200_EMU_MT
#define MB_ATOMIC_MIN_U32(ptr, val) emu_atomic_min_u32((ptr), (val))
#else
#define MB_ATOMIC_MIN_U32(ptr, val) (*(ptr) = std::min<u32>(*(ptr), (val)))
#endif
#endif

// Profiling builds (-DMODLE_B200_PROBE=1, throughput mode only: it borrows the nested phase slots
// that mode leaves unused) split the rank and LEF-BAR phases further; see scripts/gpu_probe.sh.
#ifdef MODLE_B200_PROBE
#define MB_PROBE_BEGIN() sub_begin()
#define MB_PROBE(slot) sub_lap(slot)
#else
#define MB_PROBE_BEGIN() ((void)0)
#define MB_PROBE(slot) ((void)0)
#endif

constexpr double kTwo64 = 18446744073709551616.0;
#if !MB_DEVICE_BUILD
using std::isfinite;
#endif

struct Xs {
  u64 s0, s1, s2, s3;
};
MB_FN u64 rotl64(u64 x, int k) { return (x << k) | (x >> (64 - k)); }
MB_FN u64 xs_next(Xs& g) {
  const u64 r = rotl64(g.s0 + g.s3, 23) + g.s0;
  const u64 t = g.s1 << 17;
  g.s2 ^= g.s0;
  g.s3 ^= g.s1;
  g.s1 ^= g.s2;
  g.s0 ^= g.s3;
  g.s2 ^= t;
  g.s3 = rotl64(g.s3, 45);
  return r;
}
MB_FN Xs xs_jump(const Xs& g, const u64* tbl) {
  Xs r{0, 0, 0, 0};
  const u64 w[4] = {g.s0, g.s1, g.s2, g.s3};
#pragma unroll
  for (int wi = 0; wi < 4; ++wi) {
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const u32 v = static_cast<u32>(w[wi] >> (8 * b)) & 0xFFu;
      const u64* e = tbl + (size_t(wi * 8 + b) * 256 + v) * 4;
#if MB_DEVICE_BUILD
      const ulonglong2 x = __ldg(reinterpret_cast<const ulonglong2*>(e));
      const ulonglong2 y = __ldg(reinterpret_cast<const ulonglong2*>(e) + 1);
      r.s0 ^= x.x;
      r.s1 ^= x.y;
      r.s2 ^= y.x;
      r.s3 ^= y.y;
#else
      r.s0 ^= e[0];
      r.s1 ^= e[1];
      r.s2 ^= e[2];
      r.s3 ^= e[3];
#endif
    }
  }
  return r;
}

// ---- elementary Boost.Random pieces on a raw 64-bit draw (see oracle/oracle_rng.hpp notes) ------
MB_FN bool bernoulli_raw(u64 u, double p) { return MB_U64_TO_F64(u) <= p * kTwo64; }
MB_FN double canonical_raw(u64 u) {
  double r = MB_U64_TO_F64(u) / kTwo64;
  if (r == 1.0) r -= 2.220446049250313e-16 / 2;
  return r;
}
MB_FN u64 uniform_int_bucket(u64 range) {
  const u64 brange = ~u64(0);
  u64 bucket = brange / (range + 1);
  if (brange % (range + 1) == range) ++bucket;
  return bucket;
}
// floor(x / d) for a divisor that is used many times: m = floor((2^64 - 1) / d) never exceeds
// 2^64 / d, so the high half of x * m is at most x / d and short of it by two at the very most;
// the remainder settles the rest exactly. (A 64-bit division is ~100 instructions on the device.)
struct InvU64 {
  u64 d, m;
};
MB_FN InvU64 inv_u64(u64 d) { return InvU64{d, ~u64(0) / d}; }
MB_FN u64 div_u64(u64 x, const InvU64& v) {
  u64 q = MB_UMULHI64(x, v.m);
  u64 r = x - q * v.d;
  while (r >= v.d) {
    r -= v.d;
    ++q;
  }
  return q;
}
// detail::generate_int_float_pair<double, 8>
MB_FN double int_float_pair8(u64 u, int* bucket) {
  *bucket = static_cast<int>(u & 0xFF);
  u &= ~((u64(1) << 11) - 1);
  return MB_U64_TO_F64(u >> 8) * (1.0 / 72057594037927936.0);
}
// genextreme_value_distribution (common/genextreme_value_distribution.hpp:87-105) on a canonical u
MB_FN double gev_from_canonical(double u, double mu, double sigma, double xi) {
  if (xi == 0.0) return (mu - sigma) * log(-log(u));
  return mu + (sigma * (1.0 - pow(-log(u), xi))) / xi;
}

// ---- throughput mode: counter-based draws ---------------------------------------------------
// The reference consumes ONE sequential stream per cell, which is what the default (deterministic)
// mode reproduces draw for draw. The throughput mode gives up that order, not the distributions:
// every (epoch, phase, item) owns a private sequence of 256 raw draws addressed by a 64-bit
// counter, draw = mix64((counter ^ key1) * golden + key2) -- the SplitMix64 output function over
// a per-cell keyed counter -- so no staging ring, no offset scans and no speculation/repair are
// needed, and a cell's result still depends on nothing but its task (not on the CTA width, the
// grid or the GPU count). Results are statistically equivalent to the reference's, not
// bit-identical (tests/test_throughput_mode.py holds the gate).
enum : u32 {
  kDrInit = 0,
  kDrBurnin,
  kDrBind,
  kDrSplit,
  kDrLoop,  // + kind (0 loop, 1 TAD, 2 1D occupancy): kDrLoop, kDrTad, kDrOcc
  kDrTad,
  kDrOcc,
  kDrMoves,
  kDrBarriers,
  kDrLefBar,
  kDrPrimary,
  kDrSecondary,
  kDrRelease,
};
// counter layout: epoch (24 bits) | phase (4) | item (28) | draw number within the item (8)
MB_FN u64 ctr_pack(u64 epoch, u32 phase, u32 item) {
  return (epoch << 40) | (u64(phase & 15u) << 36) | (u64(item & 0x0FFFFFFFu) << 8);
}
constexpr u64 kCtrDrawsPerItem = 255;
MB_FN u64 mix64(u64 z) {  // SplitMix64 output function
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// The per-cell simulator. All member functions are CTA-collective unless noted.
// kCtr = false: deterministic mode (the reference's draw order); true: throughput mode.
template <bool kCtr>
struct CellSimT {
  const KernelParams& P;
  const IntervalData& D;
  CellArrays A;
  CellShared& S;
  Sinks K;
  Cta cta;
  CellTaskDev task;
  u64 key1 = 0, key2 = 0;  // throughput mode: per-cell keys of the counter-based draws

  // ------------------------------------------------------------------------------ helpers
  // [lo, hi): thread tid's contiguous share of n items (n < 2^16, tid < 2^10: no overflow)
  MB_FN void chunk(int tid, u32 n, u32* lo, u32* hi) const {
    *lo = cta.div_nt(n * static_cast<u32>(tid));
    *hi = cta.div_nt(n * (static_cast<u32>(tid) + 1));
  }
  MB_FN void fault(u32 code) const {
    if (S.fault == 0) S.fault = code;  // benign race: any code is reported
  }
  // deterministic mode: `off` is an offset into the cell's stream (staged in the ring);
  // throughput mode: `off` is a counter (ctr_pack + draw number)
  MB_FN u64 raw(u64 off) const {
    if constexpr (kCtr) {
      return mix64((off ^ key1) * 0x9E3779B97F4A7C15ull + key2);
    } else {
      return MB_LD_RING_U64(A.rng_ring + (off & (2 * u64(P.rng_window) - 1)));
    }
  }
  MB_FN void ctr_keys() {
    key1 = mix64(task.rng_state[0] ^ mix64(task.rng_state[1]));
    key2 = mix64(task.rng_state[2] ^ mix64(task.rng_state[3]));
  }

  // Phase timing: thread 0 charges the SM-clock cycles since the previous lap to phase `ph`.
  // (No-op in the CPU emulation.)
  u64 t_sub = 0;
  MB_FN void sub_begin() {
#if MB_DEVICE_BUILD
    if (threadIdx.x == 0) t_sub = static_cast<u64>(clock64());
#endif
  }
  MB_FN void sub_lap(int ph) {
#if MB_DEVICE_BUILD
    if (threadIdx.x == 0) {
      const u64 now = static_cast<u64>(clock64());
      S.phase_cycles[ph] += now - t_sub;
      t_sub = now;
    }
#else
    (void)ph;
#endif
  }
  u64 t_last = 0;
  MB_FN void lap(int ph) {
#if MB_DEVICE_BUILD
    if (threadIdx.x == 0) {
      const u64 now = static_cast<u64>(clock64());
      S.phase_cycles[ph] += now - t_last;
      t_last = now;
    }
#else
    // the emulation charges CTA barriers instead of cycles (scripts/barriers_per_epoch.py)
    const u64 now = emu_barrier_count();
    emu_phase_barriers()[ph] += now - t_last;
    t_last = now;
#endif
  }

  // A serial reader of the raw stream used by the few inherently sequential samplers.
  struct Cursor {
    const CellSimT* sim;
    u64 pos, limit;
    bool overrun;
    MB_FN u64 next() {
      if (pos >= limit) {
        overrun = true;
        return 0;
      }
      return sim->raw(pos++);
    }
    MB_FN double uniform01() {  // boost::random::uniform_01<double>
      for (;;) {
        const double r = MB_U64_TO_F64(next()) * (1.0 / kTwo64);
        if (r < 1.0 || overrun) return r;
      }
    }
  };
  MB_FN Cursor cursor(u64 pos, u64 limit) const { return Cursor{this, pos, limit, false}; }
  // throughput mode: the private draw sequence of (epoch, phase, item)
  MB_FN Cursor ctr_cursor(u64 epoch, u32 phase, u32 item) const {
    const u64 c = ctr_pack(epoch, phase, item);
    return Cursor{this, c, c + kCtrDrawsPerItem, false};
  }

  // Generator sub-stream gi writes its l draws of the window that starts at stream offset wbase
  // and hops to its block of the next window.
  MB_FN void rng_generate_block(u32 gi, u64 wbase) const {
    const u32 G = P.rng_gen_threads;
    Xs g{A.rng_state[gi], A.rng_state[G + gi], A.rng_state[2 * G + gi], A.rng_state[3 * G + gi]};
    const Xs nxt = xs_jump(g, D.jump_tbl);
    const u64 mask = 2 * u64(P.rng_window) - 1;
    const u64 o0 = wbase + u64(gi) * P.rng_per_thread;
    for (u32 i = 0; i < P.rng_per_thread; i += 4) {  // one full 32-byte sector per store
      const u64 a = xs_next(g);
      const u64 b = xs_next(g);
      const u64 c = xs_next(g);
      const u64 e = xs_next(g);
      u64* dst = A.rng_ring + ((o0 + i) & mask);
#if MB_DEVICE_BUILD
      asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "l"(a), "l"(b), "l"(c),
                   "l"(e)
                   : "memory");
#else
      dst[0] = a;
      dst[1] = b;
      dst[2] = c;
      dst[3] = e;
#endif
    }
    A.rng_state[gi] = nxt.s0;
    A.rng_state[G + gi] = nxt.s1;
    A.rng_state[2 * G + gi] = nxt.s2;
    A.rng_state[3 * G + gi] = nxt.s3;
  }

  // ------------------------------------------------------------------------------ RNG staging
  // Makes raw(o) valid for every o in [S.rng_pos, need_end). The stream is produced window by
  // window (W draws): generator thread g owns the l consecutive draws [g*l, (g+1)*l) of each
  // window and hops to the next window with a precomputed GF(2) matrix (T^W).
  MB_FN void rng_ensure(u64 need_end) {
    if constexpr (!kCtr) rng_ensure_staged(need_end);  // throughput mode stages nothing
  }
  MB_FN void rng_ensure_staged(u64 need_end) {
    if (need_end > S.rng_pos + P.rng_window) {
      MB_REGION(cta, tid) {
        if (cta.leader(tid)) fault(kFaultRngWindow);
      }
      cta.sync();
      return;
    }
#if MB_DEVICE_BUILD
    const bool refill = S.rng_generated < need_end;
    u64 t_refill = 0;
    if (refill && threadIdx.x == 0) t_refill = static_cast<u64>(clock64());
#endif
    while (S.rng_generated < need_end) {
      const u64 wbase = S.rng_generated;
      cta.sync();
      MB_REGION(cta, tid) {
        const u32 G = P.rng_gen_threads;
        for (u32 gi = static_cast<u32>(tid); gi < G; gi += static_cast<u32>(cta.nt()))
          rng_generate_block(gi, wbase);
        if (cta.leader(tid)) S.rng_generated = wbase + P.rng_window;
      }
      cta.sync();
    }
#if MB_DEVICE_BUILD
    if (refill && threadIdx.x == 0)
      S.phase_cycles[kPhRngGenerate] += static_cast<u64>(clock64()) - t_refill;
#endif
  }

  // Positions the G generator sub-streams at offsets g*l of the cell's stream (leader walks the
  // stream once; W steps).
  MB_FN void rng_bootstrap() {
    MB_REGION(cta, tid) {
      if (cta.leader(tid)) {
        Xs g{task.rng_state[0], task.rng_state[1], task.rng_state[2], task.rng_state[3]};
        const u32 G = P.rng_gen_threads;
        for (u32 t = 0; t < G; ++t) {
          A.rng_state[t] = g.s0;
          A.rng_state[G + t] = g.s1;
          A.rng_state[2 * G + t] = g.s2;
          A.rng_state[3 * G + t] = g.s3;
          for (u32 i = 0; i < P.rng_per_thread; ++i) xs_next(g);
        }
        S.rng_pos = 0;
        S.rng_generated = 0;
      }
    }
    cta.sync();
  }

  // ------------------------------------------------------------------------------ serial samplers
  // boost::random::poisson_distribution<size_t,double> (inversion below 10, PTRD above)
  MB_FN u64 poisson_serial(Cursor& c, double mean) const {
    if (mean < 10) {
      double p = exp(-mean);
      u64 x = 0;
      double u = c.uniform01();
      while (u > p) {
        u = u - p;
        ++x;
        p = mean * p / static_cast<double>(x);
      }
      return x;
    }
    const double log_fact[10] = {0.0,
                                 0.0,
                                 0.69314718055994529,
                                 1.7917594692280550,
                                 3.1780538303479458,
                                 4.7874917427820458,
                                 6.5792512120101012,
                                 8.5251613610654147,
                                 10.604602902745251,
                                 12.801827480081469};
    const double smu = sqrt(mean);
    const double b = 0.931 + 2.53 * smu;
    const double a = -0.059 + 0.02483 * b;
    const double inv_alpha = 1.1239 + 1.1328 / (b - 3.4);
    const double v_r = 0.9277 - 3.6224 / (b - 2);
    for (;;) {
      if (c.overrun) return 0;
      double u;
      double v = c.uniform01();
      if (v <= 0.86 * v_r) {
        u = v / v_r - 0.43;
        return static_cast<u64>(floor((2 * a / (0.5 - fabs(u)) + b) * u + mean + 0.445));
      }
      if (v >= v_r) {
        u = c.uniform01() - 0.5;
      } else {
        u = v / v_r - 0.93;
        u = ((u < 0) ? -0.5 : 0.5) - u;
        v = c.uniform01() * v_r;
      }
      const double us = 0.5 - fabs(u);
      if (us < 0.013 && v > us) continue;
      const double k = floor((2 * a / us + b) * u + mean + 0.445);
      v = v * inv_alpha / (a / (us * us) + b);
      const double log_sqrt_2pi = 0.91893853320467267;
      if (k >= 10) {
        if (log(v * smu) <= (k + 0.5) * log(mean / k) - mean - log_sqrt_2pi + k -
                                (1 / 12. - (1 / 360. - 1 / (1260. * k * k)) / (k * k)) / k)
          return static_cast<u64>(k);
      } else if (k >= 0) {
        if (log(v) <= k * log(mean) - mean - log_fact[static_cast<int>(k)])
          return static_cast<u64>(k);
      }
    }
  }

  MB_FN static double binomial_fc(i64 k) {
    const double tbl[10] = {0.08106146679532726, 0.04134069595540929, 0.02767792568499834,
                            0.02079067210376509, 0.01664469118982119, 0.01387612882307075,
                            0.01189670994589177, 0.01041126526197209, 0.009255462182712733,
                            0.008330563433362871};
    if (k < 10) return tbl[k];
    const double ikp1 = 1.0 / static_cast<double>(k + 1);
    return (1.0 / 12 - (1.0 / 360 - (1.0 / 1260) * (ikp1 * ikp1)) * (ikp1 * ikp1)) * ikp1;
  }

  // boost::random::binomial_distribution<ptrdiff_t,double> (inversion when (t+1)p < 11, else BTRD)
  MB_FN i64 binomial_serial(Cursor& c, i64 t, double p_in) const {
    const bool flip = 0.5 < p_in;
    const double p = flip ? (1 - p_in) : p_in;
    const i64 m = static_cast<i64>(static_cast<double>(t + 1) * p);
    i64 res = 0;
    if (m < 11) {
      const double q_n = pow(1 - p, static_cast<double>(t));
      const double q = 1 - p;
      const double s = p / q;
      const double a = static_cast<double>(t + 1) * s;
      double r = q_n;
      double u = c.uniform01();
      i64 x = 0;
      while (u > r) {
        u = u - r;
        ++x;
        const double r1 = ((a / static_cast<double>(x)) - s) * r;
        if (r1 < 2.220446049250313e-16 && r1 < r) break;
        r = r1;
      }
      res = x;
    } else {
      const double td = static_cast<double>(t);
      const double r = p / (1 - p);
      const double nr = static_cast<double>(t + 1) * r;
      const double npq = td * p * (1 - p);
      const double sqrt_npq = sqrt(npq);
      const double b = 1.15 + 2.53 * sqrt_npq;
      const double a = -0.0873 + 0.0248 * b + 0.01 * p;
      const double cc = td * p + 0.5;
      const double alpha = (2.83 + 5.1 / b) * sqrt_npq;
      const double v_r = 0.92 - 4.2 / b;
      const double u_rv_r = 0.86 * v_r;
      for (;;) {
        if (c.overrun) return 0;
        double u;
        double v = c.uniform01();
        if (v <= u_rv_r) {
          u = v / v_r - 0.43;
          res = static_cast<i64>(floor((2 * a / (0.5 - fabs(u)) + b) * u + cc));
          break;
        }
        if (v >= v_r) {
          u = c.uniform01() - 0.5;
        } else {
          u = v / v_r - 0.93;
          u = ((u < 0) ? -0.5 : 0.5) - u;
          v = c.uniform01() * v_r;
        }
        const double us = 0.5 - fabs(u);
        const i64 k = static_cast<i64>(floor((2 * a / us + b) * u + cc));
        if (k < 0 || k > t) continue;
        v = v * alpha / (a / (us * us) + b);
        const double km = static_cast<double>(k > m ? k - m : m - k);
        if (km <= 15) {
          double f = 1;
          if (m < k) {
            i64 i = m;
            do {
              ++i;
              f = f * (nr / static_cast<double>(i) - r);
            } while (i != k);
          } else if (m > k) {
            i64 i = k;
            do {
              ++i;
              v = v * (nr / static_cast<double>(i) - r);
            } while (i != m);
          }
          if (v <= f) {
            res = k;
            break;
          }
          continue;
        }
        v = log(v);
        const double rho = (km / npq) * (((km / 3. + 0.625) * km + 1. / 6) / npq + 0.5);
        const double tt = -km * km / (2 * npq);
        if (v < tt - rho) {
          res = k;
          break;
        }
        if (v > tt + rho) continue;
        const i64 nm = t - m + 1;
        const double h = (static_cast<double>(m) + 0.5) *
                             log(static_cast<double>(m + 1) / (r * static_cast<double>(nm))) +
                         binomial_fc(m) + binomial_fc(t - m);
        const i64 nk = t - k + 1;
        if (v <= h +
                     static_cast<double>(t + 1) *
                         log(static_cast<double>(nm) / static_cast<double>(nk)) +
                     (static_cast<double>(k) + 0.5) *
                         log(static_cast<double>(nk) * r / static_cast<double>(k + 1)) -
                     binomial_fc(k) - binomial_fc(t - k)) {
          res = k;
          break;
        }
      }
    }
    return flip ? t - res : res;
  }

  // detail::unit_exponential_distribution<double> (256-layer ziggurat)
  MB_FN double unit_exponential_serial(Cursor& c) const {
    const double* tx = D.zig_ex;
    const double* ty = D.zig_ey;
    double shift = 0;
    for (;;) {
      if (c.overrun) return 0;
      int i;
      const double u = int_float_pair8(c.next(), &i);
      const double x = u * tx[i];
      if (x < tx[i + 1]) return shift + x;
      if (i == 0) {
        shift += tx[1];
      } else {
        const double y01 = c.uniform01();
        const double y = ty[i] + y01 * (ty[i + 1] - ty[i]);
        const double y_above_ubound = (tx[i] - tx[i + 1]) * y01 - (tx[i] - x);
        const double y_above_lbound = y - (ty[i + 1] + (tx[i + 1] - x) * ty[i + 1]);
        if (y_above_ubound < 0 && (y_above_lbound < 0 || y < exp(-x))) return x + shift;
      }
    }
  }

  // Fast path of detail::unit_normal_distribution<double>: true when the draw is accepted at once.
  MB_FN bool unit_normal_fast(u64 u, double* z) const {
    int bits;
    const double r = int_float_pair8(u, &bits);
    const int sign = (bits & 1) * 2 - 1;
    const int i = bits >> 1;
    const double x = r * A.zig_nx[i];
    *z = x * sign;
    return x < A.zig_nx[i + 1];
  }
  // Full sampler from a cursor (the first draw is taken from the cursor as well).
  MB_FN double unit_normal_serial(Cursor& c) const {
    const double* tx = A.zig_nx;
    const double* ty = D.zig_ny;
    for (;;) {
      if (c.overrun) return 0;
      int bits;
      const double r = int_float_pair8(c.next(), &bits);
      const int sign = (bits & 1) * 2 - 1;
      const int i = bits >> 1;
      const double x = r * tx[i];
      if (x < tx[i + 1]) return x * sign;
      if (i == 0) {
        const double tail_start = tx[1];
        for (;;) {
          if (c.overrun) return 0;
          const double xx = unit_exponential_serial(c) / tail_start;
          const double yy = unit_exponential_serial(c);
          if (2 * yy > xx * xx) return (xx + tail_start) * sign;
        }
      }
      const double y01 = c.uniform01();
      const double y = ty[i] + y01 * (ty[i + 1] - ty[i]);
      double y_above_ubound, y_above_lbound;
      if (tx[i] >= 1) {
        y_above_ubound = (tx[i] - tx[i + 1]) * y01 - (tx[i] - x);
        y_above_lbound = y - (ty[i] + (tx[i] - x) * ty[i] * tx[i]);
      } else {
        y_above_lbound = (tx[i] - tx[i + 1]) * y01 - (tx[i] - x);
        y_above_ubound = y - (ty[i] + (tx[i] - x) * ty[i] * tx[i]);
      }
      if (y_above_ubound < 0 && (y_above_lbound < 0 || y < exp(-(x * x / 2)))) return x * sign;
    }
  }

  // ------------------------------------------------------------------------------ init
  // State::operator=(Task) + reset_buffers (simulation.cpp:617-627,741-761) and
  // ExtrusionBarriers::init_states (extrusion_barriers.cpp:219-230).
  MB_FN void init_cell() {
    MB_REGION(cta, tid) {
      for (u32 i = tid; i < P.n_lefs; i += cta.nt()) {
        A.rev[i] = kUnbound;
        A.fwd[i] = kUnbound;
        A.ep[i] = kUnbound;
        A.rr[i] = static_cast<u16>(i);
        A.fr[i] = static_cast<u16>(i);
        A.rm[i] = 0;
        A.fm[i] = 0;
        A.rc[i] = 0;
        A.fc[i] = 0;
      }
      for (u32 i = tid; i < P.n_bar; i += cta.nt()) A.bar_pos[i] = D.bar_pos[i];
      for (u32 i = tid; i < 129; i += cta.nt()) A.zig_nx[i] = D.zig_nx[i];
      mv_clear_slow_bits(tid);
      // barrier look-up table of the LEF-BAR walk: entry e = number of barriers below the first
      // position of bucket e (binary search over the interval's own copy of the positions)
      for (u32 e = tid; e < P.lut_entries; e += cta.nt()) {
        const u64 first = u64(P.start) + (u64(e) << P.lut_shift);
        u32 a = 0, z = P.n_bar;
        while (a < z) {
          const u32 mid = (a + z) >> 1;
          if (u64(D.bar_pos[mid]) < first) {
            a = mid + 1;
          } else {
            z = mid;
          }
        }
        A.bar_lut[e] = static_cast<u16>(a);
      }
      if (cta.leader(tid)) {
        S.epoch = 0;
        S.num_burnin_epochs = 0;
        S.num_contacts = 0;
        S.lef_updates = 0;
        S.num_active = 0;
        S.burnin_completed = 0;
        S.fault = 0;
        S.hist_len = 0;
        S.hist_head = 0;
        S.done = 0;
        S.move_bound_hit = 0;
        for (int k = 0; k < kNumPhases; ++k) S.phase_cycles[k] = 0;
        if (P.burnin_history > kMaxBurninHistory || P.burnin_window + 1 >= P.burnin_history)
          fault(kFaultBurninHistory);
      }
    }
    cta.sync();
    if constexpr (kCtr) {
      MB_REGION(cta, tid) {
        for (u32 i = tid; i < P.n_bar; i += cta.nt()) {
          const double occ = D.bar_occupancy[i];
          bool act = false;
          if (occ != 0.0) act = bernoulli_raw(raw(ctr_pack(0, kDrInit, i)), occ);
          A.bar_active[i] = act ? 1 : 0;
        }
        if (cta.leader(tid)) {
          S.rng_pos = 0;
          S.rng_generated = 0;
          if (P.skip_burnin) {
            S.num_active = P.n_lefs;
            S.burnin_completed = 1;
          }
        }
      }
      cta.sync();
      return;
    }
    rng_bootstrap();

    // init_states: one Bernoulli(occupancy) per barrier whose occupancy is not 0
    rng_ensure(S.rng_pos + P.n_bar);
    PerThread<u64> cnt(cta.nt());
    MB_REGION(cta, tid) {
      u32 lo, hi;
      chunk(tid, P.n_bar, &lo, &hi);
      u64 c = 0;
      for (u32 i = lo; i < hi; ++i) c += D.bar_occupancy[i] != 0.0;
      cnt[tid] = c;
    }
    const u64 total = cta.exscan_sum(cnt);
    MB_REGION(cta, tid) {
      u32 lo, hi;
      chunk(tid, P.n_bar, &lo, &hi);
      u64 o = S.rng_pos + cnt[tid];
      for (u32 i = lo; i < hi; ++i) {
        const double occ = D.bar_occupancy[i];
        bool act = false;
        if (occ != 0.0) act = bernoulli_raw(raw(o++), occ);
        A.bar_active[i] = act ? 1 : 0;
      }
    }
    cta.sync();
    MB_REGION(cta, tid) {
      if (cta.leader(tid)) {
        S.rng_pos += total;
        if (P.skip_burnin) {
          S.num_active = P.n_lefs;
          S.burnin_completed = 1;
        }
      }
    }
    cta.sync();
  }

  // ------------------------------------------------------------------------------ burn-in
  // run_burnin (simulation.cpp:866-894), compute_loop_size_stats (:795-819),
  // evaluate_burnin (:821-864).
  MB_FN u32 count_dips(const double* hist, int tid) const {
    // hist is a ring of length cap starting at S.hist_head
    const u32 cap = P.burnin_history, w = P.burnin_window;
    u32 n = 0;
    for (u32 j = 1 + tid; j + w < cap; j += cta.nt()) {
      double a1 = 0.0, a2 = 0.0;
      for (u32 k = 0; k < w; ++k) a1 = a1 + hist[(S.hist_head + j - 1 + k) % cap];
      for (u32 k = 0; k < w; ++k) a2 = a2 + hist[(S.hist_head + j + k) % cap];
      const double n1 = a1 / static_cast<double>(w);
      const double n2 = a2 / static_cast<double>(w);
      n += n1 > n2;
    }
    return n;
  }

  MB_FN void burnin_step() {
    for (;;) {
      const bool activating = S.num_active != P.n_lefs;
      // every thread evaluates S.rng_pos here, the leader moves it in the region below: the
      // barrier has to sit between the two (scripts/lint_shared_state.py checks this pattern)
      if (activating) rng_ensure(S.rng_pos + 256);
      cta.sync();
      if (activating) {
        MB_REGION(cta, tid) {
          if (cta.leader(tid)) {
            ++S.num_burnin_epochs;
            Cursor c = kCtr ? ctr_cursor(S.num_burnin_epochs, kDrBurnin, 0)
                            : cursor(S.rng_pos, S.rng_pos + 256);
            const u64 k = poisson_serial(c, P.lef_binding_rate_burnin);
            if (c.overrun) fault(kFaultSerialDraws);
            if constexpr (!kCtr) S.rng_pos = c.pos;
            const u64 na = u64(S.num_active) + k;
            S.num_active = na < P.n_lefs ? static_cast<u32>(na) : P.n_lefs;
          }
        }
        cta.sync();
      } else {
        const u32 n = S.num_active;
        // mean loop size: the integer sum is exact in double, so any order gives the reference's
        // left-to-right result; the squared-deviation sum is reduced in a fixed tree order
        // (DESIGN.md "burn-in statistics").
        PerThread<u64> isum(cta.nt());
        MB_REGION(cta, tid) {
          u64 acc = 0;
          for (u32 i = tid; i < n; i += cta.nt())
            acc += (A.ep[i] != kUnbound) ? (A.fwd[i] - A.rev[i]) : 0u;
          isum[tid] = acc;
        }
        const u64 tot = cta.reduce_sum(isum);
        const double mean = static_cast<double>(tot) / static_cast<double>(n);
        PerThread<double> dsum(cta.nt());
        MB_REGION(cta, tid) {
          double acc = 0.0;
          for (u32 i = tid; i < n; i += cta.nt()) {
            const u32 ls = (A.ep[i] != kUnbound) ? (A.fwd[i] - A.rev[i]) : 0u;
            const double d = static_cast<double>(ls) - mean;
            acc = acc + (d * d);
          }
          dsum[tid] = acc;
        }
        const double ssd = cta.reduce_sum_f64(dsum);
        MB_REGION(cta, tid) {
          if (cta.leader(tid)) {
            ++S.num_burnin_epochs;
            const double sd = sqrt(ssd / static_cast<double>(n));
            const u32 cap = P.burnin_history;
            if (S.hist_len == cap) {
              S.hist_head = (S.hist_head + 1) % cap;
              --S.hist_len;
            }
            const u32 slot = (S.hist_head + S.hist_len) % cap;
            S.avg_hist[slot] = mean;
            S.cv_hist[slot] = sd / mean;
            ++S.hist_len;
          }
        }
        cta.sync();
        bool completed = false;
        if (S.hist_len == P.burnin_history) {
          const u32 cap = P.burnin_history, w = P.burnin_window;
          PerThread<u64> dips(cta.nt());
          MB_REGION(cta, tid) { dips[tid] = count_dips(S.cv_hist, tid); }
          const u64 n1 = cta.reduce_sum(dips);
          const double r1 = static_cast<double>(n1) / static_cast<double>(cap - w - n1);
          if (r1 >= 0.95 && r1 <= 1.05) {
            MB_REGION(cta, tid) { dips[tid] = count_dips(S.avg_hist, tid); }
            const u64 n2 = cta.reduce_sum(dips);
            const double r2 = static_cast<double>(n2) / static_cast<double>(cap - w - n2);
            completed = r2 >= 0.95 && r2 <= 1.05;
          }
        }
        completed = completed && (S.epoch > P.min_burnin_epochs);
        bool force = false;
        if (!completed && S.epoch >= P.max_burnin_epochs) force = true;
        cta.sync();
        MB_REGION(cta, tid) {
          if (cta.leader(tid)) {
            if (completed || force) S.burnin_completed = 1;
            if (force) S.num_active = P.n_lefs;
          }
        }
        cta.sync();
      }
      if (S.num_active != 0 || S.fault != 0) break;
    }
    cta.sync();
  }

  // ------------------------------------------------------------------------------ bind + rank
  // select_lefs_to_bind + bind_lefs (simulation_impl.hpp:30-91)
  MB_FN void bind_lefs() {
    const u32 n = S.num_active;
    const u64 range = u64(P.end - 1) - u64(P.start);
    const InvU64 bucket = inv_u64(range ? uniform_int_bucket(range) : 1);
    const u32 cur = static_cast<u32>(S.epoch);
    if constexpr (kCtr) {
      // Every unbound LEF draws its position from its own sequence: one pass, no repair. Few
      // LEFs are unbound in an epoch (a lane or two of a warp per trip), so a thread first notes
      // which of its LEFs are and then handles those back to back: the lanes of a warp that have
      // any then go through the draw together instead of a lane or two per trip.
      MB_REGION(cta, tid) {
        for (u32 i0 = tid; i0 < n; i0 += 32u * cta.nt()) {
          u32 todo = 0, k = 0;
          for (u32 i = i0; k < 32 && i < n; i += cta.nt(), ++k) todo |= u32(A.ep[i] == kUnbound) << k;
          while (todo) {
            k = static_cast<u32>(MB_FFS(todo)) - 1;
            todo &= todo - 1;
            const u32 i = i0 + k * cta.nt();
            u64 r = 0;
            if (range != 0) {
              Cursor c = ctr_cursor(S.epoch, kDrBind, i);
              do {
                r = div_u64(c.next(), bucket);
              } while (r > range && !c.overrun);
              if (c.overrun) fault(kFaultSerialDraws);
            }
            A.rev[i] = A.fwd[i] = static_cast<u32>(u64(P.start) + r);
            A.ep[i] = cur;
          }
        }
      }
      cta.sync();
      return;
    }
    PerThread<u64> cnt(cta.nt());
    MB_REGION(cta, tid) {
      u32 lo, hi;
      chunk(tid, n, &lo, &hi);
      u64 c = 0;
      for (u32 i = lo; i < hi; ++i) c += A.ep[i] == kUnbound;
      cnt[tid] = c;
      if (cta.leader(tid)) S.tmp_u32[0] = 0;
    }
    const u64 total = cta.exscan_sum(cnt);
    if (total == 0) return;
    if (range == 0) {
      MB_REGION(cta, tid) {
        for (u32 i = tid; i < n; i += cta.nt()) {
          if (A.ep[i] == kUnbound) {
            A.rev[i] = A.fwd[i] = P.start;
            A.ep[i] = cur;
          }
        }
      }
      cta.sync();
      return;
    }
    rng_ensure(S.rng_pos + total + 64);
    // The k-th unbound LEF takes the k-th draw unless a uniform_int rejection (about one draw in
    // 2^36 for a human chromosome) shifts the later ones: bind with the draws at their default
    // offsets and note a rejection. LEFs bound here are the only ones that carry the current
    // epoch, which is how the sequential redo below finds them again.
    MB_REGION(cta, tid) {
      u32 lo, hi;
      chunk(tid, n, &lo, &hi);
      u64 o = S.rng_pos + cnt[tid];
      bool rejected = false;
#if !MB_DEVICE_BUILD
      rejected = emu_force_bind_redo() != 0;  // (tests: the redo without a rejection)
#endif
      for (u32 i = lo; i < hi; ++i) {
        if (A.ep[i] != kUnbound) continue;
        const u64 r = div_u64(raw(o++), bucket);
        if (r > range) rejected = true;
        A.rev[i] = A.fwd[i] = static_cast<u32>(u64(P.start) + r);
        A.ep[i] = cur;
      }
      if (rejected) S.tmp_u32[0] = 1;
    }
    cta.sync();
    const bool redo = S.tmp_u32[0] != 0;
    MB_REGION(cta, tid) {
      if (cta.leader(tid)) {
        if (!redo) {
          S.rng_pos += total;
        } else {
          Cursor c = cursor(S.rng_pos, S.rng_pos + total + 64);
          for (u32 i = 0; i < n; ++i) {
            if (A.ep[i] != cur) continue;
            u64 r;
            do {
              r = div_u64(c.next(), bucket);
            } while (r > range && !c.overrun);
            A.rev[i] = A.fwd[i] = static_cast<u32>(u64(P.start) + r);
          }
          if (c.overrun) fault(kFaultSerialDraws);
          S.rng_pos = c.pos;
        }
      }
    }
    cta.sync();
  }

  // rank_lefs (simulation.cpp:410-496): total order (pos, binding epoch asc/desc, previous slot).
  template <bool kRev>
  MB_FN bool rank_less(u32 a, u32 b, const u16* prev_slot) const {
    const u32 pa = kRev ? A.rev[a] : A.fwd[a];
    const u32 pb = kRev ? A.rev[b] : A.fwd[b];
    if (pa != pb) return pa < pb;
    const u32 ea = A.ep[a], eb = A.ep[b];
    if (ea != eb) return kRev ? ea < eb : ea > eb;
    return prev_slot[a] < prev_slot[b];
  }

  // Full sort of both rank arrays (first epoch of --skip-burnin, or whenever the incremental
  // path below finds its assumptions violated). prev_r / prev_f: previous slot of every LEF.
  MB_FN void rank_lefs_full(const u16* prev_r, const u16* prev_f) {
    const u32 n = S.num_active;
#if MB_DEVICE_BUILD
    u32 npad = 1;
    while (npad < n) npad <<= 1;
    u16* kr = reinterpret_cast<u16*>(A.rm);  // npad <= 2n u16 fit in n u32
    u16* kf = reinterpret_cast<u16*>(A.fm);
    const int tid = cta.first();
    for (u32 k = tid; k < npad; k += cta.nt()) {
      kr[k] = k < n ? A.rr[k] : static_cast<u16>(0xFFFF);
      kf[k] = k < n ? A.fr[k] : static_cast<u16>(0xFFFF);
    }
    __syncthreads();
    for (u32 size = 2; size <= npad; size <<= 1) {
      for (u32 stride = size >> 1; stride > 0; stride >>= 1) {
        for (u32 t = tid; t < npad / 2; t += cta.nt()) {
          const u32 lo = 2 * t - (t & (stride - 1));
          const u32 hi = lo + stride;
          const bool asc = (lo & size) == 0;
          {
            const u16 x = kr[lo], y = kr[hi];
            // sentinel 0xFFFF sorts last
            const bool y_lt_x = (y != 0xFFFF) && (x == 0xFFFF || rank_less<true>(y, x, prev_r));
            const bool x_lt_y = (x != 0xFFFF) && (y == 0xFFFF || rank_less<true>(x, y, prev_r));
            if (asc ? y_lt_x : x_lt_y) {
              kr[lo] = y;
              kr[hi] = x;
            }
          }
          {
            const u16 x = kf[lo], y = kf[hi];
            const bool y_lt_x = (y != 0xFFFF) && (x == 0xFFFF || rank_less<false>(y, x, prev_f));
            const bool x_lt_y = (x != 0xFFFF) && (y == 0xFFFF || rank_less<false>(x, y, prev_f));
            if (asc ? y_lt_x : x_lt_y) {
              kf[lo] = y;
              kf[hi] = x;
            }
          }
        }
        __syncthreads();
      }
    }
    for (u32 k = tid; k < n; k += cta.nt()) {
      A.rr[k] = kr[k];
      A.fr[k] = kf[k];
    }
    __syncthreads();
#else
    MB_REGION(cta, tid) {
      if (cta.leader(tid)) {
        std::sort(A.rr, A.rr + n, [&](u16 a, u16 b) { return rank_less<true>(a, b, prev_r); });
        std::sort(A.fr, A.fr + n, [&](u16 a, u16 b) { return rank_less<false>(a, b, prev_f); });
      }
    }
    cta.sync();
#endif
  }

  // rank_lefs (simulation.cpp:410-496). extrude() keeps the relative order of the LEFs that were
  // not (re)bound this epoch, so the new permutation is the old one minus the changed LEFs,
  // merged with the few changed ones: compact the unchanged LEFs, rank the changed ones among
  // themselves by counting and among the unchanged ones by binary search, then place everything.
  // The result is verified (adjacent pairs under the total order); ties that extrude() created
  // between unchanged LEFs are repaired by odd-even transposition passes, and anything else
  // (more changed LEFs than the scratch holds, a still unsorted list) takes the full sort.
  MB_FN void rank_lefs() {
    const u32 n = S.num_active;
    if (n < 2) return;
    // scratch areas that are dead at this point of the epoch: moves, collision words, scratch,
    // and the bitmaps of the secondary pass (two "dirty pair" bitmaps live there)
    u16* prev_r = reinterpret_cast<u16*>(A.rc);  // previous slot of every LEF in the rev order
    u16* prev_f = reinterpret_cast<u16*>(A.fc);
    const u32 nwords = (n + 31) / 32;
    u32* dirty_r = A.bits;  // bit k: the pair of rank slots (k, k + 1) has to be looked at
    u32* dirty_f = A.bits + nwords;
    const u32 cur = static_cast<u32>(S.epoch);
    u32 max_changed = static_cast<u32>(cell_scratch_words(P.n_lefs, P.n_bar) / 6);  // see the merge
    if (max_changed > 512) max_changed = 512;
    PerThread<u64> cnt(cta.nt());
    MB_PROBE_BEGIN();
    MB_REGION(cta, tid) {
      u32 lo, hi;
      chunk(tid, n, &lo, &hi);
      u64 c = 0;
      for (u32 k = lo; k < hi; ++k) {
        const u32 r = A.rr[k], f = A.fr[k];
        prev_r[r] = static_cast<u16>(k);
        prev_f[f] = static_cast<u16>(k);
        c += A.ep[r] == cur;
        c += u64(A.ep[f] == cur) << 32;
      }
      cnt[tid] = c;
      for (u32 w = tid; w < 2 * nwords; w += cta.nt()) A.bits[w] = 0;
      if (cta.leader(tid)) S.tmp_u32[4] = S.tmp_u32[5] = S.tmp_u32[6] = 0;
    }
    const u64 tot = cta.exscan_sum(cnt);
    const u32 nc = static_cast<u32>(tot & 0xFFFFFFFFu);
    if (nc > max_changed) {
      rank_lefs_full(prev_r, prev_f);
      return;
    }
    MB_PROBE(kPhRngGenerate);
    if (nc != 0) rank_lefs_merge(cnt, nc, prev_r, prev_f);
    MB_PROBE(kPhMvEnsure);
    // Verify / repair. Units can legitimately cross during extrude() (a unit is not tested
    // against the unit behind an avoided secondary collision) and new ties need their epoch
    // order: typically a handful of adjacent pairs of a few thousand are out of order, each a
    // slot or two from its place. One sweep over all adjacent pairs marks the out-of-order ones
    // in a bitmap; after that only marked pairs are looked at: a round takes the marked pairs
    // of one parity (they are disjoint), swaps those that are out of order and marks their two
    // neighbours -- the only pairs a swap can break -- for the next round. The invariant "every
    // out-of-order pair is marked" holds throughout, so when no marks are left the permutation
    // is the sorted one (the order is total, so any correct sort gives the same result).
    MB_REGION(cta, tid) {
      u32 lo, hi;
      chunk(tid, n, &lo, &hi);
      bool any = false;
      // (the unit in slot k + 1 of one trip is the unit in slot k of the next: read once; only a
      // tie on the position needs the full comparison)
      u32 ar = lo < hi ? A.rr[lo] : 0u, af = lo < hi ? A.fr[lo] : 0u;
      u32 par = lo < hi ? A.rev[ar] : 0u, paf = lo < hi ? A.fwd[af] : 0u;
      for (u32 k = lo; k < hi && k + 1 < n; ++k) {
        const u32 br = A.rr[k + 1], bf = A.fr[k + 1];
        const u32 pbr = A.rev[br], pbf = A.fwd[bf];
        if (pbr < par || (pbr == par && rank_less<true>(br, ar, prev_r))) {
          MB_ATOMIC_OR_U32(&dirty_r[k >> 5], 1u << (k & 31));
          any = true;
        }
        if (pbf < paf || (pbf == paf && rank_less<false>(bf, af, prev_f))) {
          MB_ATOMIC_OR_U32(&dirty_f[k >> 5], 1u << (k & 31));
          any = true;
        }
        ar = br;
        af = bf;
        par = pbr;
        paf = pbf;
      }
      if (any) MB_SHARED_STORE_U32(&S.tmp_u32[4], 1u);
    }
    cta.sync();
    MB_PROBE(kPhMvScan);
    constexpr u32 kMaxRounds = 64;
    for (u32 round = 0; round < kMaxRounds; ++round) {
      // Three flags take turns: round r reads tmp_u32[4 + r % 3] ("marks were left for this
      // round"), sets the next one for round r + 1 and clears the third -- which round r - 1 read
      // before the barrier that ended it, and which nobody writes until round r + 1.
      // (Shared by design: the word read here is not the one this round's region writes.)
      if (MB_SHARED_LOAD_U32(&S.tmp_u32[4 + round % 3]) == 0) return;
      const u32 parity_mask = (round & 1) ? 0xAAAAAAAAu : 0x55555555u;
      MB_REGION(cta, tid) {
        bool left = false;
        for (u32 w = tid; w < 2 * nwords; w += cta.nt()) {
          const bool is_rev = w < nwords;
          const u32 ww = is_rev ? w : w - nwords;
          u32* dirty = is_rev ? dirty_r : dirty_f;
          u16* ranks = is_rev ? A.rr : A.fr;
          const u32 all = MB_SHARED_LOAD_U32(&dirty[ww]);
          u32 mine = all & parity_mask;
          if (all & ~parity_mask) left = true;
          if (mine == 0) continue;
          MB_ATOMIC_AND_U32(&dirty[ww], ~mine);
          while (mine) {
            const u32 bit = static_cast<u32>(MB_FFS(mine)) - 1;
            mine &= mine - 1;
            const u32 k = 32 * ww + bit;
            if (k + 1 >= n) continue;
            const u16 a = ranks[k], b = ranks[k + 1];
            const bool swap = is_rev ? rank_less<true>(b, a, prev_r) : rank_less<false>(b, a, prev_f);
            if (!swap) continue;
            ranks[k] = b;
            ranks[k + 1] = a;
            if (k > 0) MB_ATOMIC_OR_U32(&dirty[(k - 1) >> 5], 1u << ((k - 1) & 31));
            if (k + 2 < n) MB_ATOMIC_OR_U32(&dirty[(k + 1) >> 5], 1u << ((k + 1) & 31));
            left = true;
          }
        }
        if (left) MB_SHARED_STORE_U32(&S.tmp_u32[4 + (round + 1) % 3], 1u);
        if (cta.leader(tid)) MB_SHARED_STORE_U32(&S.tmp_u32[4 + (round + 2) % 3], 0u);
      }
      cta.sync();
    }
    rank_lefs_full(prev_r, prev_f);  // (not sorted after kMaxRounds: cannot happen in practice)
  }

  // The merge step of rank_lefs: cnt = per-thread exclusive counts of changed LEFs (rev order in
  // the low word, fwd order in the high word), nc = number of changed LEFs.
  MB_FN void rank_lefs_merge(const PerThread<u64>& cnt, u32 nc, const u16* prev_r,
                             const u16* prev_f) {
    const u32 n = S.num_active;
    const u32 cur = static_cast<u32>(S.epoch);
    u16* un_r = reinterpret_cast<u16*>(A.rm);  // [0, nu): unchanged LEFs in their old order
    u16* un_f = reinterpret_cast<u16*>(A.fm);
    u16* ch_r = un_r + n;  // [0, nc): changed LEFs in their old order
    u16* ch_f = un_f + n;
    const u32 nu = n - nc;
    u32* rank_r = A.scratch;            // [nc] rank of each changed LEF among the changed ones
    u32* rank_f = A.scratch + nc;
    u32* kp_r = A.scratch + 2 * nc;     // [nc] position of each changed LEF (rev == fwd == where
    u32* kp_f = A.scratch + 3 * nc;     //      it was just bound), in the order of ch_r / ch_f
    u16* pos_r = reinterpret_cast<u16*>(A.scratch + 4 * nc);  // [nc] unchanged LEFs before it
    u16* pos_f = pos_r + nc;
    u16* ts_r = pos_f + nc;  // [nc] pos_* ordered by rank_*
    u16* ts_f = ts_r + nc;
    MB_REGION(cta, tid) {
      u32 lo, hi;
      chunk(tid, n, &lo, &hi);
      u32 cr = static_cast<u32>(cnt[tid] & 0xFFFFFFFFu), cf = static_cast<u32>(cnt[tid] >> 32);
      for (u32 k = lo; k < hi; ++k) {
        const u16 r = A.rr[k], f = A.fr[k];
        if (A.ep[r] == cur) {
          kp_r[cr] = A.rev[r];
          ch_r[cr++] = r;
        } else {
          un_r[k - cr] = r;
        }
        if (A.ep[f] == cur) {
          kp_f[cf] = A.fwd[f];
          ch_f[cf++] = f;
        } else {
          un_f[k - cf] = f;
        }
      }
      for (u32 j = tid; j < 2 * nc; j += cta.nt()) A.scratch[j] = 0;
    }
    cta.sync();
    // Work items: (changed LEF j, segment s of the changed list) -> partial rank among the
    // changed ones; then one binary search over the unchanged list per changed LEF; rev order
    // first, then fwd. A changed LEF was bound this epoch: both its units sit at one position and
    // carry the current epoch, so among changed LEFs the order is (position, previous slot) --
    // and the previous slots ascend along ch_r / ch_f, i.e. ties go by list index -- while
    // against an unchanged LEF (older epoch) a tie on the position is decided by the epoch rule
    // alone: older first in the rev order, newer first in the fwd order. (The unchanged lists
    // may still hold the few out-of-order pairs the repair below removes; the search then picks
    // some slot nearby and the repair settles it, as it does for every other pair.)
    u32 nseg = static_cast<u32>(cta.nt()) / (2 * nc);
    if (nseg < 1) nseg = 1;
    if (nseg > nc) nseg = nc;
    const u32 per_dir = nc * nseg + nc;
    MB_REGION(cta, tid) {
      for (u32 w = tid; w < 2 * per_dir; w += cta.nt()) {
        const bool is_rev = w < per_dir;
        const u32 v = is_rev ? w : w - per_dir;
        const u32* kp = is_rev ? kp_r : kp_f;
        if (v < nc * nseg) {
          const u32 j = v / nseg, s = v % nseg;
          const u32 a = s * nc / nseg, b = (s + 1) * nc / nseg;
          const u32 x = kp[j];
          u32 c = 0;
          for (u32 q = a; q < b; ++q) {
            const u32 y = kp[q];
            c += (y < x) || (y == x && q < j);
          }
          if (c) MB_ATOMIC_ADD_U32(is_rev ? &rank_r[j] : &rank_f[j], c);
        } else {
          const u32 j = v - nc * nseg;
          const u32 x = kp[j];
          u32 a = 0, b = nu;
          if (is_rev) {  // unchanged LEFs with rev pos <= x order before it
            while (a < b) {
              const u32 mid = (a + b) >> 1;
              if (A.rev[un_r[mid]] <= x) {
                a = mid + 1;
              } else {
                b = mid;
              }
            }
            pos_r[j] = static_cast<u16>(a);
          } else {  // unchanged LEFs with fwd pos < x order before it
            while (a < b) {
              const u32 mid = (a + b) >> 1;
              if (A.fwd[un_f[mid]] < x) {
                a = mid + 1;
              } else {
                b = mid;
              }
            }
            pos_f[j] = static_cast<u16>(a);
          }
        }
      }
    }
    cta.sync();
    MB_REGION(cta, tid) {
      for (u32 j = tid; j < nc; j += cta.nt()) {
        ts_r[rank_r[j]] = pos_r[j];
        ts_f[rank_f[j]] = pos_f[j];
        A.rr[rank_r[j] + pos_r[j]] = ch_r[j];
        A.fr[rank_f[j] + pos_f[j]] = ch_f[j];
      }
    }
    cta.sync();
    // unchanged LEF number u lands at slot u + #{changed LEFs with pos <= u}
    MB_REGION(cta, tid) {
      u32 lo, hi;
      chunk(tid, nu, &lo, &hi);
      if (lo < hi) {
        u32 jr = 0, jf = 0;
        {
          u32 a = 0, b = nc;
          while (a < b) {
            const u32 m = (a + b) >> 1;
            if (ts_r[m] <= lo) {
              a = m + 1;
            } else {
              b = m;
            }
          }
          jr = a;
          a = 0;
          b = nc;
          while (a < b) {
            const u32 m = (a + b) >> 1;
            if (ts_f[m] <= lo) {
              a = m + 1;
            } else {
              b = m;
            }
          }
          jf = a;
        }
        for (u32 u = lo; u < hi; ++u) {
          while (jr < nc && ts_r[jr] <= u) ++jr;
          while (jf < nc && ts_f[jf] <= u) ++jf;
          A.rr[u + jr] = un_r[u];
          A.fr[u + jf] = un_f[u];
        }
      }
    }
    cta.sync();
  }

  // ------------------------------------------------------------------------------ contacts
  // ContactMatrixDense::increment (contact_matrix_dense_safe_impl.hpp:54-68,86-89)
  MB_FN void band_increment(u32 b1, u32 b2) const {
    const u32 i = b1 > b2 ? b1 - b2 : b2 - b1;
    const u32 j = b1 > b2 ? b1 : b2;
    if (i >= P.nrows) {
      MB_ATOMIC_ADD_U64(K.missed, u64(1));
      return;
    }
    MB_ATOMIC_ADD_U32(K.band + (size_t(j) * P.nrows + i), 1u);
  }

  // One sampling event starting at stream offset o (register_contacts.cpp:122-232).
  // kind 0: loop contact, 1: TAD contact, 2: 1D occupancy. Returns the number of raw draws the
  // event consumes; *b1/*b2 receive the two bins (or kUnbound when nothing is registered).
  MB_FN u32 sampling_event(u64 o, int kind, u64 limit, u32* b1, u32* b2) const {
    *b1 = kUnbound;
    *b2 = kUnbound;
    Cursor c = cursor(o, limit);
    const u32 n = S.num_active;
    u64 i = 0;
    if (n > 1) {
      // bucket * (range + 1) <= 2^64 - 1: range + 1 serves as the reciprocal of the bucket
      // (short of the quotient by (range + 1) / bucket at most: nothing for ranges below 2^32)
      const u64 range = n - 1;
      const InvU64 bucket{uniform_int_bucket(range), range + 1};
      do {
        i = div_u64(c.next(), bucket);
      } while (i > range && !c.overrun);
    }
    if (c.overrun) {
      fault(kFaultSerialDraws);
      return static_cast<u32>(c.pos - o);
    }
    const u64 sp = u64(P.start) + 1, ep = u64(P.end) - 1;
    const u32 rev = A.rev[i], fwd = A.fwd[i];
    const bool bound = A.ep[i] != kUnbound;
    if (bound && rev > sp && rev < ep && fwd > sp && fwd < ep) {
      double n1 = 0.0, n2 = 0.0;
      if (P.noisify) {
        n1 = gev_from_canonical(canonical_raw(c.next()), P.gev_mu, P.gev_sigma, P.gev_xi);
        n2 = gev_from_canonical(canonical_raw(c.next()), P.gev_mu, P.gev_sigma, P.gev_xi);
      }
      const double a = static_cast<double>(rev) - n1;
      const double b = static_cast<double>(fwd) + n2;
      const double p1 = b < a ? b : a;
      const double p2 = b < a ? a : b;
      const double sd = static_cast<double>(sp), ed = static_cast<double>(ep);
      if (p1 >= sd && p2 >= sd && p1 < ed && p2 < ed) {
        u64 x1 = static_cast<u64>(p1), x2 = static_cast<u64>(p2);
        if (kind == 1) {
          const u64 range = x2 - x1;
          u64 y1 = x1, y2 = x1;
          if (range != 0) {
            const InvU64 bucket{uniform_int_bucket(range), range + 1};
            u64 r;
            do {
              r = div_u64(c.next(), bucket);
            } while (r > range && !c.overrun);
            y1 = x1 + r;
            do {
              r = div_u64(c.next(), bucket);
            } while (r > range && !c.overrun);
            y2 = x1 + r;
          }
          x1 = y1;
          x2 = y2;
        }
        *b1 = static_cast<u32>((x1 - sp) / P.bin_size);
        *b2 = static_cast<u32>((x2 - sp) / P.bin_size);
      }
    }
    if (c.overrun) fault(kFaultSerialDraws);
    return static_cast<u32>(c.pos - o);
  }

  // Deterministic mode: the events of an epoch in the reference's order -- nloop loop contacts,
  // ntad TAD contacts, n1d 1D-occupancy events -- as one pool. Every event is evaluated at the
  // stream offset it has when all events before it consume their usual number of draws; the
  // first event that consumes a different number ends the round (events up to and including it
  // are final) and the rest is re-based in the next round. Two flag words take turns (a round's
  // flag is reset in the NEXT round's register region, after every thread has read it); the
  // caller sets S.tmp_u32[0] to all-ones before the barrier that precedes this call.
  MB_FN void sampling_pool(u32 nloop, u32 ntad, u32 n1d) {
    const u32 total = nloop + ntad + n1d;
    if (total == 0) return;
    const u32 sl = (S.num_active > 1 ? 1u : 0u) + (P.noisify ? 2u : 0u);  // loop / 1D event
    const u32 st = sl + 2;                                                  // TAD event
    // default draws of the events before pool index w
    auto draws_before = [&](u32 w) -> u64 {
      if (w <= nloop) return u64(w) * sl;
      if (w <= nloop + ntad) return u64(nloop) * sl + u64(w - nloop) * st;
      return u64(nloop) * sl + u64(ntad) * st + u64(w - nloop - ntad) * sl;
    };
    const u32 cap = ((P.n_lefs > P.n_bar ? P.n_lefs : P.n_bar) + 64) / 2;  // scratch: 2 words each
    const u32 max_by_window = (P.rng_window - 64) / (st + 1);
    u32 w0 = 0;
    for (u32 round = 0; w0 < total; ++round) {
      u32 batch = total - w0;
      if (batch > cap) batch = cap;
      if (batch > max_by_window) batch = max_by_window;
      const u64 d0 = draws_before(w0);
      const u64 base = S.rng_pos;
      const u64 limit = base + (draws_before(w0 + batch) - d0) + 48;
      rng_ensure(limit);
      u32* flag = &S.tmp_u32[(round & 1u) ? 2 : 0];
      u32* next_flag = &S.tmp_u32[(round & 1u) ? 0 : 2];
      MB_REGION(cta, tid) {
        for (u32 e = tid; e < batch; e += cta.nt()) {
          const u32 w = w0 + e;
          const int kind = w < nloop ? 0 : (w < nloop + ntad ? 1 : 2);
          u32 b1, b2;
          const u32 c = sampling_event(base + (draws_before(w) - d0), kind, limit, &b1, &b2);
          A.scratch[2 * e] = b1;
          A.scratch[2 * e + 1] = b2;
          if (c != (kind == 1 ? st : sl)) MB_ATOMIC_MIN_U32(flag, (e << 8) | (c > 255 ? 255u : c));
        }
      }
      cta.sync();
      const u32 exc = MB_SHARED_LOAD_U32(flag);
      const u32 valid = exc == 0xFFFFFFFFu ? batch : (exc >> 8) + 1;  // events final this round
      MB_REGION(cta, tid) {
        u32 registered = 0;
        for (u32 e = tid; e < valid; e += cta.nt()) {
          const u32 b1 = A.scratch[2 * e], b2 = A.scratch[2 * e + 1];
          if (b1 == kUnbound) continue;
          if (w0 + e >= nloop + ntad) {
            if (K.occ1d) {
              MB_ATOMIC_ADD_U64(K.occ1d + b1, u64(1));
              MB_ATOMIC_ADD_U64(K.occ1d + b2, u64(1));
            }
          } else {
            band_increment(b1, b2);
            ++registered;
          }
        }
        // (a 32-bit shared-memory reduction is one native instruction; the 64-bit form is a
        // compare-and-swap loop that 1,024 threads would fight over)
        if (registered) MB_ATOMIC_ADD_U32(&S.tmp_u32[7], registered);
        if (cta.leader(tid)) {
          u64 consumed = draws_before(w0 + valid) - d0;
          if (exc != 0xFFFFFFFFu) {
            if ((exc & 0xFF) == 255) fault(kFaultSerialDraws);
            consumed = (draws_before(w0 + valid - 1) - d0) + (exc & 0xFF);
          }
          S.rng_pos = base + consumed;
          MB_SHARED_STORE_U32(next_flag, 0xFFFFFFFFu);
        }
      }
      cta.sync();
      w0 += valid;
    }
  }

  // sample_and_register_contacts (register_contacts.cpp:93-120)
  MB_FN void sample_and_register_contacts() {
    u64 nev = P.contacts_per_epoch;
    if (!P.stop_on_epochs) {
      const u64 left = task.target_contacts - S.num_contacts;
      if (left < nev) nev = left;
    }
    if (nev == 0) return;
    const bool need_binomial = P.tad_to_loop != 0.0 && isfinite(P.tad_to_loop);
    if (need_binomial) rng_ensure(S.rng_pos + 256);  // reads S.rng_pos: before the barrier
    cta.sync();
    MB_REGION(cta, tid) {
      if (cta.leader(tid)) {
        u64 nloop;
        if (P.tad_to_loop == 0.0) {
          nloop = nev;
        } else if (!need_binomial) {
          nloop = 0;
        } else {
          Cursor c = kCtr ? ctr_cursor(S.epoch, kDrSplit, 0) : cursor(S.rng_pos, S.rng_pos + 256);
          nloop = static_cast<u64>(
              binomial_serial(c, static_cast<i64>(nev), 1.0 / (P.tad_to_loop + 1.0)));
          if (c.overrun) fault(kFaultSerialDraws);
          if constexpr (!kCtr) S.rng_pos = c.pos;
        }
        S.tmp_u32[1] = static_cast<u32>(nloop);
        S.tmp_u32[7] = 0;  // contacts registered this epoch
        S.tmp_u32[0] = 0xFFFFFFFFu;  // sampling_pool: no event off its usual draw count so far
      }
    }
    cta.sync();
    const u32 nloop = S.tmp_u32[1];
    const u32 ntad = static_cast<u32>(nev) - nloop;
    if constexpr (kCtr) {
      // all three kinds in one pass: an event's draws depend on (kind, number) only, so the
      // events of an epoch are one pool of independent work items
      const u32 n1d = P.track_1d ? static_cast<u32>(nev) : 0u;
      const u32 total = nloop + ntad + n1d;
      MB_REGION(cta, tid) {
        u32 registered = 0;
        for (u32 w = tid; w < total; w += cta.nt()) {
          const int kind = w < nloop ? 0 : (w < nloop + ntad ? 1 : 2);
          const u32 e = kind == 0 ? w : (kind == 1 ? w - nloop : w - nloop - ntad);
          u32 b1, b2;
          const u64 o = ctr_pack(S.epoch, kDrLoop + static_cast<u32>(kind), e);
          sampling_event(o, kind, o + kCtrDrawsPerItem, &b1, &b2);
          if (b1 == kUnbound) continue;
          if (kind == 2) {
            if (K.occ1d) {
              MB_ATOMIC_ADD_U64(K.occ1d + b1, u64(1));
              MB_ATOMIC_ADD_U64(K.occ1d + b2, u64(1));
            }
          } else {
            band_increment(b1, b2);
            ++registered;
          }
        }
        // (a 32-bit shared-memory reduction is one native instruction; the 64-bit form is a
        // compare-and-swap loop that 1,024 threads would fight over)
        if (registered) MB_ATOMIC_ADD_U32(&S.tmp_u32[7], registered);
      }
      cta.sync();
    } else {
      sampling_pool(nloop, ntad, P.track_1d ? static_cast<u32>(nev) : 0u);
    }
    MB_REGION(cta, tid) {
      if (cta.leader(tid)) S.num_contacts += S.tmp_u32[7];
    }
    cta.sync();
  }

  // ------------------------------------------------------------------------------ moves
  // std::round(x) for x >= 0 (half away from zero == half up): x - trunc(x) is exact.
  MB_FN static u32 round_nonneg_u32(double x) {
    const u32 t = static_cast<u32>(x);
    return t + ((x - static_cast<double>(t)) >= 0.5 ? 1u : 0u);
  }
  // generate_moves_helper (simulation.cpp:272-297): round(max(0, Normal(speed, sd)))
  MB_FN static u32 move_from_z(double z, double speed, double sd) {
    const double v = z * sd + speed;
    return round_nonneg_u32(v < 0.0 ? 0.0 : v);
  }
  // value of an accepted fast-path draw (no acceptance test)
  MB_FN double unit_normal_fast_value(u64 u) const {
    int bits;
    const double r = int_float_pair8(u, &bits);
    const double x = r * A.zig_nx[bits >> 1];
    return (bits & 1) ? x : -x;
  }

  // Bitmap of the draws that leave the fast ziggurat path (deterministic mode; one bit per
  // examined stream offset, at most 2.25 n_lefs + 64 of them). It borrows words [4, 12) * (n_lefs
  // / 32 + 3) of A.bits, which only the secondary pass uses otherwise: init_cell and -- after the
  // secondary pass of every epoch -- extrude_and_release clear it, so draw_normal_moves finds it
  // zero (bind, rank and the contact sampling in between leave those words alone).
  MB_FN u32* mv_slow_bits() const { return A.bits + 4 * (P.n_lefs / 32 + 3); }
  MB_FN u32 mv_slow_words() const {
    const u32 items = 2 * P.n_lefs;
    return (items + items / 8 + 64 + 31) / 32;
  }
  MB_FN void mv_clear_slow_bits(int tid) const {
    if constexpr (!kCtr) {
      u32* w = mv_slow_bits();
      const u32 nw = mv_slow_words();
      for (u32 k = static_cast<u32>(tid); k < nw; k += static_cast<u32>(cta.nt())) w[k] = 0;
    }
  }

  // One Normal(speed, sd) per item, items in stream order: the first n_rev items are the rev
  // moves of LEFs 0..n_rev-1, the others the fwd moves (generate_moves, simulation.cpp:299-330,
  // draws all rev moves and then all fwd moves). The fast ziggurat path uses exactly one draw;
  // the rare slow paths are evaluated speculatively, every one at its own offset, and stitched
  // into the stream (which of them start an item, by how much each shifts the items behind it).
  MB_FN void draw_normal_moves(u32 items, u32 n_rev, double rev_speed, double fwd_speed) {
    if constexpr (kCtr) {
      // One ziggurat sampler per item on the item's own sequence. Some 98 % of the items are
      // settled by their first draw; the others (rejection loops, exp()) are noted per thread and
      // finished after the thread's fast items, so that the lanes of a warp that have any go
      // through the slow code together instead of one or two at a time on every trip.
      sub_begin();
      const double rsd = P.rev_std, fsd = P.fwd_std;
      MB_REGION(cta, tid) {
        bool over = false;
        auto store = [&](u32 item, double z) {
          u32 mv;
          if (item < n_rev) {
            A.rm[item] = mv = move_from_z(z, rev_speed, rsd);
          } else {
            A.fm[item - n_rev] = mv = move_from_z(z, fwd_speed, fsd);
          }
          over |= mv > P.move_bound;
        };
        for (u32 i0 = tid; i0 < items; i0 += 32u * cta.nt()) {
          u32 later = 0, k = 0;
          for (u32 i = i0; k < 32 && i < items; i += cta.nt(), ++k) {
            double z;
            if (unit_normal_fast(raw(ctr_pack(S.epoch, kDrMoves, i)), &z)) {
              store(i, z);
            } else {
              later |= 1u << k;
            }
          }
          while (later) {
            k = static_cast<u32>(MB_FFS(later)) - 1;
            later &= later - 1;
            const u32 i = i0 + k * cta.nt();
            Cursor c = ctr_cursor(S.epoch, kDrMoves, i);
            const double z = unit_normal_serial(c);
            if (c.overrun) fault(kFaultSerialDraws);
            store(i, z);
          }
        }
        if (over) S.move_bound_hit = 1;
      }
      cta.sync();
      sub_lap(kPhMvFinal);
      return;
    }
    const u32 slack = items / 8 + 64;
    const u32 span = items + slack;  // offsets examined
    const u64 base = S.rng_pos;
    const u64 limit = base + span + 192;
    sub_begin();
    rng_ensure(limit);
    sub_lap(kPhMvEnsure);
    // exception records live in scratch: 4 words each {offset, draws consumed, z lo, z hi}
    u32* ex = A.scratch;
    const u32 ex_cap = ((P.n_lefs > P.n_bar ? P.n_lefs : P.n_bar) + 64) / 4;
    // Which offsets leave the fast path: consecutive threads test consecutive draws (a warp reads
    // whole lines of the ring) and set a bit per slow offset; the bitmap -- zero on entry, see
    // mv_slow_bits() -- is then turned into the ordered list by word.
    u32* slowbits = mv_slow_bits();
    const u32 slow_words = (span + 31) / 32;
    MB_REGION(cta, tid) {
      const u32 nt = static_cast<u32>(cta.nt());
      for (u32 o = static_cast<u32>(tid); o < span; o += 4 * nt) {
        // (offsets past span are read but not used: inside the staged range)
        const u64 r0 = raw(base + o), r1 = raw(base + o + nt), r2 = raw(base + o + 2 * nt),
                  r3 = raw(base + o + 3 * nt);
        auto test = [&](u32 q, u64 r) {
          double z;
          if (q < span && !unit_normal_fast(r, &z))
            MB_ATOMIC_OR_U32(&slowbits[q >> 5], 1u << (q & 31));
        };
        test(o, r0);
        test(o + nt, r1);
        test(o + 2 * nt, r2);
        test(o + 3 * nt, r3);
      }
    }
    cta.sync();
    PerThread<u64> cnt(cta.nt());
    MB_REGION(cta, tid) {
      u32 lo, hi;
      chunk(tid, slow_words, &lo, &hi);
      u32 c = 0;
      for (u32 w = lo; w < hi; ++w) c += static_cast<u32>(MB_POPC(slowbits[w]));
      cnt[tid] = c;
    }
    const u64 n_exc = cta.exscan_sum(cnt);
    sub_lap(kPhMvScan);
    if (n_exc > ex_cap) {
      MB_REGION(cta, tid) {
        if (cta.leader(tid)) fault(kFaultSerialDraws);
      }
      cta.sync();
      return;
    }
    MB_REGION(cta, tid) {
      u32 lo, hi;
      chunk(tid, slow_words, &lo, &hi);
      u32 j = static_cast<u32>(cnt[tid]);
      for (u32 w = lo; w < hi; ++w) {
        u32 m = slowbits[w];
        while (m) {
          const u32 k = static_cast<u32>(MB_FFS(m)) - 1;
          m &= m - 1;
          ex[4 * j++] = 32 * w + k;
        }
      }
    }
    cta.sync();
    MB_REGION(cta, tid) {
      for (u32 j = tid; j < n_exc; j += cta.nt()) {
        Cursor c = cursor(base + ex[4 * j], limit);
        const double z = unit_normal_serial(c);
        ex[4 * j + 1] = static_cast<u32>(c.pos - (base + ex[4 * j]));
        u64 zb;
#if MB_DEVICE_BUILD
        zb = static_cast<u64>(__double_as_longlong(z));
#else
        std::memcpy(&zb, &z, 8);
#endif
        ex[4 * j + 2] = static_cast<u32>(zb);
        ex[4 * j + 3] = static_cast<u32>(zb >> 32);
        if (c.overrun) fault(kFaultSerialDraws);
      }
    }
    cta.sync();
    // Keep the exceptions that start an item (not swallowed by an earlier slow path) and turn
    // their offsets into item indices: record j becomes {item, cumulative extra draws, z}.
    if (n_exc <= static_cast<u64>(cta.nt())) {
      // Parallel form. An exception nothing before it can reach is certainly kept and cuts the
      // list into independent clusters; clusters (1-3 records) are resolved by their heads.
      MB_REGION(cta, tid) {
        const u32 j = static_cast<u32>(tid);
        if (j < n_exc) {
          const u32 off = ex[4 * j];
          bool head = true;
          for (u32 i = j; i-- > 0;) {
            const u32 oi = ex[4 * i];
            if (off - oi >= 256u) break;  // a slow path never takes that many draws (cursor limit)
            // (the owner of record i may be setting its head flag, bit 31, right now)
            if (oi + (MB_SHARED_LOAD_U32(&ex[4 * i + 1]) & 0xFFFFu) > off) {
              head = false;
              break;
            }
          }
          if (head) MB_SHARED_OR_U32(&ex[4 * j + 1], 0x80000000u);
        }
        if (cta.leader(tid)) {
          S.tmp_u32[2] = 0;
          S.tmp_u32[3] = 0;
        }
      }
      cta.sync();
      MB_REGION(cta, tid) {
        const u32 j = static_cast<u32>(tid);
        // (a cluster's walk ends at the next head, whose owner is setting bit 30 of that word)
        if (j < n_exc && (MB_SHARED_LOAD_U32(&ex[4 * j + 1]) & 0x80000000u)) {
          u32 covered = ex[4 * j] + (MB_SHARED_LOAD_U32(&ex[4 * j + 1]) & 0xFFFFu);
          MB_SHARED_OR_U32(&ex[4 * j + 1], 0x40000000u);  // kept
          for (u32 k = j + 1; k < n_exc && !(MB_SHARED_LOAD_U32(&ex[4 * k + 1]) & 0x80000000u);
               ++k) {
            if (ex[4 * k] >= covered) {
              MB_SHARED_OR_U32(&ex[4 * k + 1], 0x40000000u);
              covered = ex[4 * k] + (MB_SHARED_LOAD_U32(&ex[4 * k + 1]) & 0xFFFFu);
            }
          }
        }
      }
      cta.sync();
      PerThread<u64> ks(cta.nt());  // low word: kept records, high word: extra draws
      MB_REGION(cta, tid) {
        const u32 j = static_cast<u32>(tid);
        u64 v = 0;
        if (j < n_exc && (ex[4 * j + 1] & 0x40000000u))
          v = u64(1) | (u64((ex[4 * j + 1] & 0xFFFFu) - 1) << 32);
        ks[tid] = v;
      }
      cta.exscan_sum(ks);
      PerThread<u64> st_a(cta.nt()), st_z(cta.nt());
      PerThread<u32> st_dest(cta.nt());
      MB_REGION(cta, tid) {
        const u32 j = static_cast<u32>(tid);
        u32 dest = 0xFFFFFFFFu;
        if (j < n_exc && (ex[4 * j + 1] & 0x40000000u)) {
          const u32 shift_before = static_cast<u32>(ks[tid] >> 32);
          const u32 item = ex[4 * j] - shift_before;
          if (item < items) {
            const u32 shift_after = shift_before + (ex[4 * j + 1] & 0xFFFFu) - 1;
            dest = static_cast<u32>(ks[tid] & 0xFFFFFFFFu);
            st_a[tid] = u64(item) | (u64(shift_after) << 32);
            st_z[tid] = u64(ex[4 * j + 2]) | (u64(ex[4 * j + 3]) << 32);
            MB_ATOMIC_MAX_U32(&S.tmp_u32[2], dest + 1);
            MB_ATOMIC_MAX_U32(&S.tmp_u32[3], shift_after);
          }
        }
        st_dest[tid] = dest;
      }
      cta.sync();
      MB_REGION(cta, tid) {
        const u32 dest = st_dest[tid];
        if (dest != 0xFFFFFFFFu) {
          ex[4 * dest] = static_cast<u32>(st_a[tid] & 0xFFFFFFFFu);
          ex[4 * dest + 1] = static_cast<u32>(st_a[tid] >> 32);
          ex[4 * dest + 2] = static_cast<u32>(st_z[tid] & 0xFFFFFFFFu);
          ex[4 * dest + 3] = static_cast<u32>(st_z[tid] >> 32);
        }
        if (cta.leader(tid) && S.tmp_u32[3] > slack) fault(kFaultRngWindow);
      }
      cta.sync();
    } else {
      MB_REGION(cta, tid) {
        if (cta.leader(tid)) {
          u32 covered = 0, shift = 0, kept = 0;
          for (u32 j = 0; j < n_exc; ++j) {
            const u32 off = ex[4 * j];
            if (off < covered) continue;
            const u32 item = off - shift;
            if (item >= items) break;
            const u32 c = ex[4 * j + 1];
            covered = off + c;
            shift += c - 1;
            ex[4 * kept] = item;
            ex[4 * kept + 1] = shift;
            ex[4 * kept + 2] = ex[4 * j + 2];
            ex[4 * kept + 3] = ex[4 * j + 3];
            ++kept;
          }
          S.tmp_u32[2] = kept;
          S.tmp_u32[3] = shift;
          if (shift > slack) fault(kFaultRngWindow);
        }
      }
      cta.sync();
    }
    const u32 kept = S.tmp_u32[2];
    sub_lap(kPhMvExceptions);
    const double rev_sd = P.rev_std, fwd_sd = P.fwd_std;
    // Item i reads the draw at offset i + (extra draws of the kept exceptions before it). Groups
    // of 32 consecutive threads (a warp) own consecutive ranges of items and go through them 32
    // items at a time, lane by lane: neighbouring lanes read neighbouring draws of the ring, and
    // every lane of a group walks the (short, shared) list of exceptions of the 32 items in
    // step, so the trip count is the same for all lanes whatever the exceptions are.
    MB_REGION(cta, tid) {
      const u32 nt = static_cast<u32>(cta.nt());
      const u32 G = (nt & 31u) == 0 ? 32u : 1u;  // (odd widths exist only in the emulation)
      const u32 ng = nt / G, g = static_cast<u32>(tid) / G, lane = static_cast<u32>(tid) % G;
      const u32 lo = static_cast<u32>(u64(items) * g / ng);
      const u32 hi = static_cast<u32>(u64(items) * (g + 1) / ng);
      // first kept exception with item >= lo
      u32 a = 0, b = kept;
      while (a < b) {
        const u32 m = (a + b) >> 1;
        if (ex[4 * m] < lo) {
          a = m + 1;
        } else {
          b = m;
        }
      }
      u32 e = a;
      u32 shift = e ? ex[4 * (e - 1) + 1] : 0;
      bool over = false;
      auto store = [&](u32 item, double z) {
        u32 mv;
        if (item < n_rev) {
          A.rm[item] = mv = move_from_z(z, rev_speed, rev_sd);
        } else {
          A.fm[item - n_rev] = mv = move_from_z(z, fwd_speed, fwd_sd);
        }
        over |= mv > P.move_bound;
      };
      // the draw an item reads when no exception lies between the start of its 32 items and it
      // -- nearly always -- is requested one trip ahead
      u64 r_ahead = lo < hi ? raw(base + shift + lo + lane) : 0;
      for (u32 i0 = lo; i0 < hi; i0 += G) {
        const u32 i = i0 + lane;
        const u32 end = i0 + G < hi ? i0 + G : hi;
        const u64 r_spec = r_ahead;
        const u32 shift_spec = shift;
        u32 my_shift = shift, my_exc = 0xFFFFFFFFu;
        while (e < kept && ex[4 * e] < end) {  // same trips for every lane of the group
          const u32 xi = ex[4 * e];
          if (xi < i) my_shift = ex[4 * e + 1];
          if (xi == i) my_exc = e;
          shift = ex[4 * e + 1];  // later items start after this item's extra draws
          ++e;
        }
        if (end < hi) r_ahead = raw(base + shift + end + lane);
        if (i >= end) continue;
        if (my_exc != 0xFFFFFFFFu) {
          const u64 zb = u64(ex[4 * my_exc + 2]) | (u64(ex[4 * my_exc + 3]) << 32);
          double z;
#if MB_DEVICE_BUILD
          z = __longlong_as_double(static_cast<long long>(zb));
#else
          std::memcpy(&z, &zb, 8);
#endif
          store(i, z);
        } else {
          const u64 r = my_shift == shift_spec ? r_spec : raw(base + my_shift + i);
          store(i, unit_normal_fast_value(r));
        }
      }
      if (over) S.move_bound_hit = 1;
    }
    cta.sync();
    MB_REGION(cta, tid) {
      if (cta.leader(tid)) S.rng_pos = base + items + S.tmp_u32[3];
    }
    cta.sync();
    sub_lap(kPhMvFinal);
  }

  // adjust_moves_of_consecutive_extr_units (simulation.cpp:350-407) + clamp_moves (:332-347).
  // The two neighbour recurrences are min-plus prefix scans; the handful of units close enough
  // to an interval end for the reference's "skip" rule to fire are finished serially.
  // Hinted forms for callers that query increasing thresholds: `from` is a lower bound of the
  // answer (the previous answer); a few linear steps, then binary search on what is left.
  MB_FN u32 count_rev_le_from(u64 thr, u32 from) const {
    u32 a = from;
    const u32 n = S.num_active;
    for (int step = 0; step < 4; ++step) {
      if (a >= n || u64(A.rev[A.rr[a]]) > thr) return a;
      ++a;
    }
    u32 b = n;
    while (a < b) {
      const u32 m = (a + b) >> 1;
      if (u64(A.rev[A.rr[m]]) <= thr) {
        a = m + 1;
      } else {
        b = m;
      }
    }
    return a;
  }
  MB_FN u32 count_fwd_lt_from(u64 thr, u32 from, u32 limit) const {  // over ranks [0, limit)
    u32 a = from;
    for (int step = 0; step < 4; ++step) {
      if (a >= limit || u64(A.fwd[A.fr[a]]) >= thr) return a;
      ++a;
    }
    u32 b = limit;
    while (a < b) {
      const u32 m = (a + b) >> 1;
      if (u64(A.fwd[A.fr[m]]) < thr) {
        a = m + 1;
      } else {
        b = m;
      }
    }
    return a;
  }
  // The same count for a caller whose previous answer `from` is close: doubling steps bracket the
  // answer and a binary search finishes inside the bracket, so the trip count follows the
  // distance (a warp runs as many trips as its farthest lane needs), not the array length.
  MB_FN u32 count_fwd_lt_near(u64 thr, u32 from, u32 limit) const {
    u32 a = from;
    if (a >= limit || u64(A.fwd[A.fr[a]]) >= thr) return a;
    // invariant: rank a is below thr
    u32 step = 1;
    while (a + step < limit && u64(A.fwd[A.fr[a + step]]) < thr) {
      a += step;
      step <<= 1;
    }
    u32 b = a + step < limit ? a + step : limit;  // rank b is not below thr (or b == limit)
    ++a;
    while (a < b) {
      const u32 m = (a + b) >> 1;
      if (u64(A.fwd[A.fr[m]]) < thr) {
        a = m + 1;
      } else {
        b = m;
      }
    }
    return a;
  }
  MB_FN u32 count_fwd_lt_from_top(u64 thr) const {  // number of fwd ranks with pos < thr
    u32 c = S.num_active;
    for (int step = 0; step < 4; ++step) {
      if (c == 0 || u64(A.fwd[A.fr[c - 1]]) < thr) return c;
      --c;
    }
    u32 a = 0, b = c;
    while (a < b) {
      const u32 m = (a + b) >> 1;
      if (u64(A.fwd[A.fr[m]]) < thr) {
        a = m + 1;
      } else {
        b = m;
      }
    }
    return a;
  }
  MB_FN u32 count_rev_le(u64 thr) const {  // number of rev ranks with pos <= thr (binary search)
    u32 a = 0, b = S.num_active;
    while (a < b) {
      const u32 m = (a + b) >> 1;
      if (u64(A.rev[A.rr[m]]) <= thr) {
        a = m + 1;
      } else {
        b = m;
      }
    }
    return a;
  }
  MB_FN u32 count_fwd_lt(u64 thr) const {  // number of fwd ranks with pos < thr
    u32 a = 0, b = S.num_active;
    while (a < b) {
      const u32 m = (a + b) >> 1;
      if (u64(A.fwd[A.fr[m]]) < thr) {
        a = m + 1;
      } else {
        b = m;
      }
    }
    return a;
  }

  MB_FN void adjust_and_clamp_moves() {
    const u32 n = S.num_active;
    // Upper bound of every move: the static bound the generator checked its output against
    // (S.move_bound_hit is set when a draw exceeded it), else the true maximum.
    u64 mmax = P.move_bound;
    if (S.move_bound_hit) {
      PerThread<u64> mx(cta.nt());
      MB_REGION(cta, tid) {
        u64 m = 0;
        for (u32 i = tid; i < n; i += cta.nt()) {
          const u64 a = A.rm[i], b = A.fm[i];
          m = a > m ? a : m;
          m = b > m ? b : m;
        }
        mx[tid] = m;
      }
      mmax = cta.reduce_max(mx);
      MB_REGION(cta, tid) {
        if (cta.leader(tid)) {
          S.move_bound_hit = 0;
          // the secondary pass parks moves in 24-bit collision indices
          if (mmax >= (u64(1) << 24)) fault(kFaultMoveRange);
        }
      }
      cta.sync();
    }
    if (n >= 2) {
      // rev units, walked 3'->5': q'[k-1] = min(q[k-1], q'[k] - 1); fwd units, walked 5'->3':
      // q'[k] = max(q[k], q'[k-1] + 1) (negated: min-plus). Both scans run together.
      // (every thread needs both counts; few units sit that close to an interval end, so a few
      // linear steps from the end in question almost always settle them)
      const u32 k_near = count_rev_le_from(u64(P.start) + mmax + n, 0);  // rev ranks < k_near: serial part
      const u32 M = n - k_near;                                  // rev ranks [k_near, n) in parallel
      const u64 far_thr = u64(P.end) - 1 > mmax + n ? u64(P.end) - 1 - mmax - n : 0;
      const u32 k_far = count_fwd_lt_from_top(far_thr);  // fwd ranks [0, k_far) in parallel
      // With x[m] the end position of the unit at scan position m, the recurrence
      // x[m] = min(q[m], x[m-1] - 1) unrolls to x[m] = min_{j <= m}(q[j] + j) - m: a plain
      // prefix minimum of q[j] + j (positions and moves fit 32 bits with room to spare).
      PerThread<i32> f(cta.nt()), g(cta.nt());
      MB_REGION(cta, tid) {
        u32 lo, hi;
        chunk(tid, M, &lo, &hi);
        i32 acc = 0x7FFFFFFF;
        for (u32 m = lo; m < hi; ++m) {
          const u32 idx = A.rr[n - 1 - m];
          const i32 r = static_cast<i32>(A.rev[idx]) - static_cast<i32>(A.rm[idx]) +
                        static_cast<i32>(m);
          acc = r < acc ? r : acc;
        }
        f[tid] = acc;
        chunk(tid, k_far, &lo, &hi);
        acc = 0x7FFFFFFF;
        for (u32 k = lo; k < hi; ++k) {
          const u32 idx = A.fr[k];
          const i32 r = static_cast<i32>(k) -
                        (static_cast<i32>(A.fwd[idx]) + static_cast<i32>(A.fm[idx]));
          acc = r < acc ? r : acc;
        }
        g[tid] = acc;
      }
      cta.exscan_min2_i32(f, g);
      MB_REGION(cta, tid) {
        u32 lo, hi;
        chunk(tid, M, &lo, &hi);
        i32 run = f[tid];
        for (u32 m = lo; m < hi; ++m) {
          const u32 idx = A.rr[n - 1 - m];
          const i32 rev = static_cast<i32>(A.rev[idx]);
          const i32 r = rev - static_cast<i32>(A.rm[idx]) + static_cast<i32>(m);
          run = r < run ? r : run;
          A.rm[idx] = static_cast<u32>(rev - (run - static_cast<i32>(m)));
        }
        chunk(tid, k_far, &lo, &hi);
        run = g[tid];
        for (u32 k = lo; k < hi; ++k) {
          const u32 idx = A.fr[k];
          const i32 fwd = static_cast<i32>(A.fwd[idx]);
          const i32 r = static_cast<i32>(k) - (fwd + static_cast<i32>(A.fm[idx]));
          run = r < run ? r : run;
          // end position = -(run - k)
          A.fm[idx] = static_cast<u32>(static_cast<i32>(k) - run - fwd);
        }
      }
      cta.sync();
      // the few units close enough to an interval end for the reference's "skip" rule to fire:
      // thread 0 finishes the rev side, thread 1 (when there is one) the fwd side
      MB_REGION(cta, tid) {
        const bool two = cta.nt() > 1;
        if (tid == 0) {
          for (u32 i = (k_near < n - 1 ? k_near : n - 1); i > 0; --i) {
            const u32 i1 = A.rr[i - 1], i2 = A.rr[i];
            if (u64(A.rev[i1]) <= u64(P.start) + A.rm[i1] ||
                u64(A.rev[i2]) <= u64(P.start) + A.rm[i2])
              continue;
            const u32 p1 = A.rev[i1] - A.rm[i1];
            const u32 p2 = A.rev[i2] - A.rm[i2];
            if (p2 <= p1) A.rm[i1] += (p1 - p2) + 1;
          }
        }
        if (tid == (two ? 1 : 0)) {
          for (u32 i = (k_far > 1 ? k_far : 1); i < n; ++i) {
            const u32 i1 = A.fr[i - 1], i2 = A.fr[i];
            if (u64(A.fwd[i1]) + A.fm[i1] > u64(P.end) - 1 ||
                u64(A.fwd[i2]) + A.fm[i2] > u64(P.end) - 1)
              continue;
            const u32 p1 = A.fwd[i1] + A.fm[i1];
            const u32 p2 = A.fwd[i2] + A.fm[i2];
            if (p1 >= p2) A.fm[i2] += (p1 - p2) + 1;
          }
        }
      }
      cta.sync();
    }
    MB_REGION(cta, tid) {
      for (u32 i = tid; i < n; i += cta.nt()) {
        const u32 rcap = A.rev[i] - P.start;
        const u32 fcap = P.end - A.fwd[i] - 1;
        if (A.rm[i] > rcap) A.rm[i] = rcap;
        if (A.fm[i] > fcap) A.fm[i] = fcap;
      }
    }
    cta.sync();
  }

  // generate_moves (simulation.cpp:299-330)
  MB_FN void generate_moves() {
    const bool done = S.burnin_completed != 0;
    const double rs = done ? P.rev_speed : P.rev_speed_burnin;
    const double fs = done ? P.fwd_speed : P.fwd_speed_burnin;
    const u32 n = S.num_active;
    const bool draw_r = P.rev_std != 0.0, draw_f = P.fwd_std != 0.0;
    if (!draw_r || !draw_f) {  // constant moves take no draws
      const u32 rmi = round_nonneg_u32(rs < 0.0 ? 0.0 : rs), fmi = round_nonneg_u32(fs < 0.0 ? 0.0 : fs);
      MB_REGION(cta, tid) {
        for (u32 i = tid; i < n; i += cta.nt()) {
          if (!draw_r) A.rm[i] = rmi;
          if (!draw_f) A.fm[i] = fmi;
        }
      }
      cta.sync();
    }
    const u32 n_rev = draw_r ? n : 0;
    const u32 items = n_rev + (draw_f ? n : 0);
    if (items) draw_normal_moves(items, n_rev, rs, fs);
    lap(kPhMovesGen);
    adjust_and_clamp_moves();
    lap(kPhMovesAdjust);
  }

  // ExtrusionBarriers::next_state (extrusion_barriers.cpp:145-161): one canonical per barrier
  MB_FN void next_barrier_states() {
    if (P.n_bar == 0) return;
    rng_ensure(S.rng_pos + P.n_bar);
    const u64 base = kCtr ? ctr_pack(S.epoch, kDrBarriers, 0) : S.rng_pos;
    MB_REGION(cta, tid) {
      for (u32 i = tid; i < P.n_bar; i += cta.nt()) {
        const double u = canonical_raw(raw(kCtr ? base + (u64(i) << 8) : base + i));
        const bool act = A.bar_active[i] != 0;
        if (!act && u > D.bar_stp_inactive[i]) {
          A.bar_active[i] = 1;
        } else if (act && u > D.bar_stp_active[i]) {
          A.bar_active[i] = 0;
        }
      }
    }
    cta.sync();  // every thread has read `base` before the leader moves the stream position
    if constexpr (!kCtr) {
      MB_REGION(cta, tid) {
        if (cta.leader(tid)) S.rng_pos = base + P.n_bar;
      }
      cta.sync();
    }
  }

  MB_FN bool bar_blocks_rev(u32 b) const { return (D.bar_dir_rev[b >> 5] >> (b & 31)) & 1u; }

  // ------------------------------------------------------------------------------ collisions
  // detect_units_at_interval_boundaries (simulation_detect_collisions.cpp:25-120); leader only.
  MB_FN void detect_boundaries_leader() {
    const u32 n = S.num_active;
    u32 n5 = 0, n3 = 0;
    const u32 first_fwd_pos = A.fwd[A.fr[0]];
    const u32 last_rev_pos = A.rev[A.rr[n - 1]];
    for (u32 i = 0; i < n; ++i) {
      const u32 idx = A.rr[i];
      const u32 pos = A.rev[idx];
      if (pos == P.start) {
        ++n5;
        A.rc[idx] = coll_make(5, kEvCollision | kEvChromBoundary);
      } else if (pos > first_fwd_pos) {
        break;
      } else if (pos - A.rm[idx] == P.start) {
        A.rc[idx] = coll_make(5, kEvCollision | kEvChromBoundary);
        ++n5;
        break;
      }
    }
    for (u32 i = n - 1; i > 0; --i) {
      const u32 idx = A.fr[i];
      const u32 pos = A.fwd[idx];
      if (pos == P.end - 1) {
        ++n3;
        A.fc[idx] = coll_make(3, kEvCollision | kEvChromBoundary);
      } else if (pos < last_rev_pos) {
        break;
      } else if (pos + A.fm[idx] == P.end - 1) {
        A.fc[idx] = coll_make(3, kEvCollision | kEvChromBoundary);
        ++n3;
        break;
      }
    }
    S.n5 = n5;
    S.n3 = n3;
  }

  // detect_lef_bar_collisions (simulation_detect_collisions.cpp:123-247). Every active barrier
  // tests exactly one unit: the first rev unit downstream of it / the last fwd unit upstream of
  // it; the closest successful barrier wins. Bernoulli draws (fractional pblock only) are taken
  // in barrier order, rev pass (ascending) then fwd pass (descending).
  // *hint: answer of the previous (smaller or equal) barrier position handled by this thread,
  // 0xFFFFFFFF = none yet
  MB_FN bool lef_bar_candidate_rev(u32 b, u32 j0, u32* unit, u32* hint) const {
    const u32 n = S.num_active;
    const u32 bp = A.bar_pos[b];
    u32 j = *hint == 0xFFFFFFFFu ? count_rev_le(bp) : count_rev_le_from(bp, *hint);
    *hint = j;  // first rank with pos > bp
    if (j < j0) j = j0;
    if (j >= n) return false;
    const u32 idx = A.rr[j];
    const u32 pos = A.rev[idx];
    if (pos <= bp) return false;
    *unit = idx;
    return pos - bp <= A.rm[idx];
  }
  MB_FN bool lef_bar_candidate_fwd(u32 b, u32 jend, u32* unit, u32* hint) const {
    const u32 bp = A.bar_pos[b];
    const u32 c = *hint == 0xFFFFFFFFu ? count_fwd_lt(bp)
                                       : count_fwd_lt_from(bp, *hint, S.num_active);
    *hint = c;  // ranks [0, c) have pos < bp
    if (c == 0) return false;
    u32 j = c - 1;
    if (j > jend) j = jend;
    const u32 idx = A.fr[j];
    const u32 pos = A.fwd[idx];
    if (pos >= bp) return false;
    *unit = idx;
    return bp - pos <= A.fm[idx];
  }

  MB_FN void detect_lef_bar_collisions() {
    const u32 n = S.num_active, nb = P.n_bar;
    if (nb == 0) return;
    const u32 j0 = S.n5 ? S.n5 - 1 : 0;
    const u32 sat3 = S.n3 ? S.n3 - 1 : 0;
    const u32 jend = n - sat3 - 1;
    const double pmaj = P.pblock_major, pmin = P.pblock_minor;
    const bool frac_maj = pmaj != 0.0 && pmaj != 1.0;
    const bool frac_min = pmin != 0.0 && pmin != 1.0;
    const u32 tmp_ev = kEvTmp | kEvCollision | kEvLefBar;
    if (!frac_maj && !frac_min) {
      // No trial needs a draw, so the search can be turned around: instead of two binary
      // searches over the (doubly indirect) rank orders per barrier, every thread walks its
      // contiguous share of the rev ranks and of the fwd ranks together with the sorted barrier
      // positions -- the number of barriers below a unit comes from a look-up table over
      // position buckets plus a step or two (or, when the table found no room in shared memory,
      // from one binary search for the first unit of the share and a forward-only cursor). The
      // barriers that test rev rank k are those with
      // pos[k-1] <= bp < pos[k] (all below pos[k] for k == j0); the closest one that is active,
      // blocks this direction with certainty and lies within the unit's move wins, exactly what
      // the per-barrier atomicMax selected. A hit overrides a boundary mark, as before.
      const bool rev_hit_maj = pmaj == 1.0, rev_hit_min = pmin == 1.0;
      const bool use_lut = P.lut_entries != 0;
      MB_REGION(cta, tid) {
        u32 lo, hi;
        if (j0 < n && (rev_hit_maj || rev_hit_min)) {
          chunk(tid, n - j0, &lo, &hi);
          u32 b = 0;  // number of barriers with bar_pos < pos of the current unit
          bool first = true;
          // position of the unit one rank down: carried along instead of re-read through rr
          u32 below = (lo < hi && lo > 0) ? A.rev[A.rr[j0 + lo - 1]] : 0u;
          for (u32 k = j0 + lo; k < j0 + hi; ++k) {
            const u32 idx = A.rr[k];
            const u32 pos = A.rev[idx];
            const u32 lower = below;
            below = pos;
            if (use_lut) {  // a bucket holds a barrier or two: no cursor, no search
              b = A.bar_lut[(pos - P.start) >> P.lut_shift];
              while (b < nb && A.bar_pos[b] < pos) ++b;
            } else if (first) {
              u32 a = 0, z = nb;
              while (a < z) {
                const u32 mid = (a + z) >> 1;
                if (A.bar_pos[mid] < pos) {
                  a = mid + 1;
                } else {
                  z = mid;
                }
              }
              b = a;
              first = false;
            } else {
              while (b < nb && A.bar_pos[b] < pos) ++b;
            }
            if (b == 0) continue;
            const u32 mv = A.rm[idx];
            for (u32 t = b; t-- > 0;) {
              const u32 bp = A.bar_pos[t];
              if (bp < lower || pos - bp > mv) break;
              if (!A.bar_active[t]) continue;
              if (bar_blocks_rev(t) ? rev_hit_maj : rev_hit_min) {
                A.rc[idx] = coll_make(t, kEvCollision | kEvLefBar);
                break;
              }
            }
          }
        }
        if (rev_hit_maj || rev_hit_min) {  // fwd units: major blocks when the barrier does NOT block rev
          chunk(tid, jend + 1, &lo, &hi);
          u32 b = 0;  // number of barriers with bar_pos <= pos of the current unit
          bool first = true;
          // the unit one rank up is read one trip ahead (its position bounds this unit's search)
          u32 idx_up = lo < hi ? A.fr[lo] : 0u;
          u32 pos_up = lo < hi ? A.fwd[idx_up] : 0u;
          for (u32 k = lo; k < hi; ++k) {
            const u32 idx = idx_up;
            const u32 pos = pos_up;
            if (k < jend) {
              idx_up = A.fr[k + 1];
              pos_up = A.fwd[idx_up];
            }
            const u32 upper = k < jend ? pos_up : 0xFFFFFFFFu;
            if (use_lut) {
              b = A.bar_lut[(pos - P.start) >> P.lut_shift];
              while (b < nb && A.bar_pos[b] <= pos) ++b;
            } else if (first) {
              u32 a = 0, z = nb;
              while (a < z) {
                const u32 mid = (a + z) >> 1;
                if (A.bar_pos[mid] <= pos) {
                  a = mid + 1;
                } else {
                  z = mid;
                }
              }
              b = a;
              first = false;
            } else {
              while (b < nb && A.bar_pos[b] <= pos) ++b;
            }
            if (b == nb) break;  // no barrier downstream of this or any later unit
            const u32 mv = A.fm[idx];
            for (u32 t = b; t < nb; ++t) {
              const u32 bp = A.bar_pos[t];
              if (bp > upper || bp - pos > mv) break;
              if (!A.bar_active[t]) continue;
              if (bar_blocks_rev(t) ? rev_hit_min : rev_hit_maj) {
                A.fc[idx] = coll_make(t, kEvCollision | kEvLefBar);
                break;
              }
            }
          }
        }
      }
      cta.sync();
      return;
    } else if constexpr (kCtr) {
      // fractional pblock: trial (barrier b, direction) reads draw 2b / 2b+1 of this epoch
      MB_REGION(cta, tid) {
        for (u32 b = tid; b < nb; b += cta.nt()) {
          if (!A.bar_active[b]) continue;
          const bool brev = bar_blocks_rev(b);
          const double pr = brev ? pmaj : pmin, pf = brev ? pmin : pmaj;
          u32 unit;
          u32 nohint_r = 0xFFFFFFFFu, nohint_f = 0xFFFFFFFFu;
          if (pr != 0.0 && lef_bar_candidate_rev(b, j0, &unit, &nohint_r) &&
              (pr == 1.0 || bernoulli_raw(raw(ctr_pack(S.epoch, kDrLefBar, 2 * b)), pr)))
            MB_ATOMIC_MAX_U32(&A.rc[unit], coll_make(b, tmp_ev));
          if (pf != 0.0 && lef_bar_candidate_fwd(b, jend, &unit, &nohint_f) &&
              (pf == 1.0 || bernoulli_raw(raw(ctr_pack(S.epoch, kDrLefBar, 2 * b + 1)), pf)))
            MB_ATOMIC_MAX_U32(&A.fc[unit], coll_make(nb - 1 - b, tmp_ev));
        }
      }
      cta.sync();
    } else {
      // general case: count the trials that need a draw, prefix-sum, then run them in order
      PerThread<u64> cnt(cta.nt());
      for (int pass = 0; pass < 2; ++pass) {  // 0: rev (ascending b), 1: fwd (descending b)
        MB_REGION(cta, tid) {
          u32 lo, hi;
          chunk(tid, nb, &lo, &hi);
          u64 c = 0;
          for (u32 t = lo; t < hi; ++t) {
            const u32 b = pass == 0 ? t : nb - 1 - t;
            if (!A.bar_active[b]) continue;
            const bool brev = bar_blocks_rev(b);
            const double pb = (pass == 0) == brev ? pmaj : pmin;
            u32 unit;
            u32 nohint = 0xFFFFFFFFu;
            const bool cand = pass == 0 ? lef_bar_candidate_rev(b, j0, &unit, &nohint)
                                        : lef_bar_candidate_fwd(b, jend, &unit, &nohint);
            if (cand && pb != 0.0 && pb != 1.0) ++c;
          }
          cnt[tid] = c;
        }
        const u64 total = cta.exscan_sum(cnt);
        rng_ensure(S.rng_pos + total);
        MB_REGION(cta, tid) {
          u32 lo, hi;
          chunk(tid, nb, &lo, &hi);
          u64 o = S.rng_pos + cnt[tid];
          for (u32 t = lo; t < hi; ++t) {
            const u32 b = pass == 0 ? t : nb - 1 - t;
            if (!A.bar_active[b]) continue;
            const bool brev = bar_blocks_rev(b);
            const double pb = (pass == 0) == brev ? pmaj : pmin;
            u32 unit;
            u32 nohint = 0xFFFFFFFFu;
            const bool cand = pass == 0 ? lef_bar_candidate_rev(b, j0, &unit, &nohint)
                                        : lef_bar_candidate_fwd(b, jend, &unit, &nohint);
            if (!cand) continue;
            bool hit;
            if (pb == 1.0) {
              hit = true;
            } else if (pb == 0.0) {
              hit = false;
            } else {
              hit = bernoulli_raw(raw(o++), pb);
            }
            if (hit) {
              if (pass == 0) {
                MB_ATOMIC_MAX_U32(&A.rc[unit], coll_make(b, tmp_ev));
              } else {
                MB_ATOMIC_MAX_U32(&A.fc[unit], coll_make(nb - 1 - b, tmp_ev));
              }
            }
          }
        }
        cta.sync();
        MB_REGION(cta, tid) {
          if (cta.leader(tid)) S.rng_pos += total;
        }
        cta.sync();
      }
    }
    // strip the temporary marker (it made LEF-BAR hits win over boundary marks in atomicMax)
    MB_REGION(cta, tid) {
      for (u32 i = tid; i < n; i += cta.nt()) {
        const u32 r = A.rc[i], f = A.fc[i];
        if (r & 0x80000000u) A.rc[i] = coll_make(coll_index(r), kEvCollision | kEvLefBar);
        if (f & 0x80000000u)
          A.fc[i] = coll_make(nb - 1 - coll_index(f), kEvCollision | kEvLefBar);
      }
    }
    cta.sync();
  }

  // compute_lef_lef_collision_pos (simulation.cpp:523-551) -> rev / fwd collision positions
  MB_FN static void lef_lef_collision_pos(u32 rev_pos, u32 fwd_pos, u32 rev_move, u32 fwd_move,
                                          u32* cp_rev, u32* cp_fwd) {
    const u64 rel = u64(rev_move) + u64(fwd_move);
    const double t = static_cast<double>(rev_pos - fwd_pos) / static_cast<double>(rel);
    const u32 cp = fwd_pos + static_cast<u32>(round(static_cast<double>(fwd_move) * t));
    if (cp == fwd_pos) {
      *cp_rev = cp + 1;
      *cp_fwd = cp;
    } else {
      *cp_rev = cp;
      *cp_fwd = cp - 1;
    }
  }

  // detect_primary_lef_lef_collisions (simulation_detect_collisions.cpp:250-397). The tested
  // pairs are the places of the merged 5'->3' order where a fwd unit is directly followed by a
  // rev unit; pairs are disjoint, so they are resolved independently. Bernoulli(1-bypass) draws
  // go to the geometrically colliding pairs in ascending order.
  // *hint: the fwd rank found for the previous (lower) rev rank of this thread, 0xFFFFFFFF = none
  MB_FN bool primary_pair(u32 j, u32 n5, u32 i2, u32* r_out, u32* f_out, u32* hint) const {
    const u32 r = A.rr[j];
    const u32 rp = A.rev[r];
    // first fwd rank in [0, i2) with pos >= rp
    u32 a;
    if (*hint == 0xFFFFFFFFu) {  // first rev rank of this thread: binary search over [0, i2)
      u32 lo = 0, hi = i2;
      while (lo < hi) {
        const u32 m = (lo + hi) >> 1;
        if (A.fwd[A.fr[m]] < rp) {
          lo = m + 1;
        } else {
          hi = m;
        }
      }
      a = lo;
    } else {
      a = count_fwd_lt_near(rp, *hint, i2);
    }
    *hint = a;
    if (a == i2 || a == 0) return false;
    const u32 f = A.fr[a - 1];
    if (j > n5 && A.rev[A.rr[j - 1]] > A.fwd[f]) return false;
    const u32 delta = rp - A.fwd[f];
    if (!(u64(delta) < u64(A.rm[r]) + u64(A.fm[f]))) return false;
    *r_out = r;
    *f_out = f;
    return true;
  }

  MB_FN void primary_apply(u32 r, u32 f) const {
    const u32 rcol = A.rc[r], fcol = A.fc[f];
    const u32 hit_r = coll_make(f, kEvCollision | kEvPrimary);
    const u32 hit_f = coll_make(r, kEvCollision | kEvPrimary);
    if (!coll_occurred(rcol) && !coll_occurred(fcol)) {  // (most pairs: no division needed)
      A.rc[r] = hit_r;
      A.fc[f] = hit_f;
      return;
    }
    if (coll_occurred(rcol) && coll_occurred(fcol)) return;
    // one unit is stalled by a barrier already: where the two would meet decides
    u32 cp_rev, cp_fwd;
    lef_lef_collision_pos(A.rev[r], A.fwd[f], A.rm[r], A.fm[f], &cp_rev, &cp_fwd);
    if (coll_occurred(rcol)) {
      // the reference asserts LEF_BAR here; a boundary mark has index 5/3 and is read the same way
      const u32 barrier_pos = A.bar_pos[coll_index(rcol)];
      if (cp_fwd > barrier_pos) A.rc[r] = hit_r;
      A.fc[f] = hit_f;
    } else {
      const u32 barrier_pos = A.bar_pos[coll_index(fcol)];
      A.rc[r] = hit_r;
      if (cp_rev < barrier_pos) A.fc[f] = hit_f;
    }
  }

  MB_FN void detect_primary_lef_lef_collisions() {
    const u32 n = S.num_active;
    const u32 n5 = S.n5, n3 = S.n3;
    if (n5 == n || n3 == n) return;
    const u32 i2 = n - (n3 ? n3 - 1 : 0);
    const u32 M = n - n5;
    // run_lef_lef_collision_trial (simulation_impl.hpp:93-96): no draw when bypass == 0 (always
    // collide) and none when 1 - bypass == 0 (Bernoulli(0) never draws and never succeeds)
    if (P.p_bypass != 0.0 && 1.0 - P.p_bypass == 0.0) return;
    const bool draws = P.p_bypass != 0.0;
    if constexpr (kCtr) {
      // the trial of the pair found at rev rank n5 + m reads draw m of this epoch: no counting pass
      MB_REGION(cta, tid) {
        u32 lo, hi;
        chunk(tid, M, &lo, &hi);
        u32 hint = 0xFFFFFFFFu;
        for (u32 m = lo; m < hi; ++m) {
          u32 r, f;
          if (!primary_pair(n5 + m, n5, i2, &r, &f, &hint)) continue;
          if (draws && !bernoulli_raw(raw(ctr_pack(S.epoch, kDrPrimary, m)), 1.0 - P.p_bypass))
            continue;
          primary_apply(r, f);
        }
      }
      cta.sync();
      return;
    }
    PerThread<u64> cnt(cta.nt());
    u64 total = 0;
    if (draws) {
      MB_REGION(cta, tid) {
        u32 lo, hi;
        chunk(tid, M, &lo, &hi);
        u64 c = 0;
        u32 hint = 0xFFFFFFFFu;
        for (u32 m = lo; m < hi; ++m) {
          u32 r, f;
          const bool is_pair = primary_pair(n5 + m, n5, i2, &r, &f, &hint);
          A.scratch[m] = is_pair ? f : 0xFFFFFFFFu;  // remembered for the pass that draws
          c += is_pair;
        }
        cnt[tid] = c;
      }
      total = cta.exscan_sum(cnt);
      rng_ensure(S.rng_pos + total);
    }
    MB_REGION(cta, tid) {
      u32 lo, hi;
      chunk(tid, M, &lo, &hi);
      u64 o = S.rng_pos + (draws ? cnt[tid] : 0);
      u32 hint = 0xFFFFFFFFu;
      for (u32 m = lo; m < hi; ++m) {
        u32 r, f;
        if (draws) {
          f = A.scratch[m];
          if (f == 0xFFFFFFFFu) continue;
          r = A.rr[n5 + m];
          if (!bernoulli_raw(raw(o++), 1.0 - P.p_bypass)) continue;
        } else if (!primary_pair(n5 + m, n5, i2, &r, &f, &hint)) {
          continue;
        }
        primary_apply(r, f);
      }
    }
    cta.sync();
    if (draws) {
      MB_REGION(cta, tid) {
        if (cta.leader(tid)) S.rng_pos += total;
      }
      cta.sync();
    }
  }

  // correct_moves_for_lef_bar_collisions (simulation_correct_moves.cpp:19-50)
  // correct_moves_for_primary_lef_lef_collisions (:53-121)
  MB_FN void correct_moves() {
    const u32 n = S.num_active;
    MB_REGION(cta, tid) {
      for (u32 i = tid; i < n; i += cta.nt()) {
        if (coll_is(A.rc[i], kEvLefBar)) A.rm[i] = (A.rev[i] - A.bar_pos[coll_index(A.rc[i])]) - 1;
        if (coll_is(A.fc[i], kEvLefBar)) A.fm[i] = (A.bar_pos[coll_index(A.fc[i])] - A.fwd[i]) - 1;
      }
    }
    cta.sync();
    // One region for both primary cases: a rev unit whose partner also carries the primary mark
    // rewrites both moves; a unit whose partner is stalled by a barrier rewrites only its own and
    // reads the partner's (set in the region above, not touched here) -- the two loops write
    // disjoint words.
    // (a thread first notes which of its rev units meet a fwd unit half-way and then works those
    // off back to back: the division sits in that part, and a warp goes through it as often as
    // its busiest lane has pairs rather than once per trip with a few lanes each)
    MB_REGION(cta, tid) {
      for (u32 r0 = tid; r0 < n; r0 += 32u * cta.nt()) {
        u32 both = 0, k = 0;
        for (u32 r = r0; k < 32 && r < n; r += cta.nt(), ++k) {
          if (!coll_is(A.rc[r], kEvPrimary)) continue;
          const u32 f = coll_index(A.rc[r]);
          if (coll_is(A.fc[f], kEvPrimary)) {
            both |= 1u << k;
          } else if (coll_is(A.fc[f], kEvLefBar)) {
            A.rm[r] = A.rev[r] - (A.fwd[f] + A.fm[f]) - 1;
          }
        }
        while (both) {
          k = static_cast<u32>(MB_FFS(both)) - 1;
          both &= both - 1;
          const u32 r = r0 + k * cta.nt();
          const u32 f = coll_index(A.rc[r]);
          u32 p1, p2;
          lef_lef_collision_pos(A.rev[r], A.fwd[f], A.rm[r], A.fm[f], &p1, &p2);
          A.rm[r] = A.rev[r] - p1;
          A.fm[f] = p2 - A.fwd[f];
        }
      }
      for (u32 f = tid; f < n; f += cta.nt()) {
        if (!coll_is(A.fc[f], kEvPrimary)) continue;
        const u32 r = coll_index(A.fc[f]);
        if (coll_is(A.rc[r], kEvLefBar)) A.fm[f] = (A.rev[r] - A.rm[r]) - A.fwd[f] - 1;
      }
    }
    cta.sync();
  }

  // process_secondary_lef_lef_collisions (simulation_detect_collisions.cpp:400-515), one
  // direction. Scan position m = 0..M-1 maps to rank `first + m` (rev pass, values as they are)
  // or `first - m` (fwd pass, values negated) so that both passes read: a stalled unit at m-1
  // stalls the free unit at m when  q[m] <= v[m-1],  after which  v[m] = min(pos[m], v[m-1]+1).
  template <bool kRevPass>
  MB_FN u32 sec_idx(u32 first, u32 m) const {
    return kRevPass ? A.rr[first + m] : A.fr[first - m];
  }
  // (32-bit signed: positions are below 2^28 and moves below 2^24)
  template <bool kRevPass>
  MB_FN i32 sec_pos(u32 idx) const {
    return kRevPass ? static_cast<i32>(A.rev[idx]) : -static_cast<i32>(A.fwd[idx]);
  }
  template <bool kRevPass>
  MB_FN i32 sec_q(u32 idx) const {  // position after the unit's current move
    return kRevPass ? static_cast<i32>(A.rev[idx]) - static_cast<i32>(A.rm[idx])
                    : -(static_cast<i32>(A.fwd[idx]) + static_cast<i32>(A.fm[idx]));
  }

  // The rev pass (5'->3' over rev ranks) and the fwd pass (3'->5' over fwd ranks) touch disjoint
  // arrays, so both run through the same steps together; only the draws couple them in the
  // deterministic mode (the fwd pass draws after the rev pass), and those are resolved in one
  // walk over the concatenated candidate list.
  struct SecDir {
    u32 first, M;
    u32* cand;   // by scan position: the unit is a candidate
    u32* head1;  // by scan position: ... directly behind an already stalled unit
  };

  // Candidates of one direction, found by walking from every unit that was stalled BEFORE this
  // pass (a "head": LEF-BAR, primary or boundary collision): the free units behind it are
  // candidates for as long as each would reach the site of the one before (q[k] <= v), assuming
  // the trials of the earlier ones succeed; the run ends at the first unit out of reach or at
  // the next head (whose owner walks on from there). Runs are disjoint, so the walks of
  // different heads write different words; a walk may leave its thread's chunk of scan
  // positions. Every candidate gets its move-if-stalled parked in the index bits of its (so far
  // empty) collision word, its bit in `cand`, and -- directly behind the head -- in `head1`.
  // Throughput mode: a candidate's trial is keyed by (pass, scan position), so the walk runs it
  // on the spot, marks only the candidates it REACHES, records the successful ones in `ok`, and
  // stops at the first failure (kNever: trials cannot succeed and draw nothing).
  template <bool kRevPass>
  MB_FN void sec_walk(const SecDir& d, int tid, bool draws, bool never, u32* ok) const {
    u32* coll = kRevPass ? A.rc : A.fc;
    u32 lo, hi;
    chunk(tid, d.M, &lo, &hi);
    // A thread first notes which of its scan positions hold a stalled unit (one in four or so) and
    // then walks from those back to back: a warp goes through the walk as often as its busiest
    // lane has heads, not once per scan position with a few lanes each.
    for (u32 m0 = lo; m0 < hi; m0 += 32) {
      u32 heads = 0;
      const u32 m1 = m0 + 32 < hi ? m0 + 32 : hi;
      for (u32 m = m0; m < m1; ++m) {
        // (a neighbouring walk may be parking a move in the index bits of this word; the event
        // bits read here do not change in this region)
        const u32 cw = MB_SHARED_LOAD_U32(&coll[sec_idx<kRevPass>(d.first, m)]);
        heads |= u32(coll_occurred(cw)) << (m - m0);
      }
      while (heads) {
        const u32 m = m0 + static_cast<u32>(MB_FFS(heads)) - 1;
        heads &= heads - 1;
        const u32 idx = sec_idx<kRevPass>(d.first, m);
        i32 v = sec_q<kRevPass>(idx);
        // One unit of a run: false when the run ends at scan position k (unit ik with collision
        // word cw, position p and target q).
        auto step = [&](u32 k, u32 ik, u32 cw, i32 p, i32 q) -> bool {
          if (coll_occurred(cw) || q > v) return false;
          const i32 mv = p - v;  // distance to the blocker's site
          // (event bits stay 0)
          MB_SHARED_STORE_U32(&coll[ik], static_cast<u32>(mv > 0 ? mv - 1 : 0));
          MB_ATOMIC_OR_U32(&d.cand[k >> 5], 1u << (k & 31));
          if (k == m + 1) MB_ATOMIC_OR_U32(&d.head1[k >> 5], 1u << (k & 31));
          if constexpr (kCtr) {
            if (never ||
                (draws && !bernoulli_raw(raw(ctr_pack(S.epoch, kDrSecondary,
                                                      ((kRevPass ? 0u : 1u) << 24) | k)),
                                         1.0 - P.p_bypass)))
              return false;
            MB_ATOMIC_OR_U32(&ok[k >> 5], 1u << (k & 31));
          }
          v = p < v + 1 ? p : v + 1;
          return true;
        };
        // Most stalled units have nobody within reach behind them: the first follower is looked at
        // on its own. For the runs that do go on, the only value carried from one unit to the next
        // is v, so the loads of the next four units are issued together before they are looked at
        // (a queue behind a barrier can be dozens of units long, and the longest one sets the pace
        // of this region).
        if (m + 1 >= d.M) continue;
        {
          const u32 i1 = sec_idx<kRevPass>(d.first, m + 1);
          if (!step(m + 1, i1, MB_SHARED_LOAD_U32(&coll[i1]), sec_pos<kRevPass>(i1),
                    sec_q<kRevPass>(i1)))
            continue;
        }
        bool open_run = true;
        for (u32 k0 = m + 2; open_run && k0 < d.M; k0 += 4) {
          u32 ik[4], cw[4];
          i32 pp[4], qq[4];
#pragma unroll
          for (u32 j = 0; j < 4; ++j)
            ik[j] = k0 + j < d.M ? sec_idx<kRevPass>(d.first, k0 + j) : idx;
#pragma unroll
          for (u32 j = 0; j < 4; ++j) {
            cw[j] = MB_SHARED_LOAD_U32(&coll[ik[j]]);
            pp[j] = sec_pos<kRevPass>(ik[j]);
            qq[j] = sec_q<kRevPass>(ik[j]);
          }
#pragma unroll
          for (u32 j = 0; j < 4; ++j) {
            if (k0 + j >= d.M || !step(k0 + j, ik[j], cw[j], pp[j], qq[j])) {
              open_run = false;
              break;
            }
          }
        }
      }
    }
  }
  // number of candidates among this thread's scan positions
  MB_FN u32 sec_count(const SecDir& d, int tid) const {
    u32 lo, hi;
    chunk(tid, d.M, &lo, &hi);
    u32 c = 0;
    for (u32 m = lo; m < hi; ++m) c += (d.cand[m >> 5] >> (m & 31)) & 1u;
    return c;
  }
  // Throughput mode: outcome of every reached candidate (`ok` is indexed by scan position).
  template <bool kRevPass>
  MB_FN void sec_apply_ctr(const SecDir& d, int tid, const u32* ok) const {
    u32* coll = kRevPass ? A.rc : A.fc;
    u32* moves = kRevPass ? A.rm : A.fm;
    u32 lo, hi;
    chunk(tid, d.M, &lo, &hi);
    for (u32 m = lo; m < hi; ++m) {
      if (!((d.cand[m >> 5] >> (m & 31)) & 1u)) continue;
      const u32 idx = sec_idx<kRevPass>(d.first, m);
      const u32 blocker = sec_idx<kRevPass>(d.first, m - 1);
      if ((ok[m >> 5] >> (m & 31)) & 1u) {
        moves[idx] = coll[idx];  // parked by sec_walk
        coll[idx] = coll_make(blocker, kEvCollision | kEvSecondary);
      } else {
        coll[idx] = coll_make(blocker, kEvSecondary);
      }
    }
  }

  // Copies the first-in-run flags of this thread's candidates to their candidate numbers.
  MB_FN void sec_number_firsts(const SecDir& d, int tid, u32 c, u32* firstc) const {
    u32 lo, hi;
    chunk(tid, d.M, &lo, &hi);
    for (u32 m = lo; m < hi; ++m) {
      if (!((d.cand[m >> 5] >> (m & 31)) & 1u)) continue;
      if ((d.head1[m >> 5] >> (m & 31)) & 1u) MB_ATOMIC_OR_U32(&firstc[c >> 5], 1u << (c & 31));
      ++c;
    }
  }
  // mode 0: outcomes from the reached / ok bitmaps; 1: every candidate stalls (bypass == 0);
  // 2: trials never succeed and draw nothing (bypass == 1): only run heads are reached.
  template <bool kRevPass>
  MB_FN void sec_apply(const SecDir& d, int tid, u32 c, int mode, const u32* reached_bits,
                       const u32* ok_bits) const {
    u32* coll = kRevPass ? A.rc : A.fc;
    u32* moves = kRevPass ? A.rm : A.fm;
    u32 lo, hi;
    chunk(tid, d.M, &lo, &hi);
    for (u32 m = lo; m < hi; ++m) {
      if (!((d.cand[m >> 5] >> (m & 31)) & 1u)) continue;
      bool reached, ok;
      if (mode == 0) {
        reached = (reached_bits[c >> 5] >> (c & 31)) & 1u;
        ok = (ok_bits[c >> 5] >> (c & 31)) & 1u;
      } else if (mode == 2) {
        reached = (d.head1[m >> 5] >> (m & 31)) & 1u;
        ok = false;
      } else {
        reached = ok = true;
      }
      ++c;
      const u32 idx = sec_idx<kRevPass>(d.first, m);
      if (!reached) {
        coll[idx] = 0;
        continue;
      }
      const u32 blocker = sec_idx<kRevPass>(d.first, m - 1);
      if (ok) {
        moves[idx] = coll[idx];  // parked by sec_classify
        coll[idx] = coll_make(blocker, kEvCollision | kEvSecondary);
      } else {
        coll[idx] = coll_make(blocker, kEvSecondary);
      }
    }
  }

  // process_secondary_lef_lef_collisions (simulation_detect_collisions.cpp:400-515)
  MB_FN void process_secondary_lef_lef_collisions() {
    const u32 n = S.num_active;
    const bool never = P.p_bypass != 0.0 && 1.0 - P.p_bypass == 0.0;  // trials fail, no draws
    const bool draws = P.p_bypass != 0.0 && !never;
    const u32 nwords = (n + 31) / 32 + 2;
    SecDir R, F;
    const u32 k0 = S.n5 > 1 ? S.n5 : 1;  // the rev pass looks at rank pairs (k-1, k), k >= k0
    R.first = k0 - 1;
    R.M = k0 < n ? n - (k0 - 1) : 0;
    const u32 sat3 = S.n3 ? S.n3 - 1 : 0;
    F.first = n - sat3 - 1;  // the fwd pass starts from this rank and walks down
    F.M = F.first + 1;
    if (R.M < 2) R.M = 0;
    if (F.M < 2) F.M = 0;
    if (R.M == 0 && F.M == 0) return;
    R.cand = A.bits;
    R.head1 = A.bits + nwords;
    F.cand = A.bits + 2 * nwords;
    F.head1 = A.bits + 3 * nwords;
    u32* bits_firstc = A.bits + 4 * nwords;   // by candidate number (rev candidates, then fwd)
    u32* bits_reached = A.bits + 6 * nwords;
    u32* bits_ok = A.bits + 8 * nwords;
    u32* bits_fail = A.bits + 10 * nwords;    // by draw

    sub_begin();
    MB_REGION(cta, tid) {
      for (u32 w = tid; w < 12 * nwords; w += cta.nt()) A.bits[w] = 0;
    }
    cta.sync();
    sub_lap(kPhSecCompose);
    if constexpr (kCtr) {
      u32* ok_r = A.bits + 8 * nwords;  // by scan position
      u32* ok_f = A.bits + 9 * nwords;
      MB_REGION(cta, tid) {
        sec_walk<true>(R, tid, draws, never, ok_r);
        sec_walk<false>(F, tid, draws, never, ok_f);
      }
      cta.sync();
      sub_lap(kPhSecClassify);
      MB_REGION(cta, tid) {
        sec_apply_ctr<true>(R, tid, ok_r);
        sec_apply_ctr<false>(F, tid, ok_f);
      }
      cta.sync();
      sub_lap(kPhSecApply);
      return;
    }
    MB_REGION(cta, tid) {
      sec_walk<true>(R, tid, draws, never, nullptr);
      sec_walk<false>(F, tid, draws, never, nullptr);
    }
    cta.sync();
    sub_lap(kPhSecScan);
    PerThread<u64> cnt(cta.nt());  // low word: rev candidates, high word: fwd candidates
    MB_REGION(cta, tid) { cnt[tid] = u64(sec_count(R, tid)) | (u64(sec_count(F, tid)) << 32); }
    const u64 tot = cta.exscan_sum(cnt);  // cnt[tid]: candidates before this thread's
    const u32 npot_r = static_cast<u32>(tot & 0xFFFFFFFFu);
    const u32 npot = npot_r + static_cast<u32>(tot >> 32);
    sub_lap(kPhSecClassify);
    if (npot == 0) return;
    u64 pregen_base = 0;
    bool pregen = false;
    if (draws) {
      rng_ensure(S.rng_pos + npot);
      MB_REGION(cta, tid) {
        sec_number_firsts(R, tid, static_cast<u32>(cnt[tid] & 0xFFFFFFFFu), bits_firstc);
        sec_number_firsts(F, tid, npot_r + static_cast<u32>(cnt[tid] >> 32), bits_firstc);
        for (u32 d = tid; d < npot; d += cta.nt()) {
          if (!bernoulli_raw(raw(S.rng_pos + d), 1.0 - P.p_bypass))
            MB_ATOMIC_OR_U32(&bits_fail[d >> 5], 1u << (d & 31));
        }
      }
      cta.sync();
      sub_lap(kPhSecDraws);
      // While warp 0 walks the candidates below, the generator threads would idle at the barrier:
      // they stage the NEXT window of the cell's stream instead, whenever the ring has room for
      // it (the window it overwrites lies wholly behind the stream position; the draws of this
      // pass have been read already). The stream is sequential, so a window staged early is never
      // wasted; S.rng_generated moves in the region after the next barrier.
      pregen_base = S.rng_generated;
      pregen = pregen_base <= S.rng_pos + P.rng_window;
      // Walk the candidates in order (all rev candidates, then all fwd candidates; the first
      // fwd candidate always starts a run). A candidate is reached when it is the first of its
      // run or the previous one was reached and its trial succeeded; only reached candidates
      // draw, so the d-th draw belongs to the d-th reached candidate.
      const u32 ncw = (npot + 31) / 32;
#if MB_DEVICE_BUILD
      // Warp 0, one lane per candidate, 32 candidates at a time: start from "every candidate
      // that can be reached is", look up each lane's draw, drop the candidates behind a failed
      // trial, and repeat until nothing changes (a fixed point satisfies the sequential
      // recurrence, and candidate i is final after i + 1 rounds at the latest).
      if (threadIdx.x < 32) {
        const u32 lane = threadIdx.x;
        const u32 lt = (1u << lane) - 1u;
        u32 d = 0, alive_in = 0;
        for (u32 w = 0; w < ncw; ++w) {
          const u32 Fw = bits_firstc[w];
          const u32 nvalid = npot - 32 * w < 32 ? npot - 32 * w : 32;
          const u32 Fle = Fw & (lt | (1u << lane));
          const u32 s = Fle ? 31u - static_cast<u32>(__clz(static_cast<int>(Fle))) : 0u;
          const u32 range = lt & ~((1u << s) - 1u);  // candidates of my run before me
          const bool can = lane < nvalid && (Fle != 0 || alive_in != 0);
          u32 Rm = __ballot_sync(0xffffffffu, can);
          u32 bad = 0;
          // outcomes of the draws d .. d + 31 (all this word can use), read once per word
          const u32 fw = __funnelshift_r(bits_fail[d >> 5], bits_fail[(d >> 5) + 1], d & 31u);
          for (int it = 0; it < 34; ++it) {
            // the d-th draw belongs to the d-th REACHED candidate
            const u32 fbit = (fw >> __popc(Rm & lt)) & 1u;
            bad = __ballot_sync(0xffffffffu, ((Rm >> lane) & 1u) && fbit);
            const u32 Rn = __ballot_sync(0xffffffffu, can && (bad & range) == 0);
            if (Rn == Rm) break;
            Rm = Rn;
          }
          if (lane == 0) {
            bits_reached[w] = Rm;
            bits_ok[w] = Rm & ~bad;
          }
          d += static_cast<u32>(__popc(Rm));
          alive_in = ((Rm & ~bad) >> 31) & 1u;
        }
        if (lane == 0) S.tmp_u32[4] = d;
      } else if (pregen) {
        for (u32 gi = threadIdx.x - 32; gi < P.rng_gen_threads; gi += blockDim.x - 32)
          rng_generate_block(gi, pregen_base);
      }
#else
      MB_REGION(cta, tid) {
        if (pregen) {
          for (u32 gi = static_cast<u32>(tid); gi < P.rng_gen_threads;
               gi += static_cast<u32>(cta.nt()))
            rng_generate_block(gi, pregen_base);
        }
        if (!cta.leader(tid)) continue;
        u32 d = 0, alive = 0;
        for (u32 w = 0; w < ncw; ++w) {
          const u32 Fw = bits_firstc[w];
          const u32 lim = npot - 32 * w < 32 ? npot - 32 * w : 32;
          u32 Rm = 0, OK = 0;
          for (u32 i = 0; i < lim; ++i) {
            const u32 reach = ((Fw >> i) & 1u) | alive;
            alive = reach & ~(bits_fail[d >> 5] >> (d & 31)) & 1u;
            Rm |= reach << i;
            OK |= alive << i;
            d += reach;
          }
          bits_reached[w] = Rm;
          bits_ok[w] = OK;
        }
        S.tmp_u32[4] = d;
      }
#endif
      cta.sync();
      sub_lap(kPhSecLeader);
    }
    const int mode = draws ? 0 : (never ? 2 : 1);
    MB_REGION(cta, tid) {
      sec_apply<true>(R, tid, static_cast<u32>(cnt[tid] & 0xFFFFFFFFu), mode, bits_reached, bits_ok);
      sec_apply<false>(F, tid, npot_r + static_cast<u32>(cnt[tid] >> 32), mode, bits_reached,
                       bits_ok);
      if (draws && cta.leader(tid)) {
        S.rng_pos += S.tmp_u32[4];
        if (pregen) S.rng_generated = pregen_base + P.rng_window;
      }
    }
    cta.sync();
    sub_lap(kPhSecApply);
  }

  // fix_secondary_lef_lef_collisions (simulation_detect_collisions.cpp:517-644). Avoided markers
  // are never adjacent, so every marker touches its own pair of rank slots and LEFs.
  MB_FN void fix_secondary_lef_lef_collisions() {
    const u32 n = S.num_active;
    const u32 k0 = S.n5 > 1 ? S.n5 : 1;
    const u32 sat3 = S.n3 ? S.n3 - 1 : 0;
    const u32 naf = n - sat3;
    // Detection and application are separate regions: a swap rewrites rank slots and collision
    // words that the neighbouring threads read while looking for their own markers. A thread
    // remembers the markers of up to 64 of its strided ranks per round (one round on the device).
    const u32 per_round = 64u * static_cast<u32>(cta.nt());
    const u32 span = n > naf ? n : naf;
    PerThread<u64> marks_r(cta.nt()), marks_f(cta.nt());
    for (u32 base = 0; base < span; base += per_round) {
      MB_REGION(cta, tid) {
        u64 mr = 0, mf = 0;
        u32 it = 0;
        for (u32 i = base + k0 + tid; i < n && it < 64; i += cta.nt(), ++it)
          if (coll_avoided(A.rc[A.rr[i]], kEvSecondary)) mr |= u64(1) << it;
        it = 0;
        for (u32 i = base + tid; i + 1 < naf && it < 64; i += cta.nt(), ++it)
          if (coll_avoided(A.fc[A.fr[i]], kEvSecondary)) mf |= u64(1) << it;
        marks_r[tid] = mr;
        marks_f[tid] = mf;
      }
      cta.sync();
      MB_REGION(cta, tid) {
        u64 mr = marks_r[tid];
        for (u32 i = base + k0 + tid; mr != 0; i += cta.nt(), mr >>= 1) {
          if (!(mr & 1)) continue;
          const u32 idx2 = A.rr[i];
          const u32 idx1 = A.rr[i - 1];
          const u32 pos1 = A.rev[idx1] - A.rm[idx1];
          u32 mv2 = 0;
          if (A.rev[idx2] > pos1 + 1) mv2 = A.rev[idx2] - (pos1 + 1);
          const u32 c2 = coll_make(idx1, kEvCollision | kEvSecondary);
          const u32 p1 = A.rev[idx1], p2 = A.rev[idx2];
          A.rev[idx1] = A.fwd[idx1] < p2 ? A.fwd[idx1] : p2;
          A.rev[idx2] = A.fwd[idx2] < p1 ? A.fwd[idx2] : p1;
          // swap collisions, moves and rank slots
          const u32 c1 = A.rc[idx1], m1 = A.rm[idx1];
          A.rc[idx1] = c2;
          A.rc[idx2] = c1;
          A.rm[idx1] = mv2;
          A.rm[idx2] = m1;
          A.rr[i - 1] = static_cast<u16>(idx2);
          A.rr[i] = static_cast<u16>(idx1);
          const u32 cap1 = A.rev[idx1] - P.start, cap2 = A.rev[idx2] - P.start;
          if (A.rm[idx1] > cap1) A.rm[idx1] = cap1;
          if (A.rm[idx2] > cap2) A.rm[idx2] = cap2;
        }
      }
      cta.sync();  // the fwd pass reads rev positions the rev pass may just have rewritten
      MB_REGION(cta, tid) {
        u64 mf = marks_f[tid];
        for (u32 i = base + tid; mf != 0; i += cta.nt(), mf >>= 1) {
          if (!(mf & 1)) continue;
          const u32 idx1 = A.fr[i];
          const u32 idx2 = A.fr[i + 1];
          const u32 pos2 = A.fwd[idx2] + A.fm[idx2];
          u32 mv1 = 0;
          if (pos2 > A.fwd[idx1] + 1) mv1 = pos2 - (A.fwd[idx1] + 1);
          const u32 c1 = coll_make(idx2, kEvCollision | kEvSecondary);
          const u32 p1 = A.fwd[idx1], p2 = A.fwd[idx2];
          A.fwd[idx1] = A.rev[idx1] > p2 ? A.rev[idx1] : p2;
          A.fwd[idx2] = A.rev[idx2] > p1 ? A.rev[idx2] : p1;
          const u32 c2 = A.fc[idx2], m2 = A.fm[idx2];
          A.fc[idx1] = c2;
          A.fc[idx2] = c1;
          A.fm[idx1] = m2;
          A.fm[idx2] = mv1;
          A.fr[i] = static_cast<u16>(idx2);
          A.fr[i + 1] = static_cast<u16>(idx1);
          const u32 cap1 = P.end - 1 - A.fwd[idx1], cap2 = P.end - 1 - A.fwd[idx2];
          if (A.fm[idx1] > cap1) A.fm[idx1] = cap1;
          if (A.fm[idx2] > cap2) A.fm[idx2] = cap2;
        }
      }
      cta.sync();
    }
  }

  // Simulation::process_collisions (simulation.cpp:763-793)
  MB_FN void process_collisions(bool with_fix) {
    const u32 n = S.num_active;
    MB_PROBE_BEGIN();
    MB_REGION(cta, tid) {
      for (u32 i = tid; i < n; i += cta.nt()) {
        A.rc[i] = 0;
        A.fc[i] = 0;
      }
    }
    cta.sync();
    MB_PROBE(kPhMvExceptions);
    MB_REGION(cta, tid) {
      if (cta.leader(tid)) detect_boundaries_leader();
    }
    cta.sync();
    MB_PROBE(kPhSecScan);
    detect_lef_bar_collisions();
    MB_PROBE(kPhSecDraws);
    lap(kPhLefBar);
    detect_primary_lef_lef_collisions();
    lap(kPhPrimary);
    correct_moves();
    lap(kPhCorrect);
    process_secondary_lef_lef_collisions();
    lap(kPhSecondary);
    if (with_fix) fix_secondary_lef_lef_collisions();
    lap(kPhFix);
  }

  // extrude (simulation.cpp:498-521) and release_lefs (:553-601)
  MB_FN void extrude_and_release() {
    const u32 n = S.num_active;
    const double base_p = S.burnin_completed ? P.p_release : P.p_release_burnin;
    const bool draws = base_p != 0.0;
    if (draws) rng_ensure(S.rng_pos + n);
    const u64 base = kCtr ? ctr_pack(S.epoch, kDrRelease, 0) : S.rng_pos;
    MB_REGION(cta, tid) {
      for (u32 i = tid; i < n; i += cta.nt()) {
        u32 rev = A.rev[i] - A.rm[i];
        u32 fwd = A.fwd[i] + A.fm[i];
        u32 ep = A.ep[i];
        if (draws) {
          int hard = 0;
          const u32 rcol = A.rc[i], fcol = A.fc[i];
          if (coll_is(rcol, kEvLefBar) && bar_blocks_rev(coll_index(rcol))) ++hard;
          if (coll_is(fcol, kEvLefBar) && !bar_blocks_rev(coll_index(fcol))) ++hard;
          const double affinity =
              hard == 0 ? 1.0 : (hard == 1 ? 1.0 / P.soft_mult : 1.0 / P.hard_mult);
          if (bernoulli_raw(raw(kCtr ? base + (u64(i) << 8) : base + i), affinity * base_p)) {
            rev = kUnbound;
            fwd = kUnbound;
            ep = kUnbound;
          }
        }
        A.rev[i] = rev;
        A.fwd[i] = fwd;
        A.ep[i] = ep;
      }
      mv_clear_slow_bits(tid);  // the secondary pass is done with A.bits
    }
    cta.sync();  // every thread has read `base` before the leader moves the stream position
    if (!kCtr && draws) {
      MB_REGION(cta, tid) {
        if (cta.leader(tid)) S.rng_pos = base + n;
      }
      cta.sync();
    }
  }

  // Simulation::dump_stats (simulation.cpp:995-1056), called where the reference calls it: after
  // extrude, before release_lefs (:968-974). Runs before extrude_and_release here, so the loop
  // sizes are taken from the positions the units are about to reach.
  MB_FN void log_epoch_state() {
    if (K.log == nullptr || S.epoch >= K.log_cap) return;
    const u32 n = S.num_active;
    PerThread<u64> a(cta.nt()), b(cta.nt()), c(cta.nt()), d(cta.nt());
    MB_REGION(cta, tid) {
      u64 st_rev = 0, st_fwd = 0, st_both = 0, bar = 0, prim = 0, sec = 0, loops = 0, occ = 0;
      for (u32 i = tid; i < n; i += cta.nt()) {
        const u32 rc = A.rc[i], fc = A.fc[i];
        const bool r = coll_occurred(rc), f = coll_occurred(fc);
        st_rev += r;
        st_fwd += f;
        st_both += r && f;
        bar += coll_is(rc, kEvLefBar) + coll_is(fc, kEvLefBar);
        prim += coll_is(rc, kEvPrimary) + coll_is(fc, kEvPrimary);
        sec += coll_is(rc, kEvSecondary) + coll_is(fc, kEvSecondary);
        if (A.rev[i] != kUnbound) loops += u64(A.fwd[i] + A.fm[i]) - u64(A.rev[i] - A.rm[i]);
      }
      for (u32 i = tid; i < P.n_bar; i += cta.nt()) occ += A.bar_active[i] != 0;
      a[tid] = st_rev | (st_fwd << 21) | (st_both << 42);
      b[tid] = bar | (prim << 21) | (sec << 42);
      c[tid] = occ;
      d[tid] = loops;
    }
    const u64 ta = cta.reduce_sum(a);
    const u64 tb = cta.reduce_sum(b);
    const u64 tc = cta.reduce_sum(c);
    const u64 td = cta.reduce_sum(d);
    MB_REGION(cta, tid) {
      if (cta.leader(tid)) {
        modle_b200_epoch_record rec;
        rec.epoch = S.epoch;
        rec.loop_size_sum = td;
        rec.burnin = S.burnin_completed ? 0u : 1u;
        rec.num_lefs = n;
        rec.barriers_occupied = static_cast<u32>(tc);
        rec.lefs_stalled_rev = static_cast<u32>(ta & 0x1FFFFFu);
        rec.lefs_stalled_fwd = static_cast<u32>((ta >> 21) & 0x1FFFFFu);
        rec.lefs_stalled_both = static_cast<u32>(ta >> 42);
        rec.lef_bar_collisions = static_cast<u32>(tb & 0x1FFFFFu);
        rec.lef_lef_primary_collisions = static_cast<u32>((tb >> 21) & 0x1FFFFFu);
        rec.lef_lef_secondary_collisions = static_cast<u32>(tb >> 42);
        rec.reserved_ = 0;
        K.log[S.epoch] = rec;
      }
    }
    cta.sync();
  }

  // ------------------------------------------------------------------------------ main loop
  // Simulation::simulate_one_cell (simulation.cpp:896-986)
  MB_FN void run() {
#if MB_DEVICE_BUILD
    const u64 t_begin = static_cast<u64>(clock64());
    t_last = t_begin;
#endif
    if constexpr (kCtr) ctr_keys();
#if !MB_DEVICE_BUILD
    t_last = emu_barrier_count();
#endif
    init_cell();
    lap(kPhInit);
    for (;;) {
      bool stop;
      if (!P.stop_on_epochs) {
        stop = S.num_contacts >= task.target_contacts;
      } else {
        stop = S.epoch - S.num_burnin_epochs >= task.target_epochs;
      }
      if (stop || S.epoch >= P.debug_max_epochs || S.fault != 0) break;
      cta.sync();
      if (!S.burnin_completed) burnin_step();
      lap(kPhBurnin);
      if (S.fault != 0) break;
      bind_lefs();
      lap(kPhBind);
      rank_lefs();
      lap(kPhRank);
      if (S.burnin_completed) {
        sample_and_register_contacts();
        lap(kPhContacts);
        if (task.target_contacts != 0 && S.num_contacts >= task.target_contacts) break;
      }
      generate_moves();
      next_barrier_states();
      lap(kPhBarriers);
      process_collisions(true);
      log_epoch_state();
      extrude_and_release();
      lap(kPhExtrudeRelease);
      MB_REGION(cta, tid) {
        if (cta.leader(tid)) {
          S.lef_updates += S.num_active;
          ++S.epoch;
        }
      }
      cta.sync();
    }
    cta.sync();
#if MB_DEVICE_BUILD
    if (threadIdx.x == 0) S.phase_cycles[kPhTotal] = static_cast<u64>(clock64()) - t_begin;
#endif
    cta.sync();
  }
};

using CellSim = CellSimT<false>;            // deterministic mode (the reference's draw order)
using CellSimThroughput = CellSimT<true>;   // throughput mode (counter-based draws)

}  // namespace modle_b200
