// TEST INFRASTRUCTURE ONLY -- CPU oracle for the modle_b200 hot path.
// Nothing under oracle/ may be imported, linked or executed by the product path
// (modle_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs use it, and only as the checker / CPU baseline.
//
// PARITY STATUS: "parity unpinned" for the random distributions. The reference draws all of
// its random numbers through un-vendored third-party code (xoshiro-cpp 1.1, Boost.Random 1.88,
// xxHash 0.8.3; /root/reference/conanfile.py:39,62,63) that is absent from this container, so
// the algorithms below are restated from their published descriptions:
//   - SplitMix64 / xoshiro256++ / jump(): Blackman & Vigna reference implementation. Pinned
//     indirectly by test "Simulation 011/012" (test/units/simulation_cpu/
//     simulation_complex_unit_test.cpp:637-756), which needs the first output of PRNG(752741483)
//     to make Bernoulli(0.75) fail. jump() is pinned against an independent GF(2) matrix power.
//   - XXH3-64 with seed: pinned against the python `xxhash` package in tests/.
//   - bernoulli / uniform_int / generate_canonical / uniform_01 / normal (ziggurat) /
//     exponential (ziggurat) / poisson (inversion + PTRD) / binomial (inversion + BTRD):
//     Boost.Random algorithms, restated from memory of boost/random/*.hpp. No in-tree test of
//     the reference pins these; see DESIGN.md.
//   - ziggurat tables: Boost holds them as 20-digit literals of the EXACT solution of the
//     ziggurat equations (not Marsaglia & Tsang's rounded r = 3.442619855899);
//     scripts/make_ziggurat_tables.py recomputes that solution with 60 digits and writes the same
//     20-digit literals into ziggurat_tables.inc; the leading entries (3.7130862467403632609,
//     3.4426198558966521214, ... / 8.6971174701310497140, 7.6971174701310497140, ...) are pinned
//     in tests/test_oracle_kats.py.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <utility>

namespace oracle {

using u64 = std::uint64_t;
using u32 = std::uint32_t;
using i64 = std::int64_t;

// --------------------------------------------------------------------------------------------
// Engine. Mirrors modle::random::PRNG(seed) (src/common/include/modle/common/random.hpp:26-30):
// SplitMix64(seed) produces the 4 state words of a xoshiro256++ generator.
// --------------------------------------------------------------------------------------------
struct SplitMix64 {
  u64 s;
  u64 next() noexcept {
    u64 z = (s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
  }
};

struct Xoshiro256pp {
  u64 s[4];
  u64 ndraws = 0;  // number of raw 64-bit draws consumed so far (for parity diagnostics only)

  static constexpr u64 rotl(u64 x, int k) noexcept { return (x << k) | (x >> (64 - k)); }

  static Xoshiro256pp from_seed(u64 seed) noexcept {
    SplitMix64 sm{seed};
    Xoshiro256pp g{};
    for (auto& w : g.s) w = sm.next();
    return g;
  }
  static Xoshiro256pp from_state(const u64* st) noexcept {
    Xoshiro256pp g{};
    for (int i = 0; i < 4; ++i) g.s[i] = st[i];
    return g;
  }

  u64 next() noexcept {
    const u64 result = rotl(s[0] + s[3], 23) + s[0];
    const u64 t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl(s[3], 45);
    ++ndraws;
    return result;
  }

  // Equivalent to 2^128 calls to next(); used by run_simulate once per cell
  // (src/libmodle/cpu/scheduler_simulate.cpp:121,158).
  void jump() noexcept {
    static constexpr u64 JUMP[] = {0x180ec6d33cfd0abaULL, 0xd5a61266f0c9392cULL,
                                   0xa9582618e03fc9aaULL, 0x39abdc4529b1661cULL};
    u64 t[4] = {0, 0, 0, 0};
    const u64 saved = ndraws;
    for (u64 jw : JUMP) {
      for (int b = 0; b < 64; ++b) {
        if (jw & (u64(1) << b)) {
          for (int i = 0; i < 4; ++i) t[i] ^= s[i];
        }
        next();
      }
    }
    for (int i = 0; i < 4; ++i) s[i] = t[i];
    ndraws = saved;
  }
};

using Rng = Xoshiro256pp;

// --------------------------------------------------------------------------------------------
// Elementary Boost.Random pieces for a 64-bit engine with min()=0, max()=2^64-1.
// --------------------------------------------------------------------------------------------
constexpr double TWO64 = 18446744073709551616.0;  // double(2^64-1) also rounds to this

// boost::random::bernoulli_distribution<double>: no draw when p == 0.
inline bool bernoulli(Rng& g, double p) noexcept {
  if (p == 0.0) return false;
  return static_cast<double>(g.next()) <= p * TWO64;
}

// boost::random::generate_canonical<double, 53>: one draw; 1.0 is mapped just below 1.
inline double canonical(Rng& g) noexcept {
  double r = static_cast<double>(g.next()) / TWO64;
  if (r == 1.0) r -= 2.220446049250313e-16 / 2;
  return r;
}

// boost::random::uniform_01<double>: redraws instead of clamping.
inline double uniform01(Rng& g) noexcept {
  for (;;) {
    const double r = static_cast<double>(g.next()) * (1.0 / TWO64);
    if (r < 1.0) return r;
  }
}

// boost::random::uniform_int_distribution<uint64>{a, b}: bucketed rejection.
inline u64 uniform_int_bucket(u64 range) noexcept {
  const u64 brange = ~u64(0);
  u64 bucket = brange / (range + 1);
  if (brange % (range + 1) == range) ++bucket;
  return bucket;
}
inline u64 uniform_int(Rng& g, u64 a, u64 b) noexcept {
  const u64 range = b - a;
  if (range == 0) return a;
  if (range == ~u64(0)) return a + g.next();
  const u64 bucket = uniform_int_bucket(range);
  for (;;) {
    const u64 r = g.next() / bucket;
    if (r <= range) return a + r;
  }
}

// detail::generate_int_float_pair<double, 8>: one draw -> (53-bit uniform in [0,1), low 8 bits).
inline std::pair<double, int> int_float_pair8(Rng& g) noexcept {
  u64 u = g.next();
  const int bucket = static_cast<int>(u & 0xFF);
  u &= ~((u64(1) << 11) - 1);  // keep 53 significant bits above the bucket
  const double r = static_cast<double>(u >> 8) * (1.0 / 72057594037927936.0);  // 2^-56
  return {r, bucket};
}

// --------------------------------------------------------------------------------------------
// Ziggurat tables (generated file; see header note).
// --------------------------------------------------------------------------------------------
namespace zigdata {
#include "ziggurat_tables.inc"
}  // namespace zigdata
struct ZigTables {
  double nx[129], ny[129];  // normal, 128 layers
  double ex[257], ey[257];  // exponential, 256 layers
  ZigTables() {
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
inline const ZigTables& zig() {
  static const ZigTables t;
  return t;
}

// detail::unit_exponential_distribution<double>
inline double unit_exponential(Rng& g) noexcept {
  const double* tx = zig().ex;
  const double* ty = zig().ey;
  double shift = 0;
  for (;;) {
    const auto [u, i] = int_float_pair8(g);
    const double x = u * tx[i];
    if (x < tx[i + 1]) return shift + x;
    if (i == 0) {
      shift += tx[1];
    } else {
      const double y01 = uniform01(g);
      const double y = ty[i] + y01 * (ty[i + 1] - ty[i]);
      const double y_above_ubound = (tx[i] - tx[i + 1]) * y01 - (tx[i] - x);
      const double y_above_lbound = y - (ty[i + 1] + (tx[i + 1] - x) * ty[i + 1]);
      if (y_above_ubound < 0 && (y_above_lbound < 0 || y < std::exp(-x))) return x + shift;
    }
  }
}

// detail::unit_normal_distribution<double>
inline double unit_normal(Rng& g) noexcept {
  const double* tx = zig().nx;
  const double* ty = zig().ny;
  for (;;) {
    const auto [u, bits] = int_float_pair8(g);
    const int sign = (bits & 1) * 2 - 1;
    const int i = bits >> 1;
    const double x = u * tx[i];
    if (x < tx[i + 1]) return x * sign;
    if (i == 0) {
      const double tail_start = tx[1];
      for (;;) {
        const double xx = unit_exponential(g) / tail_start;
        const double yy = unit_exponential(g);
        if (2 * yy > xx * xx) return (xx + tail_start) * sign;
      }
    }
    const double y01 = uniform01(g);
    const double y = ty[i] + y01 * (ty[i + 1] - ty[i]);
    double y_above_ubound, y_above_lbound;
    if (tx[i] >= 1) {
      y_above_ubound = (tx[i] - tx[i + 1]) * y01 - (tx[i] - x);
      y_above_lbound = y - (ty[i] + (tx[i] - x) * ty[i] * tx[i]);
    } else {
      y_above_lbound = (tx[i] - tx[i + 1]) * y01 - (tx[i] - x);
      y_above_ubound = y - (ty[i] + (tx[i] - x) * ty[i] * tx[i]);
    }
    if (y_above_ubound < 0 && (y_above_lbound < 0 || y < std::exp(-(x * x / 2)))) return x * sign;
  }
}

// boost::random::normal_distribution<double>{mean, sigma}
inline double normal(Rng& g, double mean, double sigma) noexcept {
  return unit_normal(g) * sigma + mean;
}

// boost::random::poisson_distribution<size_t, double>
inline u64 poisson(Rng& g, double mean) noexcept {
  if (mean < 10) {
    double p = std::exp(-mean);
    u64 x = 0;
    double u = uniform01(g);
    while (u > p) {
      u = u - p;
      ++x;
      p = mean * p / static_cast<double>(x);
    }
    return x;
  }
  static constexpr double log_fact[10] = {0.0,
                                          0.0,
                                          0.69314718055994529,
                                          1.7917594692280550,
                                          3.1780538303479458,
                                          4.7874917427820458,
                                          6.5792512120101012,
                                          8.5251613610654147,
                                          10.604602902745251,
                                          12.801827480081469};
  const double smu = std::sqrt(mean);
  const double b = 0.931 + 2.53 * smu;
  const double a = -0.059 + 0.02483 * b;
  const double inv_alpha = 1.1239 + 1.1328 / (b - 3.4);
  const double v_r = 0.9277 - 3.6224 / (b - 2);
  for (;;) {
    double u;
    double v = uniform01(g);
    if (v <= 0.86 * v_r) {
      u = v / v_r - 0.43;
      return static_cast<u64>(std::floor((2 * a / (0.5 - std::abs(u)) + b) * u + mean + 0.445));
    }
    if (v >= v_r) {
      u = uniform01(g) - 0.5;
    } else {
      u = v / v_r - 0.93;
      u = ((u < 0) ? -0.5 : 0.5) - u;
      v = uniform01(g) * v_r;
    }
    const double us = 0.5 - std::abs(u);
    if (us < 0.013 && v > us) continue;
    const double k = std::floor((2 * a / us + b) * u + mean + 0.445);
    v = v * inv_alpha / (a / (us * us) + b);
    const double log_sqrt_2pi = 0.91893853320467267;
    if (k >= 10) {
      if (std::log(v * smu) <= (k + 0.5) * std::log(mean / k) - mean - log_sqrt_2pi + k -
                                   (1 / 12. - (1 / 360. - 1 / (1260. * k * k)) / (k * k)) / k) {
        return static_cast<u64>(k);
      }
    } else if (k >= 0) {
      if (std::log(v) <= k * std::log(mean) - mean - log_fact[static_cast<int>(k)]) {
        return static_cast<u64>(k);
      }
    }
  }
}

// boost::random::binomial_distribution<ptrdiff_t, double>{t, p}
inline double binomial_fc(i64 k) noexcept {
  static constexpr double tbl[10] = {0.08106146679532726, 0.04134069595540929,
                                     0.02767792568499834, 0.02079067210376509,
                                     0.01664469118982119, 0.01387612882307075,
                                     0.01189670994589177, 0.01041126526197209,
                                     0.009255462182712733, 0.008330563433362871};
  if (k < 10) return tbl[k];
  const double ikp1 = 1.0 / static_cast<double>(k + 1);
  return (1.0 / 12 - (1.0 / 360 - (1.0 / 1260) * (ikp1 * ikp1)) * (ikp1 * ikp1)) * ikp1;
}

inline i64 binomial(Rng& g, i64 t, double p_in) noexcept {
  const bool flip = 0.5 < p_in;
  const double p = flip ? (1 - p_in) : p_in;
  const i64 m = static_cast<i64>(static_cast<double>(t + 1) * p);
  i64 res;
  if (m < 11) {
    const double q_n = std::pow(1 - p, static_cast<double>(t));
    const double q = 1 - p;
    const double s = p / q;
    const double a = static_cast<double>(t + 1) * s;
    double r = q_n;
    double u = uniform01(g);
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
    const double sqrt_npq = std::sqrt(npq);
    const double b = 1.15 + 2.53 * sqrt_npq;
    const double a = -0.0873 + 0.0248 * b + 0.01 * p;
    const double c = td * p + 0.5;
    const double alpha = (2.83 + 5.1 / b) * sqrt_npq;
    const double v_r = 0.92 - 4.2 / b;
    const double u_rv_r = 0.86 * v_r;
    for (;;) {
      double u;
      double v = uniform01(g);
      if (v <= u_rv_r) {
        u = v / v_r - 0.43;
        res = static_cast<i64>(std::floor((2 * a / (0.5 - std::abs(u)) + b) * u + c));
        break;
      }
      if (v >= v_r) {
        u = uniform01(g) - 0.5;
      } else {
        u = v / v_r - 0.93;
        u = ((u < 0) ? -0.5 : 0.5) - u;
        v = uniform01(g) * v_r;
      }
      const double us = 0.5 - std::abs(u);
      const i64 k = static_cast<i64>(std::floor((2 * a / us + b) * u + c));
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
      v = std::log(v);
      const double rho = (km / npq) * (((km / 3. + 0.625) * km + 1. / 6) / npq + 0.5);
      const double tt = -km * km / (2 * npq);
      if (v < tt - rho) {
        res = k;
        break;
      }
      if (v > tt + rho) continue;
      const i64 nm = t - m + 1;
      const double h = (static_cast<double>(m) + 0.5) *
                           std::log(static_cast<double>(m + 1) / (r * static_cast<double>(nm))) +
                       binomial_fc(m) + binomial_fc(t - m);
      const i64 nk = t - k + 1;
      if (v <= h +
                   static_cast<double>(t + 1) *
                       std::log(static_cast<double>(nm) / static_cast<double>(nk)) +
                   (static_cast<double>(k) + 0.5) *
                       std::log(static_cast<double>(nk) * r / static_cast<double>(k + 1)) -
                   binomial_fc(k) - binomial_fc(t - k)) {
        res = k;
        break;
      }
    }
  }
  return flip ? t - res : res;
}

// modle::genextreme_value_distribution<double>
// (src/common/include/modle/common/genextreme_value_distribution.hpp:87-105)
inline double genextreme(Rng& g, double mu, double sigma, double xi) noexcept {
  if (xi == 0.0) return (mu - sigma) * std::log(-std::log(canonical(g)));
  return mu + (sigma * (1.0 - std::pow(-std::log(canonical(g)), xi))) / xi;
}

// --------------------------------------------------------------------------------------------
// XXH3-64 with seed for inputs of 17..240 bytes (xxHash 0.8.3, XXH3_64bits_withSeed).
// GenomicInterval::hash feeds name ‖ u64 size ‖ u64 start ‖ u64 end (>= 25 bytes)
// (src/libmodle/internal/genome.cpp:201-224); the streaming API equals the one-shot hash.
// --------------------------------------------------------------------------------------------
namespace xxh3 {
static constexpr unsigned char kSecret[192] = {
    0xb8, 0xfe, 0x6c, 0x39, 0x23, 0xa4, 0x4b, 0xbe, 0x7c, 0x01, 0x81, 0x2c, 0xf7, 0x21, 0xad,
    0x1c, 0xde, 0xd4, 0x6d, 0xe9, 0x83, 0x90, 0x97, 0xdb, 0x72, 0x40, 0xa4, 0xa4, 0xb7, 0xb3,
    0x67, 0x1f, 0xcb, 0x79, 0xe6, 0x4e, 0xcc, 0xc0, 0xe5, 0x78, 0x82, 0x5a, 0xd0, 0x7d, 0xcc,
    0xff, 0x72, 0x21, 0xb8, 0x08, 0x46, 0x74, 0xf7, 0x43, 0x24, 0x8e, 0xe0, 0x35, 0x90, 0xe6,
    0x81, 0x3a, 0x26, 0x4c, 0x3c, 0x28, 0x52, 0xbb, 0x91, 0xc3, 0x00, 0xcb, 0x88, 0xd0, 0x65,
    0x8b, 0x1b, 0x53, 0x2e, 0xa3, 0x71, 0x64, 0x48, 0x97, 0xa2, 0x0d, 0xf9, 0x4e, 0x38, 0x19,
    0xef, 0x46, 0xa9, 0xde, 0xac, 0xd8, 0xa8, 0xfa, 0x76, 0x3f, 0xe3, 0x9c, 0x34, 0x3f, 0xf9,
    0xdc, 0xbb, 0xc7, 0xc7, 0x0b, 0x4f, 0x1d, 0x8a, 0x51, 0xe0, 0x4b, 0xcd, 0xb4, 0x59, 0x31,
    0xc8, 0x9f, 0x7e, 0xc9, 0xd9, 0x78, 0x73, 0x64, 0xea, 0xc5, 0xac, 0x83, 0x34, 0xd3, 0xeb,
    0xc3, 0xc5, 0x81, 0xa0, 0xff, 0xfa, 0x13, 0x63, 0xeb, 0x17, 0x0d, 0xdd, 0x51, 0xb7, 0xf0,
    0xda, 0x49, 0xd3, 0x16, 0x55, 0x26, 0x29, 0xd4, 0x68, 0x9e, 0x2b, 0x16, 0xbe, 0x58, 0x7d,
    0x47, 0xa1, 0xfc, 0x8f, 0xf8, 0xb8, 0xd1, 0x7a, 0xd0, 0x31, 0xce, 0x45, 0xcb, 0x3a, 0x8f,
    0x95, 0x16, 0x04, 0x28, 0xaf, 0xd7, 0xfb, 0xca, 0xbb, 0x4b, 0x40, 0x7e,
};
inline u64 rd64(const unsigned char* p) noexcept {
  u64 v = 0;
  for (int i = 7; i >= 0; --i) v = (v << 8) | p[i];
  return v;
}
inline u64 mul128_fold64(u64 a, u64 b) noexcept {
  const unsigned __int128 m = static_cast<unsigned __int128>(a) * b;
  return static_cast<u64>(m) ^ static_cast<u64>(m >> 64);
}
inline u64 avalanche(u64 h) noexcept {
  h ^= h >> 37;
  h *= 0x165667919E3779F9ULL;
  h ^= h >> 32;
  return h;
}
inline u64 mix16(const unsigned char* in, const unsigned char* sec, u64 seed) noexcept {
  return mul128_fold64(rd64(in) ^ (rd64(sec) + seed), rd64(in + 8) ^ (rd64(sec + 8) - seed));
}
// returns false when len is outside the supported 17..240 byte range
inline bool hash64(const unsigned char* in, std::size_t len, u64 seed, u64* out) noexcept {
  constexpr u64 P1 = 0x9E3779B185EBCA87ULL;
  if (len < 17 || len > 240) return false;
  u64 acc = len * P1;
  if (len <= 128) {
    if (len > 32) {
      if (len > 64) {
        if (len > 96) {
          acc += mix16(in + 48, kSecret + 96, seed);
          acc += mix16(in + len - 64, kSecret + 112, seed);
        }
        acc += mix16(in + 32, kSecret + 64, seed);
        acc += mix16(in + len - 48, kSecret + 80, seed);
      }
      acc += mix16(in + 16, kSecret + 32, seed);
      acc += mix16(in + len - 32, kSecret + 48, seed);
    }
    acc += mix16(in, kSecret, seed);
    acc += mix16(in + len - 16, kSecret + 16, seed);
    *out = avalanche(acc);
    return true;
  }
  const std::size_t nrounds = len / 16;
  for (std::size_t i = 0; i < 8; ++i) acc += mix16(in + 16 * i, kSecret + 16 * i, seed);
  acc = avalanche(acc);
  for (std::size_t i = 8; i < nrounds; ++i)
    acc += mix16(in + 16 * i, kSecret + 16 * (i - 8) + 3, seed);
  acc += mix16(in + len - 16, kSecret + 136 - 17, seed);
  *out = avalanche(acc);
  return true;
}
}  // namespace xxh3

}  // namespace oracle
