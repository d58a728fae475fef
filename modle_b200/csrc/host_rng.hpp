// Host-side seeding arithmetic of the reference's task fan-out, written for the product library
// (independent of oracle/): XXH3-64 with seed, SplitMix64, xoshiro256++ and its 2^128 jump, plus
// the GF(2) transition-matrix machinery the device RNG uses to jump ahead by a fixed distance.
//
// Reference call sites (paths relative to the reference checkout):
//   random::PRNG(seed)                 src/common/include/modle/common/random.hpp:26-30
//   GenomicInterval::hash              src/libmodle/internal/genome.cpp:201-224
//   per-cell rand_eng.jump()           src/libmodle/cpu/scheduler_simulate.cpp:108,121,158
// The algorithms themselves live in un-vendored dependencies (xoshiro-cpp 1.1, xxHash 0.8.3) and
// are restated here from their public specifications.
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace modle_b200::host {

using u64 = std::uint64_t;

struct Xoshiro {
  u64 s[4];
  static constexpr u64 rotl(u64 x, int k) { return (x << k) | (x >> (64 - k)); }
  // linear engine step; the ++ scrambler only affects the output
  void step() {
    const u64 t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl(s[3], 45);
  }
  u64 next() {
    const u64 r = rotl(s[0] + s[3], 23) + s[0];
    step();
    return r;
  }
  static Xoshiro seeded(u64 seed) {
    Xoshiro g{};
    u64 x = seed;
    for (auto& w : g.s) {
      u64 z = (x += 0x9e3779b97f4a7c15ULL);
      z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
      z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
      w = z ^ (z >> 31);
    }
    return g;
  }
  void jump() {
    static constexpr u64 poly[4] = {0x180ec6d33cfd0abaULL, 0xd5a61266f0c9392cULL,
                                    0xa9582618e03fc9aaULL, 0x39abdc4529b1661cULL};
    u64 acc[4] = {0, 0, 0, 0};
    for (u64 w : poly) {
      for (int b = 0; b < 64; ++b) {
        if ((w >> b) & 1) {
          for (int i = 0; i < 4; ++i) acc[i] ^= s[i];
        }
        step();
      }
    }
    std::memcpy(s, acc, sizeof(acc));
  }
};

// 256x256 matrix over GF(2) acting on the engine state; column j = image of basis vector e_j,
// each column stored as 4 words (same layout as the state).
struct StateMatrix {
  std::vector<std::array<u64, 4>> col;  // 256 columns
  StateMatrix() : col(256) {}

  static StateMatrix one_step() {
    StateMatrix m;
    for (int j = 0; j < 256; ++j) {
      Xoshiro g{};
      g.s[0] = g.s[1] = g.s[2] = g.s[3] = 0;
      g.s[j / 64] = u64(1) << (j % 64);
      g.step();
      m.col[j] = {g.s[0], g.s[1], g.s[2], g.s[3]};
    }
    return m;
  }
  std::array<u64, 4> apply(const std::array<u64, 4>& v) const {
    std::array<u64, 4> r{0, 0, 0, 0};
    for (int j = 0; j < 256; ++j) {
      if ((v[j / 64] >> (j % 64)) & 1) {
        for (int i = 0; i < 4; ++i) r[i] ^= col[j][i];
      }
    }
    return r;
  }
  // this ∘ other  (apply `other` first)
  StateMatrix after(const StateMatrix& other) const {
    StateMatrix r;
    for (int j = 0; j < 256; ++j) r.col[j] = apply(other.col[j]);
    return r;
  }
  static StateMatrix power(u64 n) {
    StateMatrix result;
    for (int j = 0; j < 256; ++j) {
      result.col[j] = {0, 0, 0, 0};
      result.col[j][j / 64] = u64(1) << (j % 64);
    }
    StateMatrix base = one_step();
    while (n) {
      if (n & 1) result = base.after(result);
      n >>= 1;
      if (n) base = base.after(base);
    }
    return result;
  }
};

// ---- XXH3-64 (seeded) for 1..240-byte inputs; interval keys are name + 24 bytes --------------
namespace xxh3 {
inline const unsigned char* secret() {
  static const unsigned char k[192] = {
      0xb8, 0xfe, 0x6c, 0x39, 0x23, 0xa4, 0x4b, 0xbe, 0x7c, 0x01, 0x81, 0x2c, 0xf7, 0x21, 0xad, 0x1c,
      0xde, 0xd4, 0x6d, 0xe9, 0x83, 0x90, 0x97, 0xdb, 0x72, 0x40, 0xa4, 0xa4, 0xb7, 0xb3, 0x67, 0x1f,
      0xcb, 0x79, 0xe6, 0x4e, 0xcc, 0xc0, 0xe5, 0x78, 0x82, 0x5a, 0xd0, 0x7d, 0xcc, 0xff, 0x72, 0x21,
      0xb8, 0x08, 0x46, 0x74, 0xf7, 0x43, 0x24, 0x8e, 0xe0, 0x35, 0x90, 0xe6, 0x81, 0x3a, 0x26, 0x4c,
      0x3c, 0x28, 0x52, 0xbb, 0x91, 0xc3, 0x00, 0xcb, 0x88, 0xd0, 0x65, 0x8b, 0x1b, 0x53, 0x2e, 0xa3,
      0x71, 0x64, 0x48, 0x97, 0xa2, 0x0d, 0xf9, 0x4e, 0x38, 0x19, 0xef, 0x46, 0xa9, 0xde, 0xac, 0xd8,
      0xa8, 0xfa, 0x76, 0x3f, 0xe3, 0x9c, 0x34, 0x3f, 0xf9, 0xdc, 0xbb, 0xc7, 0xc7, 0x0b, 0x4f, 0x1d,
      0x8a, 0x51, 0xe0, 0x4b, 0xcd, 0xb4, 0x59, 0x31, 0xc8, 0x9f, 0x7e, 0xc9, 0xd9, 0x78, 0x73, 0x64,
      0xea, 0xc5, 0xac, 0x83, 0x34, 0xd3, 0xeb, 0xc3, 0xc5, 0x81, 0xa0, 0xff, 0xfa, 0x13, 0x63, 0xeb,
      0x17, 0x0d, 0xdd, 0x51, 0xb7, 0xf0, 0xda, 0x49, 0xd3, 0x16, 0x55, 0x26, 0x29, 0xd4, 0x68, 0x9e,
      0x2b, 0x16, 0xbe, 0x58, 0x7d, 0x47, 0xa1, 0xfc, 0x8f, 0xf8, 0xb8, 0xd1, 0x7a, 0xd0, 0x31, 0xce,
      0x45, 0xcb, 0x3a, 0x8f, 0x95, 0x16, 0x04, 0x28, 0xaf, 0xd7, 0xfb, 0xca, 0xbb, 0x4b, 0x40, 0x7e};
  return k;
}
inline u64 le64(const unsigned char* p) {
  u64 v;
  std::memcpy(&v, p, 8);  // little-endian hosts only (x86-64 / aarch64)
  return v;
}
inline u64 fold(u64 a, u64 b) {
  const unsigned __int128 m = static_cast<unsigned __int128>(a) * b;
  return static_cast<u64>(m) ^ static_cast<u64>(m >> 64);
}
inline u64 finish(u64 h) {
  h ^= h >> 37;
  h *= 0x165667919E3779F9ULL;
  return h ^ (h >> 32);
}
inline u64 mix(const unsigned char* d, const unsigned char* k, u64 seed) {
  return fold(le64(d) ^ (le64(k) + seed), le64(d + 8) ^ (le64(k + 8) - seed));
}
inline bool hash(const unsigned char* d, std::size_t len, u64 seed, u64* out) {
  const unsigned char* k = secret();
  if (len < 17 || len > 240) return false;
  u64 acc = len * 0x9E3779B185EBCA87ULL;
  if (len <= 128) {
    const std::size_t pairs = (len - 1) / 32;  // 0..3 extra (front, back) pairs
    for (std::size_t i = pairs + 1; i-- > 0;) {
      acc += mix(d + 16 * i, k + 32 * i, seed);
      acc += mix(d + len - 16 * (i + 1), k + 32 * i + 16, seed);
    }
    *out = finish(acc);
    return true;
  }
  for (std::size_t i = 0; i < 8; ++i) acc += mix(d + 16 * i, k + 16 * i, seed);
  acc = finish(acc);
  for (std::size_t i = 8; i < len / 16; ++i) acc += mix(d + 16 * i, k + 16 * (i - 8) + 3, seed);
  acc += mix(d + len - 16, k + 136 - 17, seed);
  *out = finish(acc);
  return true;
}
}  // namespace xxh3

}  // namespace modle_b200::host
