// CTA abstraction shared by the CUDA kernel and the host-side emulation used by the CPU tests.
//
// The per-cell simulation (sim_core.hpp) is written in bulk-synchronous style: a sequence of
// "thread regions" (every thread of the CTA runs the body once) separated by CTA-wide barriers,
// plus a handful of block-wide primitives (scans over one value per thread). On the device a
// region body runs once per hardware thread; in the emulation (MODLE_B200_EMU, plain g++) the
// region is a loop over virtual thread ids, so the same source is exercised on the CPU with any
// virtual CTA width. The emulation is TEST INFRASTRUCTURE (it lives behind tests/emu) and is
// never linked into the product library.
#pragma once
#include <cstdint>

#if defined(__CUDACC__) && !defined(MODLE_B200_EMU)
#define MB_DEVICE_BUILD 1
#define MB_FN __device__ __forceinline__
#define MB_HD __host__ __device__ __forceinline__
#define MB_FN_NOINLINE __device__ __noinline__
#else
#define MB_DEVICE_BUILD 0
#define MB_FN inline
#define MB_HD inline
#define MB_FN_NOINLINE inline
#include <algorithm>
#include <cmath>
#include <cstring>
#ifdef MODLE_B200_EMU_MT
#include <pthread.h>
#endif
#endif

// Two host-side emulations of the CTA (both TEST INFRASTRUCTURE, tests/emu):
//   serial (MODLE_B200_EMU):            a region is a loop over virtual thread ids in one OS thread
//   SPMD   (MODLE_B200_EMU + _EMU_MT):  every virtual thread is an OS thread running the whole
//       per-cell program, CTA barriers are pthread barriers -- built with -fsanitize=thread this is
//       a race detector for the kernel source (missing / misplaced barriers) that needs no GPU.
#if MB_DEVICE_BUILD || defined(MODLE_B200_EMU_MT)
#define MB_SPMD 1
#else
#define MB_SPMD 0
#endif

namespace modle_b200 {

using u8 = std::uint8_t;
using u16 = std::uint16_t;
using u32 = std::uint32_t;
using u64 = std::uint64_t;
using i32 = std::int32_t;
using i64 = std::int64_t;

constexpr int kMaxThreads = 1024;
constexpr int kMaxWarps = kMaxThreads / 32;

// Scratch the block-wide primitives need (lives in shared memory on the device).
struct CtaScratch {
  u64 warp_u64[kMaxWarps];
  u64 warp_u64b[kMaxWarps];
  u64 warp_u64c[kMaxWarps];
  u64 warp_u64d[kMaxWarps];
  u64 warp_u64e[kMaxWarps];
  u64 warp_u64f[kMaxWarps];
  double warp_f64[kMaxWarps];
  u64 bcast_u64;
  double bcast_f64;
};

// "Affine-min" element x -> min(a, x + b) over signed 64-bit; closed under composition. Used for
// the neighbour scans (adjust_moves, secondary collisions). `a` large = no cap.
struct MinPlus {
  i64 a;
  i64 b;
};
constexpr i64 kMinPlusInf = (i64(1) << 60);
MB_FN MinPlus minplus_identity() { return MinPlus{kMinPlusInf, 0}; }
// b == kMinPlusInf marks a constant map x -> a (used to restart a scan at a segment head).
MB_FN MinPlus minplus_const(i64 a) { return MinPlus{a, kMinPlusInf}; }
// apply `first`, then `second`
MB_FN MinPlus minplus_then(const MinPlus& first, const MinPlus& second) {
  if (second.b >= kMinPlusInf) return second;
  MinPlus r;
  const i64 a1b2 = first.a >= kMinPlusInf ? kMinPlusInf : first.a + second.b;
  r.a = second.a < a1b2 ? second.a : a1b2;
  r.b = first.b >= kMinPlusInf ? kMinPlusInf : first.b + second.b;
  if (r.a > kMinPlusInf) r.a = kMinPlusInf;
  return r;
}
MB_FN i64 minplus_apply(const MinPlus& f, i64 x) {
  if (f.b >= kMinPlusInf) return f.a;
  const i64 y = x + f.b;
  return f.a < y ? f.a : y;
}

// Element of the secondary-collision scan (sim_core.hpp secondary_pass): a map on the state
// (alive, v) of the unit just walked. kind COND: alive' = alive && v >= T, v' = min(a, v + b);
// CONST (b == kSecConst): alive' = 1, v' = a; DEAD (b == kSecDead): alive' = 0.
struct SecOp {
  i64 T, a, b;
};
constexpr i64 kSecConst = kMinPlusInf;
constexpr i64 kSecDead = kMinPlusInf + 1;
MB_FN SecOp secop_identity() { return SecOp{-kMinPlusInf, kMinPlusInf, 0}; }
MB_FN SecOp secop_const(i64 v) { return SecOp{0, v, kSecConst}; }
MB_FN SecOp secop_dead() { return SecOp{0, 0, kSecDead}; }
// apply `f`, then `g`
MB_FN SecOp secop_then(const SecOp& f, const SecOp& g) {
  if (g.b >= kMinPlusInf) return g;
  if (f.b == kSecDead) return f;
  if (f.b == kSecConst) {
    if (f.a < g.T) return secop_dead();
    const i64 y = f.a + g.b;
    return secop_const(g.a < y ? g.a : y);
  }
  if (f.a < g.T) return secop_dead();
  SecOp r;
  const i64 t2 = g.T - f.b;
  r.T = f.T > t2 ? f.T : t2;
  const i64 y = f.a >= kMinPlusInf ? kMinPlusInf : f.a + g.b;
  r.a = g.a < y ? g.a : y;
  r.b = f.b + g.b;
  return r;
}

// One value per thread that survives across regions: a register on the device, an array indexed
// by the virtual thread id in the emulation.
template <class T>
struct PerThread {
#if MB_SPMD
  T val;
  MB_FN explicit PerThread(int) : val() {}
  MB_FN T& operator[](int) { return val; }
  MB_FN const T& operator[](int) const { return val; }
#else
  T vals[kMaxThreads];
  explicit PerThread(int n) {
    for (int i = 0; i < n; ++i) vals[i] = T();
  }
  T& operator[](int t) { return vals[t]; }
  const T& operator[](int t) const { return vals[t]; }
#endif
};

#if MB_DEVICE_BUILD

struct Cta {
  CtaScratch* scr;
  MB_FN int nt() const { return static_cast<int>(blockDim.x); }
  MB_FN int first() const { return static_cast<int>(threadIdx.x); }
  MB_FN int step() const { return static_cast<int>(blockDim.x); }
  MB_FN bool leader(int tid) const { return tid == 0; }
  MB_FN void sync() const { __syncthreads(); }
  // x / nt(); the launcher only uses power-of-two CTA widths
  MB_FN u32 div_nt(u32 x) const { return x >> (31 - __clz(static_cast<int>(blockDim.x))); }

  // exclusive prefix sum of one u64 per thread; *total receives the CTA-wide sum
  MB_FN u64 exscan_sum(int tid, u64 v, u64* total) const {
    const int lane = tid & 31, warp = tid >> 5, nw = (nt() + 31) >> 5;
    u64 inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const u64 o = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += o;
    }
    if (lane == 31) scr->warp_u64[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      u64 w = lane < nw ? scr->warp_u64[lane] : 0;
      u64 winc = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const u64 o = __shfl_up_sync(0xffffffffu, winc, d);
        if (lane >= d) winc += o;
      }
      if (lane < nw) scr->warp_u64[lane] = winc - w;
      if (lane == 31) scr->bcast_u64 = winc;
    }
    __syncthreads();
    const u64 res = scr->warp_u64[warp] + inc - v;
    if (total) *total = scr->bcast_u64;
    __syncthreads();
    return res;
  }

  MB_FN u64 reduce_max(int tid, u64 v) const {
    const int lane = tid & 31, warp = tid >> 5, nw = (nt() + 31) >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const u64 o = __shfl_xor_sync(0xffffffffu, v, d);
      v = o > v ? o : v;
    }
    if (lane == 0) scr->warp_u64[warp] = v;
    __syncthreads();
    if (warp == 0) {
      u64 w = lane < nw ? scr->warp_u64[lane] : 0;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        const u64 o = __shfl_xor_sync(0xffffffffu, w, d);
        w = o > w ? o : w;
      }
      if (lane == 0) scr->bcast_u64 = w;
    }
    __syncthreads();
    const u64 r = scr->bcast_u64;
    __syncthreads();
    return r;
  }
  MB_FN u64 reduce_min(int tid, u64 v) const { return ~reduce_max(tid, ~v); }
  MB_FN u64 reduce_sum(int tid, u64 v) const {
    u64 total;
    exscan_sum(tid, v, &total);
    return total;
  }

  // Sum of one double per thread in a FIXED order: lanes pairwise (xor butterfly is symmetric, so
  // every lane ends with the same value), then warps 0..nw-1 left to right.
  MB_FN double reduce_sum_f64(int tid, double v) const {
    const int lane = tid & 31, warp = tid >> 5, nw = (nt() + 31) >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) v = v + __shfl_xor_sync(0xffffffffu, v, d);
    if (lane == 0) scr->warp_f64[warp] = v;
    __syncthreads();
    if (tid == 0) {
      double acc = 0.0;
      for (int w = 0; w < nw; ++w) acc = acc + scr->warp_f64[w];
      scr->bcast_f64 = acc;
    }
    __syncthreads();
    const double r = scr->bcast_f64;
    __syncthreads();
    return r;
  }

  // Exclusive scan of MinPlus elements in thread order: returns the composition of the elements
  // of all lower-numbered threads (applied lowest thread first).
  MB_FN MinPlus exscan_minplus(int tid, MinPlus v) const {
    const int lane = tid & 31, warp = tid >> 5, nw = (nt() + 31) >> 5;
    MinPlus inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      MinPlus o;
      o.a = __shfl_up_sync(0xffffffffu, inc.a, d);
      o.b = __shfl_up_sync(0xffffffffu, inc.b, d);
      if (lane >= d) inc = minplus_then(o, inc);
    }
    if (lane == 31) {
      scr->warp_u64[warp] = static_cast<u64>(inc.a);
      scr->warp_u64b[warp] = static_cast<u64>(inc.b);
    }
    __syncthreads();
    if (tid == 0) {
      MinPlus acc = minplus_identity();
      for (int w = 0; w < nw; ++w) {
        const MinPlus cur{static_cast<i64>(scr->warp_u64[w]), static_cast<i64>(scr->warp_u64b[w])};
        scr->warp_u64[w] = static_cast<u64>(acc.a);
        scr->warp_u64b[w] = static_cast<u64>(acc.b);
        acc = minplus_then(acc, cur);
      }
    }
    __syncthreads();
    const MinPlus wprefix{static_cast<i64>(scr->warp_u64[warp]),
                          static_cast<i64>(scr->warp_u64b[warp])};
    // exclusive within the warp: shift the inclusive value up by one lane
    MinPlus ex;
    ex.a = __shfl_up_sync(0xffffffffu, inc.a, 1);
    ex.b = __shfl_up_sync(0xffffffffu, inc.b, 1);
    if (lane == 0) ex = minplus_identity();
    const MinPlus res = minplus_then(wprefix, ex);
    __syncthreads();
    return res;
  }

  // Two independent MinPlus scans in one go (same barriers, twice the payload).
  MB_FN void exscan_minplus2(MinPlus& va, MinPlus& vb) const {
    const int tid = first();
    const int lane = tid & 31, warp = tid >> 5, nw = (nt() + 31) >> 5;
    MinPlus ia = va, ib = vb;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      MinPlus oa, ob;
      oa.a = __shfl_up_sync(0xffffffffu, ia.a, d);
      oa.b = __shfl_up_sync(0xffffffffu, ia.b, d);
      ob.a = __shfl_up_sync(0xffffffffu, ib.a, d);
      ob.b = __shfl_up_sync(0xffffffffu, ib.b, d);
      if (lane >= d) {
        ia = minplus_then(oa, ia);
        ib = minplus_then(ob, ib);
      }
    }
    if (lane == 31) {
      scr->warp_u64[warp] = static_cast<u64>(ia.a);
      scr->warp_u64b[warp] = static_cast<u64>(ia.b);
      scr->warp_u64c[warp] = static_cast<u64>(ib.a);
      scr->warp_u64d[warp] = static_cast<u64>(ib.b);
    }
    __syncthreads();
    if (tid < 2) {
      u64* a = tid == 0 ? scr->warp_u64 : scr->warp_u64c;
      u64* b = tid == 0 ? scr->warp_u64b : scr->warp_u64d;
      MinPlus acc = minplus_identity();
      for (int w = 0; w < nw; ++w) {
        const MinPlus cur{static_cast<i64>(a[w]), static_cast<i64>(b[w])};
        a[w] = static_cast<u64>(acc.a);
        b[w] = static_cast<u64>(acc.b);
        acc = minplus_then(acc, cur);
      }
    }
    __syncthreads();
    const MinPlus wa{static_cast<i64>(scr->warp_u64[warp]), static_cast<i64>(scr->warp_u64b[warp])};
    const MinPlus wb{static_cast<i64>(scr->warp_u64c[warp]), static_cast<i64>(scr->warp_u64d[warp])};
    MinPlus ea, eb;
    ea.a = __shfl_up_sync(0xffffffffu, ia.a, 1);
    ea.b = __shfl_up_sync(0xffffffffu, ia.b, 1);
    eb.a = __shfl_up_sync(0xffffffffu, ib.a, 1);
    eb.b = __shfl_up_sync(0xffffffffu, ib.b, 1);
    if (lane == 0) {
      ea = minplus_identity();
      eb = minplus_identity();
    }
    va = minplus_then(wa, ea);
    vb = minplus_then(wb, eb);
    __syncthreads();
  }

  // Two independent exclusive prefix-MIN scans of one i32 per thread (thread order), in one go.
  // The first thread receives INT32_MAX.
  MB_FN void exscan_min2_i32(PerThread<i32>& pa, PerThread<i32>& pb) const {
    const int tid = first();
    const int lane = tid & 31, warp = tid >> 5, nw = (nt() + 31) >> 5;
    i32 a = pa.val, b = pb.val;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const i32 oa = __shfl_up_sync(0xffffffffu, a, d);
      const i32 ob = __shfl_up_sync(0xffffffffu, b, d);
      if (lane >= d) {
        a = oa < a ? oa : a;
        b = ob < b ? ob : b;
      }
    }
    if (lane == 31)
      scr->warp_u64[warp] = (u64(static_cast<u32>(a)) << 32) | u64(static_cast<u32>(b));
    __syncthreads();
    if (warp == 0) {
      const u64 w = lane < nw ? scr->warp_u64[lane] : ~u64(0);
      i32 wa = lane < nw ? static_cast<i32>(static_cast<u32>(w >> 32)) : 0x7FFFFFFF;
      i32 wb = lane < nw ? static_cast<i32>(static_cast<u32>(w)) : 0x7FFFFFFF;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const i32 oa = __shfl_up_sync(0xffffffffu, wa, d);
        const i32 ob = __shfl_up_sync(0xffffffffu, wb, d);
        if (lane >= d) {
          wa = oa < wa ? oa : wa;
          wb = ob < wb ? ob : wb;
        }
      }
      // exclusive: warp w gets the inclusive value of warp w - 1
      i32 ea = __shfl_up_sync(0xffffffffu, wa, 1);
      i32 eb = __shfl_up_sync(0xffffffffu, wb, 1);
      if (lane == 0) {
        ea = 0x7FFFFFFF;
        eb = 0x7FFFFFFF;
      }
      if (lane < nw)
        scr->warp_u64b[lane] = (u64(static_cast<u32>(ea)) << 32) | u64(static_cast<u32>(eb));
    }
    __syncthreads();
    const u64 wp = scr->warp_u64b[warp];
    const i32 wpa = static_cast<i32>(static_cast<u32>(wp >> 32));
    const i32 wpb = static_cast<i32>(static_cast<u32>(wp));
    i32 xa = __shfl_up_sync(0xffffffffu, a, 1);
    i32 xb = __shfl_up_sync(0xffffffffu, b, 1);
    if (lane == 0) {
      xa = 0x7FFFFFFF;
      xb = 0x7FFFFFFF;
    }
    pa.val = wpa < xa ? wpa : xa;
    pb.val = wpb < xb ? wpb : xb;
    __syncthreads();
  }

  // ---- collectives over PerThread values (uniform results returned to every thread) ----
  MB_FN u64 exscan_sum(PerThread<u64>& v) const {
    u64 total;
    v.val = exscan_sum(first(), v.val, &total);
    return total;
  }
  MB_FN u64 reduce_max(const PerThread<u64>& v) const { return reduce_max(first(), v.val); }
  MB_FN u64 reduce_min(const PerThread<u64>& v) const { return reduce_min(first(), v.val); }
  MB_FN u64 reduce_sum(const PerThread<u64>& v) const { return reduce_sum(first(), v.val); }
  MB_FN double reduce_sum_f64(const PerThread<double>& v) const {
    return reduce_sum_f64(first(), v.val);
  }
  MB_FN void exscan_minplus(PerThread<MinPlus>& v) const { v.val = exscan_minplus(first(), v.val); }
  MB_FN void exscan_minplus2(PerThread<MinPlus>& a, PerThread<MinPlus>& b) const {
    exscan_minplus2(a.val, b.val);
  }

  // Two independent SecOp scans in one go (same barriers, twice the payload).
  MB_FN void exscan_secop2(PerThread<SecOp>& pa, PerThread<SecOp>& pb) const {
    const int tid = first();
    const int lane = tid & 31, warp = tid >> 5, nw = (nt() + 31) >> 5;
    SecOp ia = pa.val, ib = pb.val;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      SecOp oa, ob;
      oa.T = __shfl_up_sync(0xffffffffu, ia.T, d);
      oa.a = __shfl_up_sync(0xffffffffu, ia.a, d);
      oa.b = __shfl_up_sync(0xffffffffu, ia.b, d);
      ob.T = __shfl_up_sync(0xffffffffu, ib.T, d);
      ob.a = __shfl_up_sync(0xffffffffu, ib.a, d);
      ob.b = __shfl_up_sync(0xffffffffu, ib.b, d);
      if (lane >= d) {
        ia = secop_then(oa, ia);
        ib = secop_then(ob, ib);
      }
    }
    if (lane == 31) {
      scr->warp_u64[warp] = static_cast<u64>(ia.T);
      scr->warp_u64b[warp] = static_cast<u64>(ia.a);
      scr->warp_u64c[warp] = static_cast<u64>(ia.b);
      scr->warp_u64d[warp] = static_cast<u64>(ib.T);
      scr->warp_u64e[warp] = static_cast<u64>(ib.a);
      scr->warp_u64f[warp] = static_cast<u64>(ib.b);
    }
    __syncthreads();
    if (tid < 2) {  // thread 0 chains the first scan's warp totals, thread 1 the second's
      u64* t = tid == 0 ? scr->warp_u64 : scr->warp_u64d;
      u64* a = tid == 0 ? scr->warp_u64b : scr->warp_u64e;
      u64* b = tid == 0 ? scr->warp_u64c : scr->warp_u64f;
      SecOp acc = secop_identity();
      for (int w = 0; w < nw; ++w) {
        const SecOp cur{static_cast<i64>(t[w]), static_cast<i64>(a[w]), static_cast<i64>(b[w])};
        t[w] = static_cast<u64>(acc.T);
        a[w] = static_cast<u64>(acc.a);
        b[w] = static_cast<u64>(acc.b);
        acc = secop_then(acc, cur);
      }
    }
    __syncthreads();
    const SecOp wa{static_cast<i64>(scr->warp_u64[warp]), static_cast<i64>(scr->warp_u64b[warp]),
                   static_cast<i64>(scr->warp_u64c[warp])};
    const SecOp wb{static_cast<i64>(scr->warp_u64d[warp]), static_cast<i64>(scr->warp_u64e[warp]),
                   static_cast<i64>(scr->warp_u64f[warp])};
    SecOp ea, eb;
    ea.T = __shfl_up_sync(0xffffffffu, ia.T, 1);
    ea.a = __shfl_up_sync(0xffffffffu, ia.a, 1);
    ea.b = __shfl_up_sync(0xffffffffu, ia.b, 1);
    eb.T = __shfl_up_sync(0xffffffffu, ib.T, 1);
    eb.a = __shfl_up_sync(0xffffffffu, ib.a, 1);
    eb.b = __shfl_up_sync(0xffffffffu, ib.b, 1);
    if (lane == 0) {
      ea = secop_identity();
      eb = secop_identity();
    }
    pa.val = secop_then(wa, ea);
    pb.val = secop_then(wb, eb);
    __syncthreads();
  }

  // Exclusive scan of SecOp elements in thread order (composition of all lower threads).
  MB_FN void exscan_secop(PerThread<SecOp>& pv) const {
    const int tid = first();
    const int lane = tid & 31, warp = tid >> 5, nw = (nt() + 31) >> 5;
    SecOp inc = pv.val;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      SecOp o;
      o.T = __shfl_up_sync(0xffffffffu, inc.T, d);
      o.a = __shfl_up_sync(0xffffffffu, inc.a, d);
      o.b = __shfl_up_sync(0xffffffffu, inc.b, d);
      if (lane >= d) inc = secop_then(o, inc);
    }
    if (lane == 31) {
      scr->warp_u64[warp] = static_cast<u64>(inc.T);
      scr->warp_u64b[warp] = static_cast<u64>(inc.a);
      scr->warp_u64c[warp] = static_cast<u64>(inc.b);
    }
    __syncthreads();
    if (tid == 0) {
      SecOp acc = secop_identity();
      for (int w = 0; w < nw; ++w) {
        const SecOp cur{static_cast<i64>(scr->warp_u64[w]), static_cast<i64>(scr->warp_u64b[w]),
                        static_cast<i64>(scr->warp_u64c[w])};
        scr->warp_u64[w] = static_cast<u64>(acc.T);
        scr->warp_u64b[w] = static_cast<u64>(acc.a);
        scr->warp_u64c[w] = static_cast<u64>(acc.b);
        acc = secop_then(acc, cur);
      }
    }
    __syncthreads();
    const SecOp wprefix{static_cast<i64>(scr->warp_u64[warp]), static_cast<i64>(scr->warp_u64b[warp]),
                        static_cast<i64>(scr->warp_u64c[warp])};
    SecOp ex;
    ex.T = __shfl_up_sync(0xffffffffu, inc.T, 1);
    ex.a = __shfl_up_sync(0xffffffffu, inc.a, 1);
    ex.b = __shfl_up_sync(0xffffffffu, inc.b, 1);
    if (lane == 0) ex = secop_identity();
    pv.val = secop_then(wprefix, ex);
    __syncthreads();
  }
};

#define MB_ATOMIC_MAX_U32(ptr, val) atomicMax((ptr), (val))
#define MB_ATOMIC_OR_U32(ptr, val) atomicOr((ptr), (val))
#define MB_ATOMIC_AND_U32(ptr, val) atomicAnd((ptr), (val))
#define MB_ATOMIC_ADD_U32(ptr, val) atomicAdd((ptr), (val))
#define MB_ATOMIC_ADD_U64(ptr, val) \
  atomicAdd(reinterpret_cast<unsigned long long*>(ptr), static_cast<unsigned long long>(val))
#define MB_U64_TO_F64(x) __ull2double_rn(x)
#define MB_UMULHI64(a, b) __umul64hi((a), (b))
#define MB_POPC(x) __popc(x)
#define MB_FFS(x) __ffs(static_cast<int>(x))

#else  // ---------------------------------------------------------------- emulations

// CTA barriers the device build would execute (explicit sync() calls plus the three inside every
// block-wide collective): the emulation counts them so that the barrier cost of an epoch can be
// compared between variants of the code without a GPU.
inline u64& emu_barrier_count() {
  static thread_local u64 n = 0;
  return n;
}

inline u64* emu_phase_barriers() {  // per phase of the epoch loop (sim_types.hpp kPh*)
  static thread_local u64 n[64] = {};
  return n;
}

// Order in which the emulation runs the virtual threads of a region: 0 = ascending, 1 =
// descending, 2 = a different pseudo-random permutation for every region. A region whose result
// depends on the order has a cross-thread dependency the device would need a barrier for, so the
// tests replay every case under all three (a race detector that needs no GPU).
inline int& emu_thread_order() {
  static thread_local int mode = 0;
  return mode;
}

// Tests only: make bind_lefs take its sequential redo (the path a uniform_int rejection -- one
// draw in 2^36 on a human chromosome -- sends it down) in every epoch although nothing was
// rejected; the results must not change.
inline int& emu_force_bind_redo() {
  static thread_local int on = 0;
  return on;
}

#ifdef MODLE_B200_EMU_MT  // ----------------------------------------------- SPMD emulation

// Collectives go through per-thread slots of the shared scratch (at most kMaxWarps = 32 virtual
// threads), bracketed by barriers like their device counterparts.
struct Cta {
  CtaScratch* scr;
  int nthreads;
  int my;                   // this OS thread's virtual thread id
  pthread_barrier_t* bar;
  int nt() const { return nthreads; }
  int first() const { return my; }
  int step() const { return nthreads; }
  bool leader(int tid) const { return tid == 0; }
  void sync() const {
    ++emu_barrier_count();
    pthread_barrier_wait(bar);
  }
  u32 div_nt(u32 x) const { return x / static_cast<u32>(nthreads); }

  u64 exscan_sum(PerThread<u64>& v) const {
    scr->warp_u64[my] = v.val;
    sync();
    u64 acc = 0, mine = 0;
    for (int t = 0; t < nthreads; ++t) {
      if (t == my) mine = acc;
      acc += scr->warp_u64[t];
    }
    sync();
    emu_barrier_count() += 1;  // the device version has three barriers
    v.val = mine;
    return acc;
  }
  u64 reduce_max(const PerThread<u64>& v) const {
    scr->warp_u64[my] = v.val;
    sync();
    u64 m = 0;
    for (int t = 0; t < nthreads; ++t) m = std::max(m, scr->warp_u64[t]);
    sync();
    emu_barrier_count() += 1;
    return m;
  }
  u64 reduce_min(const PerThread<u64>& v) const {
    PerThread<u64> w(0);
    w.val = ~v.val;
    return ~reduce_max(w);
  }
  u64 reduce_sum(const PerThread<u64>& v) const {
    PerThread<u64> w(0);
    w.val = v.val;
    return exscan_sum(w);
  }
  // same association order as the device (see the serial emulation below)
  double reduce_sum_f64(const PerThread<double>& v) const {
    scr->warp_f64[my] = v.val;
    sync();
    double lane[32];
    for (int l = 0; l < 32; ++l) lane[l] = l < nthreads ? scr->warp_f64[l] : 0.0;
    for (int d = 1; d < 32; d <<= 1) {
      double nxt[32];
      for (int l = 0; l < 32; ++l) nxt[l] = lane[l] + lane[l ^ d];
      for (int l = 0; l < 32; ++l) lane[l] = nxt[l];
    }
    sync();
    emu_barrier_count() += 1;
    return 0.0 + lane[0];
  }
  void exscan_minplus(PerThread<MinPlus>& v) const {
    scr->warp_u64[my] = static_cast<u64>(v.val.a);
    scr->warp_u64b[my] = static_cast<u64>(v.val.b);
    sync();
    MinPlus acc = minplus_identity(), mine = acc;
    for (int t = 0; t < nthreads; ++t) {
      if (t == my) mine = acc;
      acc = minplus_then(acc, MinPlus{static_cast<i64>(scr->warp_u64[t]),
                                      static_cast<i64>(scr->warp_u64b[t])});
    }
    sync();
    emu_barrier_count() += 1;
    v.val = mine;
  }
  void exscan_minplus2(PerThread<MinPlus>& a, PerThread<MinPlus>& b) const {
    exscan_minplus(a);
    exscan_minplus(b);
    emu_barrier_count() -= 3;  // one collective on the device
  }
  void exscan_min2_i32(PerThread<i32>& a, PerThread<i32>& b) const {
    scr->warp_u64[my] = static_cast<u64>(static_cast<u32>(a.val));
    scr->warp_u64b[my] = static_cast<u64>(static_cast<u32>(b.val));
    sync();
    i32 ma = 0x7FFFFFFF, mb = 0x7FFFFFFF;
    for (int t = 0; t < my; ++t) {
      ma = std::min(ma, static_cast<i32>(static_cast<u32>(scr->warp_u64[t])));
      mb = std::min(mb, static_cast<i32>(static_cast<u32>(scr->warp_u64b[t])));
    }
    sync();
    emu_barrier_count() += 1;  // the device version has three barriers
    a.val = ma;
    b.val = mb;
  }
  void exscan_secop(PerThread<SecOp>& v) const {
    scr->warp_u64[my] = static_cast<u64>(v.val.T);
    scr->warp_u64b[my] = static_cast<u64>(v.val.a);
    scr->warp_u64c[my] = static_cast<u64>(v.val.b);
    sync();
    SecOp acc = secop_identity(), mine = acc;
    for (int t = 0; t < nthreads; ++t) {
      if (t == my) mine = acc;
      acc = secop_then(acc, SecOp{static_cast<i64>(scr->warp_u64[t]),
                                  static_cast<i64>(scr->warp_u64b[t]),
                                  static_cast<i64>(scr->warp_u64c[t])});
    }
    sync();
    emu_barrier_count() += 1;
    v.val = mine;
  }
  void exscan_secop2(PerThread<SecOp>& a, PerThread<SecOp>& b) const {
    exscan_secop(a);
    exscan_secop(b);
    emu_barrier_count() -= 3;
  }
};

// relaxed atomics: ThreadSanitizer treats them as synchronisation-free but race-free accesses
inline u32 emu_atomic_max_u32(u32* p, u32 v) {
  u32 cur = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (cur < v && !__atomic_compare_exchange_n(p, &cur, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
  }
  return cur;
}
inline u32 emu_atomic_min_u32(u32* p, u32 v) {
  u32 cur = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (cur > v && !__atomic_compare_exchange_n(p, &cur, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
  }
  return cur;
}
#define MB_ATOMIC_MAX_U32(ptr, val) emu_atomic_max_u32((ptr), (val))
#define MB_ATOMIC_MIN_U32_(ptr, val) emu_atomic_min_u32((ptr), (val))
#define MB_ATOMIC_OR_U32(ptr, val) __atomic_fetch_or((ptr), (val), __ATOMIC_RELAXED)
#define MB_ATOMIC_AND_U32(ptr, val) __atomic_fetch_and((ptr), (val), __ATOMIC_RELAXED)
#define MB_ATOMIC_ADD_U32(ptr, val) __atomic_fetch_add((ptr), (val), __ATOMIC_RELAXED)
#define MB_ATOMIC_ADD_U64(ptr, val) __atomic_fetch_add((ptr), (val), __ATOMIC_RELAXED)
#define MB_U64_TO_F64(x) static_cast<double>(x)
#define MB_UMULHI64(a, b) \
  static_cast<unsigned long long>((static_cast<unsigned __int128>(a) * static_cast<unsigned __int128>(b)) >> 64)
#define MB_POPC(x) __builtin_popcount(x)
#define MB_FFS(x) __builtin_ffs(static_cast<int>(x))

#else  // ------------------------------------------------------------------ serial emulation


struct Cta {
  CtaScratch* scr;
  int nthreads;
  mutable u32 region_counter = 0;
  // The emulation runs a region as a loop over virtual threads, so "one value per thread"
  // primitives are split in two halves: put(tid, v) inside one region, get(tid) in a later one.
  // k-th virtual thread to run in the current region (k == nthreads ends the region)
  int order(int k) const {
    if (k >= nthreads) return nthreads;
    const int mode = emu_thread_order();
    if (mode == 0) return k;
    if (mode == 1) return nthreads - 1 - k;
    if (k == 0) ++region_counter;
    // affine permutation k -> (a * k + b) mod n with a coprime to n
    static const u32 kMul[8] = {7, 11, 13, 17, 19, 23, 29, 31};
    const u32 n = static_cast<u32>(nthreads);
    u32 a = kMul[region_counter & 7];
    while (std::__gcd(a, n) != 1) ++a;
    const u32 b = (region_counter * 2654435761u) % n;
    return static_cast<int>((u64(a) * static_cast<u32>(k) + b) % n);
  }
  int nt() const { return nthreads; }
  int first() const { return 0; }
  int step() const { return 1; }
  bool leader(int tid) const { return tid == 0; }
  void sync() const { ++emu_barrier_count(); }
  u32 div_nt(u32 x) const { return x / static_cast<u32>(nthreads); }

  u64 exscan_sum(PerThread<u64>& v) const {
    emu_barrier_count() += 3;
    u64 acc = 0;
    for (int t = 0; t < nthreads; ++t) {
      const u64 x = v[t];
      v[t] = acc;
      acc += x;
    }
    return acc;
  }
  u64 reduce_max(const PerThread<u64>& v) const {
    emu_barrier_count() += 3;
    u64 m = 0;
    for (int t = 0; t < nthreads; ++t) m = std::max(m, v[t]);
    return m;
  }
  u64 reduce_min(const PerThread<u64>& v) const {
    emu_barrier_count() += 3;
    u64 m = ~u64(0);
    for (int t = 0; t < nthreads; ++t) m = std::min(m, v[t]);
    return m;
  }
  u64 reduce_sum(const PerThread<u64>& v) const {
    emu_barrier_count() += 3;
    u64 acc = 0;
    for (int t = 0; t < nthreads; ++t) acc += v[t];
    return acc;
  }
  // same association order as the device: xor-butterfly inside each group of 32, then groups
  // left to right
  double reduce_sum_f64(const PerThread<double>& v) const {
    emu_barrier_count() += 3;
    double total = 0.0;
    for (int w = 0; w * 32 < nthreads; ++w) {
      double lane[32];
      for (int l = 0; l < 32; ++l) lane[l] = (w * 32 + l < nthreads) ? v[w * 32 + l] : 0.0;
      for (int d = 1; d < 32; d <<= 1) {
        double nxt[32];
        for (int l = 0; l < 32; ++l) nxt[l] = lane[l] + lane[l ^ d];
        for (int l = 0; l < 32; ++l) lane[l] = nxt[l];
      }
      total = total + lane[0];
    }
    return total;
  }
  void exscan_minplus(PerThread<MinPlus>& v) const {
    emu_barrier_count() += 3;
    MinPlus acc = minplus_identity();
    for (int t = 0; t < nthreads; ++t) {
      const MinPlus x = v[t];
      v[t] = acc;
      acc = minplus_then(acc, x);
    }
  }
  void exscan_minplus2(PerThread<MinPlus>& a, PerThread<MinPlus>& b) const {
    exscan_minplus(a);
    exscan_minplus(b);
    emu_barrier_count() -= 3;  // one collective on the device
  }
  void exscan_min2_i32(PerThread<i32>& a, PerThread<i32>& b) const {
    emu_barrier_count() += 3;
    i32 ma = 0x7FFFFFFF, mb = 0x7FFFFFFF;
    for (int t = 0; t < nthreads; ++t) {
      const i32 xa = a[t], xb = b[t];
      a[t] = ma;
      b[t] = mb;
      ma = std::min(ma, xa);
      mb = std::min(mb, xb);
    }
  }
  void exscan_secop2(PerThread<SecOp>& a, PerThread<SecOp>& b) const {
    exscan_secop(a);
    exscan_secop(b);
    emu_barrier_count() -= 3;  // one collective on the device
  }
  void exscan_secop(PerThread<SecOp>& v) const {
    emu_barrier_count() += 3;
    SecOp acc = secop_identity();
    for (int t = 0; t < nthreads; ++t) {
      const SecOp x = v[t];
      v[t] = acc;
      acc = secop_then(acc, x);
    }
  }
};

#define MB_ATOMIC_MAX_U32(ptr, val) (*(ptr) = std::max<u32>(*(ptr), (val)))
#define MB_ATOMIC_OR_U32(ptr, val) (*(ptr) |= (val))
#define MB_ATOMIC_AND_U32(ptr, val) (*(ptr) &= (val))
#define MB_ATOMIC_ADD_U32(ptr, val) (*(ptr) += (val))
#define MB_ATOMIC_ADD_U64(ptr, val) (*(ptr) += (val))
#define MB_U64_TO_F64(x) static_cast<double>(x)
#define MB_UMULHI64(a, b) \
  static_cast<unsigned long long>((static_cast<unsigned __int128>(a) * static_cast<unsigned __int128>(b)) >> 64)
#define MB_POPC(x) __builtin_popcount(x)
#define MB_FFS(x) __builtin_ffs(static_cast<int>(x))

#endif  // serial / SPMD emulation
#endif

// Stores and loads of words that several threads may touch in the same region BY DESIGN (flags
// every writer sets to the same value; words whose other bits a neighbour updates). Plain
// accesses on the device; relaxed atomics in the SPMD emulation, so that the race detector
// reports only what is not meant to happen.
#ifdef MODLE_B200_EMU_MT
#define MB_SHARED_STORE_U32(ptr, val) __atomic_store_n((ptr), (val), __ATOMIC_RELAXED)
#define MB_SHARED_LOAD_U32(ptr) __atomic_load_n((ptr), __ATOMIC_RELAXED)
#define MB_SHARED_OR_U32(ptr, val) __atomic_fetch_or((ptr), (val), __ATOMIC_RELAXED)
#else
#define MB_SHARED_STORE_U32(ptr, val) (*(ptr) = (val))
#define MB_SHARED_LOAD_U32(ptr) (*(ptr))
#define MB_SHARED_OR_U32(ptr, val) (*(ptr) |= (val))
#endif

// A region: every thread of the CTA executes the body once, then the CTA synchronises.
#if MB_SPMD
#define MB_REGION(cta, tid) for (int tid = (cta).first(); tid < (cta).nt(); tid += (cta).step())
#else
#define MB_REGION(cta, tid)                                                                  \
  for (int mb_k_ = 0, tid = (cta).order(0); mb_k_ < (cta).nt(); ++mb_k_, tid = (cta).order(mb_k_))
#endif

}  // namespace modle_b200
