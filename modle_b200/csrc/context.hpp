// The context behind the opaque modle_b200_context handle of the C ABI, shared by the translation
// units that implement device entry points (kernels.cu, pixels.cu).
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "../../include/modle_b200.h"
#include "launch_prep.hpp"
#include "status.hpp"

namespace modle_b200 {

#define CUDA_TRY(expr)                                                                   \
  do {                                                                                   \
    const cudaError_t e_ = (expr);                                                       \
    if (e_ != cudaSuccess)                                                               \
      return modle_b200::fail(MODLE_B200_ERR_CUDA,                                       \
                              std::string(#expr) + ": " + cudaGetErrorString(e_));       \
  } while (0)

struct PinnedBuf {  // page-locked host staging (device->host copies at full PCIe rate)
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    const cudaError_t e = cudaMallocHost(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  ~PinnedBuf() {
    if (p) cudaFreeHost(p);
  }
};

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    const cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  ~DevBuf() {
    if (p) cudaFree(p);
  }
};

}  // namespace modle_b200

using namespace modle_b200;

// Buffers one in-flight launch owns. A context keeps a few of them so that launches issued on
// different streams can overlap (the tail of one interval's cells then runs next to the head of
// the next interval's); a slot is reused once the launch that used it last has finished.
struct LaunchSlot {
  DevBuf d_queue, d_rings, d_states;                     // launch scratch
  DevBuf d_bar_pos, d_bar_dir, d_stp_a, d_stp_i, d_occ;  // per-interval arrays
  cudaEvent_t done = nullptr;
  bool in_flight = false;
};
constexpr int kLaunchSlots = 16;
constexpr int kPxSyncSlots = 8;

struct modle_b200_context {
  int device = 0;
  int num_sms = 0;
  size_t max_smem_optin = 0;
  cudaStream_t stream = nullptr;
  DevBuf d_jump[kJumpSlots];  // byte-indexed jump tables, built on first use
  ZigguratTables zig;
  DevBuf d_zig;    // nx, ny, ex, ey
  DevBuf d_phase;  // kNumPhases cycle accumulators
  LaunchSlot slots[kLaunchSlots];
  int next_slot = 0;
  // High-water sizes of the slot buffers. A slot that has to (re)allocate takes the largest size
  // any launch of this context has asked for so far, and idle slots are grown along: cudaFree /
  // cudaMalloc synchronise the device, so after the first (largest-first) launches no launch
  // should have to allocate again.
  size_t hw_rings = 0, hw_states = 0, hw_barriers = 0;
  DevBuf d_tasks, d_band, d_occ1d, d_stats, d_missed, d_snap_u64, d_snap_bar, d_log;
  PinnedBuf h_stage;  // band | occ1d | stats | missed of the host-buffer entry point
  DevBuf d_binned, d_tiles;  // binned contact register: pixel indices by tile, counts + cursors
  cudaEvent_t binned_done = nullptr;
  DevBuf d_px_band, d_px_rows, d_px_out;  // band -> pixels (pixels.cu), host-buffer entry point
  DevBuf d_px_sync[kPxSyncSlots];         // look-back scratch of k_count_row_pixels
  uint64_t px_calls = 0;
  size_t l2_bytes = 0;
  uint64_t launches = 0;
  int rng_mode = MODLE_B200_RNG_REFERENCE_ORDER;  // modle_b200_set_rng_mode
};

