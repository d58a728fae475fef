// Device half of the C ABI (include/modle_b200.h): CUDA kernels for sm_100a and their launchers.
//
//   k_simulate_cells      persistent CTAs; each CTA pops cells from a device-side queue and runs
//                         the whole per-cell loop (sim_core.hpp) with the cell's LEF / barrier
//                         state resident in shared memory. Replaces the reference worker's
//                         simulate_one_cell (src/libmodle/cpu/simulation.cpp:896-986).
//   k_register_contacts   the contact-register step in isolation: scatters (bin1, bin2) pairs
//                         into the banded matrix with red.global.add.u32
//                         (ContactMatrixDense::increment, contact_matrix_dense_safe_impl.hpp:54-89).
//
// There is no CPU fallback: without a CUDA device every entry point fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/modle_b200.h"
#include "launch_prep.hpp"
#include "sim_core.hpp"
#include "status.hpp"

namespace modle_b200 {

struct LaunchArgs {
  KernelParams kp;
  IntervalData D;
  Sinks K;
  const modle_b200_cell_task* tasks;
  modle_b200_cell_stats* stats;  // may be null
  u32 num_cells;
  u32* queue;       // device counter
  u64* ring_pool;   // grid * 2W
  u64* state_pool;  // grid * 4G
  // optional snapshot of cell 0 (debug)
  u64* phase_cycles;  // kNumPhases accumulators (SM clock cycles per phase, all cells)
  u64* snap_u64;  // 5 * n_lefs + 2
  u8* snap_bar;   // n_bar
};

// Three instantiations: <256, kSmallMinBlocks> for intervals whose state leaves room for three
// CTAs per SM, <512, 2> for two, <1024, 1> for the large ones (one CTA owns the SM's shared
// memory). Co-resident CTAs hide each other's CTA-barrier stalls.
#ifndef MODLE_B200_SMALL_MIN_BLOCKS
#define MODLE_B200_SMALL_MIN_BLOCKS 3
#endif
#ifndef MODLE_B200_LARGE_THREADS
#define MODLE_B200_LARGE_THREADS 1024
#endif
// kCtr selects the throughput mode (counter-based draws, sim_core.hpp) at compile time, so the
// deterministic kernels carry none of its code.
template <int kThreads, int kMinBlocks, bool kCtr = false>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
    k_simulate_cells(const __grid_constant__ LaunchArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  CellShared& S = *reinterpret_cast<CellShared*>(smem);
  CellArrays A = carve_cell_arrays(smem + ((sizeof(CellShared) + 15) / 16) * 16, a.kp.n_lefs,
                                   a.kp.n_bar, a.kp.lut_entries);
  A.rng_ring = a.ring_pool + size_t(blockIdx.x) * 2 * a.kp.rng_window;
  A.rng_state = a.state_pool + size_t(blockIdx.x) * 4 * a.kp.rng_gen_threads;
  __shared__ u32 s_cell;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_cell = atomicAdd(a.queue, 1u);
    __syncthreads();
    const u32 cell = s_cell;
    if (cell >= a.num_cells) break;
    const modle_b200_cell_task t = a.tasks[cell];
    const bool has_work = a.kp.stop_on_epochs || t.num_target_contacts != 0;
    CellTaskDev td;
    td.cell_id = t.cell_id;
    td.target_epochs = t.num_target_epochs;
    td.target_contacts = t.num_target_contacts;
    for (int i = 0; i < 4; ++i) td.rng_state[i] = t.rng_state[i];
    if (has_work) {
      Cta cta{&S.scratch};
      Sinks K = a.K;
      if (K.log) K.log += size_t(cell) * K.log_cap;
      CellSimT<kCtr> sim{a.kp, a.D, A, S, K, cta, td};
      sim.run();
    }
    __syncthreads();
    if (has_work && threadIdx.x < kNumPhases && a.phase_cycles)
      atomicAdd(reinterpret_cast<unsigned long long*>(a.phase_cycles + threadIdx.x),
                static_cast<unsigned long long>(S.phase_cycles[threadIdx.x]));
    if (threadIdx.x == 0 && a.stats) {
      modle_b200_cell_stats st;
      st.num_contacts = has_work ? S.num_contacts : 0;
      st.num_epochs = has_work ? S.epoch : 0;
      st.num_burnin_epochs = has_work ? S.num_burnin_epochs : 0;
      st.num_lef_updates = has_work ? S.lef_updates : 0;
      st.num_rng_draws = has_work ? S.rng_pos : 0;
      st.device_fault = has_work ? S.fault : 0;
      a.stats[cell] = st;
    }
    if (a.snap_u64 && cell == 0 && has_work) {
      const u32 n = a.kp.n_lefs;
      for (u32 i = threadIdx.x; i < n; i += blockDim.x) {
        a.snap_u64[i] = A.rev[i] == kUnbound ? ~u64(0) : u64(A.rev[i]);
        a.snap_u64[n + i] = A.fwd[i] == kUnbound ? ~u64(0) : u64(A.fwd[i]);
        a.snap_u64[2 * n + i] = A.ep[i] == kUnbound ? ~u64(0) : u64(A.ep[i]);
        a.snap_u64[3 * n + i] = A.rr[i];
        a.snap_u64[4 * n + i] = A.fr[i];
      }
      for (u32 i = threadIdx.x; i < a.kp.n_bar; i += blockDim.x) a.snap_bar[i] = A.bar_active[i];
      if (threadIdx.x == 0) {
        a.snap_u64[5 * size_t(n)] = S.num_active;
        a.snap_u64[5 * size_t(n) + 1] = S.burnin_completed;
      }
    }
  }
}

// Contact register: scatters (bin1, bin2) pairs into the band with RED.ADD.U32. Each thread
// takes four pairs per iteration through two 16-byte loads (streaming, evict-first) so that four
// independent reductions are in flight per thread; out-of-band pairs are counted once per warp.
// A pair whose larger bin lies outside the matrix (j >= ncols) would land past the buffer: the
// reference throws for it (bound_check_coords) before touching memory; here it is skipped and
// counted with the out-of-band pairs.
__device__ __forceinline__ u32 register_one(u32 b1, u32 b2, u32 nrows, u32 ncols,
                                            u32* __restrict__ band) {
  const u32 i = b1 > b2 ? b1 - b2 : b2 - b1;
  const u32 j = b1 > b2 ? b1 : b2;
  if (i >= nrows || j >= ncols) return 1;
  atomicAdd(band + (size_t(j) * nrows + i), 1u);  // result unused -> RED.E.ADD
  return 0;
}

__global__ void __launch_bounds__(256) k_register_contacts(const u32* __restrict__ bin1,
                                                           const u32* __restrict__ bin2, size_t n,
                                                           u32 nrows, u32 ncols,
                                                           u32* __restrict__ band,
                                                           u64* __restrict__ missed, int vec_ok) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  const size_t tid = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  u32 my_missed = 0;
  size_t done = 0;
  if (vec_ok) {
    const size_t n4 = n / 4;
    const uint4* v1 = reinterpret_cast<const uint4*>(bin1);
    const uint4* v2 = reinterpret_cast<const uint4*>(bin2);
    for (size_t e = tid; e < n4; e += stride) {
      const uint4 a = __ldcs(v1 + e), b = __ldcs(v2 + e);
      my_missed += register_one(a.x, b.x, nrows, ncols, band);
      my_missed += register_one(a.y, b.y, nrows, ncols, band);
      my_missed += register_one(a.z, b.z, nrows, ncols, band);
      my_missed += register_one(a.w, b.w, nrows, ncols, band);
    }
    done = n4 * 4;
  }
  for (size_t e = done + tid; e < n; e += stride)
    my_missed += register_one(__ldcs(bin1 + e), __ldcs(bin2 + e), nrows, ncols, band);
  const u32 warp_missed = __reduce_add_sync(0xffffffffu, my_missed);
  if ((threadIdx.x & 31) == 0 && warp_missed)
    atomicAdd(reinterpret_cast<unsigned long long*>(missed),
              static_cast<unsigned long long>(warp_missed));
}

// Calibration of the "atomic roofline" (regime 2), independent of the register
// kernels above: every thread derives pseudo-random pixel addresses from a counter hash (no
// memory is read) and issues plain red.global.add.u32 on them -- the rate random 4-byte
// reductions can reach over a footprint of `npx` words with nothing else in the way.
__global__ void __launch_bounds__(256) k_calibrate_red(u32* __restrict__ band, u64 npx, u64 n,
                                                       u64 seed) {
  const u64 stride = u64(gridDim.x) * blockDim.x;
  for (u64 e = u64(blockIdx.x) * blockDim.x + threadIdx.x; e < n; e += stride) {
    u64 z = (e + seed) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const u64 px = __umul64hi(z, npx);  // uniform in [0, npx)
    atomicAdd(band + px, 1u);
  }
}

// ---- binned contact register (bands much larger than the L2) ---------------------------------
// A band that does not fit the L2 turns every contact into a random DRAM read-modify-write
// (one 32-byte sector in, one out). The binned path first groups the contacts by 32 MB tile of
// the band (a counting pass, then a scatter of the 4-byte pixel indices through per-CTA
// shared-memory staging so that each tile receives coalesced runs), then replays the grouped
// indices in order: the tile being updated is L2 resident, so DRAM sees each tile once.
constexpr u32 kMaxTiles = 512;       // 2^32 pixels / 2^23 (tiles are at least 32 MB)
constexpr u32 kMinTileShift = 23;
constexpr u32 kBinThreads = 512;     // k_bin_scatter: one tile counter per thread
constexpr u32 kBinBatch = 16;        // contacts per thread per batch of k_bin_scatter
static_assert(kBinThreads == kMaxTiles, "k_bin_scatter scans one tile counter per thread");

__device__ __forceinline__ u32 band_pixel(u32 b1, u32 b2, u32 nrows, u32 ncols) {
  const u32 i = b1 > b2 ? b1 - b2 : b2 - b1;
  const u32 j = b1 > b2 ? b1 : b2;
  return (i >= nrows || j >= ncols) ? 0xFFFFFFFFu : j * nrows + i;
}

__global__ void __launch_bounds__(256) k_bin_count(const u32* __restrict__ bin1,
                                                   const u32* __restrict__ bin2, size_t n, u32 nrows,
                                                   u32 ncols, u32 tile_shift, u32* __restrict__ tile_counts,
                                                   u64* __restrict__ missed) {
  __shared__ u32 hist[kMaxTiles];
  for (u32 t = threadIdx.x; t < kMaxTiles; t += blockDim.x) hist[t] = 0;
  __syncthreads();
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  u32 my_missed = 0;
  for (size_t e = size_t(blockIdx.x) * blockDim.x + threadIdx.x; e < n; e += stride) {
    const u32 p = band_pixel(__ldcs(bin1 + e), __ldcs(bin2 + e), nrows, ncols);
    if (p == 0xFFFFFFFFu) {
      ++my_missed;
    } else {
      atomicAdd(&hist[p >> tile_shift], 1u);
    }
  }
  __syncthreads();
  for (u32 t = threadIdx.x; t < kMaxTiles; t += blockDim.x)
    if (hist[t]) atomicAdd(&tile_counts[t], hist[t]);
  const u32 warp_missed = __reduce_add_sync(0xffffffffu, my_missed);
  if ((threadIdx.x & 31) == 0 && warp_missed)
    atomicAdd(reinterpret_cast<unsigned long long*>(missed),
              static_cast<unsigned long long>(warp_missed));
}

// Exclusive scan of one value per thread over a CTA of kBinThreads threads; *total = the sum.
// Ends with the scratch free for reuse after the caller's next __syncthreads().
__device__ __forceinline__ u32 tiles_exscan(u32 mine, u32* warp_tot, u32* total) {
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u32 inc = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const u32 o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= static_cast<u32>(d)) inc += o;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  u32 wbase = 0, tot = 0;
#pragma unroll
  for (u32 w = 0; w < kBinThreads / 32; ++w) {
    const u32 v = warp_tot[w];
    if (w < warp) wbase += v;
    tot += v;
  }
  *total = tot;
  return wbase + inc - mine;
}

// exclusive scan of the tile counts into the tile cursors (one CTA, one tile per thread);
// tile_cursor[kMaxTiles] receives the total
__global__ void __launch_bounds__(kBinThreads) k_bin_offsets(const u32* __restrict__ tile_counts,
                                                             u32* __restrict__ tile_cursor) {
  __shared__ u32 warp_tot[kBinThreads / 32];
  const u32 c = tile_counts[threadIdx.x];
  u32 total;
  tile_cursor[threadIdx.x] = tiles_exscan(c, warp_tot, &total);
  if (threadIdx.x == 0) tile_cursor[kMaxTiles] = total;
}

__global__ void __launch_bounds__(kBinThreads) k_bin_scatter(const u32* __restrict__ bin1,
                                                             const u32* __restrict__ bin2, size_t n,
                                                             u32 nrows, u32 ncols, u32 tile_shift,
                                                             u32* __restrict__ tile_cursor,
                                                             u32* __restrict__ binned) {
  // Per batch of 512 x kBinBatch contacts: count per tile, exclusive scan of the counts (start of
  // each tile's run inside the batch), one global reservation per non-empty tile, sort the batch
  // by tile in shared memory, then copy it out so that consecutive threads write consecutive
  // addresses of a tile's run (~24 contacts per run at 350 tiles).
  __shared__ u32 cnt[kMaxTiles];     // contacts of this batch per tile, then fill cursor
  __shared__ u32 start[kMaxTiles];   // first slot of the tile's run in `sorted`
  __shared__ u32 gbase[kMaxTiles];   // where the run goes in `binned`
  __shared__ u32 sorted[kBinThreads * kBinBatch];
  __shared__ u32 warp_tot[kBinThreads / 32];
  const size_t batch = size_t(kBinThreads) * kBinBatch;
  cnt[threadIdx.x] = 0;
  __syncthreads();
  for (size_t b0 = size_t(blockIdx.x) * batch; b0 < n; b0 += size_t(gridDim.x) * batch) {
    u32 px[kBinBatch];
#pragma unroll
    for (u32 k = 0; k < kBinBatch; ++k) {
      const size_t e = b0 + size_t(k) * kBinThreads + threadIdx.x;  // coalesced reads
      px[k] = e < n ? band_pixel(__ldcs(bin1 + e), __ldcs(bin2 + e), nrows, ncols) : 0xFFFFFFFFu;
    }
#pragma unroll
    for (u32 k = 0; k < kBinBatch; ++k)
      if (px[k] != 0xFFFFFFFFu) atomicAdd(&cnt[px[k] >> tile_shift], 1u);
    __syncthreads();
    const u32 c = cnt[threadIdx.x];
    cnt[threadIdx.x] = 0;
    u32 total;
    const u32 ex = tiles_exscan(c, warp_tot, &total);
    start[threadIdx.x] = ex;
    gbase[threadIdx.x] = c ? atomicAdd(&tile_cursor[threadIdx.x], c) : 0u;
    __syncthreads();
#pragma unroll
    for (u32 k = 0; k < kBinBatch; ++k) {
      if (px[k] == 0xFFFFFFFFu) continue;
      const u32 t = px[k] >> tile_shift;
      sorted[start[t] + atomicAdd(&cnt[t], 1u)] = px[k];
    }
    __syncthreads();
    cnt[threadIdx.x] = 0;
    for (u32 q = threadIdx.x; q < total; q += kBinThreads) {
      const u32 v = sorted[q];
      const u32 t = v >> tile_shift;
      binned[gbase[t] + (q - start[t])] = v;
    }
    __syncthreads();
  }
}

// Replays the grouped indices tile by tile. The grid is launched cooperatively (all CTAs are
// resident) and paced by an arrival counter: a CTA starts tile k only after every CTA has finished
// tile k-2, so at most two tiles are being updated at any time and they stay L2 resident -- DRAM
// sees each tile's sectors once in and once out. Without the pacing the CTAs drift apart, the
// working set outgrows the L2 and every reduction becomes a DRAM round trip again.
__global__ void __launch_bounds__(256) k_scatter_binned(const u32* __restrict__ binned,
                                                        const u32* __restrict__ tile_counts,
                                                        const u32* __restrict__ tile_end,
                                                        u32 ntiles, u32* __restrict__ band,
                                                        u32* __restrict__ arrivals) {
  u32 k = 0;  // non-empty tiles done so far (the same sequence in every CTA)
  for (u32 t = 0; t < ntiles; ++t) {
    const u32 cnt = tile_counts[t];
    if (cnt == 0) continue;
    if (k >= 2) {
      if (threadIdx.x == 0) {
        const u32 want = (k - 1) * gridDim.x;
        u32 seen;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(arrivals) : "memory");
          if (seen < want) __nanosleep(64);
        } while (seen < want);
      }
      __syncthreads();
    }
    const u64 lo = u64(tile_end[t]) - cnt;
    const u64 a = lo + u64(cnt) * blockIdx.x / gridDim.x;
    const u64 z = lo + u64(cnt) * (blockIdx.x + 1) / gridDim.x;
    for (u64 e = a + threadIdx.x; e < z; e += blockDim.x) atomicAdd(band + __ldcs(binned + e), 1u);
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(arrivals, 1u);
    ++k;
  }
}

}  // namespace modle_b200

using namespace modle_b200;

#include "context.hpp"

extern "C" {

int modle_b200_init(modle_b200_context** out, int device) {
  if (!out) return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "ctx is NULL");
  *out = nullptr;
  int count = 0;
  const cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(MODLE_B200_ERR_NO_DEVICE,
                std::string("no CUDA device available (modle_b200 has no CPU fallback): ") +
                    cudaGetErrorString(e));
  if (device < 0 || device >= count)
    return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "device index out of range");
  CUDA_TRY(cudaSetDevice(device));
  auto* ctx = new modle_b200_context();
  ctx->device = device;
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  ctx->num_sms = prop.multiProcessorCount;
  ctx->l2_bytes = static_cast<size_t>(prop.l2CacheSize);
  CUDA_TRY(cudaEventCreateWithFlags(&ctx->binned_done, cudaEventDisableTiming));
  ctx->max_smem_optin = prop.sharedMemPerBlockOptin;
  CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  CUDA_TRY(ctx->d_zig.reserve(sizeof(double) * (129 * 2 + 257 * 2)));
  double* z = static_cast<double*>(ctx->d_zig.p);
  CUDA_TRY(cudaMemcpy(z, ctx->zig.nx, sizeof(double) * 129, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(z + 129, ctx->zig.ny, sizeof(double) * 129, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(z + 258, ctx->zig.ex, sizeof(double) * 257, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(z + 515, ctx->zig.ey, sizeof(double) * 257, cudaMemcpyHostToDevice));
  CUDA_TRY(ctx->d_phase.reserve(sizeof(u64) * kNumPhases));
  CUDA_TRY(cudaMemset(ctx->d_phase.p, 0, sizeof(u64) * kNumPhases));
  for (auto& sl : ctx->slots) CUDA_TRY(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
  // opt in to the full shared memory once (the attribute is per function and device, and several
  // contexts may launch concurrently with different sizes)
  {
    cudaFuncAttributes fa;
    CUDA_TRY(cudaFuncGetAttributes(&fa, k_simulate_cells<256, MODLE_B200_SMALL_MIN_BLOCKS>));
    CUDA_TRY(cudaFuncSetAttribute(k_simulate_cells<256, MODLE_B200_SMALL_MIN_BLOCKS>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(ctx->max_smem_optin - fa.sharedSizeBytes)));
    CUDA_TRY(cudaFuncGetAttributes(&fa, k_simulate_cells<MODLE_B200_LARGE_THREADS, 1>));
    CUDA_TRY(cudaFuncSetAttribute(k_simulate_cells<MODLE_B200_LARGE_THREADS, 1>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(ctx->max_smem_optin - fa.sharedSizeBytes)));
    CUDA_TRY(cudaFuncGetAttributes(&fa, k_simulate_cells<512, 2>));
    CUDA_TRY(cudaFuncSetAttribute(k_simulate_cells<512, 2>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(ctx->max_smem_optin - fa.sharedSizeBytes)));
    // throughput-mode instantiations (same static shared memory: one u32)
    CUDA_TRY(cudaFuncSetAttribute(k_simulate_cells<256, MODLE_B200_SMALL_MIN_BLOCKS, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(ctx->max_smem_optin - fa.sharedSizeBytes)));
    CUDA_TRY(cudaFuncSetAttribute(k_simulate_cells<MODLE_B200_LARGE_THREADS, 1, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(ctx->max_smem_optin - fa.sharedSizeBytes)));
    CUDA_TRY(cudaFuncSetAttribute(k_simulate_cells<512, 2, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(ctx->max_smem_optin - fa.sharedSizeBytes)));
    ctx->max_smem_optin -= fa.sharedSizeBytes;
  }
  *out = ctx;
  return MODLE_B200_OK;
}

void modle_b200_destroy(modle_b200_context* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (auto& sl : ctx->slots)
    if (sl.done) cudaEventDestroy(sl.done);
  if (ctx->binned_done) cudaEventDestroy(ctx->binned_done);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int modle_b200_launch_geometry(uint64_t num_lefs, uint64_t num_barriers, uint32_t* cta_threads,
                               uint32_t* cells_per_sm, uint64_t* shared_bytes_per_cell) {
  if (num_lefs == 0 || num_lefs >= 65535 || num_barriers >= (1u << 24))
    return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "num_lefs / num_barriers out of range");
  const StagingConfig sc =
      pick_staging(static_cast<u32>(num_lefs), static_cast<u32>(num_barriers));
  if (cta_threads) *cta_threads = sc.cta_threads;
  if (cells_per_sm)
    *cells_per_sm = sc.cells_per_sm == 3 ? MODLE_B200_SMALL_MIN_BLOCKS : sc.cells_per_sm;
  if (shared_bytes_per_cell)
    *shared_bytes_per_cell =
        ((sizeof(CellShared) + 15) / 16) * 16 +
        cell_array_bytes(static_cast<u32>(num_lefs), static_cast<u32>(num_barriers));
  return MODLE_B200_OK;
}

int modle_b200_set_rng_mode(modle_b200_context* ctx, int mode) {
  if (!ctx) return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "ctx is NULL");
  if (mode != MODLE_B200_RNG_REFERENCE_ORDER && mode != MODLE_B200_RNG_COUNTER)
    return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "unknown RNG mode");
  ctx->rng_mode = mode;
  return MODLE_B200_OK;
}

int modle_b200_get_rng_mode(const modle_b200_context* ctx) {
  return ctx ? ctx->rng_mode : MODLE_B200_RNG_REFERENCE_ORDER;
}

uint64_t modle_b200_kernel_launches(const modle_b200_context* ctx) {
  return ctx ? ctx->launches : 0;
}

int modle_b200_phase_cycles(modle_b200_context* ctx, uint64_t* out, size_t n, int reset) {
  if (!ctx || !out) return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "NULL argument");
  CUDA_TRY(cudaSetDevice(ctx->device));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  u64 h[kNumPhases];
  CUDA_TRY(cudaMemcpy(h, ctx->d_phase.p, sizeof(h), cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < n; ++i) out[i] = i < size_t(kNumPhases) ? h[i] : 0;
  if (reset) CUDA_TRY(cudaMemset(ctx->d_phase.p, 0, sizeof(h)));
  return MODLE_B200_OK;
}

int modle_b200_synchronize(modle_b200_context* ctx) {
  if (!ctx) return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "ctx is NULL");
  CUDA_TRY(cudaSetDevice(ctx->device));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  for (auto& sl : ctx->slots) {
    if (!sl.in_flight) continue;
    CUDA_TRY(cudaEventSynchronize(sl.done));
    sl.in_flight = false;
  }
  return MODLE_B200_OK;
}

}  // extern "C"

namespace {

// dst[i] += src[i] over a band (up to 2.9 GB): split over a few host threads, since one thread
// adds ~2 G words/s and the genome-wide run brings 370 M words back per step.
void add_into_u32(u32* dst, const u32* src, size_t n) {
  const size_t kMinPerThread = size_t(1) << 22;
  static const unsigned max_threads = [] {  // MODLE_B200_HOST_ADD_THREADS: measurement knob
    const char* e = std::getenv("MODLE_B200_HOST_ADD_THREADS");
    const int v = e ? std::atoi(e) : 4;
    return static_cast<unsigned>(std::max(1, std::min(64, v)));
  }();
  unsigned nt = std::thread::hardware_concurrency();
  nt = std::max(1u, std::min(max_threads, nt));
  nt = static_cast<unsigned>(std::min<size_t>(nt, std::max<size_t>(1, n / kMinPerThread)));
  auto work = [=](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; ++i) dst[i] += src[i];
  };
  if (nt <= 1) {
    work(0, n);
    return;
  }
  std::vector<std::thread> pool;
  for (unsigned t = 1; t < nt; ++t) pool.emplace_back(work, n * t / nt, n * (t + 1) / nt);
  work(0, n / nt);
  for (auto& th : pool) th.join();
}

void copy_into_u32(u32* dst, const u32* src, size_t n) {
  const size_t kMinPerThread = size_t(1) << 22;
  unsigned nt = std::max(1u, std::min(4u, std::thread::hardware_concurrency()));
  nt = static_cast<unsigned>(std::min<size_t>(nt, std::max<size_t>(1, n / kMinPerThread)));
  auto work = [=](size_t lo, size_t hi) { std::memcpy(dst + lo, src + lo, sizeof(u32) * (hi - lo)); };
  if (nt <= 1) {
    work(0, n);
    return;
  }
  std::vector<std::thread> pool;
  for (unsigned t = 1; t < nt; ++t) pool.emplace_back(work, n * t / nt, n * (t + 1) / nt);
  work(0, n / nt);
  for (auto& th : pool) th.join();
}

// Shared implementation of the two simulate entry points. All pointers are device pointers.
int launch_simulate(modle_b200_context* ctx, const modle_b200_sim_params* params,
                    const modle_b200_interval* interval, const modle_b200_barrier* h_barriers,
                    size_t num_barriers, const modle_b200_cell_task* d_tasks, size_t num_cells,
                    u32* d_band, u64* d_occ1d, modle_b200_cell_stats* d_stats, u64* d_missed,
                    cudaStream_t stream, u64* d_snap_u64, u8* d_snap_bar,
                    modle_b200_epoch_record* d_log = nullptr, u32 log_cap = 0) {
  const StagingConfig sc =
      pick_staging(static_cast<u32>(interval->num_lefs), static_cast<u32>(num_barriers));
  KernelParams kp;
  IntervalHostData hd;
  // shared memory a CTA of this launch class may use: what leaves room for 3 / 2 CTAs per SM,
  // or everything the device offers
  const size_t class_limit = sc.cells_per_sm == 3 ? size_t(75) * 1024
                             : sc.cells_per_sm == 2 ? size_t(113) * 1024 : ctx->max_smem_optin;
  const std::string err = prepare_interval(*params, *interval, h_barriers, num_barriers, sc, &kp,
                                           &hd, std::min(class_limit, ctx->max_smem_optin));
  if (!err.empty()) return fail(MODLE_B200_ERR_UNSUPPORTED, err);
  const size_t smem = ((sizeof(CellShared) + 15) / 16) * 16 +
                      cell_array_bytes(kp.n_lefs, kp.n_bar, kp.lut_entries);
  if (smem > ctx->max_smem_optin)
    return fail(MODLE_B200_ERR_UNSUPPORTED,
                "interval needs " + std::to_string(smem) +
                    " bytes of shared memory per cell; the device offers " +
                    std::to_string(ctx->max_smem_optin));
  if (num_cells == 0) return MODLE_B200_OK;
  if (num_cells > 0x7FFFFFFFull) return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "too many cells");

  if (!ctx->d_jump[sc.jump_slot].p) {
    std::vector<u64> table(kJumpTableWords);
    build_jump_table(sc.window, table.data());
    CUDA_TRY(ctx->d_jump[sc.jump_slot].reserve(sizeof(u64) * kJumpTableWords));
    CUDA_TRY(cudaMemcpy(ctx->d_jump[sc.jump_slot].p, table.data(), sizeof(u64) * kJumpTableWords,
                        cudaMemcpyHostToDevice));
  }
  // pick a launch slot: the idle one with the lowest index (a caller that waits for every launch
  // keeps reusing slot 0 and its buffers; overlapping launches take as many slots as are in
  // flight at once), else wait for the oldest
  LaunchSlot* sl = nullptr;
  for (int k = 0; k < kLaunchSlots && !sl; ++k) {
    LaunchSlot& c = ctx->slots[k];
    if (c.in_flight && cudaEventQuery(c.done) == cudaSuccess) c.in_flight = false;
    if (!c.in_flight) sl = &c;
  }
  if (!sl) {
    sl = &ctx->slots[ctx->next_slot];
    ctx->next_slot = (ctx->next_slot + 1) % kLaunchSlots;
    CUDA_TRY(cudaEventSynchronize(sl->done));
    sl->in_flight = false;
  }
  cudaGetLastError();  // cudaEventQuery's cudaErrorNotReady is not an error
  // per-interval arrays (pageable host -> device; small)
  const size_t nb = num_barriers;
  auto reserve_barriers = [&](LaunchSlot& x, size_t n1) -> cudaError_t {
    cudaError_t e = x.d_bar_pos.reserve(sizeof(u32) * n1);
    if (e == cudaSuccess) e = x.d_bar_dir.reserve(sizeof(u32) * (n1 / 32 + 2));
    if (e == cudaSuccess) e = x.d_stp_a.reserve(sizeof(double) * n1);
    if (e == cudaSuccess) e = x.d_stp_i.reserve(sizeof(double) * n1);
    if (e == cudaSuccess) e = x.d_occ.reserve(sizeof(double) * n1);
    return e;
  };
  if (nb + 1 > ctx->hw_barriers) {
    ctx->hw_barriers = nb + 1;
    for (auto& x : ctx->slots)  // (slots that never ran a launch allocate when they first do)
      if (!x.in_flight && &x != sl && x.d_bar_pos.p) CUDA_TRY(reserve_barriers(x, ctx->hw_barriers));
  }
  CUDA_TRY(reserve_barriers(*sl, ctx->hw_barriers));
  if (nb) {
    CUDA_TRY(cudaMemcpyAsync(sl->d_bar_pos.p, hd.bar_pos.data(), sizeof(u32) * nb,
                             cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaMemcpyAsync(sl->d_stp_a.p, hd.stp_active.data(), sizeof(double) * nb,
                             cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaMemcpyAsync(sl->d_stp_i.p, hd.stp_inactive.data(), sizeof(double) * nb,
                             cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaMemcpyAsync(sl->d_occ.p, hd.occupancy.data(), sizeof(double) * nb,
                             cudaMemcpyHostToDevice, stream));
  }
  CUDA_TRY(cudaMemcpyAsync(sl->d_bar_dir.p, hd.bar_dir_rev.data(),
                           sizeof(u32) * hd.bar_dir_rev.size(), cudaMemcpyHostToDevice, stream));

  // grid: persistent CTAs, as many as fit
  const bool ctr = ctx->rng_mode == MODLE_B200_RNG_COUNTER;
  void (*kernel)(const LaunchArgs) =
      sc.cells_per_sm == 3
          ? (ctr ? k_simulate_cells<256, MODLE_B200_SMALL_MIN_BLOCKS, true>
                 : k_simulate_cells<256, MODLE_B200_SMALL_MIN_BLOCKS>)
      : sc.cells_per_sm == 2
          ? (ctr ? k_simulate_cells<512, 2, true> : k_simulate_cells<512, 2>)
          : (ctr ? k_simulate_cells<MODLE_B200_LARGE_THREADS, 1, true>
                 : k_simulate_cells<MODLE_B200_LARGE_THREADS, 1>);
  int per_sm = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel,
                                                         static_cast<int>(sc.cta_threads), smem));
  if (per_sm < 1) return fail(MODLE_B200_ERR_UNSUPPORTED, "kernel does not fit on an SM");
  const u32 grid = static_cast<u32>(
      std::min<size_t>(num_cells, size_t(per_sm) * static_cast<size_t>(ctx->num_sms)));
  if (!ctr) {  // the throughput mode stages no draws
    const size_t rings = sizeof(u64) * 2 * size_t(sc.window) * grid;
    const size_t states = sizeof(u64) * 4 * size_t(sc.gen_threads) * grid;
    if (rings > ctx->hw_rings || states > ctx->hw_states) {
      ctx->hw_rings = std::max(ctx->hw_rings, rings);
      ctx->hw_states = std::max(ctx->hw_states, states);
      for (auto& x : ctx->slots) {
        if (x.in_flight || &x == sl || !x.d_rings.p) continue;
        CUDA_TRY(x.d_rings.reserve(ctx->hw_rings));
        CUDA_TRY(x.d_states.reserve(ctx->hw_states));
      }
    }
    CUDA_TRY(sl->d_rings.reserve(ctx->hw_rings));
    CUDA_TRY(sl->d_states.reserve(ctx->hw_states));
  }
  if (!sl->d_queue.p)
    for (auto& x : ctx->slots) CUDA_TRY(x.d_queue.reserve(sizeof(u32) * 4));
  CUDA_TRY(cudaMemsetAsync(sl->d_queue.p, 0, sizeof(u32) * 4, stream));

  LaunchArgs a;
  a.kp = kp;
  const double* z = static_cast<const double*>(ctx->d_zig.p);
  a.D.bar_pos = static_cast<const u32*>(sl->d_bar_pos.p);
  a.D.bar_dir_rev = static_cast<const u32*>(sl->d_bar_dir.p);
  a.D.bar_stp_active = static_cast<const double*>(sl->d_stp_a.p);
  a.D.bar_stp_inactive = static_cast<const double*>(sl->d_stp_i.p);
  a.D.bar_occupancy = static_cast<const double*>(sl->d_occ.p);
  a.D.zig_nx = z;
  a.D.zig_ny = z + 129;
  a.D.zig_ex = z + 258;
  a.D.zig_ey = z + 515;
  a.D.jump_tbl = static_cast<const u64*>(ctx->d_jump[sc.jump_slot].p);
  a.K.band = d_band;
  a.K.occ1d = kp.track_1d ? d_occ1d : nullptr;
  a.K.missed = d_missed;
  a.K.log = d_log;
  a.K.log_cap = log_cap;
  a.tasks = d_tasks;
  a.stats = d_stats;
  a.num_cells = static_cast<u32>(num_cells);
  a.queue = static_cast<u32*>(sl->d_queue.p);
  a.ring_pool = static_cast<u64*>(sl->d_rings.p);
  a.state_pool = static_cast<u64*>(sl->d_states.p);
  a.phase_cycles = static_cast<u64*>(ctx->d_phase.p);
  a.snap_u64 = d_snap_u64;
  a.snap_bar = d_snap_bar;
  kernel<<<grid, sc.cta_threads, smem, stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaEventRecord(sl->done, stream));
  sl->in_flight = true;
  ++ctx->launches;
  return MODLE_B200_OK;
}

}  // namespace

extern "C" {

int modle_b200_simulate_interval_device(modle_b200_context* ctx,
                                        const modle_b200_sim_params* params,
                                        const modle_b200_interval* interval,
                                        const modle_b200_barrier* barriers_host_or_dev,
                                        size_t num_barriers, const modle_b200_cell_task* d_tasks,
                                        size_t num_cells, uint32_t* d_band, uint64_t* d_occ1d,
                                        modle_b200_cell_stats* d_stats, uint64_t* d_missed_updates,
                                        void* cuda_stream) {
  if (!ctx || !params || !interval || !d_tasks || !d_band || !d_missed_updates)
    return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "NULL argument");
  CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
  // barriers may live on either side: stage a host copy (the launcher validates and repacks them)
  std::vector<modle_b200_barrier> hb(num_barriers);
  if (num_barriers) {
    cudaPointerAttributes attr;
    const cudaError_t e = cudaPointerGetAttributes(&attr, barriers_host_or_dev);
    if (e == cudaSuccess && attr.type == cudaMemoryTypeDevice) {
      CUDA_TRY(cudaMemcpy(hb.data(), barriers_host_or_dev,
                          sizeof(modle_b200_barrier) * num_barriers, cudaMemcpyDeviceToHost));
    } else {
      cudaGetLastError();
      std::memcpy(hb.data(), barriers_host_or_dev, sizeof(modle_b200_barrier) * num_barriers);
    }
  }
  return launch_simulate(ctx, params, interval, hb.data(), num_barriers, d_tasks, num_cells, d_band,
                         d_occ1d, d_stats, d_missed_updates, stream, nullptr, nullptr);
}

static int simulate_interval_host(modle_b200_context* ctx, const modle_b200_sim_params* params,
                                  const modle_b200_interval* interval,
                                  const modle_b200_barrier* barriers, size_t num_barriers,
                                  const modle_b200_cell_task* tasks, size_t num_cells,
                                  uint32_t* band_out, uint64_t* occ1d_out,
                                  modle_b200_cell_stats* stats_out, uint64_t* missed_updates_out,
                                  modle_b200_epoch_record* log_out, size_t log_cap,
                                  bool overwrite = false) {
  if (!ctx || !params || !interval || !tasks || !band_out)
    return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "NULL argument");
  if (log_cap != 0 && !log_out) return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "log_out is NULL");
  if (log_cap > 0xFFFFFFFFull)
    return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "log_capacity_per_cell too large");
  if (num_barriers && !barriers) return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "barriers is NULL");
  CUDA_TRY(cudaSetDevice(ctx->device));
  u64 nrows = 0, ncols = 0;
  modle_b200_band_shape(params, interval->end - interval->start, &nrows, &ncols);
  const size_t npx = nrows * ncols + 1;
  cudaStream_t s = ctx->stream;
  CUDA_TRY(ctx->d_tasks.reserve(sizeof(modle_b200_cell_task) * std::max<size_t>(num_cells, 1)));
  CUDA_TRY(ctx->d_band.reserve(sizeof(u32) * npx));
  CUDA_TRY(ctx->d_occ1d.reserve(sizeof(u64) * std::max<u64>(ncols, 1)));
  CUDA_TRY(ctx->d_stats.reserve(sizeof(modle_b200_cell_stats) * std::max<size_t>(num_cells, 1)));
  CUDA_TRY(ctx->d_missed.reserve(sizeof(u64)));
  CUDA_TRY(cudaMemcpyAsync(ctx->d_tasks.p, tasks, sizeof(modle_b200_cell_task) * num_cells,
                           cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemsetAsync(ctx->d_band.p, 0, sizeof(u32) * npx, s));
  CUDA_TRY(cudaMemsetAsync(ctx->d_occ1d.p, 0, sizeof(u64) * std::max<u64>(ncols, 1), s));
  CUDA_TRY(cudaMemsetAsync(ctx->d_stats.p, 0,
                           sizeof(modle_b200_cell_stats) * std::max<size_t>(num_cells, 1), s));
  CUDA_TRY(cudaMemsetAsync(ctx->d_missed.p, 0, sizeof(u64), s));
  const size_t log_bytes = sizeof(modle_b200_epoch_record) * log_cap * num_cells;
  if (log_bytes) {
    CUDA_TRY(ctx->d_log.reserve(log_bytes));
    CUDA_TRY(cudaMemsetAsync(ctx->d_log.p, 0, log_bytes, s));
  }
  const int rc = launch_simulate(
      ctx, params, interval, barriers, num_barriers,
      static_cast<const modle_b200_cell_task*>(ctx->d_tasks.p), num_cells,
      static_cast<u32*>(ctx->d_band.p), static_cast<u64*>(ctx->d_occ1d.p),
      static_cast<modle_b200_cell_stats*>(ctx->d_stats.p), static_cast<u64*>(ctx->d_missed.p), s,
      nullptr, nullptr, log_bytes ? static_cast<modle_b200_epoch_record*>(ctx->d_log.p) : nullptr,
      static_cast<u32>(log_cap));
  if (rc != MODLE_B200_OK) return rc;
  if (log_bytes)
    CUDA_TRY(cudaMemcpyAsync(log_out, ctx->d_log.p, log_bytes, cudaMemcpyDeviceToHost, s));
  // results are ADDED to the caller's buffers (the reference accumulates into a shared matrix)
  const size_t off_occ = ((sizeof(u32) * npx + 63) / 64) * 64;
  const size_t off_stats = off_occ + ((sizeof(u64) * ncols + 63) / 64) * 64;
  const size_t off_missed = off_stats + ((sizeof(modle_b200_cell_stats) * num_cells + 63) / 64) * 64;
  CUDA_TRY(ctx->h_stage.reserve(off_missed + 64));
  unsigned char* hs = static_cast<unsigned char*>(ctx->h_stage.p);
  const u32* h_band = reinterpret_cast<const u32*>(hs);
  const u64* h_occ = reinterpret_cast<const u64*>(hs + off_occ);
  const modle_b200_cell_stats* h_stats = reinterpret_cast<const modle_b200_cell_stats*>(hs + off_stats);
  const u64* h_missed = reinterpret_cast<const u64*>(hs + off_missed);
  CUDA_TRY(cudaMemcpyAsync(hs, ctx->d_band.p, sizeof(u32) * npx, cudaMemcpyDeviceToHost, s));
  if (occ1d_out)
    CUDA_TRY(cudaMemcpyAsync(hs + off_occ, ctx->d_occ1d.p, sizeof(u64) * ncols,
                             cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(hs + off_stats, ctx->d_stats.p,
                           sizeof(modle_b200_cell_stats) * num_cells, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(hs + off_missed, ctx->d_missed.p, sizeof(u64), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  static const bool skip_copy_out = [] {  // MODLE_B200_E2E_SKIP_COPY=1: measurement knob only
    const char* e = std::getenv("MODLE_B200_E2E_SKIP_COPY");
    return e && e[0] == '1';
  }();
  if (skip_copy_out) {
    // (diagnostic: how much of an end-to-end step is the host-side copy into the caller's arrays)
  } else if (overwrite) {  // the caller's buffers need not be zeroed (or even touched) beforehand
    copy_into_u32(band_out, h_band, npx);
    if (occ1d_out) std::memcpy(occ1d_out, h_occ, sizeof(u64) * ncols);
    if (missed_updates_out) *missed_updates_out = *h_missed;
  } else {
    add_into_u32(band_out, h_band, npx);
    if (occ1d_out)
      for (size_t i = 0; i < ncols; ++i) occ1d_out[i] += h_occ[i];
    if (missed_updates_out) *missed_updates_out += *h_missed;
  }
  u64 first_fault = 0;
  for (size_t i = 0; i < num_cells; ++i) {
    if (stats_out) stats_out[i] = h_stats[i];
    if (!first_fault && h_stats[i].device_fault) first_fault = h_stats[i].device_fault;
  }
  if (first_fault)
    return fail(MODLE_B200_ERR_DEVICE_FAULT,
                "kernel reported fault code " + std::to_string(first_fault));
  return MODLE_B200_OK;
}

// Sizes the buffers behind the host-buffer entry points once, for the largest interval the caller
// is going to pass (the reference sizes each worker's State the same way, up front:
// Simulation::State::resize_buffers, src/libmodle/cpu/simulation.cpp:603-627). Optional -- the
// buffers also grow on demand -- but growing means cudaFree / cudaFreeHost, which synchronise the
// whole device and stall the launches of the other worker contexts.
int modle_b200_reserve(modle_b200_context* ctx, uint64_t max_nrows, uint64_t max_ncols,
                       size_t max_cells) {
  if (!ctx) return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "ctx is NULL");
  CUDA_TRY(cudaSetDevice(ctx->device));
  const size_t npx = max_nrows * max_ncols + 1;
  const size_t cells = std::max<size_t>(max_cells, 1);
  CUDA_TRY(ctx->d_tasks.reserve(sizeof(modle_b200_cell_task) * cells));
  CUDA_TRY(ctx->d_band.reserve(sizeof(u32) * npx));
  CUDA_TRY(ctx->d_occ1d.reserve(sizeof(u64) * std::max<u64>(max_ncols, 1)));
  CUDA_TRY(ctx->d_stats.reserve(sizeof(modle_b200_cell_stats) * cells));
  CUDA_TRY(ctx->d_missed.reserve(sizeof(u64)));
  const size_t off_occ = ((sizeof(u32) * npx + 63) / 64) * 64;
  const size_t off_stats = off_occ + ((sizeof(u64) * max_ncols + 63) / 64) * 64;
  const size_t off_missed = off_stats + ((sizeof(modle_b200_cell_stats) * cells + 63) / 64) * 64;
  CUDA_TRY(ctx->h_stage.reserve(off_missed + 64));
  // the launch scratch of the slot a waiting caller keeps reusing, for the largest of the three
  // launch classes (RNG rings: 2 windows per resident CTA)
  if (ctx->rng_mode != MODLE_B200_RNG_COUNTER) {
    size_t rings = 0, states = 0;
    for (const StagingConfig& sc : {staging_small(), staging_mid(), staging_large()}) {
      const size_t grid = std::min<size_t>(cells, size_t(sc.cells_per_sm) * size_t(ctx->num_sms));
      rings = std::max(rings, sizeof(u64) * 2 * size_t(sc.window) * grid);
      states = std::max(states, sizeof(u64) * 4 * size_t(sc.gen_threads) * grid);
    }
    ctx->hw_rings = std::max(ctx->hw_rings, rings);
    ctx->hw_states = std::max(ctx->hw_states, states);
    CUDA_TRY(ctx->slots[0].d_rings.reserve(ctx->hw_rings));
    CUDA_TRY(ctx->slots[0].d_states.reserve(ctx->hw_states));
  }
  return MODLE_B200_OK;
}

int modle_b200_simulate_interval(modle_b200_context* ctx, const modle_b200_sim_params* params,
                                 const modle_b200_interval* interval,
                                 const modle_b200_barrier* barriers, size_t num_barriers,
                                 const modle_b200_cell_task* tasks, size_t num_cells,
                                 uint32_t* band_out, uint64_t* occ1d_out,
                                 modle_b200_cell_stats* stats_out, uint64_t* missed_updates_out) {
  return simulate_interval_host(ctx, params, interval, barriers, num_barriers, tasks, num_cells,
                                band_out, occ1d_out, stats_out, missed_updates_out, nullptr, 0);
}

int modle_b200_simulate_interval_overwrite(modle_b200_context* ctx,
                                           const modle_b200_sim_params* params,
                                           const modle_b200_interval* interval,
                                           const modle_b200_barrier* barriers, size_t num_barriers,
                                           const modle_b200_cell_task* tasks, size_t num_cells,
                                           uint32_t* band_out, uint64_t* occ1d_out,
                                           modle_b200_cell_stats* stats_out,
                                           uint64_t* missed_updates_out) {
  return simulate_interval_host(ctx, params, interval, barriers, num_barriers, tasks, num_cells,
                                band_out, occ1d_out, stats_out, missed_updates_out, nullptr, 0,
                                /*overwrite=*/true);
}

int modle_b200_simulate_interval_logged(modle_b200_context* ctx,
                                        const modle_b200_sim_params* params,
                                        const modle_b200_interval* interval,
                                        const modle_b200_barrier* barriers, size_t num_barriers,
                                        const modle_b200_cell_task* tasks, size_t num_cells,
                                        uint32_t* band_out, uint64_t* occ1d_out,
                                        modle_b200_cell_stats* stats_out,
                                        uint64_t* missed_updates_out,
                                        modle_b200_epoch_record* log_out,
                                        size_t log_capacity_per_cell) {
  return simulate_interval_host(ctx, params, interval, barriers, num_barriers, tasks, num_cells,
                                band_out, occ1d_out, stats_out, missed_updates_out, log_out,
                                log_capacity_per_cell);
}

int modle_b200_snapshot_cell(modle_b200_context* ctx, const modle_b200_sim_params* params,
                             const modle_b200_interval* interval,
                             const modle_b200_barrier* barriers, size_t num_barriers,
                             const modle_b200_cell_task* task, modle_b200_cell_snapshot* snapshot,
                             modle_b200_cell_stats* stats_out) {
  if (!ctx || !params || !interval || !task || !snapshot)
    return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "NULL argument");
  CUDA_TRY(cudaSetDevice(ctx->device));
  u64 nrows = 0, ncols = 0;
  modle_b200_band_shape(params, interval->end - interval->start, &nrows, &ncols);
  const size_t npx = nrows * ncols + 1;
  const size_t n = interval->num_lefs;
  cudaStream_t s = ctx->stream;
  CUDA_TRY(ctx->d_tasks.reserve(sizeof(modle_b200_cell_task)));
  CUDA_TRY(ctx->d_band.reserve(sizeof(u32) * npx));
  CUDA_TRY(ctx->d_occ1d.reserve(sizeof(u64) * std::max<u64>(ncols, 1)));
  CUDA_TRY(ctx->d_stats.reserve(sizeof(modle_b200_cell_stats)));
  CUDA_TRY(ctx->d_missed.reserve(sizeof(u64)));
  CUDA_TRY(ctx->d_snap_u64.reserve(sizeof(u64) * (5 * n + 2)));
  CUDA_TRY(ctx->d_snap_bar.reserve(num_barriers + 1));
  modle_b200_cell_task t0 = *task;
  CUDA_TRY(cudaMemcpyAsync(ctx->d_tasks.p, &t0, sizeof(t0), cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemsetAsync(ctx->d_band.p, 0, sizeof(u32) * npx, s));
  CUDA_TRY(cudaMemsetAsync(ctx->d_occ1d.p, 0, sizeof(u64) * std::max<u64>(ncols, 1), s));
  CUDA_TRY(cudaMemsetAsync(ctx->d_stats.p, 0, sizeof(modle_b200_cell_stats), s));
  CUDA_TRY(cudaMemsetAsync(ctx->d_missed.p, 0, sizeof(u64), s));
  CUDA_TRY(cudaMemsetAsync(ctx->d_snap_u64.p, 0xFF, sizeof(u64) * (5 * n + 2), s));
  CUDA_TRY(cudaStreamSynchronize(s));
  const int rc = launch_simulate(
      ctx, params, interval, barriers, num_barriers,
      static_cast<const modle_b200_cell_task*>(ctx->d_tasks.p), 1, static_cast<u32*>(ctx->d_band.p),
      static_cast<u64*>(ctx->d_occ1d.p), static_cast<modle_b200_cell_stats*>(ctx->d_stats.p),
      static_cast<u64*>(ctx->d_missed.p), s, static_cast<u64*>(ctx->d_snap_u64.p),
      static_cast<u8*>(ctx->d_snap_bar.p));
  if (rc != MODLE_B200_OK) return rc;
  std::vector<u64> h(5 * n + 2);
  modle_b200_cell_stats st{};
  CUDA_TRY(cudaMemcpyAsync(h.data(), ctx->d_snap_u64.p, sizeof(u64) * h.size(),
                           cudaMemcpyDeviceToHost, s));
  if (num_barriers)
    CUDA_TRY(cudaMemcpyAsync(snapshot->barrier_active, ctx->d_snap_bar.p, num_barriers,
                             cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(&st, ctx->d_stats.p, sizeof(st), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  for (size_t i = 0; i < n; ++i) {
    snapshot->rev_pos[i] = h[i];
    snapshot->fwd_pos[i] = h[n + i];
    snapshot->binding_epoch[i] = h[2 * n + i];
    snapshot->rev_ranks[i] = h[3 * n + i];
    snapshot->fwd_ranks[i] = h[4 * n + i];
  }
  snapshot->num_active_lefs = h[5 * n];
  snapshot->burnin_completed = h[5 * n + 1];
  if (stats_out) *stats_out = st;
  if (st.device_fault)
    return fail(MODLE_B200_ERR_DEVICE_FAULT,
                "kernel reported fault code " + std::to_string(st.device_fault));
  return MODLE_B200_OK;
}

int modle_b200_register_contacts_device(modle_b200_context* ctx, const uint32_t* d_bin1,
                                        const uint32_t* d_bin2, size_t n, uint64_t nrows,
                                        uint64_t ncols, uint32_t* d_band,
                                        uint64_t* d_missed_updates, void* cuda_stream) {
  if (!ctx || !d_bin1 || !d_bin2 || !d_band || !d_missed_updates)
    return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "NULL argument");
  if (nrows == 0 || nrows * ncols + 1 >= (u64(1) << 32))
    return fail(MODLE_B200_ERR_UNSUPPORTED, "band matrix too large for 32-bit pixel index");
  CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
  if (n == 0) return MODLE_B200_OK;
  const int threads = 256;
  const u64 npx = nrows * ncols + 1;
  // binned when the band is larger than the L2 and dense enough for tiles to be reused
  // (MODLE_B200_REGISTER_PATH=direct|binned overrides, for measurements)
  bool use_binned = npx * 4 > ctx->l2_bytes && n >= npx / 4 && n >= (size_t(1) << 20);
  if (const char* force = std::getenv("MODLE_B200_REGISTER_PATH")) {
    if (std::strcmp(force, "direct") == 0) use_binned = false;
    if (std::strcmp(force, "binned") == 0) use_binned = n >= 1;
  }
  if (use_binned && n < (size_t(1) << 32)) {
    // binned path (see k_bin_scatter). Its scratch is shared by the calls on this context.
    CUDA_TRY(cudaEventSynchronize(ctx->binned_done));
    CUDA_TRY(ctx->d_binned.reserve(sizeof(u32) * (n + 4)));
    CUDA_TRY(ctx->d_tiles.reserve(sizeof(u32) * (2 * kMaxTiles + 2)));
    u32* counts = static_cast<u32*>(ctx->d_tiles.p);
    u32* cursor = counts + kMaxTiles;
    u32* d_binned = static_cast<u32*>(ctx->d_binned.p);
    const u32 grid = static_cast<u32>(ctx->num_sms) * 8;
    u32* arrivals = cursor + kMaxTiles + 1;
    CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(u32) * (2 * kMaxTiles + 2), stream));
    // tile = 2^tile_shift pixels; the tiles being replayed at any moment must fit the L2
    u32 tile_shift = kMinTileShift;
    if (const char* ts = std::getenv("MODLE_B200_TILE_SHIFT")) tile_shift = static_cast<u32>(std::atoi(ts));
    tile_shift = std::min(26u, std::max(kMinTileShift, tile_shift));
    while ((npx >> tile_shift) >= kMaxTiles) ++tile_shift;
    k_bin_count<<<grid, threads, 0, stream>>>(d_bin1, d_bin2, n, static_cast<u32>(nrows),
                                              static_cast<u32>(ncols), tile_shift, counts,
                                              d_missed_updates);
    k_bin_offsets<<<1, kBinThreads, 0, stream>>>(counts, cursor);
    k_bin_scatter<<<static_cast<u32>(ctx->num_sms) * 2, kBinThreads, 0, stream>>>(
        d_bin1, d_bin2, n, static_cast<u32>(nrows), static_cast<u32>(ncols), tile_shift, cursor,
        d_binned);
    {
      // after k_bin_scatter cursor[t] is the END of tile t's run
      int per_sm = 0;
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_scatter_binned, threads, 0));
      const u32 coop_grid = static_cast<u32>(ctx->num_sms) * static_cast<u32>(std::min(per_sm, 8));
      const u32* c_binned = d_binned;
      const u32* c_counts = counts;
      const u32* c_end = cursor;
      u32 ntiles = static_cast<u32>(((npx - 1) >> tile_shift) + 1);
      void* args[] = {&c_binned, &c_counts, &c_end, &ntiles, &d_band, &arrivals};
      CUDA_TRY(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_scatter_binned),
                                           dim3(coop_grid), dim3(threads), args, 0, stream));
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(ctx->binned_done, stream));
    ctx->launches += 4;
    return MODLE_B200_OK;
  }
  const int vec_ok = (reinterpret_cast<uintptr_t>(d_bin1) % 16 == 0) &&
                     (reinterpret_cast<uintptr_t>(d_bin2) % 16 == 0);
  const size_t items = vec_ok ? (n + 3) / 4 : n;
  const size_t want = (items + threads - 1) / threads;
  const u32 grid = static_cast<u32>(std::min<size_t>(want, size_t(ctx->num_sms) * 8));
  k_register_contacts<<<grid, threads, 0, stream>>>(d_bin1, d_bin2, n, static_cast<u32>(nrows),
                                                    static_cast<u32>(ncols), d_band,
                                                    d_missed_updates, vec_ok);
  CUDA_TRY(cudaGetLastError());
  ++ctx->launches;
  return MODLE_B200_OK;
}

int modle_b200_calibrate_red_device(modle_b200_context* ctx, uint32_t* d_band, uint64_t num_words,
                                    uint64_t num_reductions, uint64_t seed, void* cuda_stream) {
  if (!ctx || !d_band || num_words == 0)
    return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "NULL argument");
  CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
  if (num_reductions == 0) return MODLE_B200_OK;
  k_calibrate_red<<<static_cast<u32>(ctx->num_sms) * 8, 256, 0, stream>>>(d_band, num_words,
                                                                        num_reductions, seed);
  CUDA_TRY(cudaGetLastError());
  ++ctx->launches;
  return MODLE_B200_OK;
}

}  // extern "C"
