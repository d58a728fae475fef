// Band -> sorted COO pixels on the device: the hand-off to the .cool writer.
//
// Replaces the pixel loop of modle::io::internal::append_contact_matrix_to_cooler
// (src/libmodle_io/contact_matrix_dense_io_impl.hpp:50-71):
//
//   for i in [0, ncols): for j in [i, min(ncols, i + nrows)):
//     if (n = matrix.unsafe_get(i, j)) != 0: emit {bin_offset + i, bin_offset + j, int32(n)}
//
// with unsafe_get(i, j) = band[j * nrows + (j - i)] (contact_matrix_dense_unsafe_impl.hpp:33-42,
// contact_matrix_internal_impl.hpp:19-46). The output is the same sequence, i.e. sorted by
// (bin1, bin2), as an array of hictk::ThinPixel<std::int32_t>-compatible records.
//
// Pixel row r reads band[c * (nrows + 1) - r] for c = r, r+1, ...: a stride of nrows + 1 words
// along the row, but CONSECUTIVE words across rows for a fixed column c. So a CTA owns 32 pixel
// rows (one per lane while loading: every warp load is one 128-byte run), stages 32 x 256 tiles in
// shared memory and compacts each row from there (ballot + popc, coalesced 24-byte records).
//
//   k_count_row_pixels   non-zeros per pixel row + their exclusive scan (decoupled look-back
//                        over the CTAs) -> row offsets      reads 4 B/pixel, writes 8 B/row
//   k_fill_pixels        writes the pixels                  reads 4 B/pixel, writes 24 B/non-zero
#include <cuda_runtime.h>

#include <algorithm>
#include <string>

#include "../../include/modle_b200.h"
#include "sim_types.hpp"
#include "status.hpp"

using namespace modle_b200;

#include "context.hpp"

namespace modle_b200 {

constexpr int kPxRows = 32;      // pixel rows per CTA
constexpr int kPxCols = 256;     // band columns staged per round
constexpr int kPxThreads = 256;  // 8 warps
constexpr int kPxWarps = kPxThreads / 32;

__device__ __forceinline__ u32 load_pixel(const u32* __restrict__ band, u32 nrows, u32 ncols,
                                          u32 r, u32 c) {
  // pixel (r, c) is stored iff r <= c < min(ncols, r + nrows)
  const bool valid = r < ncols && c < ncols && c >= r && c - r < nrows;
  return valid ? __ldg(band + (u64(c) * (u64(nrows) + 1) - r)) : 0u;
}

// Counts the non-zero pixels of 32 rows per CTA and turns the counts into row offsets in the same
// pass: CTAs take their row block from a ticket (so a CTA only ever waits for CTAs that already
// run) and chain their totals with a decoupled look-back -- flags[b] = status << 62 | value,
// status 1 = this block's total, 2 = inclusive prefix up to and including this block.
// `sync` = {ticket, flags[nblocks]} as 64-bit words, zeroed before the launch.
__global__ void __launch_bounds__(kPxThreads)
    k_count_row_pixels(const u32* __restrict__ band, u32 nrows, u32 ncols,
                       u64* __restrict__ row_offsets, u64* __restrict__ sync) {
  __shared__ u32 part[kPxWarps][kPxRows];
  __shared__ u32 s_block;
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0)
    s_block = static_cast<u32>(atomicAdd(reinterpret_cast<unsigned long long*>(sync), 1ull));
  __syncthreads();
  const u32 block = s_block;
  u64* flags = sync + 1;
  const u32 r0 = block * kPxRows;
  const u32 r = r0 + lane;
  const u64 cend = min(u64(ncols), u64(r0) + kPxRows - 1 + nrows);
  u32 cnt = 0;
  u64 c = u64(r0) + warp;
  for (; c + 3 * kPxWarps < cend; c += 4 * kPxWarps) {  // four independent 128-byte loads in flight
    const u32 v0 = load_pixel(band, nrows, ncols, r, u32(c));
    const u32 v1 = load_pixel(band, nrows, ncols, r, u32(c + kPxWarps));
    const u32 v2 = load_pixel(band, nrows, ncols, r, u32(c + 2 * kPxWarps));
    const u32 v3 = load_pixel(band, nrows, ncols, r, u32(c + 3 * kPxWarps));
    cnt += (v0 != 0) + (v1 != 0) + (v2 != 0) + (v3 != 0);
  }
  for (; c < cend; c += kPxWarps) cnt += load_pixel(band, nrows, ncols, r, u32(c)) != 0;
  part[warp][lane] = cnt;
  __syncthreads();
  if (warp != 0) return;
  u32 mine = 0;
#pragma unroll
  for (int w = 0; w < kPxWarps; ++w) mine += part[w][lane];
  u32 inc = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const u32 o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= u32(d)) inc += o;
  }
  const u64 total = __shfl_sync(0xffffffffu, inc, 31);
  u64 prefix = 0;
  if (lane == 0) {
    constexpr u64 kValue = (u64(1) << 62) - 1;
    if (block != 0) {
      asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(flags + block),
                   "l"((u64(1) << 62) | total)
                   : "memory");
      for (u32 i = block; i-- > 0;) {
        u64 f;
        do {
          asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(f) : "l"(flags + i) : "memory");
        } while ((f >> 62) == 0);
        prefix += f & kValue;
        if ((f >> 62) == 2) break;
      }
    }
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(flags + block),
                 "l"((u64(2) << 62) | (prefix + total))
                 : "memory");
  }
  prefix = __shfl_sync(0xffffffffu, prefix, 0);
  if (r < ncols) row_offsets[r] = prefix + inc - mine;
  if (lane == 31 && r0 + kPxRows >= ncols) row_offsets[ncols] = prefix + total;
}

__global__ void __launch_bounds__(kPxThreads)
    k_fill_pixels(const u32* __restrict__ band, u32 nrows, u32 ncols, u64 bin_offset,
                  const u64* __restrict__ row_offsets, modle_b200_pixel* __restrict__ pixels,
                  u64 capacity) {
  __shared__ u32 tile[kPxRows][kPxCols + 1];  // +1: lanes of a load write a conflict-free diagonal
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 r0 = blockIdx.x * kPxRows;
  const u64 cend = min(u64(ncols), u64(r0) + kPxRows - 1 + nrows);
  constexpr int kRowsPerWarp = kPxRows / kPxWarps;
  u64 cursor[kRowsPerWarp];  // next output slot of the rows this warp compacts (warp-uniform)
#pragma unroll
  for (int k = 0; k < kRowsPerWarp; ++k) {
    const u32 r = r0 + warp + k * kPxWarps;
    cursor[k] = r < ncols ? row_offsets[r] : 0;
  }
  for (u64 c0 = r0; c0 < cend; c0 += kPxCols) {
    // stage: lane = pixel row, each warp takes every 8th column of the round
#pragma unroll 8
    for (int k = 0; k < kPxCols / kPxWarps; ++k) {
      const u32 x = warp + k * kPxWarps;
      tile[lane][x] = load_pixel(band, nrows, ncols, r0 + lane, u32(c0 + x));
    }
    __syncthreads();
    // compact: lane = column within a 32-wide slice of the staged row
#pragma unroll
    for (int k = 0; k < kRowsPerWarp; ++k) {
      const u32 rr = warp + k * kPxWarps;
      const u32 r = r0 + rr;
      if (r >= ncols) continue;
      u64 base = cursor[k];
#pragma unroll
      for (int x0 = 0; x0 < kPxCols; x0 += 32) {
        const u32 v = tile[rr][x0 + lane];
        const u32 m = __ballot_sync(0xffffffffu, v != 0);
        if (v != 0) {
          const u64 idx = base + __popc(m & ((1u << lane) - 1u));
          if (idx < capacity) {
            u64* out = reinterpret_cast<u64*>(pixels + idx);
            out[0] = bin_offset + r;
            out[1] = bin_offset + c0 + x0 + lane;
            out[2] = u64(v);  // int32 count (two's complement of the uint32) + zeroed padding
          }
        }
        base += __popc(m);
      }
      cursor[k] = base;
    }
    __syncthreads();
  }
}

// ---- 1D LEF occupancy profile -----------------------------------------------------------------
// write_lef_occupancy_to_bwig (src/libmodle/cpu/simulation.cpp:170-197) before the bigWig call:
// profile[i] = float(double(occ[i]) / double(max(occ))). An all-zero track gives 0/0 = NaN, as in
// the reference.
__global__ void __launch_bounds__(256) k_occ_max(const u64* __restrict__ occ, size_t n,
                                                 u64* __restrict__ out_max) {
  u64 m = 0;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += size_t(gridDim.x) * blockDim.x)
    m = max(m, occ[i]);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(reinterpret_cast<unsigned long long*>(out_max), m);
}

__global__ void __launch_bounds__(256) k_occ_profile(const u64* __restrict__ occ, size_t n,
                                                     const u64* __restrict__ max_v,
                                                     float* __restrict__ out) {
  const double denom = static_cast<double>(*max_v);
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += size_t(gridDim.x) * blockDim.x)
    out[i] = static_cast<float>(static_cast<double>(occ[i]) / denom);
}

}  // namespace modle_b200

namespace {

int check_band_shape(uint64_t nrows, uint64_t ncols) {
  if (nrows == 0 || ncols == 0)
    return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "band matrix with zero rows or columns");
  if (nrows * ncols + 1 >= (u64(1) << 32))
    return fail(MODLE_B200_ERR_UNSUPPORTED, "band matrix too large for 32-bit pixel index");
  return MODLE_B200_OK;
}

}  // namespace

extern "C" {

int modle_b200_count_pixels_device(modle_b200_context* ctx, const uint32_t* d_band, uint64_t nrows,
                                   uint64_t ncols, uint64_t* d_row_offsets, void* cuda_stream) {
  if (!ctx || !d_band || !d_row_offsets)
    return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "NULL argument");
  if (const int rc = check_band_shape(nrows, ncols)) return rc;
  CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
  const u32 grid = static_cast<u32>((ncols + kPxRows - 1) / kPxRows);
  // look-back scratch (ticket + one flag per CTA); a few of them so that calls issued on
  // different streams of one context do not share one
  DevBuf& sync = ctx->d_px_sync[ctx->px_calls++ % kPxSyncSlots];
  const size_t sync_bytes = sizeof(u64) * (size_t(grid) + 1);
  if (sync.cap < sync_bytes) {
    CUDA_TRY(cudaStreamSynchronize(stream));  // the previous user of this slot may still run
    CUDA_TRY(sync.reserve(std::max(sync_bytes, size_t(1) << 16)));
  }
  CUDA_TRY(cudaMemsetAsync(sync.p, 0, sync_bytes, stream));
  k_count_row_pixels<<<grid, kPxThreads, 0, stream>>>(d_band, static_cast<u32>(nrows),
                                                      static_cast<u32>(ncols), d_row_offsets,
                                                      static_cast<u64*>(sync.p));
  CUDA_TRY(cudaGetLastError());
  ++ctx->launches;
  return MODLE_B200_OK;
}

int modle_b200_fill_pixels_device(modle_b200_context* ctx, const uint32_t* d_band, uint64_t nrows,
                                  uint64_t ncols, uint64_t bin_offset,
                                  const uint64_t* d_row_offsets, modle_b200_pixel* d_pixels,
                                  uint64_t capacity, void* cuda_stream) {
  if (!ctx || !d_band || !d_row_offsets || (!d_pixels && capacity != 0))
    return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "NULL argument");
  if (const int rc = check_band_shape(nrows, ncols)) return rc;
  CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
  if (capacity == 0) return MODLE_B200_OK;
  const u32 grid = static_cast<u32>((ncols + kPxRows - 1) / kPxRows);
  k_fill_pixels<<<grid, kPxThreads, 0, stream>>>(d_band, static_cast<u32>(nrows),
                                                 static_cast<u32>(ncols), bin_offset,
                                                 d_row_offsets, d_pixels, capacity);
  CUDA_TRY(cudaGetLastError());
  ++ctx->launches;
  return MODLE_B200_OK;
}

int modle_b200_band_to_pixels(modle_b200_context* ctx, const uint32_t* band, uint64_t nrows,
                              uint64_t ncols, uint64_t bin_offset, modle_b200_pixel* pixels_out,
                              uint64_t capacity, uint64_t* num_pixels_out) {
  if (!ctx || !band || !num_pixels_out)
    return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "NULL argument");
  if (!pixels_out && capacity != 0)
    return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "pixels_out is NULL but capacity is not 0");
  *num_pixels_out = 0;
  if (const int rc = check_band_shape(nrows, ncols)) return rc;
  CUDA_TRY(cudaSetDevice(ctx->device));
  const size_t band_bytes = sizeof(u32) * (nrows * ncols + 1);
  CUDA_TRY(ctx->d_px_band.reserve(band_bytes));
  CUDA_TRY(ctx->d_px_rows.reserve(sizeof(u64) * (ncols + 1)));
  const u32* d_band = static_cast<const u32*>(ctx->d_px_band.p);
  u64* d_rows = static_cast<u64*>(ctx->d_px_rows.p);
  CUDA_TRY(cudaMemcpyAsync(ctx->d_px_band.p, band, band_bytes, cudaMemcpyHostToDevice,
                           ctx->stream));
  if (const int rc = modle_b200_count_pixels_device(ctx, d_band, nrows, ncols, d_rows, ctx->stream))
    return rc;
  u64 nnz = 0;
  CUDA_TRY(cudaMemcpyAsync(&nnz, d_rows + ncols, sizeof(u64), cudaMemcpyDeviceToHost,
                           ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  *num_pixels_out = nnz;
  if (capacity == 0 && !pixels_out) return MODLE_B200_OK;  // size query
  if (nnz > capacity)
    return fail(MODLE_B200_ERR_INVALID_ARGUMENT,
                "pixels_out holds " + std::to_string(capacity) + " pixels, the band has " +
                    std::to_string(nnz));
  if (nnz == 0) return MODLE_B200_OK;
  CUDA_TRY(ctx->d_px_out.reserve(sizeof(modle_b200_pixel) * nnz));
  auto* d_px = static_cast<modle_b200_pixel*>(ctx->d_px_out.p);
  if (const int rc = modle_b200_fill_pixels_device(ctx, d_band, nrows, ncols, bin_offset, d_rows,
                                                   d_px, nnz, ctx->stream))
    return rc;
  CUDA_TRY(cudaMemcpyAsync(pixels_out, d_px, sizeof(modle_b200_pixel) * nnz,
                           cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return MODLE_B200_OK;
}

int modle_b200_lef_occupancy_profile_device(modle_b200_context* ctx, const uint64_t* d_occ1d,
                                            size_t n, float* d_profile, uint64_t* d_scratch_max,
                                            void* cuda_stream) {
  if (!ctx || !d_occ1d || !d_profile || !d_scratch_max)
    return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "NULL argument");
  if (n == 0) return MODLE_B200_OK;
  CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
  const u32 grid = static_cast<u32>(std::min<size_t>((n + 255) / 256, size_t(ctx->num_sms) * 8));
  CUDA_TRY(cudaMemsetAsync(d_scratch_max, 0, sizeof(u64), stream));
  k_occ_max<<<grid, 256, 0, stream>>>(d_occ1d, n, d_scratch_max);
  k_occ_profile<<<grid, 256, 0, stream>>>(d_occ1d, n, d_scratch_max, d_profile);
  CUDA_TRY(cudaGetLastError());
  ctx->launches += 2;
  return MODLE_B200_OK;
}

int modle_b200_lef_occupancy_profile(modle_b200_context* ctx, const uint64_t* occ1d, size_t n,
                                     float* profile_out) {
  if (!ctx || !occ1d || !profile_out) return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "NULL argument");
  if (n == 0) return MODLE_B200_OK;
  CUDA_TRY(cudaSetDevice(ctx->device));
  CUDA_TRY(ctx->d_px_rows.reserve(sizeof(u64) * (n + 1)));
  CUDA_TRY(ctx->d_px_out.reserve(sizeof(float) * n));
  u64* d_occ = static_cast<u64*>(ctx->d_px_rows.p);
  float* d_out = static_cast<float*>(ctx->d_px_out.p);
  CUDA_TRY(cudaMemcpyAsync(d_occ, occ1d, sizeof(u64) * n, cudaMemcpyHostToDevice, ctx->stream));
  if (const int rc = modle_b200_lef_occupancy_profile_device(ctx, d_occ, n, d_out, d_occ + n,
                                                             ctx->stream))
    return rc;
  CUDA_TRY(cudaMemcpyAsync(profile_out, d_out, sizeof(float) * n, cudaMemcpyDeviceToHost,
                           ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return MODLE_B200_OK;
}

}  // extern "C"
