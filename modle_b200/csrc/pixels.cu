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
//   k_count_row_pixels   non-zeros per pixel row            reads 4 B/pixel
//   k_scan_row_counts    exclusive scan -> row offsets      (ncols + 1 words)
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

__global__ void __launch_bounds__(kPxThreads)
    k_count_row_pixels(const u32* __restrict__ band, u32 nrows, u32 ncols,
                       u64* __restrict__ row_counts) {
  __shared__ u32 part[kPxWarps][kPxRows];
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 r0 = blockIdx.x * kPxRows;
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
  if (warp == 0 && r < ncols) {
    u32 s = 0;
#pragma unroll
    for (int w = 0; w < kPxWarps; ++w) s += part[w][lane];
    row_counts[r] = s;
  }
}

// In-place exclusive scan of counts[0 .. n) with the total left in counts[n] (one CTA; n is the
// number of pixel rows, at most a few hundred thousand).
__global__ void __launch_bounds__(1024) k_scan_row_counts(u64* __restrict__ counts, u32 n) {
  __shared__ u64 warp_sums[32];
  __shared__ u64 carry;
  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 chunk = (n + blockDim.x - 1) / blockDim.x;
  const u32 lo = min(n, tid * chunk), hi = min(n, lo + chunk);
  u64 s = 0;
  for (u32 i = lo; i < hi; ++i) s += counts[i];
  u64 incl = s;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const u64 t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= u32(d)) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const u64 w = warp_sums[lane];
    u64 wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const u64 t = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= u32(d)) wi += t;
    }
    warp_sums[lane] = wi - w;
    if (lane == 31) carry = wi;
  }
  __syncthreads();
  u64 run = warp_sums[warp] + incl - s;
  for (u32 i = lo; i < hi; ++i) {
    const u64 v = counts[i];
    counts[i] = run;
    run += v;
  }
  if (tid == 0) counts[n] = carry;
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
  k_count_row_pixels<<<grid, kPxThreads, 0, stream>>>(d_band, static_cast<u32>(nrows),
                                                      static_cast<u32>(ncols), d_row_offsets);
  k_scan_row_counts<<<1, 1024, 0, stream>>>(d_row_offsets, static_cast<u32>(ncols));
  CUDA_TRY(cudaGetLastError());
  ctx->launches += 2;
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

}  // extern "C"
