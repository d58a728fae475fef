// A C++ host driving SEVERAL GPUs through include/modle_b200.h, no Python anywhere: one thread,
// one context per GPU, the library's own shard planner, the device-resident simulate call on
// every GPU, then the one collective of the path (modle_b200_reduce_band -> ncclReduce) inside an
// NCCL group, and the root's result compared bit for bit with the unsharded single-GPU run.
// What the reference does with worker threads adding into one shared matrix
// (src/libmodle/cpu/scheduler_simulate.cpp:104-160,190-271;
// src/contact_matrix/contact_matrix_dense_safe_impl.hpp:54-68), over NVLink.
//
//   consumer_multi <ngpus> [cells]     prints {"ok": true, ...}; exit code 0 on equality
//
// TEST INFRASTRUCTURE (tests/test_gpu_multi.py builds and runs it when >= 2 GPUs are visible).
#include <cuda_runtime.h>
#include <nccl.h>

#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "modle_b200.h"

#define CHECK(call)                                                                     \
  do {                                                                                  \
    const int rc_ = (call);                                                             \
    if (rc_ != MODLE_B200_OK) {                                                         \
      std::fprintf(stderr, "%s failed: %d: %s\n", #call, rc_, modle_b200_last_error()); \
      return 2;                                                                         \
    }                                                                                   \
  } while (0)
#define CUDA(call)                                                                        \
  do {                                                                                    \
    const cudaError_t e_ = (call);                                                        \
    if (e_ != cudaSuccess) {                                                              \
      std::fprintf(stderr, "%s failed: %s\n", #call, cudaGetErrorString(e_));             \
      return 2;                                                                           \
    }                                                                                     \
  } while (0)
#define NCCL(call)                                                                        \
  do {                                                                                    \
    const ncclResult_t r_ = (call);                                                       \
    if (r_ != ncclSuccess) {                                                              \
      std::fprintf(stderr, "%s failed: %s\n", #call, ncclGetErrorString(r_));             \
      return 2;                                                                           \
    }                                                                                     \
  } while (0)

int main(int argc, char** argv) {
  const int world = argc > 1 ? std::atoi(argv[1]) : 2;
  const std::uint64_t cells = argc > 2 ? std::strtoull(argv[2], nullptr, 10) : 16;
  int ndev = 0;
  CUDA(cudaGetDeviceCount(&ndev));
  if (world < 2 || ndev < world) {
    std::fprintf(stderr, "needs %d GPUs, %d visible\n", world, ndev);
    return 77;
  }

  modle_b200_sim_params p;
  modle_b200_default_params(&p);
  p.num_cells = cells;
  p.seed = 7;
  p.target_contact_density = 0.01;
  CHECK(modle_b200_transform_params(&p, 0, 0, 0));
  const char* chrom = "chrMulti";
  modle_b200_interval iv;
  iv.chrom_size = iv.end = 40000000;
  iv.start = 0;
  iv.num_lefs = modle_b200_compute_num_lefs(&p, iv.end - iv.start);
  std::vector<modle_b200_barrier> bars;
  for (std::uint64_t pos = 31013, k = 0; pos < iv.end; pos += 57119, ++k) {
    modle_b200_barrier b;
    std::memset(&b, 0, sizeof(b));
    b.pos = pos;
    b.stp_inactive = p.barrier_not_occupied_stp;
    b.stp_active = modle_b200_stp_active_from_occupancy(b.stp_inactive, 0.65 + 0.01 * double(k % 30));
    b.blocking_direction = (k % 3 == 0) ? MODLE_B200_DIR_FWD : MODLE_B200_DIR_REV;
    bars.push_back(b);
  }
  std::vector<modle_b200_cell_task> tasks(cells);
  CHECK(modle_b200_make_cell_tasks(&p, chrom, std::strlen(chrom), &iv, tasks.data()));
  std::uint64_t nrows = 0, ncols = 0;
  modle_b200_band_shape(&p, iv.end - iv.start, &nrows, &ncols);
  const std::size_t npx = nrows * ncols + 1;

  // the plan: one interval, `cells` cells, `world` ranks -> the planner splits it by cells
  const double w = modle_b200_cell_weight(iv.num_lefs, bars.size());
  std::size_t nsh = 0;
  CHECK(modle_b200_plan_shards(&w, 1, cells, world, 0, 0.0, nullptr, 0, &nsh));
  std::vector<modle_b200_shard> shards(nsh);
  CHECK(modle_b200_plan_shards(&w, 1, cells, world, 0, 0.0, shards.data(), shards.size(), &nsh));
  const int root = shards[0].rank;
  std::uint64_t covered = 0;
  for (const auto& s : shards) covered += s.cell_hi - s.cell_lo;
  if (covered != cells || nsh < static_cast<std::size_t>(world)) {
    std::fprintf(stderr, "unexpected plan: %zu shards covering %" PRIu64 " cells\n", nsh, covered);
    return 3;
  }

  std::vector<int> devs(world);
  for (int r = 0; r < world; ++r) devs[r] = r;
  std::vector<ncclComm_t> comms(world);
  NCCL(ncclCommInitAll(comms.data(), world, devs.data()));

  std::vector<modle_b200_context*> ctx(world, nullptr);
  std::vector<cudaStream_t> streams(world);
  std::vector<std::uint32_t*> d_band(world, nullptr);
  std::vector<std::uint64_t*> d_occ(world, nullptr), d_missed(world, nullptr);
  std::vector<modle_b200_cell_stats*> d_stats(world, nullptr);
  std::vector<modle_b200_cell_task*> d_tasks(world, nullptr);
  for (int r = 0; r < world; ++r) {
    CUDA(cudaSetDevice(r));
    CHECK(modle_b200_init(&ctx[r], r));
    CUDA(cudaStreamCreateWithFlags(&streams[r], cudaStreamNonBlocking));
    CUDA(cudaMalloc(&d_band[r], npx * 4));
    CUDA(cudaMalloc(&d_occ[r], ncols * 8));
    CUDA(cudaMalloc(&d_missed[r], 8));
    CUDA(cudaMalloc(&d_stats[r], cells * sizeof(modle_b200_cell_stats)));
    CUDA(cudaMalloc(&d_tasks[r], cells * sizeof(modle_b200_cell_task)));
    CUDA(cudaMemsetAsync(d_band[r], 0, npx * 4, streams[r]));
    CUDA(cudaMemsetAsync(d_occ[r], 0, ncols * 8, streams[r]));
    CUDA(cudaMemsetAsync(d_missed[r], 0, 8, streams[r]));
    CUDA(cudaMemsetAsync(d_stats[r], 0, cells * sizeof(modle_b200_cell_stats), streams[r]));
    CUDA(cudaMemcpyAsync(d_tasks[r], tasks.data(), cells * sizeof(modle_b200_cell_task),
                         cudaMemcpyHostToDevice, streams[r]));
  }
  // every rank simulates its shards (asynchronous: all GPUs run at once)
  for (const auto& s : shards) {
    const int r = s.rank;
    CUDA(cudaSetDevice(r));
    CHECK(modle_b200_simulate_interval_device(ctx[r], &p, &iv, bars.data(), bars.size(),
                                              d_tasks[r] + s.cell_lo, s.cell_hi - s.cell_lo,
                                              d_band[r], d_occ[r], d_stats[r] + s.cell_lo,
                                              d_missed[r], streams[r]));
  }
  // the one exchange step
  NCCL(ncclGroupStart());
  for (int r = 0; r < world; ++r)
    CHECK(modle_b200_reduce_band(ctx[r], comms[r], d_band[r], nrows, ncols, d_occ[r], d_missed[r],
                                 root, streams[r]));
  NCCL(ncclGroupEnd());
  std::vector<std::uint32_t> band(npx);
  std::vector<std::uint64_t> occ(ncols);
  std::uint64_t missed = 0;
  CUDA(cudaSetDevice(root));
  CUDA(cudaMemcpyAsync(band.data(), d_band[root], npx * 4, cudaMemcpyDeviceToHost, streams[root]));
  CUDA(cudaMemcpyAsync(occ.data(), d_occ[root], ncols * 8, cudaMemcpyDeviceToHost, streams[root]));
  CUDA(cudaMemcpyAsync(&missed, d_missed[root], 8, cudaMemcpyDeviceToHost, streams[root]));
  std::uint64_t faults = 0, contacts = 0;
  for (int r = 0; r < world; ++r) {
    CUDA(cudaSetDevice(r));
    CUDA(cudaStreamSynchronize(streams[r]));
    std::vector<modle_b200_cell_stats> st(cells);
    CUDA(cudaMemcpy(st.data(), d_stats[r], cells * sizeof(modle_b200_cell_stats), cudaMemcpyDeviceToHost));
    for (const auto& x : st) {
      faults += x.device_fault != 0;
      contacts += x.num_contacts;
    }
  }

  // the unsharded run on GPU 0 through the host-buffer call
  std::vector<std::uint32_t> band1(npx, 0);
  std::vector<std::uint64_t> occ1(ncols, 0);
  std::vector<modle_b200_cell_stats> st1(cells);
  std::uint64_t missed1 = 0;
  CUDA(cudaSetDevice(0));
  CHECK(modle_b200_simulate_interval(ctx[0], &p, &iv, bars.data(), bars.size(), tasks.data(), cells,
                                     band1.data(), occ1.data(), st1.data(), &missed1));
  std::uint64_t contacts1 = 0, band_sum = 0;
  for (const auto& x : st1) contacts1 += x.num_contacts;
  for (std::uint32_t v : band) band_sum += v;
  const bool same = band == band1 && occ == occ1 && missed == missed1 && contacts == contacts1 &&
                    faults == 0 && band_sum + missed == contacts;
  for (int r = 0; r < world; ++r) {
    cudaSetDevice(r);
    cudaFree(d_band[r]);
    cudaFree(d_occ[r]);
    cudaFree(d_missed[r]);
    cudaFree(d_stats[r]);
    cudaFree(d_tasks[r]);
    modle_b200_destroy(ctx[r]);
    ncclCommDestroy(comms[r]);
  }
  std::printf("{\"ok\": %s, \"world\": %d, \"shards\": %zu, \"root\": %d, \"cells\": %" PRIu64
              ", \"contacts\": %" PRIu64 ", \"band_sum\": %" PRIu64 ", \"missed\": %" PRIu64
              ", \"reduced_bytes_per_rank\": %zu}\n",
              same ? "true" : "false", world, nsh, root, cells, contacts, band_sum, missed,
              npx * 4 + ncols * 8 + 8);
  return same ? 0 : 1;
}
