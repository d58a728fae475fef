// Multi-GPU half of the C ABI for a C++ host (include/modle_b200.h, "several GPUs"): the shard
// planner and the one collective of the path.
//
// The reference has one level of parallelism -- independent (interval, cell) tasks popped by
// worker threads that all add into the interval's shared contact matrix
// (src/libmodle/cpu/scheduler_simulate.cpp:104-160 produce, :190-271 consume;
// ContactMatrixDense::increment, src/contact_matrix/contact_matrix_dense_safe_impl.hpp:54-68).
// Over several GPUs the same tasks are dealt out as (interval, cell range) shards, one process or
// thread per GPU, and an interval whose cells ended up on several GPUs has its band summed onto one
// of them: ncclReduce(uint32, sum) -- integer sums commute, so the result is the unsharded one.
#include <dlfcn.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "context.hpp"

namespace {

struct Piece {
  uint64_t interval, lo, hi;
  int rank;
  double weight;
};

// longest-processing-time-first; returns the per-rank loads
std::vector<double> assign(std::vector<Piece>& pieces, int world) {
  std::vector<size_t> order(pieces.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) {
    const Piece &x = pieces[a], &y = pieces[b];
    if (x.weight != y.weight) return x.weight > y.weight;
    if (x.interval != y.interval) return x.interval < y.interval;
    return x.lo < y.lo;
  });
  std::vector<double> load(static_cast<size_t>(world), 0.0);
  for (size_t i : order) {
    int best = 0;
    for (int r = 1; r < world; ++r)
      if (load[static_cast<size_t>(r)] < load[static_cast<size_t>(best)]) best = r;
    pieces[i].rank = best;
    load[static_cast<size_t>(best)] += pieces[i].weight;
  }
  return load;
}

// NCCL is bound at run time (dlopen): the library has no link-time dependency on it, and a host
// that already loaded an NCCL (its own, or the one a framework bundles) gets that same copy.
struct Nccl {
  using reduce_fn = int (*)(const void*, void*, size_t, int, int, int, void*, cudaStream_t);
  using errstr_fn = const char* (*)(int);
  reduce_fn reduce = nullptr;
  errstr_fn errstr = nullptr;
  std::string error;
};

const Nccl& nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = nullptr;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) {
      n.error = std::string("cannot load NCCL: ") + dlerror();
      return;
    }
    n.reduce = reinterpret_cast<Nccl::reduce_fn>(dlsym(h, "ncclReduce"));
    n.errstr = reinterpret_cast<Nccl::errstr_fn>(dlsym(h, "ncclGetErrorString"));
    if (!n.reduce) n.error = "libnccl has no ncclReduce";
  });
  return n;
}

constexpr int kNcclUint32 = 3, kNcclUint64 = 5, kNcclSum = 0;  // nccl.h: ncclDataType_t / ncclRedOp_t

}  // namespace

extern "C" {

// SM time of one cell of an interval, relative (the planner's weight). One cell-epoch costs a
// configuration-dependent constant plus a per-LEF term, divided by the cells an SM hosts at a
// time (fitted on a B200, see modle_b200/distributed.py); epoch counts are the same for every
// interval to within a few percent, so they drop out.
double modle_b200_cell_weight(uint64_t num_lefs, uint64_t num_barriers) {
  if (num_lefs == 0) return 0.0;
  uint32_t threads = 0, per_sm = 0;
  if (modle_b200_launch_geometry(num_lefs, num_barriers, &threads, &per_sm, nullptr) != MODLE_B200_OK)
    return static_cast<double>(num_lefs);
  const double a = threads == 1024 ? 80e3 : (threads == 512 ? 46e3 : 12e3);
  const double b = threads == 1024 ? 40.0 : (threads == 512 ? 80.0 : 160.0);
  return (a + b * static_cast<double>(num_lefs)) / static_cast<double>(per_sm ? per_sm : 1);
}

int modle_b200_plan_shards(const double* cell_weights, size_t num_intervals, uint64_t num_cells,
                           int world_size, int slice_all, double tolerance,
                           modle_b200_shard* shards_out, size_t capacity, size_t* num_shards_out) {
  if (!cell_weights || !num_shards_out || world_size < 1)
    return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "NULL argument or world_size < 1");
  if (tolerance <= 0.0) tolerance = 1.10;
  if (slice_all < 0) {
    // Automatic: one cell slice of EVERY interval per rank pays when a slice still is a launch of
    // its own right -- at least ~1.5 waves of cells over the 148 SMs -- because every rank then
    // runs the same mix and twice as many launches fill each other's tails (measured on B200s,
    // C2 with 512 cells: 2 ranks 2.13 s per step sliced vs 2.85 s whole; 8 ranks 1.24 s sliced --
    // 64-cell launches are nothing but tail -- vs 0.72 s whole).
    slice_all = (world_size > 1 && num_cells / static_cast<uint64_t>(world_size) >= 222) ? 1 : 0;
  }
  std::vector<Piece> pieces;
  if (slice_all && world_size > 1) {
    for (size_t i = 0; i < num_intervals; ++i) {
      if (!(cell_weights[i] > 0.0) || num_cells == 0) continue;
      for (int k = 0; k < world_size; ++k) {
        const uint64_t lo = num_cells * static_cast<uint64_t>(k) / static_cast<uint64_t>(world_size);
        const uint64_t hi = num_cells * static_cast<uint64_t>(k + 1) / static_cast<uint64_t>(world_size);
        if (hi > lo)
          pieces.push_back(Piece{i, lo, hi, static_cast<int>((i + static_cast<size_t>(k)) %
                                                              static_cast<size_t>(world_size)),
                                 cell_weights[i] * static_cast<double>(hi - lo)});
      }
    }
  } else {
    for (size_t i = 0; i < num_intervals; ++i)
      if (cell_weights[i] > 0.0 && num_cells > 0)
        pieces.push_back(Piece{i, 0, num_cells, -1, cell_weights[i] * static_cast<double>(num_cells)});
    if (!pieces.empty()) {
      const size_t max_pieces = pieces.size() + 8 * static_cast<size_t>(world_size);
      double total = 0.0;
      for (const Piece& p : pieces) total += p.weight;
      for (;;) {
        const std::vector<double> load = assign(pieces, world_size);
        int worst = 0;
        for (int r = 1; r < world_size; ++r)
          if (load[static_cast<size_t>(r)] > load[static_cast<size_t>(worst)]) worst = r;
        if (world_size == 1 ||
            load[static_cast<size_t>(worst)] <= tolerance * total / static_cast<double>(world_size) ||
            pieces.size() >= max_pieces)
          break;
        // halve the heaviest splittable piece of the heaviest rank
        long best = -1;
        for (size_t i = 0; i < pieces.size(); ++i) {
          const Piece& p = pieces[i];
          if (p.rank != worst || p.hi - p.lo < 2) continue;
          if (best < 0) {
            best = static_cast<long>(i);
            continue;
          }
          const Piece& q = pieces[static_cast<size_t>(best)];
          if (p.weight > q.weight ||
              (p.weight == q.weight &&
               (p.interval < q.interval || (p.interval == q.interval && p.lo < q.lo))))
            best = static_cast<long>(i);
        }
        if (best < 0) break;
        Piece& p = pieces[static_cast<size_t>(best)];
        const uint64_t mid = (p.lo + p.hi) / 2;
        const double per_cell = p.weight / static_cast<double>(p.hi - p.lo);
        const Piece q{p.interval, mid, p.hi, -1, per_cell * static_cast<double>(p.hi - mid)};
        p.hi = mid;
        p.weight = per_cell * static_cast<double>(mid - p.lo);
        pieces.push_back(q);
      }
      std::stable_sort(pieces.begin(), pieces.end(), [](const Piece& x, const Piece& y) {
        return x.interval != y.interval ? x.interval < y.interval : x.lo < y.lo;
      });
    }
  }
  *num_shards_out = pieces.size();
  if (!shards_out && capacity == 0) return MODLE_B200_OK;  // size query
  if (capacity < pieces.size() || !shards_out)
    return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "shards_out too small");
  for (size_t i = 0; i < pieces.size(); ++i) {
    shards_out[i].interval = pieces[i].interval;
    shards_out[i].cell_lo = pieces[i].lo;
    shards_out[i].cell_hi = pieces[i].hi;
    shards_out[i].rank = pieces[i].rank;
    shards_out[i].reserved_ = 0;
    shards_out[i].weight = pieces[i].weight;
  }
  return MODLE_B200_OK;
}

int modle_b200_reduce_band(modle_b200_context* ctx, void* nccl_comm, uint32_t* d_band,
                           uint64_t nrows, uint64_t ncols, uint64_t* d_occ1d,
                           uint64_t* d_missed_updates, int root, void* cuda_stream) {
  if (!ctx || !nccl_comm || !d_band)
    return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "NULL argument");
  const Nccl& n = nccl();
  if (!n.reduce) return fail(MODLE_B200_ERR_UNSUPPORTED, n.error);
  CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
  auto check = [&](int rc, const char* what) -> int {
    if (rc == 0) return MODLE_B200_OK;
    return fail(MODLE_B200_ERR_CUDA, std::string("ncclReduce(") + what + "): " +
                                         (n.errstr ? n.errstr(rc) : std::to_string(rc).c_str()));
  };
  // in place: on the root the sums replace its own contribution
  if (const int rc = check(n.reduce(d_band, d_band, nrows * ncols + 1, kNcclUint32, kNcclSum, root,
                                    nccl_comm, stream), "band"))
    return rc;
  if (d_occ1d)
    if (const int rc = check(n.reduce(d_occ1d, d_occ1d, ncols, kNcclUint64, kNcclSum, root,
                                      nccl_comm, stream), "occ1d"))
      return rc;
  if (d_missed_updates)
    if (const int rc = check(n.reduce(d_missed_updates, d_missed_updates, 1, kNcclUint64, kNcclSum,
                                      root, nccl_comm, stream), "missed"))
      return rc;
  return MODLE_B200_OK;
}

}  // extern "C"
