// TEST INFRASTRUCTURE ONLY -- SPMD emulation of the kernel's per-cell code under ThreadSanitizer.
//
// Every virtual thread of the CTA is an OS thread that runs the WHOLE per-cell program
// (sim_core.hpp, compiled with MODLE_B200_EMU_MT: regions execute once per thread, CTA barriers
// are pthread barriers, block-wide collectives go through the shared scratch). Built with
// -fsanitize=thread this is a happens-before race detector for the kernel source: a shared word
// written by one virtual thread and touched by another without a CTA barrier in between is
// reported whatever the actual interleaving was -- the class of bug compute-sanitizer's racecheck
// finds on the device, found here without a GPU.
//
//   emu_mt_main <case file> <virtual threads (<= 32)> <rng mode 0|1> [<output file>]
// Case file (written by tests/test_emulation_races.py): u64 nb, u64 ncells, modle_b200_sim_params,
// modle_b200_interval, nb x modle_b200_barrier, ncells x modle_b200_cell_task.
// Prints one line per cell (stats) and FNV-1a hashes of the band / 1D occupancy, which the test
// compares with the serial emulation of the same case.
#define MODLE_B200_EMU_MT 1
#include "emu_capi.cpp"

#include <cstdio>
#include <thread>

namespace {

u64 fnv1a(const void* p, size_t n) {
  const unsigned char* b = static_cast<const unsigned char*>(p);
  u64 h = 0xcbf29ce484222325ull;
  for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 0x100000001b3ull;
  return h;
}

template <bool kCtr>
void run_cell_spmd(EmuCell& cell, const Sinks& K, int nthreads, const modle_b200_cell_task& t) {
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, nullptr, static_cast<unsigned>(nthreads));
  std::vector<std::thread> th;
  for (int k = 0; k < nthreads; ++k) {
    th.emplace_back([&, k] {
      Cta cta{&cell.shared->scratch, nthreads, k, &bar};
      CellSimT<kCtr> sim{cell.kp, cell.D, cell.A, *cell.shared, K, cta, to_task(t)};
      sim.run();
    });
  }
  for (auto& x : th) x.join();
  pthread_barrier_destroy(&bar);
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 4) {
    std::fprintf(stderr, "usage: emu_mt_main <case file> <virtual threads> <rng mode>\n");
    return 2;
  }
  const int nthreads = std::atoi(argv[2]);
  const int mode = std::atoi(argv[3]);
  if (nthreads < 1 || nthreads > kMaxWarps) return 2;
  std::FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  u64 hdr[2];
  modle_b200_sim_params p;
  modle_b200_interval iv;
  bool ok = std::fread(hdr, sizeof(u64), 2, f) == 2 && std::fread(&p, sizeof(p), 1, f) == 1 &&
            std::fread(&iv, sizeof(iv), 1, f) == 1;
  std::vector<modle_b200_barrier> bars(ok ? hdr[0] : 0);
  std::vector<modle_b200_cell_task> tasks(ok ? hdr[1] : 0);
  ok = ok && std::fread(bars.data(), sizeof(modle_b200_barrier), bars.size(), f) == bars.size() &&
       std::fread(tasks.data(), sizeof(modle_b200_cell_task), tasks.size(), f) == tasks.size();
  std::fclose(f);
  if (!ok) {
    std::fprintf(stderr, "short case file\n");
    return 2;
  }
  EmuCell cell;
  const std::string err = cell.setup(p, iv, bars.data(), bars.size(), 0);
  if (!err.empty()) {
    std::fprintf(stderr, "%s\n", err.c_str());
    return 2;
  }
  std::vector<u32> band(size_t(cell.kp.nrows) * cell.kp.ncols + 1, 0);
  std::vector<u64> occ(std::max<u32>(cell.kp.ncols, 1), 0);
  u64 missed = 0;
  Sinks K{band.data(), cell.kp.track_1d ? occ.data() : nullptr, &missed, nullptr, 0};
  for (size_t c = 0; c < tasks.size(); ++c) {
    const bool has_work = cell.kp.stop_on_epochs || tasks[c].num_target_contacts != 0;
    if (has_work) {
      if (mode == MODLE_B200_RNG_COUNTER) {
        run_cell_spmd<true>(cell, K, nthreads, tasks[c]);
      } else {
        run_cell_spmd<false>(cell, K, nthreads, tasks[c]);
      }
    }
    const CellShared& S = *cell.shared;
    std::printf("cell %zu contacts %llu epochs %llu burnin %llu lef_updates %llu draws %llu fault %u\n",
                c, has_work ? (unsigned long long)S.num_contacts : 0ull,
                has_work ? (unsigned long long)S.epoch : 0ull,
                has_work ? (unsigned long long)S.num_burnin_epochs : 0ull,
                has_work ? (unsigned long long)S.lef_updates : 0ull,
                has_work ? (unsigned long long)S.rng_pos : 0ull, has_work ? S.fault : 0u);
  }
  if (argc > 4) {  // raw band | occ for the caller to compare
    std::FILE* o = std::fopen(argv[4], "wb");
    if (!o) return 2;
    std::fwrite(band.data(), sizeof(u32), band.size(), o);
    std::fwrite(occ.data(), sizeof(u64), cell.kp.ncols, o);
    std::fclose(o);
  }
  std::printf("band %016llx occ %016llx missed %llu\n",
              (unsigned long long)fnv1a(band.data(), band.size() * sizeof(u32)),
              (unsigned long long)fnv1a(occ.data(), size_t(cell.kp.ncols) * sizeof(u64)),
              (unsigned long long)missed);
  return 0;
}
