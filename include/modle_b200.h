/* modle_b200 -- C ABI of the B200-native loop-extrusion hot path.
 *
 * Drop-in boundary for ONE path of paulsengroup/modle (reference paths are relative to the
 * reference checkout): the per-(interval, cell) loop Simulation::simulate_one_cell
 * (src/libmodle/cpu/simulation.cpp:896-986) plus the contact registration it calls
 * (src/libmodle/cpu/register_contacts.cpp:93-232), i.e. what a worker thread does between popping
 * a Task and pushing it back as COMPLETED (src/libmodle/cpu/scheduler_simulate.cpp:220-261).
 * The reference has no FFI layer; the seam is the C++ class modle::Simulation
 * (src/libmodle/cpu/include/modle/simulation.hpp:45-151). INTEGRATION.md shows the binding a
 * maintainer would add inside Simulation::run_simulate.
 *
 * Everything is plain C: POD structs, pointers and sizes. All functions return 0 on success or a
 * negative modle_b200_status; modle_b200_last_error() returns a thread-local message. There is no
 * CPU fallback: every compute entry point fails with MODLE_B200_ERR_NO_DEVICE without a CUDA GPU.
 */
#ifndef MODLE_B200_H
#define MODLE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MODLE_B200_ABI_VERSION 1

typedef enum modle_b200_status {
  MODLE_B200_OK = 0,
  MODLE_B200_ERR_INVALID_ARGUMENT = -1,
  MODLE_B200_ERR_NO_DEVICE = -2,
  MODLE_B200_ERR_CUDA = -3,
  MODLE_B200_ERR_UNSUPPORTED = -4, /* e.g. interval does not fit the shared-memory-resident kernel */
  MODLE_B200_ERR_DEVICE_FAULT = -5 /* kernel-side invariant violated (reported, never ignored) */
} modle_b200_status;

/* dna::Direction values (src/common/include/modle/common/dna.hpp:77-80). A barrier stores its
 * BLOCKING direction: motif '+' blocks REV-moving units, '-' blocks FWD-moving units
 * (src/libmodle/internal/extrusion_barriers_impl.hpp:61-72). */
enum { MODLE_B200_DIR_REV = 1, MODLE_B200_DIR_FWD = 2 };

/* Config::ContactSamplingStrategy bits (simulation_config.hpp:33-38) */
enum {
  MODLE_B200_SAMPLE_NOISIFY = 1,
  MODLE_B200_SAMPLE_TAD = 2,
  MODLE_B200_SAMPLE_LOOP = 4
};

/* Config::StoppingCriterion (simulation_config.hpp:29) */
enum { MODLE_B200_STOP_CONTACT_DENSITY = 0, MODLE_B200_STOP_SIMULATION_EPOCHS = 1 };

/* The fields of modle::Config the path reads (simulation_config.hpp:53-113), holding the values
 * AFTER Cli::transform_args (src/modle/cli.cpp:993-1016). modle_b200_default_params() +
 * modle_b200_transform_params() reproduce that arithmetic. Same names as the reference. */
typedef struct modle_b200_sim_params {
  uint64_t bin_size;
  uint64_t diagonal_width;
  uint64_t rev_extrusion_speed;
  uint64_t fwd_extrusion_speed;
  uint64_t rev_extrusion_speed_burnin;
  uint64_t fwd_extrusion_speed_burnin;
  double rev_extrusion_speed_std;
  double fwd_extrusion_speed_std;
  double prob_of_lef_release;
  double prob_of_lef_release_burnin;
  double hard_stall_lef_stability_multiplier;
  double soft_stall_lef_stability_multiplier;
  double probability_of_extrusion_unit_bypass;
  double lef_bar_major_collision_pblock;
  double lef_bar_minor_collision_pblock;
  double tad_to_loop_contact_ratio;
  double genextreme_mu;
  double genextreme_sigma;
  double genextreme_xi;
  double number_of_lefs_per_mbp;
  double target_contact_density;
  uint64_t target_simulation_epochs;
  uint64_t contact_sampling_interval;
  uint64_t avg_lef_processivity;
  uint64_t probability_normalization_factor;
  double extrusion_barrier_occupancy;
  double barrier_occupied_stp;
  double barrier_not_occupied_stp;
  double burnin_speed_coefficient;
  uint64_t burnin_history_length;
  uint64_t burnin_smoothing_window_size;
  uint64_t min_burnin_epochs;
  uint64_t max_burnin_epochs;
  uint64_t burnin_target_epochs_for_lef_activation;
  uint64_t num_cells;
  uint64_t seed;
  uint32_t contact_sampling_strategy; /* MODLE_B200_SAMPLE_* bits */
  uint32_t stopping_criterion;        /* MODLE_B200_STOP_* */
  uint32_t track_1d_lef_position;
  uint32_t skip_burnin;
  uint32_t normalize_probabilities;
  uint32_t override_extrusion_barrier_occupancy;
  uint64_t debug_max_epochs; /* not in the reference: stop every cell after this many epochs
                                (UINT64_MAX = off); used by the parity tests to bisect */
} modle_b200_sim_params;

/* GenomicInterval geometry (src/libmodle/internal/include/modle/genome.hpp:125-195). */
typedef struct modle_b200_interval {
  uint64_t chrom_size;
  uint64_t start; /* simulated range is [start, end) */
  uint64_t end;
  uint64_t num_lefs; /* Simulation::compute_num_lefs (simulation.cpp:1086-1090) */
} modle_b200_interval;

/* One extrusion barrier (ExtrusionBarrier, extrusion_barriers.hpp:32-60); array sorted by pos
 * (a position outside [start, end) is allowed: the reference keeps such barriers). */
typedef struct modle_b200_barrier {
  uint64_t pos;
  double stp_active;
  double stp_inactive;
  uint32_t blocking_direction; /* MODLE_B200_DIR_* */
  uint32_t reserved_;
} modle_b200_barrier;

/* Simulation::Task minus the interval pointer (simulation.hpp:59-69). rng_state is the serialised
 * xoshiro256++ state the reference copies into the task (scheduler_simulate.cpp:143-158). */
typedef struct modle_b200_cell_task {
  uint64_t cell_id;
  uint64_t num_target_epochs;
  uint64_t num_target_contacts;
  uint64_t rng_state[4];
} modle_b200_cell_task;

/* What the reference only logs per task (scheduler_simulate.cpp:246-251) plus bookkeeping. */
typedef struct modle_b200_cell_stats {
  uint64_t num_contacts;
  uint64_t num_epochs;
  uint64_t num_burnin_epochs;
  uint64_t num_lef_updates; /* sum over simulated epochs of the number of active LEFs */
  uint64_t num_rng_draws;   /* raw 64-bit draws consumed from the cell's stream */
  uint64_t device_fault;    /* 0, or a kernel-side fault code */
} modle_b200_cell_stats;

/* Snapshot of one cell (debug/parity aid; arrays sized num_lefs / num_barriers by the caller). */
typedef struct modle_b200_cell_snapshot {
  uint64_t* rev_pos;       /* UINT64_MAX when unbound */
  uint64_t* fwd_pos;
  uint64_t* binding_epoch; /* UINT64_MAX when unbound */
  uint64_t* rev_ranks;
  uint64_t* fwd_ranks;
  uint8_t* barrier_active;
  uint64_t num_active_lefs;
  uint64_t burnin_completed;
} modle_b200_cell_snapshot;

typedef struct modle_b200_context modle_b200_context;

/* ---- host-only helpers (no GPU needed) ------------------------------------------------- */

int modle_b200_abi_version(void);
const char* modle_b200_last_error(void);

/* Config{} defaults (simulation_config.hpp:53-113), untransformed. */
void modle_b200_default_params(modle_b200_sim_params* p);
/* Cli::transform_args (cli.cpp:886-1016). `*_given` say which CLI options were passed explicitly
 * (the reference asks CLI11 for that): rev/fwd speed and --extrusion-barrier-occupancy. */
int modle_b200_transform_params(modle_b200_sim_params* p, int rev_speed_given, int fwd_speed_given,
                                int barrier_occupancy_given);

/* Simulation::compute_num_lefs / compute_contacts_per_epoch (simulation.cpp:1076-1090). */
uint64_t modle_b200_compute_num_lefs(const modle_b200_sim_params* p, uint64_t interval_size_bp);
uint64_t modle_b200_compute_contacts_per_epoch(const modle_b200_sim_params* p, uint64_t num_lefs);
/* ContactMatrixDense geometry (contact_matrix_dense_impl.hpp:39-50): nrows, ncols; the band
 * buffer holds nrows*ncols+1 uint32 with pixel (i=|b1-b2|, j=max(b1,b2)) at j*nrows+i. */
void modle_b200_band_shape(const modle_b200_sim_params* p, uint64_t interval_size_bp,
                           uint64_t* nrows, uint64_t* ncols);

/* GenomicInterval::hash with Config::seed (genome.cpp:201-224): XXH3-64(seed) over
 * name || u64 chrom_size || u64 start || u64 end. */
int modle_b200_interval_hash(const char* chrom_name, size_t name_len, uint64_t chrom_size,
                             uint64_t start, uint64_t end, uint64_t seed, uint64_t* out);
/* random::PRNG(seed) (common/random.hpp:26-30) and Xoshiro256PlusPlus::jump(). */
void modle_b200_rng_seed(uint64_t seed, uint64_t state[4]);
uint64_t modle_b200_rng_next(uint64_t state[4]);
void modle_b200_rng_jump(uint64_t state[4]);

/* Barrier self-transition probabilities from a BED score (genome.cpp:255-271) and
 * the occupancy <-> stp identities (extrusion_barriers_impl.hpp:106-128). */
double modle_b200_stp_active_from_occupancy(double stp_inactive, double occupancy);
double modle_b200_occupancy_from_stp(double stp_active, double stp_inactive);

/* The per-interval task fan-out of Simulation::run_simulate (scheduler_simulate.cpp:104-160):
 * engine = PRNG(interval hash), per-cell targets (:129-141), one jump() per cell (:158).
 * Fills tasks[0..num_cells). */
int modle_b200_make_cell_tasks(const modle_b200_sim_params* p, const char* chrom_name,
                               size_t name_len, const modle_b200_interval* interval,
                               modle_b200_cell_task* tasks);

/* ---- genome import: chrom.sizes + barrier BED6 (+ genomic-intervals BED3) -> intervals ---------
 * Host-only. Reproduces Genome::Genome (src/libmodle/internal/genome.cpp:299-330) with the
 * parsers it uses (chrom_sizes::Parser, src/libmodle_io/chrom_sizes.cpp:24-66; bed::Parser and
 * bed::BED, src/libmodle_io/bed.cpp:44-330,411-590), Genome::map_barriers_to_intervals /
 * generate_barriers_from_bed_records (genome.cpp:423-488: pos = (start + end + 1) / 2, '.' strands
 * skipped, score -> occupancy -> stp) and Simulation's --extrusion-barrier-occupancy override
 * (src/libmodle/cpu/simulation.cpp:51-60). Plain-text files only (no libarchive here). Intervals
 * come out in the reference's processing order (chromosomes in chrom.sizes order, intervals by
 * start), barriers sorted by position; malformed input fails with
 * MODLE_B200_ERR_INVALID_ARGUMENT and the reference's diagnostic in modle_b200_last_error().
 * path_to_genomic_intervals may be NULL or "" (whole chromosomes).                              */
typedef struct modle_b200_genome modle_b200_genome;
int modle_b200_genome_import(const char* path_to_chrom_sizes, const char* path_to_extr_barriers,
                             const char* path_to_genomic_intervals,
                             const modle_b200_sim_params* params,
                             int interpret_name_field_as_puu, modle_b200_genome** out);
void modle_b200_genome_free(modle_b200_genome* genome);
size_t modle_b200_genome_num_chromosomes(const modle_b200_genome* genome);
size_t modle_b200_genome_num_intervals(const modle_b200_genome* genome);
uint64_t modle_b200_genome_num_barriers(const modle_b200_genome* genome);
/* The i-th interval. Any output pointer may be NULL. *chrom_name and *barriers stay valid until
 * modle_b200_genome_free. *bin_offset is what append_contact_matrix_to_cooler adds to the
 * interval's pixel coordinates (first bin of the chromosome + start / bin_size).               */
int modle_b200_genome_get_interval(const modle_b200_genome* genome, size_t i,
                                   const char** chrom_name, size_t* chrom_id,
                                   uint64_t* chrom_size, uint64_t* start, uint64_t* end,
                                   const modle_b200_barrier** barriers, size_t* num_barriers,
                                   uint64_t* bin_offset);

/* ---- device path ------------------------------------------------------------------------ */

/* Binds a context to CUDA device `device`; fails without a GPU (no CPU fallback). */
int modle_b200_init(modle_b200_context** ctx, int device);
void modle_b200_destroy(modle_b200_context* ctx);

/* How the cells launched through this context take their random draws.
 *   MODLE_B200_RNG_REFERENCE_ORDER (default) "deterministic mode": each cell consumes ONE
 *     xoshiro256++ stream in exactly the order simulate_one_cell does
 *     (src/libmodle/cpu/simulation.cpp:896-986 and callees), so trajectories and contact counts
 *     are bit-identical to the reference algorithm for the same seed.
 *   MODLE_B200_RNG_COUNTER "throughput mode": same distributions, but every (epoch, phase, item)
 *     reads a private counter-based sequence keyed by the cell's task state, which removes the
 *     draw staging, the offset scans and the speculation/repair the sequential order costs.
 *     Results are a pure function of the task (independent of grid, CTA width, stream and GPU
 *     count) and statistically equivalent to the reference's, NOT bit-identical to them;
 *     modle_b200_cell_stats::num_rng_draws is reported as 0.
 * The mode applies to the launches issued after the call. */
enum { MODLE_B200_RNG_REFERENCE_ORDER = 0, MODLE_B200_RNG_COUNTER = 1 };
int modle_b200_set_rng_mode(modle_b200_context* ctx, int mode);
int modle_b200_get_rng_mode(const modle_b200_context* ctx);

/* Simulates `num_cells` cells of one interval: the GPU replacement for popping num_cells Tasks
 * and running simulate_one_cell on each. HOST buffers in, HOST buffers out (copies included):
 *   band_out   nrows*ncols+1 uint32, ADDED to (caller zero-initialises; reference layout)
 *   occ1d_out  ncols uint64 or NULL, ADDED to (GenomicInterval::lef_1d_occupancy)
 *   stats_out  num_cells entries or NULL
 *   missed_updates_out  ContactMatrixDense::_updates_missed increment, or NULL            */
int modle_b200_simulate_interval(modle_b200_context* ctx, const modle_b200_sim_params* params,
                                 const modle_b200_interval* interval,
                                 const modle_b200_barrier* barriers, size_t num_barriers,
                                 const modle_b200_cell_task* tasks, size_t num_cells,
                                 uint32_t* band_out, uint64_t* occ1d_out,
                                 modle_b200_cell_stats* stats_out, uint64_t* missed_updates_out);

/* Same call, but band_out / occ1d_out / *missed_updates_out are OVERWRITTEN with this call's
 * results instead of added to: for the usual case of one call per interval (all its cells at
 * once) the caller then neither has to zero 4 x nrows x ncols bytes nor pay a read-modify-write
 * pass over them. */
int modle_b200_simulate_interval_overwrite(modle_b200_context* ctx,
                                           const modle_b200_sim_params* params,
                                           const modle_b200_interval* interval,
                                           const modle_b200_barrier* barriers, size_t num_barriers,
                                           const modle_b200_cell_task* tasks, size_t num_cells,
                                           uint32_t* band_out, uint64_t* occ1d_out,
                                           modle_b200_cell_stats* stats_out,
                                           uint64_t* missed_updates_out);

/* Optional: sizes the context's device and pinned staging buffers once for the largest interval
 * the caller will pass to the host-buffer entry points (band of max_nrows x max_ncols, max_cells
 * cells) -- what Simulation::State::resize_buffers (simulation.cpp:603-627) does for a reference
 * worker. Without it the buffers grow on demand, at the price of device-wide synchronisations. */
int modle_b200_reserve(modle_b200_context* ctx, uint64_t max_nrows, uint64_t max_ncols,
                       size_t max_cells);

/* Same, with every buffer already resident in device memory (d_* are device pointers in the
 * context's device; cuda_stream is a cudaStream_t or NULL for the context's own stream). The call
 * is asynchronous; d_missed_updates (1 uint64) and d_stats accumulate on the device.            */
int modle_b200_simulate_interval_device(modle_b200_context* ctx,
                                        const modle_b200_sim_params* params,
                                        const modle_b200_interval* interval,
                                        const modle_b200_barrier* d_barriers, size_t num_barriers,
                                        const modle_b200_cell_task* d_tasks, size_t num_cells,
                                        uint32_t* d_band, uint64_t* d_occ1d,
                                        modle_b200_cell_stats* d_stats, uint64_t* d_missed_updates,
                                        void* cuda_stream);
int modle_b200_synchronize(modle_b200_context* ctx);

/* ---- internal-state log (Config::log_model_internal_state) --------------------------------------
 * One record per simulated epoch of a cell: the quantities Simulation::dump_stats
 * (src/libmodle/cpu/simulation.cpp:995-1056) writes, taken where the reference takes them (after
 * extrude, before release_lefs, :968-974). The text columns it adds (task id, cell id, chromosome,
 * start, end) are known to the caller; effective_barrier_occupancy = barriers_occupied /
 * num_barriers and avg_loop_size = loop_size_sum / num_lefs (the integer sum is exact in double,
 * so the reference's left-to-right double accumulation gives the same value).                   */
typedef struct modle_b200_epoch_record {
  uint64_t epoch;
  uint64_t loop_size_sum;
  uint32_t burnin; /* 1 while the cell is in its burn-in phase */
  uint32_t num_lefs; /* active LEFs (lefs.size() in dump_stats) */
  uint32_t barriers_occupied;
  uint32_t lefs_stalled_rev;
  uint32_t lefs_stalled_fwd;
  uint32_t lefs_stalled_both;
  uint32_t lef_bar_collisions;
  uint32_t lef_lef_primary_collisions;
  uint32_t lef_lef_secondary_collisions;
  uint32_t reserved_;
} modle_b200_epoch_record;

/* modle_b200_simulate_interval plus the log: cell k's record of epoch e goes to
 * log_out[k * log_capacity_per_cell + e] for e < log_capacity_per_cell (later epochs are not
 * logged; stats_out[k].num_epochs says how many there were). HOST buffers.                     */
int modle_b200_simulate_interval_logged(modle_b200_context* ctx,
                                        const modle_b200_sim_params* params,
                                        const modle_b200_interval* interval,
                                        const modle_b200_barrier* barriers, size_t num_barriers,
                                        const modle_b200_cell_task* tasks, size_t num_cells,
                                        uint32_t* band_out, uint64_t* occ1d_out,
                                        modle_b200_cell_stats* stats_out,
                                        uint64_t* missed_updates_out,
                                        modle_b200_epoch_record* log_out,
                                        size_t log_capacity_per_cell);

/* Runs one cell for params->debug_max_epochs epochs and returns its state (parity bisection). */
int modle_b200_snapshot_cell(modle_b200_context* ctx, const modle_b200_sim_params* params,
                             const modle_b200_interval* interval,
                             const modle_b200_barrier* barriers, size_t num_barriers,
                             const modle_b200_cell_task* task, modle_b200_cell_snapshot* snapshot,
                             modle_b200_cell_stats* stats_out);

/* Contact-register kernel in isolation (register_contacts.cpp:149-152 + ContactMatrixDense::
 * increment, contact_matrix_dense_safe_impl.hpp:54-89): scatters n (bin1, bin2) pairs held in
 * device memory into the device band; used to replay a contact stream for the roofline run. */
int modle_b200_register_contacts_device(modle_b200_context* ctx, const uint32_t* d_bin1,
                                        const uint32_t* d_bin2, size_t n, uint64_t nrows,
                                        uint64_t ncols, uint32_t* d_band,
                                        uint64_t* d_missed_updates, void* cuda_stream);

/* Measurement aid for that kernel's roofline ("calibrate an atomic roofline on the box
 * with a micro-benchmark: uniform random red.global.add.u32 over a buffer of the same footprint"):
 * issues num_reductions reductions of +1 on pseudo-random words of d_band[0..num_words) -- the
 * addresses come from a counter hash, nothing is read -- asynchronously on `cuda_stream`. The
 * caller times it; the sum over d_band grows by num_reductions. No reference counterpart. */
int modle_b200_calibrate_red_device(modle_b200_context* ctx, uint32_t* d_band, uint64_t num_words,
                                    uint64_t num_reductions, uint64_t seed, void* cuda_stream);

/* ---- band -> sorted COO pixels: the hand-off to the .cool writer -------------------------------
 * Replaces the pixel loop of modle::io::internal::append_contact_matrix_to_cooler
 * (src/libmodle_io/contact_matrix_dense_io_impl.hpp:50-71): for i in [0, ncols), j in
 * [i, min(ncols, i + nrows)), every non-zero matrix.unsafe_get(i, j) becomes
 * {bin_offset + i, bin_offset + j, int32(count)}, in that (bin1, bin2)-sorted order. bin_offset is
 * what emplace_pixel adds (:30-43): first bin id of the chromosome + interval start / bin size.
 * modle_b200_pixel has the memory layout of hictk::ThinPixel<std::int32_t> (two uint64 bin ids,
 * int32 count, 4 bytes of tail padding), so the output can be handed to
 * hictk::cooler::File::append_pixels as is.                                                      */
typedef struct modle_b200_pixel {
  uint64_t bin1_id;
  uint64_t bin2_id;
  int32_t count;
  int32_t reserved_; /* tail padding of ThinPixel<int32_t>; written as 0 */
} modle_b200_pixel;

/* Device-resident, asynchronous on `cuda_stream`. Step 1 fills d_row_offsets[0..ncols] (ncols+1
 * uint64): d_row_offsets[r] = index of the first pixel of row r, d_row_offsets[ncols] = number of
 * non-zero pixels. Step 2 writes the pixels (at most `capacity`; size it from step 1).          */
int modle_b200_count_pixels_device(modle_b200_context* ctx, const uint32_t* d_band, uint64_t nrows,
                                   uint64_t ncols, uint64_t* d_row_offsets, void* cuda_stream);
int modle_b200_fill_pixels_device(modle_b200_context* ctx, const uint32_t* d_band, uint64_t nrows,
                                  uint64_t ncols, uint64_t bin_offset,
                                  const uint64_t* d_row_offsets, modle_b200_pixel* d_pixels,
                                  uint64_t capacity, void* cuda_stream);
/* HOST buffers in and out (copies included). *num_pixels_out is always set to the number of
 * non-zero pixels; with pixels_out == NULL and capacity == 0 the call is a size query, with a
 * too small capacity it fails with MODLE_B200_ERR_INVALID_ARGUMENT and writes nothing.          */
int modle_b200_band_to_pixels(modle_b200_context* ctx, const uint32_t* band, uint64_t nrows,
                              uint64_t ncols, uint64_t bin_offset, modle_b200_pixel* pixels_out,
                              uint64_t capacity, uint64_t* num_pixels_out);

/* ---- 1D LEF occupancy profile: what write_lef_occupancy_to_bwig hands to the bigWig writer ------
 * (src/libmodle/cpu/simulation.cpp:170-197): profile[i] = float(double(occ[i]) / double(max occ)).
 * An all-zero track yields NaN (0/0), as in the reference. Device variant: d_scratch_max is one
 * uint64 of scratch; asynchronous on `cuda_stream`.                                             */
int modle_b200_lef_occupancy_profile_device(modle_b200_context* ctx, const uint64_t* d_occ1d,
                                            size_t n, float* d_profile, uint64_t* d_scratch_max,
                                            void* cuda_stream);
int modle_b200_lef_occupancy_profile(modle_b200_context* ctx, const uint64_t* occ1d, size_t n,
                                     float* profile_out);

/* Profiling aid (the reference has none; --skip-output + perf is its recipe, cli.cpp:182-186):
 * SM-clock cycles the simulate kernel spent in each phase of the per-cell loop, summed over all
 * cells simulated on this context since the last reset. Slots, in order: init, burn-in, bind,
 * rank, contacts, move generation, move adjust/clamp, barrier states, LEF-BAR detect, primary
 * LEF-LEF detect, move correction, secondary LEF-LEF, rank fix-up, extrude+release, RNG refill
 * (nested inside the others), whole cell; then a finer split of move generation (4 slots) and
 * of the secondary pass (6 slots). Waits for the context's own stream first. */
#define MODLE_B200_NUM_PHASES 26
int modle_b200_phase_cycles(modle_b200_context* ctx, uint64_t* out, size_t n, int reset);

/* Host-only: how the library would launch cells of an interval with `num_lefs` LEFs and
 * `num_barriers` barriers -- threads per CTA (one CTA simulates one cell), cells resident per SM
 * (3, 2 or 1: what the cell's shared-memory state leaves room for) and bytes of shared memory per
 * cell. Lets a caller weigh (interval, cell) work when it spreads it over several GPUs; the
 * reference has no counterpart (its workers are interchangeable CPU threads). */
int modle_b200_launch_geometry(uint64_t num_lefs, uint64_t num_barriers, uint32_t* cta_threads,
                               uint32_t* cells_per_sm, uint64_t* shared_bytes_per_cell);

/* ---- several GPUs (one host thread or process per GPU, one context each) ------------------------
 * The reference's parallelism is its pool of worker threads popping (interval, cell) tasks that
 * all add into the interval's shared matrix (scheduler_simulate.cpp:104-160, 190-271;
 * contact_matrix_dense_safe_impl.hpp:54-68). Over several GPUs the tasks are dealt out as
 * (interval, cell range) shards; an interval whose cells sit on several GPUs has its band summed
 * onto its root with ONE reduce. Every cell keeps the task it has in the unsharded run
 * (modle_b200_make_cell_tasks over all cells, then sliced), so results do not depend on the plan. */
typedef struct modle_b200_shard {
  uint64_t interval; /* index into the caller's interval list */
  uint64_t cell_lo;  /* cells [cell_lo, cell_hi) of that interval */
  uint64_t cell_hi;
  int32_t rank;      /* GPU / process that simulates the shard */
  int32_t reserved_;
  double weight;     /* modelled cost (cell weight x cells) */
} modle_b200_shard;

/* Planner weight of ONE cell of an interval (relative SM time: launch geometry + per-LEF cost). */
double modle_b200_cell_weight(uint64_t num_lefs, uint64_t num_barriers);

/* Deals the cells of `num_intervals` intervals (`num_cells` cells each; cell_weights[i] from
 * modle_b200_cell_weight, 0 = interval skipped, e.g. no barriers) to `world_size` ranks: whole
 * intervals heaviest-first; while the heaviest rank carries more than `tolerance` (<= 0: 1.10) x
 * the mean, the heaviest piece of that rank is halved by cells. slice_all > 0: every interval is
 * cut into one cell range per rank instead (one reduce per interval); slice_all < 0: the library
 * decides (slices when a slice still holds >= 222 cells, i.e. 1.5 waves over the SMs). Shards come out sorted by
 * (interval, cell_lo); the root of an interval is the rank of its first shard. Deterministic:
 * every rank computes the same plan. shards_out == NULL with capacity 0 is a size query. */
int modle_b200_plan_shards(const double* cell_weights, size_t num_intervals, uint64_t num_cells,
                           int world_size, int slice_all, double tolerance,
                           modle_b200_shard* shards_out, size_t capacity, size_t* num_shards_out);

/* The one collective of the path: sums a split interval's device buffers onto `root` -- band
 * (nrows*ncols+1 uint32), and when not NULL the 1D track (ncols uint64) and the missed-update
 * counter (1 uint64) -- with ncclReduce(sum), in place, asynchronously on `cuda_stream`. Called
 * by every rank of `nccl_comm` (an ncclComm_t; a rank without a piece of the interval passes
 * zeroed buffers). NCCL is loaded at run time (libnccl.so.2); MODLE_B200_ERR_UNSUPPORTED if it
 * cannot be. Inside an ncclGroupStart/End when one thread drives several GPUs. */
int modle_b200_reduce_band(modle_b200_context* ctx, void* nccl_comm, uint32_t* d_band,
                           uint64_t nrows, uint64_t ncols, uint64_t* d_occ1d,
                           uint64_t* d_missed_updates, int root, void* cuda_stream);

/* Number of kernels this library has launched on the context so far (bench bookkeeping). */
uint64_t modle_b200_kernel_launches(const modle_b200_context* ctx);

#ifdef __cplusplus
}
#endif
#endif /* MODLE_B200_H */
