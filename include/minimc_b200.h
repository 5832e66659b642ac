/* minimc_b200 -- C ABI of the B200-native particle-history transport loop.
 *
 * This is the drop-in boundary for the hot path of agtumulak/minimc
 * (reference @ ed536a2).  The reference has no FFI: its seam is the C++
 * factory `Driver::Create(path)` + `virtual EstimatorSet Driver::Solve()`
 * (src/Driver.hpp:23-30), whose fixed-source implementation spawns
 * `threads` x `std::async(&FixedSource::StartWorker)` (src/FixedSource.cpp:22-36).
 * Everything below replaces what happens *inside* that Solve(): the host keeps
 * its World/Material/Nuclide/Estimator objects, flattens them ONCE into the
 * plain-old-data tables declared here (in the host containers' own iteration
 * order), and calls these entry points instead of the per-thread history loop.
 * INTEGRATION.md shows the ~80-line Driver subclass a maintainer would add.
 *
 * Conventions: plain pointers and sizes, no C++ or torch types; every function
 * returns an mmc_status (0 = ok) and never throws; mmc_last_error() gives the
 * text for the calling thread.  All tables are copied during the call, the
 * caller keeps ownership of its buffers.  There is NO CPU fallback: without a
 * CUDA device every compute entry point returns MMC_ERR_NO_DEVICE.
 */
#ifndef MINIMC_B200_H
#define MINIMC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMC_ABI_VERSION 2

typedef enum mmc_status {
  MMC_OK = 0,
  MMC_ERR_INVALID = 1,      /* malformed table / argument (message says which) */
  MMC_ERR_NO_DEVICE = 2,    /* no CUDA device: this library has no CPU path */
  MMC_ERR_CUDA = 3,         /* CUDA runtime error */
  MMC_ERR_LOST_PARTICLE = 4,/* World::FindCellContaining found no cell (World.cpp:34-36 throws) */
  MMC_ERR_CAPACITY = 5,     /* per-history secondary queue / pending-score table / trace buffer overflow */
  MMC_ERR_PHYSICS = 6       /* an assert(false) branch of the reference was reached (e.g. Particle.cpp:123) */
} mmc_status;

/* CSGSurface.cpp:107-174 */
typedef enum mmc_surface_type { MMC_SURF_SPHERE = 0, MMC_SURF_PLANEX = 1, MMC_SURF_CYLINDERX = 2 } mmc_surface_type;
/* TransportMethod.cpp:24-46 */
typedef enum mmc_tracking { MMC_TRACK_SURFACE = 0, MMC_TRACK_CELL_DELTA = 1 } mmc_tracking;
/* Particle.hpp:37-45 -- same numbering as Particle::Event */
typedef enum mmc_event {
  MMC_EV_BIRTH = 0, MMC_EV_SCATTER = 1, MMC_EV_CAPTURE = 2, MMC_EV_FISSION = 3,
  MMC_EV_SURFACE_CROSS = 4, MMC_EV_LEAK = 5, MMC_EV_VIRTUAL_COLLISION = 6
} mmc_event;
/* RNG stream.  MINSTD_COMPAT reproduces std::minstd_rand consumed through
 * libstdc++ 13 <random> (BasicTypes.hpp:27) bit for bit.  COUNTER gives every particle its own
 * counter-based stream (Philox-2x32-10: draw n of stream id is one block of 64 bits, no sequential state): the
 * stream of a source particle is the one its history's seed names, a secondary's id is 64 bits drawn from its parent
 * -- statistically equivalent to MINSTD_COMPAT (other random numbers), and like it independent of batch splits,
 * schedules and GPU counts.  In COUNTER mode mmc_site.seed / .reserved hold the low / high word of the stream id and
 * mmc_event_record.rng_state = (id low word << 32) | draws made. */
typedef enum mmc_rng_mode { MMC_RNG_MINSTD_COMPAT = 0, MMC_RNG_COUNTER = 1 } mmc_rng_mode;
/* Source.cpp:44-58 */
typedef enum mmc_direction_kind { MMC_DIR_CONSTANT = 0, MMC_DIR_ISOTROPIC = 1, MMC_DIR_ISOTROPIC_FLUX = 2 } mmc_direction_kind;
/* Bins.cpp:19-40 */
typedef enum mmc_bins_kind { MMC_BINS_NONE = 0, MMC_BINS_LINSPACE = 1, MMC_BINS_LOGSPACE = 2, MMC_BINS_BOUNDARIES = 3 } mmc_bins_kind;
/* Reaction.hpp:9-13 -- bit positions of mmc_world_desc.mg_reaction_mask */
enum { MMC_REACTION_CAPTURE = 1, MMC_REACTION_SCATTER = 2, MMC_REACTION_FISSION = 4 };
/* ScalarField.cpp:35-58 */
typedef enum mmc_field_kind { MMC_FIELD_CONSTANT = 0, MMC_FIELD_LINEAR = 1 } mmc_field_kind;

/* ---- continuous-energy tables (n_groups == 0) -------------------------------
 * Every pointwise table is the sorted, de-duplicated key/value content of the
 * std::map behind ContinuousMap (ContinuousMap.hpp:21-56): x strictly
 * increasing, lin-lin interpolation, clamped at both ends. */
typedef struct mmc_table1d {
  uint64_t n;
  const double* x;                  /* [n] */
  const double* y;                  /* [n] */
} mmc_table1d;

/* ThermalScattering::BetaPartition / AlphaPartition (ThermalScattering.hpp:62-100):
 * value(cdf_index, grid_index, T) = lerp_T( sum_r S[r] * CDF_modes[cdf][r] * grid_T_modes[grid][T_k][r] ).
 * Layouts are the row-major HDF5DataSet<D> layouts (HDF5DataSet.hpp:152-167). */
typedef struct mmc_tsl_partition {
  uint64_t n_cdf, n_grid, n_temperature, rank;
  const double* cdf;                /* [n_cdf]   CDF_modes.GetAxis(0): the F values */
  const double* grid;               /* [n_grid]  E_T_modes / beta_T_modes .GetAxis(0): incident E (MeV) or beta */
  const double* temperature;        /* [n_temperature] .GetAxis(1) */
  const double* cdf_modes;          /* [n_cdf][rank] */
  const double* singular_values;    /* [rank] */
  const double* grid_T_modes;       /* [n_grid][n_temperature][rank] */
} mmc_tsl_partition;

/* ThermalScattering (ThermalScattering.cpp:24-104). */
typedef struct mmc_tsl_desc {
  mmc_table1d majorant;             /* ThermalScattering::majorant */
  uint64_t n_energy, n_temperature, rank;
  const double* energy;             /* [n_energy]      scatter_xs_E.GetAxis(0); cutoff_energy = last */
  const double* temperature;        /* [n_temperature] scatter_xs_T.GetAxis(0) */
  const double* xs_E;               /* [n_energy][rank] */
  const double* xs_S;               /* [rank] */
  const double* xs_T;               /* [n_temperature][rank] */
  int32_t n_beta_partitions, n_alpha_partitions;
  const mmc_tsl_partition* beta_partitions;
  const mmc_tsl_partition* alpha_partitions;
  double beta_cutoff, alpha_cutoff;
  double awr;                       /* ThermalScattering::awr (the nuclide's) */
} mmc_tsl_desc;

/* One ContinuousReaction (ContinuousReaction.cpp:41-265), in XML document order. */
typedef struct mmc_ce_reaction {
  int32_t kind;                     /* MMC_REACTION_CAPTURE / _SCATTER / _FISSION */
  mmc_table1d xs;                   /* evaluation.xs */
  double temperature;               /* evaluation.temperature */
  const mmc_tsl_desc* tsl;          /* scatter only; NULL when the deck has no <tsl> */
  int32_t has_nubar;                /* fission only */
  mmc_table1d nubar;
} mmc_ce_reaction;

/* One Continuous interaction (Continuous.cpp:22-90). */
typedef struct mmc_ce_nuclide {
  double awr;                       /* <nuclide awr=...> */
  mmc_table1d total;                /* Continuous::total.xs */
  double total_temperature;
  int32_t n_reactions;
  const mmc_ce_reaction* reactions; /* [n_reactions] */
} mmc_ce_nuclide;

typedef struct mmc_ce_desc {
  const mmc_ce_nuclide* nuclides;   /* [mmc_world_desc.n_nuclides] */
} mmc_ce_desc;

/* Flattened `const World` (World.hpp).  Index order is the reference's:
 * surfaces / nuclides / materials in World creation order (World.cpp:89-168),
 * cells in XML order (first match wins, World.cpp:26-37), per-cell surfaces and
 * per-material nuclides in the order the host's std::map iterates them
 * (Cell.hpp:30, Material.hpp:20 -- pointer-keyed). */
typedef struct mmc_world_desc {
  uint32_t struct_size;             /* sizeof(mmc_world_desc) -- ABI check */
  uint32_t abi_version;             /* MMC_ABI_VERSION */

  int32_t n_surfaces;
  const int32_t* surface_type;      /* [n_surfaces] mmc_surface_type */
  const double* surface_param;      /* [n_surfaces][4] sphere: cx cy cz r; planex: c; cylinderx: r */

  int32_t n_cells;
  const int32_t* cell_material;     /* [n_cells] material index, -1 = void */
  const int32_t* cell_surface_begin;/* [n_cells+1] CSR offsets */
  const int32_t* cell_surface_index;/* [nnz] */
  const int32_t* cell_surface_sense;/* [nnz] 1: cell lies where CSGSurface::Contains() is true ("-1" in XML) */
  const int32_t* cell_field_kind;   /* [n_cells] mmc_field_kind of the temperature field (Cell.cpp:107-113) */
  const double* cell_field_param;   /* [n_cells][6] constant: c,-,-,-,upper,lower ; linear: gx,gy,gz,b,upper,lower */

  int32_t n_materials;
  const double* material_aden;      /* [n_materials] Material::number_density */
  const int32_t* material_nuclide_begin; /* [n_materials+1] CSR offsets */
  const int32_t* material_nuclide_index; /* [nnz] */
  const double* material_nuclide_afrac;  /* [nnz] normalised as Material.cpp:83-92 */

  int32_t n_nuclides;
  int32_t n_groups;                 /* G >= 1: multigroup world; 0: continuous-energy world */

  /* Multigroup.cpp:23-47,124-186.  All [n_nuclides][G] unless noted; absent
   * reactions have mask bit 0 and zero rows. */
  const uint32_t* mg_reaction_mask; /* [n_nuclides] */
  const double* mg_total;           /* CreateTotalXS: ((0+capture)+scatter)+fission over PRESENT reactions */
  const double* mg_capture;
  const double* mg_scatter;         /* column sums of the scatter matrix */
  const double* mg_fission;
  const double* mg_nubar;
  const double* mg_scatter_probs;   /* [n_nuclides][G_in][G_out] NormalizedTwoDimensional */
  const double* mg_chi;             /* [n_nuclides][G_in][G_out] */

  /* continuous-energy tables: see mmc_ce_desc (NULL for multigroup worlds) */
  const mmc_ce_desc* ce;
} mmc_world_desc;

/* Source.cpp:131-154: constant position, constant / isotropic / isotropic-flux
 * direction, constant energy, neutrons. */
typedef struct mmc_source_desc {
  double position[3];
  int32_t direction_kind;           /* mmc_direction_kind */
  double direction[3];              /* constant direction or isotropic-flux reference; normalised by the callee */
  uint64_t group;                   /* multigroup worlds: 1..G */
  double energy;                    /* continuous worlds: MeV */
} mmc_source_desc;

/* One axis of ParticleBins (Bins.cpp).  n_bins counts the two unbounded end bins. */
typedef struct mmc_bins_desc {
  int32_t kind;                     /* mmc_bins_kind */
  uint64_t n_bins;                  /* NONE: 1; LIN/LOG: bins+2; BOUNDARIES: n_boundaries+1 */
  double lower, upper, width;       /* LIN: bounds and bin width; LOG: log-space bounds and width */
  double base;                      /* LOG */
  const double* boundaries;         /* BOUNDARIES: [n_bins-1] strictly increasing */
} mmc_bins_desc;

/* CurrentEstimator (Estimator.cpp:116-151) + its ParticleBins (Bins.cpp:174-204). */
typedef struct mmc_estimator_desc {
  int32_t surface;                  /* index into mmc_world_desc surfaces */
  int32_t has_cosine_direction;     /* <cosine u v w> present */
  double cosine_direction[3];       /* normalised by the callee (Direction ctor) */
  mmc_bins_desc cosine;
  mmc_bins_desc energy;
} mmc_estimator_desc;

/* Device-side bookkeeping; also the inputs of the roofline formula
 * (BASELINE.md section 4): bytes = 72*births + 144*events + 16*scores + 144*banked. */
typedef struct mmc_counters {
  uint64_t n_histories;             /* source particles started */
  uint64_t n_births;                /* particles started: sources + secondaries */
  uint64_t n_events;                /* iterations of the Transport loop (= EstimatorSetProxy::Score calls) */
  uint64_t n_collisions;            /* real collisions (capture+scatter+fission) */
  uint64_t n_crossings;             /* surface_cross + leak */
  uint64_t n_virtual;               /* virtual collisions (cell delta tracking) */
  uint64_t n_scores;                /* non-zero scores accumulated */
  uint64_t n_secondaries;           /* fission secondaries produced */
  uint64_t n_banked;                /* sites written to the next-generation bank (k-eigenvalue) */
  uint64_t n_lost;                  /* particles outside every cell */
  uint64_t n_capacity_overflow;     /* secondary queue / pending table overflows */
  uint64_t n_physics_errors;        /* assert(false) branches reached */
} mmc_counters;

/* One event of one particle, as seen by EstimatorSetProxy::Score (TransportMethod.cpp:74). */
typedef struct mmc_event_record {
  uint64_t history;                 /* history index (seed = seed0 + history) */
  uint32_t particle;                /* ordinal of the particle within its history, reference bank order */
  int32_t event;                    /* mmc_event; MMC_EV_BIRTH records precede each particle's first event */
  uint64_t group;                   /* multigroup worlds */
  double energy;                    /* continuous worlds */
  int32_t cell;                     /* -1 before the first SetCell */
  int32_t surface;                  /* Particle::current_surface, -1 if none yet */
  double position[3];
  double direction[3];
  uint64_t rng_state;               /* minstd_rand state after the event */
} mmc_event_record;

/* opaque: owns the device copy of the tables and the scratch of a run.  ONE run at a time per world: concurrent
 * calls from several host threads serialise on the world; the asynchronous device-buffer entry points
 * (mmc_fixed_source_run_device, mmc_generation_run, ...) must be given the same stream for as long as work of an
 * earlier call may still be in flight -- two streams on one world would race on that scratch. */
typedef struct mmc_world mmc_world;

/* How the histories of a fixed-source run are scheduled on the device (results are identical either way).
 * FUSED: one persistent kernel, a particle stays in registers from birth to death (multigroup worlds always).
 * EVENT: event-split -- particle state in HBM, one flight kernel + one S(a,b) kernel per event, live particles
 *        stream-compacted in between (continuous-energy worlds; the default for them).  The device variant of the
 *        call then synchronises the stream every few passes to read the number of live histories.
 * EVENT_ONLY: as EVENT, without the hand-over of the last few thousand live histories to the fused kernel that EVENT
 *        makes once every history has started (a pass over few particles costs its launch latency, not its work). */
typedef enum mmc_schedule {
  MMC_SCHEDULE_AUTO = 0, MMC_SCHEDULE_FUSED = 1, MMC_SCHEDULE_EVENT = 2, MMC_SCHEDULE_EVENT_ONLY = 3
} mmc_schedule;

/* Optional knobs; zero-initialise for defaults. */
typedef struct mmc_run_options {
  uint32_t struct_size;             /* sizeof(mmc_run_options) */
  int32_t device;                   /* used by the host layer's Driver (which device its world is created on); the
                                     * mmc_world entry points run on the device the world was created on */
  int32_t tracking;                 /* mmc_tracking */
  int32_t rng_mode;                 /* mmc_rng_mode */
  uint32_t secondary_capacity;      /* per-history fission queue slots (default 32) */
  uint32_t pending_capacity;        /* per-history distinct scored bins (default 32) */
  uint32_t blocks_per_sm;           /* 0 = library default */
  uint32_t schedule;                /* mmc_schedule; 0 = library default */
  void* stream;                     /* cudaStream_t; NULL = the library's own stream.  For the duration of an event-split
                                     * run the library puts an L2 access-policy window over the world's tables on this
                                     * stream (and, once per world, raises the device's persisting-L2 set-aside to hold
                                     * them, cudaLimitPersistingL2CacheSize: a device-wide setting); a caller's stream
                                     * has its window cleared again before the call returns.  MMC_L2_PERSIST=0 in the
                                     * environment turns both off. */
  uint32_t event_slots;             /* MMC_SCHEDULE_EVENT: histories in flight at once (0 = library default) */
  uint32_t profile;                 /* MMC_SCHEDULE_EVENT: 1 = time every kernel with CUDA events (mmc_world_last_kernel_ms) */
} mmc_run_options;

int mmc_abi_version(void);
/* Copies the message of the calling thread's last failed call into buf. */
size_t mmc_last_error(char* buf, size_t cap);
/* Number of visible CUDA devices (0 when there is none / no driver). */
int mmc_device_count(void);

/* Validates and uploads the tables (replaces nothing in the reference: this is
 * the one-off flattening of `const World`, World.cpp:20-24).  After the upload (and after every mmc_world_update)
 * the device expands the POD factors of each thermal-scattering partition into a dense table of the sums
 * BetaPartition::Evaluate / AlphaPartition::Evaluate make (ThermalScattering.cpp:199-204,241-246; same order, same
 * values), up to 512 MB of device memory in all; the environment variable MMC_TSL_DENSE_MB changes that budget
 * (0: never expand, every reconstruction sums its rank-R terms on the fly).  Results do not depend on it. */
int mmc_world_create(const mmc_world_desc* desc, int device, mmc_world** out);
void mmc_world_destroy(mmc_world* world);
/* Uploads new table VALUES into an existing world (same shapes: the device image must have the same size), through
 * the world's pinned staging buffer; no allocation.  For callers that refresh cross sections between batches. */
int mmc_world_update(mmc_world* world, const mmc_world_desc* desc);

/* Bytes of flattened tables resident on the device for this world (what mmc_world_create copied host -> device). */
uint64_t mmc_world_bytes(const mmc_world* world);

/* Kernels launched by the last mmc_fixed_source_run[_device] on this world (1 for the fused schedule). */
uint64_t mmc_world_last_launches(const mmc_world* world);
/* With mmc_run_options.profile = 1: device time (ms, CUDA events on the run's stream) the last event-split run spent in
 * its flight kernels and in its S(a,b) kernels. */
void mmc_world_last_kernel_ms(const mmc_world* world, double* flight_ms, double* tsl_ms);
double mmc_world_last_boundary_ms(const mmc_world* world);  /* the boundary kernels (crossings, history ends, births) */

/* Total number of bins of estimator e = cosine.n_bins * energy.n_bins
 * (ParticleBins::size, Bins.cpp:192-194). */
uint64_t mmc_estimator_size(const mmc_estimator_desc* e);

/* Replaces FixedSource::Solve()/StartWorker() (FixedSource.cpp:22-77) for
 * histories [first_history, first_history + n_histories): history i is sampled
 * from `std::minstd_rand{seed0 + i}` (FixedSource.cpp:61, Source.cpp:143-154).
 * scores / square_scores are the concatenation, in estimator order, of
 * Scorable::scores / square_scores (Scorable.hpp:63-65): they are ADDED to, as
 * Scorable::operator+= does, so ranks or batches can be chained.  HOST buffers. */
int mmc_fixed_source_run(
    const mmc_world* world, const mmc_source_desc* source, const mmc_estimator_desc* estimators,
    int32_t n_estimators, uint64_t seed0, uint64_t first_history, uint64_t n_histories,
    const mmc_run_options* options, double* scores, double* square_scores, mmc_counters* counters);

/* Differential-operator sensitivities (Perturbation.cpp:68-84, Sensitivity.cpp:44-58, Particle.cpp:46-53): sensitivity
 * i is the CurrentTotalCrossSectionSensitivity of estimators[estimator] to the total cross section of nuclide
 * `nuclide` (TotalCrossSectionPerturbation).  It shares its estimator's bins; every particle accumulates
 * 1 / GetCollisionProbabilityDensity - distance per Stream in a material that holds the nuclide, and scores
 * that indirect effect times the estimator's score. */
typedef struct mmc_sensitivity_desc {
  int32_t estimator;                /* index into the estimators array */
  int32_t nuclide;                  /* index of the perturbed nuclide in the world */
} mmc_sensitivity_desc;

/* mmc_fixed_source_run plus sensitivities: sens_scores / sens_square_scores are the concatenation, in sensitivity
 * order, of Scorable::scores / square_scores of each Sensitivity (mmc_estimator_size of its estimator each), ADDED to
 * like scores.  Real-valued sums of fp64 atomics: equal to the reference up to summation order. */
int mmc_fixed_source_run_sensitivities(
    const mmc_world* world, const mmc_source_desc* source, const mmc_estimator_desc* estimators,
    int32_t n_estimators, const mmc_sensitivity_desc* sensitivities, int32_t n_sensitivities, uint64_t seed0,
    uint64_t first_history, uint64_t n_histories, const mmc_run_options* options, double* scores,
    double* square_scores, double* sens_scores, double* sens_square_scores, mmc_counters* counters);

/* Same, asynchronous on options->stream, with DEVICE tally buffers of exact
 * integer counts (the `current` score is 0 or 1 per event, so Sigma s and
 * Sigma (per-history sum)^2 are integers): d_scores / d_square_scores are
 * uint64[total bins], d_counters is mmc_counters.  Used by multi-GPU callers that
 * all-reduce the integers with NCCL before converting to double. */
int mmc_fixed_source_run_device(
    const mmc_world* world, const mmc_source_desc* source, const mmc_estimator_desc* estimators,
    int32_t n_estimators, uint64_t seed0, uint64_t first_history, uint64_t n_histories,
    const mmc_run_options* options, uint64_t* d_scores, uint64_t* d_square_scores, mmc_counters* d_counters);

/* ---- k-eigenvalue generations ----------------------------------------------------
 * The reference's KEigenvalue::Solve is a stub (KEigenvalue.cpp:36-62: the worker launch, the fission-bank merge and
 * the bank swap are commented out or TODO), so the power iteration is DEFINED here (DESIGN.md "k-eigenvalue") on top
 * of the pieces the reference does have: the initial bank of KEigenvalue.cpp:29-33 and the fission physics of
 * Multigroup::Fission / ContinuousFission::Interact.  All buffers are DEVICE buffers; calls are asynchronous on
 * options->stream. */
typedef struct mmc_site {           /* one banked particle, 64 bytes */
  double position[3];
  double direction[3];
  uint64_t energy_bits;             /* group (multigroup) or the bits of the energy in MeV (continuous) */
  uint32_t seed;                    /* std::minstd_rand{seed} of the particle (Particle.cpp:96-100) */
  int32_t reserved;
} mmc_site;

/* d_bank[i] = Source::Sample(seed0 + first_index + i) kept as a site, i in [0, n)
 * (KEigenvalue.cpp:29-33 uses seeds 1..batchsize: seed0 = 1). */
int mmc_source_bank_sample(const mmc_world* world, const mmc_source_desc* source, uint64_t seed0, uint64_t first_index,
                           uint64_t n, const mmc_run_options* options, mmc_site* d_bank);

/* One generation: transports the n_in particles of d_bank_in (each is one history for the tallies; scored only when
 * `score` is non-zero -- inactive cycles do not tally).  Fission secondaries are not followed: they are written to
 * d_bank_out in (parent index, creation ordinal) order, *d_n_out receives their number (k of the generation is
 * n_out / n_in over all ranks).  MMC_ERR_CAPACITY is reported through d_counters->n_capacity_overflow when more than
 * bank_capacity sites are produced. */
int mmc_generation_run(const mmc_world* world, const mmc_site* d_bank_in, uint64_t n_in,
                       const mmc_estimator_desc* estimators, int32_t n_estimators, int32_t score,
                       const mmc_run_options* options, mmc_site* d_bank_out, uint64_t bank_capacity, uint64_t* d_n_out,
                       uint64_t* d_scores, uint64_t* d_square_scores, mmc_counters* d_counters,
                       uint64_t* d_k_collision);
/* d_k_collision (device uint64, may be NULL; ADDED to): the generation's collision ("implicit fission") estimator of k,
 * the one KEigenvalue.hpp:33 lists as to be reimplemented -- every real collision scores nu-bar Sigma_f / Sigma_t of
 * the material at the pre-collision energy.  Fixed point, MMC_K_COLLISION_ONE per unit score, so that the sum is an
 * exact integer whatever the scheduling and the number of GPUs: k_collision = sum / MMC_K_COLLISION_ONE / n_in. */
#define MMC_K_COLLISION_ONE 268435456.0 /* 2^28 */

/* Source bank of the next generation.  The m_total fission sites of all ranks, in global order, are resampled to
 * n_total sources with a deterministic comb: source i <- site floor(i * m_total / n_total); copies of one site get
 * seeds site.seed + copy ordinal.  This call writes sources [first_out, first_out + n_out) into d_bank_next from
 * d_slice, which holds global sites [slice_first, slice_first + slice_n).  *d_errors (uint64) counts sources whose
 * site was outside the slice. */
int mmc_bank_resample(const mmc_world* world, const mmc_site* d_slice, uint64_t slice_first, uint64_t slice_n,
                      uint64_t m_total, uint64_t n_total, uint64_t first_out, uint64_t n_out,
                      const mmc_run_options* options, mmc_site* d_bank_next, uint64_t* d_errors);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink / NVSwitch -----------------------------------
 * The reference's workers are threads of one process (FixedSource.cpp:22-36) and its KEigenvalue::Solve is a stub
 * (KEigenvalue.cpp:36-62), so nothing here replaces reference code: these are the exchange steps the sharded path
 * needs.  Fixed source: rank r transports histories [r*N/P, (r+1)*N/P) (mmc_fixed_source_run_device) and the integer
 * tallies are summed once (mmc_tally_allreduce).  K-eigenvalue: rank r transports source indices [r*N/P, (r+1)*N/P) of
 * every generation (mmc_generation_run), then mmc_bank_exchange + mmc_bank_resample build its next source bank.
 * NCCL is loaded at run time (libnccl.so.2); a communicator of one rank needs no NCCL at all. */
typedef struct mmc_comm mmc_comm;
#define MMC_COMM_ID_BYTES 128
/* ncclGetUniqueId: call on one rank, hand the bytes to every rank (file, MPI, torch.distributed store, ...). */
int mmc_comm_unique_id(void* id, size_t cap);
/* ncclCommInitRank on `device` (-1: the current one); collective over the nranks processes. */
int mmc_comm_create(int nranks, int rank, const void* id, int device, mmc_comm** out);
void mmc_comm_destroy(mmc_comm* comm);
int mmc_comm_rank(const mmc_comm* comm);
int mmc_comm_size(const mmc_comm* comm);
int mmc_nccl_version(void);            /* ncclGetVersion of the library in use, 0 when NCCL is not available */
/* Device time (ms, CUDA events on the calls' stream) spent in mmc_bank_exchange so far on this rank. */
double mmc_comm_exchange_ms(mmc_comm* comm);
/* In-place sum over the ranks of n_words DEVICE uint64 words (integer tallies, counters, k sums): ncclAllReduce,
 * asynchronous on `stream` (NULL: the communicator's own stream).  The multi-GPU form of
 * `solver_estimator_set += worker_estimator_set.get()` (FixedSource.cpp:31-33); exact, so the result is the same for
 * any number of ranks. */
int mmc_tally_allreduce(mmc_comm* comm, uint64_t* d_words, size_t n_words, void* stream);
/* Host-only arithmetic of the exchange (no device, no NCCL): counts[r] = sites banked by rank r this generation, in
 * global order rank 0's sites first.  need_first / need_count: the global site range the next n_total/P sources of
 * `rank` are drawn from (source i <- site floor(i * M / n_total)).  send[2*p], send[2*p+1]: start (in this rank's
 * ordered bank) and count of the sites rank p needs from this rank; recv[2*p], recv[2*p+1]: offset in the rank's
 * slice and count of the sites it needs from rank p.  Entries with p == rank describe the local copy. */
int mmc_exchange_plan(const uint64_t* counts, int nranks, int rank, uint64_t n_total, uint64_t* need_first,
                      uint64_t* need_count, uint64_t* send, uint64_t* recv);
/* The exchange step of one k-eigenvalue generation.  d_bank_local: this rank's ordered fission bank
 * (mmc_generation_run's d_bank_out), *d_n_local its size (device).  local_status: non-zero when this rank saw an
 * error (lost particle, capacity overflow, ...).  All-gathers {size, status} of every rank into counts[nranks] /
 * statuses[nranks] (HOST, statuses may be NULL) -- one stream synchronisation -- and, unless some status is set or
 * the bank is empty (then every rank returns MMC_OK with *slice_n = 0 and the caller decides, alike on every rank),
 * moves the parts of global sites [*slice_first, *slice_first + *slice_n) that live on other ranks into d_slice with
 * grouped ncclSend / ncclRecv, asynchronously on `stream`.  Follow with mmc_bank_resample(d_slice, *slice_first, ...). */
int mmc_bank_exchange(mmc_comm* comm, const mmc_site* d_bank_local, const uint64_t* d_n_local, uint64_t local_status,
                      uint64_t n_total, mmc_site* d_slice, uint64_t slice_capacity, uint64_t* counts, uint64_t* statuses,
                      uint64_t* slice_first, uint64_t* slice_n, void* stream);
/* The stream (cudaStream_t) a world's calls run on when mmc_run_options.stream is NULL. */
void* mmc_world_stream(const mmc_world* world);

/* Device-buffer helpers so that a host without the CUDA toolkit (the C++ host of this repo is compiled by g++) can
 * drive the device-buffer entry points.  All synchronise with the world's stream. */
int mmc_device_alloc(const mmc_world* world, size_t bytes, void** d_ptr);   /* zero-initialised */
void mmc_device_free(const mmc_world* world, void* d_ptr);
int mmc_device_zero(const mmc_world* world, void* d_ptr, size_t bytes);
int mmc_device_read(const mmc_world* world, void* host_dst, const void* d_src, size_t bytes);
int mmc_device_write(const mmc_world* world, void* d_dst, const void* host_src, size_t bytes);

/* Parity hook (the reference has none; oracle/ref_harness.cpp obtains the same
 * records through an extra Estimator): per-event records of histories
 * [first_history, first_history+n_histories), grouped by history, particles in
 * the reference's bank order (FixedSource.cpp:63-72).  n_records receives the
 * number written; MMC_ERR_CAPACITY if cap was too small. */
int mmc_trace_histories(
    const mmc_world* world, const mmc_source_desc* source, uint64_t seed0, uint64_t first_history,
    uint64_t n_histories, const mmc_run_options* options, mmc_event_record* records, size_t cap,
    size_t* n_records);

/* Diagnostic (no reference counterpart): evaluates on the device the libm
 * functions the transport kernels use -- the bit-exact restatements of glibc's
 * log (fn 0: out0), sincos (fn 1: out0 = sin, out1 = cos), sin (fn 2), cos
 * (fn 3), the converged pair used for Direction(d, mu, phi) (fn 5: out0 = sin(x), out1 = cos(x) as two separate
 * libm calls return them) -- and the libstdc++ generate_canonical stream (fn 4: x[i] is a seed,
 * out0[i] the first canonical double of std::minstd_rand{seed}; fn 16 + 4: the same of the MMC_RNG_COUNTER build, the
 * first canonical double of the Philox-2x32-10 stream whose id is x[i]).  HOST buffers. */
int mmc_test_device_math(int fn, const double* x, double* out0, double* out1, size_t n);

/* Diagnostic (the reference's test_World.cpp / test_Cell.cpp / test_CSGSurface.cpp cases run through it): for each
 * of n query points, World::FindCellContaining(position) (-1 where the reference throws) and, from that cell,
 * Cell::NearestSurface(position, direction): surface index and distance (inf when no surface is ahead).
 * positions / directions are [n][3]; directions are used as given (not normalised).  HOST buffers. */
int mmc_test_geometry(const mmc_world* world, size_t n, const double* positions, const double* directions,
                      int32_t* cell, int32_t* surface, double* distance);

/* ---- host layer ---------------------------------------------------------------
 * The C++17 host (minimc_b200/host: XML deck -> World/Material/Nuclide/Source/
 * EstimatorSet -> Driver) behind C entry points, for callers that are not C++
 * (the Python tests, bench.py).  A C++ caller uses minimc_b200/host/minimc.hpp
 * directly: `minimc::Driver::Create(path)->Solve()` is the reference's own
 * call sequence (minimc.cpp:16-21). */
typedef struct mmc_driver mmc_driver;

/* Driver::Create(path) (Driver.cpp:19-35).  Construction errors carry the
 * reference's messages (mmc_last_error). */
int mmc_driver_create(const char* xml_path, mmc_driver** out);
int mmc_driver_create_from_string(const char* xml_text, mmc_driver** out);
void mmc_driver_destroy(mmc_driver* driver);
/* Knobs of the GPU path; tracking in `options` is ignored (the deck decides). */
int mmc_driver_set_options(mmc_driver* driver, const mmc_run_options* options);
/* This process transports histories [rank*N/P, (rank+1)*N/P) of the batch. */
int mmc_driver_set_shard(mmc_driver* driver, int32_t rank, int32_t world_size);
/* Driver::Solve() (Driver.hpp:30): runs on the GPU, keeps the EstimatorSet. */
int mmc_driver_solve(mmc_driver* driver);
uint64_t mmc_driver_batchsize(const mmc_driver* driver);
uint64_t mmc_driver_total_bins(const mmc_driver* driver);
/* Concatenated Scorable::scores / square_scores of the last Solve(). */
int mmc_driver_scores(const mmc_driver* driver, double* scores, double* square_scores, uint64_t n);
/* EstimatorSet::operator+= with another rank's scores (Estimator.cpp:215-231). */
int mmc_driver_add_scores(mmc_driver* driver, const double* scores, const double* square_scores, uint64_t n);
int mmc_driver_counters(const mmc_driver* driver, mmc_counters* counters);
/* The text runminimc writes to <input>.out: "<batchsize>\n" + EstimatorSet::to_string()
 * (minimc.cpp:20-21).  Returns the full length; copies at most cap-1 bytes. */
size_t mmc_driver_output(const mmc_driver* driver, char* buf, size_t cap);
/* The flattened World as JSON with C99 hex floats (test hook). */
size_t mmc_driver_world_json(const mmc_driver* driver, char* buf, size_t cap);
/* Device-resident form of Solve() for fixed-source decks (mmc_fixed_source_run_device with the driver's world,
 * source and estimators): histories [first_history, first_history + n_histories) of seed `Driver::seed`, integer
 * tallies accumulated into DEVICE buffers, asynchronous on `stream` (a cudaStream_t; NULL = the library's stream). */
int mmc_driver_run_device(mmc_driver* driver, uint64_t first_history, uint64_t n_histories, uint64_t* d_scores,
                          uint64_t* d_square_scores, mmc_counters* d_counters, void* stream);
/* Drops the device copy of the World, so that the next Solve() uploads the tables again; and its size in bytes. */
void mmc_driver_release_device(mmc_driver* driver);
/* Flattens the driver's World again and uploads it into the existing device world (mmc_world_update); creates the
 * device world when there is none. */
int mmc_driver_refresh_device(mmc_driver* driver);
uint64_t mmc_driver_table_bytes(mmc_driver* driver);
/* Kernels launched by the driver's last Solve() / run_device on this rank (mmc_world_last_launches). */
uint64_t mmc_driver_last_launches(mmc_driver* driver);
void mmc_driver_last_kernel_ms(mmc_driver* driver, double* flight_ms, double* tsl_ms);
double mmc_driver_last_boundary_ms(mmc_driver* driver);
/* Parity hook: mmc_trace_histories for histories [first, first + n) of a fixed-source deck. */
int mmc_driver_trace(mmc_driver* driver, uint64_t first_history, uint64_t n_histories, mmc_event_record* records,
                     size_t cap, size_t* n_records);
/* k-eigenvalue results of the last Solve(): mean and standard deviation of the
 * mean over active cycles, and k of every cycle (inactive first). */
int mmc_driver_keff(const mmc_driver* driver, double* k_mean, double* k_std, double* k_cycle, size_t cap, size_t* n_cycles);

/* The collision ("implicit fission", KEigenvalue.hpp:33) estimator of the last Solve(): mean and standard deviation of
 * the mean over active cycles, the value of every cycle, and the device time (ms) the bank exchanges took. */
int mmc_driver_k_collision(const mmc_driver* driver, double* k_mean, double* k_std, double* k_cycle, size_t cap,
                           size_t* n_cycles, double* exchange_ms);
/* Host time (s) the last k-eigenvalue Solve() spent in its inactive and in its active cycles; every cycle ends with a
 * device synchronisation, the active time includes the final all-reduce and the read-back of the tallies. */
int mmc_driver_cycle_seconds(const mmc_driver* driver, double* inactive_seconds, double* active_seconds);
/* Multi-process runs, one process per GPU: Solve() of a fixed-source deck transports this rank's share of the
 * histories and all-reduces the integer tallies (every rank then holds the batch's EstimatorSet); Solve() of a
 * k-eigenvalue deck shards every generation and exchanges the fission bank (mmc_bank_exchange).  The communicator is
 * not owned by the driver.  _from_environment: MMC_WORLD_SIZE, MMC_RANK, MMC_DEVICE, MMC_COMM_ID_FILE (rank 0
 * writes the NCCL unique id to that path, the others wait for it) -- a launcher without MPI or torch. */
int mmc_driver_set_comm(mmc_driver* driver, mmc_comm* comm);
int mmc_driver_init_comm_from_environment(mmc_driver* driver);

#ifdef __cplusplus
}
#endif
#endif /* MINIMC_B200_H */
