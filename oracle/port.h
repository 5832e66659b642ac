/* ORACLE TEST INFRASTRUCTURE -- not product code.
 *
 * C interface of oracle/port.cpp, the CPU restatement of the reference's
 * history transport loop (agtumulak/minimc @ ed536a2).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this; the
 * product (libminimc_b200.so) never does.  Pinned against the reference's own
 * code (oracle/_ref/ref_harness) by tests/test_oracle.py and the golden files
 * under tests/golden/.  Self-contained on purpose: it does not include the
 * product header. */
#ifndef MINIMC_ORACLE_PORT_H
#define MINIMC_ORACLE_PORT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_world {
  int32_t n_surfaces;
  const int32_t* surface_type;   /* 0 sphere, 1 planex, 2 cylinderx */
  const double* surface_param;   /* [n][4] */
  int32_t n_cells;
  const int32_t* cell_material;  /* -1 void */
  const int32_t* cell_surface_begin;
  const int32_t* cell_surface_index;
  const int32_t* cell_surface_sense;
  int32_t n_materials;
  const double* material_aden;
  const int32_t* material_nuclide_begin;
  const int32_t* material_nuclide_index;
  const double* material_nuclide_afrac;
  int32_t n_nuclides;
  int32_t n_groups;
  const uint32_t* mg_reaction_mask; /* 1 capture, 2 scatter, 4 fission */
  const double* mg_total;
  const double* mg_capture;
  const double* mg_scatter;
  const double* mg_fission;
  const double* mg_nubar;
  const double* mg_scatter_probs;   /* [n][G_in][G_out] */
  const double* mg_chi;
} orc_world;

typedef struct orc_source {
  double position[3];
  int32_t direction_kind;        /* 0 constant, 1 isotropic, 2 isotropic-flux */
  double direction[3];
  uint64_t group;
} orc_source;

typedef struct orc_bins {
  int32_t kind;                  /* 0 none, 1 linspace, 2 logspace, 3 boundaries */
  uint64_t n_bins;
  double lower, upper, width, base;
  const double* boundaries;
} orc_bins;

typedef struct orc_estimator {
  int32_t surface;
  int32_t has_cosine_direction;
  double cosine_direction[3];
  orc_bins cosine, energy;
} orc_estimator;

typedef struct orc_counters {
  uint64_t n_histories, n_births, n_events, n_collisions, n_crossings, n_virtual, n_scores, n_secondaries,
      n_lost, n_physics_errors;
} orc_counters;

typedef struct orc_record {
  uint64_t history;
  uint32_t particle;
  int32_t event;
  uint64_t group;
  int32_t cell, surface;
  double position[3], direction[3];
  uint64_t rng_state;
} orc_record;

/* Histories [first, first+n) of a fixed-source problem, seed of history i =
 * seed0 + i; tracking 0 surface, 1 cell delta; `threads` worker threads.
 * scores / square_scores (concatenated estimators) are added to.  Returns 0,
 * or 1 when a particle was lost / an assert(false) branch was reached. */
int orc_fixed_source_run(
    const orc_world* world, const orc_source* source, const orc_estimator* estimators, int32_t n_estimators,
    uint64_t seed0, uint64_t first, uint64_t n, int32_t tracking, int32_t threads, double* scores,
    double* square_scores, orc_counters* counters);

/* Per-event records in the reference's bank order; returns the number of
 * records the histories need (only min(that, cap) are written). */
size_t orc_trace(
    const orc_world* world, const orc_source* source, uint64_t seed0, uint64_t first, uint64_t n, int32_t tracking,
    orc_record* records, size_t cap);

/* K-eigenvalue power iteration as defined in DESIGN.md "k-eigenvalue" (the reference has none: parity with the
 * reference is unpinned here).  k_cycle / bank_sizes / k_collision_cycle (the collision estimator of k: nu Sigma_f / Sigma_t at every real collision,
 * summed in 2^-28 fixed point, per source) have inactive + active entries; tallies cover active cycles. */
int orc_keigenvalue_run(
    const orc_world* world, const orc_source* source, const orc_estimator* estimators, int32_t n_estimators,
    uint64_t batchsize, uint64_t inactive, uint64_t active, int32_t tracking, double* scores, double* square_scores,
    double* k_cycle, uint64_t* bank_sizes, orc_counters* counters, double* k_collision_cycle /* may be NULL */);

/* n canonical doubles + engine states from std::minstd_rand{seed}, restated. */
void orc_rng_canonical(uint64_t seed, size_t n, double* u, uint64_t* state);

#ifdef __cplusplus
}
#endif
#endif
