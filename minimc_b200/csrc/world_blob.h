// Flat device image of the reference's `const World` (World.hpp) plus the
// source and estimator descriptors of one run.  One contiguous, 16-byte aligned
// byte blob: a fixed header of counts and byte offsets followed by SoA arrays.
// The transport kernels stage the whole blob into shared memory when it fits
// (every multigroup deck of the reference is < 2 KB) and otherwise read it
// through the read-only path from global memory / L2.
#pragma once

#include <cstdint>

namespace mmc {

constexpr int kMaxEstimators = 16;
constexpr int kMaxPerturbations = 4;   // distinct perturbed nuclides per run
constexpr int kMaxSensitivities = 16;  // (estimator, perturbation) pairs per run

// Sensitivity of a `current` estimator to the total cross section of one nuclide
// (CurrentTotalCrossSectionSensitivity, Sensitivity.cpp:44-58): shares the estimator's bins.
struct SensitivitySpec {
  int32_t estimator;     // index into RunSpec::estimators
  int32_t perturbation;  // index into RunSpec::perturbed_nuclide
  uint64_t offset;       // first bin of this sensitivity in the concatenated sensitivity tallies
};

struct BinsSpec {
  int32_t kind;        // mmc_bins_kind
  uint32_t n_bins;     // including the two unbounded end bins
  double lower, upper, width, base;
  uint32_t off_boundaries;  // byte offset of double[n_bins-1] (BOUNDARIES)
  uint32_t pad;
};

struct EstimatorSpec {
  int32_t surface;
  int32_t has_direction;
  double direction[3];   // normalised (Direction ctor, Point.cpp:80-83)
  BinsSpec cosine, energy;
  uint64_t stride;       // ParticleBins::strides[0] = energy.n_bins (Bins.hpp:150-153)
  uint64_t offset;       // first bin of this estimator in the concatenated score arrays
};

struct SourceSpec {
  double position[3];
  double direction[3];   // normalised
  int32_t direction_kind;
  int32_t pad;
  uint64_t group;
  double energy;
};

// One entry of a cell's surface list with the surface itself inline (48 bytes, 16-byte aligned): Cell::Contains and
// Cell::NearestSurface walk a cell's entries with ONE load each instead of the index -> type -> parameters chain.
struct SurfaceRecord {
  double prm[4];         // CSGSurface parameters (sphere: centre + radius; planex: x; cylinderx: radius)
  int32_t type;          // mmc_surface_type
  int32_t index_sense;   // (surface index << 1) | sense
  int32_t pad[2];
};

struct WorldHeader {
  uint32_t total_bytes;
  int32_t n_surfaces, n_cells, n_materials, n_nuclides, n_groups;
  // 1 when every S(a,b) partition of the world has its dense reconstruction table (TslPartition::off_dense)
  int32_t tsl_all_dense;
  // byte offsets from the start of the blob
  uint32_t off_surface_type;    // int32[n_surfaces]
  uint32_t off_surface_param;   // double[n_surfaces][4]
  uint32_t off_cell_material;   // int32[n_cells]
  uint32_t off_cell_surf_begin; // int32[n_cells+1]
  uint32_t off_cell_surf;       // int32[nnz]  (index << 1) | sense
  uint32_t off_cell_surf_rec;   // SurfaceRecord[nnz]: the same list with each surface's type and parameters inline
  uint32_t off_cell_mask;       // uint64[n_cells][2] {mask, want}: Cell::Contains(p) <=> (inside(p) & mask) == want, where
                                // bit s of inside(p) is surfaces[s].Contains(p); 0 when the world has more than 64 surfaces
  uint32_t off_cell_field_kind; // int32[n_cells]
  uint32_t off_cell_field_param;// double[n_cells][6]
  uint32_t off_mat_aden;        // double[n_materials]
  uint32_t off_mat_nuc_begin;   // int32[n_materials+1]
  uint32_t off_mat_nuc_index;   // int32[nnz]
  uint32_t off_mat_nuc_afrac;   // double[nnz]
  uint32_t off_mg_mask;         // uint32[n_nuclides]
  uint32_t off_mg_total;        // double[n_nuclides][G]
  uint32_t off_mg_capture, off_mg_scatter, off_mg_fission, off_mg_nubar;
  uint32_t off_mg_scatter_probs, off_mg_chi;  // double[n_nuclides][G][G]
  uint32_t off_ce_nuclides;     // CeNuclide[n_nuclides]; 0 for multigroup worlds
  // every partition's S[r] * CDF_modes[cdf][r] array (TslPartition::off_scaled_cdf_modes), contiguous: the S(a,b)
  // kernel of the event-split schedule stages this arena into shared memory when it fits
  uint32_t off_sc_arena, sc_arena_bytes;
  // device-only tail of the image, [total_bytes, total_bytes + dense_bytes): the dense reconstruction tables, filled
  // on the device after every upload (kernels.cu expand_dense_kernel); never part of the host image
  uint32_t dense_bytes;
  // int32[n_cells]: index of the cell's constant temperature among the world's evaluated temperatures
  // (TslPartition::off_eval), -1 for void cells, cells with a linear temperature field and cells past kMaxEvalT
  uint32_t off_cell_eval_slot;
  int32_t n_eval_T;
  // 1 when every collision the S(a,b) kernel can be handed samples with ce::tsl_sample_direct: every material cell has
  // an evaluated temperature (off_cell_eval_slot >= 0) and every table is TslTable::direct -- the host's copy of the
  // header is corrected after the device has checked the rows (capi.cu confirm_direct); selects the S(a,b) kernel kind
  // that holds nothing but that sampler
  uint32_t tsl_all_direct;
};

// distinct constant cell temperatures a world keeps evaluated S(a,b) tables for (BASELINE configs: 1, 13, 1, 0)
constexpr int kMaxEvalT = 64;

// ---- continuous-energy tables (blob offsets; see include/minimc_b200.h mmc_ce_desc)
constexpr int kMaxCeReactions = 4;  // per nuclide: capture, scatter, fission (+1 spare)

struct Table1D {
  uint32_t n;
  uint32_t off_x, off_y;  // double[n] each
  uint32_t off_hint;      // SearchHint of x, 0 = none
};

// Bucket index over a sorted, non-negative array x[n] that narrows std::upper_bound to a few elements: the bucket
// of a value is its IEEE-754 bit pattern shifted right (exponent + leading mantissa bits: monotone in the value, a
// piecewise-linear log2), cum[k] = number of elements in buckets below k.  upper_bound(v) lies in
// [cum[k(v)], cum[k(v)+1]]; searching that range gives the same index as searching the whole array because the array
// is sorted (checked when the world is built; arrays that are not sorted get no hint and the plain search).
struct SearchHint {
  int64_t first_bucket;  // bucket number (bits >> shift) of the first positive element, minus 1
  uint32_t shift;
  uint32_t n_buckets;    // bucket 0 = everything below the first positive element
  // uint2 range[n_buckets] follows: {cum[k], cum[k + 1]}, one 8-byte load per search (the header is one 16-byte load)
};

// ThermalScattering::BetaPartition / AlphaPartition
struct TslPartition {
  uint32_t n_cdf, n_grid, n_T, rank;
  uint32_t off_cdf;        // double[n_cdf]
  uint32_t off_T;          // double[n_T]
  uint32_t off_scaled_cdf_modes;  // double[n_cdf][rank]: S[r] * CDF_modes[cdf][r] (the reference's first product)
  uint32_t off_cdf_hint;   // SearchHint of cdf, 0 = none
  uint32_t off_modes;      // double[n_grid][n_T][rank]
  uint32_t grid_begin;     // index of this partition's first grid point in the concatenated Es / betas
  uint32_t off_T_hint;     // SearchHint of T, 0 = none
  // double[n_grid][n_T][n_cdf]: the rank-R sums  sum_r S[r] * CDF_modes[cdf][r] * modes[grid][T][r]  of Evaluate
  // (ThermalScattering.cpp:199-204,241-246) for EVERY (grid point, CDF node, temperature node), each summed in the
  // reference's order -- the POD factors expanded once per upload, so that a reconstruction is two loads (T_lo, T_hi)
  // and the reference's interpolation in T.  The CDF node is the fastest axis: the last four probes of find_cdf's
  // bisections, and both ends of a try's bracket, fall into one 128-byte line per temperature row.
  // 0 = not expanded (over the budget): sum on the fly.
  uint32_t off_dense;
  // double[n_eval_T][n_grid][n_cdf]: Evaluate(cdf, grid, T_s) itself -- the two rank-R sums AND the interpolation in
  // temperature (ThermalScattering.cpp:206-214,248-255, same operations in the same order) -- for every distinct
  // constant cell temperature T_s of the world (WorldHeader::off_cell_eval_slot).  Evaluate is a pure function of
  // (cdf, grid, T) and T is a per-cell constant in BASELINE configs[1], [2], [3]: a reconstruction there is ONE load
  // and find_cdf's comparator probes one contiguous row of n_cdf doubles.  0 = none.
  uint32_t off_eval;
  // 1: every evaluated row is non-decreasing in the CDF node (checked on the device after each evaluation,
  // kernels.cu check_rows_sorted_kernel).  Kept as the condition for TslTable::direct.  (r02: find_cdf's search,
  // ce::find_cdf_bisect, now makes libstdc++'s own probes, so its index no longer depends on this; the earlier
  // search by rounds of independent loads did.  The flag stays conservative rather than being re-argued.)
  uint32_t eval_sorted;
  // double2[n_cdf + 1]: {F_lo, F_hi} of the CDF bracket whose upper index is i -- {cdf[i - 1] or 0, cdf[i] or 1}, the
  // pair ThermalScattering.cpp:313-320,438-447 reads after every search, as one 16-byte load
  uint32_t off_cdf_pairs;
  // uint8[kCdfLut]: lut[k] = std::upper_bound(cdf, k / kCdfLut) -- the sampled CDF value F is uniform, so bucket
  // floor(F * kCdfLut) almost always names the bracket outright: one byte load, then the pair above confirms it (or the
  // index moves up a node).  0 = none (more than 255 nodes, or an axis that is not sorted inside [0, 1]).
  uint32_t off_cdf_lut;
};
constexpr uint32_t kCdfLut = 1024;

// GetTotal's temperature bracket at one evaluated temperature
struct TslEvalBracket {
  uint32_t lo, hi;  // T_lo_i, T_hi_i
  double r_T;       // below_T_min ? 1 : above_T_max ? 0 : (T - T_lo) / (T_hi - T_lo)
};

// ThermalScattering
struct TslTable {
  Table1D majorant;
  uint32_t n_E, n_T, rank, off_E_hint;
  uint32_t off_E, off_T, off_xs_SE, off_Es_hint, off_xs_T;  // off_xs_SE: double[n_E][rank] = S[r] * scatter_xs_E[E][r]
  uint32_t n_beta_partitions, n_alpha_partitions;
  uint32_t off_beta_partitions, off_alpha_partitions;  // TslPartition[]
  uint32_t n_Es, off_Es;        // concatenated incident energies of the beta partitions (ThermalScattering::Es)
  uint32_t n_betas, off_betas;  // concatenated betas of the alpha partitions (ThermalScattering::betas)
  uint32_t off_betas_hint;      // SearchHints of E / Es / betas: off_E_hint, off_Es_hint, off_betas_hint
  uint32_t off_T_hint;          // SearchHint of T
  uint32_t off_xs_dense;        // double[n_E][n_T]: EvaluateInelastic at every (E, T) node, expanded like off_dense; 0 = none
  // 1: every beta and alpha partition has evaluated tables (off_eval) whose rows are all sorted (eval_sorted): a
  // collision in a cell with an evaluated temperature samples with ce::tsl_sample_direct.  Set in the image when every
  // partition got its evaluated table, cleared on the device when a partition's rows turn out unsorted.
  uint32_t direct;
  // TslEvalBracket[WorldHeader::n_eval_T]: the temperature bracket of GetTotal (ThermalScattering.cpp:126-135,152-155)
  // at each evaluated cell temperature -- T_lo_i, T_hi_i and r_T are functions of T alone; 0 = none
  uint32_t off_eval_bracket;
  double beta_cutoff, alpha_cutoff, awr, cutoff_energy;
};


// one ContinuousReaction
struct CeReaction {
  int32_t kind;      // MMC_REACTION_*
  uint32_t off_tsl;  // TslTable, 0 = none
  Table1D xs;
  double temperature;
  Table1D nubar;     // n == 0: absent
};

// one Continuous
struct CeNuclide {
  double awr;
  Table1D total;
  double total_temperature;
  int32_t n_reactions;
  uint32_t pad;
  CeReaction reactions[kMaxCeReactions];
};

// Per-run constant block (passed by value as a kernel parameter).
struct RunSpec {
  SourceSpec source;
  EstimatorSpec estimators[kMaxEstimators];
  int32_t n_estimators;
  int32_t tracking;
  uint64_t seed0;
  uint64_t first_history;
  uint64_t n_histories;
  uint64_t total_bins;
  uint32_t secondary_capacity;  // power of two
  uint32_t pending_capacity;
  uint32_t chunk;               // histories claimed per warp refill
  uint32_t world_bytes;         // size of the world blob (multiple of 16)
  uint32_t world_in_smem;       // 1: kernels stage the blob into shared memory
  uint32_t continuous_energy;   // 1: the world's nuclides are Continuous (n_groups == 0)
  // differential-operator sensitivities (Perturbation.cpp, Sensitivity.cpp); n_sensitivities == 0: none
  int32_t n_perturbations;
  int32_t n_sensitivities;
  int32_t perturbed_nuclide[kMaxPerturbations];  // TotalCrossSectionPerturbation::nuclide
  SensitivitySpec sensitivities[kMaxSensitivities];
  uint32_t sens_pending_capacity;  // per-history (bin, sum) entries of the sensitivity proxies
  uint32_t pad_sens;
};

// One expansion of POD factors into a dense table (device-side, after every upload of the image):
// out[(g * n_cdf + c) * n_T + t] = sum_r a[c * rank + r] * m[(g * n_T + t) * rank + r], r ascending from 0.0.
struct DenseJob {
  uint32_t off_a;    // double[n_cdf][rank]: S[r] * CDF_modes[cdf][r]  (or S[r] * scatter_xs_E[E][r])
  uint32_t off_m;    // double[n_grid][n_T][rank]
  uint32_t off_out;  // double[n_grid][n_cdf][n_T] or, cdf_fastest, double[n_grid][n_T][n_cdf]; in the device-only tail
  uint32_t n_grid, n_cdf, n_T, rank;
  uint32_t cdf_fastest;
};

// Evaluate(cdf, grid, T_s) of one partition for every evaluated temperature (device-side, after every upload):
// out[(s * n_grid + g) * n_cdf + c] = v_lo + (v_hi - v_lo) / dT[s] * tT[s] with
// v_k = sum_r a[c * rank + r] * m[(g * n_T + t_k[s]) * rank + r], r ascending from 0.0.
struct EvalJob {
  uint32_t off_a;    // double[n_cdf][rank]: S[r] * CDF_modes[cdf][r]
  uint32_t off_m;    // double[n_grid][n_T][rank]
  uint32_t off_out;  // double[n_slots][n_grid][n_cdf], in the device-only tail
  uint32_t n_grid, n_cdf, n_T, rank, n_slots;
  uint8_t t_lo[kMaxEvalT], t_hi[kMaxEvalT];  // T_lo_i, T_hi_i of the partition's temperature axis for each T_s
  double dT[kMaxEvalT], tT[kMaxEvalT];       // T_hi - T_lo, T_s - T_lo
  uint32_t off_sorted_flag;                  // blob offset of the partition's TslPartition::eval_sorted
  uint32_t off_direct_flag;                  // blob offset of its table's TslTable::direct
};

// One banked particle (secondary or k-eigenvalue site): 64 bytes.
struct BankSite {
  double position[3];
  double direction[3];
  uint64_t energy_bits;  // group (multigroup) or the bits of the energy in MeV
  uint32_t seed;         // argument of std::minstd_rand{seed} (Particle.cpp:96-100)
  int32_t surface;       // unused for secondaries (Particle::current_surface starts null)
};
static_assert(sizeof(BankSite) == 64, "BankSite must be 64 bytes");

}  // namespace mmc
