// C++17 host of minimc_b200: the reference's object model and driver API
// (agtumulak/minimc @ ed536a2), re-implemented so that `Driver::Solve()` runs
// the history loop on the GPU through the C ABI of include/minimc_b200.h
// instead of `threads` x std::async(StartWorker) (FixedSource.cpp:22-36).
//
// The host owns parsing, the object model, drivers and output; it holds NO
// transport physics: cross-section lookup, tracking, collisions and tallies
// exist only as CUDA kernels (minimc_b200/csrc).  Classes keep the reference's
// names and construction semantics (including its error messages) so that a
// user of the reference finds the same surface:
//
//   reference                         here
//   XMLDocument (Xerces + pugixml)    xml::Document (xml.hpp; no schema validation)
//   HDF5DataSet<D> (pandas HDF5)      TableFile (flat "MMCTAB1" file, same axes + row-major values)
//   World/Cell/CSGSurface/ScalarField same names, indices instead of shared_ptr
//   Material/Nuclide/Multigroup/Continuous/ContinuousReaction/ThermalScattering   same names, tables only
//   Source, Bins, ParticleBins, Estimator, EstimatorSet                           same names
//   Driver::Create / FixedSource / KEigenvalue                                    same names
#pragma once

#include <cstdint>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/minimc_b200.h"
#include "xml.hpp"

namespace minimc {

using Real = double;

namespace constants {
// Constants.hpp:9-38
constexpr Real room_temperature = 293.6;
}  // namespace constants

// A failed call into the C ABI: carries the mmc_status (lost particle, capacity, physics, CUDA ...).
class DeviceError : public std::runtime_error {
public:
  DeviceError(int status, const std::string& message) : std::runtime_error(message), status(status) {}
  const int status;
};

struct Point {
  Real x = 0, y = 0, z = 0;
};

// ---------------------------------------------------------------- table files
// Replaces HDF5DataSet<D> (HDF5DataSet.hpp:37-174).  libhdf5 is not available
// and every *.hdf5 of the reference is a git-lfs pointer, so tables are read
// from the flat MMCTAB1 format (DESIGN.md "table file format"): D sorted axes
// + row-major values, exactly the content of the pandas "fixed" layout.
class TableFile {
public:
  // Throws "File not found: <path>" (HDF5DataSet.hpp:100-103) and
  // "<path>: Expected <D> dimensions, but got <n>" (:113-117).
  TableFile(const std::string& path, size_t expected_dimensions);
  const std::vector<double>& GetAxis(size_t level) const { return axes.at(level); }
  std::vector<std::vector<double>> axes;
  std::vector<double> values;
};

// Sorted, de-duplicated (key, value) points of a ContinuousMap (ContinuousMap.hpp:21-56).
struct PointwiseTable {
  std::vector<double> x, y;
  static PointwiseTable FromFile(const std::string& path);  // HDF5DataSet<1>::ToContinuousMap
};

// ------------------------------------------------------------------- geometry
class CSGSurface {
public:
  // CSGSurface::Create (CSGSurface.cpp:15-44)
  static CSGSurface Create(const xml::Node& root, const std::string& name);
  std::string name;
  mmc_surface_type type = MMC_SURF_SPHERE;
  Real param[4] = {0, 0, 0, 0};  // sphere: center xyz, radius; planex: x; cylinderx: radius
};

// ScalarField / ConstantField / LinearField (ScalarField.cpp:12-58)
class ScalarField {
public:
  static ScalarField Create(const xml::Node& scalar_field_node);
  static ScalarField Constant(Real c);
  bool IsConstant() const noexcept;
  mmc_field_kind kind = MMC_FIELD_CONSTANT;
  Real c = 0;
  Point g;
  Real b = 0;
  Real upper_bound = 0, lower_bound = 0;
};

// ----------------------------------------------------------------- data model
// Multigroup (Multigroup.cpp:23-47,77-243): per-group tables, G x G matrices
// stored [incoming][outgoing] with each incoming column normalised.
class Multigroup {
public:
  explicit Multigroup(const xml::Node& particle_node);
  uint64_t max_group = 0;
  uint32_t reaction_mask = 0;
  std::vector<Real> total, capture, scatter, fission, nubar;
  std::vector<Real> scatter_probs, chi;  // [G][G]
};

// ThermalScattering tables (ThermalScattering.cpp:24-104)
class ThermalScattering {
public:
  struct Partition {
    Partition(const xml::Node& partition_node, const char* modes_attribute);
    TableFile CDF_modes, singular_values, grid_T_modes;
  };
  explicit ThermalScattering(const xml::Node& tsl_node);
  PointwiseTable majorant;
  TableFile scatter_xs_T, scatter_xs_S, scatter_xs_E;
  std::vector<Partition> beta_partitions, alpha_partitions;
  Real beta_cutoff, alpha_cutoff, awr;
};

// ContinuousReaction and subclasses (ContinuousReaction.cpp:24-265)
class ContinuousReaction {
public:
  static ContinuousReaction Create(const xml::Node& reaction_node);
  uint32_t kind = 0;  // MMC_REACTION_*
  PointwiseTable xs;
  Real temperature = 0;
  std::optional<ThermalScattering> tsl;
  std::optional<PointwiseTable> nubar;
};

// Continuous (Continuous.cpp:22-90)
class Continuous {
public:
  explicit Continuous(const xml::Node& particle_node);
  std::vector<ContinuousReaction> reactions;  // XML document order
  PointwiseTable total;
  Real total_temperature = 0;
  Real awr = 0;
};

// Nuclide + Interaction::Create (Nuclide.cpp:14-16, Interaction.cpp:16-45); neutrons only.
class Nuclide {
public:
  explicit Nuclide(const xml::Node& nuclide_node);
  std::string name;
  std::optional<Multigroup> multigroup;
  std::optional<Continuous> continuous;
};

// Material (Material.cpp:17-39,66-93)
class Material {
public:
  static const xml::Node& FindNode(const xml::Node& root, const std::string& material_name);
  Material(const xml::Node& root, const std::string& name, const std::vector<Nuclide>& all_nuclides);
  std::string name;
  // (index into World::nuclides, normalised atom fraction), iterated in World
  // creation order (the reference iterates a pointer-keyed std::map, quirk Q1)
  std::vector<std::pair<size_t, Real>> afracs;
  Real number_density = 0;
};

// Cell (Cell.cpp:17-25,57-113)
class Cell {
public:
  Cell(const xml::Node& cell_node, const std::vector<CSGSurface>& all_surfaces,
       const std::vector<Material>& all_materials, const ScalarField& global_temperature);
  std::string name;
  std::vector<std::pair<size_t, bool>> surface_senses;  // (surface index, Contains() must be) in creation order
  int material = -1;                                     // index into World::materials, -1 = void
  ScalarField temperature;
};

// World (World.cpp:20-182)
class World {
public:
  explicit World(const xml::Node& root);
  bool HasConstantTemperature() const noexcept { return temperature.IsConstant(); }
  size_t FindSurfaceIndexByName(const std::string& name) const;  // throws like World::FindSurfaceByName
  size_t FindNuclideIndexByName(const std::string& name) const;  // throws like World::FindNuclideByName
  bool IsMultigroup() const { return !nuclides.empty() ? nuclides.front().multigroup.has_value() : multigroup_groups > 0; }
  std::vector<CSGSurface> surfaces;
  std::vector<Nuclide> nuclides;
  std::vector<Material> materials;
  ScalarField temperature;
  std::vector<Cell> cells;
  uint64_t multigroup_groups = 0;  // <multigroup groups="G">; 0 for continuous decks
};

// --------------------------------------------------------------------- source
// Source (Source.cpp:131-154): constant position / energy / type, constant,
// isotropic or isotropic-flux direction.
class Source {
public:
  explicit Source(const xml::Node& source_node);
  mmc_source_desc desc{};
};

// ------------------------------------------------------------------ estimators
// Bins::Create and the concrete bins (Bins.cpp:19-170)
class Bins {
public:
  static Bins Create(const xml::Node* bins_node);
  size_t size() const noexcept;
  std::string to_string() const noexcept;
  mmc_bins_kind kind = MMC_BINS_NONE;
  size_t n_bins = 1;
  Real lower = 0, upper = 0, width = 0, base = 10;
  std::vector<Real> boundaries;
};

// ParticleBins (Bins.cpp:174-212)
class ParticleBins {
public:
  explicit ParticleBins(const xml::Node* bins_node);
  size_t size() const noexcept { return cosine.size() * energy.size(); }
  std::string to_string() const noexcept;
  std::optional<Point> direction;  // as given; normalised where it is used (Direction ctor)
  Bins cosine, energy;
};

// Scorable + Estimator + CurrentEstimator (Scorable.cpp, Estimator.cpp:21-151)
// TotalCrossSectionPerturbation (Perturbation.cpp:44-84): the only perturbation the schema allows
struct Perturbation {
  std::string name;
  size_t nuclide = 0;  // index into World::nuclides
};

// PerturbationSet (Perturbation.cpp:88-128)
class PerturbationSet {
public:
  PerturbationSet() = default;
  PerturbationSet(const xml::Node* perturbations_node, const World& world);
  const Perturbation& FindPerturbationByName(const std::string& name) const;
  std::vector<Perturbation> perturbations;
};

// CurrentTotalCrossSectionSensitivity (Sensitivity.cpp:44-58): a Scorable over its estimator's bins
struct Sensitivity {
  std::string name;  // estimator name + "::" + perturbation name (Scorable.cpp:22-29)
  size_t nuclide = 0;
  std::vector<Real> scores, square_scores;
};

class Estimator {
public:
  Estimator(const xml::Node& estimator_node, const World& world, const PerturbationSet& perturbations);
  std::string to_string(Real total_weight) const noexcept;
  Estimator& operator+=(const Estimator& other) noexcept;
  std::string name;
  ParticleBins bins;
  size_t surface = 0;
  std::vector<Real> scores, square_scores;
  std::vector<Sensitivity> sensitivities;
};

// EstimatorSet (Estimator.cpp:155-231)
class EstimatorSet {
public:
  EstimatorSet() = default;
  EstimatorSet(const xml::Node* estimators_node, const World& world, const PerturbationSet& perturbations, Real total_weight);
  size_t total_sensitivities() const noexcept;
  const Estimator& FindEstimatorByName(const std::string& name) const;
  std::string to_string() const noexcept;
  EstimatorSet& operator+=(const EstimatorSet& other);
  size_t total_bins() const noexcept;
  std::vector<Estimator> estimators;
  Real total_weight = 0;
};

// -------------------------------------------------------------------- drivers
// Device tables of a World: owns the flattened arrays and the mmc_world handle.
class DeviceWorld;

struct KResult {
  std::vector<Real> k_cycle;  // one per cycle (inactive then active)
  Real k_mean = 0, k_std = 0; // over active cycles
  std::vector<uint64_t> bank_sizes;
  // collision ("implicit fission", KEigenvalue.hpp:33) estimator: sum over real collisions of nu Sigma_f / Sigma_t, per source
  std::vector<Real> k_collision_cycle;
  Real k_collision_mean = 0, k_collision_std = 0;
  double exchange_ms = 0;     // device time of the bank exchanges (multi-rank runs)
  // host time of the inactive and of the active cycles (each cycle ends with a device synchronisation; the active
  // time includes the final all-reduce and the read-back of the tallies)
  double inactive_seconds = 0, active_seconds = 0;
};

// Driver (Driver.cpp:19-52)
class Driver {
public:
  // Parses the deck and builds the driver named by <problemtype>.
  static std::unique_ptr<Driver> Create(const std::string& xml_filepath);
  static std::unique_ptr<Driver> CreateFromString(const std::string& xml_text);
  explicit Driver(const xml::Node& root);
  virtual ~Driver() noexcept;
  // Runs the problem on the GPU (device = run_options.device) and returns the tallies.
  virtual EstimatorSet Solve() = 0;

  const World world;
  const PerturbationSet perturbations;  // Driver.hpp:35-37
  const uint64_t batchsize;
  const uint64_t seed;
  const EstimatorSet init_estimator_set;
  const size_t threads;  // parsed for schema compatibility; the GPU path does not use host worker threads
  const mmc_tracking tracking;
  // knobs for the GPU path (no reference counterpart)
  mmc_run_options run_options{};
  // rank / world_size sharding of histories [0, batchsize): rank r owns
  // [r*N/P, (r+1)*N/P); the caller reduces EstimatorSets with operator+=
  int rank = 0, world_size = 1;
  // multi-process runs (one process per GPU): the NCCL communicator of this rank, not owned unless created by
  // InitCommFromEnvironment.  SetComm also sets rank / world_size.
  mmc_comm* comm = nullptr;
  void SetComm(mmc_comm* c);
  // MMC_WORLD_SIZE > 1 in the environment: MMC_RANK, MMC_DEVICE (default: rank) and MMC_COMM_ID_FILE, a path every rank
  // can reach -- rank 0 writes the NCCL unique id there, the others wait for it.  A launcher without MPI or torch.
  void InitCommFromEnvironment();
  mmc_counters counters{};
  // Drops the device copy of the World (the next Solve() uploads it again); bytes it holds.
  void ReleaseDevice() { device_world_.reset(); }
  // Flattens the World again and uploads it into the existing device world (no allocation).
  void RefreshDevice();
  uint64_t DeviceTableBytes() { return mmc_world_bytes(device_world_handle()); }

  const mmc_world* device_world_handle();

protected:
  std::shared_ptr<DeviceWorld> device_world();

private:
  std::shared_ptr<DeviceWorld> device_world_;
  mmc_comm* owned_comm_ = nullptr;
};

// FixedSource (FixedSource.cpp:19-77)
class FixedSource : public Driver {
public:
  explicit FixedSource(const xml::Node& root);
  EstimatorSet Solve() override;
  // Device-resident Solve(): integer tallies of histories [first, first + count) accumulated into device buffers,
  // asynchronous on `stream`.
  void RunDevice(uint64_t first, uint64_t count, uint64_t* d_scores, uint64_t* d_square_scores, mmc_counters* d_counters,
                 void* stream);
  // Parity hook (mmc_trace_histories): per-event records of histories [first, first + count).
  std::vector<mmc_event_record> Trace(uint64_t first, uint64_t count, size_t cap = size_t{1} << 20);

private:
  const Source source;
  std::vector<mmc_estimator_desc> device_estimators_;  // points into init_estimator_set
};

// KEigenvalue (KEigenvalue.cpp:17-86 is a stub in the reference, SURVEY.md F1;
// the power iteration is defined in DESIGN.md "k-eigenvalue").
class KEigenvalue : public Driver {
public:
  explicit KEigenvalue(const xml::Node& root);
  EstimatorSet Solve() override;
  const uint64_t last_inactive, last_active;
  const KResult& result() const { return result_; }
  // fission-bank capacity = bank_capacity_factor * batchsize sites (k of a generation must stay below it)
  double bank_capacity_factor = 4.0;

private:
  const Source source;
  KResult result_;
};

// Flattening of a World into the C ABI descriptor; keeps the arrays alive.
class FlatWorld {
public:
  explicit FlatWorld(const World& world);
  const mmc_world_desc& desc() const { return desc_; }
  std::string to_json() const;  // hex-float dump, compared against oracle/ref_harness `dump` in tests

private:
  struct Arrays;
  std::shared_ptr<Arrays> a_;
  mmc_world_desc desc_{};
};

std::vector<mmc_estimator_desc> FlattenEstimators(const EstimatorSet& set);

}  // namespace minimc
