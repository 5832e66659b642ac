// Driver / FixedSource / KEigenvalue: the reference's batch-level seam
// (Driver.hpp:23-30).  Solve() hands the whole batch to the GPU through the C
// ABI; there is no host transport loop.
#include <cmath>
#include <stdexcept>

#include "minimc.hpp"

namespace minimc {

namespace {

[[noreturn]] void ThrowLastError(const char* where, int status) {
  char buf[512];
  mmc_last_error(buf, sizeof(buf));
  throw DeviceError(status, std::string(where) + ": " + buf);
}

mmc_tracking ParseTracking(const xml::Node& root, const World& world) {
  // TransportMethod::Create, TransportMethod.cpp:24-46
  const xml::Node* general = root.child("general");
  const xml::Node* tracking = general ? general->child("tracking") : nullptr;
  const std::string tracking_type = tracking ? tracking->text() : "";
  if (tracking_type.empty() || tracking_type == "surface") {
    if (!world.HasConstantTemperature())
      throw std::runtime_error("Surface tracking with continuous global temperature not allowed");
    return MMC_TRACK_SURFACE;
  }
  if (tracking_type == "cell delta") return MMC_TRACK_CELL_DELTA;
  throw std::runtime_error(tracking->path() + ": unknown tracking type \"" + tracking_type + "\"");
}

const xml::Node& GeneralChild(const xml::Node& root, const char* name) {
  const xml::Node* general = root.child("general");
  const xml::Node* node = general ? general->child(name) : nullptr;
  if (!node) throw std::runtime_error(std::string("/minimc/general: \"") + name + "\" node not found");
  return *node;
}

uint64_t ParseSeed(const xml::Node& root) {
  // Driver.cpp:41-44: default seed 1
  const xml::Node* general = root.child("general");
  const xml::Node* seed = general ? general->child("seed") : nullptr;
  return seed ? std::stoull(seed->text()) : 1;
}

const xml::Node& ProblemNode(const xml::Node& root, const char* name) {
  const xml::Node* problemtype = root.child("problemtype");
  const xml::Node* node = problemtype ? problemtype->child(name) : nullptr;
  if (!node) throw std::runtime_error(std::string("/minimc/problemtype: \"") + name + "\" node not found");
  return *node;
}

}  // namespace

// Owns the device copy of a World.
class DeviceWorld {
public:
  DeviceWorld(const World& world, int device) : flat{world} {
    if (const int status = mmc_world_create(&flat.desc(), device, &handle)) ThrowLastError("mmc_world_create", status);
  }
  ~DeviceWorld() { mmc_world_destroy(handle); }
  DeviceWorld(const DeviceWorld&) = delete;
  DeviceWorld& operator=(const DeviceWorld&) = delete;
  FlatWorld flat;
  mmc_world* handle = nullptr;
};

namespace {
std::unique_ptr<Driver> CreateFromRoot(const xml::Node& root) {
  // Driver::Create, Driver.cpp:19-35
  const xml::Node* problemtype = root.child("problemtype");
  const std::string problem_type = problemtype && problemtype->first_child() ? problemtype->first_child()->name() : "";
  if (problem_type == "fixedsource") return std::make_unique<FixedSource>(root);
  if (problem_type == "keigenvalue") return std::make_unique<KEigenvalue>(root);
  throw std::runtime_error("/minimc/problemtype: expected a fixedsource or keigenvalue node");
}
}  // namespace

std::unique_ptr<Driver> Driver::Create(const std::string& xml_filepath) {
  const xml::Document doc = xml::Document::FromFile(xml_filepath);
  return CreateFromRoot(doc.root());
}

std::unique_ptr<Driver> Driver::CreateFromString(const std::string& xml_text) {
  const xml::Document doc = xml::Document::FromString(xml_text);
  return CreateFromRoot(doc.root());
}

Driver::Driver(const xml::Node& root)
    : world{root}, perturbations{root.child("perturbations"), world},
      batchsize{std::stoull(GeneralChild(root, "histories").text())}, seed{ParseSeed(root)},
      init_estimator_set{root.child("estimators"), world, perturbations, static_cast<Real>(batchsize)},
      threads{std::stoul(GeneralChild(root, "threads").text())}, tracking{ParseTracking(root, world)} {
  run_options.struct_size = sizeof(mmc_run_options);
  run_options.device = -1;
  run_options.tracking = tracking;
  run_options.rng_mode = MMC_RNG_MINSTD_COMPAT;
}

Driver::~Driver() noexcept {}

std::shared_ptr<DeviceWorld> Driver::device_world() {
  if (!device_world_) device_world_ = std::make_shared<DeviceWorld>(world, run_options.device);
  return device_world_;
}

const mmc_world* Driver::device_world_handle() { return device_world()->handle; }

void Driver::RefreshDevice() {
  if (!device_world_) {
    device_world();
    return;
  }
  device_world_->flat = FlatWorld{world};
  if (const int status = mmc_world_update(device_world_->handle, &device_world_->flat.desc())) ThrowLastError("mmc_world_update", status);
}

// ---------------------------------------------------------------- FixedSource
FixedSource::FixedSource(const xml::Node& root) : Driver{root}, source{ProblemNode(root, "fixedsource")} {}

EstimatorSet FixedSource::Solve() {
  // FixedSource::Solve (FixedSource.cpp:22-36) with the worker pool replaced
  // by one call: histories [first, first + count) of this rank.
  EstimatorSet result = init_estimator_set;
  const std::vector<mmc_estimator_desc> estimators = FlattenEstimators(result);
  const uint64_t first = static_cast<uint64_t>(rank) * batchsize / static_cast<uint64_t>(world_size);
  const uint64_t last = static_cast<uint64_t>(rank + 1) * batchsize / static_cast<uint64_t>(world_size);
  std::vector<double> scores(result.total_bins(), 0.0), square_scores(result.total_bins(), 0.0);
  run_options.tracking = tracking;
  // sensitivities (Estimator::sensitivities): (estimator, perturbed nuclide) pairs, tallies concatenated in this order
  std::vector<mmc_sensitivity_desc> sensitivities;
  size_t sensitivity_bins = 0;
  for (size_t e = 0; e < result.estimators.size(); e++)
    for (const Sensitivity& s : result.estimators[e].sensitivities) {
      sensitivities.push_back(mmc_sensitivity_desc{static_cast<int32_t>(e), static_cast<int32_t>(s.nuclide)});
      sensitivity_bins += s.scores.size();
    }
  std::vector<double> sens_scores(sensitivity_bins, 0.0), sens_square_scores(sensitivity_bins, 0.0);
  const int status = mmc_fixed_source_run_sensitivities(
      device_world()->handle, &source.desc, estimators.data(), static_cast<int32_t>(estimators.size()),
      sensitivities.data(), static_cast<int32_t>(sensitivities.size()), seed, first, last - first, &run_options,
      scores.data(), square_scores.data(), sens_scores.data(), sens_square_scores.data(), &counters);
  if (status != MMC_OK) ThrowLastError("mmc_fixed_source_run", status);
  size_t offset = 0, sens_offset = 0;
  for (Estimator& e : result.estimators) {
    for (size_t i = 0; i < e.scores.size(); i++) {
      e.scores[i] += scores[offset + i];
      e.square_scores[i] += square_scores[offset + i];
    }
    offset += e.scores.size();
    for (Sensitivity& s : e.sensitivities) {
      for (size_t i = 0; i < s.scores.size(); i++) {
        s.scores[i] += sens_scores[sens_offset + i];
        s.square_scores[i] += sens_square_scores[sens_offset + i];
      }
      sens_offset += s.scores.size();
    }
  }
  return result;
}

void FixedSource::RunDevice(
    uint64_t first, uint64_t count, uint64_t* d_scores, uint64_t* d_square_scores, mmc_counters* d_counters, void* stream) {
  if (init_estimator_set.total_sensitivities())
    throw std::runtime_error("sensitivities are real-valued tallies: use Solve(), not the integer device-buffer run");
  if (device_estimators_.empty() && !init_estimator_set.estimators.empty()) device_estimators_ = FlattenEstimators(init_estimator_set);
  mmc_run_options options = run_options;
  options.tracking = tracking;
  options.stream = stream;
  const int status = mmc_fixed_source_run_device(
      device_world()->handle, &source.desc, device_estimators_.data(), static_cast<int32_t>(device_estimators_.size()), seed,
      first, count, &options, d_scores, d_square_scores, d_counters);
  if (status != MMC_OK) ThrowLastError("mmc_fixed_source_run_device", status);
}

std::vector<mmc_event_record> FixedSource::Trace(uint64_t first, uint64_t count, size_t cap) {
  std::vector<mmc_event_record> records(cap);
  size_t n = 0;
  run_options.tracking = tracking;
  const int status = mmc_trace_histories(
      device_world()->handle, &source.desc, seed, first, count, &run_options, records.data(), cap, &n);
  if (status != MMC_OK) ThrowLastError("mmc_trace_histories", status);
  records.resize(n);
  return records;
}

}  // namespace minimc
