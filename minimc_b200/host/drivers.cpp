// Driver / FixedSource / KEigenvalue: the reference's batch-level seam
// (Driver.hpp:23-30).  Solve() hands the whole batch to the GPU through the C
// ABI; there is no host transport loop.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <thread>

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

Driver::~Driver() noexcept {
  if (owned_comm_) mmc_comm_destroy(owned_comm_);
}

void Driver::SetComm(mmc_comm* c) {
  comm = c;
  rank = mmc_comm_rank(c);
  world_size = mmc_comm_size(c);
}

void Driver::InitCommFromEnvironment() {
  const char* ws = std::getenv("MMC_WORLD_SIZE");
  const int P = ws ? std::atoi(ws) : 1;
  if (P <= 1) return;
  const char* rk = std::getenv("MMC_RANK");
  const char* file = std::getenv("MMC_COMM_ID_FILE");
  if (!rk || !file) throw std::runtime_error("MMC_WORLD_SIZE > 1 needs MMC_RANK and MMC_COMM_ID_FILE");
  const int R = std::atoi(rk);
  const char* dev = std::getenv("MMC_DEVICE");
  const int device = dev ? std::atoi(dev) : R;
  run_options.device = device;
  unsigned char id[MMC_COMM_ID_BYTES];
  const std::string path = file, tmp = path + ".tmp";
  if (R == 0) {
    if (const int status = mmc_comm_unique_id(id, sizeof(id))) ThrowLastError("mmc_comm_unique_id", status);
    std::FILE* f = std::fopen(tmp.c_str(), "wb");
    if (!f || std::fwrite(id, 1, sizeof(id), f) != sizeof(id)) throw std::runtime_error("cannot write " + tmp);
    std::fclose(f);
    if (std::rename(tmp.c_str(), path.c_str()) != 0) throw std::runtime_error("cannot rename " + tmp);
  } else {
    // the file appears atomically (rename) once rank 0 has written all of it
    for (int waited_ms = 0;; waited_ms += 20) {
      if (std::FILE* f = std::fopen(path.c_str(), "rb")) {
        const size_t n = std::fread(id, 1, sizeof(id), f);
        std::fclose(f);
        if (n == sizeof(id)) break;
      }
      if (waited_ms > 120000) throw std::runtime_error("timed out waiting for " + path);
      std::this_thread::sleep_for(std::chrono::milliseconds(20));
    }
  }
  mmc_comm* c = nullptr;
  if (const int status = mmc_comm_create(P, R, id, device, &c)) ThrowLastError("mmc_comm_create", status);
  owned_comm_ = c;
  SetComm(c);
}

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
  if (comm && world_size > 1) {
    // One process per GPU: this rank's histories with DEVICE tallies, then ONE all-reduce of the packed integer words
    // [scores | squares | counters] -- the multi-GPU form of `solver_estimator_set += worker_estimator_set.get()`
    // (FixedSource.cpp:31-33).  Every rank returns the whole batch's EstimatorSet.
    if (!sensitivities.empty())
      throw std::runtime_error("sensitivities are real-valued tallies: reduce the ranks' EstimatorSets with operator+= instead");
    const mmc_world* w = device_world()->handle;
    const size_t bins = result.total_bins();
    constexpr size_t kCounterWords = sizeof(mmc_counters) / sizeof(uint64_t);
    const size_t words = 2 * bins + kCounterWords;
    void* d_ptr = nullptr;
    if (const int st = mmc_device_alloc(w, words * sizeof(uint64_t), &d_ptr)) ThrowLastError("mmc_device_alloc", st);
    uint64_t* d_words = static_cast<uint64_t*>(d_ptr);
    mmc_run_options options = run_options;
    options.stream = run_options.stream ? run_options.stream : mmc_world_stream(w);
    int st = mmc_fixed_source_run_device(w, &source.desc, estimators.data(), static_cast<int32_t>(estimators.size()), seed, first,
                                         last - first, &options, d_words, d_words + bins,
                                         reinterpret_cast<mmc_counters*>(d_words + 2 * bins));
    if (st == MMC_OK) st = mmc_tally_allreduce(comm, d_words, words, options.stream);
    std::vector<uint64_t> h_words(words);
    if (st == MMC_OK) st = mmc_device_read(w, h_words.data(), d_words, words * sizeof(uint64_t));
    mmc_device_free(w, d_ptr);
    if (st != MMC_OK) ThrowLastError("FixedSource::Solve (multi-rank)", st);
    std::memcpy(&counters, h_words.data() + 2 * bins, sizeof(counters));
    if (counters.n_lost) throw DeviceError(MMC_ERR_LOST_PARTICLE, "FixedSource::Solve: particle(s) outside every cell");
    if (counters.n_physics_errors) throw DeviceError(MMC_ERR_PHYSICS, "FixedSource::Solve: a branch the reference asserts unreachable was reached");
    if (counters.n_capacity_overflow) throw DeviceError(MMC_ERR_CAPACITY, "FixedSource::Solve: per-history capacity overflow");
    size_t at = 0;
    for (Estimator& e : result.estimators) {
      for (size_t i = 0; i < e.scores.size(); i++) {
        e.scores[i] += static_cast<Real>(h_words[at + i]);
        e.square_scores[i] += static_cast<Real>(h_words[bins + at + i]);
      }
      at += e.scores.size();
    }
    return result;
  }
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
