// KEigenvalue: power iteration on the GPU.
//
// The reference's KEigenvalue::Solve is a stub (KEigenvalue.cpp:36-62: the worker
// launch is commented out, the fission-bank merge and the bank swap are TODOs,
// active cycles never run and there is no k estimator).  What the
// reference does define is kept: the constructor's initial bank
// (KEigenvalue.cpp:29-33: `batchsize` particles from Source::Sample(s), s = 1..batchsize)
// and the fission physics.  The rest is defined in DESIGN.md "k-eigenvalue":
//
//   cycle c = 0 .. inactive + active - 1:
//     transport the N source sites (each one history); fission secondaries are
//     banked, not followed, in (source index, creation ordinal) order
//     k_c = M_c / N          (M_c sites banked, all weights 1)
//     next source bank: comb resampling, source i <- site floor(i * M_c / N),
//     copies of a site get seeds seed + copy ordinal
//     estimators score in active cycles only
//   k = mean of k_c over active cycles, sigma = sample standard deviation of the mean
//   EstimatorSet.total_weight = N * active
//
// With `rank` / `world_size` set, this process owns source indices
// [rank*N/P, (rank+1)*N/P) of every generation; the bank exchange between
// processes is done by the caller (minimc_b200/distributed.py over NCCL), this
// single-process Solve() requires world_size == 1.
#include <cmath>
#include <stdexcept>

#include "minimc.hpp"

namespace minimc {

namespace {
const xml::Node& KNode(const xml::Node& root) {
  const xml::Node* problemtype = root.child("problemtype");
  const xml::Node* node = problemtype ? problemtype->child("keigenvalue") : nullptr;
  if (!node) throw std::runtime_error("/minimc/problemtype: \"keigenvalue\" node not found");
  return *node;
}
const xml::Node& InitialSource(const xml::Node& root) {
  const xml::Node* node = KNode(root).child("initialsource");
  if (!node) throw std::runtime_error("/minimc/problemtype/keigenvalue: \"initialsource\" node not found");
  return *node;
}

void Check(int status, const char* where) {
  if (status == MMC_OK) return;
  char buf[512];
  mmc_last_error(buf, sizeof(buf));
  throw DeviceError(status, std::string(where) + ": " + buf);
}

// RAII device buffer through the C ABI helpers
class DeviceBuffer {
public:
  DeviceBuffer(const mmc_world* world, size_t bytes) : world_{world} { Check(mmc_device_alloc(world, bytes, &ptr_), "mmc_device_alloc"); }
  ~DeviceBuffer() { mmc_device_free(world_, ptr_); }
  DeviceBuffer(const DeviceBuffer&) = delete;
  DeviceBuffer& operator=(const DeviceBuffer&) = delete;
  template <typename T> T* as() const { return static_cast<T*>(ptr_); }

private:
  const mmc_world* world_;
  void* ptr_ = nullptr;
};
}  // namespace

KEigenvalue::KEigenvalue(const xml::Node& root)
    : Driver{root}, last_inactive{KNode(root).attribute_ull("inactive")},
      last_active{last_inactive + KNode(root).attribute_ull("active")}, source{InitialSource(root)} {}

EstimatorSet KEigenvalue::Solve() {
  if (init_estimator_set.total_sensitivities())
    throw std::runtime_error("/minimc/estimators: sensitivities are implemented for fixed-source problems only");
  if (world_size != 1)
    throw std::runtime_error("KEigenvalue::Solve: multi-process runs are driven by minimc_b200.distributed (NCCL bank exchange)");
  const mmc_world* w = device_world_handle();
  const uint64_t N = batchsize;
  const uint64_t active = last_active - last_inactive;
  EstimatorSet result = init_estimator_set;
  result.total_weight = static_cast<Real>(N) * static_cast<Real>(active ? active : 1);
  const std::vector<mmc_estimator_desc> estimators = FlattenEstimators(result);
  const size_t bins = result.total_bins();
  // a generation can bank more than N sites: room for k up to bank_capacity_factor
  const uint64_t capacity = static_cast<uint64_t>(bank_capacity_factor * static_cast<double>(N)) + 1024;
  DeviceBuffer bank_source{w, N * sizeof(mmc_site)}, bank_fission{w, capacity * sizeof(mmc_site)};
  DeviceBuffer d_n_out{w, sizeof(uint64_t)}, d_errors{w, sizeof(uint64_t)}, d_counters{w, sizeof(mmc_counters)};
  DeviceBuffer d_scores{w, std::max<size_t>(bins, 1) * sizeof(uint64_t)}, d_squares{w, std::max<size_t>(bins, 1) * sizeof(uint64_t)};
  run_options.tracking = tracking;
  // KEigenvalue.cpp:29-33: source.Sample(s) for s = 1 .. batchsize
  Check(mmc_source_bank_sample(w, &source.desc, 1, 0, N, &run_options, bank_source.as<mmc_site>()), "mmc_source_bank_sample");
  result_ = KResult{};
  for (uint64_t cycle = 0; cycle < last_active; cycle++) {
    const bool score = cycle >= last_inactive;
    Check(mmc_generation_run(
              w, bank_source.as<mmc_site>(), N, estimators.data(), static_cast<int32_t>(estimators.size()), score ? 1 : 0,
              &run_options, bank_fission.as<mmc_site>(), capacity, d_n_out.as<uint64_t>(), d_scores.as<uint64_t>(),
              d_squares.as<uint64_t>(), d_counters.as<mmc_counters>()),
          "mmc_generation_run");
    uint64_t M = 0;
    Check(mmc_device_read(w, &M, d_n_out.as<uint64_t>(), sizeof(M)), "mmc_device_read");
    Check(mmc_device_read(w, &counters, d_counters.as<mmc_counters>(), sizeof(counters)), "mmc_device_read");
    if (counters.n_lost) throw DeviceError(MMC_ERR_LOST_PARTICLE, "KEigenvalue::Solve: particle(s) outside every cell");
    if (counters.n_physics_errors) throw DeviceError(MMC_ERR_PHYSICS, "KEigenvalue::Solve: a branch the reference asserts unreachable was reached");
    if (counters.n_capacity_overflow || M > capacity)
      throw DeviceError(MMC_ERR_CAPACITY, "KEigenvalue::Solve: fission bank overflow (raise bank_capacity_factor / secondary_capacity)");
    result_.k_cycle.push_back(static_cast<Real>(M) / static_cast<Real>(N));
    result_.bank_sizes.push_back(M);
    if (M == 0) throw DeviceError(MMC_ERR_PHYSICS, "KEigenvalue::Solve: the fission chain died out (empty fission bank)");
    Check(mmc_bank_resample(w, bank_fission.as<mmc_site>(), 0, M, M, N, 0, N, &run_options, bank_source.as<mmc_site>(),
                            d_errors.as<uint64_t>()),
          "mmc_bank_resample");
  }
  uint64_t errors = 0;
  Check(mmc_device_read(w, &errors, d_errors.as<uint64_t>(), sizeof(errors)), "mmc_device_read");
  if (errors) throw DeviceError(MMC_ERR_INVALID, "KEigenvalue::Solve: bank resampling read outside its slice");
  // active-cycle statistics
  if (active) {
    Real sum = 0;
    for (uint64_t c = last_inactive; c < last_active; c++) sum += result_.k_cycle[c];
    result_.k_mean = sum / static_cast<Real>(active);
    Real ss = 0;
    for (uint64_t c = last_inactive; c < last_active; c++) ss += (result_.k_cycle[c] - result_.k_mean) * (result_.k_cycle[c] - result_.k_mean);
    result_.k_std = active > 1 ? std::sqrt(ss / static_cast<Real>(active * (active - 1))) : 0;
  }
  // integer tallies -> Scorable scores (exact below 2^53)
  std::vector<uint64_t> h_scores(std::max<size_t>(bins, 1)), h_squares(std::max<size_t>(bins, 1));
  Check(mmc_device_read(w, h_scores.data(), d_scores.as<uint64_t>(), h_scores.size() * sizeof(uint64_t)), "mmc_device_read");
  Check(mmc_device_read(w, h_squares.data(), d_squares.as<uint64_t>(), h_squares.size() * sizeof(uint64_t)), "mmc_device_read");
  size_t offset = 0;
  for (Estimator& e : result.estimators) {
    for (size_t i = 0; i < e.scores.size(); i++) {
      e.scores[i] += static_cast<Real>(h_scores[offset + i]);
      e.square_scores[i] += static_cast<Real>(h_squares[offset + i]);
    }
    offset += e.scores.size();
  }
  return result;
}

}  // namespace minimc
