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
// With a communicator set (Driver::SetComm: one process per GPU, NCCL), this process owns source indices
// [rank*N/P, (rank+1)*N/P) of every generation; after each generation mmc_bank_exchange all-gathers the ranks' bank
// sizes and status words and moves the few sites a rank's next sources need from its neighbours, mmc_bank_resample
// builds the next source bank, and one packed all-reduce at the end sums tallies, counters and the collision
// estimator's per-cycle sums.  Every step is order-based: k of every cycle, the banks and the tallies are identical
// for any number of ranks.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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
  if (world_size > 1 && !comm)
    throw std::runtime_error("KEigenvalue::Solve: world_size > 1 needs a communicator (Driver::SetComm / InitCommFromEnvironment)");
  const mmc_world* w = device_world_handle();
  const uint64_t N = batchsize;
  const uint64_t P = static_cast<uint64_t>(world_size), R = static_cast<uint64_t>(rank);
  // this rank's source indices of every generation: [first, first + n_local)
  const uint64_t first = R * N / P, n_local = (R + 1) * N / P - first;
  const uint64_t active = last_active - last_inactive;
  EstimatorSet result = init_estimator_set;
  result.total_weight = static_cast<Real>(N) * static_cast<Real>(active ? active : 1);
  const std::vector<mmc_estimator_desc> estimators = FlattenEstimators(result);
  const size_t bins = result.total_bins();
  // a generation can bank more than its sources: room for k up to bank_capacity_factor (per rank, like the sources)
  auto capacity_of = [&](uint64_t r) {
    const uint64_t n_r = (r + 1) * N / P - r * N / P;
    return static_cast<uint64_t>(bank_capacity_factor * static_cast<double>(std::max<uint64_t>(n_r, 1))) + 1024;
  };
  const uint64_t capacity = capacity_of(R);
  // the site range a rank's next sources are drawn from spans at most n_local * M / N + 2 sites
  const uint64_t slice_capacity = static_cast<uint64_t>(bank_capacity_factor * static_cast<double>(std::max<uint64_t>(n_local, 1))) + 1024 + 2;
  constexpr size_t kCounterWords = sizeof(mmc_counters) / sizeof(uint64_t);
  DeviceBuffer bank_source{w, std::max<uint64_t>(n_local, 1) * sizeof(mmc_site)}, bank_fission{w, capacity * sizeof(mmc_site)};
  DeviceBuffer slice{w, (P > 1 ? slice_capacity : 1) * sizeof(mmc_site)};
  DeviceBuffer d_n_out{w, sizeof(uint64_t)};
  // [scores | squares | counters | resample errors | k collision sums of the active cycles' generations]: one packed
  // all-reduce at the end
  const size_t words = 2 * bins + kCounterWords + 1;
  DeviceBuffer d_tally{w, words * sizeof(uint64_t)};
  uint64_t* d_scores = d_tally.as<uint64_t>();
  uint64_t* d_squares = d_scores + bins;
  mmc_counters* d_counters = reinterpret_cast<mmc_counters*>(d_squares + bins);
  uint64_t* d_errors = d_squares + bins + kCounterWords;
  DeviceBuffer d_k_collision{w, std::max<uint64_t>(last_active, 1) * sizeof(uint64_t)};  // one sum per cycle
  run_options.tracking = tracking;
  void* stream = run_options.stream ? run_options.stream : mmc_world_stream(w);
  // KEigenvalue.cpp:29-33: source.Sample(s) for s = 1 .. batchsize; this rank's part
  Check(mmc_source_bank_sample(w, &source.desc, 1, first, n_local, &run_options, bank_source.as<mmc_site>()), "mmc_source_bank_sample");
  result_ = KResult{};
  std::vector<uint64_t> counts(P), statuses(P);
  Check(mmc_device_read(w, &counters, d_counters, sizeof(counters)), "mmc_device_read");  // (synchronises: the source bank is sampled)
  auto t_begin = std::chrono::steady_clock::now(), t_active = t_begin;
  // MMC_K_TRACE=1: host time of every phase of every cycle on stderr (each phase ended by a device synchronisation)
  const bool trace = std::getenv("MMC_K_TRACE") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double, std::milli>(b - a).count();
  };
  for (uint64_t cycle = 0; cycle < last_active; cycle++) {
    const bool score = cycle >= last_inactive;
    const auto t0 = now();
    if (cycle == last_inactive) t_active = std::chrono::steady_clock::now();  // (the previous cycle ended with a synchronisation)
    Check(mmc_generation_run(
              w, bank_source.as<mmc_site>(), n_local, estimators.data(), static_cast<int32_t>(estimators.size()), score ? 1 : 0,
              &run_options, bank_fission.as<mmc_site>(), capacity, d_n_out.as<uint64_t>(), d_scores, d_squares, d_counters,
              d_k_collision.as<uint64_t>() + cycle),
          "mmc_generation_run");
    Check(mmc_device_read(w, &counters, d_counters, sizeof(counters)), "mmc_device_read");
    const auto t1 = now();  // (mmc_device_read synchronised: the generation and the ordering of its bank are done)
    // what this rank saw; decided alike on every rank after the all-gather below
    const uint64_t status = (counters.n_lost ? 1u : 0u) | (counters.n_physics_errors ? 2u : 0u) | (counters.n_capacity_overflow ? 4u : 0u);
    uint64_t M = 0, slice_first = 0, slice_n = 0;
    if (P > 1) {
      Check(mmc_bank_exchange(comm, bank_fission.as<mmc_site>(), d_n_out.as<uint64_t>(), status, N, slice.as<mmc_site>(),
                              slice_capacity, counts.data(), statuses.data(), &slice_first, &slice_n, stream),
            "mmc_bank_exchange");
    } else {
      Check(mmc_device_read(w, &counts[0], d_n_out.as<uint64_t>(), sizeof(uint64_t)), "mmc_device_read");
      statuses[0] = status;
    }
    const auto t2 = now();
    uint64_t any = 0;
    for (uint64_t r = 0; r < P; r++) {
      M += counts[r];
      any |= statuses[r];
      if (counts[r] > capacity_of(r)) any |= 4u;
    }
    if (any & 1u) throw DeviceError(MMC_ERR_LOST_PARTICLE, "KEigenvalue::Solve: particle(s) outside every cell");
    if (any & 2u) throw DeviceError(MMC_ERR_PHYSICS, "KEigenvalue::Solve: a branch the reference asserts unreachable was reached");
    if (any & 4u) throw DeviceError(MMC_ERR_CAPACITY, "KEigenvalue::Solve: fission bank overflow (raise bank_capacity_factor / secondary_capacity)");
    result_.k_cycle.push_back(static_cast<Real>(M) / static_cast<Real>(N));
    result_.bank_sizes.push_back(M);
    if (M == 0) throw DeviceError(MMC_ERR_PHYSICS, "KEigenvalue::Solve: the fission chain died out (empty fission bank)");
    if (P > 1)
      Check(mmc_bank_resample(w, slice.as<mmc_site>(), slice_first, slice_n, M, N, first, n_local, &run_options,
                              bank_source.as<mmc_site>(), d_errors),
            "mmc_bank_resample");
    else
      Check(mmc_bank_resample(w, bank_fission.as<mmc_site>(), 0, M, M, N, 0, N, &run_options, bank_source.as<mmc_site>(), d_errors),
            "mmc_bank_resample");
    if (trace) {
      uint64_t dummy = 0;
      Check(mmc_device_read(w, &dummy, d_n_out.as<uint64_t>(), sizeof(dummy)), "mmc_device_read");  // synchronise
      std::fprintf(stderr, "[rank %d] cycle %llu: generation %.3f ms, exchange (host view) %.3f ms, resample + pieces %.3f ms, M %llu\n",
                   rank, static_cast<unsigned long long>(cycle), ms(t0, t1), ms(t1, t2), ms(t2, now()), static_cast<unsigned long long>(M));
    }
  }
  // tallies, counters, resampling errors and the collision estimator's sums: one all-reduce each buffer
  if (P > 1) {
    Check(mmc_tally_allreduce(comm, d_tally.as<uint64_t>(), words, stream), "mmc_tally_allreduce");
    Check(mmc_tally_allreduce(comm, d_k_collision.as<uint64_t>(), last_active, stream), "mmc_tally_allreduce");
    result_.exchange_ms = mmc_comm_exchange_ms(comm);
  }
  std::vector<uint64_t> h_tally(words), h_k(std::max<uint64_t>(last_active, 1));
  Check(mmc_device_read(w, h_tally.data(), d_tally.as<uint64_t>(), words * sizeof(uint64_t)), "mmc_device_read");
  Check(mmc_device_read(w, h_k.data(), d_k_collision.as<uint64_t>(), h_k.size() * sizeof(uint64_t)), "mmc_device_read");
  const auto t_end = std::chrono::steady_clock::now();
  if (last_inactive == last_active) t_active = t_end;
  result_.inactive_seconds = std::chrono::duration<double>(t_active - t_begin).count();
  result_.active_seconds = std::chrono::duration<double>(t_end - t_active).count();
  std::memcpy(&counters, h_tally.data() + 2 * bins, sizeof(counters));
  if (h_tally[2 * bins + kCounterWords]) throw DeviceError(MMC_ERR_INVALID, "KEigenvalue::Solve: bank resampling read outside its slice");
  for (uint64_t c = 0; c < last_active; c++)
    result_.k_collision_cycle.push_back(static_cast<Real>(h_k[c]) / MMC_K_COLLISION_ONE / static_cast<Real>(N));
  // active-cycle statistics of both estimators: mean and sample standard deviation of the mean
  auto statistics = [&](const std::vector<Real>& k, Real& mean, Real& std_dev) {
    mean = std_dev = 0;
    if (!active) return;
    Real sum = 0;
    for (uint64_t c = last_inactive; c < last_active; c++) sum += k[c];
    mean = sum / static_cast<Real>(active);
    Real ss = 0;
    for (uint64_t c = last_inactive; c < last_active; c++) ss += (k[c] - mean) * (k[c] - mean);
    std_dev = active > 1 ? std::sqrt(ss / static_cast<Real>(active * (active - 1))) : 0;
  };
  statistics(result_.k_cycle, result_.k_mean, result_.k_std);
  statistics(result_.k_collision_cycle, result_.k_collision_mean, result_.k_collision_std);
  // integer tallies -> Scorable scores (exact below 2^53)
  size_t offset = 0;
  for (Estimator& e : result.estimators) {
    for (size_t i = 0; i < e.scores.size(); i++) {
      e.scores[i] += static_cast<Real>(h_tally[offset + i]);
      e.square_scores[i] += static_cast<Real>(h_tally[bins + offset + i]);
    }
    offset += e.scores.size();
  }
  return result;
}

}  // namespace minimc
