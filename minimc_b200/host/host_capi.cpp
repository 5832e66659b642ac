// C entry points of the host layer (include/minimc_b200.h "host layer").
#include <cstring>
#include <exception>
#include <string>

#include "minimc.hpp"

namespace mmc {
int set_last_error(int status, const std::string& message);  // capi.cu
}

struct mmc_driver {
  std::unique_ptr<minimc::Driver> driver;
  minimc::EstimatorSet result;
  bool solved = false;
};

namespace {

template <typename F> int Guard(F&& f) {
  try {
    return f();
  } catch (const minimc::DeviceError& e) {
    return mmc::set_last_error(e.status, e.what());
  } catch (const std::exception& e) {
    return mmc::set_last_error(MMC_ERR_INVALID, e.what());
  } catch (...) {
    return mmc::set_last_error(MMC_ERR_INVALID, "unknown C++ exception");
  }
}

size_t CopyOut(const std::string& text, char* buf, size_t cap) {
  if (buf && cap) {
    const size_t n = std::min(cap - 1, text.size());
    std::memcpy(buf, text.data(), n);
    buf[n] = 0;
  }
  return text.size();
}

}  // namespace

extern "C" {

int mmc_driver_create(const char* xml_path, mmc_driver** out) {
  return Guard([&] {
    if (!xml_path || !out) return mmc::set_last_error(MMC_ERR_INVALID, "xml_path / out is NULL");
    *out = nullptr;
    auto d = std::make_unique<mmc_driver>();
    d->driver = minimc::Driver::Create(xml_path);
    d->result = d->driver->init_estimator_set;
    *out = d.release();
    return static_cast<int>(MMC_OK);
  });
}

int mmc_driver_create_from_string(const char* xml_text, mmc_driver** out) {
  return Guard([&] {
    if (!xml_text || !out) return mmc::set_last_error(MMC_ERR_INVALID, "xml_text / out is NULL");
    *out = nullptr;
    auto d = std::make_unique<mmc_driver>();
    d->driver = minimc::Driver::CreateFromString(xml_text);
    d->result = d->driver->init_estimator_set;
    *out = d.release();
    return static_cast<int>(MMC_OK);
  });
}

void mmc_driver_destroy(mmc_driver* driver) { delete driver; }

int mmc_driver_set_options(mmc_driver* driver, const mmc_run_options* options) {
  if (!driver || !options) return mmc::set_last_error(MMC_ERR_INVALID, "driver / options is NULL");
  if (options->struct_size != sizeof(mmc_run_options)) return mmc::set_last_error(MMC_ERR_INVALID, "mmc_run_options ABI mismatch");
  const int32_t tracking = driver->driver->run_options.tracking;
  driver->driver->run_options = *options;
  driver->driver->run_options.tracking = tracking;
  return MMC_OK;
}

int mmc_driver_set_shard(mmc_driver* driver, int32_t rank, int32_t world_size) {
  if (!driver || world_size < 1 || rank < 0 || rank >= world_size)
    return mmc::set_last_error(MMC_ERR_INVALID, "bad rank / world_size");
  driver->driver->rank = rank;
  driver->driver->world_size = world_size;
  return MMC_OK;
}

int mmc_driver_solve(mmc_driver* driver) {
  return Guard([&] {
    if (!driver) return mmc::set_last_error(MMC_ERR_INVALID, "driver is NULL");
    driver->result = driver->driver->Solve();
    driver->solved = true;
    return static_cast<int>(MMC_OK);
  });
}

uint64_t mmc_driver_batchsize(const mmc_driver* driver) { return driver ? driver->driver->batchsize : 0; }

uint64_t mmc_driver_total_bins(const mmc_driver* driver) { return driver ? driver->result.total_bins() : 0; }

int mmc_driver_scores(const mmc_driver* driver, double* scores, double* square_scores, uint64_t n) {
  if (!driver || n != driver->result.total_bins() || (n && (!scores || !square_scores)))
    return mmc::set_last_error(MMC_ERR_INVALID, "bad arguments to mmc_driver_scores");
  size_t offset = 0;
  for (const auto& e : driver->result.estimators) {
    std::memcpy(scores + offset, e.scores.data(), e.scores.size() * sizeof(double));
    std::memcpy(square_scores + offset, e.square_scores.data(), e.scores.size() * sizeof(double));
    offset += e.scores.size();
  }
  return MMC_OK;
}

int mmc_driver_add_scores(mmc_driver* driver, const double* scores, const double* square_scores, uint64_t n) {
  if (!driver || n != driver->result.total_bins() || (n && (!scores || !square_scores)))
    return mmc::set_last_error(MMC_ERR_INVALID, "bad arguments to mmc_driver_add_scores");
  size_t offset = 0;
  for (auto& e : driver->result.estimators) {
    for (size_t i = 0; i < e.scores.size(); i++) {
      e.scores[i] += scores[offset + i];
      e.square_scores[i] += square_scores[offset + i];
    }
    offset += e.scores.size();
  }
  return MMC_OK;
}

int mmc_driver_counters(const mmc_driver* driver, mmc_counters* counters) {
  if (!driver || !counters) return mmc::set_last_error(MMC_ERR_INVALID, "driver / counters is NULL");
  *counters = driver->driver->counters;
  return MMC_OK;
}

size_t mmc_driver_output(const mmc_driver* driver, char* buf, size_t cap) {
  if (!driver) return 0;
  // minimc.cpp:20-21
  return CopyOut(std::to_string(driver->driver->batchsize) + "\n" + driver->result.to_string(), buf, cap);
}

size_t mmc_driver_world_json(const mmc_driver* driver, char* buf, size_t cap) {
  if (!driver) return 0;
  try {
    return CopyOut(minimc::FlatWorld{driver->driver->world}.to_json(), buf, cap);
  } catch (const std::exception& e) {
    mmc::set_last_error(MMC_ERR_INVALID, e.what());
    return 0;
  }
}

int mmc_driver_run_device(mmc_driver* driver, uint64_t first_history, uint64_t n_histories, uint64_t* d_scores,
                          uint64_t* d_square_scores, mmc_counters* d_counters, void* stream) {
  return Guard([&] {
    if (!driver) return mmc::set_last_error(MMC_ERR_INVALID, "driver is NULL");
    auto* fixed = dynamic_cast<minimc::FixedSource*>(driver->driver.get());
    if (!fixed) return mmc::set_last_error(MMC_ERR_INVALID, "not a fixed-source problem");
    fixed->RunDevice(first_history, n_histories, d_scores, d_square_scores, d_counters, stream);
    return static_cast<int>(MMC_OK);
  });
}

void mmc_driver_release_device(mmc_driver* driver) {
  if (driver) driver->driver->ReleaseDevice();
}

int mmc_driver_refresh_device(mmc_driver* driver) {
  if (!driver) return mmc::set_last_error(MMC_ERR_INVALID, "driver is NULL");
  return Guard([&] {
    driver->driver->RefreshDevice();
    return static_cast<int>(MMC_OK);
  });
}

uint64_t mmc_driver_table_bytes(mmc_driver* driver) {
  if (!driver) return 0;
  try {
    return driver->driver->DeviceTableBytes();
  } catch (const std::exception& e) {
    mmc::set_last_error(MMC_ERR_INVALID, e.what());
    return 0;
  }
}

uint64_t mmc_driver_last_launches(mmc_driver* driver) {
  if (!driver) return 0;
  try {
    return mmc_world_last_launches(driver->driver->device_world_handle());
  } catch (const std::exception& e) {
    mmc::set_last_error(MMC_ERR_INVALID, e.what());
    return 0;
  }
}

void mmc_driver_last_kernel_ms(mmc_driver* driver, double* flight_ms, double* tsl_ms) {
  if (flight_ms) *flight_ms = 0;
  if (tsl_ms) *tsl_ms = 0;
  if (!driver) return;
  try {
    mmc_world_last_kernel_ms(driver->driver->device_world_handle(), flight_ms, tsl_ms);
  } catch (const std::exception& e) {
    mmc::set_last_error(MMC_ERR_INVALID, e.what());
  }
}

double mmc_driver_last_boundary_ms(mmc_driver* driver) {
  if (!driver) return 0;
  try {
    return mmc_world_last_boundary_ms(driver->driver->device_world_handle());
  } catch (const std::exception& e) {
    mmc::set_last_error(MMC_ERR_INVALID, e.what());
    return 0;
  }
}

int mmc_driver_trace(mmc_driver* driver, uint64_t first_history, uint64_t n_histories, mmc_event_record* records,
                     size_t cap, size_t* n_records) {
  return Guard([&] {
    if (!driver || !records || !n_records) return mmc::set_last_error(MMC_ERR_INVALID, "driver / records / n_records is NULL");
    auto* fixed = dynamic_cast<minimc::FixedSource*>(driver->driver.get());
    if (!fixed) return mmc::set_last_error(MMC_ERR_INVALID, "not a fixed-source problem");
    const auto result = fixed->Trace(first_history, n_histories, cap);
    std::memcpy(records, result.data(), result.size() * sizeof(mmc_event_record));
    *n_records = result.size();
    return static_cast<int>(MMC_OK);
  });
}

int mmc_driver_set_comm(mmc_driver* driver, mmc_comm* comm) {
  if (!driver || !comm) return mmc::set_last_error(MMC_ERR_INVALID, "driver / comm is NULL");
  driver->driver->SetComm(comm);
  return MMC_OK;
}

int mmc_driver_init_comm_from_environment(mmc_driver* driver) {
  return Guard([&] {
    if (!driver) return mmc::set_last_error(MMC_ERR_INVALID, "driver is NULL");
    driver->driver->InitCommFromEnvironment();
    return static_cast<int>(MMC_OK);
  });
}

int mmc_driver_k_collision(const mmc_driver* driver, double* k_mean, double* k_std, double* k_cycle, size_t cap, size_t* n_cycles,
                           double* exchange_ms) {
  if (!driver) return mmc::set_last_error(MMC_ERR_INVALID, "driver is NULL");
  const auto* k = dynamic_cast<const minimc::KEigenvalue*>(driver->driver.get());
  if (!k) return mmc::set_last_error(MMC_ERR_INVALID, "not a k-eigenvalue problem");
  if (k_mean) *k_mean = k->result().k_collision_mean;
  if (k_std) *k_std = k->result().k_collision_std;
  if (n_cycles) *n_cycles = k->result().k_collision_cycle.size();
  if (k_cycle)
    for (size_t i = 0; i < std::min(cap, k->result().k_collision_cycle.size()); i++) k_cycle[i] = k->result().k_collision_cycle[i];
  if (exchange_ms) *exchange_ms = k->result().exchange_ms;
  return MMC_OK;
}

int mmc_driver_cycle_seconds(const mmc_driver* driver, double* inactive_seconds, double* active_seconds) {
  if (!driver) return mmc::set_last_error(MMC_ERR_INVALID, "driver is NULL");
  const auto* k = dynamic_cast<const minimc::KEigenvalue*>(driver->driver.get());
  if (!k) return mmc::set_last_error(MMC_ERR_INVALID, "not a k-eigenvalue problem");
  if (inactive_seconds) *inactive_seconds = k->result().inactive_seconds;
  if (active_seconds) *active_seconds = k->result().active_seconds;
  return MMC_OK;
}

int mmc_driver_keff(const mmc_driver* driver, double* k_mean, double* k_std, double* k_cycle, size_t cap, size_t* n_cycles) {
  if (!driver) return mmc::set_last_error(MMC_ERR_INVALID, "driver is NULL");
  const auto* k = dynamic_cast<const minimc::KEigenvalue*>(driver->driver.get());
  if (!k) return mmc::set_last_error(MMC_ERR_INVALID, "not a k-eigenvalue problem");
  if (k_mean) *k_mean = k->result().k_mean;
  if (k_std) *k_std = k->result().k_std;
  if (n_cycles) *n_cycles = k->result().k_cycle.size();
  if (k_cycle)
    for (size_t i = 0; i < std::min(cap, k->result().k_cycle.size()); i++) k_cycle[i] = k->result().k_cycle[i];
  return MMC_OK;
}

}  // extern "C"
