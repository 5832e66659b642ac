// C ABI of include/minimc_b200.h: table validation, flattening into the device
// blob, scratch management and kernel launches.  No CPU transport path exists
// in this library: without a CUDA device every compute entry point fails with
// MMC_ERR_NO_DEVICE.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <type_traits>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/minimc_b200.h"
#define MMC_DECLARE_BOTH_VARIANTS  // the launch functions of both RNG modes: mmc::lcg::*, mmc::ctr::*
#include "kernels.h"
#include "world_blob.h"

using namespace mmc;
// the launch functions that do not touch a generator are taken from the minstd build
using mmc::lcg::bank_scan_blocks;
using mmc::lcg::event_kernels_per_pass;
using mmc::lcg::launch_evaluate_rows;
using mmc::lcg::launch_expand_dense;
using mmc::lcg::launch_order_bank;
using mmc::lcg::launch_test_geometry;
// fn(args) of the RNG mode's build
#define MMC_BY_RNG(counter_, fn_, ...) ((counter_) ? mmc::ctr::fn_(__VA_ARGS__) : mmc::lcg::fn_(__VA_ARGS__))

namespace {

// histories in flight in the event-split schedule.  Measured on B200, single_zone, 2^23 histories: 2^18 slots 4.4e7,
// 2^20 8.06e7, 2^21 7.99e7, 2^22 7.5e7 hist/s (fewer slots: more, smaller passes; more slots: a longer drain tail)
constexpr uint32_t kDefaultEventSlots = 1u << 20;
// batches of 2^24 histories and more (r01j, dense tables, 2^25 histories: 2^19 slots 1.318e8, 2^20 1.387e8,
// 1.5 * 2^20 1.407e8, 2^21 1.419e8, 2^22 1.397e8 hist/s): the longer drain tail of more slots is amortised
constexpr uint32_t kLargeBatchEventSlots = 1u << 21;
constexpr uint64_t kLargeBatchHistories = 1ull << 24;
// batches below 3 * 2^20 histories (r02w/r02x, single_zone, ms per step at 2^19 | 2^20 slots: 10^6 histories -- the
// deck as shipped -- 11.28 | 11.95, 2 * 10^6 16.49 | 17.11, 2^22 28.10 | 27.64, 1.25 * 10^7 72.6 | 67.6; 2^18 and
// 3/4 * 2^20 at 10^6: 12.00 and 11.48).  Why fewer slots than histories pay there was not established (50 MB of slot
// state instead of 100 is one candidate); the choice is the measurement's.
constexpr uint32_t kSmallBatchEventSlots = 1u << 19;
constexpr uint64_t kSmallBatchHistories = 3ull << 20;
// live histories at or below which the drain of an event-split run is handed to the fused kernel
constexpr uint32_t kDefaultEventHandover = 1u << 15;

thread_local std::string g_error;

int fail(int status, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_error = buf;
  return status;
}

#define MMC_CUDA(expr)                                                                    \
  do {                                                                                    \
    const cudaError_t e_ = (expr);                                                        \
    if (e_ != cudaSuccess) return fail(MMC_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)

// Builds the blob: appends arrays at 16-byte aligned offsets.
class BlobBuilder {
public:
  BlobBuilder() { bytes.resize(sizeof(WorldHeader)); pad(); }
  template <typename T> uint32_t add(const T* src, size_t n) {
    pad();
    const uint32_t off = static_cast<uint32_t>(bytes.size());
    if (n) {
      bytes.resize(bytes.size() + n * sizeof(T));
      std::memcpy(bytes.data() + off, src, n * sizeof(T));
    }
    return off;
  }
  WorldHeader& header() { return *reinterpret_cast<WorldHeader*>(bytes.data()); }
  void pad() { bytes.resize((bytes.size() + 15) & ~size_t{15}); }
  std::vector<char> bytes;
};

}  // namespace

namespace mmc {
// used by the host layer (minimc_b200/host/host_capi.cpp) to report C++ exceptions
int set_last_error(int status, const std::string& message) {
  g_error = message;
  return status;
}
}  // namespace mmc

// Scratch recycling between worlds of one process: a Driver that is rebuilt for every batch (the e2e leg of bench.py
// does exactly that) would otherwise cudaMalloc / cudaFree ~1 GB of particle state and pending tables per solve.
// One parked buffer per (device, kind); a world takes it on its first need and parks its own on destruction.
namespace {
enum ScratchKind { kScratchSites = 0, kScratchPending = 1, kScratchEvent = 2, kScratchKinds = 3 };
struct ParkedBuffer {
  void* ptr = nullptr;
  size_t bytes = 0;
};
std::mutex g_scratch_mutex;
std::map<std::pair<int, int>, ParkedBuffer> g_scratch_parked;

// a buffer of at least `need` bytes: the parked one if it is large enough, else a fresh allocation
cudaError_t scratch_acquire(int device, ScratchKind kind, size_t need, void** ptr, size_t* bytes) {
  {
    std::lock_guard<std::mutex> lock(g_scratch_mutex);
    auto it = g_scratch_parked.find({device, kind});
    if (it != g_scratch_parked.end() && it->second.bytes >= need) {
      *ptr = it->second.ptr;
      *bytes = it->second.bytes;
      g_scratch_parked.erase(it);
      return cudaSuccess;
    }
  }
  const cudaError_t e = cudaMalloc(ptr, need);
  *bytes = e == cudaSuccess ? need : 0;
  return e;
}

void scratch_release(int device, ScratchKind kind, void* ptr, size_t bytes) {
  if (!ptr) return;
  void* drop = ptr;
  {
    std::lock_guard<std::mutex> lock(g_scratch_mutex);
    ParkedBuffer& slot = g_scratch_parked[{device, kind}];
    if (bytes > slot.bytes) {
      drop = slot.ptr;
      slot.ptr = ptr;
      slot.bytes = bytes;
    }
  }
  if (drop) cudaFree(drop);
}
}  // namespace

struct mmc_world {
  // One run at a time per world: the run entry points share the world's scratch (history counter, pending and
  // secondary tables, event-split state, tally mirrors).  A second host thread waits here; the asynchronous
  // device-buffer entry points must additionally be given ONE stream per world (include/minimc_b200.h).
  std::recursive_mutex run_mutex;
  int device = 0;
  char* d_blob = nullptr;
  char* h_blob = nullptr;  // pinned host copy of the image: the source of every upload
  // device tally / counter buffers and their pinned host mirror, reused by mmc_fixed_source_run
  unsigned long long* d_tally = nullptr;
  unsigned long long* h_tally = nullptr;
  size_t tally_words = 0;
  uint32_t blob_bytes = 0;
  WorldHeader header{};
  bool has_fission = false;
  int sm_count = 0;
  size_t smem_optin = 0;
  size_t l2_persist_bytes = 0;  // persisting-L2 set-aside asked for so far (keep_tables_in_l2)
  cudaStream_t stream = nullptr;
  // scratch reused between runs
  BankSite* d_sites = nullptr;
  size_t sites_bytes = 0;
  uint2* d_pending = nullptr;
  size_t pending_bytes = 0;
  double* d_bounds = nullptr;
  size_t bounds_bytes = 0;
  unsigned long long* d_next = nullptr;
  // event-split schedule: SoA particle state, queues, counter replicas (one allocation), pinned queue counts
  char* d_event = nullptr;
  size_t event_bytes = 0;
  unsigned int* h_event_counts = nullptr;
  uint64_t last_launches = 0;  // kernels launched by the last event-split run
  double last_flight_ms = 0, last_tsl_ms = 0, last_boundary_ms = 0;  // profile mode: device time per kernel family
  // sensitivities: per-thread pending entries, device tallies with a pinned mirror
  SensitivityPending* d_sens_pending = nullptr;
  size_t sens_pending_bytes = 0;
  double* d_sens = nullptr;
  double* h_sens = nullptr;
  size_t sens_words = 0;
  // k-eigenvalue scratch
  BankSite* d_unordered = nullptr;
  size_t unordered_bytes = 0;
  uint32_t* d_child_count = nullptr;
  unsigned long long* d_child_start = nullptr;
  unsigned long long* d_block_sums = nullptr;
  size_t parents_capacity = 0;
};

namespace {

int validate_table(const mmc_table1d& t, const char* what, int nuclide, bool required) {
  if (t.n == 0 && !required) return MMC_OK;
  if (t.n == 0 || !t.x || !t.y) return fail(MMC_ERR_INVALID, "nuclide %d: %s table is empty", nuclide, what);
  if (t.n > 0x7fffffffull) return fail(MMC_ERR_INVALID, "nuclide %d: %s table too large", nuclide, what);
  for (uint64_t i = 1; i < t.n; i++)
    if (!(t.x[i] > t.x[i - 1])) return fail(MMC_ERR_INVALID, "nuclide %d: %s table keys must be strictly increasing", nuclide, what);
  return MMC_OK;
}

int validate_partitions(const mmc_tsl_partition* p, int n, const char* what, int nuclide) {
  if (n < 1 || !p) return fail(MMC_ERR_INVALID, "nuclide %d: thermal scattering data without %s partitions", nuclide, what);
  double previous = 0;
  for (int i = 0; i < n; i++) {
    const mmc_tsl_partition& q = p[i];
    if (!q.n_cdf || !q.n_grid || !q.n_temperature || !q.rank || !q.cdf || !q.grid || !q.temperature || !q.cdf_modes ||
        !q.singular_values || !q.grid_T_modes)
      return fail(MMC_ERR_INVALID, "nuclide %d: %s partition %d has an empty table", nuclide, what, i);
    for (uint64_t k = 0; k < q.n_grid; k++) {
      // the reference asserts this ordering (ThermalScattering.cpp:46-51,79-84)
      if (!(previous < q.grid[k])) return fail(MMC_ERR_INVALID, "nuclide %d: %s partition grids must be increasing", nuclide, what);
      previous = q.grid[k];
    }
  }
  return MMC_OK;
}

int validate_ce(const mmc_world_desc* d) {
  for (int n = 0; n < d->n_nuclides; n++) {
    const mmc_ce_nuclide& c = d->ce->nuclides[n];
    if (int s = validate_table(c.total, "total", n, true)) return s;
    if (c.n_reactions < 0 || c.n_reactions > kMaxCeReactions || (c.n_reactions && !c.reactions))
      return fail(MMC_ERR_INVALID, "nuclide %d: %d reactions (at most %d supported)", n, c.n_reactions, kMaxCeReactions);
    for (int r = 0; r < c.n_reactions; r++) {
      const mmc_ce_reaction& x = c.reactions[r];
      if (x.kind != MMC_REACTION_CAPTURE && x.kind != MMC_REACTION_SCATTER && x.kind != MMC_REACTION_FISSION)
        return fail(MMC_ERR_INVALID, "nuclide %d reaction %d: unknown kind %d", n, r, x.kind);
      if (int s = validate_table(x.xs, "reaction", n, true)) return s;
      if (x.has_nubar)
        if (int s = validate_table(x.nubar, "nubar", n, true)) return s;
      if (x.tsl) {
        const mmc_tsl_desc& t = *x.tsl;
        if (x.kind != MMC_REACTION_SCATTER) return fail(MMC_ERR_INVALID, "nuclide %d: only scatter may carry thermal scattering data", n);
        if (int s = validate_table(t.majorant, "tsl majorant", n, true)) return s;
        if (!t.n_energy || !t.n_temperature || !t.rank || !t.energy || !t.temperature || !t.xs_E || !t.xs_S || !t.xs_T)
          return fail(MMC_ERR_INVALID, "nuclide %d: empty thermal scattering cross-section tables", n);
        if (int s = validate_partitions(t.beta_partitions, t.n_beta_partitions, "beta", n)) return s;
        if (int s = validate_partitions(t.alpha_partitions, t.n_alpha_partitions, "alpha", n)) return s;
      }
    }
  }
  return MMC_OK;
}

int validate_world(const mmc_world_desc* d) {
  if (!d) return fail(MMC_ERR_INVALID, "world desc is NULL");
  if (d->struct_size != sizeof(mmc_world_desc) || d->abi_version != MMC_ABI_VERSION)
    return fail(MMC_ERR_INVALID, "mmc_world_desc ABI mismatch: struct_size %u (expected %zu), abi_version %u (expected %d)",
                d->struct_size, sizeof(mmc_world_desc), d->abi_version, MMC_ABI_VERSION);
  if (d->n_surfaces < 1 || d->n_cells < 1) return fail(MMC_ERR_INVALID, "world needs at least one surface and one cell");
  if (d->n_groups < 0) return fail(MMC_ERR_INVALID, "n_groups = %d", d->n_groups);
  if (d->n_groups == 0 && d->n_nuclides > 0 && (!d->ce || !d->ce->nuclides))
    return fail(MMC_ERR_INVALID, "continuous-energy world (n_groups = 0) without mmc_ce_desc tables");
  if (d->n_materials < 0 || d->n_nuclides < 0) return fail(MMC_ERR_INVALID, "negative table size");
  if (!d->surface_type || !d->surface_param || !d->cell_material || !d->cell_surface_begin || !d->cell_surface_index ||
      !d->cell_surface_sense)
    return fail(MMC_ERR_INVALID, "geometry table pointer is NULL");
  for (int i = 0; i < d->n_surfaces; i++)
    if (d->surface_type[i] < MMC_SURF_SPHERE || d->surface_type[i] > MMC_SURF_CYLINDERX)
      return fail(MMC_ERR_INVALID, "surface %d: unknown type %d", i, d->surface_type[i]);
  if (d->cell_surface_begin[0] != 0) return fail(MMC_ERR_INVALID, "cell_surface_begin[0] must be 0");
  for (int c = 0; c < d->n_cells; c++) {
    if (d->cell_surface_begin[c + 1] <= d->cell_surface_begin[c])
      return fail(MMC_ERR_INVALID, "cell %d has no surfaces (Cell::NearestSurface would throw, Cell.cpp:44-47)", c);
    if (d->cell_material[c] < -1 || d->cell_material[c] >= d->n_materials)
      return fail(MMC_ERR_INVALID, "cell %d: material index %d out of range", c, d->cell_material[c]);
    for (int k = d->cell_surface_begin[c]; k < d->cell_surface_begin[c + 1]; k++)
      if (d->cell_surface_index[k] < 0 || d->cell_surface_index[k] >= d->n_surfaces)
        return fail(MMC_ERR_INVALID, "cell %d: surface index %d out of range", c, d->cell_surface_index[k]);
  }
  if (d->n_materials > 0) {
    if (!d->material_aden || !d->material_nuclide_begin || !d->material_nuclide_index || !d->material_nuclide_afrac)
      return fail(MMC_ERR_INVALID, "material table pointer is NULL");
    if (d->material_nuclide_begin[0] != 0) return fail(MMC_ERR_INVALID, "material_nuclide_begin[0] must be 0");
    for (int m = 0; m < d->n_materials; m++) {
      if (d->material_nuclide_begin[m + 1] < d->material_nuclide_begin[m])
        return fail(MMC_ERR_INVALID, "material %d: decreasing CSR offsets", m);
      for (int k = d->material_nuclide_begin[m]; k < d->material_nuclide_begin[m + 1]; k++)
        if (d->material_nuclide_index[k] < 0 || d->material_nuclide_index[k] >= d->n_nuclides)
          return fail(MMC_ERR_INVALID, "material %d: nuclide index %d out of range", m, d->material_nuclide_index[k]);
    }
  }
  if (d->n_groups == 0) return validate_ce(d);
  if (d->n_nuclides > 0 &&
      (!d->mg_reaction_mask || !d->mg_total || !d->mg_capture || !d->mg_scatter || !d->mg_fission || !d->mg_nubar ||
       !d->mg_scatter_probs || !d->mg_chi))
    return fail(MMC_ERR_INVALID, "multigroup table pointer is NULL");
  return MMC_OK;
}

int ensure_scratch(mmc_world* w, size_t threads, uint32_t sec_cap, uint32_t pend_cap, size_t bounds_count) {
  const size_t need_sites = threads * sec_cap * sizeof(BankSite);
  if (need_sites > w->sites_bytes) {
    scratch_release(w->device, kScratchSites, w->d_sites, w->sites_bytes);
    w->d_sites = nullptr;
    w->sites_bytes = 0;
    MMC_CUDA(scratch_acquire(w->device, kScratchSites, need_sites, reinterpret_cast<void**>(&w->d_sites), &w->sites_bytes));
  }
  const size_t need_pending = threads * pend_cap * sizeof(uint2);
  if (need_pending > w->pending_bytes) {
    scratch_release(w->device, kScratchPending, w->d_pending, w->pending_bytes);
    w->d_pending = nullptr;
    w->pending_bytes = 0;
    MMC_CUDA(scratch_acquire(w->device, kScratchPending, need_pending, reinterpret_cast<void**>(&w->d_pending), &w->pending_bytes));
  }
  const size_t need_bounds = std::max<size_t>(bounds_count, 1) * sizeof(double);
  if (need_bounds > w->bounds_bytes) {
    if (w->d_bounds) cudaFree(w->d_bounds);
    w->d_bounds = nullptr;
    w->bounds_bytes = 0;
    MMC_CUDA(cudaMalloc(&w->d_bounds, need_bounds));
    w->bounds_bytes = need_bounds;
  }
  return MMC_OK;
}

// Direction ctor: Point{x,y,z} then Normalize() (Point.cpp:80-83, 44-47)
void normalise(const double in[3], double out[3]) {
  const double n = std::sqrt(in[0] * in[0] + in[1] * in[1] + in[2] * in[2]);
  out[0] = in[0] / n;
  out[1] = in[1] / n;
  out[2] = in[2] / n;
}

int fill_bins(const mmc_bins_desc& in, BinsSpec& out, std::vector<double>& bounds, const char* what, int e) {
  out = BinsSpec{};
  out.kind = in.kind;
  switch (in.kind) {
  case MMC_BINS_NONE:
    out.n_bins = 1;
    break;
  case MMC_BINS_LINSPACE:
  case MMC_BINS_LOGSPACE:
    if (in.n_bins < 3) return fail(MMC_ERR_INVALID, "estimator %d %s: n_bins must be bins+2 >= 3", e, what);
    if (!(in.upper > in.lower)) return fail(MMC_ERR_INVALID, "estimator %d %s: max must be strictly greater than min", e, what);
    out.n_bins = static_cast<uint32_t>(in.n_bins);
    out.lower = in.lower;
    out.upper = in.upper;
    out.width = in.width;
    out.base = in.base;
    break;
  case MMC_BINS_BOUNDARIES:
    if (in.n_bins < 1 || (in.n_bins > 1 && !in.boundaries))
      return fail(MMC_ERR_INVALID, "estimator %d %s: boundaries missing", e, what);
    for (uint64_t i = 1; i + 1 < in.n_bins; i++)
      if (!(in.boundaries[i] > in.boundaries[i - 1]))
        return fail(MMC_ERR_INVALID, "estimator %d %s: nonincreasing elements found", e, what);
    out.n_bins = static_cast<uint32_t>(in.n_bins);
    out.off_boundaries = static_cast<uint32_t>(bounds.size());
    bounds.insert(bounds.end(), in.boundaries, in.boundaries + (in.n_bins - 1));
    break;
  default:
    return fail(MMC_ERR_INVALID, "estimator %d %s: unknown bins kind %d", e, what, in.kind);
  }
  return MMC_OK;
}

struct Prepared {
  RunSpec run{};
  std::vector<double> bounds;
  LaunchConfig cfg{};
  cudaStream_t stream = nullptr;
  bool counter_rng = false;     // MMC_RNG_COUNTER: the kernels of the Philox build (mmc::ctr)
  bool profile = false;         // time every kernel of the event-split schedule with CUDA events
  bool event_schedule = false;  // event-split kernels (event_loop.cu) instead of the fused kernel
  uint32_t event_slots = 0;     // histories in flight at once
  uint32_t event_handover = 0;  // live histories at or below which the drain goes to the fused kernel (0: never)
};

int prepare_run(
    mmc_world* w, const mmc_source_desc* source, const mmc_estimator_desc* estimators, int32_t n_estimators,
    uint64_t seed0, uint64_t first_history, uint64_t n_histories, const mmc_run_options* options, bool trace,
    Prepared& out, bool generation = false) {
  if (!w) return fail(MMC_ERR_INVALID, "world handle is NULL");
  if (!source) return fail(MMC_ERR_INVALID, "source desc is NULL");
  if (n_estimators < 0 || n_estimators > kMaxEstimators)
    return fail(MMC_ERR_INVALID, "n_estimators %d out of range [0, %d]", n_estimators, kMaxEstimators);
  if (n_estimators > 0 && !estimators) return fail(MMC_ERR_INVALID, "estimators is NULL");
  mmc_run_options opt{};
  if (options) {
    if (options->struct_size != sizeof(mmc_run_options))
      return fail(MMC_ERR_INVALID, "mmc_run_options ABI mismatch: struct_size %u (expected %zu)", options->struct_size,
                  sizeof(mmc_run_options));
    opt = *options;
  }
  if (opt.tracking != MMC_TRACK_SURFACE && opt.tracking != MMC_TRACK_CELL_DELTA)
    return fail(MMC_ERR_INVALID, "unknown tracking %d", opt.tracking);
  if (opt.rng_mode != MMC_RNG_MINSTD_COMPAT && opt.rng_mode != MMC_RNG_COUNTER)
    return fail(MMC_ERR_INVALID, "unknown rng_mode %d", opt.rng_mode);
  out.counter_rng = opt.rng_mode == MMC_RNG_COUNTER;
  if (const char* env = std::getenv("MMC_RNG_MODE")) {  // development override: "counter" / "minstd"
    if (!options || options->rng_mode == MMC_RNG_MINSTD_COMPAT) out.counter_rng = std::strcmp(env, "counter") == 0;
  }
  RunSpec& run = out.run;
  // source
  for (int i = 0; i < 3; i++) run.source.position[i] = source->position[i];
  run.source.direction_kind = source->direction_kind;
  if (source->direction_kind == MMC_DIR_CONSTANT || source->direction_kind == MMC_DIR_ISOTROPIC_FLUX) {
    normalise(source->direction, run.source.direction);
  } else if (source->direction_kind != MMC_DIR_ISOTROPIC) {
    return fail(MMC_ERR_INVALID, "unknown source direction kind %d", source->direction_kind);
  }
  const bool continuous_energy = w->header.n_groups == 0;
  if (!continuous_energy && (source->group < 1 || source->group > static_cast<uint64_t>(w->header.n_groups)))
    return fail(MMC_ERR_INVALID, "source group %llu outside 1..%d", static_cast<unsigned long long>(source->group),
                w->header.n_groups);
  if (continuous_energy && !(source->energy > 0)) return fail(MMC_ERR_INVALID, "source energy must be positive (MeV)");
  run.continuous_energy = continuous_energy ? 1 : 0;
  run.source.group = source->group;
  run.source.energy = source->energy;
  // estimators
  uint64_t offset = 0;
  for (int e = 0; e < n_estimators; e++) {
    const mmc_estimator_desc& in = estimators[e];
    EstimatorSpec& es = run.estimators[e];
    if (in.surface < 0 || in.surface >= w->header.n_surfaces)
      return fail(MMC_ERR_INVALID, "estimator %d: surface index %d out of range", e, in.surface);
    es.surface = in.surface;
    es.has_direction = in.has_cosine_direction ? 1 : 0;
    if (es.has_direction) normalise(in.cosine_direction, es.direction);
    if (int s = fill_bins(in.cosine, es.cosine, out.bounds, "cosine", e)) return s;
    if (int s = fill_bins(in.energy, es.energy, out.bounds, "energy", e)) return s;
    es.stride = es.energy.n_bins;
    es.offset = offset;
    offset += static_cast<uint64_t>(es.cosine.n_bins) * es.energy.n_bins;
  }
  if (offset > 0xffffffffull) return fail(MMC_ERR_INVALID, "more than 2^32 tally bins");
  run.n_estimators = n_estimators;
  run.total_bins = offset;
  run.tracking = opt.tracking;
  run.seed0 = seed0;
  run.first_history = first_history;
  run.n_histories = n_histories;
  // capacities
  uint32_t sec = opt.secondary_capacity ? opt.secondary_capacity : 32;
  if (!w->has_fission) sec = 1;
  uint32_t pow2 = 1;
  while (pow2 < sec) pow2 <<= 1;
  run.secondary_capacity = pow2;
  run.pending_capacity = opt.pending_capacity ? opt.pending_capacity : 32;
  if (n_estimators == 0) run.pending_capacity = 1;
  run.world_bytes = w->blob_bytes;
  // continuous-energy tables are read through the read-only global path (L2-resident)
  run.world_in_smem = (!trace && !continuous_energy && w->blob_bytes <= 96 * 1024) ? 1 : 0;
  // launch shape
  int per_sm = MMC_BY_RNG(out.counter_rng, max_blocks_per_sm, run.tracking, continuous_energy, generation,
                          run.world_in_smem ? run.world_bytes : 0);
  if (per_sm < 1) per_sm = 1;
  if (opt.blocks_per_sm && static_cast<int>(opt.blocks_per_sm) < per_sm) per_sm = opt.blocks_per_sm;
  long long blocks = static_cast<long long>(w->sm_count) * per_sm;
  const long long useful = static_cast<long long>((n_histories + kThreadsPerBlock - 1) / kThreadsPerBlock);
  if (blocks > useful) blocks = std::max<long long>(useful, 1);
  out.cfg.blocks = static_cast<int>(blocks);
  const uint64_t warps = static_cast<uint64_t>(blocks) * (kThreadsPerBlock / 32);
  uint64_t chunk = n_histories / (warps * 16);
  chunk = std::min<uint64_t>(std::max<uint64_t>(chunk, 32), 2048);
  run.chunk = static_cast<uint32_t>(chunk);
  out.stream = opt.stream ? static_cast<cudaStream_t>(opt.stream) : w->stream;
  // schedule: continuous-energy fixed-source runs default to the event-split kernels
  uint32_t schedule = opt.schedule, slots = opt.event_slots;
  if (const char* env = std::getenv("MMC_SCHEDULE")) {  // development override, for A/B measurements
    if (schedule == MMC_SCHEDULE_AUTO) schedule = static_cast<uint32_t>(std::atoi(env));
  }
  if (const char* env = std::getenv("MMC_EVENT_SLOTS")) {
    if (slots == 0) slots = static_cast<uint32_t>(std::strtoul(env, nullptr, 10));
  }
  out.profile = opt.profile != 0;
  if (schedule > MMC_SCHEDULE_EVENT_ONLY) return fail(MMC_ERR_INVALID, "unknown schedule %u", schedule);
  if (schedule >= MMC_SCHEDULE_EVENT && (!continuous_energy || generation || trace))
    return fail(MMC_ERR_INVALID, "MMC_SCHEDULE_EVENT is for continuous-energy fixed-source runs");
  out.event_schedule = continuous_energy && !generation && !trace && schedule != MMC_SCHEDULE_FUSED;
  if (out.event_schedule) {
    if (slots == 0)
      slots = n_histories >= kLargeBatchHistories ? kLargeBatchEventSlots
              : n_histories < kSmallBatchHistories ? kSmallBatchEventSlots : kDefaultEventSlots;
    // worlds with fission keep a secondary deque per slot: bound its memory
    if (w->has_fission) slots = std::min<uint32_t>(slots, 1u << 18);
    out.event_slots = static_cast<uint32_t>(std::min<uint64_t>(std::max<uint64_t>(n_histories, 1), slots));
    out.event_handover = schedule == MMC_SCHEDULE_EVENT_ONLY ? 0u : kDefaultEventHandover;
    if (const char* env = std::getenv("MMC_EVENT_HANDOVER")) {
      if (schedule != MMC_SCHEDULE_EVENT_ONLY) out.event_handover = static_cast<uint32_t>(std::strtoul(env, nullptr, 10));
    }
  }
  return MMC_OK;
}

// Carves the event-split schedule's device buffers out of one allocation.
struct EventBuffers {
  EventState st{};
  EventQueues q{};
  unsigned long long* counter_replicas = nullptr;
};

int ensure_event_buffers(mmc_world* w, uint32_t n_slots, EventBuffers& out) {
  const size_t n = event_padded_slots(n_slots);  // keeps every array 128-byte aligned
  const size_t need = n * (kEventStateBytesPerSlot + 4 * sizeof(uint32_t)) + 256 +
                      kCounterReplicas * sizeof(mmc_counters);
  if (need > w->event_bytes) {
    scratch_release(w->device, kScratchEvent, w->d_event, w->event_bytes);
    w->d_event = nullptr;
    w->event_bytes = 0;
    MMC_CUDA(scratch_acquire(w->device, kScratchEvent, need, reinterpret_cast<void**>(&w->d_event), &w->event_bytes));
  }
  if (!w->h_event_counts) MMC_CUDA(cudaMallocHost(&w->h_event_counts, 8 * sizeof(unsigned int)));
  char* at = w->d_event;
  auto take = [&](auto*& ptr, size_t bytes) {
    ptr = reinterpret_cast<std::remove_reference_t<decltype(ptr)>>(at);
    at += bytes;
  };
  EventState& st = out.st;
  take(st.px, n * 8), take(st.py, n * 8), take(st.pz, n * 8);
  take(st.dx, n * 8), take(st.dy, n * 8), take(st.dz, n * 8);
  take(st.energy, n * 8), take(st.tsl_T, n * 8);
  take(st.rng, n * 4), take(st.rng_k0, n * 4), take(st.rng_k1, n * 4);
  take(st.cell, n * 4), take(st.surface, n * 4), take(st.event, n * 4);
  take(st.n_pending, n * 4), take(st.dq_head, n * 4), take(st.dq_count, n * 4), take(st.tsl_off, n * 4);
  take(out.q.alive[0], n * 4), take(out.q.alive[1], n * 4), take(out.q.tsl, n * 4), take(out.q.boundary, n * 4);
  take(out.q.count, 256);
  take(out.counter_replicas, kCounterReplicas * sizeof(mmc_counters));
  return MMC_OK;
}

// The pass loop of the event-split schedule.  One pass = one event of every live history (flight kernel + S(a,b)
// kernel).  The number of live slots is read back every few passes: it bounds the next launches' grids and ends the
// loop.  Passes over an empty queue are no-ops, so checking late is harmless.
// The world's tables (image + dense / evaluated tail, ~10 MB at the reference's shapes) are gathered from by every event
// while 200 MB of slot state stream through the same L2 twice per pass: an access-policy window marks the tables
// persisting on the stream the event kernels run on, so the streaming state does not evict them.
// MMC_L2_PERSIST=0 leaves the stream alone (development A/B).
void keep_tables_in_l2(mmc_world* w, cudaStream_t stream) {
  static const bool enabled = !(std::getenv("MMC_L2_PERSIST") && std::atoi(std::getenv("MMC_L2_PERSIST")) == 0);
  if (!enabled || !stream) return;
  int max_window = 0, max_persist = 0;
  cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, w->device);
  cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, w->device);
  const size_t bytes = static_cast<size_t>(w->blob_bytes) + w->header.dense_bytes;
  if (max_window <= 0 || max_persist <= 0 || bytes == 0) return;
  const size_t window = std::min<size_t>(bytes, static_cast<size_t>(max_window));
  if (w->l2_persist_bytes < window) {
    const size_t want = std::min<size_t>(std::max<size_t>(window, size_t{16} << 20), static_cast<size_t>(max_persist));
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) {
      cudaGetLastError();
      return;
    }
    w->l2_persist_bytes = window;
  }
  cudaStreamAttrValue attr{};
  attr.accessPolicyWindow.base_ptr = w->d_blob;
  attr.accessPolicyWindow.num_bytes = window;
  attr.accessPolicyWindow.hitRatio = 1.0f;
  attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  if (cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
}

// A caller's stream gets its access policy back once the run's launches are enqueued (the window is taken at launch,
// so the kernels already in the stream keep it); the world's own stream keeps the window between runs.
void release_l2_window(const mmc_world* w, cudaStream_t stream) {
  if (!stream || stream == w->stream) return;
  cudaStreamAttrValue attr{};  // num_bytes = 0: no window
  if (cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
}

int run_event_schedule(mmc_world* w, const Prepared& p, unsigned long long* d_scores, unsigned long long* d_square,
                       mmc_counters* d_counters) {
  EventBuffers b;
  if (int s = ensure_event_buffers(w, p.event_slots, b)) return s;
  if (int s = ensure_scratch(w, p.event_slots, p.run.secondary_capacity, p.run.pending_capacity, p.bounds.size())) return s;
  if (!p.bounds.empty())
    MMC_CUDA(cudaMemcpyAsync(w->d_bounds, p.bounds.data(), p.bounds.size() * sizeof(double), cudaMemcpyHostToDevice, p.stream));
  MMC_CUDA(cudaMemsetAsync(w->d_next, 0, sizeof(unsigned long long), p.stream));
  EventTslConfig tsl;
  MMC_CUDA(MMC_BY_RNG(p.counter_rng, configure_event_tsl, w->header.sc_arena_bytes, w->smem_optin, static_cast<uint32_t>(w->sm_count), tsl));
  MMC_CUDA(MMC_BY_RNG(p.counter_rng, launch_event_init, b.st, b.q, p.event_slots, b.counter_replicas, p.stream));
  uint32_t alive = p.event_slots, pass = 0;
  w->last_launches = 1;
  keep_tables_in_l2(w, p.stream);  // (every exit below this line goes through release_l2_window)
  // profile mode: CUDA events around every kernel of every pass (flight | S(a,b)), summed after the run
  std::vector<cudaEvent_t> marks;
  auto mark = [&]() -> cudaEvent_t {
    if (!p.profile) return nullptr;
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    marks.push_back(e);
    return e;
  };
  int status = MMC_OK;
  unsigned long long claimed = 0;  // histories started so far, as of the last read
  while (alive && status == MMC_OK) {
    // passes between two reads of the queue lengths (a stream synchronisation: the GPU idles for a launch latency).
    // While more than two refills of every slot are still unstarted nothing can end, so the reads are rarer.
    const int batch = claimed + 2ull * p.event_slots < p.run.n_histories ? 32 : 8;
    for (int k = 0; k < batch && status == MMC_OK; k++, pass++) {
      cudaEvent_t inner[2] = {nullptr, nullptr};
      if (cudaEvent_t e = mark()) cudaEventRecord(e, p.stream);
      if (p.profile) inner[0] = mark(), inner[1] = mark();
      const cudaError_t err = MMC_BY_RNG(p.counter_rng, launch_event_pass,
          w->d_blob, w->header, p.run, w->d_bounds, b.st, b.q, pass, alive, w->d_sites, w->d_pending, w->d_next, d_scores, d_square,
          b.counter_replicas, tsl, p.stream, p.profile && inner[0] && inner[1] ? inner : nullptr);
      if (cudaEvent_t e = mark()) cudaEventRecord(e, p.stream);
      if (err != cudaSuccess) status = fail(MMC_ERR_CUDA, "launch_event_pass: %s", cudaGetErrorString(err));
    }
    w->last_launches += static_cast<uint64_t>(event_kernels_per_pass()) * batch;
    cudaError_t err = cudaMemcpyAsync(w->h_event_counts, b.q.count, 4 * sizeof(unsigned int), cudaMemcpyDeviceToHost, p.stream);
    if (err == cudaSuccess)
      err = cudaMemcpyAsync(w->h_event_counts + 4, w->d_next, sizeof(unsigned long long), cudaMemcpyDeviceToHost, p.stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(p.stream);
    if (err != cudaSuccess && status == MMC_OK) status = fail(MMC_ERR_CUDA, "event-split pass loop: %s", cudaGetErrorString(err));
    alive = w->h_event_counts[pass & 1u];
    std::memcpy(&claimed, w->h_event_counts + 4, sizeof(claimed));
    if (status == MMC_OK && alive && alive <= p.event_handover && claimed >= p.run.n_histories) {
      // The drain: every history has started and few are still alive.  A pass now costs its two launches' latency,
      // not their work, so the survivors are handed to the fused kernel, one thread per slot, and run to their end.
      ResumeIO resume;
      resume.slots = b.q.alive[pass & 1u];
      resume.n = b.q.count + (pass & 1u);
      resume.st = b.st;
      LaunchConfig cfg;
      cfg.blocks = static_cast<int>((alive + kThreadsPerBlock - 1) / kThreadsPerBlock);
      RunSpec run = p.run;
      run.chunk = 32;
      err = MMC_BY_RNG(p.counter_rng, launch_fixed_source, cfg, w->d_blob, run, w->d_bounds, w->d_sites, w->d_pending, w->d_next,
                       d_scores, d_square, d_counters, nullptr, p.stream, &resume);
      if (err != cudaSuccess) status = fail(MMC_ERR_CUDA, "hand-over to the fused kernel: %s", cudaGetErrorString(err));
      w->last_launches += 1;
      alive = 0;
    }
  }
  w->last_flight_ms = w->last_tsl_ms = w->last_boundary_ms = 0;
  for (size_t k = 0; k + 3 < marks.size() && status == MMC_OK; k += 4) {  // before | after flight | after boundary | after S(a,b)
    float a = 0, bd = 0, c = 0;
    if (cudaEventElapsedTime(&a, marks[k], marks[k + 1]) == cudaSuccess) w->last_flight_ms += a;
    if (cudaEventElapsedTime(&bd, marks[k + 1], marks[k + 2]) == cudaSuccess) w->last_boundary_ms += bd;
    if (cudaEventElapsedTime(&c, marks[k + 2], marks[k + 3]) == cudaSuccess) w->last_tsl_ms += c;
  }
  for (cudaEvent_t e : marks) cudaEventDestroy(e);
  if (status != MMC_OK) {
    release_l2_window(w, p.stream);
    return status;
  }
  const cudaError_t finish = MMC_BY_RNG(p.counter_rng, launch_event_finish, b.counter_replicas, d_counters, p.stream);
  release_l2_window(w, p.stream);
  MMC_CUDA(finish);
  w->last_launches += 1;
  return MMC_OK;
}

int status_from_counters(const mmc_counters& c) {
  if (c.n_lost)
    return fail(MMC_ERR_LOST_PARTICLE,
                "%llu particle(s) do not belong to any Cell. Please check that all space is either assigned a material or void.",
                static_cast<unsigned long long>(c.n_lost));
  if (c.n_physics_errors)
    return fail(MMC_ERR_PHYSICS, "%llu event(s) reached a branch the reference asserts unreachable",
                static_cast<unsigned long long>(c.n_physics_errors));
  if (c.n_capacity_overflow)
    return fail(MMC_ERR_CAPACITY,
                "%llu per-history capacity overflow(s): raise secondary_capacity / pending_capacity in mmc_run_options",
                static_cast<unsigned long long>(c.n_capacity_overflow));
  return MMC_OK;
}

}  // namespace

extern "C" {

int mmc_abi_version(void) { return MMC_ABI_VERSION; }

size_t mmc_last_error(char* buf, size_t cap) {
  if (buf && cap) {
    const size_t n = std::min(cap - 1, g_error.size());
    std::memcpy(buf, g_error.data(), n);
    buf[n] = 0;
  }
  return g_error.size();
}

int mmc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

uint64_t mmc_world_bytes(const mmc_world* world) { return world ? world->blob_bytes : 0; }

void* mmc_world_stream(const mmc_world* world) { return world ? world->stream : nullptr; }

uint64_t mmc_world_last_launches(const mmc_world* world) { return world ? world->last_launches : 0; }

void mmc_world_last_kernel_ms(const mmc_world* world, double* flight_ms, double* tsl_ms) {
  if (flight_ms) *flight_ms = world ? world->last_flight_ms : 0;
  if (tsl_ms) *tsl_ms = world ? world->last_tsl_ms : 0;
}

double mmc_world_last_boundary_ms(const mmc_world* world) { return world ? world->last_boundary_ms : 0; }

uint64_t mmc_estimator_size(const mmc_estimator_desc* e) {
  if (!e) return 0;
  auto n = [](const mmc_bins_desc& b) -> uint64_t { return b.kind == MMC_BINS_NONE ? 1 : b.n_bins; };
  return n(e->cosine) * n(e->energy);
}

namespace {
// Flattens a validated world description into the device image (world_blob.h).
// `jobs`: the dense reconstruction tables to expand on the device after the image is uploaded (world_blob.h
// DenseJob); their space is the device-only tail behind the image.
int build_world_blob(const mmc_world_desc* d, BlobBuilder& b, WorldHeader& header_out, bool& has_fission_out,
                     std::vector<DenseJob>& jobs, std::vector<EvalJob>& eval_jobs) {
  jobs.clear();
  eval_jobs.clear();
  // dense tables: offsets are known relative to the tail while the image is still growing; the fields that hold them
  // are patched once the image's size is final
  struct DensePatch { size_t field_at; uint32_t tail_offset; };
  std::vector<DensePatch> dense_patches;
  uint64_t dense_total = 0, dense_budget = 512ull << 20;
  if (const char* mb = std::getenv("MMC_TSL_DENSE_MB")) dense_budget = std::strtoull(mb, nullptr, 10) << 20;
  bool dense_complete = true;
  int dense_tables = 0;
  // reserves n doubles in the tail; false when over the budget
  auto reserve_tail = [&](uint64_t n_doubles, uint32_t& tail_offset) {
    const uint64_t bytes = (n_doubles * 8 + 15) & ~15ull;
    if (n_doubles == 0 || dense_total + bytes > dense_budget) return false;
    tail_offset = static_cast<uint32_t>(dense_total);
    dense_total += bytes;
    return true;
  };
  auto reserve_dense = [&](uint64_t n_doubles, uint32_t& tail_offset) {
    if (!reserve_tail(n_doubles, tail_offset)) {
      dense_complete = false;
      return false;
    }
    dense_tables++;
    return true;
  };
  const int G = d->n_groups;
  const int nnz_cell = d->cell_surface_begin[d->n_cells];
  const int nnz_mat = d->n_materials ? d->material_nuclide_begin[d->n_materials] : 0;
  std::vector<int32_t> packed(nnz_cell);
  for (int k = 0; k < nnz_cell; k++) packed[k] = (d->cell_surface_index[k] << 1) | (d->cell_surface_sense[k] ? 1 : 0);
  std::vector<int32_t> field_kind(d->n_cells, MMC_FIELD_CONSTANT);
  std::vector<double> field_param(static_cast<size_t>(d->n_cells) * 6, 0.0);
  if (d->cell_field_kind && d->cell_field_param) {
    std::copy(d->cell_field_kind, d->cell_field_kind + d->n_cells, field_kind.begin());
    std::copy(d->cell_field_param, d->cell_field_param + static_cast<size_t>(d->n_cells) * 6, field_param.begin());
  }
  WorldHeader h{};
  h.n_surfaces = d->n_surfaces;
  h.n_cells = d->n_cells;
  h.n_materials = d->n_materials;
  h.n_nuclides = d->n_nuclides;
  h.n_groups = G;
  h.off_surface_type = b.add(d->surface_type, d->n_surfaces);
  h.off_surface_param = b.add(d->surface_param, static_cast<size_t>(d->n_surfaces) * 4);
  h.off_cell_material = b.add(d->cell_material, d->n_cells);
  h.off_cell_surf_begin = b.add(d->cell_surface_begin, d->n_cells + 1);
  h.off_cell_surf = b.add(packed.data(), packed.size());
  std::vector<SurfaceRecord> records(nnz_cell);
  for (int k = 0; k < nnz_cell; k++) {
    const int s_index = d->cell_surface_index[k];
    SurfaceRecord& r = records[k];
    r = SurfaceRecord{};
    for (int j = 0; j < 4; j++) r.prm[j] = d->surface_param[4 * s_index + j];
    r.type = d->surface_type[s_index];
    r.index_sense = packed[k];
  }
  h.off_cell_surf_rec = b.add(records.data(), records.size());
  if (d->n_surfaces <= 64) {
    // Cell::Contains as one mask compare (Cell.cpp:27-35: every surface of the cell must have Contains(p) == sense)
    std::vector<uint64_t> masks(static_cast<size_t>(d->n_cells) * 2, 0);
    for (int c = 0; c < d->n_cells; c++) {
      uint64_t mask = 0, want = 0;
      bool contradictory = false;
      for (int k = d->cell_surface_begin[c]; k < d->cell_surface_begin[c + 1]; k++) {
        const uint64_t bit = 1ull << d->cell_surface_index[k];
        const uint64_t sense = d->cell_surface_sense[k] ? bit : 0;
        if ((mask & bit) && (want & bit) != sense) contradictory = true;  // the same surface with both senses
        mask |= bit;
        want |= sense;
      }
      if (contradictory) mask = 0, want = ~0ull;  // never equal: such a cell contains nothing
      masks[2 * c] = mask;
      masks[2 * c + 1] = want;
    }
    h.off_cell_mask = b.add(masks.data(), masks.size());
  }
  h.off_cell_field_kind = b.add(field_kind.data(), field_kind.size());
  h.off_cell_field_param = b.add(field_param.data(), field_param.size());
  // distinct constant temperatures of the material cells, in order of first appearance: the temperatures the S(a,b)
  // partitions are evaluated at once per upload (TslPartition::off_eval)
  std::vector<double> eval_T;
  std::vector<int32_t> cell_eval_slot(d->n_cells, -1);
  if (G == 0 && dense_budget > 0) {
    for (int c = 0; c < d->n_cells; c++) {
      if (d->cell_material[c] < 0 || field_kind[c] != MMC_FIELD_CONSTANT) continue;
      const double T = field_param[static_cast<size_t>(c) * 6];
      size_t k = 0;
      while (k < eval_T.size() && std::memcmp(&eval_T[k], &T, sizeof(double)) != 0) k++;
      if (k == eval_T.size()) {
        if (eval_T.size() == static_cast<size_t>(kMaxEvalT)) continue;
        eval_T.push_back(T);
      }
      cell_eval_slot[c] = static_cast<int32_t>(k);
    }
  }
  h.off_cell_eval_slot = b.add(cell_eval_slot.data(), cell_eval_slot.size());
  h.n_eval_T = static_cast<int32_t>(eval_T.size());
  const int32_t zero_begin[1] = {0};
  h.off_mat_aden = b.add(d->material_aden, d->n_materials);
  h.off_mat_nuc_begin = d->n_materials ? b.add(d->material_nuclide_begin, d->n_materials + 1) : b.add(zero_begin, 1);
  h.off_mat_nuc_index = b.add(d->material_nuclide_index, nnz_mat);
  h.off_mat_nuc_afrac = b.add(d->material_nuclide_afrac, nnz_mat);
  const size_t ng = static_cast<size_t>(d->n_nuclides) * G;
  h.off_mg_mask = b.add(d->mg_reaction_mask, G > 0 ? d->n_nuclides : 0);
  h.off_mg_total = b.add(d->mg_total, ng);
  h.off_mg_capture = b.add(d->mg_capture, ng);
  h.off_mg_scatter = b.add(d->mg_scatter, ng);
  h.off_mg_fission = b.add(d->mg_fission, ng);
  h.off_mg_nubar = b.add(d->mg_nubar, ng);
  h.off_mg_scatter_probs = b.add(d->mg_scatter_probs, ng * G);
  h.off_mg_chi = b.add(d->mg_chi, ng * G);
  bool has_fission = false;
  for (int n = 0; G > 0 && n < d->n_nuclides; n++) has_fission = has_fission || (d->mg_reaction_mask[n] & MMC_REACTION_FISSION);
  if (G == 0 && d->n_nuclides > 0) {
    // SearchHint of a sorted, non-negative axis (world_blob.h); 0 for short or unsorted arrays
    auto hint = [&b](const double* x, uint64_t n) -> uint32_t {
      if (n < 8) return 0;
      for (uint64_t i = 0; i < n; i++)
        if (!(x[i] >= 0) || (i && x[i] < x[i - 1]) || !std::isfinite(x[i])) return 0;
      uint64_t first_positive = 0;
      while (first_positive < n && x[first_positive] == 0) first_positive++;
      if (first_positive == n) return 0;
      auto bits = [](double v) { int64_t u; std::memcpy(&u, &v, 8); return u; };
      uint32_t shift = 52 - 6;  // at most 64 buckets per octave, fewer until the index fits 512 buckets
      int64_t lo = 0, hi = 0;
      for (;; shift++) {
        lo = bits(x[first_positive]) >> shift;
        hi = bits(x[n - 1]) >> shift;
        if (hi - lo + 2 <= 512 || shift == 62) break;
      }
      SearchHint sh{};
      sh.first_bucket = lo - 1;
      sh.shift = shift;
      sh.n_buckets = static_cast<uint32_t>(hi - lo + 2);
      // cum[k] = number of elements whose bucket (clamped like the device does) is below k
      std::vector<uint32_t> cum(sh.n_buckets + 1, 0);
      for (uint64_t i = 0; i < n; i++) {
        int64_t k = (bits(x[i]) >> shift) - sh.first_bucket;
        k = std::min<int64_t>(std::max<int64_t>(k, 0), sh.n_buckets - 1);
        cum[k + 1]++;
      }
      for (uint32_t k = 0; k < sh.n_buckets; k++) cum[k + 1] += cum[k];
      // stored as pairs {cum[k], cum[k + 1]}: a bucket's range is one 8-byte load
      std::vector<uint32_t> pairs(2 * static_cast<size_t>(sh.n_buckets));
      for (uint32_t k = 0; k < sh.n_buckets; k++) pairs[2 * k] = cum[k], pairs[2 * k + 1] = cum[k + 1];
      std::vector<char> image(sizeof(SearchHint) + pairs.size() * sizeof(uint32_t));
      std::memcpy(image.data(), &sh, sizeof(SearchHint));
      std::memcpy(image.data() + sizeof(SearchHint), pairs.data(), pairs.size() * sizeof(uint32_t));
      return b.add(image.data(), image.size());
    };
    auto table = [&b, &hint](const mmc_table1d& t) {
      Table1D out{};
      out.n = static_cast<uint32_t>(t.n);
      if (t.n) {
        out.off_x = b.add(t.x, t.n);
        out.off_y = b.add(t.y, t.n);
        out.off_hint = hint(t.x, t.n);
      }
      return out;
    };
    // the scaled CDF modes of every partition first, back to back (WorldHeader::off_sc_arena):
    // S[r] * CDF_modes[cdf][r] is the first product of Evaluate's left-to-right expression
    // (ThermalScattering.cpp:199-204,241-246), one IEEE multiply (this file is built with -ffp-contract=off)
    std::vector<uint32_t> sc_offsets;
    size_t sc_next = 0;
    {
      b.pad();
      h.off_sc_arena = static_cast<uint32_t>(b.bytes.size());
      auto add_scaled = [&](const mmc_tsl_partition* p, int n) {
        for (int i = 0; i < n; i++) {
          const mmc_tsl_partition& q = p[i];
          std::vector<double> scaled(q.n_cdf * q.rank);
          for (uint64_t c = 0; c < q.n_cdf; c++)
            for (uint64_t r = 0; r < q.rank; r++) scaled[c * q.rank + r] = q.singular_values[r] * q.cdf_modes[c * q.rank + r];
          sc_offsets.push_back(b.add(scaled.data(), scaled.size()));
        }
      };
      for (int n = 0; n < d->n_nuclides; n++)
        for (int r = 0; r < d->ce->nuclides[n].n_reactions; r++)
          if (const mmc_tsl_desc* t = d->ce->nuclides[n].reactions[r].tsl) {
            add_scaled(t->beta_partitions, t->n_beta_partitions);
            add_scaled(t->alpha_partitions, t->n_alpha_partitions);
          }
      b.pad();
      h.sc_arena_bytes = static_cast<uint32_t>(b.bytes.size()) - h.off_sc_arena;
    }
    auto partitions = [&](const mmc_tsl_partition* p, int n, std::vector<double>& concatenated) {
      std::vector<TslPartition> out(n);
      std::vector<uint32_t> tail(n, 0), tail_eval(n, 0);
      std::vector<char> expanded(n, 0), evaluated(n, 0);
      std::vector<size_t> job_of(n, 0);  // index of partition i's EvalJob
      for (int i = 0; i < n; i++) {
        const mmc_tsl_partition& q = p[i];
        TslPartition& o = out[i];
        o.n_cdf = static_cast<uint32_t>(q.n_cdf);
        o.n_grid = static_cast<uint32_t>(q.n_grid);
        o.n_T = static_cast<uint32_t>(q.n_temperature);
        o.rank = static_cast<uint32_t>(q.rank);
        o.off_cdf = b.add(q.cdf, q.n_cdf);
        o.off_cdf_hint = hint(q.cdf, q.n_cdf);
        {
          std::vector<double> pairs(2 * (q.n_cdf + 1));
          for (uint64_t c = 0; c <= q.n_cdf; c++) {
            pairs[2 * c] = c != 0 ? q.cdf[c - 1] : 0.0;
            pairs[2 * c + 1] = c != q.n_cdf ? q.cdf[c] : 1.0;
          }
          o.off_cdf_pairs = b.add(pairs.data(), pairs.size());
          bool usable = q.n_cdf <= 255;
          for (uint64_t c = 0; c < q.n_cdf && usable; c++)
            usable = q.cdf[c] >= 0.0 && q.cdf[c] <= 1.0 && (c == 0 || q.cdf[c - 1] <= q.cdf[c]);
          o.off_cdf_lut = 0;
          if (usable) {
            std::vector<uint8_t> lut(kCdfLut);
            for (uint32_t k = 0; k < kCdfLut; k++)
              lut[k] = static_cast<uint8_t>(std::upper_bound(q.cdf, q.cdf + q.n_cdf, static_cast<double>(k) / kCdfLut) - q.cdf);
            o.off_cdf_lut = b.add(lut.data(), lut.size());
          }
        }
        o.off_T = b.add(q.temperature, q.n_temperature);
        o.off_T_hint = hint(q.temperature, q.n_temperature);
        o.off_scaled_cdf_modes = sc_offsets[sc_next++];  // in the arena, same traversal order
        o.off_modes = b.add(q.grid_T_modes, q.n_grid * q.n_temperature * q.rank);
        o.grid_begin = static_cast<uint32_t>(concatenated.size());
        concatenated.insert(concatenated.end(), q.grid, q.grid + q.n_grid);
        o.off_dense = 0;
        if (reserve_dense(q.n_grid * q.n_cdf * q.n_temperature, tail[i])) {
          expanded[i] = 1;
          jobs.push_back(DenseJob{o.off_scaled_cdf_modes, o.off_modes, tail[i], o.n_grid, o.n_cdf, o.n_T, o.rank, 1u});
        }
        o.off_eval = 0;
        o.eval_sorted = 0;
        if (!eval_T.empty() && q.n_temperature <= 255 &&
            reserve_tail(eval_T.size() * q.n_grid * q.n_cdf, tail_eval[i])) {
          evaluated[i] = 1;
          o.eval_sorted = std::getenv("MMC_TSL_SORTED_SEARCH") && std::atoi(std::getenv("MMC_TSL_SORTED_SEARCH")) == 0 ? 0u : 1u;
          EvalJob job{};
          job.off_a = o.off_scaled_cdf_modes;
          job.off_m = o.off_modes;
          job.off_out = tail_eval[i];
          job.n_grid = o.n_grid, job.n_cdf = o.n_cdf, job.n_T = o.n_T, job.rank = o.rank;
          job.n_slots = static_cast<uint32_t>(eval_T.size());
          for (size_t s = 0; s < eval_T.size(); s++) {
            // "Find index of Temperature above and below target Temperature", ThermalScattering.cpp:188-196,230-238
            const double* Ts = q.temperature;
            const size_t candidate = static_cast<size_t>(std::upper_bound(Ts, Ts + q.n_temperature, eval_T[s]) - Ts);
            const size_t T_hi_i = candidate == q.n_temperature ? candidate - 1 : candidate;
            const size_t T_lo_i = T_hi_i == 0 ? T_hi_i : T_hi_i - 1;
            job.t_hi[s] = static_cast<uint8_t>(T_hi_i);
            job.t_lo[s] = static_cast<uint8_t>(T_lo_i);
            job.dT[s] = Ts[T_hi_i] - Ts[T_lo_i];
            job.tT[s] = eval_T[s] - Ts[T_lo_i];
          }
          job_of[i] = eval_jobs.size();
          eval_jobs.push_back(job);
        }
      }
      const uint32_t at = b.add(out.data(), out.size());
      for (int i = 0; i < n; i++) {
        if (expanded[i]) dense_patches.push_back({at + i * sizeof(TslPartition) + offsetof(TslPartition, off_dense), tail[i]});
        if (evaluated[i]) {
          dense_patches.push_back({at + i * sizeof(TslPartition) + offsetof(TslPartition, off_eval), tail_eval[i]});
          // the job of this partition learns where the partition's eval_sorted flag lives
          eval_jobs[job_of[i]].off_sorted_flag =
              static_cast<uint32_t>(at + i * sizeof(TslPartition) + offsetof(TslPartition, eval_sorted));
        }
      }
      return at;
    };
    std::vector<CeNuclide> nuclides(d->n_nuclides);
    for (int n = 0; n < d->n_nuclides; n++) {
      const mmc_ce_nuclide& c = d->ce->nuclides[n];
      CeNuclide& o = nuclides[n];
      o = CeNuclide{};
      o.awr = c.awr;
      o.total = table(c.total);
      o.total_temperature = c.total_temperature;
      o.n_reactions = c.n_reactions;
      for (int r = 0; r < c.n_reactions; r++) {
        const mmc_ce_reaction& x = c.reactions[r];
        CeReaction& xr = o.reactions[r];
        xr.kind = x.kind;
        xr.xs = table(x.xs);
        xr.temperature = x.temperature;
        if (x.has_nubar) xr.nubar = table(x.nubar);
        has_fission = has_fission || x.kind == MMC_REACTION_FISSION;
        if (x.tsl) {
          const mmc_tsl_desc& t = *x.tsl;
          TslTable tt{};
          tt.majorant = table(t.majorant);
          tt.n_E = static_cast<uint32_t>(t.n_energy);
          tt.n_T = static_cast<uint32_t>(t.n_temperature);
          tt.rank = static_cast<uint32_t>(t.rank);
          tt.off_E = b.add(t.energy, t.n_energy);
          tt.off_E_hint = hint(t.energy, t.n_energy);
          tt.off_T = b.add(t.temperature, t.n_temperature);
          tt.off_T_hint = hint(t.temperature, t.n_temperature);
          // S[r] * scatter_xs_E[E][r]: the first product of EvaluateInelastic (ThermalScattering.cpp:264-267)
          std::vector<double> xs_SE(t.n_energy * t.rank);
          for (uint64_t e = 0; e < t.n_energy; e++)
            for (uint64_t r = 0; r < t.rank; r++) xs_SE[e * t.rank + r] = t.xs_S[r] * t.xs_E[e * t.rank + r];
          tt.off_xs_SE = b.add(xs_SE.data(), xs_SE.size());
          tt.off_xs_T = b.add(t.xs_T, t.n_temperature * t.rank);
          std::vector<double> Es, betas;
          const size_t first_eval_job = eval_jobs.size();
          tt.n_beta_partitions = t.n_beta_partitions;
          tt.off_beta_partitions = partitions(t.beta_partitions, t.n_beta_partitions, Es);
          tt.n_alpha_partitions = t.n_alpha_partitions;
          tt.off_alpha_partitions = partitions(t.alpha_partitions, t.n_alpha_partitions, betas);
          // every partition evaluated (and, checked on the device, sorted): collisions at an evaluated temperature
          // sample with ce::tsl_sample_direct
          const bool sorted_search = !(std::getenv("MMC_TSL_SORTED_SEARCH") && std::atoi(std::getenv("MMC_TSL_SORTED_SEARCH")) == 0);
          tt.direct = sorted_search && !eval_T.empty() &&
                              eval_jobs.size() - first_eval_job == static_cast<size_t>(t.n_beta_partitions + t.n_alpha_partitions)
                          ? 1u : 0u;
          tt.n_Es = static_cast<uint32_t>(Es.size());
          tt.off_Es = b.add(Es.data(), Es.size());
          tt.off_Es_hint = hint(Es.data(), Es.size());
          tt.n_betas = static_cast<uint32_t>(betas.size());
          tt.off_betas = b.add(betas.data(), betas.size());
          tt.off_betas_hint = hint(betas.data(), betas.size());
          tt.beta_cutoff = t.beta_cutoff;
          tt.alpha_cutoff = t.alpha_cutoff;
          tt.awr = t.awr;
          tt.cutoff_energy = t.energy[t.n_energy - 1];  // ThermalScattering.hpp:125-126
          if (!eval_T.empty()) {
            // "Find index of Temperature above and below" + r_T of GetTotal at every evaluated temperature
            // (ThermalScattering.cpp:126-135,152-155; one IEEE subtraction pair and division, -ffp-contract=off)
            std::vector<TslEvalBracket> brackets(eval_T.size());
            for (size_t s = 0; s < eval_T.size(); s++) {
              const double* Ts = t.temperature;
              const size_t candidate = static_cast<size_t>(std::upper_bound(Ts, Ts + t.n_temperature, eval_T[s]) - Ts);
              const bool above_max = candidate == t.n_temperature;
              const size_t hi = above_max ? candidate - 1 : candidate;
              const bool below_min = hi == 0;
              const size_t lo = below_min ? hi : hi - 1;
              brackets[s].lo = static_cast<uint32_t>(lo);
              brackets[s].hi = static_cast<uint32_t>(hi);
              brackets[s].r_T = below_min ? 1.0 : above_max ? 0.0 : (eval_T[s] - Ts[lo]) / (Ts[hi] - Ts[lo]);
            }
            tt.off_eval_bracket = b.add(brackets.data(), brackets.size());
          }
          uint32_t xs_tail = 0;
          const bool xs_expanded = reserve_dense(t.n_energy * t.n_temperature, xs_tail);
          if (xs_expanded) jobs.push_back(DenseJob{tt.off_xs_SE, tt.off_xs_T, xs_tail, 1u, tt.n_E, tt.n_T, tt.rank, 0u});
          xr.off_tsl = b.add(&tt, 1);
          for (size_t k = first_eval_job; k < eval_jobs.size(); k++)
            eval_jobs[k].off_direct_flag = xr.off_tsl + static_cast<uint32_t>(offsetof(TslTable, direct));
          if (xs_expanded) dense_patches.push_back({xr.off_tsl + offsetof(TslTable, off_xs_dense), xs_tail});
        }
      }
    }
    h.off_ce_nuclides = b.add(nuclides.data(), nuclides.size());
  }
  b.pad();
  if (b.bytes.size() + dense_total > 0xfffffff0ull) return fail(MMC_ERR_INVALID, "world tables exceed 4 GiB");
  h.total_bytes = static_cast<uint32_t>(b.bytes.size());
  h.dense_bytes = static_cast<uint32_t>(dense_total);
  h.tsl_all_dense = dense_tables > 0 && dense_complete;
  // candidates for the direct-only S(a,b) kernel: every material cell at an evaluated temperature, every table direct
  // (the device still has to confirm that the evaluated rows are sorted: confirm_direct)
  {
    bool all = G == 0 && !eval_jobs.empty();
    for (int c = 0; c < d->n_cells && all; c++) all = d->cell_material[c] < 0 || cell_eval_slot[c] >= 0;
    for (const EvalJob& job : eval_jobs) {
      uint32_t flag = 0;
      std::memcpy(&flag, b.bytes.data() + job.off_direct_flag, sizeof(flag));
      all = all && flag == 1u;
    }
    h.tsl_all_direct = all ? 1u : 0u;
  }
  for (const DensePatch& patch : dense_patches) {
    const uint32_t absolute = h.total_bytes + patch.tail_offset;
    std::memcpy(b.bytes.data() + patch.field_at, &absolute, sizeof(uint32_t));
  }
  for (DenseJob& job : jobs) job.off_out += h.total_bytes;
  for (EvalJob& job : eval_jobs) job.off_out += h.total_bytes;
  b.header() = h;
  header_out = h;
  has_fission_out = has_fission;
  return MMC_OK;
}

// After the evaluated rows were checked on the device (check_rows_sorted_kernel clears TslTable::direct of a table with
// an unsorted row): the host's copy of the header keeps tsl_all_direct only if every table is still direct.
cudaError_t confirm_direct(const char* d_blob, const std::vector<EvalJob>& eval_jobs, WorldHeader& h) {
  if (!h.tsl_all_direct) return cudaSuccess;
  uint32_t last = 0;
  for (const EvalJob& job : eval_jobs) {
    if (job.off_direct_flag == last) continue;
    last = job.off_direct_flag;
    uint32_t flag = 0;
    const cudaError_t e = cudaMemcpy(&flag, d_blob + job.off_direct_flag, sizeof(flag), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return e;
    if (flag != 1u) h.tsl_all_direct = 0;
  }
  return cudaSuccess;
}
}  // namespace

int mmc_world_create(const mmc_world_desc* d, int device, mmc_world** out) {
  if (!out) return fail(MMC_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (int s = validate_world(d)) return s;
  if (mmc_device_count() < 1)
    return fail(MMC_ERR_NO_DEVICE, "no CUDA device visible: minimc_b200 has no CPU transport path");
  if (device < 0) MMC_CUDA(cudaGetDevice(&device));
  MMC_CUDA(cudaSetDevice(device));
  BlobBuilder b;
  WorldHeader h{};
  bool has_fission = false;
  std::vector<DenseJob> dense_jobs;
  std::vector<EvalJob> eval_jobs;
  if (int s = build_world_blob(d, b, h, has_fission, dense_jobs, eval_jobs)) return s;

  auto* w = new mmc_world;
  w->device = device;
  w->header = h;
  w->blob_bytes = h.total_bytes;
  w->has_fission = has_fission;
  cudaDeviceProp prop{};
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e == cudaSuccess) e = cudaMalloc(&w->d_blob, static_cast<size_t>(w->blob_bytes) + h.dense_bytes);
  if (e == cudaSuccess) e = cudaMallocHost(&w->h_blob, w->blob_bytes);  // pinned staging copy of the image
  if (e == cudaSuccess) std::memcpy(w->h_blob, b.bytes.data(), w->blob_bytes);
  if (e == cudaSuccess) e = cudaMemcpy(w->d_blob, w->h_blob, w->blob_bytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&w->d_next, sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = launch_expand_dense(w->d_blob, dense_jobs.data(), dense_jobs.size(), w->stream);
  if (e == cudaSuccess) e = launch_evaluate_rows(w->d_blob, eval_jobs.data(), eval_jobs.size(), w->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(w->stream);
  if (e == cudaSuccess) e = confirm_direct(w->d_blob, eval_jobs, w->header);
  if (e != cudaSuccess) {
    mmc_world_destroy(w);
    return fail(MMC_ERR_CUDA, "mmc_world_create: %s", cudaGetErrorString(e));
  }
  w->sm_count = prop.multiProcessorCount;
  w->smem_optin = prop.sharedMemPerBlockOptin;
  *out = w;
  return MMC_OK;
}

int mmc_world_update(mmc_world* w, const mmc_world_desc* d) {
  if (!w) return fail(MMC_ERR_INVALID, "world handle is NULL");
  if (int s = validate_world(d)) return s;
  BlobBuilder b;
  WorldHeader h{};
  bool has_fission = false;
  std::vector<DenseJob> dense_jobs;
  std::vector<EvalJob> eval_jobs;
  if (int s = build_world_blob(d, b, h, has_fission, dense_jobs, eval_jobs)) return s;
  if (h.total_bytes != w->blob_bytes || h.dense_bytes != w->header.dense_bytes || has_fission != w->has_fission ||
      h.n_groups != w->header.n_groups)
    return fail(MMC_ERR_INVALID, "mmc_world_update: the new tables have a different shape (%u bytes, world holds %u): "
                "create a new world", h.total_bytes, w->blob_bytes);
  MMC_CUDA(cudaSetDevice(w->device));
  std::memcpy(w->h_blob, b.bytes.data(), w->blob_bytes);
  MMC_CUDA(cudaMemcpyAsync(w->d_blob, w->h_blob, w->blob_bytes, cudaMemcpyHostToDevice, w->stream));
  MMC_CUDA(launch_expand_dense(w->d_blob, dense_jobs.data(), dense_jobs.size(), w->stream));
  MMC_CUDA(launch_evaluate_rows(w->d_blob, eval_jobs.data(), eval_jobs.size(), w->stream));
  MMC_CUDA(cudaStreamSynchronize(w->stream));
  MMC_CUDA(confirm_direct(w->d_blob, eval_jobs, h));
  w->header = h;
  return MMC_OK;
}

void mmc_world_destroy(mmc_world* w) {
  if (!w) return;
  cudaSetDevice(w->device);
  cudaDeviceSynchronize();  // work in flight (on any stream a caller passed) may still use the scratch parked below
  if (w->stream) cudaStreamDestroy(w->stream);
  cudaFree(w->d_blob);
  if (w->h_blob) cudaFreeHost(w->h_blob);
  cudaFree(w->d_tally);
  if (w->h_tally) cudaFreeHost(w->h_tally);
  scratch_release(w->device, kScratchSites, w->d_sites, w->sites_bytes);
  scratch_release(w->device, kScratchPending, w->d_pending, w->pending_bytes);
  cudaFree(w->d_bounds);
  cudaFree(w->d_next);
  cudaFree(w->d_sens_pending);
  cudaFree(w->d_sens);
  if (w->h_sens) cudaFreeHost(w->h_sens);
  scratch_release(w->device, kScratchEvent, w->d_event, w->event_bytes);
  if (w->h_event_counts) cudaFreeHost(w->h_event_counts);
  cudaFree(w->d_unordered);
  cudaFree(w->d_child_count);
  cudaFree(w->d_child_start);
  cudaFree(w->d_block_sums);
  delete w;
}

int mmc_fixed_source_run_device(
    const mmc_world* world, const mmc_source_desc* source, const mmc_estimator_desc* estimators, int32_t n_estimators,
    uint64_t seed0, uint64_t first_history, uint64_t n_histories, const mmc_run_options* options, uint64_t* d_scores,
    uint64_t* d_square_scores, mmc_counters* d_counters) {
  auto* w = const_cast<mmc_world*>(world);
  if (!w) return fail(MMC_ERR_INVALID, "world handle is NULL");
  std::lock_guard<std::recursive_mutex> run_lock(w->run_mutex);
  Prepared p;
  if (int s = prepare_run(w, source, estimators, n_estimators, seed0, first_history, n_histories, options, false, p)) return s;
  if (!d_counters) return fail(MMC_ERR_INVALID, "d_counters is NULL");
  if (p.run.total_bins && (!d_scores || !d_square_scores)) return fail(MMC_ERR_INVALID, "tally buffers are NULL");
  MMC_CUDA(cudaSetDevice(w->device));
  if (n_histories == 0) return MMC_OK;
  if (p.event_schedule)
    return run_event_schedule(w, p, reinterpret_cast<unsigned long long*>(d_scores),
                              reinterpret_cast<unsigned long long*>(d_square_scores), d_counters);
  w->last_launches = 1;
  const size_t threads = static_cast<size_t>(p.cfg.blocks) * kThreadsPerBlock;
  if (int s = ensure_scratch(w, threads, p.run.secondary_capacity, p.run.pending_capacity, p.bounds.size())) return s;
  if (!p.bounds.empty())
    MMC_CUDA(cudaMemcpyAsync(w->d_bounds, p.bounds.data(), p.bounds.size() * sizeof(double), cudaMemcpyHostToDevice, p.stream));
  MMC_CUDA(cudaMemsetAsync(w->d_next, 0, sizeof(unsigned long long), p.stream));
  MMC_CUDA(MMC_BY_RNG(p.counter_rng, launch_fixed_source,
      p.cfg, w->d_blob, p.run, w->d_bounds, w->d_sites, w->d_pending, w->d_next,
      reinterpret_cast<unsigned long long*>(d_scores), reinterpret_cast<unsigned long long*>(d_square_scores), d_counters,
      nullptr, p.stream));
  return MMC_OK;
}

static_assert(sizeof(mmc_site) == sizeof(BankSite), "mmc_site and BankSite must have the same layout");

int mmc_source_bank_sample(const mmc_world* world, const mmc_source_desc* source, uint64_t seed0, uint64_t first_index,
                           uint64_t n, const mmc_run_options* options, mmc_site* d_bank) {
  auto* w = const_cast<mmc_world*>(world);
  if (!w) return fail(MMC_ERR_INVALID, "world handle is NULL");
  std::lock_guard<std::recursive_mutex> run_lock(w->run_mutex);
  Prepared p;
  if (int s = prepare_run(w, source, nullptr, 0, seed0, first_index, n, options, true, p)) return s;
  if (n && !d_bank) return fail(MMC_ERR_INVALID, "d_bank is NULL");
  MMC_CUDA(cudaSetDevice(w->device));
  MMC_CUDA(MMC_BY_RNG(p.counter_rng, launch_source_bank, p.run, reinterpret_cast<BankSite*>(d_bank), p.stream));
  return MMC_OK;
}

int mmc_generation_run(const mmc_world* world, const mmc_site* d_bank_in, uint64_t n_in,
                       const mmc_estimator_desc* estimators, int32_t n_estimators, int32_t score,
                       const mmc_run_options* options, mmc_site* d_bank_out, uint64_t bank_capacity, uint64_t* d_n_out,
                       uint64_t* d_scores, uint64_t* d_square_scores, mmc_counters* d_counters, uint64_t* d_k_collision) {
  auto* w = const_cast<mmc_world*>(world);
  if (!w) return fail(MMC_ERR_INVALID, "world handle is NULL");
  std::lock_guard<std::recursive_mutex> run_lock(w->run_mutex);
  if (!w) return fail(MMC_ERR_INVALID, "world handle is NULL");
  // the source of a generation is the bank: a placeholder source satisfies prepare_run's checks
  mmc_source_desc placeholder{};
  placeholder.direction_kind = MMC_DIR_ISOTROPIC;
  placeholder.group = 1;
  placeholder.energy = 1.0;
  Prepared p;
  if (int s = prepare_run(w, &placeholder, estimators, score ? n_estimators : 0, 0, 0, n_in, options, false, p, true)) return s;
  if (!d_counters || !d_n_out) return fail(MMC_ERR_INVALID, "d_counters / d_n_out is NULL");
  if (n_in && (!d_bank_in || !d_bank_out)) return fail(MMC_ERR_INVALID, "bank buffers are NULL");
  if (p.run.total_bins && (!d_scores || !d_square_scores)) return fail(MMC_ERR_INVALID, "tally buffers are NULL");
  MMC_CUDA(cudaSetDevice(w->device));
  MMC_CUDA(cudaMemsetAsync(d_n_out, 0, sizeof(uint64_t), p.stream));
  if (n_in == 0) return MMC_OK;
  // every thread needs room for the secondaries of one fission
  if (w->has_fission) p.run.secondary_capacity = std::max<uint32_t>(p.run.secondary_capacity, 16);
  const size_t threads = static_cast<size_t>(p.cfg.blocks) * kThreadsPerBlock;
  if (int s = ensure_scratch(w, threads, p.run.secondary_capacity, p.run.pending_capacity, p.bounds.size())) return s;
  const size_t need_unordered = std::max<uint64_t>(bank_capacity, 1) * sizeof(BankSite);
  if (need_unordered > w->unordered_bytes) {
    cudaFree(w->d_unordered);
    w->d_unordered = nullptr;
    w->unordered_bytes = 0;
    MMC_CUDA(cudaMalloc(&w->d_unordered, need_unordered));
    w->unordered_bytes = need_unordered;
  }
  if (n_in > w->parents_capacity) {
    cudaFree(w->d_child_count);
    cudaFree(w->d_child_start);
    cudaFree(w->d_block_sums);
    w->d_child_count = nullptr;
    w->d_child_start = nullptr;
    w->d_block_sums = nullptr;
    w->parents_capacity = 0;
    MMC_CUDA(cudaMalloc(&w->d_child_count, n_in * sizeof(uint32_t)));
    MMC_CUDA(cudaMalloc(&w->d_child_start, n_in * sizeof(unsigned long long)));
    MMC_CUDA(cudaMalloc(&w->d_block_sums, (static_cast<size_t>(bank_scan_blocks(n_in)) + 1) * sizeof(unsigned long long)));
    w->parents_capacity = n_in;
  }
  if (!p.bounds.empty())
    MMC_CUDA(cudaMemcpyAsync(w->d_bounds, p.bounds.data(), p.bounds.size() * sizeof(double), cudaMemcpyHostToDevice, p.stream));
  MMC_CUDA(cudaMemsetAsync(w->d_next, 0, sizeof(unsigned long long), p.stream));
  MMC_CUDA(cudaMemsetAsync(w->d_child_count, 0, n_in * sizeof(uint32_t), p.stream));
  GenerationIO io;
  io.in = reinterpret_cast<const BankSite*>(d_bank_in);
  io.out = w->d_unordered;
  io.capacity = bank_capacity;
  io.n_out = reinterpret_cast<unsigned long long*>(d_n_out);
  io.child_count = w->d_child_count;
  io.child_start = w->d_child_start;
  io.k_collision = reinterpret_cast<unsigned long long*>(d_k_collision);
  MMC_CUDA(MMC_BY_RNG(p.counter_rng, launch_fixed_source,
      p.cfg, w->d_blob, p.run, w->d_bounds, w->d_sites, w->d_pending, w->d_next,
      reinterpret_cast<unsigned long long*>(d_scores), reinterpret_cast<unsigned long long*>(d_square_scores), d_counters,
      &io, p.stream));
  MMC_CUDA(launch_order_bank(
      w->d_child_count, w->d_child_start, n_in, w->d_block_sums, w->d_unordered, reinterpret_cast<BankSite*>(d_bank_out),
      reinterpret_cast<unsigned long long*>(d_n_out), p.stream));
  return MMC_OK;
}

int mmc_bank_resample(const mmc_world* world, const mmc_site* d_slice, uint64_t slice_first, uint64_t slice_n,
                      uint64_t m_total, uint64_t n_total, uint64_t first_out, uint64_t n_out,
                      const mmc_run_options* options, mmc_site* d_bank_next, uint64_t* d_errors) {
  if (!world) return fail(MMC_ERR_INVALID, "world handle is NULL");
  if (m_total == 0 || n_total == 0) return fail(MMC_ERR_INVALID, "empty fission bank: the chain died out (m_total = 0)");
  if (first_out + n_out > n_total) return fail(MMC_ERR_INVALID, "output range outside [0, n_total)");
  if (n_out && (!d_slice || !d_bank_next || !d_errors)) return fail(MMC_ERR_INVALID, "bank buffers are NULL");
  if (options && options->struct_size != sizeof(mmc_run_options)) return fail(MMC_ERR_INVALID, "mmc_run_options ABI mismatch");
  MMC_CUDA(cudaSetDevice(world->device));
  cudaStream_t stream = options && options->stream ? static_cast<cudaStream_t>(options->stream) : world->stream;
  const bool counter_rng = (options && options->rng_mode == MMC_RNG_COUNTER) ||
                           (std::getenv("MMC_RNG_MODE") && std::strcmp(std::getenv("MMC_RNG_MODE"), "counter") == 0);
  MMC_CUDA(MMC_BY_RNG(counter_rng, launch_resample_bank,
      reinterpret_cast<const BankSite*>(d_slice), slice_first, slice_n, m_total, n_total, first_out, n_out,
      reinterpret_cast<BankSite*>(d_bank_next), reinterpret_cast<unsigned long long*>(d_errors), stream));
  return MMC_OK;
}

int mmc_fixed_source_run_sensitivities(
    const mmc_world* world, const mmc_source_desc* source, const mmc_estimator_desc* estimators, int32_t n_estimators,
    const mmc_sensitivity_desc* sensitivities, int32_t n_sensitivities, uint64_t seed0, uint64_t first_history,
    uint64_t n_histories, const mmc_run_options* options, double* scores, double* square_scores, double* sens_scores,
    double* sens_square_scores, mmc_counters* counters) {
  if (n_sensitivities == 0)
    return mmc_fixed_source_run(world, source, estimators, n_estimators, seed0, first_history, n_histories, options,
                                scores, square_scores, counters);
  if (!world) return fail(MMC_ERR_INVALID, "world handle is NULL");
  auto* w = const_cast<mmc_world*>(world);
  std::lock_guard<std::recursive_mutex> run_lock(w->run_mutex);
  if (n_sensitivities < 0 || n_sensitivities > kMaxSensitivities)
    return fail(MMC_ERR_INVALID, "n_sensitivities %d out of range [0, %d]", n_sensitivities, kMaxSensitivities);
  if (!sensitivities || !sens_scores || !sens_square_scores)
    return fail(MMC_ERR_INVALID, "sensitivities / sens_scores / sens_square_scores is NULL");
  if (n_estimators > 0 && (!scores || !square_scores)) return fail(MMC_ERR_INVALID, "scores / square_scores is NULL");
  mmc_run_options opt{};
  if (options) opt = *options;
  opt.struct_size = sizeof(mmc_run_options);
  opt.schedule = MMC_SCHEDULE_FUSED;  // the sensitivity proxies live in the fused kernel
  Prepared p;
  if (int s = prepare_run(w, source, estimators, n_estimators, seed0, first_history, n_histories, &opt, false, p)) return s;
  // distinct perturbed nuclides, sensitivity offsets
  RunSpec& run = p.run;
  uint64_t sens_bins = 0;
  for (int i = 0; i < n_sensitivities; i++) {
    const mmc_sensitivity_desc& sd = sensitivities[i];
    if (sd.estimator < 0 || sd.estimator >= n_estimators)
      return fail(MMC_ERR_INVALID, "sensitivity %d: estimator index %d out of range", i, sd.estimator);
    if (sd.nuclide < 0 || sd.nuclide >= w->header.n_nuclides)
      return fail(MMC_ERR_INVALID, "sensitivity %d: nuclide index %d out of range", i, sd.nuclide);
    int k = 0;
    while (k < run.n_perturbations && run.perturbed_nuclide[k] != sd.nuclide) k++;
    if (k == run.n_perturbations) {
      if (k == kMaxPerturbations) return fail(MMC_ERR_INVALID, "more than %d distinct perturbed nuclides", kMaxPerturbations);
      run.perturbed_nuclide[run.n_perturbations++] = sd.nuclide;
    }
    run.sensitivities[i].estimator = sd.estimator;
    run.sensitivities[i].perturbation = k;
    run.sensitivities[i].offset = sens_bins;
    sens_bins += mmc_estimator_size(&estimators[sd.estimator]);
  }
  run.n_sensitivities = n_sensitivities;
  run.sens_pending_capacity = run.pending_capacity * static_cast<uint32_t>(n_sensitivities);
  MMC_CUDA(cudaSetDevice(w->device));
  if (n_histories == 0) return MMC_OK;
  // launch shape of the kPerturb instantiation
  int per_sm = MMC_BY_RNG(p.counter_rng, max_blocks_per_sm, run.tracking, run.continuous_energy != 0, false,
                          run.world_in_smem ? run.world_bytes : 0, true);
  if (per_sm < 1) per_sm = 1;
  long long blocks = static_cast<long long>(w->sm_count) * per_sm;
  const long long useful = static_cast<long long>((n_histories + kThreadsPerBlock - 1) / kThreadsPerBlock);
  if (blocks > useful) blocks = std::max<long long>(useful, 1);
  p.cfg.blocks = static_cast<int>(blocks);
  const size_t threads = static_cast<size_t>(p.cfg.blocks) * kThreadsPerBlock;
  if (int s = ensure_scratch(w, threads, run.secondary_capacity, run.pending_capacity, p.bounds.size())) return s;
  const size_t need_pending = threads * run.sens_pending_capacity * sizeof(SensitivityPending);
  if (need_pending > w->sens_pending_bytes) {
    cudaFree(w->d_sens_pending);
    w->d_sens_pending = nullptr;
    w->sens_pending_bytes = 0;
    MMC_CUDA(cudaMalloc(&w->d_sens_pending, need_pending));
    w->sens_pending_bytes = need_pending;
  }
  // device tallies: [sens scores | sens squares] doubles, [scores | squares | counters] integers
  const uint64_t total_bins = run.total_bins;
  constexpr size_t kCounterWords = sizeof(mmc_counters) / sizeof(unsigned long long);
  const size_t words = 2 * sens_bins + 2 * total_bins + kCounterWords;
  if (words > w->sens_words) {
    cudaFree(w->d_sens);
    if (w->h_sens) cudaFreeHost(w->h_sens);
    w->d_sens = w->h_sens = nullptr;
    w->sens_words = 0;
    MMC_CUDA(cudaMalloc(&w->d_sens, words * sizeof(double)));
    MMC_CUDA(cudaMallocHost(&w->h_sens, words * sizeof(double)));
    w->sens_words = words;
  }
  cudaStream_t stream = p.stream;
  MMC_CUDA(cudaMemsetAsync(w->d_sens, 0, words * sizeof(double), stream));
  if (!p.bounds.empty())
    MMC_CUDA(cudaMemcpyAsync(w->d_bounds, p.bounds.data(), p.bounds.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
  MMC_CUDA(cudaMemsetAsync(w->d_next, 0, sizeof(unsigned long long), stream));
  SensitivityIO io;
  io.pending = w->d_sens_pending;
  io.scores = w->d_sens;
  io.square_scores = w->d_sens + sens_bins;
  auto* d_tally = reinterpret_cast<unsigned long long*>(w->d_sens + 2 * sens_bins);
  auto* d_counters = reinterpret_cast<mmc_counters*>(d_tally + 2 * total_bins);
  w->last_launches = 1;
  MMC_CUDA(MMC_BY_RNG(p.counter_rng, launch_fixed_source, p.cfg, w->d_blob, run, w->d_bounds, w->d_sites, w->d_pending,
                      w->d_next, d_tally, d_tally + total_bins, d_counters, nullptr, stream, nullptr, &io));
  MMC_CUDA(cudaMemcpyAsync(w->h_sens, w->d_sens, words * sizeof(double), cudaMemcpyDeviceToHost, stream));
  MMC_CUDA(cudaStreamSynchronize(stream));
  for (uint64_t i = 0; i < sens_bins; i++) {
    sens_scores[i] += w->h_sens[i];
    sens_square_scores[i] += w->h_sens[sens_bins + i];
  }
  const auto* h_tally = reinterpret_cast<const unsigned long long*>(w->h_sens + 2 * sens_bins);
  for (uint64_t i = 0; i < total_bins; i++) {
    scores[i] += static_cast<double>(h_tally[i]);
    square_scores[i] += static_cast<double>(h_tally[total_bins + i]);
  }
  mmc_counters h_counters{};
  std::memcpy(&h_counters, h_tally + 2 * total_bins, sizeof(mmc_counters));
  if (counters) *counters = h_counters;
  return status_from_counters(h_counters);
}

int mmc_fixed_source_run(
    const mmc_world* world, const mmc_source_desc* source, const mmc_estimator_desc* estimators, int32_t n_estimators,
    uint64_t seed0, uint64_t first_history, uint64_t n_histories, const mmc_run_options* options, double* scores,
    double* square_scores, mmc_counters* counters) {
  if (!world) return fail(MMC_ERR_INVALID, "world handle is NULL");
  uint64_t total_bins = 0;
  for (int e = 0; e < n_estimators && estimators; e++) total_bins += mmc_estimator_size(&estimators[e]);
  if (total_bins && (!scores || !square_scores)) return fail(MMC_ERR_INVALID, "scores / square_scores is NULL");
  MMC_CUDA(cudaSetDevice(world->device));
  mmc_run_options opt{};
  if (options) opt = *options;
  opt.struct_size = sizeof(mmc_run_options);
  cudaStream_t stream = opt.stream ? static_cast<cudaStream_t>(opt.stream) : world->stream;
  // [scores | square_scores | counters] in one device buffer with a pinned host mirror, both kept by the world
  auto* w = const_cast<mmc_world*>(world);
  std::lock_guard<std::recursive_mutex> run_lock(w->run_mutex);
  constexpr size_t kCounterWords = sizeof(mmc_counters) / sizeof(unsigned long long);
  const size_t words = 2 * total_bins + kCounterWords;
  if (words > w->tally_words) {
    cudaFree(w->d_tally);
    if (w->h_tally) cudaFreeHost(w->h_tally);
    w->d_tally = w->h_tally = nullptr;
    w->tally_words = 0;
    MMC_CUDA(cudaMalloc(&w->d_tally, words * sizeof(unsigned long long)));
    MMC_CUDA(cudaMallocHost(&w->h_tally, words * sizeof(unsigned long long)));
    w->tally_words = words;
  }
  unsigned long long* d_tally = w->d_tally;
  mmc_counters* d_counters = reinterpret_cast<mmc_counters*>(d_tally + 2 * total_bins);
  MMC_CUDA(cudaMemsetAsync(d_tally, 0, words * sizeof(unsigned long long), stream));
  int status = mmc_fixed_source_run_device(
      world, source, estimators, n_estimators, seed0, first_history, n_histories, &opt,
      reinterpret_cast<uint64_t*>(d_tally), reinterpret_cast<uint64_t*>(d_tally + total_bins), d_counters);
  if (status != MMC_OK) return status;
  MMC_CUDA(cudaMemcpyAsync(w->h_tally, d_tally, words * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
  MMC_CUDA(cudaStreamSynchronize(stream));
  const unsigned long long* h_tally = w->h_tally;
  mmc_counters h_counters{};
  std::memcpy(&h_counters, h_tally + 2 * total_bins, sizeof(mmc_counters));
  // Scorable::operator+= (Scorable.cpp:37-48): integer-valued, exact below 2^53
  for (uint64_t i = 0; i < total_bins; i++) {
    scores[i] += static_cast<double>(h_tally[i]);
    square_scores[i] += static_cast<double>(h_tally[total_bins + i]);
  }
  if (counters) *counters = h_counters;
  return status_from_counters(h_counters);
}

int mmc_trace_histories(
    const mmc_world* world, const mmc_source_desc* source, uint64_t seed0, uint64_t first_history, uint64_t n_histories,
    const mmc_run_options* options, mmc_event_record* records, size_t cap, size_t* n_records) {
  auto* w = const_cast<mmc_world*>(world);
  if (!w) return fail(MMC_ERR_INVALID, "world handle is NULL");
  std::lock_guard<std::recursive_mutex> run_lock(w->run_mutex);
  Prepared p;
  if (int s = prepare_run(w, source, nullptr, 0, seed0, first_history, n_histories, options, true, p)) return s;
  if (!records || !n_records) return fail(MMC_ERR_INVALID, "records / n_records is NULL");
  *n_records = 0;
  MMC_CUDA(cudaSetDevice(w->device));
  // the single trace thread gets a roomy secondary queue
  p.run.secondary_capacity = std::max<uint32_t>(p.run.secondary_capacity, 1024);
  if (int s = ensure_scratch(w, 1, p.run.secondary_capacity, 1, 0)) return s;
  mmc_event_record* d_records = nullptr;
  unsigned long long* d_n = nullptr;
  mmc_counters* d_counters = nullptr;
  MMC_CUDA(cudaMalloc(&d_records, std::max<size_t>(cap, 1) * sizeof(mmc_event_record)));
  cudaError_t e = cudaMalloc(&d_n, sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMalloc(&d_counters, sizeof(mmc_counters));
  if (e == cudaSuccess) e = cudaMemsetAsync(d_counters, 0, sizeof(mmc_counters), p.stream);
  if (e == cudaSuccess)
    e = MMC_BY_RNG(p.counter_rng, launch_trace, w->d_blob, p.run, w->d_sites, d_records, cap, d_n, d_counters, p.stream);
  unsigned long long n = 0;
  mmc_counters h_counters{};
  if (e == cudaSuccess) e = cudaMemcpyAsync(&n, d_n, sizeof(n), cudaMemcpyDeviceToHost, p.stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&h_counters, d_counters, sizeof(h_counters), cudaMemcpyDeviceToHost, p.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(p.stream);
  if (e == cudaSuccess)
    e = cudaMemcpy(records, d_records, std::min<size_t>(n, cap) * sizeof(mmc_event_record), cudaMemcpyDeviceToHost);
  cudaFree(d_records);
  cudaFree(d_n);
  cudaFree(d_counters);
  if (e != cudaSuccess) return fail(MMC_ERR_CUDA, "mmc_trace_histories: %s", cudaGetErrorString(e));
  *n_records = static_cast<size_t>(std::min<unsigned long long>(n, cap));
  if (n > cap) return fail(MMC_ERR_CAPACITY, "trace needs %llu records but cap is %zu", n, cap);
  return status_from_counters(h_counters);
}

int mmc_device_alloc(const mmc_world* world, size_t bytes, void** d_ptr) {
  if (!world || !d_ptr) return fail(MMC_ERR_INVALID, "world / d_ptr is NULL");
  *d_ptr = nullptr;
  MMC_CUDA(cudaSetDevice(world->device));
  MMC_CUDA(cudaMalloc(d_ptr, std::max<size_t>(bytes, 1)));
  MMC_CUDA(cudaMemsetAsync(*d_ptr, 0, std::max<size_t>(bytes, 1), world->stream));
  MMC_CUDA(cudaStreamSynchronize(world->stream));
  return MMC_OK;
}

void mmc_device_free(const mmc_world* world, void* d_ptr) {
  if (!world || !d_ptr) return;
  cudaSetDevice(world->device);
  cudaFree(d_ptr);
}

int mmc_device_zero(const mmc_world* world, void* d_ptr, size_t bytes) {
  if (!world || (bytes && !d_ptr)) return fail(MMC_ERR_INVALID, "world / d_ptr is NULL");
  MMC_CUDA(cudaSetDevice(world->device));
  MMC_CUDA(cudaMemsetAsync(d_ptr, 0, bytes, world->stream));
  return MMC_OK;
}

int mmc_device_read(const mmc_world* world, void* host_dst, const void* d_src, size_t bytes) {
  if (!world || (bytes && (!host_dst || !d_src))) return fail(MMC_ERR_INVALID, "bad arguments");
  MMC_CUDA(cudaSetDevice(world->device));
  MMC_CUDA(cudaMemcpyAsync(host_dst, d_src, bytes, cudaMemcpyDeviceToHost, world->stream));
  MMC_CUDA(cudaStreamSynchronize(world->stream));
  return MMC_OK;
}

int mmc_device_write(const mmc_world* world, void* d_dst, const void* host_src, size_t bytes) {
  if (!world || (bytes && (!d_dst || !host_src))) return fail(MMC_ERR_INVALID, "bad arguments");
  MMC_CUDA(cudaSetDevice(world->device));
  MMC_CUDA(cudaMemcpyAsync(d_dst, host_src, bytes, cudaMemcpyHostToDevice, world->stream));
  MMC_CUDA(cudaStreamSynchronize(world->stream));
  return MMC_OK;
}

int mmc_test_geometry(const mmc_world* world, size_t n, const double* positions, const double* directions,
                      int32_t* cell, int32_t* surface, double* distance) {
  if (!world || !positions || !directions || !cell || !surface || !distance) return fail(MMC_ERR_INVALID, "bad arguments");
  if (n == 0) return MMC_OK;
  MMC_CUDA(cudaSetDevice(world->device));
  double *d_pos = nullptr, *d_dir = nullptr, *d_dist = nullptr;
  int32_t *d_cell = nullptr, *d_surf = nullptr;
  cudaError_t e = cudaMalloc(&d_pos, 3 * n * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&d_dir, 3 * n * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&d_dist, n * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&d_cell, n * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc(&d_surf, n * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMemcpy(d_pos, positions, 3 * n * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_dir, directions, 3 * n * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = launch_test_geometry(world->d_blob, n, d_pos, d_dir, d_cell, d_surf, d_dist, nullptr);
  if (e == cudaSuccess) e = cudaMemcpy(cell, d_cell, n * sizeof(int32_t), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(surface, d_surf, n * sizeof(int32_t), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(distance, d_dist, n * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d_pos);
  cudaFree(d_dir);
  cudaFree(d_dist);
  cudaFree(d_cell);
  cudaFree(d_surf);
  if (e != cudaSuccess) return fail(MMC_ERR_CUDA, "mmc_test_geometry: %s", cudaGetErrorString(e));
  return MMC_OK;
}

int mmc_test_device_math(int fn, const double* x, double* out0, double* out1, size_t n) {
  if (fn < 0 || (fn > 5 && fn != 16 + 4) || !x || !out0 || ((fn == 1 || fn == 5) && !out1)) return fail(MMC_ERR_INVALID, "bad arguments");
  if (mmc_device_count() < 1) return fail(MMC_ERR_NO_DEVICE, "no CUDA device visible");
  if (n == 0) return MMC_OK;
  double *d_x = nullptr, *d_0 = nullptr, *d_1 = nullptr;
  cudaError_t e = cudaMalloc(&d_x, n * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&d_0, n * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&d_1, n * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpy(d_x, x, n * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = fn >= 16 ? mmc::ctr::launch_test_math(fn - 16, d_x, d_0, d_1, n, nullptr) : mmc::lcg::launch_test_math(fn, d_x, d_0, d_1, n, nullptr);
  if (e == cudaSuccess) e = cudaMemcpy(out0, d_0, n * sizeof(double), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && (fn == 1 || fn == 5)) e = cudaMemcpy(out1, d_1, n * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d_x);
  cudaFree(d_0);
  cudaFree(d_1);
  if (e != cudaSuccess) return fail(MMC_ERR_CUDA, "mmc_test_device_math: %s", cudaGetErrorString(e));
  return MMC_OK;
}

}  // extern "C"
