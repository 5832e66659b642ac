// Host-side launch interface of kernels.cu (internal; the public boundary is
// include/minimc_b200.h).
#pragma once

#include <cuda_runtime.h>

#include "../../include/minimc_b200.h"
#include "world_blob.h"

namespace mmc {

constexpr int kThreadsPerBlock = 256;

struct LaunchConfig {
  int blocks;
};

// Fission-bank plumbing of one k-eigenvalue generation (device pointers).
struct GenerationIO {
  const BankSite* in = nullptr;              // source bank of this generation, [n_histories]
  BankSite* out = nullptr;                   // unordered fission bank, [capacity]
  unsigned long long capacity = 0;
  unsigned long long* n_out = nullptr;       // sites claimed so far
  uint32_t* child_count = nullptr;           // [n_histories] secondaries of each source particle
  unsigned long long* child_start = nullptr; // [n_histories] where that particle's run starts in `out`
  unsigned long long* k_collision = nullptr; // sum of the collision estimator's scores, fixed point (MMC_K_COLLISION_ONE); may be null
};

// fills the dense reconstruction tables of the device-only tail of the image (world_blob.h DenseJob)
cudaError_t launch_expand_dense(char* world_d, const DenseJob* jobs, size_t n_jobs, cudaStream_t stream);
// fills the evaluated S(a,b) tables (world_blob.h EvalJob, TslPartition::off_eval)
cudaError_t launch_evaluate_rows(char* world_d, const EvalJob* jobs, size_t n_jobs, cudaStream_t stream);

cudaError_t launch_source_bank(const RunSpec& run, BankSite* bank, cudaStream_t stream);
uint32_t bank_scan_blocks(uint64_t n_parents);
cudaError_t launch_order_bank(
    const uint32_t* child_count, const unsigned long long* child_start, uint64_t n_parents, unsigned long long* block_sums,
    const BankSite* unordered, BankSite* ordered, unsigned long long* n_sites, cudaStream_t stream);
cudaError_t launch_resample_bank(
    const BankSite* slice, uint64_t slice_first, uint64_t slice_n, uint64_t m_total, uint64_t n_total, uint64_t first_out,
    uint64_t n_out, BankSite* next, unsigned long long* errors, cudaStream_t stream);

cudaError_t launch_trace(
    const char* world_d, const RunSpec& run, BankSite* site_scratch, mmc_event_record* records, unsigned long long cap,
    unsigned long long* n_records, mmc_counters* counters, cudaStream_t stream);

cudaError_t launch_test_math(int fn, const double* x_d, double* out0_d, double* out1_d, size_t n, cudaStream_t stream);

cudaError_t launch_test_geometry(
    const char* world_d, size_t n, const double* pos_d, const double* dir_d, int32_t* cell_d, int32_t* surface_d,
    double* distance_d, cudaStream_t stream);

// ---- event-split schedule (event_loop.cu) -----------------------------------------
// Particle state of n_slots concurrently running histories as structure-of-arrays in HBM.  A slot is one history
// context (current particle, per-history pending-score table, secondary deque); the flight and the S(a,b) kernels
// stream the arrays they need, coalesced over the compacted queues of slot indices.
struct EventState {
  double *px, *py, *pz, *dx, *dy, *dz, *energy;  // [n_slots]
  double* tsl_T;                                 // temperature at the pending S(a,b) collision
  uint32_t* rng;                                 // minstd_rand state
  int32_t *cell, *surface, *event;
  uint32_t* n_pending;                           // entries used in the slot's pending-score table
  uint32_t *dq_head, *dq_count;                  // secondary deque (worlds with fission)
  uint32_t* tsl_off;                             // blob offset of the TslTable of the pending S(a,b) collision
};
constexpr size_t kEventStateBytesPerSlot = 8 * 8 + 8 * 4;

// Sensitivity tallies of a run with perturbations (fused kernel, kPerturb): real-valued, fp64 atomics.
struct SensitivityPending {
  uint32_t bin;  // index into the concatenated sensitivity tallies
  uint32_t pad;
  double sum;    // this history's score in that bin so far
};
struct SensitivityIO {
  SensitivityPending* pending = nullptr;  // [threads][RunSpec::sens_pending_capacity]
  double* scores = nullptr;               // [total sensitivity bins]
  double* square_scores = nullptr;
};

// Hand-over from the event-split schedule to the fused kernel (the tail of a run, when too few histories are alive
// to fill the GPU and every pass costs its launch latency): thread t < n adopts slot slots[t] -- its particle, pending
// table and secondary deque -- and runs it, and any history it can still claim, to the end.
struct ResumeIO {
  const uint32_t* slots = nullptr;  // compacted live slots
  const unsigned int* n = nullptr;  // how many (device-side count)
  EventState st{};
};

struct EventQueues {
  uint32_t* alive[2];   // compacted slot indices, ping-pong between passes
  uint32_t* tsl;        // slots whose collision awaits S(a,b) sampling in this pass
  uint32_t* boundary;   // slots whose particle is on a surface or dead: Cell lookup, tallies, next particle / history
  unsigned int* count;  // [0..1] alive counts, [2..3] tsl counts (by pass parity), [4] chunk counter of the S(a,b) kernel, [5..6] boundary counts, [7] chunk counter of the boundary queue
};

constexpr int kCounterReplicas = 64;

// array length for n_slots slots: whole CTAs of the flight kernel may read (not use) queue entries past the end
inline uint32_t event_padded_slots(uint32_t n_slots) { return (n_slots + 511u) & ~511u; }

struct EventTslConfig {
  uint32_t sm_count = 0;
  uint32_t sc_arena_bytes = 0;
  bool shared_sc = false;  // the S*CDF_modes arena fits into shared memory next to the per-lane mode rows
};

cudaError_t launch_event_init(const EventState& st, const EventQueues& q, uint32_t n_slots,
                              unsigned long long* counter_replicas, cudaStream_t stream);
// one pass = one event of every live slot: the flight kernel, then the boundary kernel and the S(a,b) kernel over the
// slots it queued
cudaError_t launch_event_pass(
    const char* world_d, const WorldHeader& header, const RunSpec& run, const double* bounds_d, const EventState& st, const EventQueues& q,
    uint32_t pass, uint32_t alive_upper_bound, BankSite* site_scratch, uint2* pending_scratch,
    unsigned long long* next_history, unsigned long long* scores, unsigned long long* square_scores,
    unsigned long long* counter_replicas, const EventTslConfig& tsl, cudaStream_t stream,
    const cudaEvent_t* marks = nullptr);  // profile mode: marks[0] after the flight kernel, marks[1] after the boundary kernel
int event_kernels_per_pass();
// shared-memory plan of the S(a,b) kernel for a world (opts the kernels into their dynamic shared memory)
cudaError_t configure_event_tsl(uint32_t sc_arena_bytes, size_t smem_optin, uint32_t sm_count, EventTslConfig& out);
cudaError_t launch_event_finish(const unsigned long long* counter_replicas, mmc_counters* counters, cudaStream_t stream);

// generation == nullptr: fixed source; otherwise one k-eigenvalue generation over generation->in
cudaError_t launch_fixed_source(
    const LaunchConfig& cfg, const char* world_d, const RunSpec& run, const double* bounds_d, BankSite* site_scratch,
    uint2* pending_scratch, unsigned long long* next_history, unsigned long long* scores,
    unsigned long long* square_scores, mmc_counters* counters, const GenerationIO* generation, cudaStream_t stream,
    const ResumeIO* resume = nullptr, const SensitivityIO* sensitivity = nullptr);

// occupancy query for the fused kernel
int max_blocks_per_sm(int tracking, bool continuous_energy, bool generation, size_t smem, bool perturb = false);

}  // namespace mmc
