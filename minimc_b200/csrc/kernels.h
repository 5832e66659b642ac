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

// ---- event-split schedule (event_loop.cu) -----------------------------------------
// Particle state of n_slots concurrently running histories as structure-of-arrays in HBM.  A slot is one history
// context (current particle, per-history pending-score table, secondary deque); the flight and the S(a,b) kernels
// stream the arrays they need, coalesced over the compacted queues of slot indices.
struct EventState {
  double *px, *py, *pz, *dx, *dy, *dz, *energy;  // [n_slots]
  double* tsl_T;                                 // temperature at the pending S(a,b) collision
  uint32_t* rng;                                 // minstd_rand state; counter mode: draws made from the stream
  uint32_t *rng_k0, *rng_k1;                     // counter mode: the stream's 64-bit id
  int32_t *cell, *surface, *event;
  uint32_t* n_pending;                           // entries used in the slot's pending-score table
  uint32_t *dq_head, *dq_count;                  // secondary deque (worlds with fission)
  uint32_t* tsl_off;                             // blob offset of the TslTable of the pending S(a,b) collision
};
constexpr size_t kEventStateBytesPerSlot = 8 * 8 + 10 * 4;

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

}  // namespace mmc

// The launch functions exist once per RNG mode (mmc_rng_mode): kernels.cu and event_loop.cu are compiled twice, with
// MMC_COUNTER_RNG = 0 into namespace mmc::lcg (std::minstd_rand, bit-exact with the reference) and = 1 into
// namespace mmc::ctr (Philox-2x32-10 per particle, transport.cuh).  capi.cu declares both and dispatches.
#ifndef MMC_COUNTER_RNG
#define MMC_COUNTER_RNG 0
#endif
#ifdef MMC_DECLARE_BOTH_VARIANTS
#define MMC_VARIANT_NS lcg
#include "kernel_launches.inc"
#undef MMC_VARIANT_NS
#define MMC_VARIANT_NS ctr
#include "kernel_launches.inc"
#undef MMC_VARIANT_NS
#else
#if MMC_COUNTER_RNG
#define MMC_VARIANT_NS ctr
#else
#define MMC_VARIANT_NS lcg
#endif
#include "kernel_launches.inc"
#endif
