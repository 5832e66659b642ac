// Event-split schedule of the history loop for continuous-energy worlds (sm_100a).
//
// The fused kernel (kernels.cu) keeps a particle in registers for its whole life.
// For the S(a,b) decks that puts ~14 k instructions (230 KB) of divergent code in
// one kernel: ncu shows warps waiting on instruction fetch as often as on memory.
// Here one PASS advances every live history by one event in three small pieces of code (two launches: the boundary
// work runs as the prologue of the persistent S(a,b) kernel unless built with -DMMC_EV_MERGE_BOUNDARY=0):
//
//   event_flight_kernel    cross-section lookup, distance to collision and to the nearest surface, the flight;
//                          at a collision: nuclide and reaction choice, capture / free-gas scatter / fission.
//                          A collision that chose thermal scattering is not sampled here (-> S(a,b) queue); a
//                          particle that reached a surface, or died, is not finished here (-> boundary queue).
//   boundary_chunk         dense over the boundary queue: Cell lookup after a crossing, leak, tallies; next
//                          particle of the history's bank, next history (Source::Sample), retirement of the slot.
//   event_tsl_kernel       ThermalScattering::Scatter (SampleBeta, SampleAlpha, Particle::Scatter) for the
//                          S(a,b) queue -- every lane of every warp runs the POD sampler.
//
// Particle state streams through HBM as structure-of-arrays (EventState, kernels.h) indexed by SLOT; a slot is one
// history context (current particle, pending-score table, secondary deque), so the per-history bookkeeping of
// FixedSource.cpp:48-72 stays slot-local exactly as it is lane-local in the fused kernel.  Between kernels the live,
// S(a,b) and boundary slots are stream-compacted: ballot + popc inside a warp, the eight warp totals scanned in
// shared memory, ONE global atomic per CTA and queue, order inside a CTA preserved.  The arithmetic per particle is
// the same device code (transport.cuh, physics_ce.cuh) in the same order, so results stay bit-identical to the
// fused kernel and to the reference; tallies are integer sums and do not depend on the schedule.
#include <cstdint>
#include <cuda_runtime.h>

#include "kernel_common.cuh"

// flight kernel CTA: threads, and resident CTAs per SM the register allocation is tuned for (768 threads per SM)
#ifndef MMC_EV_FLIGHT_THREADS
#define MMC_EV_FLIGHT_THREADS 128
#endif
#ifndef MMC_EV_FLIGHT_BLOCKS
#define MMC_EV_FLIGHT_BLOCKS (768 / MMC_EV_FLIGHT_THREADS)
#endif
// The S(a,b) kernel runs one persistent CTA per SM.  Measured on B200 (single_zone, hist/s): 768 threads with the
// mode rows in shared memory 6.90e7; 512 threads with the mode rows in 40 registers per lane 7.18e7 (384: 6.64e7,
// 640: 6.52e7).  Shared-memory rows keep the kernel at the 128 B/clk/SM shared-memory roof (ncu: 73 % of peak L1
// wavefronts, 63 % of them shared); register rows trade a third of the warps for a third of the wavefronts.
#ifndef MMC_EV_TSL_ROWS_IN_REGS
#define MMC_EV_TSL_ROWS_IN_REGS 1
#endif
#ifndef MMC_EV_TSL_THREADS
#define MMC_EV_TSL_THREADS (MMC_EV_TSL_ROWS_IN_REGS ? 512 : 768)
#endif

// Slot state is touched once per kernel and handed to the next kernel through L2: loads and stores bypass L1
// (ld.global.cg / st.global.cg), which stays with the cross-section and S(a,b) table gathers.
// the S(a,b) kernel of worlds whose partitions are all expanded into dense tables (ce::DenseRows): no mode rows in
// registers, so more, thinner threads to cover the L2 latency of the table gathers
#ifndef MMC_EV_TSL_DENSE_THREADS
#define MMC_EV_TSL_DENSE_THREADS 768
#endif
// the S(a,b) kernel of worlds whose every collision samples over evaluated rows (WorldHeader::tsl_all_direct): nothing
// but ce::tsl_sample_direct -- no sampler state machine, no spills
#ifndef MMC_EV_TSL_DIRECT_THREADS
#define MMC_EV_TSL_DIRECT_THREADS 1024
#endif
// share of a pass's S(a,b) queue that the warps of the persistent kernel take as fixed contiguous pieces (no atomics)
#ifndef MMC_EV_TSL_STATIC_PCT
#define MMC_EV_TSL_STATIC_PCT 75
#endif
#ifndef MMC_STATE_CACHE_GLOBAL
#define MMC_STATE_CACHE_GLOBAL 1
#endif
#if MMC_STATE_CACHE_GLOBAL
#define MMC_LD(x) __ldcg(&(x))
#define MMC_ST(x, v) __stcg(&(x), (v))
#else
#define MMC_LD(x) (x)
#define MMC_ST(x, v) ((x) = (v))
#endif

namespace mmc {
namespace MMC_VARIANT_NS {  // lcg or ctr: this file is compiled once per RNG mode (kernels.h)

namespace {

// a slot's generator: the minstd state, or (counter mode) the draws made and the stream's id
__device__ __forceinline__ Rng load_rng(const EventState& st, uint32_t slot) {
  Rng r;
  r.x = MMC_LD(st.rng[slot]);
#if MMC_COUNTER_RNG
  r.k0 = MMC_LD(st.rng_k0[slot]);
  r.k1 = MMC_LD(st.rng_k1[slot]);
#endif
  return r;
}
// kStream: the particle may be a new one (births in the boundary work); a flight or a scatter keeps its stream
template <bool kStream>
__device__ __forceinline__ void store_rng(const EventState& st, uint32_t slot, const Rng& r) {
  MMC_ST(st.rng[slot], r.x);
#if MMC_COUNTER_RNG
  if (kStream) {
    MMC_ST(st.rng_k0[slot], r.k0);
    MMC_ST(st.rng_k1[slot], r.k1);
  }
#endif
}

constexpr int kFlightThreads = MMC_EV_FLIGHT_THREADS;
constexpr int kWarpsPerBlock = kFlightThreads / 32;
constexpr int kTslThreads = MMC_EV_TSL_THREADS;
constexpr int kTslDenseThreads = MMC_EV_TSL_DENSE_THREADS;
constexpr int kTslDirectThreads = MMC_EV_TSL_DIRECT_THREADS;
// kinds of the S(a,b) kernel: where a reconstruction reads from
enum : int { kTslRowsGlobalSc = 0, kTslRowsSharedSc = 1, kTslDense = 2, kTslDirect = 3 };
__host__ __device__ constexpr int tsl_threads_of(int kind) {
  return kind == kTslDirect ? kTslDirectThreads : kind == kTslDense ? kTslDenseThreads : kTslThreads;
}
constexpr size_t kTslRowBytes = MMC_EV_TSL_ROWS_IN_REGS ? 0 : 10 * kTslThreads * sizeof(double2);

// exclusive prefix of this warp among the CTA's warp totals, and the CTA total
__device__ __forceinline__ uint32_t warp_prefix(const uint32_t* totals, uint32_t warp, uint32_t& block_total) {
  uint32_t prefix = 0, total = 0;
#pragma unroll
  for (int k = 0; k < kWarpsPerBlock; k++) {
    const uint32_t v = totals[k];
    if (static_cast<uint32_t>(k) < warp) prefix += v;
    total += v;
  }
  block_total = total;
  return prefix;
}

__global__ void event_init_kernel(const __grid_constant__ EventState st, const __grid_constant__ EventQueues q,
                                  uint32_t n_slots, uint32_t n_padded, unsigned long long* counter_replicas) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_slots) {
    st.event[i] = MMC_EV_CAPTURE;  // "dead": the first pass refills every slot
    // the flight kernel loads a slot's particle together with its event code, before it knows the slot is dead
    st.px[i] = st.py[i] = st.pz[i] = st.dx[i] = st.dy[i] = st.dz[i] = st.energy[i] = st.tsl_T[i] = 0.0;
    st.rng[i] = st.rng_k0[i] = st.rng_k1[i] = 0;
    st.cell[i] = st.surface[i] = -1;
    st.tsl_off[i] = 0;
    st.n_pending[i] = 0;
    st.dq_head[i] = 0;
    st.dq_count[i] = 0;
  }
  if (i < n_padded) {  // queue entries past the end must still be valid slots (event_flight_kernel reads them)
    q.alive[0][i] = i < n_slots ? i : 0u;
    q.alive[1][i] = 0;
  }
  if (i < kCounterReplicas * kNumCounters) counter_replicas[i] = 0;
  if (i == 0) {
    q.count[0] = n_slots;
    q.count[1] = q.count[2] = q.count[3] = q.count[4] = q.count[5] = q.count[6] = q.count[7] = 0;
  }
}

// CTA-wide stream compaction of up to kQueues flags per thread: ballot + popc inside a warp, the warp totals scanned
// in shared memory, ONE global atomic per CTA and queue, order inside the CTA preserved.  Returns each thread's
// position in queue k (valid where its flag is set).
template <int kQueues>
struct CtaCompactor {
  uint32_t totals[kQueues][kWarpsPerBlock];
  uint32_t base[kQueues];
};

template <int kQueues, typename BeforeBarrier, typename AfterBarrier>
__device__ __forceinline__ void cta_compact(
    CtaCompactor<kQueues>& sm, const bool (&flag)[kQueues], unsigned int* const (&counter)[kQueues], uint32_t (&position)[kQueues],
    BeforeBarrier before_barrier, AfterBarrier after_barrier) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t lanes_below = (1u << lane) - 1u;
  unsigned mask[kQueues];
#pragma unroll
  for (int k = 0; k < kQueues; k++) {
    mask[k] = __ballot_sync(kFull, flag[k]);
    if (lane == 0) sm.totals[k][warp] = __popc(mask[k]);
  }
  before_barrier();
  __syncthreads();
  after_barrier();
  if (threadIdx.x < kQueues) {
    uint32_t total = 0;
    for (int wi = 0; wi < kWarpsPerBlock; wi++) total += sm.totals[threadIdx.x][wi];
    sm.base[threadIdx.x] = total ? atomicAdd(counter[threadIdx.x], total) : 0u;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kQueues; k++) {
    uint32_t total;
    position[k] = sm.base[k] + warp_prefix(sm.totals[k], warp, total) + __popc(mask[k] & lanes_below);
  }
}

// per-CTA counter flush: the 0/1-per-lane counters packed four to a word, one warp reduction (REDUX) per word, the
// warp sums through shared memory, one atomic per non-zero counter and CTA into one of 64 replicas.  In two halves
// around a CTA barrier the caller has anyway (the flight kernel's is the first barrier of its stream compaction).
__device__ __forceinline__ void stage_counters_cta(const ThreadCounters& c, bool has_secondaries, uint4* s_packed) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t pa = c.events | (c.collisions << 8) | (c.crossings << 16) | (c.virtuals << 24);
  const uint32_t pb = c.histories | (c.births << 8) | (c.lost << 16) | (c.physics << 24);
  const uint32_t pd = c.scores | (c.capacity << 16);  // at most kMaxEstimators + 1 each per lane
  const uint32_t sa = __reduce_add_sync(kFull, pa), sb = __reduce_add_sync(kFull, pb), sd = __reduce_add_sync(kFull, pd);
  const uint32_t ss = has_secondaries ? __reduce_add_sync(kFull, c.secondaries) : 0u;
  if (lane == 0) s_packed[warp] = make_uint4(sa, sb, sd, ss);
}
// after the barrier: threads first_thread .. first_thread + kNumCounters - 1 add one counter each
__device__ __forceinline__ void commit_counters_cta(const uint4* s_packed, uint32_t first_thread, unsigned long long* counter_replicas) {
  if (threadIdx.x >= first_thread && threadIdx.x < first_thread + kNumCounters) {
    // mmc_counters order: histories, births, events, collisions, crossings, virtual, scores, secondaries, banked,
    // lost, capacity, physics -> (word, shift, mask) of the packed sums
    const uint32_t k = threadIdx.x - first_thread;
    const uint32_t word = (k == 2 || k == 3 || k == 4 || k == 5) ? 0u : (k == 0 || k == 1 || k == 9 || k == 11) ? 1u : (k == 6 || k == 10) ? 2u : 3u;
    const uint32_t shift = k == 3 ? 8u : k == 4 ? 16u : k == 5 ? 24u : k == 1 ? 8u : k == 9 ? 16u : k == 11 ? 24u : k == 10 ? 16u : 0u;
    const uint32_t mask = word < 2u ? 0xffu : word == 2u ? 0xffffu : 0xffffffffu;
    uint32_t sum = 0;
    if (k != 8) {
      for (int wi = 0; wi < kWarpsPerBlock; wi++) {
        const uint4 v = s_packed[wi];
        const uint32_t field = word == 0u ? v.x : word == 1u ? v.y : word == 2u ? v.z : v.w;
        sum += (field >> shift) & mask;
      }
    }
    if (sum) atomicAdd(counter_replicas + (blockIdx.x % kCounterReplicas) * kNumCounters + k, static_cast<unsigned long long>(sum));
  }
}

// ---- flight: one Transport-loop iteration of every live particle up to the point where the event's kind is known
template <int kTracking>
__global__ void __launch_bounds__(kFlightThreads, MMC_EV_FLIGHT_BLOCKS) event_flight_kernel(
    const char* __restrict__ world_g, const __grid_constant__ WorldHeader header, const __grid_constant__ RunSpec run,
    const __grid_constant__ EventState st, const __grid_constant__ EventQueues q, uint32_t pass,
    BankSite* __restrict__ site_scratch, unsigned long long* counter_replicas) {
  __shared__ CtaCompactor<3> s_compact;  // live slots, S(a,b) slots, boundary slots
  __shared__ uint4 s_packed[kWarpsPerBlock];

  const uint32_t parity = pass & 1u;
  const uint32_t first = blockIdx.x * kFlightThreads;
  const uint32_t i = first + threadIdx.x;
  // Two dependent round trips to HBM instead of four: the queue entry is read together with the queue length (the
  // queues are zero-filled past their end, so a stale entry is a valid slot), and the whole particle together with
  // the slot's event code (a dead slot's particle is loaded in vain).
  const uint32_t slot = q.alive[parity][i];
  const uint32_t n = q.count[parity];
  Particle p;
  p.event = MMC_LD(st.event[slot]);
  p.px = MMC_LD(st.px[slot]), p.py = MMC_LD(st.py[slot]), p.pz = MMC_LD(st.pz[slot]);
  p.dx = MMC_LD(st.dx[slot]), p.dy = MMC_LD(st.dy[slot]), p.dz = MMC_LD(st.dz[slot]);
  p.energy = MMC_LD(st.energy[slot]);
  p.group = 0;
  p.rng = load_rng(st, slot);
  p.cell = MMC_LD(st.cell[slot]);
  p.surface = -1;  // written only by a crossing; tallies read it in the boundary kernel
  if (blockIdx.x == 0 && threadIdx.x == 0) q.count[4] = q.count[7] = 0;  // chunk counters of this pass's S(a,b) kernel
  if (first >= n) return;  // CTA-uniform
  const bool valid = i < n;
  const WorldView w(world_g, &header);
  const bool has_secondaries = run.secondary_capacity > 1;

  const bool retired = !valid || p.event == kEvRetired;
  const bool alive = !retired && is_alive(p.event);
  SiteDeque dq;
  dq.slots = site_scratch + static_cast<size_t>(slot) * run.secondary_capacity;
  dq.mask = run.secondary_capacity - 1;
  dq.head = 0;
  dq.count = 0;
  if (alive && has_secondaries) {
    dq.head = MMC_LD(st.dq_head[slot]);
    dq.count = MMC_LD(st.dq_count[slot]);
  }
  ThreadCounters c;
  StepOut o;
  o.secondaries = 0;
  o.need_direction = false;
  o.need_tsl = o.need_cross = false;
  o.error_physics = o.error_capacity = o.error_lost = false;
  if (alive) {
    // p.cell >= 0: the boundary kernel looked the Cell up at birth and after every crossing
    transport_step<kTracking, true, false, true, false, true>(w, p, dq, o);
    count_event(c, p, o);
    MMC_ST(st.px[slot], p.px), MMC_ST(st.py[slot], p.py), MMC_ST(st.pz[slot], p.pz);
    // direction and energy change in this kernel only where a scatter is sampled here (free gas): a flight to a
    // surface, a virtual collision and a collision whose S(a,b) sampling is still to come leave both as loaded
    if (p.event == MMC_EV_SCATTER && !o.need_tsl) {
      MMC_ST(st.dx[slot], p.dx), MMC_ST(st.dy[slot], p.dy), MMC_ST(st.dz[slot], p.dz);
      MMC_ST(st.energy[slot], p.energy);
    }
    store_rng<false>(st, slot, p.rng);
    MMC_ST(st.event[slot], p.event);
    if (o.need_cross) MMC_ST(st.surface[slot], p.surface);
    if (has_secondaries) {
      MMC_ST(st.dq_head[slot], dq.head);
      MMC_ST(st.dq_count[slot], dq.count);
    }
    if (o.need_tsl) {
      // (the direct S(a,b) kernel reads the temperature of an evaluated cell from the cell's field record)
      if (!header.tsl_all_direct) MMC_ST(st.tsl_T[slot], o.tsl_T);
      MMC_ST(st.tsl_off[slot], o.tsl_off);
    }
  }
  // ---- stream compaction into the three queues of this pass
  //   live:     every slot that has not retired (it is looked at again next pass)
  //   S(a,b):   the collision chose thermal scattering
  //   boundary: the particle is on a surface (Cell lookup, leak, tallies), or the slot's particle is dead (capture,
  //             leak, fission; a dead slot of the previous pass, e.g. all of them in pass 0): next particle / history
  const bool flag[3] = {!retired, alive && o.need_tsl, !retired && (o.need_cross || !is_alive(p.event))};
  unsigned int* const counter[3] = {&q.count[parity ^ 1u], &q.count[2u + parity], &q.count[5u + parity]};
  uint32_t position[3];
  // the counters ride on the compaction's first barrier: staged before it, added (by the second warp) after it
  cta_compact<3>(s_compact, flag, counter, position,
                 [&]() { stage_counters_cta(c, has_secondaries, s_packed); },
                 [&]() { commit_counters_cta(s_packed, kFlightThreads >= 64 ? 32u : 0u, counter_replicas); });
  if (flag[0]) q.alive[parity ^ 1u][position[0]] = slot;
  if (flag[1]) q.tsl[position[1]] = slot;
  if (flag[2]) q.boundary[position[2]] = slot;
}

// ---- boundary: the rare ends of an event, run densely over the slots the flight kernel queued.
//   * a particle on a surface: World::FindCellContaining, surface_cross or leak (TransportMethod.cpp:68-73),
//     EstimatorSetProxy::Score (:74) -- `current` estimators score nothing else;
//   * a dead particle (captured or fissioned in this pass's flight, leaked just above, or dead since an earlier
//     pass): the next particle of the slot's history from its bank (FixedSource.cpp:63-71), else a new history
//     (one atomic per warp claims the indices), Source::Sample, and the Cell of the newborn (TransportMethod.cpp:55);
//     a slot that finds no history left retires.
// In a steady-state single_zone pass 8 % of the live slots come here (4 % crossings, 4 % history ends); inside the
// flight kernel these branches ran with one or two lanes of a warp while the others waited.
//
// boundary_chunk: one warp, 32 consecutive queue entries (lane i: entry base + i; `valid` false past the end).
// Every lane of the warp must call it (ballots, REDUX).
struct BoundaryArgs {
  const double* bounds;
  BankSite* site_scratch;
  uint2* pending_scratch;
  unsigned long long* next_history;
  unsigned long long* scores;
  unsigned long long* square_scores;
};

__device__ __forceinline__ void boundary_chunk(
    const WorldView& w, const RunSpec& run, const EventState& st, const BoundaryArgs& a, uint32_t slot, bool valid,
    unsigned long long* counter_replica) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t lanes_below = (1u << lane) - 1u;
  const bool has_secondaries = run.secondary_capacity > 1;

  Particle p;
  p.event = valid ? MMC_LD(st.event[slot]) : kEvRetired;
  p.px = MMC_LD(st.px[slot]), p.py = MMC_LD(st.py[slot]), p.pz = MMC_LD(st.pz[slot]);
  p.dx = MMC_LD(st.dx[slot]), p.dy = MMC_LD(st.dy[slot]), p.dz = MMC_LD(st.dz[slot]);
  p.energy = MMC_LD(st.energy[slot]);
  p.group = 0;
  p.rng = load_rng(st, slot);
  p.cell = MMC_LD(st.cell[slot]);
  p.surface = MMC_LD(st.surface[slot]);
  uint32_t n_pending = run.n_estimators ? MMC_LD(st.n_pending[slot]) : 0u;
  SiteDeque dq;
  dq.slots = a.site_scratch + static_cast<size_t>(slot) * run.secondary_capacity;
  dq.mask = run.secondary_capacity - 1;
  dq.head = 0;
  dq.count = 0;
  if (valid && has_secondaries) {
    dq.head = MMC_LD(st.dq_head[slot]);
    dq.count = MMC_LD(st.dq_count[slot]);
  }
  ThreadCounters c;

  // ---- the second half of a crossing, and its tallies
  const bool crossing = p.event == kEvCrossPending;
  bool lost = false;
  if (crossing) {
    StepOut o;
    o.error_lost = false;
    finish_crossing(w, p, o);
    lost = o.error_lost;
    c.crossings++;  // count_event: surface_cross or leak (a lost particle is recorded as a leak)
    c.lost += lost;
  }
  uint2* pending = a.pending_scratch + static_cast<size_t>(slot) * run.pending_capacity;
  for (int32_t e = 0; e < run.n_estimators; e++) {
    uint64_t bin = 0;
    const bool hit = crossing && !lost && estimator_score<true>(run.estimators[e], a.bounds, p, bin);
    const unsigned hit_mask = __ballot_sync(kFull, hit);
    if (hit) {
      // incremental CommitHistory: the k-th hit of a history in a bin adds 1 and 2k - 1 (see kernels.cu)
      uint32_t k = 0, s = 0;
      for (; s < n_pending; s++)
        if (pending[s].x == static_cast<uint32_t>(bin)) break;
      if (s < n_pending) {
        k = pending[s].y;
        pending[s].y = k + 1;
      } else if (n_pending < run.pending_capacity) {
        pending[n_pending++] = make_uint2(static_cast<uint32_t>(bin), 1u);
      } else {
        c.capacity++;
      }
      c.scores++;
      const unsigned peers = __match_any_sync(hit_mask, bin);
      const uint32_t sq = __reduce_add_sync(peers, 2u * k + 1u);
      if (lane == static_cast<uint32_t>(__ffs(peers) - 1)) {
        atomicAdd(a.scores + bin, static_cast<unsigned long long>(__popc(peers)));
        atomicAdd(a.square_scores + bin, static_cast<unsigned long long>(sq));
      }
    }
  }

  // ---- a dead particle: the next one of the history's bank, else a new history
  bool alive = valid && is_alive(p.event);
  bool born = false;
  if (valid && !alive && dq.count) {
    dq.count--;
    load_site(dq.slots[(dq.head + dq.count) & dq.mask], p);
    alive = born = true;
    c.births++;
  }
  const bool need = valid && !alive;
  const unsigned need_mask = __ballot_sync(kFull, need);
  unsigned long long claim_base = 0;
  if (need_mask) {
    if (lane == 0) {
      claim_base = *reinterpret_cast<volatile unsigned long long*>(a.next_history);
      if (claim_base < run.n_histories)
        claim_base = atomicAdd(a.next_history, static_cast<unsigned long long>(__popc(need_mask)));
    }
    claim_base = __shfl_sync(kFull, claim_base, 0);
  }
  if (need) {
    const uint64_t idx = claim_base + __popc(need_mask & lanes_below);
    if (idx < run.n_histories) {
      n_pending = 0;  // a new scoring proxy starts empty: FixedSource.cpp:48
      sample_source(run.source, run.seed0 + run.first_history + idx, p);
      alive = born = true;
      c.histories++;
      c.births++;
    } else {
      p.event = kEvRetired;  // no work left for this slot
    }
  }
  if (born) {
    // TransportMethod.cpp:55: p.SetCell(w.FindCellContaining(p.GetPosition())); a newborn outside every Cell is what
    // the fused kernel records as one lost event (a leak)
    p.cell = find_cell(w, p.px, p.py, p.pz);
    if (p.cell < 0) {
      p.event = MMC_EV_LEAK;
      c.events++;
      c.crossings++;
      c.lost++;
    }
  }
  if (valid) {
    MMC_ST(st.px[slot], p.px), MMC_ST(st.py[slot], p.py), MMC_ST(st.pz[slot], p.pz);
    MMC_ST(st.dx[slot], p.dx), MMC_ST(st.dy[slot], p.dy), MMC_ST(st.dz[slot], p.dz);
    MMC_ST(st.energy[slot], p.energy);
    store_rng<true>(st, slot, p.rng);
    MMC_ST(st.cell[slot], p.cell);
    MMC_ST(st.surface[slot], p.surface);
    MMC_ST(st.event[slot], p.event);
    if (run.n_estimators) MMC_ST(st.n_pending[slot], n_pending);
    if (has_secondaries) {
      MMC_ST(st.dq_head[slot], dq.head);
      MMC_ST(st.dq_count[slot], dq.count);
    }
  }
  // ---- counters: packed, one warp reduction per word, lane 0 adds the non-zero ones to this warp's replica
  const uint32_t pa = c.events | (c.collisions << 8) | (c.crossings << 16) | (c.virtuals << 24);
  const uint32_t pb = c.histories | (c.births << 8) | (c.lost << 16) | (c.physics << 24);
  const uint32_t pd = c.scores | (c.capacity << 16);
  const uint32_t sa = __reduce_add_sync(kFull, pa), sb = __reduce_add_sync(kFull, pb), sd = __reduce_add_sync(kFull, pd);
  if (lane == 0) {
    const uint32_t sums[kNumCounters] = {sb & 0xffu, (sb >> 8) & 0xffu, sa & 0xffu, (sa >> 8) & 0xffu, (sa >> 16) & 0xffu, sa >> 24,
                                         sd & 0xffffu, 0u, 0u, (sb >> 16) & 0xffu, sd >> 16, sb >> 24};
#pragma unroll
    for (int k = 0; k < kNumCounters; k++)
      if (sums[k]) atomicAdd(counter_replica + k, static_cast<unsigned long long>(sums[k]));
  }
}

// MMC_EV_MERGE_BOUNDARY = 1 (default): the warps of the persistent S(a,b) kernel work off the boundary queue first, then
// the S(a,b) queue -- the boundary work is a short latency chain over few slots (28 us as a kernel of its own,
// 12 % issue slots); inside the persistent kernel it overlaps with the S(a,b) sampling of the other warps.
// 0: a kernel of its own between the flight and the S(a,b) kernel (kept for profiling it in isolation).
#ifndef MMC_EV_MERGE_BOUNDARY
#define MMC_EV_MERGE_BOUNDARY 1
#endif

#if !MMC_EV_MERGE_BOUNDARY
__global__ void __launch_bounds__(kFlightThreads) event_boundary_kernel(
    const char* __restrict__ world_g, const __grid_constant__ WorldHeader header, const __grid_constant__ RunSpec run,
    const __grid_constant__ EventState st, const __grid_constant__ EventQueues q, uint32_t pass,
    const __grid_constant__ BoundaryArgs args, unsigned long long* counter_replicas) {
  const uint32_t n = q.count[5u + (pass & 1u)];
  const uint32_t i = blockIdx.x * kFlightThreads + threadIdx.x;
  if ((i & ~31u) >= n) return;  // warp-uniform
  const bool valid = i < n;
  const WorldView w(world_g, &header);
  boundary_chunk(w, run, st, args, valid ? q.boundary[i] : 0u, valid,
                 counter_replicas + ((i >> 5) % kCounterReplicas) * kNumCounters);
}
#endif

// ThermalScattering::Scatter (ThermalScattering.cpp:159-171) for the slots the flight kernel queued.
//
// ncu on B200 of the first version of this kernel (one thread per queue entry, every table read through L1): 88 % of
// peak L1 WAVEFRONTS, 98 % L1 hit rate -- every lane gathers its own 80-byte rows, so one load instruction costs up
// to 32 cache-line lookups.  The gather roof, not fp64 issue and not HBM.  Hence one persistent CTA per SM that keeps
// what is gathered out of L1:
//   (1) each lane's two mode rows in 40 registers (ce::RegisterRows, the default, 512 threads) or in a
//       [pair][thread] column of shared memory (ce::SharedRows, 768 threads: 4 conflict-free wavefronts per read);
//   (2) when it fits, the arena of every partition's S*CDF_modes rows in shared memory (WorldHeader::off_sc_arena,
//       44 KB at the reference's table shapes; rows are 80 bytes apart, i.e. 5 sixteen-byte bank groups -- odd -- so
//       rows that differ modulo 8 never conflict).
// Warps claim chunks of 32 queue entries from a counter; state loads and stores are coalesced over the compacted queue.
template <int kKind>
__global__ void __launch_bounds__(tsl_threads_of(kKind), 1) event_tsl_kernel(
    const char* __restrict__ world_g, const __grid_constant__ WorldHeader header, const __grid_constant__ RunSpec run,
    const __grid_constant__ EventState st, const __grid_constant__ EventQueues q, uint32_t pass,
    const __grid_constant__ BoundaryArgs args, unsigned long long* counter_replicas) {
  constexpr bool kSharedSc = kKind == kTslRowsSharedSc;
  constexpr uint32_t kThreads = tsl_threads_of(kKind);
  extern __shared__ __align__(16) char smem[];
  [[maybe_unused]] double2* s_rows = reinterpret_cast<double2*>(smem);  // SharedRows: double2[10][kTslThreads]
  [[maybe_unused]] char* s_sc = smem + kTslRowBytes;
  const uint32_t parity = pass & 1u;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    q.count[parity] = 0;             // the alive queue this pass consumed: the next pass appends to it
    q.count[2u + (parity ^ 1u)] = 0; // the S(a,b) queue of the next pass
    q.count[5u + (parity ^ 1u)] = 0; // the boundary queue of the next pass
  }
  const uint32_t n = q.count[2u + parity];
  [[maybe_unused]] constexpr uint32_t kWarps = kThreads / 32;
#if !MMC_EV_MERGE_BOUNDARY
  if (blockIdx.x * kWarps * 32u >= n) return;  // CTA-uniform: not even the first warp has work
#endif
  const WorldView w(world_g, &header);
  if (kSharedSc) {
    const uint4* src = reinterpret_cast<const uint4*>(world_g + w.h->off_sc_arena);
    uint4* dst = reinterpret_cast<uint4*>(s_sc);
    for (uint32_t k = threadIdx.x; k < w.h->sc_arena_bytes / 16; k += kThreads) dst[k] = __ldg(src + k);
    __syncthreads();
  }
  const uint32_t lane = threadIdx.x & 31u;
#if MMC_EV_MERGE_BOUNDARY
  {
    // the boundary queue first: chunks of 32 entries claimed from a counter (zeroed by the flight kernel)
    const uint32_t n_boundary = q.count[5u + parity];
    unsigned long long* replica = counter_replicas + ((blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) % kCounterReplicas) * kNumCounters;
    while (true) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(&q.count[7], 32u);
      base = __shfl_sync(kFull, base, 0);
      if (base >= n_boundary) break;
      const bool valid = base + lane < n_boundary;
      boundary_chunk(w, run, st, args, valid ? q.boundary[base + lane] : 0u, valid, replica);
    }
  }
#endif
  auto make_rows = [&]() {
    if constexpr (kKind == kTslDense || kKind == kTslDirect) return ce::DenseRows{};
#if MMC_EV_TSL_ROWS_IN_REGS
    else return ce::RegisterRows<kSharedSc>(s_sc, w.h->off_sc_arena);
#else
    else return ce::SharedRows<kTslThreads, kSharedSc>(s_rows, s_sc, w.h->off_sc_arena);
#endif
  };
  [[maybe_unused]] auto rows = make_rows();
  // warps claim chunks of 32 queue entries from one counter: a scatter takes 10-30 rounds of reconstructions, so a
  // static split leaves the unlucky warps running alone at the end of every pass (chunks of 64: no faster)
  // Guided self-scheduling: a claim takes 1/(2 x warps) of what is left, between 32 and 256 entries -- a quarter of the
  // same-address atomics of fixed 32-entry chunks (ncu r02a: 11 % of this kernel's stall samples sat on that atomic),
  // and still single chunks at the end of the pass, where balance matters.
  // The first MMC_EV_TSL_STATIC_PCT per cent of the queue are dealt out without any atomic -- warp k takes entries
  // [k * share, (k + 1) * share) -- and only the rest is claimed dynamically: that rest evens out what the static
  // shares leave uneven (a share is ~10 chunks, the spread of their sum a few per cent), with a fraction of the claims.
  const uint32_t total_warps = gridDim.x * (blockDim.x >> 5);  // (a thin CTA when few slots are alive: launch_event_pass)
  const uint32_t share = static_cast<uint32_t>((static_cast<uint64_t>(n) * MMC_EV_TSL_STATIC_PCT / 100u) / total_warps) & ~31u;
  const uint32_t static_total = share * total_warps;
  uint32_t claim_next = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * share, claim_end = claim_next + share;
  while (true) {
    if (claim_next >= claim_end) {
      uint32_t base = 0, size = 0;
      if (lane == 0) {
        const uint32_t seen = static_total + *reinterpret_cast<volatile unsigned int*>(&q.count[4]);
        const uint32_t left = seen < n ? n - seen : 0u;
        size = left / (2u * total_warps);
        size = size > 256u ? 256u : size < 32u ? 32u : size & ~31u;
        base = static_total + atomicAdd(&q.count[4], size);
      }
      base = __shfl_sync(kFull, base, 0);
      size = __shfl_sync(kFull, size, 0);
      if (base >= n) break;
      claim_next = base;
      claim_end = base + size < n ? base + size : n;
    }
    const uint32_t i = claim_next + lane;
    claim_next += 32u;
    if (i < claim_end) {  // no early `continue`: every lane must come back to the shuffle above
    const uint32_t slot = q.tsl[i];
    Particle p;
    p.energy = MMC_LD(st.energy[slot]);
    p.rng = load_rng(st, slot);
    const int32_t cell = MMC_LD(st.cell[slot]);
    // ScalarField::at of a constant field is its value (ce::cell_temperature): the flight kernel does not pass it on
    const double T = kKind == kTslDirect ? __ldg(w.at<double>(w.h->off_cell_field_param) + 6 * cell) : MMC_LD(st.tsl_T[slot]);
    const TslTable& t = *w.at<TslTable>(MMC_LD(st.tsl_off[slot]));
    bool error = false;
    double mu = 0, E_p = 0;
    const int32_t eval_slot = ce::cell_eval_slot(w, cell);
    if constexpr (kKind == kTslDirect) {
      // WorldHeader::tsl_all_direct promises both; a collision that breaks the promise is reported, not sampled
      if (eval_slot >= 0 && t.direct) ce::tsl_sample_direct<true>(w, t, p.rng, p.energy, T, eval_slot, error, mu, E_p);
      else error = true;
    } else if constexpr (kKind == kTslDense) {
      // every partition has its dense table (WorldHeader::tsl_all_dense): a row is evaluated (one load per node) or
      // two rows of the dense table + the interpolation in T -- the samplers written straight down for both
      ce::tsl_sample_direct<false>(w, t, p.rng, p.energy, T, eval_slot, error, mu, E_p);
    } else {
      ce::tsl_sample(w, t, p.rng, p.energy, T, eval_slot, error, rows, mu, E_p);
    }
    if (!error) {
      // the direction is only needed now: it stays in memory while the sampler's state fills the registers
      p.dx = MMC_LD(st.dx[slot]), p.dy = MMC_LD(st.dy[slot]), p.dz = MMC_LD(st.dz[slot]);
      ce::particle_scatter(p, mu, E_p);
      MMC_ST(st.dx[slot], p.dx), MMC_ST(st.dy[slot], p.dy), MMC_ST(st.dz[slot], p.dz);
      MMC_ST(st.energy[slot], p.energy);
    }
    store_rng<false>(st, slot, p.rng);
    if (error) {
      // the reference throws here (-> std::terminate); the flight kernel counted the collision already
      MMC_ST(st.event[slot], MMC_EV_CAPTURE);
      unsigned long long* mine = counter_replicas + (blockIdx.x % kCounterReplicas) * kNumCounters;
      atomicAdd(mine + 3, ~0ull);  // n_collisions - 1
      atomicAdd(mine + 11, 1ull);  // n_physics_errors + 1
    }
    }
  }
}

__global__ void event_finish_kernel(const unsigned long long* counter_replicas, mmc_counters* counters) {
  const uint32_t k = threadIdx.x;
  if (k >= kNumCounters) return;
  unsigned long long sum = 0;
  for (int r = 0; r < kCounterReplicas; r++) sum += counter_replicas[r * kNumCounters + k];
  if (sum) atomicAdd(reinterpret_cast<unsigned long long*>(counters) + k, sum);
}

}  // namespace

static_assert(sizeof(mmc_counters) == kNumCounters * sizeof(uint64_t), "mmc_counters layout");

cudaError_t launch_event_init(const EventState& st, const EventQueues& q, uint32_t n_slots,
                              unsigned long long* counter_replicas, cudaStream_t stream) {
  const uint32_t n_padded = event_padded_slots(n_slots);
  const uint32_t n = n_padded > kCounterReplicas * kNumCounters ? n_padded : kCounterReplicas * kNumCounters;
  event_init_kernel<<<(n + 255) / 256, 256, 0, stream>>>(st, q, n_slots, n_padded, counter_replicas);
  return cudaGetLastError();
}

cudaError_t launch_event_pass(
    const char* world_d, const WorldHeader& header, const RunSpec& run, const double* bounds_d, const EventState& st, const EventQueues& q,
    uint32_t pass, uint32_t alive_upper_bound, BankSite* site_scratch, uint2* pending_scratch,
    unsigned long long* next_history, unsigned long long* scores, unsigned long long* square_scores,
    unsigned long long* counter_replicas, const EventTslConfig& tsl, cudaStream_t stream, const cudaEvent_t* marks) {
  const uint32_t blocks = (alive_upper_bound + kFlightThreads - 1) / kFlightThreads;
  if (blocks == 0) return cudaSuccess;
  if (run.tracking == MMC_TRACK_CELL_DELTA)
    event_flight_kernel<MMC_TRACK_CELL_DELTA><<<blocks, kFlightThreads, 0, stream>>>(
        world_d, header, run, st, q, pass, site_scratch, counter_replicas);
  else
    event_flight_kernel<MMC_TRACK_SURFACE><<<blocks, kFlightThreads, 0, stream>>>(
        world_d, header, run, st, q, pass, site_scratch, counter_replicas);
  if (marks) cudaEventRecord(marks[0], stream);
  BoundaryArgs args;
  args.bounds = bounds_d;
  args.site_scratch = site_scratch;
  args.pending_scratch = pending_scratch;
  args.next_history = next_history;
  args.scores = scores;
  args.square_scores = square_scores;
#if !MMC_EV_MERGE_BOUNDARY
  event_boundary_kernel<<<blocks, kFlightThreads, 0, stream>>>(world_d, header, run, st, q, pass, args, counter_replicas);
#endif
  if (marks) cudaEventRecord(marks[1], stream);
  // S(a,b) kernel: persistent, at most one CTA per SM, no more CTAs than the queues can feed
  const uint32_t per_cta = header.tsl_all_direct ? kTslDirectThreads : header.tsl_all_dense ? kTslDenseThreads : kTslThreads;
  uint32_t tsl_blocks = (alive_upper_bound + per_cta - 1) / per_cta;
  if (tsl_blocks > tsl.sm_count) tsl_blocks = tsl.sm_count;
  // Few slots alive (the ramp-down of a batch): a pass costs the latency of one chunk, and that is lowest when the
  // chunks are spread over every SM -- one CTA per SM with as few warps as the work needs -- instead of filling a few
  // SMs with 32 warps each.  (Only the kinds without per-CTA shared-memory rows.)
  uint32_t thin_threads = 0;
  if ((header.tsl_all_direct || header.tsl_all_dense) && alive_upper_bound < per_cta * tsl.sm_count) {
    thin_threads = ((alive_upper_bound + tsl.sm_count - 1) / tsl.sm_count + 31u) & ~31u;
    if (thin_threads < 64u) thin_threads = 64u;
    tsl_blocks = tsl.sm_count;
  }
  if (header.tsl_all_direct)
    event_tsl_kernel<kTslDirect><<<tsl_blocks, thin_threads ? thin_threads : kTslDirectThreads, 0, stream>>>(world_d, header, run, st, q, pass, args, counter_replicas);
  else if (header.tsl_all_dense)
    event_tsl_kernel<kTslDense><<<tsl_blocks, thin_threads ? thin_threads : kTslDenseThreads, 0, stream>>>(world_d, header, run, st, q, pass, args, counter_replicas);
  else if (tsl.shared_sc)
    event_tsl_kernel<kTslRowsSharedSc><<<tsl_blocks, kTslThreads, kTslRowBytes + tsl.sc_arena_bytes, stream>>>(
        world_d, header, run, st, q, pass, args, counter_replicas);
  else
    event_tsl_kernel<kTslRowsGlobalSc><<<tsl_blocks, kTslThreads, kTslRowBytes, stream>>>(world_d, header, run, st, q, pass, args, counter_replicas);
  return cudaGetLastError();
}

int event_kernels_per_pass() { return MMC_EV_MERGE_BOUNDARY ? 2 : 3; }

cudaError_t configure_event_tsl(uint32_t sc_arena_bytes, size_t smem_optin, uint32_t sm_count, EventTslConfig& out) {
  out.sm_count = sm_count;
  out.sc_arena_bytes = sc_arena_bytes;
  out.shared_sc = sc_arena_bytes > 0 && kTslRowBytes + sc_arena_bytes <= smem_optin;
  cudaError_t e = cudaFuncSetAttribute(event_tsl_kernel<kTslRowsGlobalSc>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kTslRowBytes));
  if (e == cudaSuccess && out.shared_sc)
    e = cudaFuncSetAttribute(event_tsl_kernel<kTslRowsSharedSc>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(kTslRowBytes + sc_arena_bytes));
  return e;
}

cudaError_t launch_event_finish(const unsigned long long* counter_replicas, mmc_counters* counters, cudaStream_t stream) {
  event_finish_kernel<<<1, 32, 0, stream>>>(counter_replicas, counters);
  return cudaGetLastError();
}

}  // namespace MMC_VARIANT_NS
}  // namespace mmc
