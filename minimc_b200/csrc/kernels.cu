// sm_100a kernels of the history transport loop.
//
// fixed_source_kernel: one fused, persistent kernel that replaces
// FixedSource::StartWorker (FixedSource.cpp:40-77).  Particle state lives in
// registers for its whole life; the world tables are staged once per CTA into
// shared memory.  Lanes whose history has ended are refilled in place: every
// loop iteration first tops the warp up with new histories (claimed in chunks
// from one global counter with a single atomic per chunk, handed out with
// ballot/popc), then every lane advances its particle by exactly one event.
// That keeps all 32 lanes on the expensive fp64 event code instead of idling
// until the longest history of the warp finishes.
//
// Tallies: the `current` score is 0/1 per event, so Sigma s and
// Sigma_h (Sigma_e s)^2 are integers.  Each lane keeps the per-history pending
// table of ScorableProxy (Scorable.cpp:81-99) as (bin, hits) pairs in a private
// slice of global scratch; the k-th hit of a history in a bin adds 1 to
// scores[bin] and k^2-(k-1)^2 = 2k-1 to square_scores[bin], which is what
// CommitHistory (Scorable.cpp:101-106) yields after the history.  Lanes that
// hit the same bin in the same iteration are combined with __match_any_sync and
// one 64-bit integer atomic per distinct bin; integer sums make the result
// independent of scheduling, GPU count and atomic ordering.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "kernel_common.cuh"

// resident CTAs per SM the register allocation is tuned for (measured choices)
#ifndef MMC_CE_BLOCKS_PER_SM
#define MMC_CE_BLOCKS_PER_SM 3
#endif
#ifndef MMC_MG_BLOCKS_PER_SM
#define MMC_MG_BLOCKS_PER_SM 3
#endif

namespace mmc {
namespace MMC_VARIANT_NS {  // lcg or ctr: this file is compiled once per RNG mode (kernels.h)

namespace {

__device__ __forceinline__ void flush_counter(uint64_t* dst, uint32_t v) {
  v = __reduce_add_sync(kFull, v);
  if ((threadIdx.x & 31) == 0 && v)
    atomicAdd(reinterpret_cast<unsigned long long*>(dst), static_cast<unsigned long long>(v));
}

// Stage the world blob into shared memory (16-byte vectors).
__device__ __forceinline__ const char* stage_world(const char* world_g, uint32_t bytes, bool in_smem, char* smem) {
  if (!in_smem) return world_g;
  const uint4* src = reinterpret_cast<const uint4*>(world_g);
  uint4* dst = reinterpret_cast<uint4*>(smem);
  for (uint32_t i = threadIdx.x; i < bytes / 16; i += blockDim.x) dst[i] = src[i];
  __syncthreads();
  return smem;
}

}  // namespace

// kGeneration = false: fixed source (above).  kGeneration = true: one generation
// of the k-eigenvalue power iteration (DESIGN.md "k-eigenvalue"): history idx
// starts from bank.in[idx] instead of Source::Sample, and fission secondaries
// are not followed but appended to the fission bank -- one warp-aggregated
// atomic per warp claims the slots, each parent's run is contiguous, and
// (count, start) per parent let order_bank_kernel restore the deterministic
// (parent index, creation ordinal) order.
//
// kPerturb = true: differential-operator sensitivities (Perturbation.cpp, Sensitivity.cpp).  Every particle carries
// one indirect effect per perturbation, reset when the particle is taken from the bank (FixedSource.cpp:65) and
// grown by every Stream (transport.cuh perturb_stream); where an estimator scores, each of its sensitivities scores
// indirect effect x estimator score into the sensitivity's per-history pending table (ScorableProxy::Score,
// Scorable.cpp:81-99), committed -- sum and sum squared -- when the history ends (CommitHistory, :101-106).  These
// tallies are real-valued: fp64 atomics, order-dependent in the last bits exactly as the reference's own worker merge.
//
// kDrain = true: the instantiation the event-split schedule hands its last few thousand histories to (ResumeIO).  Few
// warps, each alone on its scheduler: the time of that kernel is the longest remaining history times the latency of
// one event, so it is compiled for one CTA per SM -- all the registers it wants, no spills -- instead of three.
template <int kTracking, bool kCE, bool kGeneration, bool kPerturb, bool kDrain = false>
__global__ void __launch_bounds__(kThreadsPerBlock, kDrain ? 1 : kCE ? MMC_CE_BLOCKS_PER_SM : MMC_MG_BLOCKS_PER_SM) fixed_source_kernel(
    const char* __restrict__ world_g, const __grid_constant__ RunSpec run, const double* __restrict__ bounds,
    BankSite* __restrict__ site_scratch, uint2* __restrict__ pending_scratch, unsigned long long* next_history,
    unsigned long long* scores, unsigned long long* square_scores, mmc_counters* counters,
    const __grid_constant__ GenerationIO bank, const __grid_constant__ ResumeIO resume,
    const __grid_constant__ SensitivityIO sens) {
  extern __shared__ __align__(16) char smem[];
  const WorldView w(stage_world(world_g, run.world_bytes, run.world_in_smem != 0, smem));

  const uint32_t lane = threadIdx.x & 31;
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  // resume (ResumeIO): this thread continues slot resume.slots[tid] of the event-split schedule
  const bool resuming = !kGeneration && resume.n != nullptr;
  const bool adopt = resuming && tid < *resume.n;
  const size_t scratch = adopt ? resume.slots[tid] : tid;
  SiteDeque dq;
  dq.slots = site_scratch + scratch * run.secondary_capacity;
  dq.mask = run.secondary_capacity - 1;
  dq.head = 0;
  dq.count = 0;
  uint2* pending = pending_scratch + scratch * run.pending_capacity;
  uint32_t n_pending = 0;
  // sensitivities: the particle's indirect effects and the history's pending (bin, sum) entries
  double indirect[kPerturb ? kMaxPerturbations : 1];
  SensitivityPending* sens_pending = kPerturb ? sens.pending + scratch * run.sens_pending_capacity : nullptr;
  uint32_t n_sens_pending = 0;
  const PerturbContext pc{run.perturbed_nuclide, run.n_perturbations, indirect};
  auto reset_indirect = [&]() {
    if (kPerturb)
      for (int k = 0; k < kMaxPerturbations; k++) indirect[k] = 0.0;
  };
  auto commit_sensitivities = [&]() {  // ScorableProxy::CommitHistory of every sensitivity proxy
    if (kPerturb) {
      for (uint32_t k = 0; k < n_sens_pending; k++) {
        const SensitivityPending e = sens_pending[k];
        atomicAdd(sens.scores + e.bin, e.sum);
        atomicAdd(sens.square_scores + e.bin, __dmul_rn(e.sum, e.sum));
      }
      n_sens_pending = 0;
    }
  };
  reset_indirect();

  Particle p;
  p.event = MMC_EV_CAPTURE;  // "dead": forces a refill
  bool done = resuming && !adopt;  // no more work for this lane (a resumed run has no scratch for extra threads)
  if (adopt) {
    const EventState& st = resume.st;
    p.event = st.event[scratch];
    if (p.event == kEvRetired) p.event = MMC_EV_CAPTURE;  // a slot that already found no history left: dead
    p.px = st.px[scratch], p.py = st.py[scratch], p.pz = st.pz[scratch];
    p.dx = st.dx[scratch], p.dy = st.dy[scratch], p.dz = st.dz[scratch];
    p.energy = st.energy[scratch];
    p.group = 0;
    p.rng.x = st.rng[scratch];
#if MMC_COUNTER_RNG
    p.rng.k0 = st.rng_k0[scratch], p.rng.k1 = st.rng_k1[scratch];
#endif
    p.cell = st.cell[scratch];
    p.surface = st.surface[scratch];
    if (run.n_estimators) n_pending = st.n_pending[scratch];
    if (run.secondary_capacity > 1) {
      dq.head = st.dq_head[scratch];
      dq.count = st.dq_count[scratch];
    }
  }
  uint64_t w_next = 0, w_end = 0;  // warp-uniform chunk of history indices
  uint64_t history = 0;            // index of this lane's current history
  // isotropic direction owed to this lane: bit 0 = its multigroup scatter of the previous iteration, bit 1 = it
  // was just born from an isotropic source.  Both are the same code on the lane's rng, so they run converged.
  uint32_t owed = 0;
  ThreadCounters c;
  [[maybe_unused]] unsigned long long k_fixed = 0;  // this thread's collision-estimator scores (generations)
  // generations: the fission secondaries this lane's dead particle still owes the bank.  The reference makes them in
  // a loop inside Fission (Multigroup.cpp:267-287, ContinuousReaction.cpp:252-265); with one lane in four fissioning
  // per collision that loop would run in every warp at a quarter of its lanes.  Here the fission only draws its yield
  // and claims its run of the bank; the sites are made one per loop iteration from the parent's rng, in the same
  // draw order, converged with the isotropic directions the other lanes owe (scatter, birth).
  [[maybe_unused]] uint32_t sites_owed = 0, sites_made = 0;
  [[maybe_unused]] unsigned long long site_start = 0;
  [[maybe_unused]] int32_t fission_nuclide = -1;

  while (true) {
    // ---- refill: next particle of the current history, else a new history
    bool alive = is_alive(p.event);
    if (!kGeneration && !alive && !done && !dq.empty()) {
      // bank.back(): FixedSource.cpp:63-71
      dq.count--;
      load_site(dq.slots[(dq.head + dq.count) & dq.mask], p);
      reset_indirect();  // bank.back().SetPerturbations(perturbations)
      alive = true;
      c.births++;
    }
    const bool need = !alive && !done && !(kGeneration && sites_owed);
    const unsigned need_mask = __ballot_sync(kFull, need);
    if (need_mask) {
      const uint32_t n = __popc(need_mask);
      const uint32_t rank = __popc(need_mask & ((1u << lane) - 1u));
      const uint64_t avail = w_end - w_next;
      uint64_t idx;
      bool valid;
      if (n > avail) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(next_history, static_cast<unsigned long long>(run.chunk));
        base = __shfl_sync(kFull, base, 0);
        const uint64_t new_begin = base < run.n_histories ? base : run.n_histories;
        const uint64_t new_end = base + run.chunk < run.n_histories ? base + run.chunk : run.n_histories;
        if (rank < avail) {
          idx = w_next + rank;
          valid = true;
        } else {
          idx = new_begin + (rank - avail);
          valid = idx < new_end;
        }
        w_next = new_begin + (n - avail);
        w_end = new_end;
        if (w_next > w_end) w_next = w_end;
      } else {
        idx = w_next + rank;
        valid = true;
        w_next += n;
      }
      if (need) {
        commit_sensitivities();  // the previous history of this lane is over
        if (valid) {
          // scoring_proxy of the previous history is committed (incrementally);
          // a new proxy starts empty: FixedSource.cpp:48
          n_pending = 0;
          reset_indirect();
          history = idx;
          if (kGeneration) load_site(bank.in[idx], p);
          else if (sample_source(run.source, run.seed0 + run.first_history + idx, p, true)) owed |= 2u;
          alive = true;
          c.histories++;
          c.births++;
        } else {
          done = true;
        }
      }
    }
    if (__all_sync(kFull, done)) break;

    // ---- isotropic directions owed to scattered and newly born lanes (Point.cpp:89-96)
    if (kGeneration) {
      if (owed || sites_owed) {
        const bool make = sites_owed != 0;
        // Multigroup::Fission: the secondary's group from chi (no group found: no secondary, Multigroup.cpp:278-283)
        uint64_t child_group = 0;
        bool site_ok = true;
        if (make && !kCE) {
          const int32_t G = w.h->n_groups;
          const double* chi = w.at<double>(w.h->off_mg_chi) + (static_cast<size_t>(fission_nuclide) * G + (p.group - 1)) * G;
          const double t = p.rng.canonical();
          double a = 0.0;
          site_ok = false;
          for (int32_t g = 0; g < G; g++) {
            a = __dadd_rn(a, chi[g]);
            if (a > t) {
              child_group = static_cast<uint64_t>(g + 1);
              site_ok = true;
              break;
            }
          }
        }
        if (site_ok) {
          double ix, iy, iz;
          isotropic_direction(p.rng, ix, iy, iz);
          if (make) {
            BankSite s;
            s.position[0] = p.px, s.position[1] = p.py, s.position[2] = p.pz;
            s.direction[0] = ix, s.direction[1] = iy, s.direction[2] = iz;
            s.energy_bits = kCE ? static_cast<uint64_t>(__double_as_longlong(p.energy)) : child_group;
            store_seed(s, p.rng.spawn());  // Particle::BankSecondaries, Particle.cpp:96-100
            bank.out[site_start + sites_made] = s;
            sites_made++;
          } else {
            p.dx = ix, p.dy = iy, p.dz = iz;
            if (owed & 2u) finish_source(p);
          }
        }
        if (make) {
          if (--sites_owed == 0) {
            bank.child_count[history] = sites_made;
            c.secondaries += sites_made;
            c.banked += sites_made;
          }
        } else {
          owed = 0;
        }
      }
    } else if (owed) {
      isotropic_direction(p.rng, p.dx, p.dy, p.dz);
      if (owed & 2u) finish_source(p);
      owed = 0;
    }

    // ---- one event per live lane
    StepOut o;
    o.secondaries = 0;
    o.need_direction = false;
    o.error_physics = o.error_capacity = o.error_lost = false;
    o.need_tsl = o.need_cross = false;
    if (alive) {
      if (p.cell < 0) {
        // TransportMethod.cpp:55: p.SetCell(w.FindCellContaining(p.GetPosition()))
        p.cell = find_cell(w, p.px, p.py, p.pz);
        if (p.cell < 0) {
          o.error_lost = true;
          p.event = MMC_EV_LEAK;
        }
      }
      [[maybe_unused]] const uint64_t group_before = p.group;
      [[maybe_unused]] const double energy_before = p.energy;
      if (p.cell >= 0) transport_step<kTracking, kCE, true, false, kPerturb, false, kGeneration>(w, p, dq, o, &pc);
      if (kGeneration && bank.k_collision &&
          (p.event == MMC_EV_SCATTER || p.event == MMC_EV_CAPTURE || p.event == MMC_EV_FISSION) && !o.error_physics) {
        // collision estimator of k at the collision site, pre-collision energy (see implicit_fission_score)
        const int32_t mat = w.at<int32_t>(w.h->off_cell_material)[p.cell];
        const double T = kCE ? ce::cell_temperature(w, p.cell, p.px, p.py, p.pz) : 0.0;
        k_fixed += static_cast<unsigned long long>(
            __double2ll_rn(__dmul_rn(implicit_fission_score<kCE>(w, mat, group_before, energy_before, T), 268435456.0)));
      }
      if (o.need_direction) owed |= 1u;
      count_event(c, p, o);
    }
    if (kGeneration) {
      // ---- a fission claims its run of the fission bank: one aggregated atomic per warp, each parent's run contiguous
      const uint32_t mine = alive ? o.pending_yield : 0u;
      const unsigned any = __ballot_sync(kFull, mine != 0);
      if (any) {
        uint32_t inclusive = mine;
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t up = __shfl_up_sync(kFull, inclusive, d);
          if (lane >= static_cast<uint32_t>(d)) inclusive += up;
        }
        const uint32_t total = __shfl_sync(kFull, inclusive, 31);
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(bank.n_out, static_cast<unsigned long long>(total));
        base = __shfl_sync(kFull, base, 0);
        if (mine) {
          const unsigned long long start = base + (inclusive - mine);
          if (start + mine <= bank.capacity) {
            site_start = start;
            sites_owed = mine;
            sites_made = 0;
            fission_nuclide = o.fission_nuclide;
            bank.child_start[history] = start;
          } else {
            c.capacity++;
          }
        }
      }
    }

    // ---- EstimatorSetProxy::Score(p): TransportMethod.cpp:74
    for (int32_t e = 0; e < run.n_estimators; e++) {
      uint64_t bin = 0;
      const bool hit = alive && !o.error_lost && estimator_score<kCE>(run.estimators[e], bounds, p, bin);
      const unsigned hit_mask = __ballot_sync(kFull, hit);
      if (hit) {
        // hits of this history in this bin so far
        uint32_t k = 0, slot = 0;
        for (; slot < n_pending; slot++)
          if (pending[slot].x == static_cast<uint32_t>(bin)) break;
        if (slot < n_pending) {
          k = pending[slot].y;
          pending[slot].y = k + 1;
        } else if (n_pending < run.pending_capacity) {
          pending[n_pending++] = make_uint2(static_cast<uint32_t>(bin), 1u);
        } else {
          c.capacity++;
        }
        c.scores++;
        const unsigned peers = __match_any_sync(hit_mask, bin);
        const uint32_t sq = __reduce_add_sync(peers, 2u * k + 1u);
        if (lane == static_cast<uint32_t>(__ffs(peers) - 1)) {
          atomicAdd(scores + bin, static_cast<unsigned long long>(__popc(peers)));
          atomicAdd(square_scores + bin, static_cast<unsigned long long>(sq));
        }
        if (kPerturb) {
          // CurrentTotalCrossSectionSensitivity::GetScore = indirect effect x estimator score (= 1 here);
          // ScorableProxy::Score skips a zero score and otherwise adds to the history's entry of the bin
          for (int32_t si = 0; si < run.n_sensitivities; si++) {
            const SensitivitySpec& sp = run.sensitivities[si];
            if (sp.estimator != e) continue;
            const double value = indirect[sp.perturbation];
            if (value == 0) continue;
            const uint32_t sens_bin = static_cast<uint32_t>(sp.offset + (bin - run.estimators[e].offset));
            uint32_t slot_i = 0;
            for (; slot_i < n_sens_pending; slot_i++)
              if (sens_pending[slot_i].bin == sens_bin) break;
            if (slot_i < n_sens_pending) {
              sens_pending[slot_i].sum = __dadd_rn(sens_pending[slot_i].sum, value);
            } else if (n_sens_pending < run.sens_pending_capacity) {
              sens_pending[n_sens_pending].bin = sens_bin;
              sens_pending[n_sens_pending].sum = value;
              n_sens_pending++;
            } else {
              c.capacity++;
            }
          }
        }
      }
    }
  }
  commit_sensitivities();
  if (kGeneration && bank.k_collision) {
    for (int d = 16; d > 0; d >>= 1) k_fixed += __shfl_down_sync(kFull, k_fixed, d);
    if (lane == 0 && k_fixed) atomicAdd(bank.k_collision, k_fixed);
  }

  flush_counter(&counters->n_histories, c.histories);
  flush_counter(&counters->n_births, c.births);
  flush_counter(&counters->n_events, c.events);
  flush_counter(&counters->n_collisions, c.collisions);
  flush_counter(&counters->n_crossings, c.crossings);
  flush_counter(&counters->n_virtual, c.virtuals);
  flush_counter(&counters->n_scores, c.scores);
  flush_counter(&counters->n_secondaries, c.secondaries);
  flush_counter(&counters->n_banked, c.banked);
  flush_counter(&counters->n_lost, c.lost);
  flush_counter(&counters->n_capacity_overflow, c.capacity);
  flush_counter(&counters->n_physics_errors, c.physics);
}

// Parity hook: one thread walks histories sequentially and records every event
// in the reference's order (see mmc_trace_histories).
template <int kTracking, bool kCE>
__global__ void trace_kernel(
    const char* __restrict__ world_g, const __grid_constant__ RunSpec run, BankSite* site_scratch,
    mmc_event_record* records, unsigned long long cap, unsigned long long* n_records, mmc_counters* counters) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const WorldView w(world_g);
  SiteDeque dq;
  dq.slots = site_scratch;
  dq.mask = run.secondary_capacity - 1;
  unsigned long long n = 0;
  ThreadCounters c;
  auto emit = [&](uint64_t history, uint32_t particle, const Particle& p) {
    if (n < cap) {
      mmc_event_record& r = records[n];
      r.history = history;
      r.particle = particle;
      r.event = p.event;
      r.group = kCE ? 0 : p.group;
      r.energy = kCE ? p.energy : 0.0;
      r.cell = p.cell;
      r.surface = p.surface;
      r.position[0] = p.px;
      r.position[1] = p.py;
      r.position[2] = p.pz;
      r.direction[0] = p.dx;
      r.direction[1] = p.dy;
      r.direction[2] = p.dz;
      r.rng_state = p.rng.state64();
    }
    n++;
  };
  for (uint64_t h = 0; h < run.n_histories; h++) {
    const uint64_t history = run.first_history + h;
    dq.head = 0;
    dq.count = 0;
    Particle p;
    sample_source(run.source, run.seed0 + history, p);
    c.histories++;
    uint32_t ordinal = 0;
    while (true) {
      c.births++;
      emit(history, ordinal, p);
      StepOut o;
      p.cell = find_cell(w, p.px, p.py, p.pz);
      if (p.cell < 0) {
        c.lost++;
      } else {
        while (is_alive(p.event)) {
          transport_step<kTracking, kCE>(w, p, dq, o);
          count_event(c, p, o);
          emit(history, ordinal, p);
        }
      }
      if (dq.empty()) break;
      dq.count--;
      load_site(dq.slots[(dq.head + dq.count) & dq.mask], p);
      ordinal++;
    }
  }
  *n_records = n;
  counters->n_histories = c.histories;
  counters->n_births = c.births;
  counters->n_events = c.events;
  counters->n_collisions = c.collisions;
  counters->n_crossings = c.crossings;
  counters->n_virtual = c.virtuals;
  counters->n_secondaries = c.secondaries;
  counters->n_lost = c.lost;
  counters->n_capacity_overflow = c.capacity;
  counters->n_physics_errors = c.physics;
}

__global__ void test_math_kernel(int fn, const double* x, double* out0, double* out1, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (fn == 0) {
    out0[i] = glibc::log(x[i]);
  } else if (fn == 1) {
    glibc::sincos(x[i], &out0[i], &out1[i]);
  } else if (fn == 2) {
    out0[i] = glibc::sin(x[i]);
  } else if (fn == 3) {
    out0[i] = glibc::cos(x[i]);
  } else if (fn == 5) {
    glibc::sin_and_cos(x[i], &out0[i], &out1[i]);
  } else {
    Rng rng = Rng::seeded(static_cast<uint64_t>(x[i]));
    out0[i] = rng.canonical();
  }
}

// ---- k-eigenvalue bank kernels -------------------------------------------------
// Initial source bank: site i = Source::Sample(seed0 + first + i), kept as a
// bank site (KEigenvalue.cpp:29-33 samples seeds 1..batchsize).
__global__ void source_bank_kernel(const __grid_constant__ RunSpec run, BankSite* bank) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= run.n_histories) return;
  Particle p;
  sample_source(run.source, run.seed0 + run.first_history + i, p);
  BankSite s;
  s.position[0] = p.px, s.position[1] = p.py, s.position[2] = p.pz;
  s.direction[0] = p.dx, s.direction[1] = p.dy, s.direction[2] = p.dz;
  s.energy_bits = run.continuous_energy ? static_cast<uint64_t>(__double_as_longlong(p.energy)) : p.group;
  store_seed(s, p.rng.stream());  // the particle's generator has made no draw yet: its seed is its state
  bank[i] = s;
}

// Exclusive scan of child_count over the parents (three launches: block sums,
// scan of the block sums, per-block scan + gather of each parent's run), which
// puts the fission bank in (parent index, creation ordinal) order whatever
// order the warps claimed their slots in.
constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;  // parents per thread
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* smem_warp, uint32_t& block_total) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inclusive = v;
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t up = __shfl_up_sync(kFull, inclusive, d);
    if (lane >= static_cast<uint32_t>(d)) inclusive += up;
  }
  if (lane == 31) smem_warp[warp] = inclusive;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < kScanThreads / 32 ? smem_warp[lane] : 0u;
    uint32_t winc = w;
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t up = __shfl_up_sync(kFull, winc, d);
      if (lane >= static_cast<uint32_t>(d)) winc += up;
    }
    if (lane < kScanThreads / 32) smem_warp[lane] = winc - w;
    if (lane == 31) smem_warp[32] = winc;
  }
  __syncthreads();
  block_total = smem_warp[32];
  return smem_warp[warp] + inclusive - v;
}

__global__ void __launch_bounds__(kScanThreads) bank_block_sums_kernel(
    const uint32_t* __restrict__ child_count, uint64_t n_parents, unsigned long long* block_sums) {
  __shared__ uint32_t smem_warp[33];
  const uint64_t first = static_cast<uint64_t>(blockIdx.x) * kScanTile + threadIdx.x * kScanItems;
  uint32_t v = 0;
  for (int k = 0; k < kScanItems; k++)
    if (first + k < n_parents) v += child_count[first + k];
  uint32_t total;
  block_exclusive_scan(v, smem_warp, total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) bank_scan_block_sums_kernel(unsigned long long* block_sums, uint32_t n_blocks,
                                                                           unsigned long long* n_sites) {
  // one block walks the block sums in tiles (n_blocks <= n_parents / 1024)
  __shared__ uint32_t smem_warp[33];
  __shared__ unsigned long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n_blocks; base += kScanThreads) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < n_blocks ? static_cast<uint32_t>(block_sums[i]) : 0u;
    uint32_t total;
    const uint32_t exclusive = block_exclusive_scan(v, smem_warp, total);
    if (i < n_blocks) block_sums[i] = carry + exclusive;
    __syncthreads();
    if (threadIdx.x == 0) carry += total;
    __syncthreads();
  }
  // the bank's size: the sites MADE (a claimed run can end in unused slots: a secondary whose group walk over chi
  // found no group is not made, Multigroup.cpp:278-283), not the slots claimed
  if (threadIdx.x == 0 && n_sites) *n_sites = carry;
}

__global__ void __launch_bounds__(kScanThreads) order_bank_kernel(
    const uint32_t* __restrict__ child_count, const unsigned long long* __restrict__ child_start, uint64_t n_parents,
    const unsigned long long* __restrict__ block_offsets, const BankSite* __restrict__ unordered, BankSite* __restrict__ ordered) {
  __shared__ uint32_t smem_warp[33];
  const uint64_t first = static_cast<uint64_t>(blockIdx.x) * kScanTile + threadIdx.x * kScanItems;
  uint32_t counts[kScanItems];
  uint32_t v = 0;
  for (int k = 0; k < kScanItems; k++) {
    counts[k] = first + k < n_parents ? child_count[first + k] : 0u;
    v += counts[k];
  }
  uint32_t total;
  unsigned long long dst = block_offsets[blockIdx.x] + block_exclusive_scan(v, smem_warp, total);
  for (int k = 0; k < kScanItems; k++) {
    if (counts[k]) {
      const unsigned long long src = child_start[first + k];
      for (uint32_t j = 0; j < counts[k]; j++) ordered[dst + j] = unordered[src + j];
      dst += counts[k];
    }
  }
}

// Source bank of the next generation: N sites drawn from the M ordered fission
// sites with a deterministic comb, source i <- site floor(i * M / N); copies of
// one site get consecutive seeds (site seed + copy ordinal).
__global__ void resample_bank_kernel(
    const BankSite* __restrict__ slice, uint64_t slice_first, uint64_t slice_n, uint64_t m_total, uint64_t n_total,
    uint64_t first_out, uint64_t n_out, BankSite* __restrict__ next, unsigned long long* errors) {
  const uint64_t t = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n_out) return;
  const uint64_t i = first_out + t;
  // floor(i * M / N) and ceil(j * N / M) with 128-bit products
  const unsigned __int128 im = static_cast<unsigned __int128>(i) * m_total;
  const uint64_t j = static_cast<uint64_t>(im / n_total);
  const unsigned __int128 jn = static_cast<unsigned __int128>(j) * n_total;
  const uint64_t i0 = static_cast<uint64_t>((jn + m_total - 1) / m_total);
  if (j < slice_first || j >= slice_first + slice_n) {
    atomicAdd(errors, 1ull);
    return;
  }
  BankSite s = slice[j - slice_first];
  store_seed(s, load_seed(s) + (i - i0));  // (minstd mode: the sum wraps to 32 bits, as the seed word does)
  next[t] = s;
}

// One thread per element of a dense table: the rank-R sum in the reference's order (0.0 + p0, + p1, ...), the same
// IEEE operations as ce::pod_evaluate / ce::evaluate_inelastic make on the fly, so that a table entry IS the value
// the on-the-fly path computes.
__global__ void expand_dense_kernel(char* world, const DenseJob job) {
  const uint64_t n = static_cast<uint64_t>(job.n_grid) * job.n_cdf * job.n_T;
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t t = static_cast<uint32_t>(i % job.n_T);
  const uint32_t c = static_cast<uint32_t>((i / job.n_T) % job.n_cdf);
  const uint32_t g = static_cast<uint32_t>(i / (static_cast<uint64_t>(job.n_T) * job.n_cdf));
  const double* a = reinterpret_cast<const double*>(world + job.off_a) + static_cast<size_t>(c) * job.rank;
  const double* m = reinterpret_cast<const double*>(world + job.off_m) + (static_cast<size_t>(g) * job.n_T + t) * job.rank;
  double sum = 0;
  for (uint32_t r = 0; r < job.rank; r++) sum = __dadd_rn(sum, __dmul_rn(a[r], m[r]));
  const uint64_t at = job.cdf_fastest ? (static_cast<uint64_t>(g) * job.n_T + t) * job.n_cdf + c : i;
  reinterpret_cast<double*>(world + job.off_out)[at] = sum;
}

cudaError_t launch_expand_dense(char* world_d, const DenseJob* jobs, size_t n_jobs, cudaStream_t stream) {
  for (size_t k = 0; k < n_jobs; k++) {
    const uint64_t n = static_cast<uint64_t>(jobs[k].n_grid) * jobs[k].n_cdf * jobs[k].n_T;
    if (n == 0) continue;
    expand_dense_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(world_d, jobs[k]);
  }
  return cudaGetLastError();
}

// One thread per (temperature slot, grid point, CDF node): BetaPartition::Evaluate / AlphaPartition::Evaluate
// (ThermalScattering.cpp:183-215,225-256) at the slot's temperature -- both rank-R sums in the reference's order from
// 0.0 and the interpolation v_lo + (v_hi - v_lo) / (T_hi - T_lo) * (T - T_lo), the IEEE operations ce::pod_evaluate
// makes on the fly (__ddiv_rn is the correctly rounded quotient ce::divide_by reproduces).
__global__ void evaluate_rows_kernel(char* world, const __grid_constant__ EvalJob job) {
  const uint32_t s = blockIdx.y;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= job.n_grid * job.n_cdf) return;
  const uint32_t c = i % job.n_cdf, g = i / job.n_cdf;
  const double* a = reinterpret_cast<const double*>(world + job.off_a) + static_cast<size_t>(c) * job.rank;
  const double* m = reinterpret_cast<const double*>(world + job.off_m);
  const double* hi = m + (static_cast<size_t>(g) * job.n_T + job.t_hi[s]) * job.rank;
  const double* lo = m + (static_cast<size_t>(g) * job.n_T + job.t_lo[s]) * job.rank;
  double v_hi = 0, v_lo = 0;
  for (uint32_t r = 0; r < job.rank; r++) {
    v_hi = __dadd_rn(v_hi, __dmul_rn(a[r], hi[r]));
    v_lo = __dadd_rn(v_lo, __dmul_rn(a[r], lo[r]));
  }
  reinterpret_cast<double*>(world + job.off_out)[(static_cast<size_t>(s) * job.n_grid + g) * job.n_cdf + c] =
      __dadd_rn(v_lo, __dmul_rn(__ddiv_rn(__dsub_rn(v_hi, v_lo), job.dT[s]), job.tT[s]));
}

// One thread per evaluated row (temperature slot, grid point): clears the partition's eval_sorted flag (set in the
// uploaded image) when the row is not non-decreasing -- NaN included -- in the CDF node.
__global__ void check_rows_sorted_kernel(char* world, const __grid_constant__ EvalJob job) {
  const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= job.n_slots * job.n_grid) return;
  const double* v = reinterpret_cast<const double*>(world + job.off_out) + static_cast<size_t>(row) * job.n_cdf;
  bool sorted = true;
  for (uint32_t c = 0; c + 1 < job.n_cdf; c++) sorted = sorted && v[c] <= v[c + 1];
  if (!sorted || !(v[0] == v[0])) {
    *reinterpret_cast<uint32_t*>(world + job.off_sorted_flag) = 0u;
    *reinterpret_cast<uint32_t*>(world + job.off_direct_flag) = 0u;
  }
}

cudaError_t launch_evaluate_rows(char* world_d, const EvalJob* jobs, size_t n_jobs, cudaStream_t stream) {
  for (size_t k = 0; k < n_jobs; k++) {
    const uint32_t n = jobs[k].n_grid * jobs[k].n_cdf;
    if (n == 0 || jobs[k].n_slots == 0) continue;
    evaluate_rows_kernel<<<dim3((n + 255) / 256, jobs[k].n_slots), 256, 0, stream>>>(world_d, jobs[k]);
    const uint32_t rows = jobs[k].n_slots * jobs[k].n_grid;
    check_rows_sorted_kernel<<<(rows + 127) / 128, 128, 0, stream>>>(world_d, jobs[k]);
  }
  return cudaGetLastError();
}

cudaError_t launch_source_bank(const RunSpec& run, BankSite* bank, cudaStream_t stream) {
  if (run.n_histories == 0) return cudaSuccess;
  source_bank_kernel<<<static_cast<unsigned>((run.n_histories + 255) / 256), 256, 0, stream>>>(run, bank);
  return cudaGetLastError();
}

uint32_t bank_scan_blocks(uint64_t n_parents) { return static_cast<uint32_t>((n_parents + kScanTile - 1) / kScanTile); }

cudaError_t launch_order_bank(
    const uint32_t* child_count, const unsigned long long* child_start, uint64_t n_parents, unsigned long long* block_sums,
    const BankSite* unordered, BankSite* ordered, unsigned long long* n_sites, cudaStream_t stream) {
  if (n_parents == 0) return cudaSuccess;
  const uint32_t blocks = bank_scan_blocks(n_parents);
  bank_block_sums_kernel<<<blocks, kScanThreads, 0, stream>>>(child_count, n_parents, block_sums);
  bank_scan_block_sums_kernel<<<1, kScanThreads, 0, stream>>>(block_sums, blocks, n_sites);
  order_bank_kernel<<<blocks, kScanThreads, 0, stream>>>(child_count, child_start, n_parents, block_sums, unordered, ordered);
  return cudaGetLastError();
}

cudaError_t launch_resample_bank(
    const BankSite* slice, uint64_t slice_first, uint64_t slice_n, uint64_t m_total, uint64_t n_total, uint64_t first_out,
    uint64_t n_out, BankSite* next, unsigned long long* errors, cudaStream_t stream) {
  if (n_out == 0) return cudaSuccess;
  resample_bank_kernel<<<static_cast<unsigned>((n_out + 255) / 256), 256, 0, stream>>>(
      slice, slice_first, slice_n, m_total, n_total, first_out, n_out, next, errors);
  return cudaGetLastError();
}

__global__ void test_geometry_kernel(
    const char* __restrict__ world_g, size_t n, const double* pos, const double* dir, int32_t* cell, int32_t* surface,
    double* distance) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const WorldView w(world_g);
  Particle p;
  p.px = pos[3 * i], p.py = pos[3 * i + 1], p.pz = pos[3 * i + 2];
  p.dx = dir[3 * i], p.dy = dir[3 * i + 1], p.dz = dir[3 * i + 2];
  p.cell = find_cell(w, p.px, p.py, p.pz);
  cell[i] = p.cell;
  surface[i] = -1;
  distance[i] = __longlong_as_double(0x7ff0000000000000ll);
  if (p.cell >= 0) {
    int32_t nearest;
    distance[i] = nearest_surface(w, p, nearest);
    surface[i] = nearest;
  }
}

cudaError_t launch_test_geometry(
    const char* world_d, size_t n, const double* pos_d, const double* dir_d, int32_t* cell_d, int32_t* surface_d,
    double* distance_d, cudaStream_t stream) {
  test_geometry_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, stream>>>(
      world_d, n, pos_d, dir_d, cell_d, surface_d, distance_d);
  return cudaGetLastError();
}

cudaError_t launch_test_math(int fn, const double* x_d, double* out0_d, double* out1_d, size_t n, cudaStream_t stream) {
  test_math_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(fn, x_d, out0_d, out1_d, n);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ launchers
namespace {
// picks the instantiation for (tracking, energy mode, fixed source | generation | fixed source with sensitivities)
template <typename F> auto dispatch_history_kernel(int tracking, bool ce, bool generation, bool perturb, F&& f) {
  const bool delta = tracking == MMC_TRACK_CELL_DELTA;
  if (generation) {
    if (ce) return delta ? f(fixed_source_kernel<MMC_TRACK_CELL_DELTA, true, true, false>)
                         : f(fixed_source_kernel<MMC_TRACK_SURFACE, true, true, false>);
    return delta ? f(fixed_source_kernel<MMC_TRACK_CELL_DELTA, false, true, false>)
                 : f(fixed_source_kernel<MMC_TRACK_SURFACE, false, true, false>);
  }
  if (perturb) {
    if (ce) return delta ? f(fixed_source_kernel<MMC_TRACK_CELL_DELTA, true, false, true>)
                         : f(fixed_source_kernel<MMC_TRACK_SURFACE, true, false, true>);
    return delta ? f(fixed_source_kernel<MMC_TRACK_CELL_DELTA, false, false, true>)
                 : f(fixed_source_kernel<MMC_TRACK_SURFACE, false, false, true>);
  }
  if (ce) return delta ? f(fixed_source_kernel<MMC_TRACK_CELL_DELTA, true, false, false>)
                       : f(fixed_source_kernel<MMC_TRACK_SURFACE, true, false, false>);
  return delta ? f(fixed_source_kernel<MMC_TRACK_CELL_DELTA, false, false, false>)
               : f(fixed_source_kernel<MMC_TRACK_SURFACE, false, false, false>);
}
}  // namespace

cudaError_t launch_fixed_source(
    const LaunchConfig& cfg, const char* world_d, const RunSpec& run, const double* bounds_d, BankSite* site_scratch,
    uint2* pending_scratch, unsigned long long* next_history, unsigned long long* scores,
    unsigned long long* square_scores, mmc_counters* counters, const GenerationIO* generation, cudaStream_t stream,
    const ResumeIO* resume, const SensitivityIO* sensitivity) {
  const size_t smem = run.world_in_smem ? run.world_bytes : 0;
  const GenerationIO io = generation ? *generation : GenerationIO{};
  const ResumeIO rs = resume ? *resume : ResumeIO{};
  const SensitivityIO se = sensitivity ? *sensitivity : SensitivityIO{};
  if (resume && run.continuous_energy && !generation && !sensitivity) {  // the drain of an event-split run
    auto drain = run.tracking == MMC_TRACK_CELL_DELTA ? fixed_source_kernel<MMC_TRACK_CELL_DELTA, true, false, false, true>
                                                      : fixed_source_kernel<MMC_TRACK_SURFACE, true, false, false, true>;
    drain<<<cfg.blocks, kThreadsPerBlock, smem, stream>>>(
        world_d, run, bounds_d, site_scratch, pending_scratch, next_history, scores, square_scores, counters, io, rs, se);
    return cudaGetLastError();
  }
  return dispatch_history_kernel(run.tracking, run.continuous_energy != 0, generation != nullptr, sensitivity != nullptr, [&](auto kernel) -> cudaError_t {
    if (smem > 48 * 1024) {
      const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return e;
    }
    kernel<<<cfg.blocks, kThreadsPerBlock, smem, stream>>>(
        world_d, run, bounds_d, site_scratch, pending_scratch, next_history, scores, square_scores, counters, io, rs, se);
    return cudaGetLastError();
  });
}

cudaError_t launch_trace(
    const char* world_d, const RunSpec& run, BankSite* site_scratch, mmc_event_record* records, unsigned long long cap,
    unsigned long long* n_records, mmc_counters* counters, cudaStream_t stream) {
  auto go = [&](auto kernel) {
    kernel<<<1, 1, 0, stream>>>(world_d, run, site_scratch, records, cap, n_records, counters);
    return cudaGetLastError();
  };
  if (run.continuous_energy) {
    if (run.tracking == MMC_TRACK_CELL_DELTA) return go(trace_kernel<MMC_TRACK_CELL_DELTA, true>);
    return go(trace_kernel<MMC_TRACK_SURFACE, true>);
  }
  if (run.tracking == MMC_TRACK_CELL_DELTA) return go(trace_kernel<MMC_TRACK_CELL_DELTA, false>);
  return go(trace_kernel<MMC_TRACK_SURFACE, false>);
}

int max_blocks_per_sm(int tracking, bool continuous_energy, bool generation, size_t smem, bool perturb) {
  return dispatch_history_kernel(tracking, continuous_energy, generation, perturb, [&](auto kernel) {
    int n = 0;
    // the query answers 0 for dynamic shared memory above the function's current limit: opt in first
    if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kThreadsPerBlock, smem);
    return n;
  });
}

}  // namespace MMC_VARIANT_NS
}  // namespace mmc
