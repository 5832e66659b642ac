// Pieces shared by the fused history kernel (kernels.cu) and the event-split
// kernels (event_loop.cu): bank-site loading, per-thread event counters.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "kernels.h"
#include "transport.cuh"

namespace mmc {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ void load_site(const BankSite& s, Particle& p) {
  p.px = s.position[0];
  p.py = s.position[1];
  p.pz = s.position[2];
  p.dx = s.direction[0];
  p.dy = s.direction[1];
  p.dz = s.direction[2];
  p.group = s.energy_bits;
  p.energy = __longlong_as_double(static_cast<long long>(s.energy_bits));
  p.rng = Rng::seeded(load_seed(s));
  p.cell = -1;
  p.surface = -1;
  p.event = MMC_EV_BIRTH;
}

// the fields of mmc_counters, in declaration order
constexpr int kNumCounters = 12;

struct ThreadCounters {
  uint32_t histories = 0, births = 0, events = 0, collisions = 0, crossings = 0, virtuals = 0, scores = 0,
           secondaries = 0, banked = 0, lost = 0, capacity = 0, physics = 0;
};

__device__ __forceinline__ void count_event(ThreadCounters& c, const Particle& p, const StepOut& o) {
  c.events++;
  c.collisions += (p.event == MMC_EV_SCATTER || p.event == MMC_EV_CAPTURE || p.event == MMC_EV_FISSION) &&
                  !o.error_physics;
  c.crossings += (p.event == MMC_EV_SURFACE_CROSS || p.event == MMC_EV_LEAK) && !o.error_physics;
  c.virtuals += p.event == MMC_EV_VIRTUAL_COLLISION;
  c.secondaries += o.secondaries;
  c.lost += o.error_lost;
  c.capacity += o.error_capacity;
  c.physics += o.error_physics;
}

}  // namespace mmc
