// Device-side restatement of the reference's per-event physics for multigroup
// worlds, written for sm_100a.  Every function cites the reference lines whose
// arithmetic it must reproduce operation for operation: this translation unit
// is compiled with -fmad=false so that each + - * / sqrt is one IEEE-754
// round-to-nearest operation, exactly as g++ emits for baseline x86-64
// (SURVEY.md F5, H1).  log / sincos / sin / cos are glibc's, restated bit for
// bit in glibc_math.h, because the reference's absolute 10-eps surface nudge
// makes cell assignment depend on the last bit of a position.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/minimc_b200.h"
#include "glibc_math.h"
#include "world_blob.h"

namespace mmc {

// ---------------------------------------------------------------- world view
// Typed view over the staged blob (shared or global memory).
struct WorldView {
  const char* base;
  const WorldHeader* h;
  __device__ __forceinline__ explicit WorldView(const char* b)
      : base(b), h(reinterpret_cast<const WorldHeader*>(b)) {}
  // header in kernel-parameter (constant) space: every w.h->off_* becomes a constant-bank operand instead of a
  // dependent global load in front of each table access
  __device__ __forceinline__ WorldView(const char* b, const WorldHeader* header) : base(b), h(header) {}
  template <typename T> __device__ __forceinline__ const T* at(uint32_t off) const {
    return reinterpret_cast<const T*>(base + off);
  }
};

// ----------------------------------------------------------------------- RNG
// std::minstd_rand = linear_congruential_engine<uint_fast32_t, 48271, 0, 2^31-1>
// (BasicTypes.hpp:27).  Seeding: x = seed mod m, and 0 -> 1 (libstdc++
// random.tcc, linear_congruential_engine::seed).
constexpr uint32_t kLcgM = 2147483647u;

__host__ __device__ __forceinline__ uint32_t lcg_seed(uint64_t seed) {
  const uint32_t x = static_cast<uint32_t>(seed % kLcgM);
  return x == 0 ? 1u : x;
}

__host__ __device__ __forceinline__ uint32_t lcg_next(uint32_t x) {
  // 48271 * x < 2^47; reduce modulo the Mersenne prime 2^31-1 by folding
  const uint64_t p = static_cast<uint64_t>(x) * 48271u;
  uint32_t r = static_cast<uint32_t>(p & kLcgM) + static_cast<uint32_t>(p >> 31);
  return r >= kLcgM ? r - kLcgM : r;
}

#ifndef MMC_COUNTER_RNG
#define MMC_COUNTER_RNG 0
#endif

#if MMC_COUNTER_RNG
// MMC_COUNTER_RNG = 1 (mmc_rng_mode MMC_RNG_COUNTER of include/minimc_b200.h; this translation unit is compiled once per mode): a
// counter-based generator per particle instead of the reference's sequential std::minstd_rand -- Philox-2x32-10
// (Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3", SC'11; restated from the paper).  A
// stream is a 64-bit id (the Philox key and the high word of its counter); draw number n of the stream is the block
// philox(key = id.lo, counter = (n, id.hi)) -- 64 fresh bits, no state to advance, no division in canonical().  Where
// the reference constructs a generator from a seed (Source::Sample's std::minstd_rand{seed + i}, the Particle ctor's
// rng{seed}, Particle::BankSecondaries' seed = parent's next draw, Source.cpp:143-154, Particle.cpp:96-100) this mode
// starts the stream whose id is that seed, taken as the 64 bits of one block.  Results are statistically equivalent
// to the minstd mode (not bit-identical: other random numbers) and, like it, independent of batch splits, schedules
// and GPU counts: a particle's stream is a pure function of its history's seed and its ancestry.
struct Rng {
  uint32_t x;       // draws made so far (the counter's low word)
  uint32_t k0, k1;  // the stream's 64-bit id
  __device__ __forceinline__ uint2 block() {
    uint32_t c0 = x, c1 = k1, key = k0;
#pragma unroll
    for (int round = 0; round < 10; round++) {
      const uint32_t hi = __umulhi(0xD256D193u, c0), lo = 0xD256D193u * c0;
      c0 = hi ^ key ^ c1;
      c1 = lo;
      key += 0x9E3779B9u;
    }
    x++;
    return make_uint2(c0, c1);
  }
  __device__ __forceinline__ uint32_t raw() { return block().x; }
  // 53 random bits scaled to [0, 1): every value is a multiple of 2^-53 below 1
  __device__ __forceinline__ double canonical() {
    const uint2 b = block();
    const uint64_t bits = (static_cast<uint64_t>(b.x) << 32 | b.y) >> 11;
    return __dmul_rn(static_cast<double>(bits), 1.1102230246251565404e-16);
  }
  // the seed handed to a new generator (Particle ctor, BankSecondaries): 64 bits
  __device__ __forceinline__ uint64_t spawn() {
    const uint2 b = block();
    return static_cast<uint64_t>(b.x) << 32 | b.y;
  }
  __device__ __forceinline__ static Rng seeded(uint64_t seed) {
    Rng r;
    r.x = 0;
    r.k0 = static_cast<uint32_t>(seed);
    r.k1 = static_cast<uint32_t>(seed >> 32);
    return r;
  }
  __device__ __forceinline__ uint64_t stream() const { return static_cast<uint64_t>(k1) << 32 | k0; }
  // for event records: the stream's low word and the draws made
  __device__ __forceinline__ uint64_t state64() const { return static_cast<uint64_t>(k0) << 32 | x; }
};
#else
struct Rng {
  uint32_t x;
  __device__ __forceinline__ uint32_t raw() {
    x = lcg_next(x);
    return x;
  }
  // std::generate_canonical<double, 53>(minstd_rand): two draws,
  // sum = (x1-1) + (x2-1)*R, R = 2147483646; result sum / R^2 with R^2 rounded
  // from long double to double (bits/random.tcc:3349).
  __device__ __forceinline__ double canonical() {
    const uint32_t x1 = raw();
    const uint32_t x2 = raw();
    const double sum =
        __dadd_rn(static_cast<double>(x1 - 1u), __dmul_rn(static_cast<double>(x2 - 1u), 2147483646.0));
    double u = __ddiv_rn(sum, 4611686009837453312.0);
    if (u >= 1.0) u = 0.99999999999999988897769753748;  // nextafter(1, 0)
    return u;
  }
  // the seed handed to a new generator (Particle ctor, BankSecondaries): the engine's next 32-bit output
  __device__ __forceinline__ uint64_t spawn() { return raw(); }
  // std::minstd_rand{seed}
  __device__ __forceinline__ static Rng seeded(uint64_t seed) { return Rng{lcg_seed(seed)}; }
  __device__ __forceinline__ uint64_t stream() const { return x; }
  __device__ __forceinline__ uint64_t state64() const { return x; }
};
#endif

// A generator's seed in a bank site: the 32-bit seed of std::minstd_rand{seed} (Particle.cpp:96-100), or, in counter
// mode, the stream's 64-bit id in the seed and surface words (Particle::current_surface of a banked particle is null).
__device__ __forceinline__ void store_seed(BankSite& s, uint64_t seed) {
  s.seed = static_cast<uint32_t>(seed);
  s.surface = MMC_COUNTER_RNG ? static_cast<int32_t>(static_cast<uint32_t>(seed >> 32)) : -1;
}
__device__ __forceinline__ uint64_t load_seed(const BankSite& s) {
  return MMC_COUNTER_RNG ? (static_cast<uint64_t>(static_cast<uint32_t>(s.surface)) << 32 | s.seed) : s.seed;
}

// ------------------------------------------------------------------ particle
struct Particle {
  double px, py, pz;
  double dx, dy, dz;
  uint64_t group;   // multigroup: 1..G
  double energy;    // continuous energy: MeV
  Rng rng;
  int32_t cell;
  int32_t surface;  // Particle::current_surface (persists across collisions)
  int32_t event;    // mmc_event
};

__device__ __forceinline__ bool is_alive(int32_t event) {
  // Particle.cpp:126-129
  return event != MMC_EV_CAPTURE && event != MMC_EV_LEAK && event != MMC_EV_FISSION;
}

// Direction(RNG&): Point.cpp:89-96
__device__ __forceinline__ void isotropic_direction(Rng& rng, double& x, double& y, double& z) {
  x = __dadd_rn(__dmul_rn(rng.canonical(), 2.0), -1.0);
  const double sin_theta = __dsqrt_rn(__dsub_rn(1.0, __dmul_rn(x, x)));
  const double phi = __dmul_rn(rng.canonical(), 6.283185307179586476925286766559);
  double s, c;
  glibc::sincos(phi, &s, &c);  // g++ -O3 merges the reference's cos/sin pair into one sincos call
  y = __dmul_rn(sin_theta, c);
  z = __dmul_rn(sin_theta, s);
}

__device__ __forceinline__ void normalize(double& x, double& y, double& z) {
  // Point::Normalize: *this /= sqrt(Dot(*this)), Point.cpp:44-47
  const double n = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
  // (+-0) / n == (+-0) for every n > 0.  The cross product with a coordinate axis (rotate_direction) always has one
  // component that is exactly zero, and a zero QUOTIENT sends CUDA's fp64 division down its slow path (a ~300
  // instruction subroutine: 5 % of the S(a,b) kernel's instructions).  Such a lane divides 0 + n by n instead (fast path)
  // and keeps its zero: same result, no branch.
  const bool zx = x == 0.0 && n > 0.0, zy = y == 0.0 && n > 0.0, zz = z == 0.0 && n > 0.0;
  // (the dividend is formed by an addition, not a select: the compiler folds `zx ? n : x` back to x because the
  // quotient of such a lane is not used)
  const double qx = __ddiv_rn(__dadd_rn(x, zx ? n : 0.0), n), qy = __ddiv_rn(__dadd_rn(y, zy ? n : 0.0), n),
               qz = __ddiv_rn(__dadd_rn(z, zz ? n : 0.0), n);
  x = zx ? x : qx;
  y = zy ? y : qy;
  z = zz ? z : qz;
}

// Direction(const Direction& d, mu, phi): Point.cpp:98-121
static __device__ __noinline__ void rotate_direction(
    double dx, double dy, double dz, double mu, double phi, double& ox, double& oy, double& oz) {
  const bool off_xaxis = dx <= 0.9 && dx > -0.9;
  const double ax = off_xaxis ? 1.0 : 0.0, ay = off_xaxis ? 0.0 : 1.0, az = 0.0;
  // u = Direction{d.Cross(axis)}
  double ux = __dsub_rn(__dmul_rn(dy, az), __dmul_rn(dz, ay));
  double uy = __dsub_rn(__dmul_rn(dz, ax), __dmul_rn(dx, az));
  double uz = __dsub_rn(__dmul_rn(dx, ay), __dmul_rn(dy, ax));
  normalize(ux, uy, uz);
  // v = Direction{d.Cross(u)}
  double vx = __dsub_rn(__dmul_rn(dy, uz), __dmul_rn(dz, uy));
  double vy = __dsub_rn(__dmul_rn(dz, ux), __dmul_rn(dx, uz));
  double vz = __dsub_rn(__dmul_rn(dx, uy), __dmul_rn(dy, ux));
  normalize(vx, vy, vz);
  // separate cos and sin calls in the reference's object code (Point.cpp:117-118)
  double s, c;
  glibc::sin_and_cos(phi, &s, &c);
  const double sq = __dsqrt_rn(__dsub_rn(1.0, __dmul_rn(mu, mu)));
  const double uc = __dmul_rn(sq, c), vc = __dmul_rn(sq, s);
  // (u_comp + v_comp) + d_comp, then Direction(Point&&) normalises
  ox = __dadd_rn(__dadd_rn(__dmul_rn(ux, uc), __dmul_rn(vx, vc)), __dmul_rn(dx, mu));
  oy = __dadd_rn(__dadd_rn(__dmul_rn(uy, uc), __dmul_rn(vy, vc)), __dmul_rn(dy, mu));
  oz = __dadd_rn(__dadd_rn(__dmul_rn(uz, uc), __dmul_rn(vz, vc)), __dmul_rn(dz, mu));
  normalize(ox, oy, oz);
}

// ------------------------------------------------------------------ geometry
// CSGSurface::SolveQuadratic, CSGSurface.cpp:71-97
__device__ __forceinline__ double solve_quadratic(double a, double b, double c) {
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  const double disc = __dsub_rn(__dmul_rn(b, b), __dmul_rn(__dmul_rn(4.0, a), c));
  if (disc <= 0) return inf;
  const double sq = __dsqrt_rn(disc);
  const double two_a = __dmul_rn(2.0, a), two_c = __dmul_rn(2.0, c);
  double lesser, greater;
  if (b > 0) {
    const double t = __dsub_rn(-b, sq);
    lesser = __ddiv_rn(t, two_a);
    greater = __ddiv_rn(two_c, t);
  } else {
    const double t = __dadd_rn(-b, sq);
    lesser = __ddiv_rn(two_c, t);
    greater = __ddiv_rn(t, two_a);
  }
  if (lesser > 0) return lesser;
  if (greater > 0) return greater;
  return inf;
}

__device__ __forceinline__ double surface_distance(
    int32_t type, const double* prm, double px, double py, double pz, double dx, double dy, double dz) {
  if (type == MMC_SURF_SPHERE) {
    // Sphere::Distance, CSGSurface.cpp:107-111
    const double ox = __dsub_rn(px, prm[0]), oy = __dsub_rn(py, prm[1]), oz = __dsub_rn(pz, prm[2]);
    const double od = __dadd_rn(__dadd_rn(__dmul_rn(ox, dx), __dmul_rn(oy, dy)), __dmul_rn(oz, dz));
    const double oo = __dadd_rn(__dadd_rn(__dmul_rn(ox, ox), __dmul_rn(oy, oy)), __dmul_rn(oz, oz));
    return solve_quadratic(1.0, __dmul_rn(2.0, od), __dsub_rn(oo, __dmul_rn(prm[3], prm[3])));
  } else if (type == MMC_SURF_PLANEX) {
    // PlaneX::Distance, CSGSurface.cpp:125-134.  Dot(Point{1,0,0}) = x*1 + y*0 + z*0
    // equals x for finite y, z (only the sign of a zero can differ, which no
    // comparison below observes).
    const double d = __ddiv_rn(__dsub_rn(prm[0], px), dx);
    return d > 0 ? d : __longlong_as_double(0x7ff0000000000000ll);
  } else {
    // CylinderX::Distance, CSGSurface.cpp:148-164
    const double pp = __dadd_rn(__dadd_rn(__dmul_rn(px, px), __dmul_rn(py, py)), __dmul_rn(pz, pz));
    const double pw = px, dw = dx;  // Dot with the axis Direction{1,0,0}
    const double pd = __dadd_rn(__dadd_rn(__dmul_rn(px, dx), __dmul_rn(py, dy)), __dmul_rn(pz, dz));
    return solve_quadratic(
        __dsub_rn(1.0, __dmul_rn(dw, dw)), __dmul_rn(2.0, __dsub_rn(pd, __dmul_rn(pw, dw))),
        __dsub_rn(__dsub_rn(pp, __dmul_rn(pw, pw)), __dmul_rn(prm[0], prm[0])));
  }
}

__device__ __forceinline__ bool surface_contains(int32_t type, const double* prm, double px, double py, double pz) {
  if (type == MMC_SURF_SPHERE) {
    // Sphere::Contains, CSGSurface.cpp:113-116
    const double ox = __dsub_rn(px, prm[0]), oy = __dsub_rn(py, prm[1]), oz = __dsub_rn(pz, prm[2]);
    const double oo = __dadd_rn(__dadd_rn(__dmul_rn(ox, ox), __dmul_rn(oy, oy)), __dmul_rn(oz, oz));
    return oo < __dmul_rn(prm[3], prm[3]);
  } else if (type == MMC_SURF_PLANEX) {
    // PlaneX::Contains, CSGSurface.cpp:136-138
    return px < prm[0];
  } else {
    // CylinderX::Contains, CSGSurface.cpp:166-174 -- quirk Q2: sqrt(r_perp^2) < radius^2
    const double pp = __dadd_rn(__dadd_rn(__dmul_rn(px, px), __dmul_rn(py, py)), __dmul_rn(pz, pz));
    const double pw = px;
    return __dsqrt_rn(__dsub_rn(pp, __dmul_rn(pw, pw))) < __dmul_rn(prm[0], prm[0]);
  }
}

// World::FindCellContaining, World.cpp:26-37: first cell in XML order whose
// every surface satisfies Contains(p) == sense (Cell::Contains, Cell.cpp:27-35).
// Returns -1 where the reference throws.
__device__ inline int32_t find_cell(const WorldView& w, double px, double py, double pz) {
  if (w.h->off_cell_mask) {
    // Contains() of every surface once (independent loads, no early exit: the lanes of a warp stay together), then
    // each Cell is one mask compare; the first match in XML order, as the reference's scan
    const int32_t* type = w.at<int32_t>(w.h->off_surface_type);
    const double* prm = w.at<double>(w.h->off_surface_param);
    unsigned long long inside = 0;
    for (int32_t s = 0; s < w.h->n_surfaces; s++)
      inside |= static_cast<unsigned long long>(surface_contains(type[s], prm + 4 * s, px, py, pz)) << s;
    const ulonglong2* masks = w.at<ulonglong2>(w.h->off_cell_mask);
    for (int32_t c = 0; c < w.h->n_cells; c++) {
      const ulonglong2 m = masks[c];  // {mask, want}
      if ((inside & m.x) == m.y) return c;
    }
    return -1;
  }
  const int32_t* begin = w.at<int32_t>(w.h->off_cell_surf_begin);
  const SurfaceRecord* rec = w.at<SurfaceRecord>(w.h->off_cell_surf_rec);
  int32_t k = begin[0];
  for (int32_t c = 0; c < w.h->n_cells; c++) {
    const int32_t end = begin[c + 1];
    bool inside = true;
    for (; k < end && inside; k++) {
      const SurfaceRecord& r = rec[k];
      inside = surface_contains(r.type, r.prm, px, py, pz) == ((r.index_sense & 1) != 0);
    }
    if (inside) return c;
    k = end;
  }
  return -1;
}

// Cell::NearestSurface, Cell.cpp:37-51: std::min_element keeps the FIRST minimum.
__device__ __forceinline__ double nearest_surface(const WorldView& w, const Particle& p, int32_t& nearest) {
  const int32_t* begin = w.at<int32_t>(w.h->off_cell_surf_begin);
  const SurfaceRecord* rec = w.at<SurfaceRecord>(w.h->off_cell_surf_rec);
  const int32_t b = begin[p.cell], e = begin[p.cell + 1];
  double best = surface_distance(rec[b].type, rec[b].prm, p.px, p.py, p.pz, p.dx, p.dy, p.dz);
  nearest = rec[b].index_sense >> 1;
  for (int32_t k = b + 1; k < e; k++) {
    const double d = surface_distance(rec[k].type, rec[k].prm, p.px, p.py, p.pz, p.dx, p.dy, p.dz);
    if (d < best) {
      best = d;
      nearest = rec[k].index_sense >> 1;
    }
  }
  return best;
}

// Particle::Stream, Particle.cpp:46-53: position += direction * distance
__device__ __forceinline__ void stream(Particle& p, double d) {
  p.px = __dadd_rn(p.px, __dmul_rn(p.dx, d));
  p.py = __dadd_rn(p.py, __dmul_rn(p.dy, d));
  p.pz = __dadd_rn(p.pz, __dmul_rn(p.dz, d));
}

// ------------------------------------------------------------------- source
// Source::Sample, Source.cpp:143-154
// With defer_isotropic the isotropic direction is NOT drawn here: the particle's rng is left holding the SOURCE
// stream (std::minstd_rand{seed}) and the function returns true; the caller then runs
// isotropic_direction(p.rng, ...) followed by finish_source(p), which together are the rest of Source::Sample.
__device__ inline bool sample_source(const SourceSpec& src, uint64_t seed, Particle& p, bool defer_isotropic = false) {
  Rng rng = Rng::seeded(seed);
  p.px = src.position[0];
  p.py = src.position[1];
  p.pz = src.position[2];
  p.group = src.group;
  p.energy = src.energy;
  p.cell = -1;
  p.surface = -1;
  p.event = MMC_EV_BIRTH;
  if (src.direction_kind == MMC_DIR_ISOTROPIC) {
    if (defer_isotropic) {
      p.rng = rng;
      return true;
    }
    isotropic_direction(rng, p.dx, p.dy, p.dz);
  } else if (src.direction_kind == MMC_DIR_ISOTROPIC_FLUX) {
    // IsotropicFlux::Sample, Source.cpp:124-129
    const double mu = __dsqrt_rn(rng.canonical());
    const double phi = __dmul_rn(rng.canonical(), 6.283185307179586476925286766559);
    rotate_direction(src.direction[0], src.direction[1], src.direction[2], mu, phi, p.dx, p.dy, p.dz);
  } else {
    p.dx = src.direction[0];
    p.dy = src.direction[1];
    p.dz = src.direction[2];
  }
  p.rng = Rng::seeded(rng.spawn());  // Particle ctor: rng{seed}
  return false;
}

// auto sampled_seed = rng(); Particle{..., sampled_seed} (Source.cpp:149-153)
__device__ __forceinline__ void finish_source(Particle& p) { p.rng = Rng::seeded(p.rng.spawn()); }

// ------------------------------------------------------- secondary particles
// Per-thread ring deque in global scratch reproducing the bank order of
// FixedSource.cpp:63-72 / Bank.cpp:5-8: particles are taken from the back,
// a dead particle's secondaries are spliced, in creation order, to the front.
struct SiteDeque {
  BankSite* slots;   // capacity entries (power of two)
  uint32_t mask;
  uint32_t head;     // index of the front element
  uint32_t count;
  __device__ __forceinline__ bool empty() const { return count == 0; }
};

// ----------------------------------------------------------- multigroup step
struct StepOut {
  uint32_t secondaries;  // produced by this event
  bool need_direction;   // a multigroup scatter left its isotropic direction to the caller (kDeferDirection)
  bool error_physics;
  bool error_capacity;
  bool error_lost;
  // event-split schedule (event_loop.cu): the collision chose a thermal-scattering scatter and left the S(a,b)
  // sampling to the caller; the table's blob offset and the cell temperature at the collision site
  bool need_tsl;
  uint32_t tsl_off;
  double tsl_T;
  // event-split schedule: the particle streamed onto a surface and left Cell lookup, event code and tallies to the
  // boundary kernel (p.event == kEvCrossPending meanwhile)
  bool need_cross;
  // generation kernels (kDeferFission): a fission computed its yield and left the secondaries to the caller, which
  // produces them one per loop iteration from the dead parent's rng (converged with the other lanes' isotropic
  // directions) -- fission_nuclide names the chi rows (multigroup)
  uint32_t pending_yield;
  int32_t fission_nuclide;
};

// event codes used only between the kernels of one event-split pass (never in records or counters)
constexpr int32_t kEvCrossPending = 64;  // streamed to a surface: World::FindCellContaining still to do
constexpr int32_t kEvRetired = 65;       // the slot found no history left: it leaves the live queue

// Material::GetMicroscopicTotal (Material.cpp:53-62): accumulate afrac*total
// from 0 in afracs order.  For multigroup GetMajorant == GetTotal
// (Multigroup.cpp:40-43).
__device__ __forceinline__ double material_micro_total(const WorldView& w, int32_t mat, uint64_t group) {
  const int32_t* nb = w.at<int32_t>(w.h->off_mat_nuc_begin);
  const int32_t* ni = w.at<int32_t>(w.h->off_mat_nuc_index);
  const double* af = w.at<double>(w.h->off_mat_nuc_afrac);
  const double* total = w.at<double>(w.h->off_mg_total);
  const int32_t G = w.h->n_groups;
  double acc = 0.0;
  for (int32_t k = nb[mat]; k < nb[mat + 1]; k++)
    acc = __dadd_rn(acc, __dmul_rn(af[k], total[ni[k] * G + (group - 1)]));
  return acc;
}

// Multigroup::Interact and its Capture/Scatter/Fission (Multigroup.cpp:49-73,
// 245-287) after Particle::SampleNuclide (Particle.cpp:110-124).
template <bool kDeferDirection, bool kDeferFission = false>
__device__ inline void collide_multigroup(
    const WorldView& w, Particle& p, int32_t mat, double micro_total, SiteDeque& dq, StepOut& out) {
  const int32_t G = w.h->n_groups;
  const int32_t gi = static_cast<int32_t>(p.group - 1);
  // --- SampleNuclide
  const int32_t* nb = w.at<int32_t>(w.h->off_mat_nuc_begin);
  const int32_t* ni = w.at<int32_t>(w.h->off_mat_nuc_index);
  const double* af = w.at<double>(w.h->off_mat_nuc_afrac);
  const double* total = w.at<double>(w.h->off_mg_total);
  const double nuc_threshold = __dmul_rn(micro_total, p.rng.canonical());
  int32_t nuc = -1;
  double acc = 0.0;
  for (int32_t k = nb[mat]; k < nb[mat + 1]; k++) {
    acc = __dadd_rn(acc, __dmul_rn(af[k], total[ni[k] * G + gi]));
    if (acc > nuc_threshold) {
      nuc = ni[k];
      break;
    }
  }
  if (nuc < 0) {  // assert(false), Particle.cpp:123
    out.error_physics = true;
    p.event = MMC_EV_CAPTURE;
    return;
  }
  // --- Interact: reactions in enum order capture, scatter, fission
  const uint32_t mask = w.at<uint32_t>(w.h->off_mg_mask)[nuc];
  const double threshold = __dmul_rn(p.rng.canonical(), total[nuc * G + gi]);
  double racc = 0.0;
  int reaction = -1;
  if (mask & MMC_REACTION_CAPTURE) {
    racc = __dadd_rn(racc, w.at<double>(w.h->off_mg_capture)[nuc * G + gi]);
    if (racc > threshold) reaction = 0;
  }
  if (reaction < 0 && (mask & MMC_REACTION_SCATTER)) {
    racc = __dadd_rn(racc, w.at<double>(w.h->off_mg_scatter)[nuc * G + gi]);
    if (racc > threshold) reaction = 1;
  }
  if (reaction < 0 && (mask & MMC_REACTION_FISSION)) {
    racc = __dadd_rn(racc, w.at<double>(w.h->off_mg_fission)[nuc * G + gi]);
    if (racc > threshold) reaction = 2;
  }
  if (reaction == 0) {
    p.event = MMC_EV_CAPTURE;
  } else if (reaction == 1) {
    // Multigroup::Scatter, Multigroup.cpp:249-265
    p.event = MMC_EV_SCATTER;
    const double t = p.rng.canonical();
    const double* probs = w.at<double>(w.h->off_mg_scatter_probs) + (static_cast<size_t>(nuc) * G + gi) * G;
    double a = 0.0;
    int32_t g = 0;
    for (; g < G; g++) {
      a = __dadd_rn(a, probs[g]);
      if (a > t) break;
    }
    if (g == G) {  // assert(false), Multigroup.cpp:264
      out.error_physics = true;
      p.event = MMC_EV_CAPTURE;
      return;
    }
    p.group = static_cast<uint64_t>(g + 1);
    // SetDirectionIsotropic (Multigroup.cpp:258-259): the next draws of p.rng; the fused kernel runs it together
    // with the source directions of newly born lanes so that the warp stays converged on the sincos
    if (kDeferDirection) out.need_direction = true;
    else isotropic_direction(p.rng, p.dx, p.dy, p.dz);
  } else if (reaction == 2) {
    // Multigroup::Fission, Multigroup.cpp:267-287
    p.event = MMC_EV_FISSION;
    const double nubar = w.at<double>(w.h->off_mg_nubar)[nuc * G + gi];
    const uint64_t yield = static_cast<uint64_t>(__dadd_rn(nubar, p.rng.canonical()));
    if (kDeferFission) {
      out.pending_yield = static_cast<uint32_t>(yield);
      out.fission_nuclide = nuc;
      return;
    }
    const double* chi = w.at<double>(w.h->off_mg_chi) + (static_cast<size_t>(nuc) * G + gi) * G;
    // Secondaries are spliced to the FRONT of the bank in creation order
    // (Bank.cpp:5-8).  Each one is pushed in front of the previous one, then
    // the new front run is reversed.
    uint32_t produced = 0;
    for (uint64_t i = 0; i < yield; i++) {
      const double t = p.rng.canonical();
      double a = 0.0;
      for (int32_t g = 0; g < G; g++) {
        a = __dadd_rn(a, chi[g]);
        if (a > t) {
          BankSite s;
          s.position[0] = p.px;
          s.position[1] = p.py;
          s.position[2] = p.pz;
          isotropic_direction(p.rng, s.direction[0], s.direction[1], s.direction[2]);
          s.energy_bits = static_cast<uint64_t>(g + 1);
          store_seed(s, p.rng.spawn());  // Particle::BankSecondaries, Particle.cpp:96-100
          if (dq.count > dq.mask) {
            out.error_capacity = true;
          } else {
            dq.head = (dq.head - 1u) & dq.mask;
            dq.slots[dq.head] = s;
            dq.count++;
            produced++;
          }
          break;
        }
      }
    }
    for (uint32_t lo = 0, hi = produced; lo + 1 < hi; lo++) {
      hi--;
      const BankSite tmp = dq.slots[(dq.head + lo) & dq.mask];
      dq.slots[(dq.head + lo) & dq.mask] = dq.slots[(dq.head + hi) & dq.mask];
      dq.slots[(dq.head + hi) & dq.mask] = tmp;
    }
    out.secondaries = produced;
  } else {  // assert(false), Multigroup.cpp:72
    out.error_physics = true;
    p.event = MMC_EV_CAPTURE;
  }
}

}  // namespace mmc

#include "physics_ce.cuh"

namespace mmc {

// TotalCrossSectionPerturbation::Stream (Perturbation.cpp:68-80), called by Particle::Stream (Particle.cpp:46-53)
// before the position moves: every perturbation whose nuclide is in the material the particle streams through
// adds 1 / GetCollisionProbabilityDensity(p) - distance to its indirect effect.  GetCollisionProbabilityDensity is the
// MICROSCOPIC total (surface tracking) or majorant (cell delta tracking) of the material, without the number density
// (TransportMethod.cpp:79-82,124-127): reproduced as it is.
struct PerturbContext {
  const int32_t* perturbed_nuclide;  // RunSpec::perturbed_nuclide
  int32_t n;
  double* indirect;                  // the particle's indirect effects, one per perturbation
};

__device__ __forceinline__ void perturb_stream(const WorldView& w, const PerturbContext& pc, int32_t mat, double micro, double distance) {
  const int32_t* nb = w.at<int32_t>(w.h->off_mat_nuc_begin);
  const int32_t* ni = w.at<int32_t>(w.h->off_mat_nuc_index);
  for (int32_t k = 0; k < pc.n; k++) {
    bool present = false;
    for (int32_t j = nb[mat]; j < nb[mat + 1]; j++) present = present || ni[j] == pc.perturbed_nuclide[k];
    if (present) pc.indirect[k] = __dadd_rn(pc.indirect[k], __dsub_rn(__ddiv_rn(1.0, micro), distance));
  }
}

// One iteration of SurfaceTracking::Transport (TransportMethod.cpp:56-75) or
// CellDeltaTracking::Transport (TransportMethod.cpp:92-120).  kCE selects the
// Continuous (true) or Multigroup (false) Interaction of the world's nuclides.
// the second half of a surface crossing (TransportMethod.cpp:68-73,100-105): the new Cell, leak or crossing
__device__ __forceinline__ void finish_crossing(const WorldView& w, Particle& p, StepOut& out) {
  p.cell = find_cell(w, p.px, p.py, p.pz);
  if (p.cell < 0) {
    out.error_lost = true;
    p.event = MMC_EV_LEAK;
    return;
  }
  p.event = w.at<int32_t>(w.h->off_cell_material)[p.cell] >= 0 ? MMC_EV_SURFACE_CROSS : MMC_EV_LEAK;
}

template <int kTracking, bool kCE, bool kDeferDirection = false, bool kDeferTsl = false, bool kPerturb = false,
          bool kDeferCross = false, bool kDeferFission = false>
__device__ __forceinline__ void transport_step(
    const WorldView& w, Particle& p, SiteDeque& dq, StepOut& out, const PerturbContext* pc = nullptr) {
  out.secondaries = 0;
  out.need_direction = false;
  out.need_tsl = false;
  out.need_cross = false;
  out.pending_yield = 0;
  out.fission_nuclide = -1;
  out.error_physics = out.error_capacity = out.error_lost = false;
  const int32_t mat = w.at<int32_t>(w.h->off_cell_material)[p.cell];
  if (mat < 0) {  // born in a void cell: the reference dereferences a null Material
    out.error_physics = true;
    p.event = MMC_EV_LEAK;
    return;
  }
  // GetCollisionProbabilityDensity: total (surface tracking) or majorant (cell delta tracking)
  double micro, majorant;
  ce::NuclideEval ev;
  if (kCE) {
    bool error = false;
    const double T = ce::cell_temperature(w, p.cell, p.px, p.py, p.pz);
    if (kTracking == MMC_TRACK_CELL_DELTA) {
      majorant = ce::material_majorant(w, mat, p.energy, T, ce::cell_temperature_upper(w, p.cell), error);
      micro = 0;  // evaluated below, only when a collision is a candidate
    } else {
      micro = majorant = ce::material_total(w, mat, p.energy, T, error, &ev, ce::cell_eval_slot(w, p.cell));
    }
    if (error) {
      out.error_physics = true;
      p.event = MMC_EV_CAPTURE;
      return;
    }
  } else {
    micro = majorant = material_micro_total(w, mat, p.group);  // Multigroup::GetMajorant == GetTotal
  }
  const double lambda = __dmul_rn(w.at<double>(w.h->off_mat_aden)[mat], majorant);
  // std::exponential_distribution: -log(1 - u) / lambda
  const double d_coll = __ddiv_rn(-glibc::log(__dsub_rn(1.0, p.rng.canonical())), lambda);
  int32_t nearest;
  const double d_surf = nearest_surface(w, p, nearest);
  bool cross;
  if (kTracking == MMC_TRACK_SURFACE) cross = !(d_coll < d_surf);
  else cross = d_surf < d_coll;
  if (cross) {
    if (kPerturb) perturb_stream(w, *pc, mat, majorant, __dadd_rn(d_surf, 2.220446049250313e-15));
    stream(p, __dadd_rn(d_surf, 2.220446049250313e-15));  // constants::nudge = 10 * epsilon
    p.surface = nearest;
    if (kDeferCross) {
      out.need_cross = true;
      p.event = kEvCrossPending;
      return;
    }
    finish_crossing(w, p, out);
  } else {
    bool real = true;
    if (kTracking == MMC_TRACK_CELL_DELTA) {
      // bernoulli_distribution{total / majorant}: u < p, both evaluated BEFORE the particle streams
      if (kCE) {
        bool error = false;
        micro = ce::material_total(w, mat, p.energy, ce::cell_temperature(w, p.cell, p.px, p.py, p.pz), error, &ev,
                                   ce::cell_eval_slot(w, p.cell));
        if (error) {
          out.error_physics = true;
          p.event = MMC_EV_CAPTURE;
          return;
        }
      }
      real = p.rng.canonical() < __ddiv_rn(micro, majorant);
    }
    if (kPerturb) perturb_stream(w, *pc, mat, majorant, d_coll);
    stream(p, d_coll);
    if (real) {
      if (kCE) ce::collide_continuous<kDeferTsl, kDeferFission>(w, p, mat, dq, out, ev);
      else collide_multigroup<kDeferDirection, kDeferFission>(w, p, mat, micro, dq, out);
    } else {
      p.event = MMC_EV_VIRTUAL_COLLISION;
    }
  }
}

// Collision ("implicit fission") estimator of k, the one KEigenvalue.hpp:33 lists as to be reimplemented: every real
// collision scores nu-bar Sigma_f / Sigma_t of the material at the particle's pre-collision energy, whatever the
// reaction sampled next.  Sums in nuclide order from 0.0; the caller adds the score as a fixed-point integer
// (MMC_K_COLLISION_ONE = 2^28) so that the generation's sum is independent of scheduling and GPU count.
template <bool kCE>
__device__ inline double implicit_fission_score(const WorldView& w, int32_t mat, uint64_t group, double E, double T) {
  const int32_t* nb = w.at<int32_t>(w.h->off_mat_nuc_begin);
  const int32_t* ni = w.at<int32_t>(w.h->off_mat_nuc_index);
  const double* af = w.at<double>(w.h->off_mat_nuc_afrac);
  double nu_fission = 0, total = 0;
  if (kCE) {
    const CeNuclide* nuclides = w.at<CeNuclide>(w.h->off_ce_nuclides);
    bool error = false;
    for (int32_t k = nb[mat]; k < nb[mat + 1]; k++) {
      const CeNuclide& n = nuclides[ni[k]];
      total = __dadd_rn(total, __dmul_rn(af[k], ce::nuclide_total(w, n, E, T, error)));
      for (int32_t r = 0; r < n.n_reactions; r++)
        if (n.reactions[r].kind == MMC_REACTION_FISSION && n.reactions[r].nubar.n)
          nu_fission = __dadd_rn(nu_fission, __dmul_rn(af[k], __dmul_rn(ce::table_at(w, n.reactions[r].nubar, E),
                                                                      ce::table_at(w, n.reactions[r].xs, E))));
    }
  } else {
    const int32_t G = w.h->n_groups;
    const double* tot = w.at<double>(w.h->off_mg_total);
    const double* fis = w.at<double>(w.h->off_mg_fission);
    const double* nubar = w.at<double>(w.h->off_mg_nubar);
    const uint32_t* mask = w.at<uint32_t>(w.h->off_mg_mask);
    for (int32_t k = nb[mat]; k < nb[mat + 1]; k++) {
      const int32_t i = ni[k] * G + static_cast<int32_t>(group - 1);
      total = __dadd_rn(total, __dmul_rn(af[k], tot[i]));
      if (mask[ni[k]] & MMC_REACTION_FISSION) nu_fission = __dadd_rn(nu_fission, __dmul_rn(af[k], __dmul_rn(nubar[i], fis[i])));
    }
  }
  return total > 0 ? __ddiv_rn(nu_fission, total) : 0.0;
}

// ------------------------------------------------------------------- tallies
// Bins::GetIndex for each concrete type, Bins.cpp:72-82,112-123,158-162
static __device__ __noinline__ uint64_t bins_index(const BinsSpec& b, const double* bounds, double v) {
  switch (b.kind) {
  case MMC_BINS_LINSPACE:
    if (v < b.lower) return 0;
    if (v >= b.upper) return b.n_bins - 1;
    return static_cast<uint64_t>(__dadd_rn(__ddiv_rn(__dsub_rn(v, b.lower), b.width), 1.0));
  case MMC_BINS_LOGSPACE: {
    const double lv = __ddiv_rn(glibc::log(v), glibc::log(b.base));
    if (lv < b.lower) return 0;
    if (lv >= b.upper) return b.n_bins - 1;
    return static_cast<uint64_t>(__dadd_rn(__ddiv_rn(__dsub_rn(lv, b.lower), b.width), 1.0));
  }
  case MMC_BINS_BOUNDARIES: {
    // std::upper_bound: first boundary > v
    const double* a = bounds + b.off_boundaries;
    uint32_t lo = 0, hi = b.n_bins - 1;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (v < a[mid]) hi = mid;
      else lo = mid + 1;
    }
    return lo;
  }
  default:
    return 0;
  }
}

// CurrentEstimator::GetScore (Estimator.cpp:142-151) + ParticleBins::GetIndex
// (Bins.cpp:196-204).  Returns true and the flattened bin when the score is 1.
template <bool kCE>
__device__ __forceinline__ bool estimator_score(
    const EstimatorSpec& e, const double* bounds, const Particle& p, uint64_t& bin) {
  if (p.surface != e.surface || (p.event != MMC_EV_SURFACE_CROSS && p.event != MMC_EV_LEAK)) return false;
  uint64_t ci = 0;
  if (e.has_direction) {
    const double mu = __dadd_rn(
        __dadd_rn(__dmul_rn(e.direction[0], p.dx), __dmul_rn(e.direction[1], p.dy)), __dmul_rn(e.direction[2], p.dz));
    ci = bins_index(e.cosine, bounds, mu);
  }
  // std::visit(VisitEnergy(), p.GetEnergy()): the energy, or the group as a double (Bins.cpp:196-204)
  const uint64_t ei = bins_index(e.energy, bounds, kCE ? p.energy : static_cast<double>(p.group));
  bin = e.offset + e.stride * ci + ei;
  return true;
}

}  // namespace mmc
