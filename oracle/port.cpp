// ORACLE TEST INFRASTRUCTURE -- not product code.
//
// CPU restatement of the reference's history transport loop on flat tables
// (agtumulak/minimc @ ed536a2; every function cites the file:line it follows).
// Plain scalar C++: one operation per statement, compiled with
// -ffp-contract=off for baseline x86-64, so every + - * / sqrt is a single
// IEEE-754 round-to-nearest operation as in the reference build.  log, sin, cos
// come from glibc, as they do for the reference.
//
// Parity status: PINNED.  tests/test_oracle.py checks this file against
//   * the reference's own code built in oracle/_ref (event traces, tallies),
//   * the goldens of SURVEY.md section 8(c) (G1-G5) stored in tests/golden/,
//   * libstdc++'s std::minstd_rand / generate_canonical on this box.
//
// Third-party arithmetic restated here: libstdc++ 13 <random>
// (linear_congruential_engine, generate_canonical, uniform_real_distribution,
// exponential_distribution, bernoulli_distribution) -- bits/random.h and
// bits/random.tcc of GCC 13.3.
#include "port.h"

#include <atomic>
#include <cmath>
#include <cstring>
#include <deque>
#include <limits>
#include <map>
#include <thread>
#include <vector>

namespace {

constexpr double kInf = std::numeric_limits<double>::infinity();
constexpr double kPi = 3.14159265358979323846;           // M_PI, Constants.hpp:13
constexpr double kNudge = 10 * 2.220446049250313e-16;    // Constants.hpp:15

enum Event { birth = 0, scatter, capture, fission, surface_cross, leak, virtual_collision };  // Particle.hpp:37-45

// std::minstd_rand: x <- 48271 x mod (2^31 - 1); seed s -> s mod m, 0 -> 1
struct MinStd {
  uint64_t x;
  explicit MinStd(uint64_t seed) {
    x = seed % 2147483647ull;
    if (x == 0) x = 1;
  }
  uint64_t operator()() {
    x = (x * 48271ull) % 2147483647ull;
    return x;
  }
};

// std::generate_canonical<double, 53>: k = 2 draws, range R = max - min + 1
double Canonical(MinStd& rng) {
  const double r = 2147483646.0;
  double sum = 0.0;
  double tmp = 1.0;
  sum += static_cast<double>(rng() - 1) * tmp;
  tmp *= r;
  sum += static_cast<double>(rng() - 1) * tmp;
  tmp = static_cast<double>(static_cast<long double>(r) * static_cast<long double>(r));
  double ret = sum / tmp;
  if (ret >= 1.0) ret = std::nextafter(1.0, 0.0);
  return ret;
}

// std::uniform_real_distribution{a, b}: u * (b - a) + a
double Uniform(MinStd& rng, double a, double b) { return Canonical(rng) * (b - a) + a; }

struct Vec {
  double x, y, z;
};
double Dot(const Vec& a, const Vec& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // Point.cpp:49-51
Vec Cross(const Vec& a, const Vec& b) {                                                 // Point.cpp:53-56
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
Vec Normalized(Vec v) {  // Point.cpp:44-47
  const double n = std::sqrt(Dot(v, v));
  return {v.x / n, v.y / n, v.z / n};
}
// Direction(RNG&), Point.cpp:89-96
Vec IsotropicDirection(MinStd& rng) {
  Vec d;
  d.x = Uniform(rng, -1., +1.);
  const double sin_theta = std::sqrt(1 - d.x * d.x);
  const double phi = Uniform(rng, 0., 2 * kPi);
  d.y = sin_theta * std::cos(phi);
  d.z = sin_theta * std::sin(phi);
  return d;
}
// Direction(d, mu, phi), Point.cpp:98-121
Vec Rotated(const Vec& d, double mu, double phi) {
  const bool off_xaxis = d.x <= 0.9 && d.x > -0.9;
  const Vec axis = off_xaxis ? Vec{1, 0, 0} : Vec{0, 1, 0};
  const Vec u = Normalized(Cross(d, axis));
  const Vec v = Normalized(Cross(d, u));
  const double a = std::sqrt(1 - mu * mu) * std::cos(phi);
  const double b = std::sqrt(1 - mu * mu) * std::sin(phi);
  const Vec uc{u.x * a, u.y * a, u.z * a}, vc{v.x * b, v.y * b, v.z * b}, dc{d.x * mu, d.y * mu, d.z * mu};
  const Vec uv{uc.x + vc.x, uc.y + vc.y, uc.z + vc.z};
  return Normalized({uv.x + dc.x, uv.y + dc.y, uv.z + dc.z});
}

struct Particle {
  Vec position{0, 0, 0};
  Vec direction{1, 0, 0};
  uint64_t group = 1;
  int32_t cell = -1;
  MinStd rng{1};
  int event = birth;
  int32_t current_surface = -1;
  std::vector<Particle> secondaries;
  bool IsAlive() const { return event != capture && event != leak && event != fission; }  // Particle.cpp:126-129
};

// CSGSurface::SolveQuadratic, CSGSurface.cpp:71-97
double SolveQuadratic(double a, double b, double c) {
  const double discriminant = b * b - 4 * a * c;
  if (discriminant <= 0) return kInf;
  const double lesser = b > 0 ? (-b - std::sqrt(discriminant)) / (2 * a) : (2 * c) / (-b + std::sqrt(discriminant));
  const double greater = b > 0 ? (2 * c) / (-b - std::sqrt(discriminant)) : (-b + std::sqrt(discriminant)) / (2 * a);
  if (lesser > 0) return lesser;
  if (greater > 0) return greater;
  return kInf;
}

struct World {
  const orc_world& w;
  // CSGSurface.cpp:107-111, 125-134, 148-164
  double Distance(int32_t s, const Vec& p, const Vec& d) const {
    const double* q = w.surface_param + 4 * s;
    switch (w.surface_type[s]) {
    case 0: {
      const Vec oc{p.x - q[0], p.y - q[1], p.z - q[2]};
      return SolveQuadratic(1, 2 * Dot(oc, d), Dot(oc, oc) - q[3] * q[3]);
    }
    case 1: {
      const double v_x = q[0] - Dot(p, Vec{1, 0, 0});
      const double d_x = Dot(d, Vec{1, 0, 0});
      const double dist = v_x / d_x;
      return dist > 0 ? dist : kInf;
    }
    default: {
      const Vec ax{1, 0, 0};
      const double pp = Dot(p, p), pw = Dot(p, ax), dw = Dot(d, ax), pd = Dot(p, d);
      return SolveQuadratic(1 - dw * dw, 2 * (pd - pw * dw), pp - pw * pw - q[0] * q[0]);
    }
    }
  }
  // CSGSurface.cpp:113-116, 136-138, 166-174 (quirk Q2 kept)
  bool Contains(int32_t s, const Vec& p) const {
    const double* q = w.surface_param + 4 * s;
    switch (w.surface_type[s]) {
    case 0: {
      const Vec pc{p.x - q[0], p.y - q[1], p.z - q[2]};
      return Dot(pc, pc) < q[3] * q[3];
    }
    case 1:
      return Dot(p, Vec{1, 0, 0}) < q[0];
    default: {
      const Vec ax{1, 0, 0};
      const double pp = Dot(p, p), pw = Dot(p, ax);
      return std::sqrt(pp - pw * pw) < q[0] * q[0];
    }
    }
  }
  // World::FindCellContaining, World.cpp:26-37 + Cell::Contains, Cell.cpp:27-35
  int32_t FindCellContaining(const Vec& p) const {
    for (int32_t c = 0; c < w.n_cells; c++) {
      bool all = true;
      for (int32_t k = w.cell_surface_begin[c]; k < w.cell_surface_begin[c + 1] && all; k++)
        all = Contains(w.cell_surface_index[k], p) == (w.cell_surface_sense[k] != 0);
      if (all) return c;
    }
    return -1;
  }
  // Cell::NearestSurface, Cell.cpp:37-51 (std::min_element: first minimum)
  double NearestSurface(int32_t cell, const Vec& p, const Vec& d, int32_t& nearest) const {
    int32_t best_k = w.cell_surface_begin[cell];
    for (int32_t k = best_k + 1; k < w.cell_surface_begin[cell + 1]; k++)
      if (Distance(w.cell_surface_index[k], p, d) < Distance(w.cell_surface_index[best_k], p, d)) best_k = k;
    nearest = w.cell_surface_index[best_k];
    return Distance(nearest, p, d);
  }
  double NuclideTotal(int32_t nuc, uint64_t g) const { return w.mg_total[nuc * w.n_groups + (g - 1)]; }
  // Material::GetMicroscopicTotal / Majorant, Material.cpp:41-62 (Multigroup majorant == total)
  double MicroscopicTotal(int32_t mat, uint64_t g) const {
    double acc = 0;
    for (int32_t k = w.material_nuclide_begin[mat]; k < w.material_nuclide_begin[mat + 1]; k++)
      acc = acc + w.material_nuclide_afrac[k] * NuclideTotal(w.material_nuclide_index[k], g);
    return acc;
  }
  double ImplicitFission(int32_t mat, uint64_t g) const {
    double nu_fission = 0, total = 0;
    for (int32_t k = w.material_nuclide_begin[mat]; k < w.material_nuclide_begin[mat + 1]; k++) {
      const int32_t nuc = w.material_nuclide_index[k];
      const int32_t i = nuc * w.n_groups + static_cast<int32_t>(g - 1);
      total = total + w.material_nuclide_afrac[k] * w.mg_total[i];
      if (w.mg_reaction_mask[nuc] & 4u) nu_fission = nu_fission + w.material_nuclide_afrac[k] * (w.mg_nubar[i] * w.mg_fission[i]);
    }
    return total > 0 ? nu_fission / total : 0.0;
  }
};

struct Errors {
  uint64_t lost = 0, physics = 0;
};

// Particle::SampleNuclide, Particle.cpp:110-124
int32_t SampleNuclide(const World& W, Particle& p, int32_t mat) {
  const orc_world& w = W.w;
  const double threshold = W.MicroscopicTotal(mat, p.group) * Canonical(p.rng);
  double accumulated = 0;
  for (int32_t k = w.material_nuclide_begin[mat]; k < w.material_nuclide_begin[mat + 1]; k++) {
    accumulated += w.material_nuclide_afrac[k] * W.NuclideTotal(w.material_nuclide_index[k], p.group);
    if (accumulated > threshold) return w.material_nuclide_index[k];
  }
  return -1;
}

// Multigroup::Interact + Capture / Scatter / Fission, Multigroup.cpp:49-73, 245-287
void Interact(const World& W, Particle& p, int32_t nuc, Errors& err, orc_counters& cnt) {
  const orc_world& w = W.w;
  const int32_t G = w.n_groups;
  const size_t row = static_cast<size_t>(nuc) * G + (p.group - 1);
  const double threshold = Canonical(p.rng) * W.NuclideTotal(nuc, p.group);
  double accumulated = 0;
  const uint32_t mask = w.mg_reaction_mask[nuc];
  const double* xs[3] = {w.mg_capture, w.mg_scatter, w.mg_fission};
  for (int reaction = 0; reaction < 3; reaction++) {  // std::map<Reaction,...>: enum order
    if (!(mask & (1u << reaction))) continue;
    accumulated += xs[reaction][row];
    if (!(accumulated > threshold)) continue;
    if (reaction == 0) {
      p.event = capture;
    } else if (reaction == 1) {
      p.event = scatter;
      const double t = Canonical(p.rng);
      double acc = 0;
      const double* probs = w.mg_scatter_probs + row * G;
      for (int32_t g = 1; g <= G; g++) {
        acc += probs[g - 1];
        if (acc > t) {
          p.group = g;
          p.direction = IsotropicDirection(p.rng);
          return;
        }
      }
      err.physics++;  // assert(false)
      p.event = capture;
    } else {
      p.event = fission;
      const size_t yield = static_cast<size_t>(w.mg_nubar[row] + Canonical(p.rng));
      const double* chi = w.mg_chi + row * G;
      for (size_t i = 0; i < yield; i++) {
        const double t = Canonical(p.rng);
        double acc = 0;
        for (int32_t g = 1; g <= G; g++) {
          acc += chi[g - 1];
          if (acc > t) {
            // p.BankSecondaries(Direction{p.rng}, g): Particle.cpp:96-100
            Particle s;
            s.position = p.position;
            s.direction = IsotropicDirection(p.rng);
            s.group = g;
            s.cell = p.cell;
            s.rng = MinStd{p.rng()};
            p.secondaries.push_back(s);
            cnt.n_secondaries++;
            break;
          }
        }
      }
    }
    return;
  }
  err.physics++;  // assert(false)
  p.event = capture;
}

void Stream(Particle& p, double distance) {  // Particle.cpp:46-53
  p.position.x += p.direction.x * distance;
  p.position.y += p.direction.y * distance;
  p.position.z += p.direction.z * distance;
}

// SurfaceTracking::Transport (TransportMethod.cpp:52-77) and
// CellDeltaTracking::Transport (TransportMethod.cpp:88-122).  `score` is
// EstimatorSetProxy::Score.
template <typename ScoreFn>
void Transport(const World& W, Particle& p, int tracking, Errors& err, orc_counters& cnt, ScoreFn&& score,
               uint64_t* k_collision_fixed = nullptr) {
  const orc_world& w = W.w;
  p.cell = W.FindCellContaining(p.position);
  if (p.cell < 0) {
    err.lost++;
    return;
  }
  while (p.IsAlive()) {
    const int32_t mat = w.cell_material[p.cell];
    if (mat < 0) {  // null Material dereference in the reference
      err.physics++;
      return;
    }
    const double micro = W.MicroscopicTotal(mat, p.group);
    const double lambda = w.material_aden[mat] * micro;
    // std::exponential_distribution{lambda}
    const double distance_to_collision = -std::log(1.0 - Canonical(p.rng)) / lambda;
    int32_t nearest;
    const double distance_to_surface = W.NearestSurface(p.cell, p.position, p.direction, nearest);
    const bool cross = tracking == 0 ? !(distance_to_collision < distance_to_surface)
                                     : distance_to_surface < distance_to_collision;
    cnt.n_events++;
    if (cross) {
      Stream(p, distance_to_surface + kNudge);
      p.cell = W.FindCellContaining(p.position);
      p.current_surface = nearest;
      if (p.cell < 0) {
        err.lost++;
        p.event = leak;
        return;
      }
      p.event = w.cell_material[p.cell] >= 0 ? surface_cross : leak;
      cnt.n_crossings++;
    } else {
      bool real = true;
      if (tracking == 1) {
        // std::bernoulli_distribution{total / majorant}: u < p
        const double prob = W.MicroscopicTotal(mat, p.group) / W.MicroscopicTotal(mat, p.group);
        real = Canonical(p.rng) < prob;
      }
      Stream(p, distance_to_collision);
      if (real) {
        // collision ("implicit fission", KEigenvalue.hpp:33) estimator of k: nu Sigma_f / Sigma_t of the material at the
        // pre-collision group, as a fixed-point integer (2^28 per unit) -- DESIGN.md "k-eigenvalue"
        if (k_collision_fixed) *k_collision_fixed += static_cast<uint64_t>(std::llrint(W.ImplicitFission(mat, p.group) * 268435456.0));
        const int32_t nuc = SampleNuclide(W, p, mat);
        if (nuc < 0) {
          err.physics++;
          p.event = capture;
        } else {
          Interact(W, p, nuc, err, cnt);
        }
        cnt.n_collisions++;
      } else {
        p.event = virtual_collision;
        cnt.n_virtual++;
      }
    }
    score(p);
  }
}

// Source::Sample, Source.cpp:143-154
Particle SampleSource(const orc_source& s, uint64_t seed) {
  MinStd rng{seed};
  Particle p;
  p.position = {s.position[0], s.position[1], s.position[2]};
  const Vec ref = Normalized({s.direction[0], s.direction[1], s.direction[2]});
  if (s.direction_kind == 1) {
    p.direction = IsotropicDirection(rng);
  } else if (s.direction_kind == 2) {
    const double mu = std::sqrt(Canonical(rng));  // Source.cpp:124-129
    const double phi = Uniform(rng, 0., 2 * kPi);
    p.direction = Rotated(ref, mu, phi);
  } else {
    p.direction = ref;
  }
  p.group = s.group;
  p.rng = MinStd{rng()};
  return p;
}

// Bins::GetIndex, Bins.cpp:45,72-82,112-123,158-162
size_t BinsIndex(const orc_bins& b, double v) {
  switch (b.kind) {
  case 1:
    if (v < b.lower) return 0;
    if (v >= b.upper) return b.n_bins - 1;
    return static_cast<size_t>((v - b.lower) / b.width + 1);
  case 2: {
    const double log_v = std::log(v) / std::log(b.base);
    if (log_v < b.lower) return 0;
    if (log_v >= b.upper) return b.n_bins - 1;
    return static_cast<size_t>((log_v - b.lower) / b.width + 1);
  }
  case 3: {
    size_t i = 0;
    while (i < b.n_bins - 1 && !(v < b.boundaries[i])) i++;  // upper_bound
    return i;
  }
  default:
    return 0;
  }
}

size_t BinsSize(const orc_bins& b) { return b.kind == 0 ? 1 : b.n_bins; }

struct Tally {
  const orc_estimator* estimators;
  int32_t n;
  std::vector<size_t> offset;
  std::vector<Vec> direction;
  size_t total = 0;
  Tally(const orc_estimator* e, int32_t n) : estimators(e), n(n) {
    for (int32_t i = 0; i < n; i++) {
      offset.push_back(total);
      total += BinsSize(e[i].cosine) * BinsSize(e[i].energy);
      direction.push_back(Normalized({e[i].cosine_direction[0], e[i].cosine_direction[1], e[i].cosine_direction[2]}));
    }
  }
  // ScorableProxy::Score (Scorable.cpp:81-99) with CurrentEstimator::GetScore
  // (Estimator.cpp:142-151) and ParticleBins::GetIndex (Bins.cpp:196-204)
  void Score(const Particle& p, std::map<size_t, double>& pending, orc_counters& cnt) const {
    for (int32_t i = 0; i < n; i++) {
      const orc_estimator& e = estimators[i];
      if (!(p.current_surface == e.surface && (p.event == surface_cross || p.event == leak))) continue;
      const size_t c_i = e.has_cosine_direction ? BinsIndex(e.cosine, Dot(direction[i], p.direction)) : 0;
      const size_t e_i = BinsIndex(e.energy, static_cast<double>(p.group));
      pending[offset[i] + BinsSize(e.energy) * c_i + e_i] += 1;
      cnt.n_scores++;
    }
  }
};

void Add(orc_counters& a, const orc_counters& b) {
  a.n_histories += b.n_histories;
  a.n_births += b.n_births;
  a.n_events += b.n_events;
  a.n_collisions += b.n_collisions;
  a.n_crossings += b.n_crossings;
  a.n_virtual += b.n_virtual;
  a.n_scores += b.n_scores;
  a.n_secondaries += b.n_secondaries;
  a.n_lost += b.n_lost;
  a.n_physics_errors += b.n_physics_errors;
}

}  // namespace

extern "C" {

int orc_fixed_source_run(
    const orc_world* world, const orc_source* source, const orc_estimator* estimators, int32_t n_estimators,
    uint64_t seed0, uint64_t first, uint64_t n, int32_t tracking, int32_t threads, double* scores,
    double* square_scores, orc_counters* counters) {
  const World W{*world};
  const Tally tally{estimators, n_estimators};
  if (threads < 1) threads = 1;
  std::atomic<uint64_t> histories_elapsed{0};
  std::vector<std::vector<double>> t_scores(threads, std::vector<double>(tally.total, 0.0));
  std::vector<std::vector<double>> t_squares(threads, std::vector<double>(tally.total, 0.0));
  std::vector<orc_counters> t_counters(threads, orc_counters{});
  // FixedSource::StartWorker, FixedSource.cpp:40-77
  auto worker = [&](int t) {
    orc_counters& cnt = t_counters[t];
    Errors err;
    while (true) {
      std::map<size_t, double> pending;  // scoring_proxy
      const uint64_t elapsed = histories_elapsed++;
      if (elapsed >= n) break;
      cnt.n_histories++;
      std::deque<Particle> bank;
      bank.push_back(SampleSource(*source, seed0 + first + elapsed));
      while (!bank.empty()) {
        Particle& p = bank.back();
        cnt.n_births++;
        Transport(W, p, tracking, err, cnt, [&](const Particle& q) { tally.Score(q, pending, cnt); });
        // MoveSecondariesTo: splice to the front, in creation order (Bank.cpp:5-8)
        std::vector<Particle> secondaries = std::move(p.secondaries);
        bank.pop_back();
        bank.insert(bank.begin(), secondaries.begin(), secondaries.end());
      }
      // ScorableProxy::CommitHistory, Scorable.cpp:101-106
      for (const auto& [index, score] : pending) {
        t_scores[t][index] += score;
        t_squares[t][index] += score * score;
      }
    }
    cnt.n_lost = err.lost;
    cnt.n_physics_errors = err.physics;
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < threads; t++) pool.emplace_back(worker, t);
  worker(0);
  for (auto& th : pool) th.join();
  orc_counters total{};
  for (int t = 0; t < threads; t++) {
    Add(total, t_counters[t]);
    for (size_t i = 0; i < tally.total; i++) {  // Scorable::operator+=
      scores[i] += t_scores[t][i];
      square_scores[i] += t_squares[t][i];
    }
  }
  if (counters) *counters = total;
  return (total.n_lost || total.n_physics_errors) ? 1 : 0;
}

size_t orc_trace(
    const orc_world* world, const orc_source* source, uint64_t seed0, uint64_t first, uint64_t n, int32_t tracking,
    orc_record* records, size_t cap) {
  const World W{*world};
  size_t count = 0;
  orc_counters cnt{};
  Errors err;
  auto emit = [&](uint64_t history, uint32_t ordinal, const Particle& p) {
    if (count < cap) {
      orc_record& r = records[count];
      r.history = history;
      r.particle = ordinal;
      r.event = p.event;
      r.group = p.group;
      r.cell = p.cell;
      r.surface = p.current_surface;
      r.position[0] = p.position.x;
      r.position[1] = p.position.y;
      r.position[2] = p.position.z;
      r.direction[0] = p.direction.x;
      r.direction[1] = p.direction.y;
      r.direction[2] = p.direction.z;
      r.rng_state = p.rng.x;
    }
    count++;
  };
  for (uint64_t h = first; h < first + n; h++) {
    std::deque<Particle> bank;
    bank.push_back(SampleSource(*source, seed0 + h));
    uint32_t ordinal = 0;
    while (!bank.empty()) {
      Particle& p = bank.back();
      emit(h, ordinal, p);
      Transport(W, p, tracking, err, cnt, [&](const Particle& q) { emit(h, ordinal, q); });
      std::vector<Particle> secondaries = std::move(p.secondaries);
      bank.pop_back();
      bank.insert(bank.begin(), secondaries.begin(), secondaries.end());
      ordinal++;
    }
  }
  return count;
}

// K-eigenvalue power iteration as DEFINED by this repo (DESIGN.md "k-eigenvalue";
// the reference's KEigenvalue::Solve is a stub, KEigenvalue.cpp:36-62 -- parity
// with the reference is therefore UNPINNED for this function; it pins the CUDA
// path against an independent CPU statement of the same definition and against
// analytic k-infinity).  Sequential on purpose: the definition is order-based.
int orc_keigenvalue_run(
    const orc_world* world, const orc_source* source, const orc_estimator* estimators, int32_t n_estimators,
    uint64_t batchsize, uint64_t inactive, uint64_t active, int32_t tracking, double* scores, double* square_scores,
    double* k_cycle, uint64_t* bank_sizes, orc_counters* counters, double* k_collision_cycle) {
  const World W{*world};
  const Tally tally{estimators, n_estimators};
  orc_counters cnt{};
  Errors err;
  struct Site {
    Vec position, direction;
    uint64_t group;
    uint64_t seed;
  };
  // KEigenvalue.cpp:29-33: source.Sample(s), s = 1 .. batchsize
  std::vector<Site> bank;
  for (uint64_t s = 1; s <= batchsize; s++) {
    const Particle p = SampleSource(*source, s);
    bank.push_back({p.position, p.direction, p.group, p.rng.x});
  }
  for (uint64_t cycle = 0; cycle < inactive + active; cycle++) {
    const bool score = cycle >= inactive;
    uint64_t k_fixed = 0;
    std::vector<Site> fission_bank;
    for (const Site& site : bank) {
      Particle p;
      p.position = site.position;
      p.direction = site.direction;
      p.group = site.group;
      p.rng = MinStd{site.seed};
      std::map<size_t, double> pending;
      cnt.n_histories++;
      cnt.n_births++;
      orc_counters scratch{};
      Transport(W, p, tracking, err, cnt, [&](const Particle& q) {
        if (score) tally.Score(q, pending, cnt);
        else tally.Score(q, pending, scratch);
      }, &k_fixed);
      if (score)
        for (const auto& [index, s] : pending) {
          scores[index] += s;
          square_scores[index] += s * s;
        }
      for (const Particle& child : p.secondaries)  // creation order
        fission_bank.push_back({child.position, child.direction, child.group, child.rng.x});
    }
    const uint64_t M = fission_bank.size(), N = batchsize;
    k_cycle[cycle] = static_cast<double>(M) / static_cast<double>(N);
    bank_sizes[cycle] = M;
    if (k_collision_cycle) k_collision_cycle[cycle] = static_cast<double>(k_fixed) / 268435456.0 / static_cast<double>(N);
    if (M == 0) break;
    // comb resampling: source i <- site floor(i * M / N); copies get seed + copy ordinal
    std::vector<Site> next;
    for (uint64_t i = 0; i < N; i++) {
      const uint64_t j = static_cast<uint64_t>(static_cast<unsigned __int128>(i) * M / N);
      const uint64_t i0 = static_cast<uint64_t>((static_cast<unsigned __int128>(j) * N + M - 1) / M);
      Site s = fission_bank[j];
      s.seed = static_cast<uint32_t>(s.seed + (i - i0));
      next.push_back(s);
    }
    bank.swap(next);
  }
  cnt.n_lost = err.lost;
  cnt.n_physics_errors = err.physics;
  if (counters) *counters = cnt;
  return (cnt.n_lost || cnt.n_physics_errors) ? 1 : 0;
}

void orc_rng_canonical(uint64_t seed, size_t n, double* u, uint64_t* state) {
  MinStd rng{seed};
  for (size_t i = 0; i < n; i++) {
    u[i] = Canonical(rng);
    state[i] = rng.x;
  }
}

}  // extern "C"
