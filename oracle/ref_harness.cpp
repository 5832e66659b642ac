// ORACLE TEST INFRASTRUCTURE -- not product code.
//
// Harness around the UNMODIFIED reference sources (/root/reference/src/*.cpp,
// agtumulak/minimc @ ed536a2) linked behind the shims in oracle/shim/. Built by
// oracle/Makefile into oracle/_ref/ref_harness (git-ignored, travels with
// gpurun). This translation unit alone is compiled with -fno-access-control so
// it can read private members of the reference classes; no reference source is
// copied or edited.
//
// Commands
//   run   <deck.xml>                 Driver::Create(path)->Solve() exactly as
//                                    minimc.cpp:16-21; prints the .out text to
//                                    stdout and "solve_seconds=<s>" to stderr.
//   trace <deck.xml> <first> <count> per-event records of histories
//                                    [first, first+count) produced by the
//                                    reference's own TransportMethod::Transport
//                                    through an extra Estimator whose GetScore()
//                                    logs the Particle (Scorable.cpp:81-84 calls
//                                    it once per event) and returns 0.
//   dump  <deck.xml>                 the World, Source and Estimators flattened
//                                    in the reference containers' own iteration
//                                    order (quirk Q1), as JSON with C99 hex
//                                    floats, to pin the product's flattening.
//   rng   <seed> <n>                 n canonical doubles from minstd_rand{seed}
//                                    (libstdc++ generate_canonical), hex floats.
#include "Bins.hpp"
#include "CSGSurface.hpp"
#include "Cell.hpp"
#include "Driver.hpp"
#include "Estimator.hpp"
#include "FixedSource.hpp"
#include "Material.hpp"
#include "Multigroup.hpp"
#include "Nuclide.hpp"
#include "Particle.hpp"
#include "Perturbation.hpp"
#include "ScalarField.hpp"
#include "Source.hpp"
#include "TransportMethod.hpp"
#include "World.hpp"
#include "XMLDocument.hpp"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <variant>

namespace {

std::string Hex(double v) {
  char buf[64];
  std::snprintf(buf, sizeof(buf), "%a", v);
  return buf;
}

unsigned long RngState(const RNG& rng) {
  std::stringstream ss;
  ss << rng;  // linear_congruential_engine streams its state
  unsigned long s;
  ss >> s;
  return s;
}

const World* g_world = nullptr;
long g_history = 0;
long g_particle = 0;

long CellIndex(const Particle& p) { return p.cell ? p.cell - g_world->cells.data() : -1; }

long SurfaceIndex(const Particle& p) {
  if (!p.current_surface) return -1;
  for (size_t i = 0; i < g_world->surfaces.size(); i++)
    if (g_world->surfaces[i] == p.current_surface) return static_cast<long>(i);
  return -2;
}

void PrintParticle(const char* tag, const Particle& p) {
  const auto& e = p.GetEnergy();
  std::printf(
      "%s %ld %ld %d %s %ld %ld %s %s %s %s %s %s %lu\n", tag, g_history, g_particle,
      static_cast<int>(p.event),
      e.index() == 1 ? ("g" + std::to_string(std::get<Group>(e))).c_str()
                     : ("e" + Hex(std::get<ContinuousEnergy>(e))).c_str(),
      CellIndex(p), SurfaceIndex(p), Hex(p.position.x).c_str(), Hex(p.position.y).c_str(),
      Hex(p.position.z).c_str(), Hex(p.direction.x).c_str(), Hex(p.direction.y).c_str(),
      Hex(p.direction.z).c_str(), RngState(p.rng));
}

// Logs every event of every particle; scores nothing.
class TraceEstimator : public Estimator {
public:
  explicit TraceEstimator(const PerturbationSet& perturbations)
      : Estimator{pugi::xml_node{}, perturbations} {}
  TraceEstimator(const TraceEstimator& other) : Estimator{other} {}
  std::unique_ptr<Estimator> Clone() const noexcept override {
    return std::make_unique<TraceEstimator>(*this);
  }
  Real GetScore(const Particle& p) const noexcept override {
    PrintParticle("E", p);
    return 0;
  }
};

int Run(const char* path) {
  auto driver = Driver::Create(path);
  // the reference prints progress with '\r' from inside the history loop, from
  // every worker thread: discard it through a stateless (thread-safe) buffer
  struct NullBuffer : std::streambuf {
    int overflow(int c) override { return traits_type::not_eof(c); }
    std::streamsize xsputn(const char*, std::streamsize n) override { return n; }
  } sink;
  auto* old = std::cout.rdbuf(&sink);
  const auto t0 = std::chrono::steady_clock::now();
  const auto result = driver->Solve();
  const auto t1 = std::chrono::steady_clock::now();
  std::cout.rdbuf(old);
  std::cout << driver->batchsize << std::endl;
  std::cout << result.to_string();
  std::fprintf(stderr, "solve_seconds=%.6f\n", std::chrono::duration<double>(t1 - t0).count());
  return 0;
}

int Trace(const char* path, long first, long count) {
  XMLDocument doc{path};
  FixedSource driver{doc.root};
  g_world = &driver.world;
  auto& estimators = const_cast<std::vector<std::unique_ptr<Estimator>>&>(
      driver.init_estimator_set.estimators);
  estimators.push_back(std::make_unique<TraceEstimator>(driver.perturbations));
  for (long h = first; h < first + count; h++) {
    g_history = h;
    g_particle = 0;
    auto proxy = driver.init_estimator_set.GetProxy();
    // FixedSource.cpp:59-72 restated so that records can be grouped
    Bank bank;
    bank.emplace_back(driver.source.Sample(driver.seed + h));
    while (!bank.empty()) {
      PrintParticle("B", bank.back());
      bank.back().SetPerturbations(driver.perturbations);
      bank.back().Transport(proxy, driver.world);
      bank.back().MoveSecondariesTo(bank);
      bank.pop_back();
      g_particle++;
    }
  }
  return 0;
}

void DumpVector(const char* key, const std::vector<Real>& v, bool last = false) {
  std::printf("\"%s\": [", key);
  for (size_t i = 0; i < v.size(); i++) std::printf("%s\"%s\"", i ? ", " : "", Hex(v[i]).c_str());
  std::printf("]%s", last ? "" : ", ");
}

int Dump(const char* path) {
  XMLDocument doc{path};
  FixedSource driver{doc.root};
  const World& w = driver.world;
  auto surface_index = [&w](const std::shared_ptr<const CSGSurface>& s) {
    for (size_t i = 0; i < w.surfaces.size(); i++)
      if (w.surfaces[i] == s) return static_cast<long>(i);
    return -1L;
  };
  auto nuclide_index = [&w](const std::shared_ptr<const Nuclide>& s) {
    for (size_t i = 0; i < w.nuclides.size(); i++)
      if (w.nuclides[i] == s) return static_cast<long>(i);
    return -1L;
  };
  auto material_index = [&w](const std::shared_ptr<const Material>& s) {
    if (!s) return -1L;
    for (size_t i = 0; i < w.materials.size(); i++)
      if (w.materials[i] == s) return static_cast<long>(i);
    return -2L;
  };
  std::printf("{\n\"batchsize\": %lu, \"seed\": %lu, \"threads\": %zu,\n", driver.batchsize, driver.seed,
              driver.threads);
  std::printf("\"tracking\": \"%s\",\n",
              dynamic_cast<const SurfaceTracking*>(Particle::transport_method.get()) ? "surface" : "cell delta");
  std::printf("\"surfaces\": [\n");
  for (size_t i = 0; i < w.surfaces.size(); i++) {
    const auto* s = w.surfaces[i].get();
    std::printf("  {\"name\": \"%s\", ", s->name.c_str());
    if (const auto* sp = dynamic_cast<const Sphere*>(s)) {
      std::printf("\"type\": \"sphere\", \"params\": [\"%s\", \"%s\", \"%s\", \"%s\"]}", Hex(sp->center.x).c_str(),
                  Hex(sp->center.y).c_str(), Hex(sp->center.z).c_str(), Hex(sp->radius).c_str());
    } else if (const auto* pl = dynamic_cast<const PlaneX*>(s)) {
      std::printf("\"type\": \"planex\", \"params\": [\"%s\"]}", Hex(pl->c).c_str());
    } else if (const auto* cy = dynamic_cast<const CylinderX*>(s)) {
      std::printf("\"type\": \"cylinderx\", \"params\": [\"%s\"]}", Hex(cy->radius).c_str());
    }
    std::printf("%s\n", i + 1 < w.surfaces.size() ? "," : "");
  }
  std::printf("],\n\"nuclides\": [\n");
  for (size_t i = 0; i < w.nuclides.size(); i++) {
    const auto& n = *w.nuclides[i];
    std::printf("  {\"name\": \"%s\", ", n.name.c_str());
    const auto* mg = dynamic_cast<const Multigroup*>(n.xs.at(Particle::Type::neutron).get());
    if (mg) {
      const auto G = mg->max_group;
      std::printf("\"kind\": \"multigroup\", \"groups\": %lu, ", G);
      DumpVector("total", mg->total.elements);
      std::printf("\"reactions\": {");
      bool first = true;
      for (const auto& [reaction, xs] : mg->reactions) {
        std::printf("%s", first ? "" : ", ");
        first = false;
        const char* name = reaction == Reaction::capture ? "capture" : reaction == Reaction::scatter ? "scatter" : "fission";
        DumpVector(name, xs.elements, true);
      }
      std::printf("}, ");
      auto dump2d = [G](const char* key, const auto& opt) {
        std::vector<Real> flat;  // [g_in-1][g_out-1]
        if (opt.has_value())
          for (Group gi = 1; gi <= G; gi++)
            for (Group go = 1; go <= G; go++) flat.push_back(opt.value().at(gi).at(go));
        DumpVector(key, flat);
      };
      dump2d("scatter_probs", mg->scatter_probs);
      dump2d("chi", mg->chi);
      DumpVector("nubar", mg->nubar.has_value() ? mg->nubar.value().elements : std::vector<Real>{}, true);
    } else {
      std::printf("\"kind\": \"continuous\"");
    }
    std::printf("}%s\n", i + 1 < w.nuclides.size() ? "," : "");
  }
  std::printf("],\n\"materials\": [\n");
  for (size_t i = 0; i < w.materials.size(); i++) {
    const auto& m = *w.materials[i];
    std::printf("  {\"name\": \"%s\", \"aden\": \"%s\", \"afracs\": [", m.name.c_str(), Hex(m.number_density).c_str());
    bool first = true;
    for (const auto& [nuclide, afrac] : m.afracs) {  // pointer order: quirk Q1
      std::printf("%s[%ld, \"%s\"]", first ? "" : ", ", nuclide_index(nuclide), Hex(afrac).c_str());
      first = false;
    }
    std::printf("]}%s\n", i + 1 < w.materials.size() ? "," : "");
  }
  std::printf("],\n\"cells\": [\n");
  for (size_t i = 0; i < w.cells.size(); i++) {
    const auto& c = w.cells[i];
    std::printf("  {\"name\": \"%s\", \"material\": %ld, \"temperature_upper\": \"%s\", \"surfaces\": [", c.name.c_str(),
                material_index(c.material), Hex(c.temperature->upper_bound).c_str());
    bool first = true;
    for (const auto& [surface, sense] : c.surface_senses) {  // pointer order: quirk Q1
      std::printf("%s[%ld, %d]", first ? "" : ", ", surface_index(surface), sense ? 1 : 0);
      first = false;
    }
    std::printf("]}%s\n", i + 1 < w.cells.size() ? "," : "");
  }
  std::printf("],\n\"estimators\": [\n");
  const auto& estimators = driver.init_estimator_set.estimators;
  for (size_t i = 0; i < estimators.size(); i++) {
    const auto* ce = dynamic_cast<const CurrentEstimator*>(estimators[i].get());
    std::printf("  {\"name\": \"%s\", \"surface\": %ld, \"n_bins\": %zu}%s\n", ce->name.c_str(),
                surface_index(ce->surface), ce->bins->size(), i + 1 < estimators.size() ? "," : "");
  }
  std::printf("]\n}\n");
  return 0;
}

int Rng(unsigned long seed, long n) {
  RNG rng{seed};
  for (long i = 0; i < n; i++) {
    const double u = std::generate_canonical<double, 53>(rng);
    std::printf("%s %lu\n", Hex(u).c_str(), RngState(rng));
  }
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  try {
    if (argc >= 3 && !std::strcmp(argv[1], "run")) return Run(argv[2]);
    if (argc >= 5 && !std::strcmp(argv[1], "trace")) return Trace(argv[2], std::atol(argv[3]), std::atol(argv[4]));
    if (argc >= 3 && !std::strcmp(argv[1], "dump")) return Dump(argv[2]);
    if (argc >= 4 && !std::strcmp(argv[1], "rng")) return Rng(std::strtoul(argv[2], nullptr, 10), std::atol(argv[3]));
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 3;
  }
  std::fprintf(stderr, "usage: ref_harness run|trace|dump|rng ...\n");
  return 2;
}
