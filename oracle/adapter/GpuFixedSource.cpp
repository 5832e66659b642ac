// ORACLE TEST INFRASTRUCTURE -- not product code, and the model of what a maintainer of agtumulak/minimc would add.
//
// INTEGRATION.md, way B, compiled for real: the REFERENCE'S OWN classes (/root/reference/src, unmodified, behind the
// shims of oracle/shim) keep everything -- XML, World, Source, EstimatorSet, the .out text -- and one new Driver
// subclass, GpuFixedSource, replaces what happens inside FixedSource::Solve() (src/FixedSource.cpp:22-36: `threads` x
// std::async(StartWorker)) by ONE call into the C ABI of include/minimc_b200.h:
//
//     Flatten(world)  ->  mmc_world_create  ->  mmc_fixed_source_run  ->  Scorable::scores / square_scores
//
// Built by oracle/Makefile (target `adapter`) into oracle/_ref/gpu_adapter, linked against
// minimc_b200/libminimc_b200.so.  Usage: gpu_adapter <deck.xml> prints what runminimc writes to <deck>.out
// (minimc.cpp:20-21).  tests/test_gpu_adapter.py asserts that this text equals the golden .out files the reference
// binary wrote, byte for byte.
//
// This file is compiled with -fno-access-control because the reference keeps its tables private; in the reference
// tree the same code needs `friend class GpuFixedSource;` in Multigroup, Continuous, ContinuousReaction,
// ThermalScattering, HDF5DataSet, Source, Scorable, ParticleBins and the Bins classes (or accessors).
#include "Bins.hpp"
#include "CSGSurface.hpp"
#include "Cell.hpp"
#include "Continuous.hpp"
#include "ContinuousReaction.hpp"
#include "Driver.hpp"
#include "Estimator.hpp"
#include "FixedSource.hpp"
#include "Material.hpp"
#include "Multigroup.hpp"
#include "Nuclide.hpp"
#include "Particle.hpp"
#include "ScalarField.hpp"
#include "Source.hpp"
#include "ThermalScattering.hpp"
#include "TransportMethod.hpp"
#include "World.hpp"
#include "XMLDocument.hpp"

#include "../../include/minimc_b200.h"

#include <algorithm>
#include <cstdio>
#include <deque>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

[[noreturn]] void ThrowLastError(const char* where) {
  char buf[1024];
  mmc_last_error(buf, sizeof(buf));
  throw std::runtime_error(std::string(where) + ": " + buf);
}

// Owns every array mmc_world_desc points into.
struct FlatWorld {
  std::vector<int32_t> surface_type, cell_material, cell_surface_begin{0}, cell_surface_index, cell_surface_sense,
      cell_field_kind, material_nuclide_begin{0}, material_nuclide_index;
  std::vector<double> surface_param, cell_field_param, material_aden, material_nuclide_afrac;
  std::vector<uint32_t> mg_reaction_mask;
  std::vector<double> mg_total, mg_capture, mg_scatter, mg_fission, mg_nubar, mg_scatter_probs, mg_chi;
  // continuous energy: std::map contents as sorted (key, value) arrays
  std::deque<std::vector<double>> arrays;
  std::deque<std::vector<mmc_ce_reaction>> reactions;
  std::deque<std::vector<mmc_tsl_partition>> partitions;
  std::deque<mmc_tsl_desc> tsl;
  std::vector<mmc_ce_nuclide> nuclides;
  mmc_ce_desc ce{};
  mmc_world_desc desc{};

  mmc_table1d Table(const ContinuousMap<ContinuousEnergy, Real>& map) {
    std::vector<double> x, y;
    for (const auto& [key, value] : map.elements) {
      x.push_back(key);
      y.push_back(value);
    }
    arrays.push_back(std::move(x));
    const double* px = arrays.back().data();
    arrays.push_back(std::move(y));
    return mmc_table1d{arrays.back().size(), px, arrays.back().data()};
  }
};

template <typename Vec, typename Ptr> int32_t IndexOf(const Vec& vec, const Ptr& ptr) {
  return static_cast<int32_t>(std::find(vec.begin(), vec.end(), ptr) - vec.begin());
}

template <typename P> mmc_tsl_partition Partition(const P& p, const HDF5DataSet<3>& grid_T_modes) {
  mmc_tsl_partition d{};
  d.n_cdf = p.CDF_modes.GetAxis(0).size();
  d.n_grid = grid_T_modes.GetAxis(0).size();
  d.n_temperature = grid_T_modes.GetAxis(1).size();
  d.rank = p.singular_values.GetAxis(0).size();
  d.cdf = p.CDF_modes.GetAxis(0).data();
  d.grid = grid_T_modes.GetAxis(0).data();
  d.temperature = grid_T_modes.GetAxis(1).data();
  d.cdf_modes = p.CDF_modes.values.data();
  d.singular_values = p.singular_values.values.data();
  d.grid_T_modes = grid_T_modes.values.data();
  return d;
}

std::unique_ptr<FlatWorld> Flatten(const World& w) {
  auto flat = std::make_unique<FlatWorld>();
  FlatWorld& f = *flat;
  for (const auto& s : w.surfaces) {
    double prm[4] = {0, 0, 0, 0};
    if (const auto* sphere = dynamic_cast<const Sphere*>(s.get())) {
      f.surface_type.push_back(MMC_SURF_SPHERE);
      prm[0] = sphere->center.x, prm[1] = sphere->center.y, prm[2] = sphere->center.z, prm[3] = sphere->radius;
    } else if (const auto* plane = dynamic_cast<const PlaneX*>(s.get())) {
      f.surface_type.push_back(MMC_SURF_PLANEX);
      prm[0] = plane->c;
    } else if (const auto* cylinder = dynamic_cast<const CylinderX*>(s.get())) {
      f.surface_type.push_back(MMC_SURF_CYLINDERX);
      prm[0] = cylinder->radius;
    } else {
      throw std::runtime_error("surface \"" + s->name + "\": no device counterpart");
    }
    f.surface_param.insert(f.surface_param.end(), prm, prm + 4);
  }
  for (const auto& c : w.cells) {
    f.cell_material.push_back(c.material ? IndexOf(w.materials, c.material) : -1);
    // World creation order of the surfaces, the order this repo's host canonicalises the pointer-keyed map to (quirk
    // Q1): observable only through ties between equidistant surfaces
    std::vector<std::pair<int32_t, bool>> senses;
    for (const auto& [surface, is_within] : c.surface_senses) senses.emplace_back(IndexOf(w.surfaces, surface), is_within);
    std::sort(senses.begin(), senses.end());
    for (const auto& [index, is_within] : senses) {
      f.cell_surface_index.push_back(index);
      f.cell_surface_sense.push_back(is_within ? 1 : 0);
    }
    f.cell_surface_begin.push_back(static_cast<int32_t>(f.cell_surface_index.size()));
    const ScalarField* t = c.temperature.get();
    if (const auto* constant = dynamic_cast<const ConstantField*>(t)) {
      f.cell_field_kind.push_back(MMC_FIELD_CONSTANT);
      f.cell_field_param.insert(f.cell_field_param.end(), {constant->c, 0, 0, 0, t->upper_bound, t->lower_bound});
    } else if (const auto* linear = dynamic_cast<const LinearField*>(t)) {
      f.cell_field_kind.push_back(MMC_FIELD_LINEAR);
      f.cell_field_param.insert(f.cell_field_param.end(), {linear->g.x, linear->g.y, linear->g.z, linear->b, t->upper_bound, t->lower_bound});
    } else {
      throw std::runtime_error("cell \"" + c.name + "\": temperature field without a device counterpart");
    }
  }
  for (const auto& m : w.materials) {
    f.material_aden.push_back(m->number_density);
    std::vector<std::pair<int32_t, double>> afracs;  // normalised by Material (Material.cpp:83-92); creation order (Q1)
    for (const auto& [nuclide, afrac] : m->afracs) afracs.emplace_back(IndexOf(w.nuclides, nuclide), afrac);
    std::sort(afracs.begin(), afracs.end());
    for (const auto& [index, afrac] : afracs) {
      f.material_nuclide_index.push_back(index);
      f.material_nuclide_afrac.push_back(afrac);
    }
    f.material_nuclide_begin.push_back(static_cast<int32_t>(f.material_nuclide_index.size()));
  }
  int32_t n_groups = 0;
  for (const auto& nuclide : w.nuclides) {
    const Interaction* xs = nuclide->xs.at(Particle::Type::neutron).get();
    if (const auto* mg = dynamic_cast<const Multigroup*>(xs)) {
      const Group G = mg->max_group;
      n_groups = static_cast<int32_t>(G);
      uint32_t mask = 0;
      auto row = [&](Reaction reaction, uint32_t bit, std::vector<double>& out) {
        const auto it = mg->reactions.find(reaction);
        if (it != mg->reactions.end()) {
          mask |= bit;
          out.insert(out.end(), it->second.elements.begin(), it->second.elements.end());
        } else {
          out.insert(out.end(), G, 0.0);
        }
      };
      f.mg_total.insert(f.mg_total.end(), mg->total.elements.begin(), mg->total.elements.end());
      row(Reaction::capture, MMC_REACTION_CAPTURE, f.mg_capture);
      row(Reaction::scatter, MMC_REACTION_SCATTER, f.mg_scatter);
      row(Reaction::fission, MMC_REACTION_FISSION, f.mg_fission);
      f.mg_reaction_mask.push_back(mask);
      auto matrix = [G](const auto& optional, std::vector<double>& out) {  // [g_in][g_out]
        for (Group gi = 1; gi <= G; gi++)
          for (Group go = 1; go <= G; go++) out.push_back(optional.has_value() ? optional.value().at(gi).at(go) : 0.0);
      };
      matrix(mg->scatter_probs, f.mg_scatter_probs);
      matrix(mg->chi, f.mg_chi);
      for (Group g = 1; g <= G; g++) f.mg_nubar.push_back(mg->nubar.has_value() ? mg->nubar.value().at(g) : 0.0);
    } else if (const auto* ce = dynamic_cast<const Continuous*>(xs)) {
      f.reactions.emplace_back();
      double awr = 0;
      for (const auto& reaction : ce->reactions) {  // XML document order
        mmc_ce_reaction d{};
        d.xs = f.Table(reaction->evaluation.xs);
        d.temperature = reaction->evaluation.temperature;
        if (dynamic_cast<const ContinuousCapture*>(reaction.get())) {
          d.kind = MMC_REACTION_CAPTURE;
        } else if (const auto* scatter = dynamic_cast<const ContinuousScatter*>(reaction.get())) {
          d.kind = MMC_REACTION_SCATTER;
          awr = scatter->awr;
          if (scatter->tsl.has_value()) {
            const ThermalScattering& t = scatter->tsl.value();
            mmc_tsl_desc desc{};
            desc.majorant = f.Table(t.majorant);
            desc.n_energy = t.scatter_xs_E.GetAxis(0).size();
            desc.n_temperature = t.scatter_xs_T.GetAxis(0).size();
            desc.rank = t.scatter_xs_S.GetAxis(0).size();
            desc.energy = t.scatter_xs_E.GetAxis(0).data();
            desc.temperature = t.scatter_xs_T.GetAxis(0).data();
            desc.xs_E = t.scatter_xs_E.values.data();
            desc.xs_S = t.scatter_xs_S.values.data();
            desc.xs_T = t.scatter_xs_T.values.data();
            f.partitions.emplace_back();
            for (const auto& p : t.beta_partitions) f.partitions.back().push_back(Partition(p, p.E_T_modes));
            desc.n_beta_partitions = static_cast<int32_t>(t.beta_partitions.size());
            desc.beta_partitions = f.partitions.back().data();
            f.partitions.emplace_back();
            for (const auto& p : t.alpha_partitions) f.partitions.back().push_back(Partition(p, p.beta_T_modes));
            desc.n_alpha_partitions = static_cast<int32_t>(t.alpha_partitions.size());
            desc.alpha_partitions = f.partitions.back().data();
            desc.beta_cutoff = t.beta_cutoff;
            desc.alpha_cutoff = t.alpha_cutoff;
            desc.awr = t.awr;
            f.tsl.push_back(desc);
            d.tsl = &f.tsl.back();
          }
        } else if (const auto* fission = dynamic_cast<const ContinuousFission*>(reaction.get())) {
          d.kind = MMC_REACTION_FISSION;
          if (fission->nubar.has_value()) {
            d.has_nubar = 1;
            d.nubar = f.Table(fission->nubar.value());
          }
        }
        f.reactions.back().push_back(d);
      }
      mmc_ce_nuclide n{};
      n.awr = awr;  // <nuclide awr=...>: the reference hands it to ContinuousScatter only (free gas, thermal scattering)
      n.total = f.Table(ce->total.xs);
      n.total_temperature = ce->total.temperature;
      n.n_reactions = static_cast<int32_t>(f.reactions.back().size());
      n.reactions = f.reactions.back().data();
      f.nuclides.push_back(n);
    } else {
      throw std::runtime_error("nuclide \"" + nuclide->name + "\": no neutron interaction with a device counterpart");
    }
  }
  mmc_world_desc& d = f.desc;
  d.struct_size = sizeof(mmc_world_desc);
  d.abi_version = MMC_ABI_VERSION;
  d.n_surfaces = static_cast<int32_t>(w.surfaces.size());
  d.surface_type = f.surface_type.data();
  d.surface_param = f.surface_param.data();
  d.n_cells = static_cast<int32_t>(w.cells.size());
  d.cell_material = f.cell_material.data();
  d.cell_surface_begin = f.cell_surface_begin.data();
  d.cell_surface_index = f.cell_surface_index.data();
  d.cell_surface_sense = f.cell_surface_sense.data();
  d.cell_field_kind = f.cell_field_kind.data();
  d.cell_field_param = f.cell_field_param.data();
  d.n_materials = static_cast<int32_t>(w.materials.size());
  d.material_aden = f.material_aden.data();
  d.material_nuclide_begin = f.material_nuclide_begin.data();
  d.material_nuclide_index = f.material_nuclide_index.data();
  d.material_nuclide_afrac = f.material_nuclide_afrac.data();
  d.n_nuclides = static_cast<int32_t>(w.nuclides.size());
  d.n_groups = n_groups;
  if (n_groups > 0) {
    d.mg_reaction_mask = f.mg_reaction_mask.data();
    d.mg_total = f.mg_total.data();
    d.mg_capture = f.mg_capture.data();
    d.mg_scatter = f.mg_scatter.data();
    d.mg_fission = f.mg_fission.data();
    d.mg_nubar = f.mg_nubar.data();
    d.mg_scatter_probs = f.mg_scatter_probs.data();
    d.mg_chi = f.mg_chi.data();
  } else {
    f.ce.nuclides = f.nuclides.data();
    d.ce = &f.ce;
  }
  return flat;
}

mmc_source_desc FlattenSource(const Source& source) {
  mmc_source_desc d{};
  const auto* position = dynamic_cast<const ConstantDistribution<Point>*>(source.position.get());
  if (!position) throw std::runtime_error("source position: only <constant> has a device counterpart");
  d.position[0] = position->constant.x, d.position[1] = position->constant.y, d.position[2] = position->constant.z;
  if (const auto* constant = dynamic_cast<const ConstantDistribution<Direction>*>(source.direction.get())) {
    d.direction_kind = MMC_DIR_CONSTANT;
    d.direction[0] = constant->constant.x, d.direction[1] = constant->constant.y, d.direction[2] = constant->constant.z;
  } else if (dynamic_cast<const IsotropicDistribution*>(source.direction.get())) {
    d.direction_kind = MMC_DIR_ISOTROPIC;
  } else if (const auto* flux = dynamic_cast<const IsotropicFlux*>(source.direction.get())) {
    d.direction_kind = MMC_DIR_ISOTROPIC_FLUX;
    d.direction[0] = flux->reference.x, d.direction[1] = flux->reference.y, d.direction[2] = flux->reference.z;
  } else {
    throw std::runtime_error("source direction: no device counterpart");
  }
  const auto* energy = dynamic_cast<const ConstantDistribution<Energy>*>(source.energy.get());
  if (!energy) throw std::runtime_error("source energy: only <constant> has a device counterpart");
  d.group = 1;
  if (std::holds_alternative<Group>(energy->constant)) d.group = std::get<Group>(energy->constant);
  else d.energy = std::get<ContinuousEnergy>(energy->constant);
  return d;
}

mmc_bins_desc FlattenBins(const Bins* bins) {
  mmc_bins_desc d{};
  if (const auto* lin = dynamic_cast<const LinspaceBins*>(bins)) {
    d.kind = MMC_BINS_LINSPACE;
    d.n_bins = lin->n_bins, d.lower = lin->lower_bound, d.upper = lin->upper_bound, d.width = lin->bin_width;
  } else if (const auto* log = dynamic_cast<const LogspaceBins*>(bins)) {
    d.kind = MMC_BINS_LOGSPACE;
    d.n_bins = log->n_bins, d.base = log->base, d.lower = log->log_lower_bound, d.upper = log->log_upper_bound,
    d.width = log->log_bin_width;
  } else if (const auto* boundary = dynamic_cast<const BoundaryBins*>(bins)) {
    d.kind = MMC_BINS_BOUNDARIES;
    d.n_bins = boundary->boundaries.size() + 1;
    d.boundaries = boundary->boundaries.data();
  } else {
    d.kind = MMC_BINS_NONE;
    d.n_bins = 1;
  }
  return d;
}

}  // namespace

// FixedSource with its worker pool replaced by the GPU (src/FixedSource.cpp:22-77).
class GpuFixedSource : public FixedSource {
public:
  using FixedSource::FixedSource;

  EstimatorSet Solve() override {
    const auto flat = Flatten(world);
    mmc_world* handle = nullptr;
    if (mmc_world_create(&flat->desc, /*device*/ -1, &handle) != MMC_OK) ThrowLastError("mmc_world_create");
    const mmc_source_desc src = FlattenSource(source);
    std::vector<mmc_estimator_desc> estimators;
    size_t total_bins = 0;
    for (const auto& estimator : init_estimator_set.estimators) {
      const auto* current = dynamic_cast<const CurrentEstimator*>(estimator.get());
      if (!current) throw std::runtime_error("estimator \"" + estimator->name + "\": only <current> has a device counterpart");
      if (!estimator->sensitivities.empty()) throw std::runtime_error("this adapter leaves sensitivities to the reference");
      mmc_estimator_desc d{};
      d.surface = IndexOf(world.surfaces, current->surface);
      const ParticleBins& bins = *current->bins;
      d.has_cosine_direction = bins.direction.has_value() ? 1 : 0;
      if (bins.direction.has_value())
        d.cosine_direction[0] = bins.direction->x, d.cosine_direction[1] = bins.direction->y, d.cosine_direction[2] = bins.direction->z;
      d.cosine = FlattenBins(bins.cosine.get());
      d.energy = FlattenBins(bins.energy.get());
      estimators.push_back(d);
      total_bins += current->bins->size();
    }
    mmc_run_options options{};
    options.struct_size = sizeof(options);
    options.device = -1;
    options.tracking = dynamic_cast<const CellDeltaTracking*>(Particle::transport_method.get()) ? MMC_TRACK_CELL_DELTA : MMC_TRACK_SURFACE;
    options.secondary_capacity = 256;
    std::vector<double> scores(total_bins, 0.0), square_scores(total_bins, 0.0);
    mmc_counters counters{};
    const int status = mmc_fixed_source_run(handle, &src, estimators.data(), static_cast<int32_t>(estimators.size()), seed,
                                            /*first_history*/ 0, batchsize, &options, scores.data(), square_scores.data(), &counters);
    mmc_world_destroy(handle);
    if (status != MMC_OK) ThrowLastError("mmc_fixed_source_run");
    // solver_estimator_set += worker_estimator_set (FixedSource.cpp:31-33): the GPU is the one worker
    EstimatorSet result = init_estimator_set;
    size_t offset = 0;
    for (auto& estimator : result.estimators) {
      for (size_t i = 0; i < estimator->scores.size(); i++) {
        estimator->scores[i] += scores[offset + i];
        estimator->square_scores[i] += square_scores[offset + i];
      }
      offset += estimator->scores.size();
    }
    std::fprintf(stderr, "gpu_adapter: %llu histories, %llu events on the device\n",
                 static_cast<unsigned long long>(counters.n_histories), static_cast<unsigned long long>(counters.n_events));
    return result;
  }
};

// Driver::Create (src/Driver.cpp:19-35) with the one changed line: a fixed-source deck gets the GPU driver.
std::unique_ptr<Driver> CreateGpuDriver(const std::filesystem::path& xml_filepath) {
  auto doc = std::make_unique<XMLDocument>(xml_filepath);
  const std::string problem_type = doc->root.child("problemtype").first_child().name();
  if (problem_type == "fixedsource") return std::make_unique<GpuFixedSource>(doc->root);
  throw std::runtime_error("gpu_adapter: only <fixedsource> decks (the reference's KEigenvalue::Solve is a stub)");
}

int main(int argc, char** argv) {
  if (argc != 2) {
    std::fprintf(stderr, "usage: gpu_adapter <deck.xml>\n");
    return 2;
  }
  try {
    auto driver = CreateGpuDriver(argv[1]);
    const auto result = driver->Solve();
    // minimc.cpp:20-21
    std::cout << driver->batchsize << std::endl;
    std::cout << result.to_string();
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 3;
  }
  return 0;
}
