// World -> plain-old-data tables of the C ABI (include/minimc_b200.h).  Walks
// the host objects in their own iteration order; nothing here touches CUDA.
#include <cstdio>
#include <sstream>

#include "minimc.hpp"

namespace minimc {

struct FlatWorld::Arrays {
  std::vector<int32_t> surface_type;
  std::vector<double> surface_param;
  std::vector<int32_t> cell_material, cell_surface_begin, cell_surface_index, cell_surface_sense, cell_field_kind;
  std::vector<double> cell_field_param;
  std::vector<double> material_aden;
  std::vector<int32_t> material_nuclide_begin, material_nuclide_index;
  std::vector<double> material_nuclide_afrac;
  std::vector<uint32_t> mg_reaction_mask;
  std::vector<double> mg_total, mg_capture, mg_scatter, mg_fission, mg_nubar, mg_scatter_probs, mg_chi;
  // continuous energy: the descriptor structs point into the World's own tables
  std::vector<mmc_ce_nuclide> ce_nuclides;
  std::vector<std::vector<mmc_ce_reaction>> ce_reactions;
  std::vector<std::unique_ptr<mmc_tsl_desc>> tsl;
  std::vector<std::vector<mmc_tsl_partition>> tsl_partitions;
  mmc_ce_desc ce{};
};

namespace {

mmc_table1d Table(const PointwiseTable& t) { return mmc_table1d{t.x.size(), t.x.data(), t.y.data()}; }

mmc_tsl_partition Partition(const ThermalScattering::Partition& p) {
  mmc_tsl_partition d{};
  d.n_cdf = p.CDF_modes.GetAxis(0).size();
  d.n_grid = p.grid_T_modes.GetAxis(0).size();
  d.n_temperature = p.grid_T_modes.GetAxis(1).size();
  d.rank = p.singular_values.GetAxis(0).size();
  d.cdf = p.CDF_modes.GetAxis(0).data();
  d.grid = p.grid_T_modes.GetAxis(0).data();
  d.temperature = p.grid_T_modes.GetAxis(1).data();
  d.cdf_modes = p.CDF_modes.values.data();
  d.singular_values = p.singular_values.values.data();
  d.grid_T_modes = p.grid_T_modes.values.data();
  return d;
}

}  // namespace

FlatWorld::FlatWorld(const World& w) : a_{std::make_shared<Arrays>()} {
  Arrays& a = *a_;
  for (const CSGSurface& s : w.surfaces) {
    a.surface_type.push_back(s.type);
    a.surface_param.insert(a.surface_param.end(), s.param, s.param + 4);
  }
  a.cell_surface_begin.push_back(0);
  for (const Cell& c : w.cells) {
    a.cell_material.push_back(c.material);
    for (const auto& [index, is_within] : c.surface_senses) {
      a.cell_surface_index.push_back(static_cast<int32_t>(index));
      a.cell_surface_sense.push_back(is_within ? 1 : 0);
    }
    a.cell_surface_begin.push_back(static_cast<int32_t>(a.cell_surface_index.size()));
    const ScalarField& t = c.temperature;
    a.cell_field_kind.push_back(t.kind);
    if (t.kind == MMC_FIELD_CONSTANT) a.cell_field_param.insert(a.cell_field_param.end(), {t.c, 0, 0, 0, t.upper_bound, t.lower_bound});
    else a.cell_field_param.insert(a.cell_field_param.end(), {t.g.x, t.g.y, t.g.z, t.b, t.upper_bound, t.lower_bound});
  }
  a.material_nuclide_begin.push_back(0);
  for (const Material& m : w.materials) {
    a.material_aden.push_back(m.number_density);
    for (const auto& [index, afrac] : m.afracs) {
      a.material_nuclide_index.push_back(static_cast<int32_t>(index));
      a.material_nuclide_afrac.push_back(afrac);
    }
    a.material_nuclide_begin.push_back(static_cast<int32_t>(a.material_nuclide_index.size()));
  }
  const bool multigroup = w.multigroup_groups > 0;
  for (const Nuclide& n : w.nuclides) {
    if (multigroup) {
      const Multigroup& mg = n.multigroup.value();
      a.mg_reaction_mask.push_back(mg.reaction_mask);
      a.mg_total.insert(a.mg_total.end(), mg.total.begin(), mg.total.end());
      a.mg_capture.insert(a.mg_capture.end(), mg.capture.begin(), mg.capture.end());
      a.mg_scatter.insert(a.mg_scatter.end(), mg.scatter.begin(), mg.scatter.end());
      a.mg_fission.insert(a.mg_fission.end(), mg.fission.begin(), mg.fission.end());
      a.mg_nubar.insert(a.mg_nubar.end(), mg.nubar.begin(), mg.nubar.end());
      a.mg_scatter_probs.insert(a.mg_scatter_probs.end(), mg.scatter_probs.begin(), mg.scatter_probs.end());
      a.mg_chi.insert(a.mg_chi.end(), mg.chi.begin(), mg.chi.end());
    } else {
      const Continuous& ce = n.continuous.value();
      a.ce_reactions.emplace_back();
      for (const ContinuousReaction& r : ce.reactions) {
        mmc_ce_reaction d{};
        d.kind = static_cast<int32_t>(r.kind);
        d.xs = Table(r.xs);
        d.temperature = r.temperature;
        if (r.tsl) {
          const ThermalScattering& t = *r.tsl;
          auto desc = std::make_unique<mmc_tsl_desc>();
          desc->majorant = Table(t.majorant);
          desc->n_energy = t.scatter_xs_E.GetAxis(0).size();
          desc->n_temperature = t.scatter_xs_T.GetAxis(0).size();
          desc->rank = t.scatter_xs_S.GetAxis(0).size();
          desc->energy = t.scatter_xs_E.GetAxis(0).data();
          desc->temperature = t.scatter_xs_T.GetAxis(0).data();
          desc->xs_E = t.scatter_xs_E.values.data();
          desc->xs_S = t.scatter_xs_S.values.data();
          desc->xs_T = t.scatter_xs_T.values.data();
          a.tsl_partitions.emplace_back();
          for (const auto& p : t.beta_partitions) a.tsl_partitions.back().push_back(Partition(p));
          desc->n_beta_partitions = static_cast<int32_t>(t.beta_partitions.size());
          desc->beta_partitions = a.tsl_partitions.back().data();
          a.tsl_partitions.emplace_back();
          for (const auto& p : t.alpha_partitions) a.tsl_partitions.back().push_back(Partition(p));
          desc->n_alpha_partitions = static_cast<int32_t>(t.alpha_partitions.size());
          desc->alpha_partitions = a.tsl_partitions.back().data();
          desc->beta_cutoff = t.beta_cutoff;
          desc->alpha_cutoff = t.alpha_cutoff;
          desc->awr = t.awr;
          d.tsl = desc.get();
          a.tsl.push_back(std::move(desc));
        }
        if (r.nubar) {
          d.has_nubar = 1;
          d.nubar = Table(*r.nubar);
        }
        a.ce_reactions.back().push_back(d);
      }
    }
  }
  if (!multigroup) {
    for (size_t i = 0; i < w.nuclides.size(); i++) {
      const Continuous& ce = w.nuclides[i].continuous.value();
      mmc_ce_nuclide d{};
      d.awr = ce.awr;
      d.total = Table(ce.total);
      d.total_temperature = ce.total_temperature;
      d.n_reactions = static_cast<int32_t>(a.ce_reactions[i].size());
      d.reactions = a.ce_reactions[i].data();
      a.ce_nuclides.push_back(d);
    }
    a.ce.nuclides = a.ce_nuclides.data();
  }

  mmc_world_desc& d = desc_;
  d.struct_size = sizeof(mmc_world_desc);
  d.abi_version = MMC_ABI_VERSION;
  d.n_surfaces = static_cast<int32_t>(w.surfaces.size());
  d.surface_type = a.surface_type.data();
  d.surface_param = a.surface_param.data();
  d.n_cells = static_cast<int32_t>(w.cells.size());
  d.cell_material = a.cell_material.data();
  d.cell_surface_begin = a.cell_surface_begin.data();
  d.cell_surface_index = a.cell_surface_index.data();
  d.cell_surface_sense = a.cell_surface_sense.data();
  d.cell_field_kind = a.cell_field_kind.data();
  d.cell_field_param = a.cell_field_param.data();
  d.n_materials = static_cast<int32_t>(w.materials.size());
  d.material_aden = a.material_aden.data();
  d.material_nuclide_begin = a.material_nuclide_begin.data();
  d.material_nuclide_index = a.material_nuclide_index.data();
  d.material_nuclide_afrac = a.material_nuclide_afrac.data();
  d.n_nuclides = static_cast<int32_t>(w.nuclides.size());
  d.n_groups = static_cast<int32_t>(w.multigroup_groups);
  if (multigroup) {
    d.mg_reaction_mask = a.mg_reaction_mask.data();
    d.mg_total = a.mg_total.data();
    d.mg_capture = a.mg_capture.data();
    d.mg_scatter = a.mg_scatter.data();
    d.mg_fission = a.mg_fission.data();
    d.mg_nubar = a.mg_nubar.data();
    d.mg_scatter_probs = a.mg_scatter_probs.data();
    d.mg_chi = a.mg_chi.data();
  } else {
    d.ce = &a.ce;
  }
}

namespace {
std::string Hex(double v) {
  char buf[64];
  std::snprintf(buf, sizeof(buf), "\"%a\"", v);
  return buf;
}
template <typename T> std::string List(const std::vector<T>& v) {
  std::ostringstream s;
  s << "[";
  for (size_t i = 0; i < v.size(); i++) {
    if (i) s << ", ";
    if constexpr (std::is_floating_point_v<T>) s << Hex(v[i]);
    else s << v[i];
  }
  s << "]";
  return s.str();
}
}  // namespace

std::string FlatWorld::to_json() const {
  const Arrays& a = *a_;
  std::ostringstream s;
  s << "{\n";
  s << "\"n_groups\": " << desc_.n_groups << ",\n";
  s << "\"surface_type\": " << List(a.surface_type) << ",\n";
  s << "\"surface_param\": " << List(a.surface_param) << ",\n";
  s << "\"cell_material\": " << List(a.cell_material) << ",\n";
  s << "\"cell_surface_begin\": " << List(a.cell_surface_begin) << ",\n";
  s << "\"cell_surface_index\": " << List(a.cell_surface_index) << ",\n";
  s << "\"cell_surface_sense\": " << List(a.cell_surface_sense) << ",\n";
  s << "\"cell_field_kind\": " << List(a.cell_field_kind) << ",\n";
  s << "\"cell_field_param\": " << List(a.cell_field_param) << ",\n";
  s << "\"material_aden\": " << List(a.material_aden) << ",\n";
  s << "\"material_nuclide_begin\": " << List(a.material_nuclide_begin) << ",\n";
  s << "\"material_nuclide_index\": " << List(a.material_nuclide_index) << ",\n";
  s << "\"material_nuclide_afrac\": " << List(a.material_nuclide_afrac) << ",\n";
  s << "\"mg_reaction_mask\": " << List(a.mg_reaction_mask) << ",\n";
  s << "\"mg_total\": " << List(a.mg_total) << ",\n";
  s << "\"mg_capture\": " << List(a.mg_capture) << ",\n";
  s << "\"mg_scatter\": " << List(a.mg_scatter) << ",\n";
  s << "\"mg_fission\": " << List(a.mg_fission) << ",\n";
  s << "\"mg_nubar\": " << List(a.mg_nubar) << ",\n";
  s << "\"mg_scatter_probs\": " << List(a.mg_scatter_probs) << ",\n";
  s << "\"mg_chi\": " << List(a.mg_chi) << ",\n";
  s << "\"ce_nuclides\": " << a.ce_nuclides.size() << "\n";
  s << "}\n";
  return s.str();
}

std::vector<mmc_estimator_desc> FlattenEstimators(const EstimatorSet& set) {
  auto bins = [](const Bins& b) {
    mmc_bins_desc d{};
    d.kind = b.kind;
    d.n_bins = b.n_bins;
    d.lower = b.lower;
    d.upper = b.upper;
    d.width = b.width;
    d.base = b.base;
    d.boundaries = b.boundaries.empty() ? nullptr : b.boundaries.data();
    return d;
  };
  std::vector<mmc_estimator_desc> result;
  for (const Estimator& e : set.estimators) {
    mmc_estimator_desc d{};
    d.surface = static_cast<int32_t>(e.surface);
    d.has_cosine_direction = e.bins.direction ? 1 : 0;
    if (e.bins.direction) {
      d.cosine_direction[0] = e.bins.direction->x;
      d.cosine_direction[1] = e.bins.direction->y;
      d.cosine_direction[2] = e.bins.direction->z;
    }
    d.cosine = bins(e.bins.cosine);
    d.energy = bins(e.bins.energy);
    result.push_back(d);
  }
  return result;
}

}  // namespace minimc
