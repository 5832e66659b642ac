// Object model of the host: construction from the XML deck.  No transport
// physics lives here (see minimc.hpp).
#include "minimc.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <numeric>
#include <sstream>
#include <stdexcept>

namespace minimc {

namespace {

// `while (stream >> value)` of the reference's list parsing (Multigroup.cpp:82-86, Bins.cpp:136-150)
std::vector<Real> ParseRealList(const std::string& text) {
  std::vector<Real> result;
  std::stringstream list{text};
  Real element;
  while (list >> element) result.push_back(element);
  return result;
}

template <typename Range, typename GetName> std::string QuotedNames(const Range& range, GetName get_name) {
  std::string result;
  for (const auto& item : range) result += "\"" + get_name(item) + "\", ";
  return result;
}

uint64_t GroupStructureSize(const xml::Node& any_node) {
  // Multigroup::GroupStructureSize, Multigroup.cpp:168-175
  const xml::Node& root = any_node.root();
  const xml::Node* nuclides = root.child("nuclides");
  const xml::Node* multigroup = nuclides ? nuclides->child("multigroup") : nullptr;
  return multigroup ? multigroup->attribute_ull("groups") : 0;
}

const xml::Node& Require(const xml::Node* node, const xml::Node& parent, const char* name) {
  if (!node) throw std::runtime_error(parent.path() + ": \"" + name + "\" node not found");
  return *node;
}

}  // namespace

// ---------------------------------------------------------------- table files
TableFile::TableFile(const std::string& path, size_t expected_dimensions) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("File not found: " + path);
  char magic[8];
  f.read(magic, 8);
  if (!f || std::memcmp(magic, "MMCTAB1\0", 8) != 0)
    throw std::runtime_error(
        path + ": not an MMCTAB1 table file (pandas-HDF5 input must be converted first; see DESIGN.md)");
  uint64_t ndim = 0;
  f.read(reinterpret_cast<char*>(&ndim), 8);
  if (!f || ndim == 0 || ndim > 8) throw std::runtime_error(path + ": corrupt table header");
  if (ndim != expected_dimensions)
    throw std::runtime_error(
        path + ": Expected " + std::to_string(expected_dimensions) + " dimensions, but got " + std::to_string(ndim));
  std::vector<uint64_t> shape(ndim);
  f.read(reinterpret_cast<char*>(shape.data()), static_cast<std::streamsize>(8 * ndim));
  uint64_t total = 1;
  for (const uint64_t n : shape) {
    if (n == 0 || n > (uint64_t{1} << 32)) throw std::runtime_error(path + ": corrupt table shape");
    axes.emplace_back(n);
    f.read(reinterpret_cast<char*>(axes.back().data()), static_cast<std::streamsize>(8 * n));
    total *= n;
  }
  values.resize(total);
  f.read(reinterpret_cast<char*>(values.data()), static_cast<std::streamsize>(8 * total));
  if (!f) throw std::runtime_error(path + ": truncated table file");
}

PointwiseTable PointwiseTable::FromFile(const std::string& path) {
  // HDF5DataSet<1>::ToContinuousMap inserts (axis[i], values[i]) into a
  // std::map with operator[]: sorted by key, the last duplicate wins.
  const TableFile file{path, 1};
  std::map<double, double> elements;
  const auto& keys = file.GetAxis(0);
  for (size_t i = 0; i < keys.size(); i++) elements[keys[i]] = file.values[i];
  PointwiseTable t;
  for (const auto& [k, v] : elements) {
    t.x.push_back(k);
    t.y.push_back(v);
  }
  return t;
}

// ------------------------------------------------------------------- geometry
CSGSurface CSGSurface::Create(const xml::Node& root, const std::string& name) {
  const xml::Node* surfaces_node = root.child("surfaces");
  const xml::Node* surface_node = surfaces_node ? surfaces_node->child_by_attribute("name", name) : nullptr;
  if (!surface_node) {
    std::string names;
    if (surfaces_node) names = QuotedNames(surfaces_node->children(), [](const auto& n) { return n->attribute("name"); });
    throw std::runtime_error("Surface node \"" + name + "\" not found. Must be one of: [" + names + "]");
  }
  CSGSurface s;
  s.name = surface_node->attribute("name");
  const std::string& type = surface_node->name();
  if (type == "sphere") {
    // Sphere::Sphere, CSGSurface.cpp:101-105
    const xml::Node& center = Require(surface_node->child("center"), *surface_node, "center");
    const xml::Node& radius = Require(surface_node->child("radius"), *surface_node, "radius");
    s.type = MMC_SURF_SPHERE;
    s.param[0] = center.attribute_double("x");
    s.param[1] = center.attribute_double("y");
    s.param[2] = center.attribute_double("z");
    s.param[3] = radius.attribute_double("r");
  } else if (type == "planex") {
    s.type = MMC_SURF_PLANEX;  // CSGSurface.cpp:120-123
    s.param[0] = surface_node->attribute_double("x");
  } else if (type == "cylinderx") {
    s.type = MMC_SURF_CYLINDERX;  // CSGSurface.cpp:142-146
    s.param[0] = surface_node->attribute_double("r");
  } else {
    throw std::runtime_error(surface_node->path() + ": unknown surface type");
  }
  return s;
}

ScalarField ScalarField::Create(const xml::Node& node) {
  ScalarField f;
  if (node.name() == "constant") {
    f = Constant(node.attribute_double("c"));
  } else if (node.name() == "linear") {
    const xml::Node& bounds = Require(node.child("bounds"), node, "bounds");
    const xml::Node& gradient = Require(node.child("gradient"), node, "gradient");
    const xml::Node& intercept = Require(node.child("intercept"), node, "intercept");
    f.kind = MMC_FIELD_LINEAR;
    f.upper_bound = bounds.attribute_double("upper");
    f.lower_bound = bounds.attribute_double("lower");
    f.g = Point{gradient.attribute_double("x"), gradient.attribute_double("y"), gradient.attribute_double("z")};
    f.b = intercept.attribute_double("b");
  } else {
    throw std::runtime_error(node.path() + ": unknown scalar field type");
  }
  return f;
}

ScalarField ScalarField::Constant(Real c) {
  ScalarField f;
  f.kind = MMC_FIELD_CONSTANT;
  f.c = c;
  f.upper_bound = f.lower_bound = c;
  return f;
}

bool ScalarField::IsConstant() const noexcept {
  // ConstantField::IsConstant / LinearField::IsConstant (ScalarField.cpp:43,54-56)
  return kind == MMC_FIELD_CONSTANT || (g.x == 0 && g.y == 0 && g.z == 0);
}

// ----------------------------------------------------------------- multigroup
namespace {

std::vector<Real> OneDimensional(const xml::Node& groupxs_node) {
  const uint64_t G = GroupStructureSize(groupxs_node);
  std::vector<Real> elements = ParseRealList(groupxs_node.text());
  if (elements.size() != G)
    throw std::runtime_error(
        groupxs_node.path() + ": Expected " + std::to_string(G) + " entries but got " + std::to_string(elements.size()));
  return elements;
}

// [incoming gp][outgoing g] = flattened[G*(g-1) + (gp-1)] (Multigroup.cpp:124-147)
std::vector<Real> TwoDimensional(const xml::Node& groupxs_node) {
  const uint64_t G = GroupStructureSize(groupxs_node);
  const std::vector<Real> flattened = ParseRealList(groupxs_node.text());
  if (flattened.size() != G * G)
    throw std::runtime_error(
        groupxs_node.path() + ": Expected " + std::to_string(G * G) + " entries but got " +
        std::to_string(flattened.size()));
  std::vector<Real> columns(G * G);
  for (uint64_t gp = 0; gp < G; gp++)
    for (uint64_t g = 0; g < G; g++) columns[gp * G + g] = flattened[G * g + gp];
  return columns;
}

std::vector<Real> NormalizedTwoDimensional(const xml::Node& groupxs_node) {
  // Multigroup.cpp:168-186: columns whose sum is exactly 0 are left alone
  const uint64_t G = GroupStructureSize(groupxs_node);
  std::vector<Real> columns = TwoDimensional(groupxs_node);
  for (uint64_t gp = 0; gp < G; gp++) {
    Real column_sum = 0;
    for (uint64_t g = 0; g < G; g++) column_sum = column_sum + columns[gp * G + g];
    if (column_sum == 0.) continue;
    for (uint64_t g = 0; g < G; g++) columns[gp * G + g] /= column_sum;
  }
  return columns;
}

uint32_t ReactionBit(const xml::Node& reaction_node) {
  // ToReaction, Reaction.cpp
  const std::string& name = reaction_node.name();
  if (name == "capture") return MMC_REACTION_CAPTURE;
  if (name == "scatter") return MMC_REACTION_SCATTER;
  if (name == "fission") return MMC_REACTION_FISSION;
  throw std::runtime_error(reaction_node.path() + ": Unrecognized reaction name: " + name);
}

}  // namespace

Multigroup::Multigroup(const xml::Node& particle_node) : max_group{GroupStructureSize(particle_node)} {
  const uint64_t G = max_group;
  capture.assign(G, 0);
  scatter.assign(G, 0);
  fission.assign(G, 0);
  nubar.assign(G, 0);
  scatter_probs.assign(G * G, 0);
  chi.assign(G * G, 0);
  const xml::Node* fission_node = particle_node.child("fission");
  if (fission_node && fission_node->child("nubar")) nubar = OneDimensional(*fission_node->child("nubar"));
  if (fission_node && fission_node->child("chi")) chi = NormalizedTwoDimensional(*fission_node->child("chi"));
  if (const xml::Node* scatter_node = particle_node.child("scatter"))
    scatter_probs = NormalizedTwoDimensional(*scatter_node);
  // CreateReactions (Multigroup.cpp:206-222): std::map::emplace keeps the first node of each kind
  for (const auto& reaction_node : particle_node.children()) {
    const uint32_t bit = ReactionBit(*reaction_node);
    if (reaction_mask & bit) continue;
    reaction_mask |= bit;
    if (bit == MMC_REACTION_SCATTER) {
      // CreateScatterXS: column sums of the raw matrix
      const std::vector<Real> columns = TwoDimensional(*reaction_node);
      for (uint64_t gp = 0; gp < G; gp++) {
        Real sum = 0;
        for (uint64_t g = 0; g < G; g++) sum = sum + columns[gp * G + g];
        scatter[gp] = sum;
      }
    } else if (bit == MMC_REACTION_FISSION) {
      fission = OneDimensional(Require(reaction_node->child("xs"), *reaction_node, "xs"));
    } else {
      capture = OneDimensional(*reaction_node);
    }
  }
  // CreateTotalXS (Multigroup.cpp:224-237): accumulate from 0 over the map in enum order
  total.assign(G, 0);
  for (const uint32_t bit : {MMC_REACTION_CAPTURE, MMC_REACTION_SCATTER, MMC_REACTION_FISSION}) {
    if (!(reaction_mask & bit)) continue;
    const std::vector<Real>& xs = bit == MMC_REACTION_CAPTURE ? capture : bit == MMC_REACTION_SCATTER ? scatter : fission;
    for (uint64_t g = 0; g < G; g++) total[g] = total[g] + xs[g];
  }
}

// ----------------------------------------------------------- continuous energy
ThermalScattering::Partition::Partition(const xml::Node& partition_node, const char* modes_attribute)
    : CDF_modes{partition_node.attribute("CDF"), 2}, singular_values{partition_node.attribute("S"), 1},
      grid_T_modes{partition_node.attribute(modes_attribute), 3} {
  const size_t rank = singular_values.GetAxis(0).size();
  if (CDF_modes.GetAxis(1).size() != rank || grid_T_modes.GetAxis(2).size() != rank)
    throw std::runtime_error(partition_node.path() + ": CDF, S and " + modes_attribute + " tables disagree on the POD rank");
}

namespace {
std::vector<ThermalScattering::Partition> ReadPartitions(const xml::Node* list_node, const char* modes_attribute) {
  std::vector<ThermalScattering::Partition> result;
  if (list_node)
    for (const auto& partition_node : list_node->children()) result.emplace_back(*partition_node, modes_attribute);
  return result;
}
}  // namespace

ThermalScattering::ThermalScattering(const xml::Node& tsl_node)
    : majorant{PointwiseTable::FromFile(tsl_node.attribute("majorant"))},
      scatter_xs_T{tsl_node.attribute("total_T"), 2}, scatter_xs_S{tsl_node.attribute("total_S"), 1},
      scatter_xs_E{tsl_node.attribute("total_E"), 2},
      beta_partitions{ReadPartitions(tsl_node.child("beta_partitions"), "E_T")},
      alpha_partitions{ReadPartitions(tsl_node.child("alpha_partitions"), "beta_T")},
      beta_cutoff{tsl_node.attribute_double("beta_cutoff")}, alpha_cutoff{tsl_node.attribute_double("alpha_cutoff")},
      // tsl -> scatter -> neutron -> nuclide (ThermalScattering.cpp:104)
      awr{tsl_node.parent()->parent()->parent()->attribute_double("awr")} {
  // the reference asserts these orderings (ThermalScattering.cpp:46-51,79-84)
  Real previous = 0;
  for (const auto& partition : beta_partitions)
    for (const Real E : partition.grid_T_modes.GetAxis(0)) {
      if (!(previous < E)) throw std::runtime_error(tsl_node.path() + ": beta partition incident energies must be increasing");
      previous = E;
    }
  previous = 0;
  for (const auto& partition : alpha_partitions)
    for (const Real beta : partition.grid_T_modes.GetAxis(0)) {
      if (!(previous < beta)) throw std::runtime_error(tsl_node.path() + ": alpha partition betas must be increasing");
      previous = beta;
    }
  const size_t rank = scatter_xs_S.GetAxis(0).size();
  if (scatter_xs_E.GetAxis(1).size() != rank || scatter_xs_T.GetAxis(1).size() != rank)
    throw std::runtime_error(tsl_node.path() + ": total_E, total_S and total_T tables disagree on the POD rank");
}

ContinuousReaction ContinuousReaction::Create(const xml::Node& reaction_node) {
  ContinuousReaction r;
  r.kind = ReactionBit(reaction_node);
  // capture: the node itself is the evaluation; scatter / fission: its <xs> child
  // (ContinuousReaction.cpp:62-63,76-79,244-251)
  const xml::Node& evaluation_node =
      r.kind == MMC_REACTION_CAPTURE ? reaction_node : Require(reaction_node.child("xs"), reaction_node, "xs");
  r.xs = PointwiseTable::FromFile(evaluation_node.attribute("file"));
  r.temperature = evaluation_node.attribute_double("temperature");
  if (r.kind == MMC_REACTION_SCATTER) {
    if (const xml::Node* tsl_node = reaction_node.child("tsl")) {
      // ContinuousScatter::ReadPandasSAB, ContinuousReaction.cpp:193-205
      if (tsl_node->parent()->parent()->name() != "neutron")
        throw std::runtime_error(tsl_node->path() + ": Only neutrons may have a thermal scattering library node");
      r.tsl.emplace(*tsl_node);
    }
  } else if (r.kind == MMC_REACTION_FISSION) {
    if (const xml::Node* nubar_node = reaction_node.child("nubar"))
      r.nubar = PointwiseTable::FromFile(nubar_node->attribute("file"));
  }
  return r;
}

Continuous::Continuous(const xml::Node& particle_node) {
  // CreateReactions (Continuous.cpp:74-85): document order, <total> skipped
  for (const auto& reaction_node : particle_node.children()) {
    if (reaction_node->name() == "total") continue;
    reactions.push_back(ContinuousReaction::Create(*reaction_node));
  }
  const xml::Node& total_node = Require(particle_node.child("total"), particle_node, "total");
  total = PointwiseTable::FromFile(total_node.attribute("file"));
  total_temperature = total_node.attribute_double("temperature");
  awr = particle_node.parent()->attribute_double("awr");
}

Nuclide::Nuclide(const xml::Node& nuclide_node) : name{nuclide_node.attribute("name")} {
  // Interaction::Create, Interaction.cpp:16-45
  const xml::Node& root = nuclide_node.root();
  const xml::Node* general = root.child("general");
  const xml::Node* particles = general ? general->child("particles") : nullptr;
  std::stringstream particle_name_list{particles ? particles->text() : ""};
  std::string particle_name;
  while (particle_name_list >> particle_name) {
    const xml::Node* particle_node = nuclide_node.child(particle_name);
    if (!particle_node) throw std::runtime_error(nuclide_node.path() + ": \"" + particle_name + "\" node not found");
    if (particle_name != "neutron")
      throw std::runtime_error(nuclide_node.path() + ": only neutron transport is implemented on the GPU path");
    const std::string& energy_type = nuclide_node.parent()->name();
    if (energy_type == "multigroup") multigroup.emplace(*particle_node);
    else if (energy_type == "continuous") continuous.emplace(*particle_node);
    else throw std::runtime_error(nuclide_node.path() + ": unknown energy type " + energy_type);
  }
}

// ------------------------------------------------------------------- material
const xml::Node& Material::FindNode(const xml::Node& root, const std::string& material_name) {
  const xml::Node* materials_node = root.child("materials");
  const xml::Node* material_node = materials_node ? materials_node->child_by_attribute("name", material_name) : nullptr;
  if (!material_node) {
    std::string names;
    if (materials_node) names = QuotedNames(materials_node->children(), [](const auto& n) { return n->attribute("name"); });
    throw std::runtime_error("Material node \"" + material_name + "\" not found. Must be one of: [" + names + "]");
  }
  return *material_node;
}

Material::Material(const xml::Node& root, const std::string& name, const std::vector<Nuclide>& all_nuclides)
    : name{name}, number_density{FindNode(root, name).attribute_double("aden")} {
  // AssignNuclides (Material.cpp:66-93).  std::map::emplace keeps the first
  // entry of a repeated nuclide; iteration is in World creation order here.
  const xml::Node& material_node = FindNode(root, name);
  for (const auto& nuclide_node : material_node.children()) {
    const std::string nuclide_name = nuclide_node->attribute("name");
    const auto it = std::find_if(
        all_nuclides.cbegin(), all_nuclides.cend(), [&](const Nuclide& n) { return n.name == nuclide_name; });
    if (it == all_nuclides.cend()) throw std::runtime_error("Nuclide \"" + nuclide_name + "\" not found");
    const size_t index = static_cast<size_t>(it - all_nuclides.cbegin());
    if (std::none_of(afracs.cbegin(), afracs.cend(), [&](const auto& e) { return e.first == index; }))
      afracs.emplace_back(index, nuclide_node->attribute_double("afrac"));
  }
  std::sort(afracs.begin(), afracs.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
  Real sum_afrac = 0;
  for (const auto& e : afracs) sum_afrac = sum_afrac + e.second;
  for (auto& e : afracs) e.second = e.second / sum_afrac;
}

// ----------------------------------------------------------------------- cell
Cell::Cell(
    const xml::Node& cell_node, const std::vector<CSGSurface>& all_surfaces, const std::vector<Material>& all_materials,
    const ScalarField& global_temperature)
    : name{cell_node.attribute("name")},
      // AssignTemperature, Cell.cpp:103-113
      temperature{
          cell_node.has_attribute("temperature") ? ScalarField::Constant(cell_node.attribute_double("temperature"))
                                                 : global_temperature} {
  // AssignSurfaceSenses, Cell.cpp:57-85
  for (const auto& surface_node : cell_node.children()) {
    const std::string surface_name = surface_node->attribute("name");
    const auto it = std::find_if(
        all_surfaces.cbegin(), all_surfaces.cend(), [&](const CSGSurface& s) { return s.name == surface_name; });
    if (it == all_surfaces.cend()) throw std::runtime_error(surface_node->path() + ": unknown surface " + surface_name);
    const std::string sense = surface_node->attribute("sense");
    bool is_within;
    if (sense == "-1") is_within = true;
    else if (sense == "+1") is_within = false;
    else throw std::runtime_error(surface_node->path() + ": sense must be -1 or +1");
    const size_t index = static_cast<size_t>(it - all_surfaces.cbegin());
    if (std::none_of(surface_senses.cbegin(), surface_senses.cend(), [&](const auto& e) { return e.first == index; }))
      surface_senses.emplace_back(index, is_within);
  }
  std::sort(surface_senses.begin(), surface_senses.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
  // AssignMaterial, Cell.cpp:87-101: a node without a name is a <void>
  if (cell_node.has_attribute("name")) {
    const std::string material_name = cell_node.attribute("material");
    const auto it = std::find_if(
        all_materials.cbegin(), all_materials.cend(), [&](const Material& m) { return m.name == material_name; });
    if (it == all_materials.cend()) throw std::runtime_error(cell_node.path() + ": unknown material " + material_name);
    material = static_cast<int>(it - all_materials.cbegin());
  }
}

// ---------------------------------------------------------------------- world
World::World(const xml::Node& root) : temperature{ScalarField::Constant(constants::room_temperature)} {
  const xml::Node& cells_node = Require(root.child("cells"), root, "cells");
  multigroup_groups = GroupStructureSize(root);
  // CreateCSGSurfaces, World.cpp:89-107
  for (const auto& cell_node : cells_node.children())
    for (const auto& surface_node : cell_node->children()) {
      const std::string surface_name = surface_node->attribute("name");
      if (std::none_of(surfaces.cbegin(), surfaces.cend(), [&](const CSGSurface& s) { return s.name == surface_name; }))
        surfaces.push_back(CSGSurface::Create(root, surface_name));
    }
  // CreateNuclides, World.cpp:109-149
  for (const auto& cell_node : cells_node.children()) {
    if (cell_node->name() == "void") continue;
    const std::string cell_material_name = cell_node->attribute("material");
    const xml::Node& material_node = Material::FindNode(root, cell_material_name);
    for (const auto& material_nuclide_node : material_node.children()) {
      const std::string nuclide_name = material_nuclide_node->attribute("name");
      if (std::any_of(nuclides.cbegin(), nuclides.cend(), [&](const Nuclide& n) { return n.name == nuclide_name; }))
        continue;
      const xml::Node& nuclides_node = Require(root.child("nuclides"), root, "nuclides");
      const xml::Node* energy_type_node = nuclides_node.first_child();
      const xml::Node* nuclide_node = energy_type_node ? energy_type_node->child_by_attribute("name", nuclide_name) : nullptr;
      if (!nuclide_node) {
        std::string names;
        if (energy_type_node)
          names = QuotedNames(energy_type_node->children(), [](const auto& n) { return n->attribute("name"); });
        throw std::runtime_error("Nuclide node \"" + nuclide_name + "\" not found. Must be one of: [" + names + "]");
      }
      nuclides.emplace_back(*nuclide_node);
    }
  }
  // CreateMaterials, World.cpp:151-171
  for (const auto& cell_node : cells_node.children()) {
    if (cell_node->name() == "void") continue;
    const std::string material_name = cell_node->attribute("material");
    if (std::none_of(materials.cbegin(), materials.cend(), [&](const Material& m) { return m.name == material_name; }))
      materials.emplace_back(root, material_name, nuclides);
  }
  // CreateTemperature, World.cpp:173-179
  if (const xml::Node* temperature_node = root.child("temperature")) {
    if (!temperature_node->first_child()) throw std::runtime_error(temperature_node->path() + ": empty temperature node");
    temperature = ScalarField::Create(*temperature_node->first_child());
  }
  // CreateCells, World.cpp:181-191
  for (const auto& cell_node : cells_node.children()) cells.emplace_back(*cell_node, surfaces, materials, temperature);
}

size_t World::FindSurfaceIndexByName(const std::string& name) const {
  const auto it = std::find_if(surfaces.cbegin(), surfaces.cend(), [&](const CSGSurface& s) { return s.name == name; });
  if (it == surfaces.cend())
    throw std::runtime_error(
        "Surface \"" + name + "\" not found. Must be one of: [" +
        QuotedNames(surfaces, [](const CSGSurface& s) { return s.name; }) + "]");
  return static_cast<size_t>(it - surfaces.cbegin());
}

size_t World::FindNuclideIndexByName(const std::string& name) const {
  const auto it = std::find_if(nuclides.cbegin(), nuclides.cend(), [&](const Nuclide& n) { return n.name == name; });
  if (it == nuclides.cend())
    throw std::runtime_error(
        "Nuclide \"" + name + "\" not found. Must be one of: [" +
        QuotedNames(nuclides, [](const Nuclide& n) { return n.name; }) + "]");
  return static_cast<size_t>(it - nuclides.cbegin());
}

// --------------------------------------------------------------------- source
Source::Source(const xml::Node& source_node) {
  // Distribution<T>::Create, Source.cpp:28-101
  const xml::Node& position = Require(source_node.child("position"), source_node, "position");
  const xml::Node& direction = Require(source_node.child("direction"), source_node, "direction");
  const xml::Node& energy = Require(source_node.child("energy"), source_node, "energy");
  const xml::Node& particle_type = Require(source_node.child("particletype"), source_node, "particletype");
  const xml::Node* p = position.first_child();
  if (!p || p->name() != "constant") throw std::runtime_error(position.path() + ": only a constant position is supported");
  desc.position[0] = p->attribute_double("x");
  desc.position[1] = p->attribute_double("y");
  desc.position[2] = p->attribute_double("z");
  const xml::Node* d = direction.first_child();
  if (!d) throw std::runtime_error(direction.path() + ": empty direction node");
  desc.direction[0] = 1;
  if (d->name() == "constant" || d->name() == "isotropic-flux") {
    desc.direction_kind = d->name() == "constant" ? MMC_DIR_CONSTANT : MMC_DIR_ISOTROPIC_FLUX;
    desc.direction[0] = d->attribute_double("x");
    desc.direction[1] = d->attribute_double("y");
    desc.direction[2] = d->attribute_double("z");
  } else if (d->name() == "isotropic") {
    desc.direction_kind = MMC_DIR_ISOTROPIC;
  } else {
    throw std::runtime_error(d->path() + ": unknown direction distribution");
  }
  const xml::Node* e = energy.first_child();
  if (!e || e->name() != "constant") throw std::runtime_error(energy.path() + ": only a constant energy is supported");
  const std::string energy_value = e->attribute("energy");
  if (GroupStructureSize(source_node) > 0) desc.group = std::stoull(energy_value);
  else desc.energy = std::stod(energy_value);
  const xml::Node* t = particle_type.first_child();
  if (!t || t->attribute("type") != "neutron")
    throw std::runtime_error(particle_type.path() + ": only neutron sources are implemented on the GPU path");
}

// ----------------------------------------------------------------------- bins
Bins Bins::Create(const xml::Node* bins_node) {
  Bins b;
  if (!bins_node) return b;  // NoBins
  const std::string& type = bins_node->name();
  if (type == "linspace" || type == "logspace") {
    b.kind = type == "linspace" ? MMC_BINS_LINSPACE : MMC_BINS_LOGSPACE;
    b.n_bins = bins_node->attribute_ull("bins") + 2;
    b.base = bins_node->attribute_double("base", 10);
    b.lower = bins_node->attribute_double("min");
    b.upper = bins_node->attribute_double("max");
    if (b.upper <= b.lower) throw std::runtime_error(bins_node->path() + ": max must be strictly greater than min");
    b.width = (b.upper - b.lower) / static_cast<Real>(b.n_bins - 2);
  } else if (type == "boundaries") {
    b.kind = MMC_BINS_BOUNDARIES;
    Real prev_boundary = -std::numeric_limits<Real>::infinity();
    for (const Real boundary : ParseRealList(bins_node->text())) {
      if (boundary <= prev_boundary)
        throw std::runtime_error(
            bins_node->path() + ": nonincreasing elements found: " + std::to_string(prev_boundary) + " " +
            std::to_string(boundary));
      b.boundaries.push_back(boundary);
      prev_boundary = boundary;
    }
    b.n_bins = b.boundaries.size() + 1;
  } else {
    throw std::runtime_error(bins_node->path() + ": unknown bins type");
  }
  return b;
}

size_t Bins::size() const noexcept { return n_bins; }

std::string Bins::to_string() const noexcept {
  // NoBins / LinspaceBins / LogspaceBins / BoundaryBins ::to_string (Bins.cpp:45,84-91,125-132,164-170)
  if (kind == MMC_BINS_NONE) return "none";
  std::stringstream sstream;
  sstream << std::scientific;
  if (kind == MMC_BINS_LINSPACE)
    for (size_t i = 0; i < n_bins - 1; i++) sstream << lower + i * width << ", ";
  else if (kind == MMC_BINS_LOGSPACE)
    for (size_t i = 0; i < n_bins - 1; i++) sstream << std::pow(base, lower + i * width) << ", ";
  else
    for (const Real boundary : boundaries) sstream << boundary << ", ";
  return sstream.str();
}

ParticleBins::ParticleBins(const xml::Node* bins_node) {
  const xml::Node* cosine_node = bins_node ? bins_node->child("cosine") : nullptr;
  const xml::Node* energy_node = bins_node ? bins_node->child("energy") : nullptr;
  if (cosine_node)
    direction = Point{cosine_node->attribute_double("u"), cosine_node->attribute_double("v"), cosine_node->attribute_double("w")};
  cosine = Bins::Create(cosine_node ? cosine_node->first_child() : nullptr);
  energy = Bins::Create(energy_node ? energy_node->first_child() : nullptr);
}

std::string ParticleBins::to_string() const noexcept {
  std::string result;
  result += "cosine\n------\n" + cosine.to_string() + "\n\n";
  result += "energy\n------\n" + energy.to_string() + "\n\n";
  return result;
}

// ---------------------------------------------------------------- perturbations
PerturbationSet::PerturbationSet(const xml::Node* perturbations_node, const World& world) {
  // PerturbationSet::PerturbationSet + Perturbation::Create (Perturbation.cpp:22-34,90-99)
  if (!perturbations_node) return;
  for (const auto& node : perturbations_node->children()) {
    if (node->name() != "total") throw std::runtime_error(node->path() + ": unknown perturbation type");
    perturbations.push_back(Perturbation{node->attribute("name"), world.FindNuclideIndexByName(node->attribute("nuclide"))});
  }
}

const Perturbation& PerturbationSet::FindPerturbationByName(const std::string& name) const {
  // Perturbation.cpp:101-119
  const auto it = std::find_if(perturbations.cbegin(), perturbations.cend(), [&](const Perturbation& p) { return p.name == name; });
  if (it == perturbations.cend())
    throw std::runtime_error(
        "Perturbation \"" + name + "\" not found. Must be one of: [" +
        QuotedNames(perturbations, [](const Perturbation& p) { return p.name; }) + "]");
  return *it;
}

// ------------------------------------------------------------------ estimators
Estimator::Estimator(const xml::Node& estimator_node, const World& world, const PerturbationSet& perturbations)
    : name{estimator_node.attribute("name")}, bins{estimator_node.child("bins")},
      surface{world.FindSurfaceIndexByName(estimator_node.attribute("surface"))}, scores(bins.size(), 0),
      square_scores(bins.size(), 0) {
  if (estimator_node.name() != "current") throw std::runtime_error(estimator_node.path() + ": unknown estimator type");
  // Estimator::Estimator + Sensitivity::Create (Estimator.cpp:83-92, Sensitivity.cpp:13-21)
  if (const xml::Node* sensitivities_node = estimator_node.child("sensitivities"))
    for (const auto& node : sensitivities_node->children()) {
      const Perturbation& perturbation = perturbations.FindPerturbationByName(node->attribute("name"));
      sensitivities.push_back(Sensitivity{
          name + "::" + perturbation.name, perturbation.nuclide, std::vector<Real>(bins.size(), 0),
          std::vector<Real>(bins.size(), 0)});
    }
}

namespace {
// Scorable::GetScoreAsString, Scorable.cpp:51-70
std::string ScoreAsString(const std::vector<Real>& scores, const std::vector<Real>& square_scores, Real total_weight) {
  std::stringstream sstream;
  sstream << "mean\n----\n";
  sstream << std::scientific;
  for (const Real score : scores) sstream << score / total_weight << ", ";
  sstream << "\n\n";
  sstream << "std dev\n-------\n";
  for (size_t i = 0; i < scores.size(); i++)
    sstream << std::sqrt(square_scores[i] - scores[i] * scores[i] / total_weight) / total_weight << ", ";
  sstream << "\n\n";
  return sstream.str();
}
}  // namespace

std::string Estimator::to_string(Real total_weight) const noexcept {
  // Estimator::to_string (Estimator.cpp:48-57) and Sensitivity::to_string (Sensitivity.cpp:25-28)
  std::string result;
  result += name + "\n" + std::string(name.size(), '=') + "\n\n";
  result += bins.to_string();
  result += ScoreAsString(scores, square_scores, total_weight);
  for (const Sensitivity& s : sensitivities)
    result += s.name + "\n" + std::string(s.name.size(), '=') + "\n\n" + ScoreAsString(s.scores, s.square_scores, total_weight) + "\n";
  return result;
}

Estimator& Estimator::operator+=(const Estimator& other) noexcept {
  for (size_t i = 0; i < scores.size(); i++) {
    scores[i] += other.scores[i];
    square_scores[i] += other.square_scores[i];
  }
  // Estimator.cpp:62-70: sensitivities are matched by name
  for (Sensitivity& s : sensitivities) {
    const auto matched = std::find_if(other.sensitivities.cbegin(), other.sensitivities.cend(), [&](const Sensitivity& o) { return o.name == s.name; });
    if (matched == other.sensitivities.cend()) continue;
    for (size_t i = 0; i < s.scores.size(); i++) {
      s.scores[i] += matched->scores[i];
      s.square_scores[i] += matched->square_scores[i];
    }
  }
  return *this;
}

EstimatorSet::EstimatorSet(const xml::Node* estimators_node, const World& world, const PerturbationSet& perturbations, Real total_weight)
    : total_weight{total_weight} {
  if (estimators_node)
    for (const auto& estimator_node : estimators_node->children()) estimators.emplace_back(*estimator_node, world, perturbations);
}

size_t EstimatorSet::total_sensitivities() const noexcept {
  size_t n = 0;
  for (const auto& e : estimators) n += e.sensitivities.size();
  return n;
}

const Estimator& EstimatorSet::FindEstimatorByName(const std::string& name) const {
  const auto it = std::find_if(estimators.cbegin(), estimators.cend(), [&](const Estimator& e) { return e.name == name; });
  if (it == estimators.cend())
    throw std::runtime_error(
        "Estimator \"" + name + "\" notfound. Must be one of: [" +
        QuotedNames(estimators, [](const Estimator& e) { return e.name; }) + "]");
  return *it;
}

std::string EstimatorSet::to_string() const noexcept {
  std::string result;
  for (const auto& estimator : estimators) result += "\n" + estimator.to_string(total_weight) + "\n";
  return result;
}

EstimatorSet& EstimatorSet::operator+=(const EstimatorSet& other) {
  for (auto& estimator : estimators) {
    const auto matched = std::find_if(
        other.estimators.cbegin(), other.estimators.cend(), [&](const Estimator& o) { return o.name == estimator.name; });
    if (matched == other.estimators.cend()) throw std::runtime_error("Estimator not found: " + estimator.name);
    estimator += *matched;
  }
  return *this;
}

size_t EstimatorSet::total_bins() const noexcept {
  size_t n = 0;
  for (const auto& e : estimators) n += e.bins.size();
  return n;
}

}  // namespace minimc
