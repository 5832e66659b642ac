"""ORACLE TEST INFRASTRUCTURE -- not product code.

XML deck -> flat tables, restating the reference's construction code with
xml.etree + numpy, independently of the product's C++ host flattener:

  World::CreateCSGSurfaces / CreateNuclides / CreateMaterials / CreateCells   World.cpp:89-182
  Cell::AssignSurfaceSenses / AssignMaterial                                   Cell.cpp:57-101
  Material::AssignNuclides (afrac normalisation)                               Material.cpp:66-93
  Multigroup (OneDimensional / TwoDimensional / Normalized / total / scatter)  Multigroup.cpp:77-243
  Source::Source / Distribution<T>::Create                                     Source.cpp:28-141
  ParticleBins / LinspaceBins / LogspaceBins / BoundaryBins                    Bins.cpp:57-190
  Driver::Driver (histories, seed, threads, tracking)                          Driver.cpp:37-52

Iteration order of the reference's pointer-keyed std::maps (quirk Q1) is taken
to be World creation order; tests/golden/*_world.json, dumped from the
reference's own objects by oracle/ref_harness `dump`, pins that choice.
"""
from __future__ import annotations

import xml.etree.ElementTree as ET

import numpy as np

SURFACE_TYPES = {"sphere": 0, "planex": 1, "cylinderx": 2}
REACTION_BITS = {"capture": 1, "scatter": 2, "fission": 4}


def _floats(text):
    return [float(t) for t in (text or "").split()]


def _find_by_name(parent, name, what):
    for child in parent:
        if child.get("name") == name:
            return child
    raise RuntimeError(f'{what} node "{name}" not found')


def _multigroup(particle_node, G):
    """Multigroup::Multigroup for one nuclide; returns dict of arrays."""
    out = {
        "mask": 0, "capture": np.zeros(G), "scatter": np.zeros(G), "fission": np.zeros(G), "nubar": np.zeros(G),
        "scatter_probs": np.zeros((G, G)), "chi": np.zeros((G, G)),
    }

    def one_d(node):
        v = _floats(node.text)
        if len(v) != G:
            raise RuntimeError(f"Expected {G} entries but got {len(v)}")
        return np.array(v)

    def two_d(node):
        flat = _floats(node.text)
        if len(flat) != G * G:
            raise RuntimeError(f"Expected {G * G} entries but got {len(flat)}")
        # column gp (incoming group) = [flat[G*(g-1) + (gp-1)] for g = 1..G]
        return np.array([[flat[G * g + gp] for g in range(G)] for gp in range(G)])

    def normalised(cols):
        cols = cols.copy()
        for gp in range(G):
            s = 0.0
            for g in range(G):
                s = s + cols[gp, g]
            if s != 0.0:
                for g in range(G):
                    cols[gp, g] = cols[gp, g] / s
        return cols

    for reaction in particle_node:
        name = reaction.tag
        if name not in REACTION_BITS:
            raise RuntimeError(f"Unrecognized reaction name: {name}")
        if out["mask"] & REACTION_BITS[name]:
            continue  # std::map::emplace keeps the first
        out["mask"] |= REACTION_BITS[name]
        if name == "capture":
            out["capture"] = one_d(reaction)
        elif name == "scatter":
            cols = two_d(reaction)
            sums = np.zeros(G)
            for gp in range(G):
                s = 0.0
                for g in range(G):
                    s = s + cols[gp, g]
                sums[gp] = s
            out["scatter"] = sums
            out["scatter_probs"] = normalised(cols)
        else:
            out["fission"] = one_d(reaction.find("xs"))
    fission = particle_node.find("fission")
    if fission is not None:
        if fission.find("nubar") is not None:
            out["nubar"] = one_d(fission.find("nubar"))
        if fission.find("chi") is not None:
            out["chi"] = normalised(two_d(fission.find("chi")))
    scatter = particle_node.find("scatter")
    if scatter is not None:
        out["scatter_probs"] = normalised(two_d(scatter))
    # CreateTotalXS: accumulate over the map in enum order
    total = np.zeros(G)
    for name in ("capture", "scatter", "fission"):
        if out["mask"] & REACTION_BITS[name]:
            total = total + out[name]
    out["total"] = total
    return out


def _bins(node):
    """Bins::Create on the first child of a <cosine>/<energy> node."""
    child = None if node is None or len(node) == 0 else node[0]
    if child is None:
        return {"kind": 0, "n_bins": 1}
    if child.tag in ("linspace", "logspace"):
        bins = int(child.get("bins"))
        lower, upper = float(child.get("min")), float(child.get("max"))
        if upper <= lower:
            raise RuntimeError("max must be strictly greater than min")
        spec = {"kind": 1 if child.tag == "linspace" else 2, "n_bins": bins + 2, "lower": lower, "upper": upper,
                "width": (upper - lower) / bins, "base": float(child.get("base", 10))}
        return spec
    if child.tag == "boundaries":
        b = _floats(child.text)
        for lo, hi in zip(b, b[1:]):
            if hi <= lo:
                raise RuntimeError("nonincreasing elements found")
        return {"kind": 3, "n_bins": len(b) + 1, "boundaries": np.array(b)}
    raise RuntimeError(child.tag)


def flatten(xml_path_or_text, is_text=False):
    root = ET.fromstring(xml_path_or_text) if is_text else ET.parse(xml_path_or_text).getroot()
    cells_node, surfaces_node = root.find("cells"), root.find("surfaces")
    materials_node, nuclides_node = root.find("materials"), root.find("nuclides")
    energy_node = nuclides_node[0]
    if energy_node.tag != "multigroup":
        raise NotImplementedError("oracle/flatten.py handles multigroup decks")
    G = int(energy_node.get("groups"))

    # World::CreateCSGSurfaces
    surface_names = []
    for cell in cells_node:
        for s in cell:
            if s.get("name") not in surface_names:
                surface_names.append(s.get("name"))
    surface_type, surface_param = [], []
    for name in surface_names:
        node = _find_by_name(surfaces_node, name, "Surface")
        surface_type.append(SURFACE_TYPES[node.tag])
        if node.tag == "sphere":
            c = node.find("center")
            prm = [float(c.get("x")), float(c.get("y")), float(c.get("z")), float(node.find("radius").get("r"))]
        elif node.tag == "planex":
            prm = [float(node.get("x")), 0.0, 0.0, 0.0]
        else:
            prm = [float(node.get("r")), 0.0, 0.0, 0.0]
        surface_param.append(prm)

    # World::CreateNuclides
    nuclide_names = []
    for cell in cells_node:
        if cell.tag == "void":
            continue
        mat = _find_by_name(materials_node, cell.get("material"), "Material")
        for n in mat:
            if n.get("name") not in nuclide_names:
                nuclide_names.append(n.get("name"))
    particles = root.find("general").find("particles").text.split()
    if particles != ["neutron"]:
        raise NotImplementedError("only neutron decks")
    nuclides = []
    for name in nuclide_names:
        node = _find_by_name(energy_node, name, "Nuclide")
        pnode = node.find("neutron")
        if pnode is None:
            raise RuntimeError('"neutron" node not found')
        nuclides.append(_multigroup(pnode, G))

    # World::CreateMaterials
    material_names = []
    for cell in cells_node:
        if cell.tag != "void" and cell.get("material") not in material_names:
            material_names.append(cell.get("material"))
    aden, nuc_begin, nuc_index, nuc_afrac = [], [0], [], []
    for name in material_names:
        mat = _find_by_name(materials_node, name, "Material")
        aden.append(float(mat.get("aden")))
        entries = {}
        for n in mat:
            idx = nuclide_names.index(n.get("name"))
            entries.setdefault(idx, float(n.get("afrac")))
        order = sorted(entries)  # pointer order ~ World creation order (Q1)
        s = 0.0
        for idx in order:
            s = s + entries[idx]
        for idx in order:
            nuc_index.append(idx)
            nuc_afrac.append(entries[idx] / s)
        nuc_begin.append(len(nuc_index))

    # World::CreateCells
    cell_material, surf_begin, surf_index, surf_sense = [], [0], [], []
    for cell in cells_node:
        cell_material.append(-1 if cell.get("name") is None else material_names.index(cell.get("material")))
        entries = {}
        for s in cell:
            sense = s.get("sense")
            assert sense in ("-1", "+1")
            entries.setdefault(surface_names.index(s.get("name")), 1 if sense == "-1" else 0)
        for idx in sorted(entries):
            surf_index.append(idx)
            surf_sense.append(entries[idx])
        surf_begin.append(len(surf_index))
    n_cells = len(cell_material)

    def stack(key, shape):
        return np.array([n[key] for n in nuclides], dtype=np.float64).reshape((len(nuclides),) + shape)

    world = {
        "n_groups": G,
        "surface_type": np.array(surface_type, np.int32),
        "surface_param": np.array(surface_param, np.float64).reshape(-1),
        "cell_material": np.array(cell_material, np.int32),
        "cell_surface_begin": np.array(surf_begin, np.int32),
        "cell_surface_index": np.array(surf_index, np.int32),
        "cell_surface_sense": np.array(surf_sense, np.int32),
        "cell_field_kind": np.zeros(n_cells, np.int32),
        "cell_field_param": np.zeros(n_cells * 6, np.float64),
        "material_aden": np.array(aden, np.float64),
        "material_nuclide_begin": np.array(nuc_begin, np.int32),
        "material_nuclide_index": np.array(nuc_index, np.int32),
        "material_nuclide_afrac": np.array(nuc_afrac, np.float64),
        "mg_reaction_mask": np.array([n["mask"] for n in nuclides], np.uint32),
        "mg_total": stack("total", (G,)).reshape(-1),
        "mg_capture": stack("capture", (G,)).reshape(-1),
        "mg_scatter": stack("scatter", (G,)).reshape(-1),
        "mg_fission": stack("fission", (G,)).reshape(-1),
        "mg_nubar": stack("nubar", (G,)).reshape(-1),
        "mg_scatter_probs": stack("scatter_probs", (G, G)).reshape(-1),
        "mg_chi": stack("chi", (G, G)).reshape(-1),
    }
    names = {"surfaces": surface_names, "nuclides": nuclide_names, "materials": material_names,
             "cells": [c.get("name") or "" for c in cells_node]}

    # Driver::Driver
    general = root.find("general")
    tracking = (general.findtext("tracking") or "surface").strip()
    problem = root.find("problemtype")[0]
    src_node = problem if problem.tag == "fixedsource" else problem.find("initialsource")
    run = {
        "histories": int(general.findtext("histories")),
        "seed": int(general.findtext("seed") or "1"),
        "threads": int(general.findtext("threads")),
        "tracking": {"surface": 0, "cell delta": 1}[tracking],
        "problemtype": problem.tag,
        "inactive": int(problem.get("inactive", 0)),
        "active": int(problem.get("active", 0)),
    }

    # Source::Source
    pos = src_node.find("position")[0]
    dnode = src_node.find("direction")[0]
    source = {
        "position": (float(pos.get("x")), float(pos.get("y")), float(pos.get("z"))),
        "direction_kind": {"constant": 0, "isotropic": 1, "isotropic-flux": 2}[dnode.tag],
        "direction": (1.0, 0.0, 0.0),
        "group": int(src_node.find("energy")[0].get("energy")),
    }
    if dnode.tag != "isotropic":
        source["direction"] = (float(dnode.get("x")), float(dnode.get("y")), float(dnode.get("z")))

    # EstimatorSet::EstimatorSet
    estimators = []
    est_node = root.find("estimators")
    for e in (est_node if est_node is not None else []):
        assert e.tag == "current"
        bins = e.find("bins")
        cosine = None if bins is None else bins.find("cosine")
        energy = None if bins is None else bins.find("energy")
        spec = {
            "name": e.get("name"),
            "surface": surface_names.index(e.get("surface")),
            "cosine_direction": None if cosine is None else
            (float(cosine.get("u")), float(cosine.get("v")), float(cosine.get("w"))),
            "cosine": _bins(cosine),
            "energy": _bins(energy),
        }
        spec["n_bins"] = spec["cosine"]["n_bins"] * spec["energy"]["n_bins"]
        estimators.append(spec)
    return {"world": world, "names": names, "run": run, "source": source, "estimators": estimators}
