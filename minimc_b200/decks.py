"""Input decks, generated.

The reference's decks live in /root/reference/test and /root/reference/benchmarks,
which do not exist on the GPU box, so the decks this repo measures and tests are
generated here from compact Python descriptions in the reference's XML schema
(test/minimc.xsd).  `critical`, `three_shells` and `leakage_sphere` describe the
same problems as the reference's test/multigroup_critical.xml,
test/multigroup.xml and test/point_source_leakage.xml (tests/test_decks.py
checks, when the reference tree is present, that they flatten to identical
tables); the others are derived decks that exercise fission chains, planes,
cylinders and the isotropic-flux source.
"""
from __future__ import annotations

from xml.sax.saxutils import quoteattr


def _general(histories, threads, seed, tracking, chunksize=100):
    s = f"<general>\n  <particles>neutron</particles>\n  <histories>{histories}</histories>\n"
    s += f"  <threads>{threads}</threads>\n  <chunksize>{chunksize}</chunksize>\n"
    if seed is not None:
        s += f"  <seed>{seed}</seed>\n"
    if tracking is not None:
        s += f"  <tracking>{tracking}</tracking>\n"
    return s + "</general>\n"


def _matrix(rows):
    return "\n".join("          " + " ".join(repr(float(v)) if not isinstance(v, str) else v for v in r) for r in rows)


def _nuclide(name, capture=None, scatter=None, fission=None):
    """capture: list[G]; scatter: G x G rows (row = outgoing group); fission: dict(xs, nubar, chi rows)."""
    s = f"    <nuclide name={quoteattr(name)}>\n      <neutron>\n"
    if capture is not None:
        s += "        <capture>" + " ".join(map(str, capture)) + "</capture>\n"
    if scatter is not None:
        s += "        <scatter>\n" + _matrix(scatter) + "\n        </scatter>\n"
    if fission is not None:
        s += "        <fission>\n          <xs>" + " ".join(map(str, fission["xs"])) + "</xs>\n"
        s += "          <nubar>" + " ".join(map(str, fission["nubar"])) + "</nubar>\n"
        s += "          <chi>\n" + _matrix(fission["chi"]) + "\n          </chi>\n        </fission>\n"
    return s + "      </neutron>\n    </nuclide>\n"


def _source(position=(0, 0, 0), direction="isotropic", group=1, tag="fixedsource", attrs=""):
    if direction == "isotropic":
        d = "<isotropic/>"
    elif direction[0] == "flux":
        d = '<isotropic-flux x="%s" y="%s" z="%s"/>' % tuple(direction[1])
    else:
        d = '<constant x="%s" y="%s" z="%s"/>' % tuple(direction)
    body = (
        f'    <position>\n      <constant x="{position[0]}" y="{position[1]}" z="{position[2]}"/>\n    </position>\n'
        f"    <direction>\n      {d}\n    </direction>\n"
        f'    <energy>\n      <constant energy="{group}"/>\n    </energy>\n'
        '    <particletype>\n      <constant type="neutron"/>\n    </particletype>\n')
    if tag == "fixedsource":
        return f"<problemtype>\n  <fixedsource>\n{body}  </fixedsource>\n</problemtype>\n"
    return f"<problemtype>\n  <keigenvalue {attrs}>\n  <initialsource>\n{body}  </initialsource>\n  </keigenvalue>\n</problemtype>\n"


def _bins_xml(spec):
    kind = spec[0]
    if kind == "linspace":
        return '<linspace min="%s" max="%s" bins="%d"/>' % spec[1:]
    if kind == "logspace":
        return '<logspace min="%s" max="%s" bins="%d" base="%s"/>' % spec[1:]
    if kind == "boundaries":
        return "<boundaries>" + " ".join(map(str, spec[1])) + "</boundaries>"
    raise ValueError(kind)


def _estimators(estimators):
    """estimators: list of dict(name, surface, cosine=(u,v,w,binspec)|None, energy=binspec|None)."""
    if not estimators:
        return ""
    s = "<estimators>\n"
    for e in estimators:
        s += f'  <current name={quoteattr(e["name"])} surface={quoteattr(e["surface"])}>\n'
        if e.get("sensitivities"):
            s += "    <sensitivities>\n"
            s += "".join(f"      <perturbation name={quoteattr(n)}/>\n" for n in e["sensitivities"])
            s += "    </sensitivities>\n"
        if e.get("cosine") or e.get("energy"):
            s += "    <bins>\n"
            if e.get("cosine"):
                u, v, w, spec = e["cosine"]
                s += f'      <cosine u="{u}" v="{v}" w="{w}">\n        {_bins_xml(spec)}\n      </cosine>\n'
            if e.get("energy"):
                s += f"      <energy>\n        {_bins_xml(e['energy'])}\n      </energy>\n"
            s += "    </bins>\n"
        s += "  </current>\n"
    return s + "</estimators>\n"


def _perturbations(perturbations):
    """perturbations: list of (name, nuclide) -> <perturbations><total .../></perturbations>."""
    if not perturbations:
        return ""
    return ("<perturbations>\n" + "".join(f"  <total name={quoteattr(n)} nuclide={quoteattr(nuc)}/>\n" for n, nuc in perturbations)
            + "</perturbations>\n")


def _deck(general, groups, nuclides, materials, surfaces, cells, problem, estimators, perturbations=None):
    return (
        '<minimc\n  xmlns:xsi="http://www.w3.org/2001/XMLSchema-instance"\n'
        '  xsi:noNamespaceSchemaLocation="minimc.xsd">\n'
        + general
        + f'<nuclides>\n  <multigroup groups="{groups}">\n' + "".join(nuclides) + "  </multigroup>\n</nuclides>\n"
        + "<materials>\n" + "".join(materials) + "</materials>\n"
        + "<surfaces>\n" + "".join(surfaces) + "</surfaces>\n"
        + "<cells>\n" + "".join(cells) + "</cells>\n"
        + problem + _estimators(estimators) + _perturbations(perturbations) + "</minimc>\n")


def _material(name, aden, nuclides):
    s = f'  <material name={quoteattr(name)} aden="{aden}">\n'
    for n, afrac in nuclides:
        s += f'    <nuclide name={quoteattr(n)} afrac="{afrac}"/>\n'
    return s + "  </material>\n"


def _sphere(name, r, center=(0, 0, 0)):
    return (f'  <sphere name={quoteattr(name)}>\n    <center x="{center[0]}" y="{center[1]}" z="{center[2]}"/>\n'
            f'    <radius r="{r}"/>\n  </sphere>\n')


def _cell(name, material, surfaces):
    head = f'  <cell name={quoteattr(name)} material={quoteattr(material)}>\n' if name else "  <void>\n"
    body = "".join(f'    <surface name={quoteattr(n)} sense="{s}"/>\n' for n, s in surfaces)
    return head + body + ("  </cell>\n" if name else "  </void>\n")


# ------------------------------------------------------------------ the decks
def critical(histories=100000, threads=2, seed=None, tracking=None, estimators=None):
    """BASELINE config C1 / M1: one-group infinite medium, c = 0.25 (test/multigroup_critical.xml)."""
    return _deck(
        _general(histories, threads, seed, tracking), 1,
        [_nuclide("fissile", capture=[0.75], scatter=[[0.25]], fission={"xs": [0], "nubar": [1], "chi": [[1.0]]})],
        [_material("fissile", 1, [("fissile", 1.0)])],
        [_sphere("sphere", "1e10")],
        [_cell("sphere", "fissile", [("sphere", "-1")]), _cell(None, None, [("sphere", "+1")])],
        _source(), estimators)


def leakage_sphere(histories=1000, threads=2, seed=None, tracking=None):
    """Pure absorber sphere of radius 1 mfp with a leakage estimator (test/point_source_leakage.xml)."""
    return _deck(
        _general(histories, threads, seed, tracking, chunksize=500), 1,
        [_nuclide("absorber", capture=[1])],
        [_material("absorber", 1, [("absorber", 1)])],
        [_sphere("sphere", 1)],
        [_cell("sphere", "absorber", [("sphere", "-1")]), _cell(None, None, [("sphere", "+1")])],
        _source(), [{"name": "leakage", "surface": "sphere"}])


THREE_SHELL_ESTIMATORS = [
    {"name": "inner", "surface": "inner shell", "energy": ("boundaries", [1.5])},
    {"name": "middle", "surface": "middle shell", "energy": ("boundaries", [1.5])},
    {"name": "outer", "surface": "outer shell", "cosine": (1, 0, 0, ("linspace", -1, 1, 4)),
     "energy": ("boundaries", [1.5])},
]


def three_shells(histories=100000, threads=2, seed=None, tracking=None, estimators=None, perturbations=None):
    """M2: two groups, two-nuclide water, three concentric shells (test/multigroup.xml); `estimators`
    defaults to none as in the reference file, THREE_SHELL_ESTIMATORS reproduces golden G2."""
    return _deck(
        _general(histories, threads, seed, tracking, chunksize=1000), 2,
        [_nuclide("hydrogen", capture=[0, 1], scatter=[[0, 0], [1, 0]]),
         _nuclide("oxygen", capture=[0.5, 0.5], scatter=[[0.5, 0.0], [0.5, 0.5]]),
         _nuclide("uranium235", capture=[0.33, 0.67], scatter=[[1, 0], [0, 1]],
                  fission={"xs": [0, 1], "nubar": [0, 2.43], "chi": [[0, 0.5], [0, 0.5]]})],
        [_material("water", 2, [("hydrogen", 0.67), ("oxygen", 0.33)]),
         _material("hydrogen", 1, [("hydrogen", 0.1)]),
         _material("oxygen", 1, [("oxygen", 0.1)])],
        [_sphere("inner shell", 1), _sphere("middle shell", 2), _sphere("outer shell", 3)],
        [_cell("pit", "hydrogen", [("inner shell", "-1")]),
         _cell("inner shell", "water", [("inner shell", "+1"), ("middle shell", "-1")]),
         _cell("outer shell", "hydrogen", [("middle shell", "+1"), ("outer shell", "-1")]),
         _cell(None, None, [("outer shell", "+1")])],
        _source(), estimators, perturbations)


# N4: differential-operator sensitivities of the three-shell currents to the total cross sections of
# hydrogen (in every material) and oxygen (in the water shell only)
SENSITIVITY_PERTURBATIONS = [("h-total", "hydrogen"), ("o-total", "oxygen")]
SENSITIVITY_ESTIMATORS = [
    dict(THREE_SHELL_ESTIMATORS[0], sensitivities=["h-total", "o-total"]),
    THREE_SHELL_ESTIMATORS[1],
    dict(THREE_SHELL_ESTIMATORS[2], sensitivities=["o-total"]),
]


def sensitivity_shells(histories=20000, threads=2, seed=None, tracking=None):
    """three_shells with sensitivities of the inner and outer currents to total cross-section perturbations."""
    return three_shells(histories=histories, threads=threads, seed=seed, tracking=tracking,
                        estimators=SENSITIVITY_ESTIMATORS, perturbations=SENSITIVITY_PERTURBATIONS)


def fissile_slab(histories=20000, threads=2, seed=None, tracking=None):
    """Derived deck: subcritical two-group fissile slab between x planes, bounded sideways by a sphere;
    fission chains run inside each fixed-source history (FixedSource.cpp:63-72), an isotropic-flux source
    enters from the left face.  Exercises PlaneX, fission banking, cosine + logspace/linspace/boundary bins."""
    fuel = _nuclide("fuel", capture=[0.2, 0.6], scatter=[[0.4, 0.0], [0.3, 0.5]],
                    fission={"xs": [0.1, 0.4], "nubar": [2.6, 2.43], "chi": [[0.9, 0.8], [0.1, 0.2]]})
    moderator = _nuclide("moderator", capture=[0.02, 0.1], scatter=[[0.3, 0.0], [0.6, 0.9]])
    surfaces = ['  <planex name="left" x="0"/>\n', '  <planex name="mid" x="1.5"/>\n',
                '  <planex name="right" x="3"/>\n', _sphere("wall", 2.5, (1.5, 0, 0))]
    cells = [
        _cell("fuel", "fuel", [("left", "+1"), ("mid", "-1"), ("wall", "-1")]),
        _cell("moderator", "mix", [("mid", "+1"), ("right", "-1"), ("wall", "-1")]),
        _cell(None, None, [("left", "-1")]),
        _cell(None, None, [("right", "+1")]),
        _cell(None, None, [("wall", "+1")]),
    ]
    estimators = [
        {"name": "transmitted", "surface": "right", "cosine": (1, 0, 0, ("linspace", 0, 1, 5)),
         "energy": ("linspace", 0.5, 2.5, 2)},
        {"name": "reflected", "surface": "left", "energy": ("logspace", -1, 1, 4, 10)},
        {"name": "interface", "surface": "mid", "cosine": (1, 0, 0, ("boundaries", [-0.5, 0.0, 0.5]))},
        {"name": "side", "surface": "wall"},
    ]
    return _deck(
        _general(histories, threads, seed, tracking), 2, [fuel, moderator],
        [_material("fuel", 1.0, [("fuel", 1.0)]), _material("mix", 0.8, [("moderator", 3.0), ("fuel", 1.0)])],
        surfaces, cells, _source(position=(1e-9, 0, 0), direction=("flux", (1, 0, 0)), group=1), estimators)


def pipe(histories=20000, threads=2, seed=None, tracking=None):
    """Derived deck: a scattering pipe (CylinderX r = 0.8 about the x axis) cut by three planes.  Keeps the
    reference's CylinderX::Contains quirk Q2 observable (it compares sqrt(r_perp^2) with r^2 = 0.64,
    CSGSurface.cpp:166-174, while Distance uses r = 0.8): a particle that crosses the middle plane at
    0.64 <= r_perp < 0.8 is in neither pipe cell and leaks through the plane.  No fission: a secondary born
    in that annulus would start in a void cell, which is undefined behaviour in the reference."""
    a = _nuclide("steel", capture=[0.05, 0.15], scatter=[[0.6, 0.0], [0.25, 0.7]])
    b = _nuclide("water", capture=[0.01, 0.3], scatter=[[0.2, 0.0], [1.0, 1.4]])
    surfaces = ['  <planex name="inlet" x="-2"/>\n', '  <planex name="joint" x="0.5"/>\n',
                '  <planex name="outlet" x="4"/>\n', '  <cylinderx name="wall" r="0.8"/>\n']
    cells = [
        _cell("upstream", "steel", [("inlet", "+1"), ("joint", "-1"), ("wall", "-1")]),
        _cell("downstream", "wet", [("joint", "+1"), ("outlet", "-1"), ("wall", "-1")]),
        _cell(None, None, [("inlet", "-1")]),
        _cell(None, None, [("outlet", "+1")]),
        _cell(None, None, [("wall", "+1")]),
    ]
    estimators = [
        {"name": "wall", "surface": "wall", "cosine": (1, 0, 0, ("linspace", -1, 1, 10)),
         "energy": ("boundaries", [1.5])},
        {"name": "joint", "surface": "joint", "cosine": (1, 0, 0, ("linspace", -1, 1, 2))},
        {"name": "outlet", "surface": "outlet"},
        {"name": "inlet", "surface": "inlet", "energy": ("logspace", -1, 1, 2, 10)},
    ]
    return _deck(
        _general(histories, threads, seed, tracking), 2, [a, b],
        [_material("steel", 0.9, [("steel", 1.0)]), _material("wet", 1.1, [("water", 2.0), ("steel", 1.0)])],
        surfaces, cells, _source(position=(-1.0, 0.1, 0), direction="isotropic", group=1), estimators)


def offcentre_spheres(histories=20000, threads=2, seed=None, tracking=None):
    """Derived deck: two off-centre overlapping spheres (first-match cell search, World.cpp:26-37), constant
    direction source, three groups with up-scatter."""
    a = _nuclide("a", capture=[0.1, 0.2, 0.3], scatter=[[0.5, 0.1, 0.0], [0.3, 0.4, 0.2], [0.1, 0.3, 0.6]])
    b = _nuclide("b", capture=[0.3, 0.3, 0.9], scatter=[[0.2, 0.0, 0.0], [0.2, 0.2, 0.0], [0.1, 0.4, 0.3]])
    return _deck(
        _general(histories, threads, seed, tracking), 3, [a, b],
        [_material("ma", 0.7, [("a", 1.0)]), _material("mb", 1.3, [("b", 2.0), ("a", 1.0)])],
        [_sphere("s1", 1.5, (0.2, 0.1, -0.1)), _sphere("s2", 2.0, (1.0, 0.0, 0.3)), _sphere("s3", 4.0, (0.5, 0, 0))],
        [_cell("core", "ma", [("s1", "-1")]),
         _cell("lobe", "mb", [("s2", "-1"), ("s3", "-1")]),
         _cell("shell", "ma", [("s3", "-1")]),
         _cell(None, None, [("s3", "+1")])],
        _source(position=(0.1, 0.0, 0.0), direction=(1, 2, 3), group=2),
        [{"name": "out", "surface": "s3", "cosine": (0, 0, 1, ("linspace", -1, 1, 8)),
          "energy": ("boundaries", [1.5, 2.5])},
         {"name": "s1", "surface": "s1", "energy": ("linspace", 0.5, 3.5, 3)}])


# ---------------------------------------------------------------- k-eigenvalue decks
# None of the reference's decks is a k-eigenvalue problem and its KEigenvalue::Solve is a stub;
# these derived decks have analytic answers (SURVEY.md 8(d) "M1k").
def k_unity(histories=20000, threads=2, inactive=2, active=4, tracking=None):
    """One-group infinite medium, capture 0, scatter 0.25, fission 0.75, nubar 1: every history ends in a fission
    that yields exactly size_t(1 + U) = 1 site, so every generation banks exactly `histories` sites: k = 1 with zero
    variance -- an exact integer-bookkeeping check of banking, ordering and resampling."""
    return _deck(
        _general(histories, threads, None, tracking), 1,
        [_nuclide("fuel", scatter=[[0.25]], fission={"xs": [0.75], "nubar": [1], "chi": [[1.0]]})],
        [_material("fuel", 1, [("fuel", 1.0)])],
        [_sphere("sphere", "1e10")],
        [_cell("sphere", "fuel", [("sphere", "-1")]), _cell(None, None, [("sphere", "+1")])],
        _source(tag="keigenvalue", attrs=f'inactive="{inactive}" active="{active}"'), None)


def k_infinite(histories=50000, threads=2, inactive=3, active=12, tracking=None):
    """One-group infinite medium, capture 0.5, scatter 0.25, fission 0.25, nubar 2.43:
    k_inf = nubar * Sigma_f / Sigma_a = 2.43 * 0.25 / 0.75 = 0.81."""
    return _deck(
        _general(histories, threads, None, tracking), 1,
        [_nuclide("fuel", capture=[0.5], scatter=[[0.25]], fission={"xs": [0.25], "nubar": [2.43], "chi": [[1.0]]})],
        [_material("fuel", 1, [("fuel", 1.0)])],
        [_sphere("sphere", "1e10")],
        [_cell("sphere", "fuel", [("sphere", "-1")]), _cell(None, None, [("sphere", "+1")])],
        _source(tag="keigenvalue", attrs=f'inactive="{inactive}" active="{active}"'), None)


def k_slab(histories=20000, threads=2, inactive=3, active=6, tracking=None):
    """Two-group fuel + moderator slab with leakage (the fissile_slab geometry) as a k-eigenvalue problem with
    `current` estimators on its faces: exercises banking in a finite geometry, both groups of chi and tallies that
    accumulate over active cycles only."""
    text = fissile_slab(histories=histories, threads=threads, tracking=tracking)
    a, b = text.index("<problemtype>"), text.index("</problemtype>") + len("</problemtype>\n")
    return text[:a] + _source(position=(0.75, 0, 0), direction="isotropic", group=1, tag="keigenvalue",
                              attrs=f'inactive="{inactive}" active="{active}"') + text[b:]


KDECKS = {"k_unity": k_unity, "k_infinite": k_infinite, "k_slab": k_slab}

SENSITIVITY_DECKS = {"sensitivity_shells": sensitivity_shells}

DECKS = {
    "critical": critical,
    "leakage_sphere": leakage_sphere,
    "three_shells": three_shells,
    "fissile_slab": fissile_slab,
    "pipe": pipe,
    "offcentre_spheres": offcentre_spheres,
}
