"""Continuous-energy / thermal-scattering decks with SYNTHETIC nuclear data.

Every *.hdf5 the reference's benchmarks/*.xml need is a git-lfs pointer, so the continuous-energy
path is exercised -- by the GPU code and by the reference binary alike -- on tables generated here, in the
MMCTAB1 flat format (oracle/shim/H5Cpp.h and minimc_b200/host read the same files):

    char magic[8] = "MMCTAB1\\0"; uint64 ndim; uint64 shape[ndim]; double axis_i[shape[i]]...; double values[prod(shape)]

which is the content of the pandas "fixed" HDF5 layout HDF5DataSet<D> reads (HDF5DataSet.hpp:88-130): D sorted axes
plus row-major values.  Shapes follow the reference's data: H-1 pointwise grids, a rank-R POD of the
inelastic cross section [nE][R] x [R] x [nT][R], four beta partitions (CDF [nF][R], S [R], E_T [nE_p][nT][R]) and four
alpha partitions (CDF [nF][R], S [R], beta_T [nB_p][nT][R]); `size="full"` uses the coarse shapes of the reference's
sweep log, (nF, nT, nE_p) = (54,17,7), (57,18,10), (64,17,21), (97,18,294), R = 10.

The generator uses ONLY + - * / sqrt on float64 (IEEE-exact in numpy on every machine): no exp/log/pow, no SVD, no
random numbers.  The same call therefore writes bit-identical files in the build container and on the GPU box, which
is what lets golden traces produced here by the reference binary be compared there.

The physics is a smooth caricature of H in H2O (1/v capture, ~20 b scattering with a thermal rise, down-scatter
bounded by -E/kT, quantile functions monotone in F); parity, not realism, is the point.
"""
from __future__ import annotations

import os
import struct
from pathlib import Path
from xml.sax.saxutils import quoteattr

import numpy as np

BOLTZMANN = 8.617333262145e-11  # MeV / K, Constants.hpp:24
BETA_CUTOFF = 20.0
ALPHA_CUTOFF = 1636.7475317348378
AWR_H = 0.99916733

SIZES = {
    # name: (pointwise n, ratio), (tsl nE, ratio), nT, rank, beta partitions (nF, nT, nE_p), alpha partitions (nF, nT, nB_p)
    "small": {"pointwise": (60, 1.62), "tsl_E": (64, 1.14), "n_T": 6, "rank": 4,
              "beta": [(9, 5, 6), (10, 6, 10), (11, 5, 16), (13, 6, 32)],  # 64 incident energies, ratio 1.14
              "alpha": [(9, 5, 5), (10, 6, 8), (11, 5, 12), (13, 6, 20)]},
    "full": {"pointwise": (350, 1.0845), "tsl_E": (1000, 1.0123), "n_T": 18, "rank": 10,
             "beta": [(54, 17, 7), (57, 18, 10), (64, 17, 21), (97, 18, 294)],
             "alpha": [(54, 17, 12), (57, 18, 24), (64, 17, 48), (97, 18, 96)]},
}
TSL_CUTOFF_ENERGY = 2.0e-6  # MeV: last point of the scatter_xs_E axis (ThermalScattering.hpp:125-126)


from .mmctab import write_table  # noqa: E402,F401  (the MMCTAB1 writer lives with the converter)


def _geometric(first: float, ratio: float, n: int) -> np.ndarray:
    out = np.empty(n)
    out[0] = first
    for i in range(1, n):
        out[i] = out[i - 1] * ratio
    return out


def _geometric_to(last: float, ratio: float, n: int) -> np.ndarray:
    """n points ending exactly at `last`, each the next one divided by ratio."""
    out = np.empty(n)
    out[-1] = last
    for i in range(n - 2, -1, -1):
        out[i] = out[i + 1] / ratio
    return out


def _temperatures(n: int, lo=273.6, hi=800.0) -> np.ndarray:
    step = (hi - lo) / (n - 1)
    return np.array([lo + k * step for k in range(n)])


def _wiggle(i, r, m):
    """Small integer-valued pattern in [-1, 1] (exact arithmetic)."""
    return ((np.asarray(i) * (r + 1)) % m - (m - 1) / 2.0) / ((m - 1) / 2.0)


def _inelastic_modes(E, T, rank):
    """sigma_inel(E_i, T_k) = sum_r S[r] * XE[i][r] * XT[k][r]."""
    nE, nT = len(E), len(T)
    S = np.zeros(rank)
    XE, XT = np.zeros((nE, rank)), np.zeros((nT, rank))
    S[0] = 1.0
    XE[:, 0] = 20.4 * (1.0 + 2.53e-8 / (E + 5.0e-9))
    XT[:, 0] = 1.0
    if rank > 1:
        S[1] = 0.1
        XE[:, 1] = 40.0 * 1.0e-8 / (E + 1.0e-8)
        XT[:, 1] = (T - 273.6) / 100.0
    for r in range(2, rank):
        S[r] = 0.02 / r
        XE[:, r] = _wiggle(np.arange(nE), r, 7)
        XT[:, r] = _wiggle(np.arange(nT), r, 5)
    return S, XE, XT


def _cdf_axis(n: int) -> np.ndarray:
    # n values strictly inside (0, 1), denser towards 1
    i = np.arange(1, n + 1, dtype=np.float64)
    x = i / (n + 1.0)
    return x * (2.0 - x) * (1.0 - 0.5 / (n + 1.0))


def _beta_partition(F, E, T, rank):
    """beta(F_i, E_j, T_k) = sum_r S[r] * C[i][r] * M[j][k][r]: -E/kT (1-F)^2 + up-scatter quantile."""
    S = np.zeros(rank)
    C = np.zeros((len(F), rank))
    M = np.zeros((len(E), len(T), rank))
    kT = BOLTZMANN * T
    S[0] = 1.0
    C[:, 0] = (1.0 - F) * (1.0 - F)
    M[:, :, 0] = -(E[:, None] / kT[None, :])
    if rank > 1:
        S[1] = 1.0
        C[:, 1] = 0.05 + 0.45 * F * F / (1.04 - F)  # floor 0.05 > E_min / kT: below-grid energies never reject
        M[:, :, 1] = 1.0 + 0.2 * E[:, None] / (E[:, None] + kT[None, :])
    for r in range(2, rank):
        S[r] = 1.0e-4 / r
        C[:, r] = _wiggle(np.arange(len(F)), r, 5) * F * (1.0 - F)
        M[:, :, r] = _wiggle(np.arange(len(E))[:, None] + np.arange(len(T))[None, :], r, 3)
    return S, C, M


def _alpha_partition(F, B, T, rank):
    """alpha(F_i, beta_j, T_k): quantile function, increasing in F, growing with beta, falling with T."""
    S = np.zeros(rank)
    C = np.zeros((len(F), rank))
    M = np.zeros((len(B), len(T), rank))
    S[0] = 1.0
    C[:, 0] = 40.0 * F * F / (1.1 - F)
    M[:, :, 0] = (1.0 + B[:, None] / 8.0) * (300.0 / T[None, :])
    if rank > 1:
        S[1] = 0.5
        C[:, 1] = F
        M[:, :, 1] = B[:, None] / (1.0 + B[:, None]) + 0.0 * T[None, :]
    for r in range(2, rank):
        S[r] = 1.0e-4 / r
        C[:, r] = _wiggle(np.arange(len(F)), r, 5) * F * (1.0 - F)
        M[:, :, r] = _wiggle(np.arange(len(B))[:, None] + np.arange(len(T))[None, :], r, 3)
    return S, C, M


def generate_tables(directory, size: str = "small") -> dict:
    """Writes every table of one synthetic 'hydrogen in water' nuclide + a fissile heavy nuclide into `directory`
    (created if needed) and returns {logical name: path}."""
    cfg = SIZES[size]
    d = Path(directory)
    d.mkdir(parents=True, exist_ok=True)
    paths = {}

    def put(name, axes, values):
        paths[name] = os.fspath(d / f"{name}.mmctab")
        write_table(paths[name], axes, values)

    # ---- H-1 pointwise (ContinuousEvaluation): capture 1/v, elastic with a thermal rise, total = sum
    n, ratio = cfg["pointwise"]
    E = _geometric(1.0e-11, ratio, n)
    capture = 0.332 * np.sqrt(2.53e-8 / E)
    for T_eval, tag in ((293.6, "293K"), (623.6, "623K")):
        kT = BOLTZMANN * T_eval
        elastic = 20.4 * (1.0 + kT / (2.0 * (E + kT / 10.0)))
        put(f"H1_elastic_{tag}", [E], elastic)
        put(f"H1_total_{tag}", [E], capture + elastic)
    put("H1_gamma", [E], capture)

    # ---- heavy fissile nuclide (free gas above 500 kT / awr is off; exercises ContinuousFission)
    cap_u = 2.0 * np.sqrt(2.53e-8 / E) + 0.5
    ela_u = 10.0 + 0.0 * E
    fis_u = 12.0 * np.sqrt(2.53e-8 / E) + 1.0
    put("U_gamma", [E], cap_u)
    put("U_elastic", [E], ela_u)
    put("U_fission", [E], fis_u)
    put("U_total", [E], cap_u + ela_u + fis_u)
    put("U_nubar", [np.array([1.0e-11, 1.0e-6, 1.0, 20.0])], np.array([2.43, 2.43, 2.6, 5.0]))
    # oxygen-like scatterer (awr 15.86): free gas only below 500 kT / awr
    put("O_gamma", [E], 1.0e-4 * np.sqrt(2.53e-8 / E))
    put("O_elastic", [E], 3.8 + 0.0 * E)
    put("O_total", [E], 1.0e-4 * np.sqrt(2.53e-8 / E) + 3.8)

    # ---- thermal scattering: POD of the inelastic cross section
    nE, ratioE = cfg["tsl_E"]
    Et = _geometric_to(TSL_CUTOFF_ENERGY, ratioE, nE)
    Tt = _temperatures(cfg["n_T"])
    R = cfg["rank"]
    S, XE, XT = _inelastic_modes(Et, Tt, R)
    put("scatter_xs_E", [Et, np.arange(R, dtype=np.float64)], XE)
    put("scatter_xs_S", [np.arange(R, dtype=np.float64)], S)
    put("scatter_xs_T", [Tt, np.arange(R, dtype=np.float64)], XT)
    # majorant = 1.05 * max_T sigma_inel on the same grid (the reference's file name says safety factor 1.05)
    sig = np.zeros((nE, len(Tt)))
    for r in range(R):
        sig = sig + S[r] * XE[:, r][:, None] * XT[:, r][None, :]
    put("majorant", [Et], 1.05 * sig.max(axis=1))

    # ---- beta partitions: consecutive slices of one increasing incident-energy grid ending at the cutoff
    n_total = sum(p[2] for p in cfg["beta"])
    ratio_b = {"small": 1.14, "full": 1.0235}[size]  # first energy ~5e-10 / ~9e-10 MeV < 0.05 k T_min
    Eb = _geometric_to(TSL_CUTOFF_ENERGY, ratio_b, n_total)
    begin = 0
    for i, (nF, nT, nEp) in enumerate(cfg["beta"]):
        F = _cdf_axis(nF)
        Tp = _temperatures(nT)
        Sb, C, M = _beta_partition(F, Eb[begin:begin + nEp], Tp, R)
        put(f"beta_{i}_CDF", [F, np.arange(R, dtype=np.float64)], C)
        put(f"beta_{i}_S", [np.arange(R, dtype=np.float64)], Sb)
        put(f"beta_{i}_E_T", [Eb[begin:begin + nEp], Tp, np.arange(R, dtype=np.float64)], M)
        begin += nEp

    # ---- alpha partitions: consecutive slices of one increasing beta grid reaching past E_cutoff / (k T_min)
    n_total = sum(p[2] for p in cfg["alpha"])
    ratio_a = {"small": 1.53, "full": 1.1085}[size]
    Bg = _geometric(1.0e-6, ratio_a, n_total)
    assert Bg[-1] > TSL_CUTOFF_ENERGY / (BOLTZMANN * 273.6) and Bg[-1] > BETA_CUTOFF, Bg[-1]
    begin = 0
    for i, (nF, nT, nBp) in enumerate(cfg["alpha"]):
        F = _cdf_axis(nF)
        Tp = _temperatures(nT)
        Sa, C, M = _alpha_partition(F, Bg[begin:begin + nBp], Tp, R)
        put(f"alpha_{i}_CDF", [F, np.arange(R, dtype=np.float64)], C)
        put(f"alpha_{i}_S", [np.arange(R, dtype=np.float64)], Sa)
        put(f"alpha_{i}_beta_T", [Bg[begin:begin + nBp], Tp, np.arange(R, dtype=np.float64)], M)
        begin += nBp
    paths["_n_partitions"] = (len(cfg["beta"]), len(cfg["alpha"]))
    return paths


# ------------------------------------------------------------------------------------------------ decks
def _tables(table_dir):
    """Paths by logical name WITHOUT requiring the files to exist (decks can name missing files on purpose)."""
    d = Path(table_dir)
    return lambda name: os.fspath(d / f"{name}.mmctab")


def _hydrogen(table_dir, T_eval_tag="623K", T_eval=623.6, tsl=True, name="hydrogen in water"):
    t = _tables(table_dir)
    s = f'    <nuclide name={quoteattr(name)} awr="{AWR_H}">\n      <neutron>\n'
    s += f'        <total file="{t("H1_total_" + T_eval_tag)}" temperature="{T_eval}"/>\n'
    s += f'        <capture file="{t("H1_gamma")}" temperature="{T_eval}"/>\n'
    s += f'        <scatter>\n          <xs file="{t("H1_elastic_" + T_eval_tag)}" temperature="{T_eval}"/>\n'
    if tsl:
        s += (f'          <tsl majorant="{t("majorant")}" total_T="{t("scatter_xs_T")}" total_S="{t("scatter_xs_S")}" '
              f'total_E="{t("scatter_xs_E")}" beta_cutoff="{BETA_CUTOFF}" alpha_cutoff="{ALPHA_CUTOFF}">\n')
        s += "            <beta_partitions>\n"
        for i in range(4):
            s += (f'              <partition CDF="{t(f"beta_{i}_CDF")}" S="{t(f"beta_{i}_S")}" '
                  f'E_T="{t(f"beta_{i}_E_T")}"/>\n')
        s += "            </beta_partitions>\n            <alpha_partitions>\n"
        for i in range(4):
            s += (f'              <partition CDF="{t(f"alpha_{i}_CDF")}" S="{t(f"alpha_{i}_S")}" '
                  f'beta_T="{t(f"alpha_{i}_beta_T")}"/>\n')
        s += "            </alpha_partitions>\n          </tsl>\n"
    s += "        </scatter>\n      </neutron>\n    </nuclide>\n"
    return s


def _simple_nuclide(table_dir, name, prefix, awr, fission=False, T_eval=293.6):
    t = _tables(table_dir)
    s = f'    <nuclide name={quoteattr(name)} awr="{awr}">\n      <neutron>\n'
    s += f'        <total file="{t(prefix + "_total")}" temperature="{T_eval}"/>\n'
    s += f'        <capture file="{t(prefix + "_gamma")}" temperature="{T_eval}"/>\n'
    s += f'        <scatter>\n          <xs file="{t(prefix + "_elastic")}" temperature="{T_eval}"/>\n        </scatter>\n'
    if fission:
        s += (f'        <fission>\n          <xs file="{t(prefix + "_fission")}" temperature="{T_eval}"/>\n'
              f'          <nubar file="{t(prefix + "_nubar")}" temperature="{T_eval}"/>\n        </fission>\n')
    s += "      </neutron>\n    </nuclide>\n"
    return s


def _general(histories, threads, seed, tracking):
    s = f"<general>\n  <particles>neutron</particles>\n  <histories>{histories}</histories>\n"
    s += f"  <threads>{threads}</threads>\n  <chunksize>100000</chunksize>\n"
    if seed is not None:
        s += f"  <seed>{seed}</seed>\n"
    if tracking is not None:
        s += f"  <tracking>{tracking}</tracking>\n"
    return s + "</general>\n"


def _source(energy, position=(0, 0, 0), direction=(1, 0, 0)):
    d = "<isotropic/>" if direction == "isotropic" else '<constant x="%s" y="%s" z="%s"/>' % tuple(direction)
    return ("<problemtype>\n  <fixedsource>\n"
            f'    <position>\n      <constant x="{position[0]}" y="{position[1]}" z="{position[2]}"/>\n    </position>\n'
            f"    <direction>\n      {d}\n    </direction>\n"
            f'    <energy>\n      <constant energy="{energy!r}"/>\n    </energy>\n'
            '    <particletype>\n      <constant type="neutron"/>\n    </particletype>\n'
            "  </fixedsource>\n</problemtype>\n")


def energy_boundaries(n: int, lo=1.388794e-11, hi=2.3e-6) -> list:
    """n log-spaced leakage-spectrum bin edges from lo to about hi, like the reference's benchmark decks (201 edges
    1.39e-11 .. 2.26e-6 MeV, ~6 % apart).  The ratio comes from a bisection in exact arithmetic (no pow)."""
    target = hi / lo
    low, high = 1.0, 1.0e3
    for _ in range(100):  # bisection on ratio^(n-1) = target
        ratio = 0.5 * (low + high)
        p = 1.0
        for _ in range(n - 1):
            p = p * ratio
        if p < target:
            low = ratio
        else:
            high = ratio
    return [float(v) for v in _geometric(lo, ratio, n)]


def _current(name, surface, energy_bounds=None, cosine=None):
    s = f'  <current name={quoteattr(name)} surface={quoteattr(surface)}>\n    <bins>\n'
    if cosine is not None:
        (u, v, w), lo, hi, bins = cosine
        s += (f'      <cosine u="{u}" v="{v}" w="{w}">\n        <linspace min="{lo}" max="{hi}" bins="{bins}"/>\n'
              "      </cosine>\n")
    if energy_bounds is not None:
        s += "      <energy>\n        <boundaries>" + " ".join(repr(b) for b in energy_bounds) + "</boundaries>\n      </energy>\n"
    return s + "    </bins>\n  </current>\n"


def slab_deck(table_dir, *, histories=1000, threads=1, seed=None, tracking=None, temperature=450.0, thickness=4.33,
              T_eval=("293K", 293.6), energy=0.56e-6, n_energy_bins=40) -> str:
    """benchmarks/single_zone.xml: one slab of H-in-H2O at a constant cell temperature, surface tracking by default,
    leakage spectrum on the far plane.  T_eval 293.6 K < 450 K makes the tabulated total invalid (Q5): total = sum of
    reactions, free-gas adjusted above the TSL cutoff."""
    s = "<minimc>\n" + _general(histories, threads, seed, tracking)
    s += "<nuclides>\n  <continuous>\n" + _hydrogen(table_dir, *T_eval) + "  </continuous>\n</nuclides>\n"
    s += '<materials>\n  <material name="hydrogen in water" aden="0.066854">\n    <nuclide name="hydrogen in water" afrac="1"/>\n  </material>\n</materials>\n'
    s += f'<surfaces>\n  <planex name="plane-1" x="-1e-6"/>\n  <planex name="plane-2" x="{thickness}"/>\n</surfaces>\n'
    s += ('<cells>\n  <void>\n    <surface name="plane-1" sense="-1"/>\n  </void>\n'
          f'  <cell name="segment-1" material="hydrogen in water" temperature="{temperature}">\n'
          '    <surface name="plane-1" sense="+1"/>\n    <surface name="plane-2" sense="-1"/>\n  </cell>\n'
          '  <void>\n    <surface name="plane-2" sense="+1"/>\n  </void>\n</cells>\n')
    s += _source(energy)
    bounds = energy_boundaries(n_energy_bins)
    s += "<estimators>\n" + _current("leakage", "plane-2", bounds, cosine=((1, 0, 0), 0, 1, 4))
    s += _current("reflected", "plane-1", bounds) + "</estimators>\n</minimc>\n"
    return s


def sensitivity_slab_deck(table_dir, *, histories=1000, threads=1, seed=None, tracking=None) -> str:
    """N4: the single-zone slab with the sensitivity of both currents to the total cross section of
    hydrogen in water (continuous energy + thermal scattering; with `cell delta` tracking the indirect effect uses
    the majorant)."""
    s = slab_deck(table_dir, histories=histories, threads=threads, seed=seed, tracking=tracking, n_energy_bins=8)
    s = s.replace("    <bins>\n", '    <sensitivities>\n      <perturbation name="h-total"/>\n    </sensitivities>\n    <bins>\n')
    return s.replace("</estimators>\n", '</estimators>\n<perturbations>\n  <total name="h-total" nuclide="hydrogen in water"/>\n</perturbations>\n')


def single_zone_benchmark_deck(table_dir, *, histories=1000000, threads=8, seed=None) -> str:
    """benchmarks/single_zone.xml as shipped (BASELINE configs[1]) with the table paths rewritten: 5 cm slab, global
    constant temperature 450 K, data evaluated at 293.6 K, surface tracking, one `current` estimator on the right
    plane with 101 linspace cosine bins x 99 linspace energy bins (103 x 101 = 10 403 bins with the end bins)."""
    s = "<minimc>\n" + _general(histories, threads, seed, "surface").replace("<chunksize>100000", "<chunksize>10000")
    s += "<nuclides>\n  <continuous>\n" + _hydrogen(table_dir, "293K", 293.6) + "  </continuous>\n</nuclides>\n"
    s += '<materials>\n  <material name="hydrogen in water" aden="0.066854">\n    <nuclide name="hydrogen in water" afrac="1"/>\n  </material>\n</materials>\n'
    s += '<surfaces>\n  <planex name="left-plane" x="-1e-6"/>\n  <planex name="right-plane" x="5.0"/>\n</surfaces>\n'
    s += ('<cells>\n  <void>\n    <surface name="left-plane" sense="-1"/>\n  </void>\n'
          '  <cell name="segment" material="hydrogen in water">\n'
          '    <surface name="left-plane" sense="+1"/>\n    <surface name="right-plane" sense="-1"/>\n  </cell>\n'
          '  <void>\n    <surface name="right-plane" sense="+1"/>\n  </void>\n</cells>\n')
    s += _source(0.56e-6)
    s += '<temperature>\n  <constant c="450.0"/>\n</temperature>\n'
    s += ('<estimators>\n  <current name="leakage" surface="right-plane">\n    <bins>\n'
          '      <cosine u="1.0" v="0.0" w="0.0">\n        <linspace min="0" max="1.01" bins="101" />\n      </cosine>\n'
          '      <energy>\n        <linspace min="1e-11" max="0.8e-6" bins="99" />\n      </energy>\n'
          '    </bins>\n  </current>\n</estimators>\n</minimc>\n')
    return s


def multi_zone_deck(table_dir, *, histories=1000, threads=1, seed=None, tracking="surface", n_energy_bins=40) -> str:
    """benchmarks/multi_zone.xml: 13 slab segments at 300..600 K between 14 planes, tabulated data at 623.6 K."""
    planes = [-1e-6, 0.33, 0.67, 1.00, 1.33, 1.67, 2.00, 2.33, 2.67, 3.00, 3.33, 3.67, 4.00, 4.33]
    temps = [300, 323.6, 350.0, 373.6, 400.0, 423.6, 450.0, 473.6, 500.0, 523.6, 550.0, 573.6, 600.0]
    s = "<minimc>\n" + _general(histories, threads, seed, tracking)
    s += "<nuclides>\n  <continuous>\n" + _hydrogen(table_dir, "623K", 623.6) + "  </continuous>\n</nuclides>\n"
    s += '<materials>\n  <material name="hydrogen in water" aden="0.066854">\n    <nuclide name="hydrogen in water" afrac="1"/>\n  </material>\n</materials>\n'
    s += "<surfaces>\n" + "".join(f'  <planex name="plane-{i + 1}" x="{x}"/>\n' for i, x in enumerate(planes)) + "</surfaces>\n"
    s += '<cells>\n  <void>\n    <surface name="plane-1" sense="-1"/>\n  </void>\n'
    for i, T in enumerate(temps):
        s += (f'  <cell name="segment-{i + 1}" material="hydrogen in water" temperature="{T}">\n'
              f'    <surface name="plane-{i + 1}" sense="+1"/>\n    <surface name="plane-{i + 2}" sense="-1"/>\n  </cell>\n')
    s += '  <void>\n    <surface name="plane-14" sense="+1"/>\n  </void>\n</cells>\n'
    s += _source(0.56e-6)
    bounds = energy_boundaries(n_energy_bins)
    s += "<estimators>\n" + _current("leakage", "plane-14", bounds) + "</estimators>\n</minimc>\n"
    return s


def continuous_temperature_deck(table_dir, *, histories=1000, threads=1, seed=None, n_energy_bins=40) -> str:
    """benchmarks/continuous_temperature.xml: one slab, cell delta tracking, T(x) = 300 + 69.28406467 x with declared
    bounds [300, 600], tabulated data at 623.6 K."""
    s = "<minimc>\n" + _general(histories, threads, seed, "cell delta")
    s += "<nuclides>\n  <continuous>\n" + _hydrogen(table_dir, "623K", 623.6) + "  </continuous>\n</nuclides>\n"
    s += '<materials>\n  <material name="hydrogen in water" aden="0.066854">\n    <nuclide name="hydrogen in water" afrac="1"/>\n  </material>\n</materials>\n'
    s += '<surfaces>\n  <planex name="plane-1" x="-1e-6"/>\n  <planex name="plane-14" x="4.33"/>\n</surfaces>\n'
    s += ('<cells>\n  <void>\n    <surface name="plane-1" sense="-1"/>\n  </void>\n'
          '  <cell name="segment-1" material="hydrogen in water">\n'
          '    <surface name="plane-1" sense="+1"/>\n    <surface name="plane-14" sense="-1"/>\n  </cell>\n'
          '  <void>\n    <surface name="plane-14" sense="+1"/>\n  </void>\n</cells>\n')
    s += _source(0.56e-6)
    s += ('<temperature>\n  <linear>\n    <bounds lower="300" upper="600"/>\n    <intercept b="300"/>\n'
          '    <gradient x="69.28406467" y="0" z="0"/>\n  </linear>\n</temperature>\n')
    bounds = energy_boundaries(n_energy_bins)
    s += "<estimators>\n" + _current("leakage", "plane-14", bounds) + "</estimators>\n</minimc>\n"
    return s


def mixed_temperature_deck(table_dir, *, histories=1000, threads=1, seed=None, n_energy_bins=40) -> str:
    """Three slab segments under cell delta tracking: the outer two carry their own constant temperature
    (Cell::AssignTemperature, Cell.cpp:107-113: a ConstantField), the middle one falls back to the global linear field
    of continuous_temperature.xml -- evaluated S(a,b) rows (one load) and two-row reconstructions (interpolated in T at
    every call) inside one world."""
    s = "<minimc>\n" + _general(histories, threads, seed, "cell delta")
    s += "<nuclides>\n  <continuous>\n" + _hydrogen(table_dir, "623K", 623.6) + "  </continuous>\n</nuclides>\n"
    s += '<materials>\n  <material name="hydrogen in water" aden="0.066854">\n    <nuclide name="hydrogen in water" afrac="1"/>\n  </material>\n</materials>\n'
    s += ('<surfaces>\n  <planex name="plane-1" x="-1e-6"/>\n  <planex name="plane-2" x="1.4"/>\n'
          '  <planex name="plane-3" x="3.0"/>\n  <planex name="plane-4" x="4.33"/>\n</surfaces>\n')
    s += ('<cells>\n  <void>\n    <surface name="plane-1" sense="-1"/>\n  </void>\n'
          '  <cell name="segment-1" material="hydrogen in water" temperature="350">\n'
          '    <surface name="plane-1" sense="+1"/>\n    <surface name="plane-2" sense="-1"/>\n  </cell>\n'
          '  <cell name="segment-2" material="hydrogen in water">\n'
          '    <surface name="plane-2" sense="+1"/>\n    <surface name="plane-3" sense="-1"/>\n  </cell>\n'
          '  <cell name="segment-3" material="hydrogen in water" temperature="525.5">\n'
          '    <surface name="plane-3" sense="+1"/>\n    <surface name="plane-4" sense="-1"/>\n  </cell>\n'
          '  <void>\n    <surface name="plane-4" sense="+1"/>\n  </void>\n</cells>\n')
    s += _source(0.56e-6)
    # A gentle gradient: a particle that grazes a plane (|dx| * nudge below half an ulp of the plane's x: the reference's
    # absolute nudge, Constants.hpp:15, does not move it across) stays in its Cell while it wanders through its
    # neighbours; with the 69.28 K/cm of continuous_temperature.xml T(x) then leaves the tables' temperature grid on
    # the left and the reference's SampleBeta throws (std::terminate).  5 K/cm keeps T inside [273.6, 800] K for |x| < 29 cm.
    s += ('<temperature>\n  <linear>\n    <bounds lower="300" upper="600"/>\n    <intercept b="420"/>\n'
          '    <gradient x="5.0" y="0" z="0"/>\n  </linear>\n</temperature>\n')
    bounds = energy_boundaries(n_energy_bins)
    s += "<estimators>\n" + _current("leakage", "plane-4", bounds) + _current("reflected", "plane-1", bounds) + "</estimators>\n</minimc>\n"
    return s


def broomstick_deck(table_dir, *, histories=1000, threads=1, seed=None, temperature=450.0, n_energy_bins=24,
                    n_cosine_bins=12) -> str:
    """benchmarks/broomstick.xml: a cylinder of radius 1e-6 along x between two planes; a particle born on the axis
    collides at most once before leaving through the cylinder (double-differential scattering kernel on cosine x energy
    bins)."""
    s = "<minimc>\n" + _general(histories, threads, seed, "surface")
    s += "<nuclides>\n  <continuous>\n" + _hydrogen(table_dir, "293K", 293.6) + "  </continuous>\n</nuclides>\n"
    s += '<materials>\n  <material name="hydrogen in water" aden="0.066854">\n    <nuclide name="hydrogen in water" afrac="1"/>\n  </material>\n</materials>\n'
    s += ('<surfaces>\n  <cylinderx name="stick" r="1e-6"/>\n  <planex name="back" x="-1e-6"/>\n'
          '  <planex name="front" x="1e4"/>\n</surfaces>\n')
    s += ('<cells>\n'
          f'  <cell name="broomstick" material="hydrogen in water" temperature="{temperature}">\n'
          '    <surface name="stick" sense="-1"/>\n    <surface name="back" sense="+1"/>\n    <surface name="front" sense="-1"/>\n  </cell>\n'
          '  <void>\n    <surface name="stick" sense="+1"/>\n  </void>\n'
          '  <void>\n    <surface name="back" sense="-1"/>\n  </void>\n'
          '  <void>\n    <surface name="front" sense="+1"/>\n  </void>\n</cells>\n')
    s += _source(0.56e-6)
    bounds = energy_boundaries(n_energy_bins)
    s += "<estimators>\n" + _current("kernel", "stick", bounds, cosine=((1, 0, 0), -1, 1, n_cosine_bins)) + "</estimators>\n</minimc>\n"
    return s


def free_gas_sphere_deck(table_dir, *, histories=1000, threads=1, seed=None, tracking=None, energy=1.0e-3,
                         pellet_temperature=900) -> str:
    """A test/continuous.xml-like problem without thermal scattering data: concentric spheres of hydrogen (free gas
    always, awr < 1), an oxygen-like scatterer (free gas only below 500 kT / awr) and a fissile heavy nuclide
    (ContinuousFission with a nubar table), isotropic source.  Cell temperatures above and below T_eval exercise both
    branches of ContinuousEvaluation::IsValid (Q5)."""
    s = "<minimc>\n" + _general(histories, threads, seed, tracking)
    s += "<nuclides>\n  <continuous>\n"
    s += _hydrogen(table_dir, "293K", 293.6, tsl=False, name="hydrogen")
    s += _simple_nuclide(table_dir, "oxygen", "O", 15.8575107)
    s += _simple_nuclide(table_dir, "uranium235", "U", 233.024791, fission=True)
    s += "  </continuous>\n</nuclides>\n"
    s += ('<materials>\n  <material name="water" aden="0.1">\n    <nuclide name="hydrogen" afrac="0.67"/>\n'
          '    <nuclide name="oxygen" afrac="0.33"/>\n  </material>\n'
          '  <material name="fuel" aden="0.02">\n    <nuclide name="uranium235" afrac="0.2"/>\n'
          '    <nuclide name="oxygen" afrac="0.8"/>\n  </material>\n</materials>\n')
    s += ('<surfaces>\n  <sphere name="pellet">\n    <center x="0" y="0" z="0"/>\n    <radius r="1.5"/>\n  </sphere>\n'
          '  <sphere name="moderator">\n    <center x="0" y="0" z="0"/>\n    <radius r="6"/>\n  </sphere>\n</surfaces>\n')
    s += (f'<cells>\n  <cell name="pellet" material="fuel" temperature="{pellet_temperature}">\n    <surface name="pellet" sense="-1"/>\n  </cell>\n'
          '  <cell name="moderator" material="water" temperature="280">\n    <surface name="pellet" sense="+1"/>\n'
          '    <surface name="moderator" sense="-1"/>\n  </cell>\n'
          '  <void>\n    <surface name="moderator" sense="+1"/>\n  </void>\n</cells>\n')
    s += _source(energy, direction="isotropic")
    bounds = energy_boundaries(16, lo=1e-10, hi=2.0)
    s += "<estimators>\n" + _current("leakage", "moderator", bounds) + _current("interface", "pellet") + "</estimators>\n</minimc>\n"
    return s


def thermal_fissile_sphere_deck(table_dir, **kw) -> str:
    """free_gas_sphere_deck with a thermal (0.0253 eV) source in the fuel pellet: ContinuousFission::Interact
    (ContinuousReaction.cpp:252-265) happens within the first histories, so the golden traces hold fission events and
    banked secondaries (FixedSource.cpp:63-72).  Both cells are colder than 1.01 x the evaluations' 293.6 K, so
    ContinuousEvaluation::IsValid holds (Q5) and no cross section takes the erf/exp free-gas adjustment
    (ContinuousReaction.cpp:225-238): the deck is bit-exact under both tracking modes, free-gas scatters included."""
    kw.setdefault("energy", 2.53e-8)
    kw.setdefault("pellet_temperature", 290)
    return free_gas_sphere_deck(table_dir, **kw)


def fissile_sphere_keigenvalue_deck(table_dir, *, histories=20000, threads=1, inactive=3, active=8, tracking=None,
                                    pellet_temperature=290, pellet_radius=12.0) -> str:
    """A continuous-energy k-eigenvalue problem (ContinuousFission::Interact in generation mode,
    ContinuousReaction.cpp:252-265): the fuel / moderator spheres of thermal_fissile_sphere_deck with a larger pellet,
    isotropic thermal initial source at the centre.  The reference's KEigenvalue::Solve is a stub, so there is no
    reference result: the analog k (sites banked per source) and the collision estimator of k must agree."""
    text = free_gas_sphere_deck(table_dir, histories=histories, threads=threads, tracking=tracking, energy=2.53e-8,
                                pellet_temperature=pellet_temperature)
    text = text.replace('<radius r="1.5"/>', f'<radius r="{pellet_radius}"/>').replace('<radius r="6"/>', f'<radius r="{pellet_radius + 6}"/>')
    text = text.replace("  <fixedsource>\n", f'  <keigenvalue inactive="{inactive}" active="{active}">\n  <initialsource>\n')
    return text.replace("  </fixedsource>\n", "  </initialsource>\n  </keigenvalue>\n")


CE_DECKS = {
    "single_zone": slab_deck,
    "multi_zone": multi_zone_deck,
    "continuous_temperature": continuous_temperature_deck,
    "broomstick": broomstick_deck,
    "free_gas_sphere": free_gas_sphere_deck,
    "thermal_fissile_sphere": thermal_fissile_sphere_deck,
}
