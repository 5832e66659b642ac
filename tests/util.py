"""Shared helpers of the test-suite: deck -> oracle problem / product objects."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

from minimc_b200 import capi, decks
from oracle import flatten as oflatten
from oracle import port_py

GOLDEN = Path(__file__).resolve().parent / "golden"
REFERENCE_ROOT = Path("/root/reference")


def flat_from_xml(text: str) -> dict:
    return oflatten.flatten(text, is_text=True)


def oracle_problem(flat: dict) -> port_py.Problem:
    return port_py.Problem(flat)


def product_world(flat: dict) -> capi.World:
    w = flat["world"]
    return capi.World(capi.FlatWorld(**w))


def product_source(flat: dict) -> capi.SourceDesc:
    s = flat["source"]
    return capi.source_desc(position=s["position"], direction_kind=s["direction_kind"], direction=s["direction"],
                            group=s["group"])


def product_estimators(flat: dict) -> capi.Estimators:
    specs = []
    for e in flat["estimators"]:
        def conv(b):
            if b["kind"] == 0:
                return (capi.BINS_NONE, {})
            if b["kind"] in (1, 2):
                return (b["kind"], {"bins": b["n_bins"] - 2, "lower": b["lower"], "upper": b["upper"],
                                    "base": b.get("base", 10.0)})
            return (capi.BINS_BOUNDARIES, {"boundaries": b["boundaries"]})
        specs.append({"surface": e["surface"], "cosine_direction": e["cosine_direction"], "cosine": conv(e["cosine"]),
                      "energy": conv(e["energy"])})
    return capi.Estimators(specs)


def record_tuple(r, with_cell=True):
    """(history, particle, event, group, cell, surface, rng_state) of a ctypes record or parsed dict."""
    if isinstance(r, dict):
        t = (r["history"], r["particle"], r["event"], r["group"], r["cell"], r["surface"], r["rng_state"])
    else:
        t = (int(r.history), int(r.particle), int(r.event), int(r.group), int(r.cell), int(r.surface),
             int(r.rng_state))
    if not with_cell or t[2] == 0:  # the cell of a birth record is not part of the contract
        t = t[:4] + (None,) + t[5:]
    return t


def record_vectors(r):
    if isinstance(r, dict):
        return np.array(r["position"]), np.array(r["direction"])
    return np.array(list(r.position)), np.array(list(r.direction))


def ulp_distance(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Distance in units in the last place between float64 arrays (same sign assumed near zero)."""
    ia = np.ascontiguousarray(a, np.float64).view(np.int64)
    ib = np.ascontiguousarray(b, np.float64).view(np.int64)
    ia = np.where(ia < 0, np.int64(-(2 ** 63)) - ia, ia)
    ib = np.where(ib < 0, np.int64(-(2 ** 63)) - ib, ib)
    return np.abs(ia - ib)


def load_golden(name: str):
    with open(GOLDEN / name) as f:
        return json.load(f)


TRACKING = {"surface": None, "delta": "cell delta"}
DECK_CASES = [(name, trk) for name in decks.DECKS for trk in TRACKING]
TRACE_HISTORIES = 48


def deck_text(name: str, tracking: str, **kw) -> str:
    if name == "three_shells":
        kw.setdefault("estimators", decks.THREE_SHELL_ESTIMATORS)
    return decks.DECKS[name](tracking=TRACKING[tracking], **kw)


def golden_trace(name: str, tracking: str):
    return port_py.parse_trace((GOLDEN / f"{name}__{tracking}.trace").read_text())


def golden_out(name: str, tracking: str):
    return port_py.parse_out((GOLDEN / f"{name}__{tracking}.out").read_text())


def format_out_values(flat: dict, scores, squares):
    """Formats scores the way Scorable::GetScoreAsString does (Scorable.cpp:51-70): mean = score / N,
    std dev = sqrt(square - score^2 / N) / N, both printed with std::scientific (%e)."""
    N = float(flat["run"]["histories"])
    result, off = {}, 0
    for e in flat["estimators"]:
        n = e["n_bins"]
        sc, sq = scores[off:off + n], squares[off:off + n]
        result[e["name"]] = {
            "mean": ["%e" % (v / N) for v in sc],
            "std dev": ["%e" % (np.sqrt(q - v * v / N) / N) for v, q in zip(sc, sq)],
        }
        off += n
    return result


# ---------------------------------------------------------------- continuous-energy cases
CE_TRACE_HISTORIES = 16
CE_HISTORIES = 4000
# decks whose every arithmetic step is restated bit for bit on the device; free_gas_sphere with surface tracking
# reaches the free-gas cross-section adjustment (erf, exp: CUDA's, ulp-level differences) -- see physics_ce.cuh
CE_EXACT = {("single_zone", "surface"), ("single_zone", "delta"), ("multi_zone", "surface"), ("multi_zone", "delta"),
            ("continuous_temperature", "delta"), ("broomstick", "surface"), ("free_gas_sphere", "delta"),
            ("thermal_fissile_sphere", "surface"), ("thermal_fissile_sphere", "delta")}


def ce_cases(table_dir, histories=CE_HISTORIES, seed=None):
    """(deck name, tracking tag, XML text) of every continuous-energy parity case."""
    from minimc_b200 import ce_decks
    cases = []
    for name, fn in ce_decks.CE_DECKS.items():
        for tag, tracking in TRACKING.items():
            kw = {"histories": histories, "threads": 4, "seed": seed}
            if name in ("single_zone", "free_gas_sphere", "thermal_fissile_sphere"):
                kw["tracking"] = tracking
            elif name == "multi_zone":
                kw["tracking"] = tracking or "surface"
            elif name == "continuous_temperature":
                if tag == "surface":
                    continue  # cell delta tracking only (non-constant temperature)
            elif tag == "delta":
                continue
            cases.append((name, tag, fn(table_dir, **kw)))
    return cases


CE_CASE_IDS = [("single_zone", "surface"), ("single_zone", "delta"), ("multi_zone", "surface"), ("multi_zone", "delta"),
               ("continuous_temperature", "delta"), ("broomstick", "surface"), ("free_gas_sphere", "surface"),
               ("free_gas_sphere", "delta"), ("thermal_fissile_sphere", "surface"), ("thermal_fissile_sphere", "delta")]


# ---------------------------------------------------------------- sensitivities
SENS_HISTORIES = 20000


def sensitivity_cases(table_dir, histories=SENS_HISTORIES, seed=None):
    """(deck name, tracking tag, XML text) of every differential-operator sensitivity parity case.  One reference
    worker thread: the reference's real-valued sums then have one fixed order."""
    from minimc_b200 import ce_decks
    cases = []
    for tag, tracking in TRACKING.items():
        cases.append(("sensitivity_shells", tag, decks.sensitivity_shells(histories=histories, threads=1, seed=seed, tracking=tracking)))
        cases.append(("sensitivity_slab", tag, ce_decks.sensitivity_slab_deck(table_dir, histories=histories // 5, threads=1,
                                                                           seed=seed, tracking=tracking)))
    return cases


SENS_CASE_IDS = [(name, tag) for name in ("sensitivity_shells", "sensitivity_slab") for tag in TRACKING]
