"""ORACLE TEST INFRASTRUCTURE -- not product code.

ctypes binding of oracle/liboracle_port.so (oracle/port.h) for tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "liboracle_port.so"
REF_HARNESS = HERE / "_ref" / "ref_harness"

_pd, _pi, _pu = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint32)


class OrcWorld(C.Structure):
    _fields_ = [
        ("n_surfaces", C.c_int32), ("surface_type", _pi), ("surface_param", _pd),
        ("n_cells", C.c_int32), ("cell_material", _pi), ("cell_surface_begin", _pi), ("cell_surface_index", _pi),
        ("cell_surface_sense", _pi),
        ("n_materials", C.c_int32), ("material_aden", _pd), ("material_nuclide_begin", _pi),
        ("material_nuclide_index", _pi), ("material_nuclide_afrac", _pd),
        ("n_nuclides", C.c_int32), ("n_groups", C.c_int32), ("mg_reaction_mask", _pu), ("mg_total", _pd),
        ("mg_capture", _pd), ("mg_scatter", _pd), ("mg_fission", _pd), ("mg_nubar", _pd), ("mg_scatter_probs", _pd),
        ("mg_chi", _pd),
    ]


class OrcSource(C.Structure):
    _fields_ = [("position", C.c_double * 3), ("direction_kind", C.c_int32), ("direction", C.c_double * 3),
                ("group", C.c_uint64)]


class OrcBins(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_bins", C.c_uint64), ("lower", C.c_double), ("upper", C.c_double),
                ("width", C.c_double), ("base", C.c_double), ("boundaries", _pd)]


class OrcEstimator(C.Structure):
    _fields_ = [("surface", C.c_int32), ("has_cosine_direction", C.c_int32), ("cosine_direction", C.c_double * 3),
                ("cosine", OrcBins), ("energy", OrcBins)]


class OrcCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "n_histories", "n_births", "n_events", "n_collisions", "n_crossings", "n_virtual", "n_scores",
        "n_secondaries", "n_lost", "n_physics_errors")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class OrcRecord(C.Structure):
    _fields_ = [("history", C.c_uint64), ("particle", C.c_uint32), ("event", C.c_int32), ("group", C.c_uint64),
                ("cell", C.c_int32), ("surface", C.c_int32), ("position", C.c_double * 3),
                ("direction", C.c_double * 3), ("rng_state", C.c_uint64)]


_lib = None


def build():
    subprocess.run(["make", "-s", "-C", os.fspath(HERE), "port"], check=True)


def load():
    global _lib
    if _lib is None:
        if not LIB.exists():
            build()
        lib = C.CDLL(os.fspath(LIB))
        lib.orc_fixed_source_run.restype = C.c_int
        lib.orc_fixed_source_run.argtypes = [
            C.POINTER(OrcWorld), C.POINTER(OrcSource), C.POINTER(OrcEstimator), C.c_int32, C.c_uint64, C.c_uint64,
            C.c_uint64, C.c_int32, C.c_int32, _pd, _pd, C.POINTER(OrcCounters)]
        lib.orc_trace.restype = C.c_size_t
        lib.orc_trace.argtypes = [C.POINTER(OrcWorld), C.POINTER(OrcSource), C.c_uint64, C.c_uint64, C.c_uint64,
                                  C.c_int32, C.POINTER(OrcRecord), C.c_size_t]
        lib.orc_keigenvalue_run.restype = C.c_int
        lib.orc_keigenvalue_run.argtypes = [
            C.POINTER(OrcWorld), C.POINTER(OrcSource), C.POINTER(OrcEstimator), C.c_int32, C.c_uint64, C.c_uint64,
            C.c_uint64, C.c_int32, _pd, _pd, _pd, C.POINTER(C.c_uint64), C.POINTER(OrcCounters), _pd]
        lib.orc_rng_canonical.restype = None
        lib.orc_rng_canonical.argtypes = [C.c_uint64, C.c_size_t, _pd, C.POINTER(C.c_uint64)]
        _lib = lib
    return _lib


_CT = {np.dtype(np.int32): C.c_int32, np.dtype(np.uint32): C.c_uint32, np.dtype(np.float64): C.c_double}


class Problem:
    """Oracle-side view of a flattened deck (the dict oracle/flatten.py returns)."""

    def __init__(self, flat: dict):
        self.flat = flat
        w = flat["world"]
        self._keep = {}
        ow = OrcWorld()
        for name, _ in OrcWorld._fields_:
            if name.startswith("n_"):
                continue
            a = np.ascontiguousarray(w[name])
            self._keep[name] = a
            setattr(ow, name, a.ctypes.data_as(C.POINTER(_CT[a.dtype])))
        ow.n_surfaces = len(w["surface_type"])
        ow.n_cells = len(w["cell_material"])
        ow.n_materials = len(w["material_aden"])
        ow.n_nuclides = len(w["mg_reaction_mask"])
        ow.n_groups = w["n_groups"]
        self.world = ow
        s = flat["source"]
        src = OrcSource()
        src.position = (C.c_double * 3)(*s["position"])
        src.direction_kind = s["direction_kind"]
        src.direction = (C.c_double * 3)(*s["direction"])
        src.group = s["group"]
        self.source = src
        specs = flat["estimators"]
        self.estimators = (OrcEstimator * max(len(specs), 1))()
        self.n_estimators = len(specs)
        self.total_bins = 0
        for i, e in enumerate(specs):
            oe = self.estimators[i]
            oe.surface = e["surface"]
            oe.has_cosine_direction = 0 if e["cosine_direction"] is None else 1
            if e["cosine_direction"] is not None:
                oe.cosine_direction = (C.c_double * 3)(*e["cosine_direction"])
            for axis in ("cosine", "energy"):
                b, ob = e[axis], getattr(oe, axis)
                ob.kind, ob.n_bins = b["kind"], b["n_bins"]
                ob.lower, ob.upper = b.get("lower", 0.0), b.get("upper", 0.0)
                ob.width, ob.base = b.get("width", 0.0), b.get("base", 10.0)
                if b["kind"] == 3:
                    arr = np.ascontiguousarray(b["boundaries"], np.float64)
                    self._keep[(i, axis)] = arr
                    ob.boundaries = arr.ctypes.data_as(_pd)
            self.total_bins += e["n_bins"]

    def run(self, n_histories=None, *, seed0=None, first=0, tracking=None, threads=1):
        run = self.flat["run"]
        n = run["histories"] if n_histories is None else n_histories
        seed0 = run["seed"] if seed0 is None else seed0
        tracking = run["tracking"] if tracking is None else tracking
        scores = np.zeros(max(self.total_bins, 1))
        squares = np.zeros(max(self.total_bins, 1))
        counters = OrcCounters()
        status = load().orc_fixed_source_run(
            C.byref(self.world), C.byref(self.source), self.estimators, self.n_estimators, seed0, first, n, tracking,
            threads, scores.ctypes.data_as(_pd), squares.ctypes.data_as(_pd), C.byref(counters))
        return scores[:self.total_bins], squares[:self.total_bins], counters.as_dict(), status

    def keigenvalue(self, batchsize=None, inactive=None, active=None, tracking=None):
        """The power iteration of DESIGN.md "k-eigenvalue" on the CPU: (scores, squares, k per cycle, bank sizes,
        counters, status)."""
        run = self.flat["run"]
        n = run["histories"] if batchsize is None else batchsize
        inactive = run["inactive"] if inactive is None else inactive
        active = run["active"] if active is None else active
        tracking = run["tracking"] if tracking is None else tracking
        scores = np.zeros(max(self.total_bins, 1))
        squares = np.zeros(max(self.total_bins, 1))
        k = np.zeros(inactive + active)
        sizes = np.zeros(inactive + active, np.uint64)
        counters = OrcCounters()
        self.k_collision = np.zeros(inactive + active)  # the collision estimator of k, per cycle
        status = load().orc_keigenvalue_run(
            C.byref(self.world), C.byref(self.source), self.estimators, self.n_estimators, n, inactive, active,
            tracking, scores.ctypes.data_as(_pd), squares.ctypes.data_as(_pd), k.ctypes.data_as(_pd),
            sizes.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(counters), self.k_collision.ctypes.data_as(_pd))
        return scores[:self.total_bins], squares[:self.total_bins], k, sizes, counters.as_dict(), status

    def trace(self, first, n, *, seed0=None, tracking=None, cap=1 << 16):
        run = self.flat["run"]
        seed0 = run["seed"] if seed0 is None else seed0
        tracking = run["tracking"] if tracking is None else tracking
        records = (OrcRecord * cap)()
        count = load().orc_trace(C.byref(self.world), C.byref(self.source), seed0, first, n, tracking, records, cap)
        if count > cap:
            raise RuntimeError(f"trace needs {count} records")
        return [records[i] for i in range(count)]


def rng_canonical(seed: int, n: int):
    u = np.zeros(n)
    state = np.zeros(n, np.uint64)
    load().orc_rng_canonical(seed, n, u.ctypes.data_as(_pd), state.ctypes.data_as(C.POINTER(C.c_uint64)))
    return u, state


# --------------------------------------------------------------- reference
def ref_available() -> bool:
    return REF_HARNESS.exists() and os.access(REF_HARNESS, os.X_OK)


def ref_run(deck_path) -> tuple[str, float]:
    """Runs the reference's own Driver::Create(path)->Solve(); returns (.out text, solve seconds)."""
    p = subprocess.run([os.fspath(REF_HARNESS), "run", os.fspath(deck_path)], capture_output=True, text=True,
                       check=True)
    seconds = float(p.stderr.strip().split("solve_seconds=")[-1])
    return p.stdout, seconds


def ref_trace(deck_path, first: int, count: int):
    p = subprocess.run([os.fspath(REF_HARNESS), "trace", os.fspath(deck_path), str(first), str(count)],
                       capture_output=True, text=True, check=True)
    return parse_trace(p.stdout)


def ref_dump(deck_path) -> str:
    return subprocess.run([os.fspath(REF_HARNESS), "dump", os.fspath(deck_path)], capture_output=True, text=True,
                          check=True).stdout


def parse_trace(text: str):
    """Parses ref_harness `trace` lines into dicts comparable with OrcRecord / mmc_event_record."""
    out = []
    for line in text.splitlines():
        f = line.split()
        if not f or f[0] not in ("B", "E"):
            continue
        out.append({
            "tag": f[0], "history": int(f[1]), "particle": int(f[2]), "event": int(f[3]),
            "group": int(f[4][1:]) if f[4][0] == "g" else None,
            "energy": float.fromhex(f[4][1:]) if f[4][0] == "e" else None,
            "cell": int(f[5]), "surface": int(f[6]),
            "position": tuple(float.fromhex(x) for x in f[7:10]),
            "direction": tuple(float.fromhex(x) for x in f[10:13]),
            "rng_state": int(f[13]),
        })
    return out


def parse_out(text: str):
    """Parses the reference's .out text (minimc.cpp:20-21, Estimator.cpp:48-57, Scorable.cpp:51-70) into
    {estimator name: {"mean": [...], "std dev": [...]}} plus the batch size."""
    lines = text.splitlines()
    batch = int(lines[0])
    result = {}
    i = 1
    while i < len(lines):
        if i + 1 < len(lines) and lines[i] and lines[i + 1] and set(lines[i + 1]) == {"="}:
            name = lines[i]
            entry = {}
            j = i + 2
            while j < len(lines) and not (j + 1 < len(lines) and lines[j] and lines[j + 1] and set(lines[j + 1]) == {"="}):
                if lines[j] in ("mean", "std dev", "cosine", "energy") and j + 2 < len(lines):
                    vals = [v.strip() for v in lines[j + 2].split(",") if v.strip()]
                    entry[lines[j]] = vals
                    j += 3
                else:
                    j += 1
            result[name] = entry
            i = j
        else:
            i += 1
    return batch, result
