"""ctypes binding of include/minimc_b200.h (the C ABI of libminimc_b200.so).

Python is plumbing here: the product is the shared library.  Importing this
module never falls back to a CPU implementation -- if the library is missing,
`load()` raises, and on a box without a GPU every compute call fails with
MMC_ERR_NO_DEVICE.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

ABI_VERSION = 2
LIB_PATH = Path(__file__).resolve().parent / "libminimc_b200.so"

# enums of include/minimc_b200.h
OK, ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_LOST_PARTICLE, ERR_CAPACITY, ERR_PHYSICS = range(7)
SURF_SPHERE, SURF_PLANEX, SURF_CYLINDERX = 0, 1, 2
TRACK_SURFACE, TRACK_CELL_DELTA = 0, 1
EV_BIRTH, EV_SCATTER, EV_CAPTURE, EV_FISSION, EV_SURFACE_CROSS, EV_LEAK, EV_VIRTUAL_COLLISION = range(7)
RNG_MINSTD_COMPAT, RNG_COUNTER = 0, 1
DIR_CONSTANT, DIR_ISOTROPIC, DIR_ISOTROPIC_FLUX = 0, 1, 2
BINS_NONE, BINS_LINSPACE, BINS_LOGSPACE, BINS_BOUNDARIES = 0, 1, 2, 3
REACTION_CAPTURE, REACTION_SCATTER, REACTION_FISSION = 1, 2, 4
FIELD_CONSTANT, FIELD_LINEAR = 0, 1

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int32)
_pu = C.POINTER(C.c_uint32)


class WorldDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("abi_version", C.c_uint32),
        ("n_surfaces", C.c_int32), ("surface_type", _pi), ("surface_param", _pd),
        ("n_cells", C.c_int32), ("cell_material", _pi), ("cell_surface_begin", _pi),
        ("cell_surface_index", _pi), ("cell_surface_sense", _pi), ("cell_field_kind", _pi),
        ("cell_field_param", _pd),
        ("n_materials", C.c_int32), ("material_aden", _pd), ("material_nuclide_begin", _pi),
        ("material_nuclide_index", _pi), ("material_nuclide_afrac", _pd),
        ("n_nuclides", C.c_int32), ("n_groups", C.c_int32),
        ("mg_reaction_mask", _pu), ("mg_total", _pd), ("mg_capture", _pd), ("mg_scatter", _pd),
        ("mg_fission", _pd), ("mg_nubar", _pd), ("mg_scatter_probs", _pd), ("mg_chi", _pd),
        ("ce", C.c_void_p),
    ]


class SourceDesc(C.Structure):
    _fields_ = [
        ("position", C.c_double * 3), ("direction_kind", C.c_int32), ("direction", C.c_double * 3),
        ("group", C.c_uint64), ("energy", C.c_double),
    ]


class BinsDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("n_bins", C.c_uint64), ("lower", C.c_double), ("upper", C.c_double),
        ("width", C.c_double), ("base", C.c_double), ("boundaries", _pd),
    ]


class EstimatorDesc(C.Structure):
    _fields_ = [
        ("surface", C.c_int32), ("has_cosine_direction", C.c_int32), ("cosine_direction", C.c_double * 3),
        ("cosine", BinsDesc), ("energy", BinsDesc),
    ]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "n_histories", "n_births", "n_events", "n_collisions", "n_crossings", "n_virtual", "n_scores",
        "n_secondaries", "n_banked", "n_lost", "n_capacity_overflow", "n_physics_errors")]

    def as_dict(self) -> dict:
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class EventRecord(C.Structure):
    _fields_ = [
        ("history", C.c_uint64), ("particle", C.c_uint32), ("event", C.c_int32), ("group", C.c_uint64),
        ("energy", C.c_double), ("cell", C.c_int32), ("surface", C.c_int32), ("position", C.c_double * 3),
        ("direction", C.c_double * 3), ("rng_state", C.c_uint64),
    ]


class RunOptions(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("device", C.c_int32), ("tracking", C.c_int32), ("rng_mode", C.c_int32),
        ("secondary_capacity", C.c_uint32), ("pending_capacity", C.c_uint32), ("blocks_per_sm", C.c_uint32),
        ("schedule", C.c_uint32), ("stream", C.c_void_p), ("event_slots", C.c_uint32), ("profile", C.c_uint32),
    ]


SCHEDULE_AUTO, SCHEDULE_FUSED, SCHEDULE_EVENT, SCHEDULE_EVENT_ONLY = 0, 1, 2, 3  # mmc_schedule


# every symbol include/minimc_b200.h declares
EXPORTS = (
    "mmc_abi_version", "mmc_last_error", "mmc_device_count", "mmc_world_create", "mmc_world_destroy",
    "mmc_estimator_size", "mmc_fixed_source_run", "mmc_fixed_source_run_sensitivities", "mmc_fixed_source_run_device", "mmc_trace_histories",
    "mmc_test_device_math", "mmc_test_geometry",
    "mmc_source_bank_sample", "mmc_generation_run", "mmc_bank_resample",
    "mmc_device_alloc", "mmc_device_free", "mmc_device_zero", "mmc_device_read", "mmc_device_write",
    # host layer
    "mmc_driver_create", "mmc_driver_create_from_string", "mmc_driver_destroy", "mmc_driver_set_options",
    "mmc_driver_set_shard", "mmc_driver_solve", "mmc_driver_batchsize", "mmc_driver_total_bins", "mmc_driver_scores",
    "mmc_driver_add_scores", "mmc_driver_counters", "mmc_driver_output", "mmc_driver_world_json", "mmc_driver_keff",
    "mmc_driver_trace", "mmc_driver_run_device", "mmc_driver_release_device", "mmc_driver_table_bytes", "mmc_world_bytes",
    "mmc_world_update", "mmc_driver_refresh_device", "mmc_world_last_launches", "mmc_driver_last_launches", "mmc_world_last_kernel_ms", "mmc_driver_last_kernel_ms", "mmc_world_last_boundary_ms", "mmc_driver_last_boundary_ms",
    # multi-GPU
    "mmc_comm_unique_id", "mmc_comm_create", "mmc_comm_destroy", "mmc_comm_rank", "mmc_comm_size", "mmc_nccl_version",
    "mmc_comm_exchange_ms", "mmc_tally_allreduce", "mmc_exchange_plan", "mmc_bank_exchange", "mmc_world_stream",
    "mmc_driver_k_collision", "mmc_driver_cycle_seconds", "mmc_driver_set_comm", "mmc_driver_init_comm_from_environment",
)

COMM_ID_BYTES = 128
K_COLLISION_ONE = float(1 << 28)

_lib = None


class MinimcError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"minimc_b200 status {status}: {message}")
        self.status = status
        self.message = message


def load() -> C.CDLL:
    """Loads libminimc_b200.so; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(minimc_b200 has no CPU or pure-Python transport path)")
    lib = C.CDLL(os.fspath(LIB_PATH))
    lib.mmc_abi_version.restype = C.c_int
    lib.mmc_last_error.restype = C.c_size_t
    lib.mmc_last_error.argtypes = [C.c_char_p, C.c_size_t]
    lib.mmc_device_count.restype = C.c_int
    lib.mmc_world_create.restype = C.c_int
    lib.mmc_world_create.argtypes = [C.POINTER(WorldDesc), C.c_int, C.POINTER(C.c_void_p)]
    lib.mmc_world_destroy.restype = None
    lib.mmc_world_destroy.argtypes = [C.c_void_p]
    lib.mmc_estimator_size.restype = C.c_uint64
    lib.mmc_estimator_size.argtypes = [C.POINTER(EstimatorDesc)]
    run_args = [C.c_void_p, C.POINTER(SourceDesc), C.POINTER(EstimatorDesc), C.c_int32, C.c_uint64, C.c_uint64,
                C.c_uint64, C.POINTER(RunOptions)]
    lib.mmc_fixed_source_run.restype = C.c_int
    lib.mmc_fixed_source_run.argtypes = run_args + [_pd, _pd, C.POINTER(Counters)]
    lib.mmc_fixed_source_run_device.restype = C.c_int
    lib.mmc_fixed_source_run_device.argtypes = run_args + [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mmc_trace_histories.restype = C.c_int
    lib.mmc_trace_histories.argtypes = [
        C.c_void_p, C.POINTER(SourceDesc), C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(RunOptions),
        C.POINTER(EventRecord), C.c_size_t, C.POINTER(C.c_size_t)]
    lib.mmc_test_device_math.restype = C.c_int
    lib.mmc_test_device_math.argtypes = [C.c_int, _pd, _pd, _pd, C.c_size_t]
    lib.mmc_source_bank_sample.restype = C.c_int
    lib.mmc_source_bank_sample.argtypes = [C.c_void_p, C.POINTER(SourceDesc), C.c_uint64, C.c_uint64, C.c_uint64,
                                           C.POINTER(RunOptions), C.c_void_p]
    lib.mmc_generation_run.restype = C.c_int
    lib.mmc_generation_run.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(EstimatorDesc), C.c_int32, C.c_int32,
                                       C.POINTER(RunOptions), C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p]
    lib.mmc_bank_resample.restype = C.c_int
    lib.mmc_bank_resample.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                                      C.c_uint64, C.POINTER(RunOptions), C.c_void_p, C.c_void_p]
    lib.mmc_device_alloc.restype = C.c_int
    lib.mmc_device_alloc.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
    lib.mmc_device_free.restype = None
    lib.mmc_device_free.argtypes = [C.c_void_p, C.c_void_p]
    lib.mmc_device_zero.restype = C.c_int
    lib.mmc_device_zero.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.mmc_device_read.restype = C.c_int
    lib.mmc_device_read.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.mmc_device_write.restype = C.c_int
    lib.mmc_device_write.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.mmc_test_geometry.restype = C.c_int
    lib.mmc_test_geometry.argtypes = [C.c_void_p, C.c_size_t, _pd, _pd, _pi, _pi, _pd]
    lib.mmc_driver_create.restype = C.c_int
    lib.mmc_driver_create.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    lib.mmc_driver_create_from_string.restype = C.c_int
    lib.mmc_driver_create_from_string.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    lib.mmc_driver_destroy.restype = None
    lib.mmc_driver_destroy.argtypes = [C.c_void_p]
    lib.mmc_driver_set_options.restype = C.c_int
    lib.mmc_driver_set_options.argtypes = [C.c_void_p, C.POINTER(RunOptions)]
    lib.mmc_driver_set_shard.restype = C.c_int
    lib.mmc_driver_set_shard.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
    lib.mmc_driver_solve.restype = C.c_int
    lib.mmc_driver_solve.argtypes = [C.c_void_p]
    lib.mmc_driver_batchsize.restype = C.c_uint64
    lib.mmc_driver_batchsize.argtypes = [C.c_void_p]
    lib.mmc_driver_total_bins.restype = C.c_uint64
    lib.mmc_driver_total_bins.argtypes = [C.c_void_p]
    lib.mmc_driver_scores.restype = C.c_int
    lib.mmc_driver_scores.argtypes = [C.c_void_p, _pd, _pd, C.c_uint64]
    lib.mmc_driver_add_scores.restype = C.c_int
    lib.mmc_driver_add_scores.argtypes = [C.c_void_p, _pd, _pd, C.c_uint64]
    lib.mmc_driver_counters.restype = C.c_int
    lib.mmc_driver_counters.argtypes = [C.c_void_p, C.POINTER(Counters)]
    lib.mmc_driver_output.restype = C.c_size_t
    lib.mmc_driver_output.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    lib.mmc_driver_world_json.restype = C.c_size_t
    lib.mmc_driver_world_json.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    lib.mmc_driver_trace.restype = C.c_int
    lib.mmc_driver_trace.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(EventRecord), C.c_size_t,
                                     C.POINTER(C.c_size_t)]
    lib.mmc_driver_run_device.restype = C.c_int
    lib.mmc_driver_run_device.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mmc_driver_release_device.restype = None
    lib.mmc_driver_release_device.argtypes = [C.c_void_p]
    lib.mmc_driver_table_bytes.restype = C.c_uint64
    lib.mmc_driver_table_bytes.argtypes = [C.c_void_p]
    lib.mmc_world_bytes.restype = C.c_uint64
    lib.mmc_world_bytes.argtypes = [C.c_void_p]
    lib.mmc_driver_refresh_device.restype = C.c_int
    lib.mmc_driver_refresh_device.argtypes = [C.c_void_p]
    lib.mmc_world_update.restype = C.c_int
    lib.mmc_world_update.argtypes = [C.c_void_p, C.c_void_p]
    for fn in (lib.mmc_world_last_launches, lib.mmc_driver_last_launches):
        fn.restype = C.c_uint64
        fn.argtypes = [C.c_void_p]
    for fn in (lib.mmc_world_last_boundary_ms, lib.mmc_driver_last_boundary_ms):
        fn.restype = C.c_double
        fn.argtypes = [C.c_void_p]
    for fn in (lib.mmc_world_last_kernel_ms, lib.mmc_driver_last_kernel_ms):
        fn.restype = None
        fn.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.mmc_driver_keff.restype = C.c_int
    lib.mmc_driver_keff.argtypes = [C.c_void_p, _pd, _pd, _pd, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.mmc_driver_k_collision.restype = C.c_int
    lib.mmc_driver_k_collision.argtypes = [C.c_void_p, _pd, _pd, _pd, C.c_size_t, C.POINTER(C.c_size_t), _pd]
    lib.mmc_driver_cycle_seconds.restype = C.c_int
    lib.mmc_driver_cycle_seconds.argtypes = [C.c_void_p, _pd, _pd]
    lib.mmc_driver_set_comm.restype = C.c_int
    lib.mmc_driver_set_comm.argtypes = [C.c_void_p, C.c_void_p]
    lib.mmc_driver_init_comm_from_environment.restype = C.c_int
    lib.mmc_driver_init_comm_from_environment.argtypes = [C.c_void_p]
    # multi-GPU
    lib.mmc_comm_unique_id.restype = C.c_int
    lib.mmc_comm_unique_id.argtypes = [C.c_void_p, C.c_size_t]
    lib.mmc_comm_create.restype = C.c_int
    lib.mmc_comm_create.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
    lib.mmc_comm_destroy.restype = None
    lib.mmc_comm_destroy.argtypes = [C.c_void_p]
    lib.mmc_comm_rank.restype = C.c_int
    lib.mmc_comm_rank.argtypes = [C.c_void_p]
    lib.mmc_comm_size.restype = C.c_int
    lib.mmc_comm_size.argtypes = [C.c_void_p]
    lib.mmc_nccl_version.restype = C.c_int
    lib.mmc_comm_exchange_ms.restype = C.c_double
    lib.mmc_comm_exchange_ms.argtypes = [C.c_void_p]
    lib.mmc_tally_allreduce.restype = C.c_int
    lib.mmc_tally_allreduce.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    _pu64 = C.POINTER(C.c_uint64)
    lib.mmc_exchange_plan.restype = C.c_int
    lib.mmc_exchange_plan.argtypes = [_pu64, C.c_int, C.c_int, C.c_uint64, _pu64, _pu64, _pu64, _pu64]
    lib.mmc_bank_exchange.restype = C.c_int
    lib.mmc_bank_exchange.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64,
                                      _pu64, _pu64, _pu64, _pu64, C.c_void_p]
    lib.mmc_world_stream.restype = C.c_void_p
    lib.mmc_world_stream.argtypes = [C.c_void_p]
    if lib.mmc_abi_version() != ABI_VERSION:
        raise ImportError(f"ABI mismatch: library {lib.mmc_abi_version()} vs binding {ABI_VERSION}")
    _lib = lib
    return lib


def last_error() -> str:
    buf = C.create_string_buffer(1024)
    load().mmc_last_error(buf, len(buf))
    return buf.value.decode()


def check(status: int) -> None:
    if status != OK:
        raise MinimcError(status, last_error())


def _arr(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


class FlatWorld:
    """The flat tables of mmc_world_desc as numpy arrays (see the header for the
    meaning and order of every array).  Built by the host flattener
    (minimc_b200.host) or, in tests, by oracle/flatten.py."""

    FIELDS = {
        "surface_type": np.int32, "surface_param": np.float64, "cell_material": np.int32,
        "cell_surface_begin": np.int32, "cell_surface_index": np.int32, "cell_surface_sense": np.int32,
        "cell_field_kind": np.int32, "cell_field_param": np.float64, "material_aden": np.float64,
        "material_nuclide_begin": np.int32, "material_nuclide_index": np.int32,
        "material_nuclide_afrac": np.float64, "mg_reaction_mask": np.uint32, "mg_total": np.float64,
        "mg_capture": np.float64, "mg_scatter": np.float64, "mg_fission": np.float64, "mg_nubar": np.float64,
        "mg_scatter_probs": np.float64, "mg_chi": np.float64,
    }

    def __init__(self, n_groups: int, **arrays):
        self.n_groups = int(n_groups)
        for name, dtype in self.FIELDS.items():
            setattr(self, name, _arr(arrays[name], dtype).reshape(-1))
        self.n_surfaces = len(self.surface_type)
        self.n_cells = len(self.cell_material)
        self.n_materials = len(self.material_aden)
        self.n_nuclides = len(self.mg_reaction_mask)

    def desc(self) -> WorldDesc:
        d = WorldDesc()
        d.struct_size = C.sizeof(WorldDesc)
        d.abi_version = ABI_VERSION
        d.n_surfaces, d.n_cells = self.n_surfaces, self.n_cells
        d.n_materials, d.n_nuclides, d.n_groups = self.n_materials, self.n_nuclides, self.n_groups
        for name, dtype in self.FIELDS.items():
            ctype = {np.int32: C.c_int32, np.uint32: C.c_uint32, np.float64: C.c_double}[dtype]
            setattr(d, name, _ptr(getattr(self, name), ctype))
        d.ce = None
        return d


def source_desc(position=(0.0, 0.0, 0.0), direction_kind=DIR_ISOTROPIC, direction=(1.0, 0.0, 0.0), group=1,
                energy=0.0) -> SourceDesc:
    s = SourceDesc()
    s.position = (C.c_double * 3)(*position)
    s.direction_kind = direction_kind
    s.direction = (C.c_double * 3)(*direction)
    s.group = group
    s.energy = energy
    return s


def bins_desc(kind=BINS_NONE, *, bins=0, lower=0.0, upper=0.0, base=10.0, boundaries=None):
    """Mirrors the constructors of Bins.cpp:57-70,96-110,139-154.  Returns the
    descriptor and the numpy array that must outlive it."""
    b = BinsDesc()
    b.kind = kind
    keep = None
    if kind == BINS_NONE:
        b.n_bins = 1
    elif kind in (BINS_LINSPACE, BINS_LOGSPACE):
        b.n_bins = bins + 2
        b.lower, b.upper, b.base = lower, upper, base
        b.width = (upper - lower) / bins  # (upper_bound - lower_bound) / (n_bins - 2)
    elif kind == BINS_BOUNDARIES:
        keep = _arr(boundaries, np.float64)
        b.n_bins = len(keep) + 1
        b.boundaries = _ptr(keep, C.c_double)
    else:
        raise ValueError(kind)
    return b, keep


class Estimators:
    """An array of mmc_estimator_desc plus the buffers it points into."""

    def __init__(self, specs):
        # specs: list of dicts {surface, cosine_direction|None, cosine: (kind, kwargs), energy: (kind, kwargs)}
        self._keep = []
        self.array = (EstimatorDesc * max(len(specs), 1))()
        self.n = len(specs)
        self.sizes = []
        for i, s in enumerate(specs):
            e = self.array[i]
            e.surface = s["surface"]
            cd = s.get("cosine_direction")
            e.has_cosine_direction = 1 if cd is not None else 0
            if cd is not None:
                e.cosine_direction = (C.c_double * 3)(*cd)
            ck, ckw = s.get("cosine", (BINS_NONE, {}))
            ek, ekw = s.get("energy", (BINS_NONE, {}))
            e.cosine, k1 = bins_desc(ck, **ckw)
            e.energy, k2 = bins_desc(ek, **ekw)
            self._keep += [k1, k2]
            self.sizes.append(int(e.cosine.n_bins * e.energy.n_bins))
        self.total_bins = sum(self.sizes)


def device_math(fn: int, x):
    """mmc_test_device_math: (out0, out1) evaluated on the GPU."""
    x = _arr(x, np.float64)
    out0, out1 = np.zeros_like(x), np.zeros_like(x)
    check(load().mmc_test_device_math(fn, _ptr(x, C.c_double), _ptr(out0, C.c_double), _ptr(out1, C.c_double),
                                      x.size))
    return out0, out1


class World:
    """Owns an mmc_world handle (the device copy of the tables)."""

    def __init__(self, flat: FlatWorld, device: int = -1):
        self.flat = flat
        self._handle = C.c_void_p()
        desc = flat.desc()
        check(load().mmc_world_create(C.byref(desc), device, C.byref(self._handle)))

    def close(self):
        if self._handle:
            load().mmc_world_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _options(tracking, secondary_capacity, pending_capacity, blocks_per_sm, stream, schedule=0, event_slots=0):
        o = RunOptions()
        o.struct_size = C.sizeof(RunOptions)
        o.device = -1
        o.tracking = tracking
        o.rng_mode = RNG_MINSTD_COMPAT
        o.secondary_capacity = secondary_capacity
        o.pending_capacity = pending_capacity
        o.blocks_per_sm = blocks_per_sm
        o.stream = stream
        o.schedule = schedule
        o.event_slots = event_slots
        return o

    def fixed_source_run(self, source: SourceDesc, estimators: Estimators, seed0: int, first_history: int,
                         n_histories: int, *, tracking=TRACK_SURFACE, secondary_capacity=0, pending_capacity=0,
                         blocks_per_sm=0, scores=None, square_scores=None, schedule=0, event_slots=0):
        """mmc_fixed_source_run with host buffers.  Returns (scores, square_scores, counters dict)."""
        nb = max(estimators.total_bins, 1)
        scores = np.zeros(nb) if scores is None else scores
        square_scores = np.zeros(nb) if square_scores is None else square_scores
        counters = Counters()
        o = self._options(tracking, secondary_capacity, pending_capacity, blocks_per_sm, None, schedule, event_slots)
        check(load().mmc_fixed_source_run(
            self._handle, C.byref(source), estimators.array, estimators.n, seed0, first_history, n_histories,
            C.byref(o), _ptr(scores, C.c_double), _ptr(square_scores, C.c_double), C.byref(counters)))
        return scores[:estimators.total_bins], square_scores[:estimators.total_bins], counters.as_dict()

    def fixed_source_run_device(self, source, estimators, seed0, first_history, n_histories, d_scores, d_square,
                                d_counters, *, tracking=TRACK_SURFACE, secondary_capacity=0, pending_capacity=0,
                                blocks_per_sm=0, stream=None, schedule=0, event_slots=0):
        """mmc_fixed_source_run_device: raw device pointers (ints), on `stream`."""
        o = self._options(tracking, secondary_capacity, pending_capacity, blocks_per_sm, stream, schedule, event_slots)
        check(load().mmc_fixed_source_run_device(
            self._handle, C.byref(source), estimators.array, estimators.n, seed0, first_history, n_histories,
            C.byref(o), d_scores, d_square, d_counters))

    SITE_BYTES = 64  # sizeof(mmc_site)

    def source_bank_sample(self, source, seed0, first_index, n, d_bank, *, stream=None):
        """mmc_source_bank_sample into a device buffer (raw pointer)."""
        o = self._options(TRACK_SURFACE, 0, 0, 0, stream)
        check(load().mmc_source_bank_sample(self._handle, C.byref(source), seed0, first_index, n, C.byref(o), d_bank))

    def generation_run(self, d_bank_in, n_in, estimators, score, d_bank_out, bank_capacity, d_n_out, d_scores, d_square,
                       d_counters, *, tracking=TRACK_SURFACE, secondary_capacity=0, pending_capacity=0, stream=None,
                       d_k_collision=None):
        """mmc_generation_run: raw device pointers (ints), asynchronous on `stream`."""
        o = self._options(tracking, secondary_capacity, pending_capacity, 0, stream)
        check(load().mmc_generation_run(
            self._handle, d_bank_in, n_in, estimators.array, estimators.n, 1 if score else 0, C.byref(o), d_bank_out,
            bank_capacity, d_n_out, d_scores, d_square, d_counters, d_k_collision))

    def bank_resample(self, d_slice, slice_first, slice_n, m_total, n_total, first_out, n_out, d_bank_next, d_errors, *,
                      stream=None):
        o = self._options(TRACK_SURFACE, 0, 0, 0, stream)
        check(load().mmc_bank_resample(self._handle, d_slice, slice_first, slice_n, m_total, n_total, first_out, n_out,
                                       C.byref(o), d_bank_next, d_errors))

    def geometry(self, positions, directions):
        """mmc_test_geometry: (cell, nearest surface, distance) per query point, evaluated on the GPU."""
        pos, dirs = _arr(positions, np.float64).reshape(-1, 3), _arr(directions, np.float64).reshape(-1, 3)
        n = len(pos)
        cell, surface, distance = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n)
        check(load().mmc_test_geometry(self._handle, n, _ptr(pos, C.c_double), _ptr(dirs, C.c_double),
                                       _ptr(cell, C.c_int32), _ptr(surface, C.c_int32), _ptr(distance, C.c_double)))
        return cell, surface, distance

    def trace(self, source, seed0, first_history, n_histories, *, tracking=TRACK_SURFACE, cap=1 << 16):
        records = (EventRecord * cap)()
        n = C.c_size_t()
        o = self._options(tracking, 0, 0, 0, None)
        check(load().mmc_trace_histories(
            self._handle, C.byref(source), seed0, first_history, n_histories, C.byref(o), records, cap, C.byref(n)))
        return [records[i] for i in range(n.value)]


def exchange_plan(counts, n_total: int, rank: int) -> dict:
    """mmc_exchange_plan (host arithmetic only: no device, no NCCL), in the shape of distributed.exchange_plan."""
    P = len(counts)
    c = (C.c_uint64 * P)(*counts)
    first, count = C.c_uint64(), C.c_uint64()
    send, recv = (C.c_uint64 * (2 * P))(), (C.c_uint64 * (2 * P))()
    check(load().mmc_exchange_plan(c, P, rank, n_total, C.byref(first), C.byref(count), send, recv))
    return {"need": (first.value, count.value),
            "send": [(p, send[2 * p], send[2 * p + 1]) for p in range(P) if send[2 * p + 1]],
            "recv": [(p, recv[2 * p], recv[2 * p + 1]) for p in range(P) if recv[2 * p + 1]],
            "m_total": sum(counts)}


class Comm:
    """mmc_comm: this rank's NCCL communicator (one process per GPU)."""

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(COMM_ID_BYTES)
        check(load().mmc_comm_unique_id(buf, COMM_ID_BYTES))
        return buf.raw

    def __init__(self, nranks: int, rank: int, unique_id: bytes | None, device: int = -1):
        self._handle = C.c_void_p()
        check(load().mmc_comm_create(nranks, rank, unique_id, device, C.byref(self._handle)))

    @classmethod
    def from_torch(cls, device: int = -1, group=None):
        """Rank, size and the unique id's broadcast taken from an initialised torch.distributed group (plumbing only:
        the communicator and every collective on it are the library's own NCCL calls)."""
        import torch.distributed as dist
        rank, size = dist.get_rank(group), dist.get_world_size(group)
        box = [cls.unique_id() if rank == 0 and size > 1 else None]
        if size > 1:
            dist.broadcast_object_list(box, src=0, group=group)
        return cls(size, rank, box[0], device)

    @property
    def rank(self) -> int:
        return load().mmc_comm_rank(self._handle)

    @property
    def size(self) -> int:
        return load().mmc_comm_size(self._handle)

    @property
    def exchange_ms(self) -> float:
        return float(load().mmc_comm_exchange_ms(self._handle))

    def allreduce(self, d_words: int, n_words: int, stream=None):
        check(load().mmc_tally_allreduce(self._handle, d_words, n_words, stream))

    def close(self):
        if self._handle:
            load().mmc_comm_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Driver:
    """The C++ host's Driver (minimc_b200/host/minimc.hpp) through the C entry points: parses an XML deck exactly
    as the reference's Driver::Create does, and Solve() runs the batch on the GPU."""

    def __init__(self, path=None, *, text=None):
        self._handle = C.c_void_p()
        if text is not None:
            check(load().mmc_driver_create_from_string(text.encode(), C.byref(self._handle)))
        else:
            check(load().mmc_driver_create(os.fspath(path).encode(), C.byref(self._handle)))

    def close(self):
        if self._handle:
            load().mmc_driver_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def batchsize(self) -> int:
        return int(load().mmc_driver_batchsize(self._handle))

    @property
    def total_bins(self) -> int:
        return int(load().mmc_driver_total_bins(self._handle))

    def set_options(self, *, device=-1, secondary_capacity=0, pending_capacity=0, blocks_per_sm=0, rng_mode=RNG_MINSTD_COMPAT,
                    schedule=0, event_slots=0, profile=0):
        o = RunOptions()
        o.struct_size = C.sizeof(RunOptions)
        o.device = device
        o.rng_mode = rng_mode
        o.secondary_capacity = secondary_capacity
        o.pending_capacity = pending_capacity
        o.blocks_per_sm = blocks_per_sm
        o.schedule = schedule
        o.event_slots = event_slots
        o.profile = profile
        check(load().mmc_driver_set_options(self._handle, C.byref(o)))

    def set_shard(self, rank: int, world_size: int):
        check(load().mmc_driver_set_shard(self._handle, rank, world_size))

    def set_comm(self, comm: "Comm"):
        """Multi-process Solve(): this rank's share of the batch / of every generation, NCCL exchange inside."""
        self._comm = comm  # keep it alive
        check(load().mmc_driver_set_comm(self._handle, comm._handle))

    def cycle_seconds(self):
        """(inactive, active) host seconds of the last k-eigenvalue Solve()."""
        a, b = C.c_double(), C.c_double()
        check(load().mmc_driver_cycle_seconds(self._handle, C.byref(a), C.byref(b)))
        return a.value, b.value

    def k_collision(self):
        """(mean, std of the mean, per-cycle values, exchange ms) of the collision estimator of k of the last Solve()."""
        k_mean, k_std, n, ms = C.c_double(), C.c_double(), C.c_size_t(), C.c_double()
        check(load().mmc_driver_k_collision(self._handle, C.byref(k_mean), C.byref(k_std), None, 0, C.byref(n), C.byref(ms)))
        cycles = np.zeros(max(n.value, 1))
        check(load().mmc_driver_k_collision(self._handle, C.byref(k_mean), C.byref(k_std), _ptr(cycles, C.c_double), n.value,
                                            C.byref(n), C.byref(ms)))
        return k_mean.value, k_std.value, cycles[:n.value], ms.value

    def solve(self):
        check(load().mmc_driver_solve(self._handle))
        return self.scores()

    def scores(self):
        n = self.total_bins
        scores, squares = np.zeros(max(n, 1)), np.zeros(max(n, 1))
        check(load().mmc_driver_scores(self._handle, _ptr(scores, C.c_double), _ptr(squares, C.c_double), n))
        return scores[:n], squares[:n]

    def add_scores(self, scores, squares):
        scores, squares = _arr(scores, np.float64), _arr(squares, np.float64)
        check(load().mmc_driver_add_scores(self._handle, _ptr(scores, C.c_double), _ptr(squares, C.c_double), scores.size))

    def counters(self) -> dict:
        c = Counters()
        check(load().mmc_driver_counters(self._handle, C.byref(c)))
        return c.as_dict()

    def _text(self, fn) -> str:
        n = fn(self._handle, None, 0)
        buf = C.create_string_buffer(n + 1)
        fn(self._handle, buf, n + 1)
        return buf.value.decode()

    def output(self) -> str:
        """The text of <input>.out (minimc.cpp:20-21)."""
        return self._text(load().mmc_driver_output)

    def world_json(self) -> dict:
        import json
        text = self._text(load().mmc_driver_world_json)
        if not text:
            raise MinimcError(ERR_INVALID, last_error())
        return json.loads(text)

    def run_device(self, first_history, n_histories, d_scores, d_square, d_counters, stream=None):
        """mmc_driver_run_device: raw device pointers (ints), asynchronous on `stream` (a cudaStream_t handle)."""
        check(load().mmc_driver_run_device(self._handle, first_history, n_histories, d_scores, d_square, d_counters, stream))

    def release_device(self):
        load().mmc_driver_release_device(self._handle)

    def refresh_device(self):
        """Flattens the World again and uploads the tables into the existing device world (no allocation)."""
        check(load().mmc_driver_refresh_device(self._handle))

    @property
    def table_bytes(self) -> int:
        return int(load().mmc_driver_table_bytes(self._handle))

    @property
    def last_launches(self) -> int:
        return int(load().mmc_driver_last_launches(self._handle))

    def last_kernel_ms(self):
        """(flight ms, S(a,b) ms, boundary ms) of the last event-split run made with set_options(profile=1)."""
        a, b = C.c_double(), C.c_double()
        load().mmc_driver_last_kernel_ms(self._handle, C.byref(a), C.byref(b))
        return a.value, b.value, float(load().mmc_driver_last_boundary_ms(self._handle))

    def trace(self, first_history: int, n_histories: int, cap=1 << 18):
        records = (EventRecord * cap)()
        n = C.c_size_t()
        check(load().mmc_driver_trace(self._handle, first_history, n_histories, records, cap, C.byref(n)))
        return [records[i] for i in range(n.value)]

    def keff(self):
        k_mean, k_std, n = C.c_double(), C.c_double(), C.c_size_t()
        check(load().mmc_driver_keff(self._handle, C.byref(k_mean), C.byref(k_std), None, 0, C.byref(n)))
        cycles = np.zeros(max(n.value, 1))
        check(load().mmc_driver_keff(self._handle, C.byref(k_mean), C.byref(k_std), _ptr(cycles, C.c_double), n.value,
                                     C.byref(n)))
        return k_mean.value, k_std.value, cycles[:n.value]
