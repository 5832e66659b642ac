"""The C-ABI library on a box without a GPU: it loads, exports every symbol include/minimc_b200.h declares,
validates tables, and refuses to compute (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import util
from minimc_b200 import capi

HEADER = Path(__file__).resolve().parent.parent / "include" / "minimc_b200.h"


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(mmc_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    names = declared_functions()
    assert len(names) >= 9
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(capi.EXPORTS) == names
    assert lib.mmc_abi_version() == capi.ABI_VERSION


def test_struct_layouts_match_the_header():
    """sizeof of every struct as the C compiler sees it (compiled probe) equals the ctypes mirror."""
    import subprocess, tempfile, os
    src = '#include <stdio.h>\n#include "minimc_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n",' \
          "sizeof(mmc_world_desc),sizeof(mmc_source_desc),sizeof(mmc_bins_desc),sizeof(mmc_estimator_desc)," \
          "sizeof(mmc_counters),sizeof(mmc_event_record),sizeof(mmc_run_options));return 0;}\n"
    with tempfile.TemporaryDirectory() as d:
        c = Path(d) / "probe.c"
        c.write_text(src)
        exe = Path(d) / "probe"
        subprocess.run(["gcc", "-I", os.fspath(HEADER.parent), "-o", os.fspath(exe), os.fspath(c)], check=True)
        sizes = [int(v) for v in subprocess.run([os.fspath(exe)], capture_output=True, text=True).stdout.split()]
    mirror = [capi.WorldDesc, capi.SourceDesc, capi.BinsDesc, capi.EstimatorDesc, capi.Counters, capi.EventRecord,
              capi.RunOptions]
    assert sizes == [C.sizeof(m) for m in mirror]


def _flat():
    return util.flat_from_xml(util.deck_text("three_shells", "surface"))


def test_world_validation_errors():
    lib = capi.load()
    flat = _flat()
    fw = capi.FlatWorld(**flat["world"])
    handle = C.c_void_p()
    d = fw.desc()
    d.struct_size = 8
    assert lib.mmc_world_create(C.byref(d), -1, C.byref(handle)) == capi.ERR_INVALID
    assert "ABI mismatch" in capi.last_error()
    bad = dict(flat["world"])
    bad["cell_material"] = np.array([0, 1, 7, -1], np.int32)
    d = capi.FlatWorld(**bad).desc()
    assert lib.mmc_world_create(C.byref(d), -1, C.byref(handle)) == capi.ERR_INVALID
    assert "material index 7 out of range" in capi.last_error()
    bad = dict(flat["world"])
    bad["surface_type"] = np.array([0, 5, 0], np.int32)
    d = capi.FlatWorld(**bad).desc()
    assert lib.mmc_world_create(C.byref(d), -1, C.byref(handle)) == capi.ERR_INVALID
    assert "unknown type" in capi.last_error()


def test_no_cpu_fallback_without_a_device():
    lib = capi.load()
    if lib.mmc_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(capi.MinimcError) as err:
        util.product_world(_flat())
    assert err.value.status == capi.ERR_NO_DEVICE
    assert "no CPU transport path" in err.value.message
    with pytest.raises(capi.MinimcError) as err:
        capi.device_math(0, np.ones(4))
    assert err.value.status == capi.ERR_NO_DEVICE


def test_estimator_size():
    lib = capi.load()
    est = util.product_estimators(_flat())
    assert est.sizes == [2, 2, 12]
    assert [lib.mmc_estimator_size(C.byref(est.array[i])) for i in range(3)] == [2, 2, 12]


def test_product_never_links_the_oracle():
    """The shipped library must not reference anything under oracle/ (voids parity claims otherwise)."""
    import subprocess
    out = subprocess.run(["ldd", str(capi.LIB_PATH)], capture_output=True, text=True).stdout
    assert "oracle" not in out
    syms = subprocess.run(["nm", "-D", str(capi.LIB_PATH)], capture_output=True, text=True).stdout
    assert "orc_" not in syms
