"""minimc_b200/csrc/glibc_math.h -- the restatement of glibc's log / sincos / sin / cos used inside the kernels --
against the libm of this box, bit for bit.  The CPU test builds the same header for the host (g++ -mfma) and
sweeps ~10^7 arguments per branch; the GPU test evaluates the device build through mmc_test_device_math."""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

from minimc_b200 import capi

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "native" / "glibc_math_host.cpp"

CASES = [  # (what, lo, hi, mode)  what: 0 log, 1 sincos, 2 sin, 3 cos; mode 1 = uniform in bit pattern
    (0, 1e-300, 1.0, 0), (0, 0.93, 1.07, 0), (0, 2.0 ** -70, 2.0 ** 70, 1), (0, 1.0, 1e6, 0),
    (1, 0.0, 2 * np.pi, 0), (1, -7.0, 7.0, 0), (1, 2.0 ** -30, 0.9, 1), (1, 0.0, 1e5, 0), (1, 0.8, 2.5, 0),
    (2, 0.0, 2 * np.pi, 0), (2, -7.0, 7.0, 0), (2, 2.0 ** -30, 0.9, 1), (2, 0.0, 1e5, 0),
    (3, 0.0, 2 * np.pi, 0), (3, -7.0, 7.0, 0), (3, 2.0 ** -30, 0.9, 1), (3, 0.0, 1e5, 0),
    # 4: the converged sin_and_cos pair against two separate libm calls
    (4, 0.0, 2 * np.pi, 0), (4, -7.0, 7.0, 0), (4, 2.0 ** -30, 0.9, 1), (4, 0.0, 1e5, 0), (4, 0.8, 2.5, 0),
]


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    so = tmp_path_factory.mktemp("native") / "glibc_math_host.so"
    subprocess.run(["g++", "-std=c++17", "-O2", "-mfma", "-ffp-contract=off", "-fPIC", "-shared",
                    "-DMMC_HAVE_SIN_COS", "-o", os.fspath(so), os.fspath(SRC)], check=True)
    lib = C.CDLL(os.fspath(so))
    lib.mmc_host_fuzz.restype = C.c_uint64
    lib.mmc_host_fuzz.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, C.c_uint64, C.c_uint64,
                                  C.POINTER(C.c_double)]
    return lib


@pytest.mark.parametrize("what,lo,hi,mode", CASES)
def test_host_restatement_equals_libm(host_lib, what, lo, hi, mode):
    first = C.c_double()
    bad = host_lib.mmc_host_fuzz(what, lo, hi, mode, 3_000_000, 20260101 + what, C.byref(first))
    assert bad == 0, f"first mismatch at {first.value.hex()}"


def _libm():
    m = C.CDLL("libm.so.6")
    for f in ("log", "sin", "cos"):
        getattr(m, f).restype = C.c_double
        getattr(m, f).argtypes = [C.c_double]
    return m


@pytest.mark.gpu
def test_device_math_equals_libm():
    m = _libm()
    rs = np.random.default_rng(11)
    n = 100_000
    x = np.concatenate([rs.random(n), 1 - rs.random(n) * 0.12, 1 + rs.random(n) * 0.07, rs.random(n) * 1e4 + 1e-12,
                        np.array([1.0, 0.5, 2.0, 10.0, 3.0, 1e-20])])
    got = capi.device_math(0, x)[0]
    want = np.array([m.log(float(v)) for v in x])
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
    x = np.concatenate([rs.random(n) * 2 * np.pi, rs.random(n) * 14 - 7, rs.random(n) * 1e5, rs.random(n) * 0.2,
                        np.array([0.0, 1e-9, 0.126, 0.855469, 2.426265, np.pi, 2 * np.pi])])
    ws = np.array([m.sin(float(v)) for v in x])
    wc = np.array([m.cos(float(v)) for v in x])
    s, c = capi.device_math(1, x)
    assert np.array_equal(s.view(np.uint64), ws.view(np.uint64))
    assert np.array_equal(c.view(np.uint64), wc.view(np.uint64))
    assert np.array_equal(capi.device_math(2, x)[0].view(np.uint64), ws.view(np.uint64))
    assert np.array_equal(capi.device_math(3, x)[0].view(np.uint64), wc.view(np.uint64))
    s, c = capi.device_math(5, x)  # the converged (sin, cos) pair of Direction(d, mu, phi)
    assert np.array_equal(s.view(np.uint64), ws.view(np.uint64))
    assert np.array_equal(c.view(np.uint64), wc.view(np.uint64))


@pytest.mark.gpu
def test_device_rng_equals_libstdcxx():
    from oracle import port_py
    seeds = np.array([1, 2, 3, 0, 2147483647, 2147483648, 12345, 987654321, 4294967297], np.float64)
    got = capi.device_math(4, seeds)[0]
    want = np.array([port_py.rng_canonical(int(s), 1)[0][0] for s in seeds])
    assert np.array_equal(got, want)
