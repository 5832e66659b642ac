"""The invariant behind TslPartition::off_cdf_lut (csrc/world_blob.h; built in csrc/capi.cu, read by
ce::tsl_try_evaluated in csrc/physics_ce.cuh): for a sorted CDF axis inside [0, 1],

    lut[k] = std::upper_bound(cdf, k / 1024)          (k / 1024 and F * 1024 are exact: powers of two)

is never above std::upper_bound(cdf, F) for any F of bucket k = floor(F * 1024), and walking up from lut[k] while F is
not below the node ends exactly at std::upper_bound(cdf, F) -- the index ThermalScattering.cpp:292,434 computes.

This is a CPU restatement of the RULE (numpy), kept as a tripwire for the rule and for the table generator; the code
itself is pinned by the GPU parity tests against the live reference (tests/test_gpu_full_shape.py), which run through
the lookup table on every S(a,b) try."""
import numpy as np
import pytest

from minimc_b200 import ce_decks

K = 1024


def build_lut(cdf):
    """csrc/capi.cu: usable only for at most 255 nodes, sorted, inside [0, 1]; None otherwise (the device then
    searches with the hinted upper_bound)."""
    cdf = np.asarray(cdf, np.float64)
    if len(cdf) > 255 or np.any(cdf < 0) or np.any(cdf > 1) or np.any(np.diff(cdf) < 0):
        return None
    return np.searchsorted(cdf, np.arange(K) / K, side="right").astype(np.uint8)


def device_search(cdf, lut, F):
    """ce::tsl_try_evaluated: the pair table's F_hi is cdf[first], 1.0 past the end."""
    first = int(lut[int(F * K)])
    while first < len(cdf) and not (F < cdf[first]):
        first += 1
    return first


def probes(cdf):
    rng = np.random.default_rng(1234)
    edges = np.arange(K) / K
    around = np.concatenate([edges, np.nextafter(edges, 0.0), np.nextafter(edges, 1.0)])
    nodes = np.concatenate([cdf, np.nextafter(cdf, 0.0), np.nextafter(cdf, 1.0)])
    F = np.concatenate([around, nodes, rng.random(20000), [0.0, np.nextafter(1.0, 0.0)]])
    return F[(F >= 0.0) & (F < 1.0)]


@pytest.mark.parametrize("n", [2, 13, 32, 97, 255])
def test_lookup_then_walk_is_upper_bound_on_the_generated_axes(n):
    cdf = ce_decks._cdf_axis(n)
    lut = build_lut(cdf)
    assert lut is not None, "the generator's CDF axes are sorted and inside [0, 1]"
    for F in probes(cdf):
        want = int(np.searchsorted(cdf, F, side="right"))
        assert lut[int(F * K)] <= want
        assert device_search(cdf, lut, F) == want


def test_clustered_and_repeated_nodes():
    # many nodes inside one bucket, repeated nodes, nodes at 0 and at 1: the walk does the work the table cannot
    cdf = np.sort(np.concatenate([[0.0, 0.0, 1.0], 0.5 + np.arange(40) * 1e-6, [0.25, 0.25, 0.25], np.linspace(0.9, 1.0, 30)]))
    lut = build_lut(cdf)
    assert lut is not None
    for F in probes(cdf):
        assert device_search(cdf, lut, F) == int(np.searchsorted(cdf, F, side="right"))


@pytest.mark.parametrize("cdf", [np.linspace(0, 1, 256), [0.1, 0.05, 0.9], [-0.1, 0.5], [0.5, 1.5]])
def test_axes_the_table_is_not_built_for(cdf):
    assert build_lut(cdf) is None
