"""Synthetic continuous-energy tables and their host-side loading (no GPU)."""
import hashlib

import numpy as np
import pytest

import util
from minimc_b200 import capi, ce_decks

# sha256 over the sorted "small" table files: the generator uses only + - * / sqrt, so this digest is the same on
# every machine; golden traces under tests/golden/ce were produced by the reference on exactly these bytes
SMALL_TABLES_SHA256 = "ca12aca75553f7df"


def _digest(paths):
    h = hashlib.sha256()
    for key in sorted(k for k in paths if not k.startswith("_")):
        h.update(open(paths[key], "rb").read())
    return h.hexdigest()


def test_generator_is_deterministic(tmp_path):
    a = ce_decks.generate_tables(tmp_path / "a", "small")
    b = ce_decks.generate_tables(tmp_path / "b", "small")
    assert _digest(a) == _digest(b)
    assert _digest(a)[:16] == SMALL_TABLES_SHA256


def test_tables_satisfy_the_samplers_invariants(tmp_path):
    """What ThermalScattering's constructor asserts (ThermalScattering.cpp:46-51,79-84) and what its samplers
    need: increasing concatenated grids, the beta grid ending at the cutoff energy, the alpha grid reaching past
    E_cutoff / (k T_min) and beta_cutoff, monotone quantile reconstructions."""
    import struct

    def read(path):
        raw = open(path, "rb").read()
        assert raw[:8] == b"MMCTAB1\0"
        ndim = struct.unpack_from("<Q", raw, 8)[0]
        shape = struct.unpack_from(f"<{ndim}Q", raw, 16)
        off = 16 + 8 * ndim
        axes = []
        for n in shape:
            axes.append(np.frombuffer(raw, np.float64, n, off))
            off += 8 * n
        return axes, np.frombuffer(raw, np.float64, int(np.prod(shape)), off).reshape(shape)

    for size in ("small", "full"):
        p = ce_decks.generate_tables(tmp_path / size, size)
        Es = np.concatenate([read(p[f"beta_{i}_E_T"])[0][0] for i in range(4)])
        betas = np.concatenate([read(p[f"alpha_{i}_beta_T"])[0][0] for i in range(4)])
        assert np.all(np.diff(Es) > 0) and np.all(np.diff(betas) > 0)
        assert Es[-1] == ce_decks.TSL_CUTOFF_ENERGY == read(p["scatter_xs_E"])[0][0][-1]
        assert betas[-1] > ce_decks.TSL_CUTOFF_ENERGY / (ce_decks.BOLTZMANN * 273.6) and betas[-1] > ce_decks.BETA_CUTOFF
        for kind, modes in (("beta", "E_T"), ("alpha", "beta_T")):
            for i in range(4):
                (F, _), C = read(p[f"{kind}_{i}_CDF"])
                (_,), S = read(p[f"{kind}_{i}_S"])
                (_, T, _), M = read(p[f"{kind}_{i}_{modes}"])
                assert np.all(np.diff(F) > 0) and 0 < F[0] and F[-1] < 1
                q = np.einsum("r,fr,gtr->fgt", S, C, M)  # reconstructed quantile function
                assert np.all(np.diff(q, axis=0) > 0), (size, kind, i)
        sizes = ce_decks.SIZES[size]
        assert [read(p[f"beta_{i}_CDF"])[1].shape[0] for i in range(4)] == [b[0] for b in sizes["beta"]]


@pytest.mark.parametrize("name", list(ce_decks.CE_DECKS))
def test_host_loads_continuous_decks(tmp_path, name):
    ce_decks.generate_tables(tmp_path, "small")
    drv = capi.Driver(text=ce_decks.CE_DECKS[name](tmp_path, histories=10))
    w = drv.world_json()
    assert w["n_groups"] == 0 and w["ce_nuclides"] >= 1 and w["mg_total"] == []
    assert drv.batchsize == 10


def test_table_file_errors(tmp_path):
    """HDF5DataSet.hpp:100-117: missing file and wrong dimensionality."""
    p = ce_decks.generate_tables(tmp_path, "small")
    text = ce_decks.slab_deck(tmp_path, histories=10)
    with pytest.raises(capi.MinimcError) as e:
        capi.Driver(text=text.replace(p["H1_gamma"], p["scatter_xs_E"]))
    assert e.value.message == p["scatter_xs_E"] + ": Expected 1 dimensions, but got 2"
    with pytest.raises(capi.MinimcError) as e:
        capi.Driver(text=text.replace(p["H1_gamma"], str(tmp_path / "missing.mmctab")))
    assert e.value.message == "File not found: " + str(tmp_path / "missing.mmctab")
    (tmp_path / "lfs_pointer.hdf5").write_text("version https://git-lfs.github.com/spec/v1\n")
    with pytest.raises(capi.MinimcError) as e:
        capi.Driver(text=text.replace(p["H1_gamma"], str(tmp_path / "lfs_pointer.hdf5")))
    assert "not an MMCTAB1 table file" in e.value.message
