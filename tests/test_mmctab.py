"""pandas -> MMCTAB1 converter: frames built the way the reference's writer builds them
(pyminimc/pyminimc/util.py:112-150: SVD factors as one-column DataFrames over MultiIndex.from_product) go through
minimc_b200.mmctab and come back equal; the files are byte-identical to what the deck generator writes directly, and
the C++ host (TableFile behind the HDF5DataSet role, HDF5DataSet.hpp:88-130) loads a deck made of converted files."""
import numpy as np
import pandas as pd
import pytest

from minimc_b200 import capi, ce_decks, mmctab


def to_svd_dfs(df, order):
    """The layout of pyminimc.util.to_svd_dfs (util.py:112-150), restated."""
    U, S, Vt = np.linalg.svd(df, full_matrices=False)
    U_df = pd.DataFrame({"coefficient": U[:, :order].flatten()},
                        index=pd.MultiIndex.from_product([df.index, range(order)], names=["CDF", "order"]))
    S_df = pd.DataFrame({"coefficient": S[:order]}, index=pd.MultiIndex.from_product([range(order)], names=["order"]))
    V_df = pd.DataFrame({"coefficient": Vt[:order, :].T.flatten()},
                        index=pd.MultiIndex.from_product([df.columns.unique(0), df.columns.unique(1), range(order)],
                                                         names=[df.columns.unique(0).name, df.columns.unique(1).name, "order"]))
    return U_df, S_df, V_df


def test_svd_frames_round_trip(tmp_path):
    rng = np.random.default_rng(7)
    cdf = pd.Index(np.linspace(0.01, 0.99, 23), name="CDF")
    columns = pd.MultiIndex.from_product([np.geomspace(1e-9, 2e-6, 11), np.linspace(273.6, 800.0, 5)], names=["E", "T"])
    frame = pd.DataFrame(rng.random((len(cdf), len(columns))), index=cdf, columns=columns)
    for name, part in zip(("CDF", "S", "E_T"), to_svd_dfs(frame, order=4)):
        path = tmp_path / f"beta_0_{name}.mmctab"
        mmctab.from_pandas(part, path)
        axes, values = mmctab.read_table(path)
        assert [len(a) for a in axes] == [len(level) for level in part.index.levels]
        assert values.shape == tuple(len(level) for level in part.index.levels)
        back = mmctab.to_pandas(path, names=part.index.names)
        assert np.array_equal(back["coefficient"].to_numpy(), part["coefficient"].to_numpy())
        assert all(np.array_equal(np.asarray(a, float), np.asarray(b, float)) for a, b in zip(back.index.levels, part.index.levels))
    # the reconstruction from the converted factors is the one from the frames
    U, S, V = (mmctab.read_table(tmp_path / f"beta_0_{n}.mmctab")[1] for n in ("CDF", "S", "E_T"))
    rebuilt = np.einsum("cr,r,etr->cet", U, S, V).reshape(len(cdf), -1)
    U_df, S_df, V_df = to_svd_dfs(frame, order=4)
    expect = (U_df["coefficient"].unstack().to_numpy() * S_df["coefficient"].to_numpy()) @ V_df["coefficient"].unstack().to_numpy().T
    assert np.allclose(rebuilt, expect, rtol=0, atol=1e-13)


def test_rejects_what_hdf5dataset_could_not_index(tmp_path):
    ragged = pd.Series([1.0, 2.0, 3.0], index=pd.MultiIndex.from_tuples([(1.0, 0), (1.0, 1), (2.0, 0)]))
    with pytest.raises(ValueError, match="product"):
        mmctab.from_pandas(ragged, tmp_path / "x.mmctab")
    with pytest.raises(ValueError, match="increasing"):
        mmctab.from_pandas(pd.Series([1.0, 2.0], index=[2.0, 1.0]), tmp_path / "x.mmctab")
    with pytest.raises(ValueError, match="one value column"):
        mmctab.from_pandas(pd.DataFrame({"a": [1.0], "b": [2.0]}), tmp_path / "x.mmctab")


def test_converted_tables_equal_generated_tables_and_load_in_the_host(tmp_path):
    """Every table of the synthetic nuclide, as pandas frames, through the converter: byte-identical files, and the C++
    host builds its World from them (construction only: no device needed)."""
    direct, converted = tmp_path / "direct", tmp_path / "converted"
    converted.mkdir()
    paths = ce_decks.generate_tables(direct, "small")
    for name, path in paths.items():
        if name.startswith("_"):
            continue
        axes, values = mmctab.read_table(path)
        frame = pd.DataFrame({"coefficient": values.reshape(-1)}, index=pd.MultiIndex.from_product(axes))
        mmctab.from_pandas(frame, converted / f"{name}.mmctab")
        assert (converted / f"{name}.mmctab").read_bytes() == open(path, "rb").read(), name
    drv = capi.Driver(text=ce_decks.slab_deck(converted, histories=10))
    assert drv.batchsize == 10 and drv.total_bins > 0
    world = drv.world_json()
    assert world["n_nuclides"] == 1 if "n_nuclides" in world else True
