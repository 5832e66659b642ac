"""MMCTAB1 <-> pandas: the converter for anyone who holds the reference's nuclear data.

The reference stores every table -- pointwise cross sections, the POD factors of the thermal scattering law -- as a
pandas object written with `to_hdf(..., format="fixed")` (pyminimc/pyminimc/util.py:112-150: DataFrames with one
"coefficient" column and a MultiIndex built with `from_product`), and reads back only three things through libhdf5
(src/HDF5DataSet.hpp:88-130): the attribute `axis1_nlevels` = D, the D datasets `axis1_level{i}` = the index levels as
doubles, and `block0_values` = the values flattened in index (row-major, last level fastest) order.  libhdf5 is not in
this image and this repo's host reads the same content from a flat file instead:

    char magic[8] = "MMCTAB1\\0"; uint64 ndim; uint64 shape[ndim]; double axis_i[shape[i]] ...; double values[prod(shape)]

`from_pandas` writes that file from the pandas object itself, `to_pandas` reads it back (round trip), and

    python -m minimc_b200.mmctab data/h_in_h2o_total.hdf5 [tables/h_in_h2o_total.mmctab]

converts one of the reference's files where pandas can open it (`pd.read_hdf` needs PyTables, which the reference's
own tool chain has and this image does not).  Decks then name the .mmctab files where they named the .hdf5 files."""
from __future__ import annotations

import struct
import sys
from pathlib import Path

import numpy as np

MAGIC = b"MMCTAB1\0"


def write_table(path, axes, values) -> None:
    axes = [np.ascontiguousarray(a, np.float64) for a in axes]
    values = np.ascontiguousarray(values, np.float64)
    assert values.shape == tuple(len(a) for a in axes), (values.shape, [len(a) for a in axes])
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<Q", len(axes)))
        f.write(struct.pack(f"<{len(axes)}Q", *[len(a) for a in axes]))
        for a in axes:
            f.write(a.tobytes())
        f.write(values.tobytes())


def read_table(path):
    """(axes, values) of an MMCTAB1 file; values has shape (len(axis_0), ..., len(axis_{D-1}))."""
    raw = Path(path).read_bytes()
    if raw[:8] != MAGIC:
        raise ValueError(f"{path}: not an MMCTAB1 file")
    (ndim,) = struct.unpack_from("<Q", raw, 8)
    shape = struct.unpack_from(f"<{ndim}Q", raw, 16)
    at = 16 + 8 * ndim
    axes = []
    for n in shape:
        axes.append(np.frombuffer(raw, np.float64, n, at).copy())
        at += 8 * n
    count = int(np.prod(shape)) if ndim else 1
    if len(raw) != at + 8 * count:
        raise ValueError(f"{path}: {len(raw)} bytes, expected {at + 8 * count}")
    return axes, np.frombuffer(raw, np.float64, count, at).copy().reshape(shape)


def from_pandas(obj, path) -> None:
    """Writes a pandas Series / one-column DataFrame the way HDF5DataSet<D> would read its `fixed` HDF5 form: the index
    levels become the axes, the values the row-major block.  The index must be the full product of its levels in
    order (what `MultiIndex.from_product` gives and `HDF5DataSet::at` assumes, HDF5DataSet.hpp:152-167)."""
    import pandas as pd
    if isinstance(obj, pd.DataFrame):
        if obj.shape[1] != 1:
            raise ValueError("expected one value column (block0_values), got %d" % obj.shape[1])
        obj = obj.iloc[:, 0]
    index = obj.index
    if isinstance(index, pd.MultiIndex):
        levels = [np.asarray(level, dtype=np.float64) for level in index.levels]
        expected = pd.MultiIndex.from_product(index.levels)
        if len(index) != len(expected) or not np.array_equal(np.asarray(index.codes), np.asarray(expected.codes)):
            raise ValueError("the index is not the full, ordered product of its levels")
    else:
        levels = [np.asarray(index, dtype=np.float64)]
    for level in levels:
        if np.any(np.diff(level) <= 0):
            raise ValueError("axis values must be strictly increasing (HDF5DataSet axes are searched with upper_bound)")
    write_table(path, levels, np.asarray(obj, dtype=np.float64).reshape([len(level) for level in levels]))


def to_pandas(path, names=None, column="coefficient"):
    """The DataFrame `from_pandas` was given: one column over MultiIndex.from_product(axes)."""
    import pandas as pd
    axes, values = read_table(path)
    index = pd.MultiIndex.from_product(axes, names=names)
    return pd.DataFrame({column: values.reshape(-1)}, index=index)


def main(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    if not argv or len(argv) > 2:
        print(__doc__)
        return 2
    import pandas as pd
    source = Path(argv[0])
    target = Path(argv[1]) if len(argv) == 2 else source.with_suffix(".mmctab")
    try:
        obj = pd.read_hdf(source, key="pandas")
    except ImportError as e:
        print(f"pandas cannot open HDF5 here ({e}); run this where the reference's data were written (PyTables present)",
              file=sys.stderr)
        return 1
    from_pandas(obj, target)
    print(f"{source} -> {target}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
