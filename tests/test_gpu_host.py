"""The reference's public surface on the GPU path: Driver::Create(deck)->Solve() through the C++ host, the
runminimc CLI and its .out file, and the reference's own geometry unit tests (test_World.cpp, test_Cell.cpp,
test_CSGSurface.cpp, test_FixedSource.cpp) evaluated on the device."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

import util
from minimc_b200 import capi

pytestmark = pytest.mark.gpu
INF = float("inf")


@pytest.mark.parametrize("name,tracking", util.DECK_CASES)
def test_driver_output_equals_reference_out_file(name, tracking):
    """minimc.cpp:16-21: the text written to <input>.out, byte for byte, against the file the reference wrote."""
    drv = capi.Driver(text=util.deck_text(name, tracking))
    drv.set_options(secondary_capacity=256)
    drv.solve()
    assert drv.output() == (util.GOLDEN / f"{name}__{tracking}.out").read_text()


def test_cli_writes_reference_out_file(tmp_path):
    deck = tmp_path / "three_shells.xml"
    deck.write_text(util.deck_text("three_shells", "surface"))
    cli = Path(capi.LIB_PATH).parent / "runminimc_b200"
    p = subprocess.run([str(cli), str(deck)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert p.stdout.startswith("Welcome to MiniMC!")
    assert (tmp_path / "three_shells.out").read_text() == (util.GOLDEN / "three_shells__surface.out").read_text()
    # construction errors reach the user with the reference's message
    bad = tmp_path / "bad.xml"
    bad.write_text(util.deck_text("leakage_sphere", "surface").replace('material="absorber"', 'material="nope"', 1))
    p = subprocess.run([str(cli), str(bad)], capture_output=True, text=True)
    assert p.returncode == 1 and 'Material node "nope" not found' in p.stderr


def test_sharded_drivers_sum_to_the_whole():
    """EstimatorSet::operator+= over ranks (FixedSource.cpp:31-33): any split of the batch gives the same tallies."""
    text = util.deck_text("fissile_slab", "delta")
    whole = capi.Driver(text=text)
    whole.set_options(secondary_capacity=256)
    scores, squares = whole.solve()
    acc = capi.Driver(text=text)
    for rank in range(5):
        part = capi.Driver(text=text)
        part.set_options(secondary_capacity=256)
        part.set_shard(rank, 5)
        acc.add_scores(*part.solve())
    assert np.array_equal(acc.scores()[0], scores) and np.array_equal(acc.scores()[1], squares)
    assert acc.output() == whole.output()


def _three_shell_world():
    return util.product_world(util.flat_from_xml(util.deck_text("three_shells", "surface")))


def test_reference_world_cell_lookup():
    """test_World.cpp:46-53: on-surface points belong to the outer cell (Contains is strict)."""
    w = _three_shell_world()
    pts = [(0, 0, 0), (0.5, 0, 0), (1, 0, 0), (1.5, 0, 0), (2, 0, 0), (2.5, 0, 0), (3, 0, 0), (100, 0, 0)]
    cell, _, _ = w.geometry(pts, [(1, 0, 0)] * len(pts))
    assert cell.tolist() == [0, 0, 1, 1, 2, 2, 3, 3]


def test_reference_sphere_distances():
    """test_CSGSurface.cpp:23-56, evaluated from the pit cell (inner shell is its only surface) and from the
    inner-shell cell (nearest of inner / middle shell)."""
    w = _three_shell_world()
    # points inside the pit: distance to the r = 1 sphere
    cell, surface, dist = w.geometry([(0, 0, 0), (0, 0, -0.5)], [(0, 0, 1), (0, 0, 1)])
    assert cell.tolist() == [0, 0] and surface.tolist() == [0, 0]
    assert dist.tolist() == [1.0, 1.5]
    # the seven cases of the reference, on the r = 1 sphere, from whichever cell holds the point:
    # (2,0,0)->z: no intersection; (1,0,0)->z grazing; (0,0,-2)->z hits at 1; (0,0,-1)->z leaving: 2;
    # (0,0,1)->z and (0,0,2)->z: heading away
    flat = util.flat_from_xml(util.deck_text("leakage_sphere", "surface"))  # one sphere r = 1, void outside
    s = util.product_world(flat)
    pts = [(2, 0, 0), (1, 0, 0), (0, 0, -2), (0, 0, -1), (0, 0, 0), (0, 0, 1), (0, 0, 2)]
    _, _, d = s.geometry(pts, [(0, 0, 1)] * 7)
    assert d.tolist() == [INF, INF, 1.0, 2.0, 1.0, INF, INF]


def test_reference_cell_nearest_surface():
    """test_Cell.cpp:22-27: from (1.5,0,0) heading +x inside the inner-shell cell the nearest surface is the middle
    shell at 0.5."""
    w = _three_shell_world()
    cell, surface, dist = w.geometry([(1.5, 0, 0), (1.5, 0, 0)], [(1, 0, 0), (-1, 0, 0)])
    assert cell.tolist() == [1, 1]
    assert surface.tolist() == [1, 0] and dist.tolist() == [0.5, 0.5]


def test_reference_point_source_leakage():
    """test_FixedSource.cpp:13-26 as shipped: 1000 histories, leakage = exp(-1) within 3 sigma."""
    drv = capi.Driver(text=util.deck_text("leakage_sphere", "surface"))
    scores, _ = drv.solve()
    n, p = drv.batchsize, np.exp(-1)
    assert n == 1000
    assert abs(scores[0] / n - p) / p < 3 * np.sqrt((1 - p) / (p * n))


def test_refresh_device_reuploads_tables_in_place():
    """mmc_world_update through Driver::RefreshDevice: the World is flattened again and uploaded into the existing
    device world; solving before and after gives the same text, and solving twice accumulates nothing stale."""
    drv = capi.Driver(text=util.deck_text("three_shells", "surface"))
    drv.set_options(secondary_capacity=256)
    drv.solve()
    first = drv.output()
    for _ in range(3):
        drv.refresh_device()
        drv.solve()
        assert drv.output() == first
    assert first == (util.GOLDEN / "three_shells__surface.out").read_text()


def test_world_update_rejects_a_different_shape():
    """mmc_world_update only replaces values: tables of another size need a new world."""
    a = util.flat_from_xml(util.deck_text("three_shells", "surface"))
    b = util.flat_from_xml(util.deck_text("leakage_sphere", "surface"))
    world = util.product_world(a)
    import ctypes as C
    other, same = capi.FlatWorld(**b["world"]), capi.FlatWorld(**a["world"])
    other_desc, same_desc = other.desc(), same.desc()
    assert capi.load().mmc_world_update(world._handle, C.byref(other_desc)) == capi.ERR_INVALID
    assert capi.load().mmc_world_update(world._handle, C.byref(same_desc)) == 0


@pytest.mark.parametrize("n_planes", [9, 64, 65, 90])
def test_cell_lookup_in_a_stack_of_slabs(n_planes):
    """World::FindCellContaining (World.cpp:26-37) over a stack of slabs between n planes: worlds of up to 64
    surfaces evaluate every surface once and compare one mask per Cell, larger worlds walk each Cell's surface list;
    both must give the first Cell in creation order that contains the point.  A slab listed again at the END, and a
    Cell that demands both senses of one plane (it contains nothing), check the order and the contradiction rule."""
    base = util.flat_from_xml(util.deck_text("three_shells", "surface"))["world"]
    xs = np.arange(n_planes, dtype=np.float64)
    world = dict(base)
    world["surface_type"] = np.full(n_planes, 1, np.int32)  # MMC_SURF_PLANEX
    prm = np.zeros((n_planes, 4))
    prm[:, 0] = xs
    world["surface_param"] = prm.reshape(-1)
    begin, index, sense, material = [0], [], [], []

    def cell(pairs, mat):
        for s, sign in pairs:
            index.append(s), sense.append(sign)
        begin.append(len(index)), material.append(mat)

    cell([(0, 1)], -1)                                   # x < plane 0 (void)
    cell([(3, 0), (3, 1)], 0)                            # both senses of plane 3: contains nothing
    for k in range(n_planes - 1):
        cell([(k, 0), (k + 1, 1)], 0)                    # plane k <= x < plane k+1  (Contains is strict `<`)
    cell([(n_planes - 1, 0)], -1)                        # beyond the last plane (void)
    cell([(2, 0), (3, 1)], 0)                            # a copy of slab 2..3, later in the order: never chosen
    n_cells = len(material)
    world["cell_surface_begin"], world["cell_surface_index"], world["cell_surface_sense"] = begin, index, sense
    world["cell_material"] = material
    world["cell_field_kind"] = np.zeros(n_cells, np.int32)
    world["cell_field_param"] = np.zeros(n_cells * 6)
    w = capi.World(capi.FlatWorld(**world))
    rng = np.random.default_rng(7)
    pts = np.concatenate([rng.uniform(-2, n_planes + 1, 500), xs, xs - 1e-12, xs + 1e-12])
    got, _, _ = w.geometry([(x, 0.3, -0.2) for x in pts], [(1, 0, 0)] * len(pts))
    # expected: void 0 for x < 0; slab k -> cell 2 + k; beyond -> n_planes + 1
    expected = np.where(pts < xs[0], 0, np.where(pts >= xs[-1], n_planes + 1, 2 + np.floor(pts).astype(int)))
    assert got.tolist() == expected.tolist()
