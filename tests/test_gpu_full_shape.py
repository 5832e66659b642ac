"""Parity on the configurations the published numbers are measured on: the BASELINE decks at the
reference's full table shapes (rank 10, partitions up to 97 x 18 x 294: ce_decks "full"), library-default schedule
(event-split kernels, dense / evaluated S(a,b) tables, hand-over of the drain), against the LIVE reference binary
(oracle/_ref/ref_harness = /root/reference/src compiled behind shims; it travels with the repo):

  * the .out text of Driver::Solve() byte for byte at 3*10^4 histories,
  * the event records (event, cell, surface, rng state, position, direction, energy) of 200 histories bit for bit.

bench.py's workloads are these decks: single_zone_benchmark_deck (BASELINE configs[1], benchmarks/single_zone.xml:73-118
as shipped: 103 x 101 bins), multi_zone (201 energy boundaries), broomstick (238 x 182), continuous_temperature (201)."""
import numpy as np
import pytest

from minimc_b200 import capi, ce_decks
from oracle import port_py

pytestmark = pytest.mark.gpu

HISTORIES = 30_000
TRACED = 200

FULL_DECKS = {
    "single_zone": lambda d, **kw: ce_decks.single_zone_benchmark_deck(d, threads=8, **kw),
    "multi_zone": lambda d, **kw: ce_decks.multi_zone_deck(d, threads=8, n_energy_bins=201, **kw),
    "broomstick": lambda d, **kw: ce_decks.broomstick_deck(d, threads=8, n_energy_bins=238, n_cosine_bins=182, **kw),
    "continuous_temperature": lambda d, **kw: ce_decks.continuous_temperature_deck(d, threads=8, n_energy_bins=201, **kw),
    # evaluated rows (cells with their own constant temperature) and two-row reconstructions (global linear field) mixed
    "mixed_temperature": lambda d, **kw: ce_decks.mixed_temperature_deck(d, threads=8, n_energy_bins=201, **kw),
}


@pytest.fixture(scope="module")
def full_tables(tmp_path_factory):
    d = tmp_path_factory.mktemp("full_tables")
    ce_decks.generate_tables(d, "full")
    return d


def _need_reference():
    if not port_py.ref_available():
        pytest.skip("oracle/_ref/ref_harness was not built")


def _records_equal(mine, ref):
    assert len(mine) == len(ref)
    for a, b in zip(mine, ref):
        ta = (int(a.history), int(a.particle), int(a.event), int(a.cell), int(a.surface), int(a.rng_state))
        tb = (b["history"], b["particle"], b["event"], b["cell"], b["surface"], b["rng_state"])
        if a.event == 0:  # the cell of a birth record is not part of the contract
            ta, tb = ta[:3] + ta[4:], tb[:3] + tb[4:]
        assert ta == tb
        va = np.array(list(a.position) + list(a.direction) + [a.energy])
        vb = np.array(list(b["position"]) + list(b["direction"]) + [b["energy"]])
        assert np.array_equal(va, vb), (ta, va, vb)


@pytest.mark.parametrize("name", list(FULL_DECKS))
def test_full_shape_out_is_byte_identical_to_live_reference(full_tables, tmp_path, name):
    _need_reference()
    path = tmp_path / f"{name}.xml"
    path.write_text(FULL_DECKS[name](full_tables, histories=HISTORIES))
    drv = capi.Driver(path)  # library defaults: schedule, slots, tables
    drv.solve()
    c = drv.counters()
    assert c["n_histories"] == HISTORIES and c["n_lost"] == c["n_physics_errors"] == c["n_capacity_overflow"] == 0
    if name != "broomstick":  # (a broomstick history is one flight: nothing is left alive after two passes)
        assert drv.last_launches > 1  # the event-split kernels ran, not the fused kernel
    assert drv.output() == port_py.ref_run(path)[0]


@pytest.mark.parametrize("name", list(FULL_DECKS))
def test_full_shape_event_traces_are_bit_exact(full_tables, tmp_path, name):
    _need_reference()
    path = tmp_path / f"{name}.xml"
    path.write_text(FULL_DECKS[name](full_tables, histories=HISTORIES, seed=987654321))
    drv = capi.Driver(path)
    _records_equal(drv.trace(5000, TRACED, cap=1 << 18), port_py.ref_trace(path, 5000, TRACED))


@pytest.mark.parametrize("name", ["single_zone", "multi_zone", "continuous_temperature", "mixed_temperature"])
def test_full_shape_refilled_slots_and_fresh_seed(full_tables, tmp_path, name):
    """Few slots (every slot is refilled ~10 times, the queues end in ragged warps) and a seed no fixture holds."""
    _need_reference()
    path = tmp_path / f"{name}.xml"
    path.write_text(FULL_DECKS[name](full_tables, histories=HISTORIES, seed=20261018))
    drv = capi.Driver(path)
    drv.set_options(schedule=capi.SCHEDULE_EVENT, event_slots=3001)
    drv.solve()
    assert drv.output() == port_py.ref_run(path)[0]


def test_evaluated_tables_change_nothing(full_tables, tmp_path, monkeypatch):
    """Evaluated S(a,b) tables (world_blob.h TslPartition::off_eval: Evaluate(cdf, grid, T_cell) made once per upload)
    against the two-row dense tables and the on-the-fly rank-R sums: multi_zone (13 temperatures) and mixed_temperature at
    full shape with every table (default), a 3 MB budget (the dense tables of the small partitions and some evaluated
    tables fit, the rest is summed on the fly: the kinds mix inside one sampler) and none (MMC_TSL_DENSE_MB=0) --
    identical tallies, counters and traces under the event-split and the fused schedule."""
    n = 100_000
    for name in ("multi_zone", "mixed_temperature"):
        text = FULL_DECKS[name](full_tables, histories=n)
        results = []
        for budget in (None, "3", "0"):
            if budget is None:
                monkeypatch.delenv("MMC_TSL_DENSE_MB", raising=False)
            else:
                monkeypatch.setenv("MMC_TSL_DENSE_MB", budget)
            for schedule in (capi.SCHEDULE_EVENT, capi.SCHEDULE_FUSED):
                drv = capi.Driver(text=text)  # the budget is read when the device world is built
                drv.set_options(schedule=schedule)
                scores, squares = drv.solve()
                c = drv.counters()
                assert c["n_histories"] == n and c["n_lost"] == c["n_physics_errors"] == 0
                records = [(int(r.history), int(r.event), int(r.rng_state), float(r.energy), tuple(r.position), tuple(r.direction))
                           for r in capi.Driver(text=text).trace(0, 40, cap=1 << 16)]
                results.append((scores, squares, c, records))
        monkeypatch.delenv("MMC_TSL_DENSE_MB", raising=False)
        for scores, squares, c, records in results[1:]:
            assert np.array_equal(scores, results[0][0]) and np.array_equal(squares, results[0][1])
            assert c == results[0][2]
            assert records == results[0][3]


def test_sorted_row_search_changes_nothing(full_tables, monkeypatch):
    """The S(a,b) kernel kind that holds only the direct samplers over evaluated rows (WorldHeader::tsl_all_direct:
    one load per reconstruction, the CDF lookup table, ce::find_cdf_bisect) against the kind a world falls back to when
    its tables are not TslTable::direct (MMC_TSL_SORTED_SEARCH=0 clears the flag: the dense kind, direct samplers over
    two-row reconstructions): identical tallies, counters and traces on single_zone and multi_zone at full shape,
    event-split and fused schedule.  (Until r02m the default searched sorted rows by rounds of independent loads, which
    is what the environment variable is named after.)"""
    n = 100_000
    for name in ("single_zone", "multi_zone"):
        text = FULL_DECKS[name](full_tables, histories=n)
        results = []
        for flag in (None, "0"):
            if flag is None:
                monkeypatch.delenv("MMC_TSL_SORTED_SEARCH", raising=False)
            else:
                monkeypatch.setenv("MMC_TSL_SORTED_SEARCH", flag)
            for schedule in (capi.SCHEDULE_EVENT, capi.SCHEDULE_FUSED):
                drv = capi.Driver(text=text)  # the flag is read when the device world is built
                drv.set_options(schedule=schedule)
                scores, squares = drv.solve()
                records = [(int(r.history), int(r.event), int(r.rng_state), float(r.energy), tuple(r.position), tuple(r.direction))
                           for r in capi.Driver(text=text).trace(0, 40, cap=1 << 16)]
                results.append((scores, squares, drv.counters(), records))
        monkeypatch.delenv("MMC_TSL_SORTED_SEARCH", raising=False)
        for scores, squares, c, records in results[1:]:
            assert np.array_equal(scores, results[0][0]) and np.array_equal(squares, results[0][1])
            assert c == results[0][2] and records == results[0][3]
