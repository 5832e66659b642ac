"""Continuous-energy + thermal-scattering parity of the CUDA path (through the C++ host's Driver and the C ABI) with
the reference's own code: golden traces / .out files written by oracle/_ref/ref_harness (tests/golden/generate.py)
on the synthetic tables of minimc_b200/ce_decks.py, and -- when the prebuilt reference binary travelled with the
repo -- the reference itself on fresh seeds.

Contract: for every deck in util.CE_EXACT the event sequence, energies, positions, directions and RNG states are
BIT-EXACT and the .out text is identical; free_gas_sphere with surface tracking evaluates erf/exp (CUDA's, not
glibc's) in the free-gas cross-section adjustment, so there the integer fields must match and floating-point fields
agree to 1e-12 relative; its tallies agree within 3 sigma (north_star's stated tolerance)."""
import numpy as np
import pytest

import util
from minimc_b200 import capi, ce_decks
from oracle import port_py

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tables(tmp_path_factory):
    d = tmp_path_factory.mktemp("tables")
    ce_decks.generate_tables(d, "small")
    return d


def _case(tables, name, tag, **kw):
    for n, t, text in util.ce_cases(tables, **kw):
        if (n, t) == (name, tag):
            return text
    raise KeyError((name, tag))


def _compare_traces(mine, ref, exact):
    assert len(mine) == len(ref)
    for a, b in zip(mine, ref):
        ta = (int(a.history), int(a.particle), int(a.event), int(a.cell), int(a.surface), int(a.rng_state))
        tb = (b["history"], b["particle"], b["event"], b["cell"], b["surface"], b["rng_state"])
        if a.event == 0:  # the cell of a birth record is not part of the contract
            ta, tb = ta[:3] + ta[4:], tb[:3] + tb[4:]
        assert ta == tb
        va = np.array(list(a.position) + list(a.direction) + [a.energy])
        vb = np.array(list(b["position"]) + list(b["direction"]) + [b["energy"]])
        if exact:
            assert np.array_equal(va, vb), (ta, va, vb)
        else:
            assert np.allclose(va, vb, rtol=1e-12, atol=1e-13), (ta, va, vb)


@pytest.mark.parametrize("name,tag", util.CE_CASE_IDS)
def test_ce_event_traces_match_reference_golden(tables, name, tag):
    drv = capi.Driver(text=_case(tables, name, tag))
    drv.set_options(secondary_capacity=256)
    mine = drv.trace(0, util.CE_TRACE_HISTORIES, cap=1 << 18)
    ref = port_py.parse_trace((util.GOLDEN / "ce" / f"{name}__{tag}.trace").read_text())
    _compare_traces(mine, ref, (name, tag) in util.CE_EXACT)


# (schedule, event_slots): the library default (event-split for continuous energy, the drain handed to the fused
# kernel), the fused persistent kernel alone, the event-split kernels to the last history with a slot count that is
# neither a multiple of the warp nor of the CTA (slots are refilled many times, the compacted queues end in ragged
# warps), and the hand-over from ragged queues
SCHEDULES = {"default": (capi.SCHEDULE_AUTO, 0), "fused": (capi.SCHEDULE_FUSED, 0),
             "event_ragged": (capi.SCHEDULE_EVENT_ONLY, 301), "event_handover": (capi.SCHEDULE_EVENT, 1001)}


@pytest.mark.parametrize("schedule", list(SCHEDULES))
@pytest.mark.parametrize("name,tag", util.CE_CASE_IDS)
def test_ce_tallies_match_reference_out(tables, name, tag, schedule):
    drv = capi.Driver(text=_case(tables, name, tag))
    drv.set_options(secondary_capacity=256, schedule=SCHEDULES[schedule][0], event_slots=SCHEDULES[schedule][1])
    scores, squares = drv.solve()
    c = drv.counters()
    assert (drv.last_launches == 1) == (schedule == "fused")
    assert c["n_histories"] == util.CE_HISTORIES and c["n_lost"] == 0 and c["n_physics_errors"] == 0
    golden = (util.GOLDEN / "ce" / f"{name}__{tag}.out").read_text()
    if (name, tag) in util.CE_EXACT:
        assert drv.output() == golden  # the whole .out text, byte for byte
    else:
        _, ref = port_py.parse_out(golden)
        _, mine = port_py.parse_out(drv.output())
        n = util.CE_HISTORIES
        for est in ref:
            m, r = np.array(mine[est]["mean"], float), np.array(ref[est]["mean"], float)
            s = np.sqrt(np.array(mine[est]["std dev"], float) ** 2 + np.array(ref[est]["std dev"], float) ** 2)
            assert np.all(np.abs(m - r) <= 3 * s + 1.0 / n), est


@pytest.mark.parametrize("name,tag", util.CE_CASE_IDS)
def test_ce_fresh_seed_against_live_reference(tables, tmp_path, name, tag):
    """The reference binary itself (prebuilt in oracle/_ref, no /root/reference needed) on a seed no fixture holds."""
    if not port_py.ref_available():
        pytest.skip("oracle/_ref/ref_harness was not built")
    text = _case(tables, name, tag, histories=3000, seed=424242)
    path = tmp_path / "deck.xml"
    path.write_text(text)
    drv = capi.Driver(path)
    drv.set_options(secondary_capacity=256)
    exact = (name, tag) in util.CE_EXACT
    _compare_traces(drv.trace(100, 120, cap=1 << 18), port_py.ref_trace(path, 100, 120), exact)
    drv.solve()
    out, _ = port_py.ref_run(path)
    if exact:
        assert drv.output() == out


def test_ce_full_size_tables_properties(tmp_path):
    """The real table shapes (rank 10, partitions up to 97 x 18 x 294) at 2*10^5 histories of the
    continuous_temperature deck (BASELINE config C5): every history ends in exactly one capture or leak, the result
    is independent of how the batch is split across ranks, and a 3000-history prefix equals the live reference."""
    ce_decks.generate_tables(tmp_path, "full")
    n = 200_000
    text = ce_decks.continuous_temperature_deck(tmp_path, histories=n, threads=4)
    drv = capi.Driver(text=text)
    scores, squares = drv.solve()
    c = drv.counters()
    assert c["n_histories"] == c["n_births"] == n
    assert c["n_lost"] == c["n_physics_errors"] == c["n_capacity_overflow"] == 0
    assert c["n_events"] == c["n_collisions"] + c["n_crossings"] + c["n_virtual"]
    assert c["n_crossings"] <= n and np.all(squares == scores)  # at most one leak per history, score 1
    total = np.zeros_like(scores)
    for rank in range(3):
        part = capi.Driver(text=text)
        part.set_shard(rank, 3)
        s, _ = part.solve()
        total += s
    assert np.array_equal(total, scores)
    if port_py.ref_available():
        path = tmp_path / "deck.xml"
        path.write_text(ce_decks.continuous_temperature_deck(tmp_path, histories=3000, threads=4))
        small = capi.Driver(path)
        small.solve()
        assert small.output() == port_py.ref_run(path)[0]


def test_ce_schedules_agree_at_scale(tmp_path):
    """Full-shape tables, 3*10^5 histories of single_zone (BASELINE configs[1]) and of continuous_temperature: the
    event-split schedule (with and without the hand-over of the drain, ragged slot count) and the fused kernel
    give identical integer tallies and counters -- scheduling never changes a history."""
    ce_decks.generate_tables(tmp_path, "full")
    n = 300_000
    for text in (ce_decks.single_zone_benchmark_deck(tmp_path, histories=n, threads=4),
                 ce_decks.continuous_temperature_deck(tmp_path, histories=n, threads=4)):
        results = []
        for schedule, slots in ((capi.SCHEDULE_FUSED, 0), (capi.SCHEDULE_EVENT, 0), (capi.SCHEDULE_EVENT, 65537),
                                (capi.SCHEDULE_EVENT_ONLY, 40001)):
            drv = capi.Driver(text=text)
            drv.set_options(schedule=schedule, event_slots=slots)
            scores, squares = drv.solve()
            c = drv.counters()
            assert c["n_histories"] == n and c["n_lost"] == c["n_physics_errors"] == c["n_capacity_overflow"] == 0
            results.append((scores, squares, c))
        for scores, squares, c in results[1:]:
            assert np.array_equal(scores, results[0][0]) and np.array_equal(squares, results[0][1])
            assert c == results[0][2]


def test_ce_dense_tables_change_nothing(tmp_path, monkeypatch):
    """The dense reconstruction tables (world_blob.h TslPartition::off_dense, expanded on the device after every upload)
    against the on-the-fly rank-R sums they replace: full-shape tables, 2*10^5 histories of single_zone and of
    continuous_temperature with every partition expanded (default), with none (MMC_TSL_DENSE_MB=0) and with a budget
    that only fits the small partitions (1 MB: expanded and summed rows mix inside one sampler) -- identical
    tallies, counters and event traces, under the event-split and the fused schedule."""
    ce_decks.generate_tables(tmp_path, "full")
    n = 200_000
    for text in (ce_decks.single_zone_benchmark_deck(tmp_path, histories=n, threads=4),
                 ce_decks.continuous_temperature_deck(tmp_path, histories=n, threads=4)):
        results = []
        for budget in (None, "0", "1"):
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
                           for r in capi.Driver(text=text).trace(0, 40, cap=1 << 16)]  # (the trace kernel has its own schedule)
                results.append((scores, squares, c, records))
        monkeypatch.delenv("MMC_TSL_DENSE_MB", raising=False)
        for scores, squares, c, records in results[1:]:
            assert np.array_equal(scores, results[0][0]) and np.array_equal(squares, results[0][1])
            assert c == results[0][2]
            assert records == results[0][3]


def test_ce_refresh_device_expands_the_dense_tables_again(tables):
    """mmc_world_update on a continuous-energy world: the image is uploaded into the existing device world and the dense
    reconstruction tables behind it are expanded again; solving after each refresh reproduces the golden .out."""
    name, tag = "single_zone", "surface"
    drv = capi.Driver(text=_case(tables, name, tag))
    drv.set_options(secondary_capacity=256)
    golden = (util.GOLDEN / "ce" / f"{name}__{tag}.out").read_text()
    drv.solve()
    assert drv.output() == golden
    for _ in range(2):
        drv.refresh_device()
        drv.solve()
        assert drv.output() == golden


def test_ce_lost_particles_are_counted_alike_by_every_schedule(tables):
    """A slab whose far void is missing (crossing the right plane finds no Cell) and, second, a source outside every
    Cell: World::FindCellContaining throws in the reference; here the run reports MMC_ERR_LOST_PARTICLE and the
    fused, event-split and event-only schedules count the same events, crossings and lost particles."""
    text = ce_decks.slab_deck(tables, histories=3000, threads=1)
    last_void = text.rindex("  <void>")
    no_far_void = text[:last_void] + text[text.index("</void>\n", last_void) + len("</void>\n"):]
    outside = text.replace('<constant x="0" y="0" z="0"/>', '<constant x="-5" y="0" z="0"/>', 1)
    outside = outside[:outside.index("  <void>")] + outside[outside.index("</void>\n") + len("</void>\n"):]
    for deck in (no_far_void, outside):
        seen = []
        for schedule in (capi.SCHEDULE_FUSED, capi.SCHEDULE_EVENT, capi.SCHEDULE_EVENT_ONLY):
            drv = capi.Driver(text=deck)
            drv.set_options(schedule=schedule, event_slots=777)
            with pytest.raises(capi.MinimcError) as e:
                drv.solve()
            assert e.value.status == capi.ERR_LOST_PARTICLE
            seen.append(drv.counters())
        assert seen[0]["n_lost"] > 0 and seen[0]["n_histories"] == 3000
        assert seen[1] == seen[0] and seen[2] == seen[0]


def test_ce_resample_limit_is_reported(tables):
    """A source far above every table (20 MeV neutron in the slab) still runs; an absurd temperature below every
    partition's grid makes BetaPartition::Evaluate divide 0 by 0 and the resample limit trip: the reference throws
    from a noexcept function (std::terminate); here the run reports MMC_ERR_PHYSICS."""
    text = ce_decks.slab_deck(tables, histories=2000, temperature=100.0)
    for schedule in (capi.SCHEDULE_FUSED, capi.SCHEDULE_EVENT):
        drv = capi.Driver(text=text)
        drv.set_options(schedule=schedule)
        with pytest.raises(capi.MinimcError) as e:
            drv.solve()
        assert e.value.status == capi.ERR_PHYSICS
