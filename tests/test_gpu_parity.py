"""Parity of the CUDA path (through the C ABI) with the oracle and the reference's golden fixtures.

Contract (multigroup, minstd_compat RNG): event sequence, groups, cells, surfaces, RNG state after every event,
positions and directions are BIT-EXACT; tallies (Sigma s, Sigma (per-history sum)^2) and all counters are exact
integers."""
import numpy as np
import pytest

import util
from minimc_b200 import capi

pytestmark = pytest.mark.gpu


def _product(flat):
    return util.product_world(flat), util.product_source(flat), util.product_estimators(flat)


def _assert_same_records(mine, ref):
    assert len(mine) == len(ref)
    for a, b in zip(mine, ref):
        assert util.record_tuple(a) == util.record_tuple(b)
        pa, da = util.record_vectors(a)
        pb, db = util.record_vectors(b)
        assert np.array_equal(pa, pb), (util.record_tuple(a), pa, pb)
        assert np.array_equal(da, db), (util.record_tuple(a), da, db)


@pytest.mark.parametrize("name,tracking", util.DECK_CASES)
def test_event_traces_match_reference_golden(name, tracking):
    flat = util.flat_from_xml(util.deck_text(name, tracking))
    world, src, _ = _product(flat)
    mine = world.trace(src, flat["run"]["seed"], 0, util.TRACE_HISTORIES, tracking=flat["run"]["tracking"],
                       cap=1 << 18)
    _assert_same_records(mine, util.golden_trace(name, tracking))


@pytest.mark.parametrize("name,tracking", util.DECK_CASES)
def test_event_traces_match_oracle_on_fresh_seed(name, tracking):
    flat = util.flat_from_xml(util.deck_text(name, tracking, seed=77777))
    world, src, _ = _product(flat)
    mine = world.trace(src, 77777, 1000, 400, tracking=flat["run"]["tracking"], cap=1 << 18)
    ref = util.oracle_problem(flat).trace(1000, 400, cap=1 << 18)
    _assert_same_records(mine, ref)


@pytest.mark.parametrize("name,tracking", util.DECK_CASES)
def test_tallies_match_reference_out_and_oracle(name, tracking):
    flat = util.flat_from_xml(util.deck_text(name, tracking))
    world, src, est = _product(flat)
    run = flat["run"]
    scores, squares, counters = world.fixed_source_run(src, est, run["seed"], 0, run["histories"],
                                                       tracking=run["tracking"], secondary_capacity=256)
    # the reference's own .out (formatted as Scorable::GetScoreAsString does)
    _, ref_out = util.golden_out(name, tracking)
    mine_out = util.format_out_values(flat, scores, squares)
    for estimator in mine_out:
        assert mine_out[estimator]["mean"] == ref_out[estimator]["mean"]
        assert mine_out[estimator]["std dev"] == ref_out[estimator]["std dev"]
    # the oracle: exact integers, and every counter
    o_scores, o_squares, o_counters, status = util.oracle_problem(flat).run(threads=4)
    assert status == 0
    assert np.array_equal(scores, o_scores) and np.array_equal(squares, o_squares)
    for key, value in o_counters.items():
        assert counters[key] == value, key


def test_batches_and_ranks_chain_exactly():
    """Histories [0,N) in one call == any split into batches accumulated into the same buffers (what a
    multi-GPU run does): Scorable::operator+= semantics, exact integers."""
    flat = util.flat_from_xml(util.deck_text("fissile_slab", "surface", histories=30000))
    world, src, est = _product(flat)
    whole = world.fixed_source_run(src, est, 1, 0, 30000, secondary_capacity=256)
    scores = np.zeros(est.total_bins)
    squares = np.zeros(est.total_bins)
    total = {}
    for first, n in ((0, 1), (1, 9999), (10000, 12345), (22345, 7655)):
        _, _, c = world.fixed_source_run(src, est, 1, first, n, secondary_capacity=256, scores=scores, square_scores=squares)
        for k, v in c.items():
            total[k] = total.get(k, 0) + v
    assert np.array_equal(scores, whole[0]) and np.array_equal(squares, whole[1])
    assert total == whole[2]


def test_result_independent_of_launch_shape():
    flat = util.flat_from_xml(util.deck_text("three_shells", "delta"))
    world, src, est = _product(flat)
    a = world.fixed_source_run(src, est, 1, 0, 50000, tracking=1, blocks_per_sm=1)
    b = world.fixed_source_run(src, est, 1, 0, 50000, tracking=1, blocks_per_sm=0)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]


def test_empty_and_tiny_inputs():
    flat = util.flat_from_xml(util.deck_text("leakage_sphere", "surface"))
    world, src, est = _product(flat)
    s, q, c = world.fixed_source_run(src, est, 1, 0, 0)
    assert s.tolist() == [0.0] and c["n_histories"] == 0
    s, q, c = world.fixed_source_run(src, est, 1, 0, 1)
    o = util.oracle_problem(flat).run(1)
    assert s.tolist() == o[0].tolist() and c["n_events"] == o[2]["n_events"] == 1
    # seeds beyond 2^31 - 1 wrap like std::minstd_rand's seeding does
    big = 2147483647 * 3 + 5
    s, q, c = world.fixed_source_run(src, est, big, 0, 2000)
    o = util.oracle_problem(flat).run(2000, seed0=big)
    assert s.tolist() == o[0].tolist() and q.tolist() == o[1].tolist()


def test_capacity_overflow_is_reported_not_hidden():
    flat = util.flat_from_xml(util.deck_text("fissile_slab", "surface"))
    world, src, est = _product(flat)
    with pytest.raises(capi.MinimcError) as err:
        world.fixed_source_run(src, est, 1, 0, 20000, secondary_capacity=1)
    assert err.value.status == capi.ERR_CAPACITY


def test_lost_particle_is_reported():
    """A world whose cells do not cover space: World::FindCellContaining throws in the reference."""
    flat = util.flat_from_xml(util.deck_text("leakage_sphere", "surface"))
    w = dict(flat["world"])
    # drop the void cell outside the sphere
    w["cell_material"] = w["cell_material"][:1]
    w["cell_surface_begin"] = w["cell_surface_begin"][:2]
    w["cell_surface_index"] = w["cell_surface_index"][:1]
    w["cell_surface_sense"] = w["cell_surface_sense"][:1]
    w["cell_field_kind"] = w["cell_field_kind"][:1]
    w["cell_field_param"] = w["cell_field_param"][:6]
    world = capi.World(capi.FlatWorld(**w))
    with pytest.raises(capi.MinimcError) as err:
        world.fixed_source_run(util.product_source(flat), util.product_estimators(flat), 1, 0, 1000)
    assert err.value.status == capi.ERR_LOST_PARTICLE


def test_full_size_properties_m1():
    """BASELINE config C1 at 10^8 histories: size-independent properties.  One-group infinite medium with
    c = 0.25: every history ends in a capture, collisions per history -> 1/(1-c) = 4/3, nothing leaks."""
    flat = util.flat_from_xml(util.deck_text("critical", "surface"))
    world, src, _ = _product(flat)
    est = capi.Estimators([{"surface": 0}])
    n = 100_000_000
    scores, squares, c = world.fixed_source_run(src, est, 1, 0, n)
    assert scores[0] == 0 and squares[0] == 0
    assert c["n_histories"] == n and c["n_births"] == n and c["n_crossings"] == 0 and c["n_events"] == c["n_collisions"]
    mean = c["n_collisions"] / n
    sigma = np.sqrt(0.25 / 0.75 ** 2 / n)  # geometric distribution
    assert abs(mean - 4 / 3) < 5 * sigma
    # additivity over a split at an arbitrary point (what ranks do)
    _, _, c1 = world.fixed_source_run(src, est, 1, 0, 33_333_333)
    _, _, c2 = world.fixed_source_run(src, est, 1, 33_333_333, n - 33_333_333)
    assert c1["n_events"] + c2["n_events"] == c["n_events"]


def test_full_size_leakage_probability():
    """test_FixedSource.cpp:13-26 at 10^7 histories: leakage = e^-1 within 3 sigma (test/Statistics.hpp)."""
    flat = util.flat_from_xml(util.deck_text("leakage_sphere", "surface"))
    world, src, est = _product(flat)
    n = 10_000_000
    scores, squares, c = world.fixed_source_run(src, est, 1, 0, n)
    p = np.exp(-1)
    assert abs(scores[0] / n - p) / p < 3 * np.sqrt((1 - p) / (p * n))
    assert squares[0] == scores[0]  # at most one leak per history
    assert c["n_crossings"] == scores[0] == c["n_scores"]
