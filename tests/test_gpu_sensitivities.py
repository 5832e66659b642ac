"""Differential-operator sensitivities (Perturbation.cpp:68-84, Sensitivity.cpp:44-58,
Particle.cpp:46-53) on the CUDA path against the reference's own output.

The reference has no test or deck for them (parity unpinned by its tests); the fixtures in tests/golden/sens/ are
.out files the reference binary wrote (tests/golden/generate.py sens) for a three-shell multigroup deck and a
continuous-energy + S(a,b) slab, both tracking modes, one worker thread.  The estimators' own (integer) tallies must
be identical; a sensitivity is a real-valued sum whose last bits depend on summation order -- in the reference on
which worker ran which history -- so means and standard deviations must agree to the 7 printed digits up to one
unit in the last place (relative 2e-6)."""
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
    for n, t, text in util.sensitivity_cases(tables, **kw):
        if (n, t) == (name, tag):
            return text
    raise KeyError((name, tag))


def _compare(mine_text, ref_text):
    batch_m, mine = port_py.parse_out(mine_text)
    batch_r, ref = port_py.parse_out(ref_text)
    assert batch_m == batch_r and list(mine) == list(ref)
    n_sens = 0
    for name in ref:
        for key in ("mean", "std dev"):
            m, r = np.array(mine[name][key], float), np.array(ref[name][key], float)
            if "::" in name:  # a Sensitivity: real-valued
                n_sens += 1
                assert np.allclose(m, r, rtol=2e-6, atol=0), (name, key, m, r)
                assert np.array_equal(m == 0, r == 0), (name, key)
            else:             # an Estimator: integer tallies, identical text
                assert mine[name][key] == ref[name][key], (name, key)
    assert n_sens > 0


@pytest.mark.parametrize("name,tag", util.SENS_CASE_IDS)
def test_sensitivities_match_reference_out(tables, name, tag):
    drv = capi.Driver(text=_case(tables, name, tag))
    drv.set_options(secondary_capacity=256)
    drv.solve()
    c = drv.counters()
    assert c["n_lost"] == c["n_physics_errors"] == c["n_capacity_overflow"] == 0
    _compare(drv.output(), (util.GOLDEN / "sens" / f"{name}__{tag}.out").read_text())


@pytest.mark.parametrize("name,tag", util.SENS_CASE_IDS)
def test_sensitivities_fresh_seed_against_live_reference(tables, tmp_path, name, tag):
    if not port_py.ref_available():
        pytest.skip("oracle/_ref/ref_harness was not built")
    path = tmp_path / "deck.xml"
    path.write_text(_case(tables, name, tag, histories=5000, seed=97531))
    drv = capi.Driver(path)
    drv.set_options(secondary_capacity=256)
    drv.solve()
    _compare(drv.output(), port_py.ref_run(path)[0])


def test_sensitivity_sharding_adds_up(tables):
    """Rank shards of the histories add up to the single-rank result: exactly for the estimators, to rounding for
    the sensitivities (Estimator::operator+=, Estimator.cpp:59-71)."""
    text = _case(tables, "sensitivity_shells", "surface")
    whole = capi.Driver(text=text)
    whole.solve()
    _, ref = port_py.parse_out(whole.output())
    parts = []
    for rank in range(3):
        part = capi.Driver(text=text)
        part.set_shard(rank, 3)
        part.solve()
        parts.append(port_py.parse_out(part.output())[1])
    for name in ref:
        total = sum(np.array(p[name]["mean"], float) for p in parts)
        assert np.allclose(total, np.array(ref[name]["mean"], float), rtol=1e-5, atol=1e-12), name
