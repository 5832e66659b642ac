"""K-eigenvalue power iteration on the GPU (C++ host KEigenvalue::Solve over mmc_generation_run /
mmc_bank_resample).

The reference's KEigenvalue::Solve is a stub (KEigenvalue.cpp:36-62): PARITY WITH THE REFERENCE IS
UNPINNED for k-eigenvalue.  What is pinned: (1) the CUDA path against the oracle's independent CPU statement of the
same definition (oracle/port.cpp orc_keigenvalue_run) -- k of every cycle, every bank size, all tallies and counters
exact, since the definition is order-based and every step is integer or bit-exact fp64; (2) analytic k-infinity."""
import numpy as np
import pytest

import util
from minimc_b200 import capi, decks

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(decks.KDECKS))
@pytest.mark.parametrize("tracking", ["surface", "delta"])
def test_power_iteration_matches_oracle_exactly(name, tracking):
    text = decks.KDECKS[name](tracking=util.TRACKING[tracking])
    flat = util.flat_from_xml(text)
    o_scores, o_squares, o_k, o_sizes, o_counters, status = util.oracle_problem(flat).keigenvalue()
    assert status == 0
    drv = capi.Driver(text=text)
    scores, squares = drv.solve()
    k_mean, k_std, k_cycle = drv.keff()
    assert np.array_equal(k_cycle, o_k)
    assert np.array_equal(scores, o_scores) and np.array_equal(squares, o_squares)
    c = drv.counters()
    for key in ("n_histories", "n_births", "n_events", "n_collisions", "n_crossings", "n_virtual", "n_scores",
                "n_secondaries"):
        assert c[key] == o_counters[key], key
    assert c["n_banked"] == int(o_sizes.sum())
    run = flat["run"]
    active = o_k[run["inactive"]:]
    assert k_mean == pytest.approx(active.mean(), rel=1e-15)
    assert k_std == pytest.approx(active.std(ddof=1) / np.sqrt(len(active)), rel=1e-12, abs=1e-300)


def test_k_unity_is_exact():
    """nubar = 1, no capture: every source particle ends in one fission with exactly one site."""
    drv = capi.Driver(text=decks.k_unity(histories=300_000, inactive=1, active=5))
    drv.solve()
    k_mean, k_std, k_cycle = drv.keff()
    assert k_cycle.tolist() == [1.0] * 6 and k_mean == 1.0 and k_std == 0.0
    c = drv.counters()
    assert c["n_banked"] == 6 * 300_000 == c["n_histories"]


def test_k_infinite_agrees_with_analytic_value():
    """k_inf = nubar * Sigma_f / Sigma_a = 0.81; 40 active cycles of 2*10^5: within 4 sigma of the cycle scatter."""
    drv = capi.Driver(text=decks.k_infinite(histories=200_000, inactive=5, active=40))
    drv.solve()
    k_mean, k_std, k_cycle = drv.keff()
    assert len(k_cycle) == 45 and k_std > 0
    assert abs(k_mean - 0.81) < 4 * k_std
    assert k_std < 2e-3


def test_out_file_normalisation_counts_active_histories():
    """EstimatorSet.total_weight = batchsize * active cycles, so means are per source particle of active cycles."""
    text = decks.k_slab(histories=10_000, inactive=2, active=3)
    drv = capi.Driver(text=text)
    scores, _ = drv.solve()
    from oracle import port_py
    batch, parsed = port_py.parse_out(drv.output())
    assert batch == 10_000
    side = [float(v) for v in parsed["side"]["mean"]]
    assert side[0] == pytest.approx(scores[-1] / 30_000, rel=1e-6)


def test_fission_bank_overflow_is_reported():
    text = decks.k_infinite(histories=20_000, inactive=0, active=1).replace("<nubar>2.43</nubar>", "<nubar>40</nubar>")
    drv = capi.Driver(text=text)
    with pytest.raises(capi.MinimcError) as e:
        drv.solve()
    assert e.value.status == capi.ERR_CAPACITY


def test_python_distributed_driver_equals_cpp_host_at_one_rank():
    """minimc_b200.distributed.KEigenvalue (the multi-GPU orchestration; here world_size 1) and the C++ host's
    KEigenvalue::Solve run the same device entry points: identical k per cycle and tallies."""
    import torch
    from minimc_b200 import distributed
    text = decks.k_slab(histories=30_000, inactive=2, active=4)
    flat = util.flat_from_xml(text)
    drv = capi.Driver(text=text)
    scores, squares = drv.solve()
    _, _, k_cycle = drv.keff()
    world = util.product_world(flat)
    run = flat["run"]
    kd = distributed.KEigenvalue(world, util.product_source(flat), util.product_estimators(flat), run["histories"],
                                 run["inactive"], run["active"], tracking=run["tracking"])
    out = kd.solve(device=torch.device("cuda", 0))
    assert np.array_equal(out["k_cycle"], k_cycle)
    assert np.array_equal(out["scores"], scores) and np.array_equal(out["square_scores"], squares)
