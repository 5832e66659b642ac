"""MMC_RNG_COUNTER (north_star: "per-particle counter-based RNG"): Philox-2x32-10 per particle
instead of the reference's sequential std::minstd_rand (BasicTypes.hpp:27).  The reference has no such mode, so the
contract is statistical -- every bin of every BASELINE configuration within the combined 3 sigma of the minstd run
(which is bit-exact with the reference) -- plus what a counter-based generator is for: the result does not depend on
how the batch is split over launches, schedules or GPUs, because a particle's stream is a pure function of its
history's seed and its ancestry."""
import numpy as np
import pytest

from minimc_b200 import capi, ce_decks, decks

pytestmark = pytest.mark.gpu

M, W = 0xD256D193, 0x9E3779B9


def philox2x32_10(c0, c1, key):
    """Philox-2x32-10 restated from Salmon et al., SC'11 (pinned below by the Random123 known-answer vectors)."""
    for _ in range(10):
        p = M * c0
        c0, c1 = ((p >> 32) ^ key ^ c1) & 0xFFFFFFFF, p & 0xFFFFFFFF
        key = (key + W) & 0xFFFFFFFF
    return c0, c1


def test_philox_known_answers():
    assert philox2x32_10(0, 0, 0) == (0xFF1DAE59, 0x6CD10DF2)
    assert philox2x32_10(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF) == (0x2C3F628B, 0xAB4FD7AD)
    assert philox2x32_10(0x243F6A88, 0x85A308D3, 0x13198A2E) == (0xDD7CE038, 0xF62A4C12)


def test_device_generator_equals_philox():
    """mmc_test_device_math fn 16 + 4: the first canonical double of the stream whose id is x[i]."""
    seeds = np.array([0, 1, 2, 12345, 2 ** 31 - 1, 2 ** 32 + 7, 2 ** 40 + 3, 2 ** 52 + 11], dtype=np.float64)
    got, _ = capi.device_math(16 + 4, seeds)
    for s, u in zip(seeds, got):
        s = int(s)
        c0, c1 = philox2x32_10(0, s >> 32, s & 0xFFFFFFFF)
        assert u == ((c0 << 32 | c1) >> 11) * 2.0 ** -53
    # uniformity of the first draw over consecutive stream ids (the ids source particles get)
    u, _ = capi.device_math(16 + 4, np.arange(1, 400_001, dtype=np.float64))
    assert abs(u.mean() - 0.5) < 4 / np.sqrt(12 * len(u)) and 0 <= u.min() and u.max() < 1
    hist, _ = np.histogram(u, bins=100, range=(0, 1))
    chi2 = ((hist - len(u) / 100) ** 2 / (len(u) / 100)).sum()
    assert 50 < chi2 < 160  # 99 degrees of freedom


@pytest.fixture(scope="module")
def full_tables(tmp_path_factory):
    d = tmp_path_factory.mktemp("full_tables_ctr")
    ce_decks.generate_tables(d, "full")
    return d


def _baseline_configs(full_tables):
    n = 300_000
    return {
        "C1 multigroup_critical": decks.critical(histories=4 * n, estimators=[{"name": "leakage", "surface": "sphere"}]),
        "M2 three_shells": decks.DECKS["three_shells"](histories=2 * n, estimators=decks.THREE_SHELL_ESTIMATORS),
        "C2 single_zone": ce_decks.single_zone_benchmark_deck(full_tables, histories=n, threads=1),
        "C3 multi_zone": ce_decks.multi_zone_deck(full_tables, histories=n, n_energy_bins=201),
        "C4 broomstick": ce_decks.broomstick_deck(full_tables, histories=4 * n, n_energy_bins=238, n_cosine_bins=182),
        "C5 continuous_temperature": ce_decks.continuous_temperature_deck(full_tables, histories=n, n_energy_bins=201),
        "fissile_slab (fixed source, banked secondaries)": decks.DECKS["fissile_slab"](histories=n),
    }


@pytest.mark.parametrize("name", ["C1 multigroup_critical", "M2 three_shells", "C2 single_zone", "C3 multi_zone", "C4 broomstick",
                                  "C5 continuous_temperature", "fissile_slab (fixed source, banked secondaries)"])
def test_counter_mode_agrees_with_minstd_within_3_sigma(full_tables, name):
    text = _baseline_configs(full_tables)[name]
    results = {}
    for mode in (capi.RNG_MINSTD_COMPAT, capi.RNG_COUNTER):
        drv = capi.Driver(text=text)
        drv.set_options(rng_mode=mode, secondary_capacity=256)
        scores, squares = drv.solve()
        c = drv.counters()
        assert c["n_lost"] == c["n_physics_errors"] == c["n_capacity_overflow"] == 0
        results[mode] = (scores, squares, c)
    (s0, q0, c0), (s1, q1, c1) = results[capi.RNG_MINSTD_COMPAT], results[capi.RNG_COUNTER]
    n = c0["n_histories"]
    assert c1["n_histories"] == n
    assert not np.array_equal(s0, s1) or s0.sum() == 0  # other random numbers
    # Scorable::GetScoreAsString (Scorable.cpp:51-70): mean = s / N, std dev of the mean = sqrt(q - s^2 / N) / N
    mean0, mean1 = s0 / n, s1 / n
    var = (np.maximum(q0 - s0 * s0 / n, 0) + np.maximum(q1 - s1 * s1 / n, 0)) / (n * n)
    populated = (s0 + s1) >= 40  # bins with enough hits for a normal approximation
    z = (mean1 - mean0)[populated] / np.sqrt(var[populated])
    if name.startswith("C1"):  # (nothing leaks from the sphere of radius 1e10: the event counts below are the check)
        assert len(z) == 0 and s0.sum() == s1.sum() == 0
        z = np.zeros(1)
    # no bin far out: the largest of n standard normal deviates is about sqrt(2 ln n) (4.3 for broomstick's ~10^4 bins)
    assert np.abs(z).max() < np.sqrt(2 * np.log(max(len(z), 2))) + 1.5, (name, np.abs(z).max(), len(z))
    assert (np.abs(z) > 3).mean() <= 0.01 + 3 / max(len(z), 1), name  # 0.27 % expected beyond 3 sigma
    assert 0.75 < (z * z).mean() < 1.3 or len(z) < 30, (name, (z * z).mean())
    # event counts per history agree too (3 sigma of a Poisson-like count, generously)
    for key in ("n_events", "n_collisions", "n_crossings"):
        assert abs(c0[key] - c1[key]) < 6 * np.sqrt(max(c0[key], 1)) + 6 * np.sqrt(n), key


def test_counter_mode_is_independent_of_splits_and_schedules(full_tables):
    """The same batch as one launch, as three shards (what three GPUs would run), under the fused and the event-split
    schedule with few slots: identical integer tallies and counters."""
    cases = ((ce_decks.single_zone_benchmark_deck(full_tables, histories=60_000, threads=1),
              ((capi.SCHEDULE_AUTO, 0, 0), (capi.SCHEDULE_FUSED, 0, 0), (capi.SCHEDULE_EVENT, 4099, 0))),
             (decks.DECKS["fissile_slab"](histories=90_000), ((capi.SCHEDULE_AUTO, 0, 0), (capi.SCHEDULE_FUSED, 0, 1))))
    for text, schedules in cases:
        runs = []
        for schedule, slots, blocks_per_sm in schedules:
            drv = capi.Driver(text=text)
            drv.set_options(rng_mode=capi.RNG_COUNTER, schedule=schedule, event_slots=slots, secondary_capacity=256,
                            blocks_per_sm=blocks_per_sm)
            scores, squares = drv.solve()
            runs.append((scores, squares, drv.counters()))
        total = np.zeros_like(runs[0][0])
        for rank in range(3):
            part = capi.Driver(text=text)
            part.set_options(rng_mode=capi.RNG_COUNTER, secondary_capacity=256)
            part.set_shard(rank, 3)
            total += part.solve()[0]
        for scores, squares, c in runs[1:]:
            assert np.array_equal(scores, runs[0][0]) and np.array_equal(squares, runs[0][1]) and c == runs[0][2]
        assert np.array_equal(total, runs[0][0])


def test_counter_mode_keigenvalue():
    """k-eigenvalue generations in counter mode: the analytic k_inf = 0.81 within 4 sigma, both estimators."""
    drv = capi.Driver(text=decks.k_infinite(histories=200_000, inactive=5, active=40))
    drv.set_options(rng_mode=capi.RNG_COUNTER)
    drv.solve()
    k_mean, k_std, k_cycle = drv.keff()
    kc_mean, kc_std, _, _ = drv.k_collision()
    assert abs(k_mean - 0.81) < 4 * k_std and abs(kc_mean - 0.81) < 4 * kc_std and 0 < kc_std < k_std < 2e-3
    ref = capi.Driver(text=decks.k_infinite(histories=200_000, inactive=5, active=40))
    ref.solve()
    assert not np.array_equal(ref.keff()[2], k_cycle)
