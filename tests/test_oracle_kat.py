"""CPU: both oracle restatements against the hand-derived known-answer vectors and each other."""
import numpy as np
import pytest

import kat


KATS = kat.all_kats()


@pytest.mark.parametrize("k", KATS, ids=[k.name for k in KATS])
def test_kat_c_oracle(k, oracle_c):
    r = oracle_c.decide(k.util, k.power, k.eligible, k.created_ts, k.cutoff_ts, k.power_threshold)
    assert np.array_equal(r["candidate_bits"], kat.expected_bits(k.candidate)), k.why
    assert np.array_equal(r["decision_bits"], kat.expected_bits(k.decision)), k.why
    assert r["n_candidates"] == sum(k.candidate) and r["n_decisions"] == sum(k.decision)
    if k.series_max is not None:
        assert kat.smax_equal(r["series_max"], k.series_max)
    if k.n_series is not None:
        assert r["n_series"] == k.n_series
    if "value" in k.extra:  # reported value = max / 100 (query.promql.j2:20, lib.rs:184)
        assert float(r["series_max"][0, 0]) / 100 == pytest.approx(k.extra["value"])


@pytest.mark.parametrize("k", KATS, ids=[k.name for k in KATS])
def test_kat_numpy_oracle(k, oracle_np):
    r = oracle_np.decide(k.util, k.power, k.eligible, k.created_ts, k.cutoff_ts, k.power_threshold)
    assert np.array_equal(r["candidate_bits"], kat.expected_bits(k.candidate)), k.why
    assert np.array_equal(r["decision_bits"], kat.expected_bits(k.decision)), k.why
    if k.series_max is not None:
        assert kat.smax_equal(r["series_max"], k.series_max)
    if k.n_series is not None:
        assert r["n_series"] == k.n_series


def test_max_over_time_first_sample_rule(oracle_c):
    # Prometheus funcMaxOverTime: start from the first sample, replace on `cur > max`
    assert oracle_c.max_over_time([3, 1, 2]) == 3.0
    assert oracle_c.max_over_time([np.nan, -5, -7]) == -5.0
    assert np.isnan(oracle_c.max_over_time([np.nan] * 4))
    assert oracle_c.max_over_time([0.0]) == 0.0
    assert np.isnan(oracle_c.max_over_time([]))


def _random_case(rng, P, G, T, with_power, with_gates):
    u = rng.choice(np.array([0, 0, 0, 0, 1, 50, 100, np.nan, -0.0, -3], np.float32), size=(P, G, T),
                   p=[.55, .1, .1, .1, .01, .01, .01, .1, .01, .01])
    # make a good share of rows fully idle so verdicts are mixed
    idle_rows = rng.random((P, G)) < 0.45
    u[idle_rows] = np.where(rng.random((int(idle_rows.sum()), T)) < 0.05, np.nan, 0).astype(np.float32)
    kw = {}
    if with_power:
        w = rng.choice(np.array([40, 60, 149.99, np.nan], np.float32), size=(P, G, T),
                       p=[.5, .44, .02, .04])
        hot = np.flatnonzero(rng.random(P) < 0.4)       # these pods get one hot sample somewhere
        w[hot, rng.integers(0, G, hot.size), rng.integers(0, T, hot.size)] = rng.choice(
            np.array([150, 150.01, 400], np.float32), size=hot.size)
        kw["power"] = w
        kw["power_threshold"] = 150.0
    if with_gates:
        kw["eligible"] = (rng.random(P) < 0.9).astype(np.uint8)
        kw["created_ts"] = rng.integers(1000, 2000, P).astype(np.int64)
        kw["cutoff_ts"] = 1500
    return u, kw


@pytest.mark.parametrize("P,G,T", [(1, 1, 1), (7, 3, 5), (64, 4, 33), (257, 8, 100), (1000, 4, 180)])
@pytest.mark.parametrize("with_power", [False, True])
@pytest.mark.parametrize("with_gates", [False, True])
def test_two_restatements_agree(P, G, T, with_power, with_gates, oracle_c, oracle_np):
    rng = np.random.default_rng(P * 1000 + G * 10 + T + with_power * 7 + with_gates * 3)
    u, kw = _random_case(rng, P, G, T, with_power, with_gates)
    a = oracle_c.decide(u, **kw)
    b = oracle_np.decide(u, **kw)
    for key in ("decision_bits", "candidate_bits"):
        assert np.array_equal(a[key], b[key]), key
    for key in ("n_series", "n_candidates", "n_decisions"):
        assert a[key] == b[key], key
    assert kat.smax_equal(a["series_max"], b["series_max"])
    # verdicts are mixed, not degenerate
    if P >= 64:
        assert 0 < a["n_candidates"] < P


def test_threaded_oracle_matches_single(oracle_c):
    rng = np.random.default_rng(5)
    u, kw = _random_case(rng, 1003, 4, 64, True, True)
    a = oracle_c.decide(u, **kw)
    for n in (2, 3, 8, 33):
        b = oracle_c.decide(u, n_threads=n, **kw)
        assert np.array_equal(a["decision_bits"], b["decision_bits"])
        assert np.array_equal(a["candidate_bits"], b["candidate_bits"])
        assert (a["n_series"], a["n_candidates"], a["n_decisions"]) == (
            b["n_series"], b["n_candidates"], b["n_decisions"])
        assert kat.smax_equal(a["series_max"], b["series_max"])


def test_empty_window(oracle_c, oracle_np):
    u = np.zeros((0, 4, 16), np.float32)
    a = oracle_c.decide(u)
    b = oracle_np.decide(u)
    assert a["decision_bits"].size == 0 and b["decision_bits"].size == 0
    assert a["n_decisions"] == 0 and b["n_decisions"] == 0
