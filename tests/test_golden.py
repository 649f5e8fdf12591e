"""CPU: both oracles reproduce the checked-in config-#1 idle set (tests/golden/make_golden.py)."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c1_idle_set.npz")


def test_c1_idle_set(oracle_c, oracle_np):
    g = np.load(GOLD)
    seed, P, G, T = int(g["seed"]), int(g["P"]), int(g["G"]), int(g["T"])
    assert (seed, P, G, T) == (0x5EED0001, 100, 4, 1800)
    thr = float(g["power_threshold"])
    for orc in (oracle_c, oracle_np):
        u = orc.synth_fill(seed, 0, 0, P, G, T)
        w = orc.synth_fill(seed, 1, 0, P, G, T)
        e = orc.synth_eligible(seed, 0, P)
        r = orc.decide(u, None, e)
        assert np.array_equal(r["decision_bits"], g["decision_bits"])
        assert np.array_equal(r["candidate_bits"], g["candidate_bits"])
        assert r["n_series"] == int(g["n_series"])
        rp = orc.decide(u, w, e, power_threshold=thr)
        assert np.array_equal(rp["decision_bits"], g["decision_bits_power"])
        assert rp["n_series"] == int(g["n_series_power"])
    s = oracle_c.decide_synth(seed, 0, P, G, T, use_elig=True)
    assert np.array_equal(s["decision_bits"], g["decision_bits"])
    assert 0 < g["idle_pods_power"].size < g["idle_pods"].size < P
