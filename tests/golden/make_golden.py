"""Generates tests/golden/c1_idle_set.npz — BASELINE config #1 ("dry-run, 100-pod synthetic DCGM
fixture, 30-min window, CPU aggregation").

The reference holds no fixture for this path and cannot be built or imported here (Rust + a
remote Prometheus server), so the golden idle set is produced by the C oracle
and accepted only if the independent numpy oracle reproduces it bit for bit.  Run from the
repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_c, oracle_np  # noqa: E402

SEED, P, G, T, THR = 0x5EED0001, 100, 4, 1800, 150.0


def main():
    u = oracle_c.synth_fill(SEED, 0, 0, P, G, T)
    w = oracle_c.synth_fill(SEED, 1, 0, P, G, T)
    e = oracle_c.synth_eligible(SEED, 0, P)
    assert np.all((u == oracle_np.synth_fill(SEED, 0, 0, P, G, T)) | np.isnan(u))
    a = oracle_c.decide(u, None, e)
    b = oracle_np.decide(u, None, e)
    ap = oracle_c.decide(u, w, e, power_threshold=THR)
    bp = oracle_np.decide(u, w, e, power_threshold=THR)
    for x, y in ((a, b), (ap, bp)):
        assert np.array_equal(x["decision_bits"], y["decision_bits"])
        assert np.array_equal(x["candidate_bits"], y["candidate_bits"])
        assert x["n_series"] == y["n_series"]
    idle = np.flatnonzero(oracle_np.unpack_bits(a["decision_bits"], P))
    idle_p = np.flatnonzero(oracle_np.unpack_bits(ap["decision_bits"], P))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c1_idle_set.npz")
    np.savez(out, seed=np.uint64(SEED), P=P, G=G, T=T, power_threshold=THR,
             decision_bits=a["decision_bits"], candidate_bits=a["candidate_bits"],
             n_series=a["n_series"], idle_pods=idle,
             decision_bits_power=ap["decision_bits"], candidate_bits_power=ap["candidate_bits"],
             n_series_power=ap["n_series"], idle_pods_power=idle_p,
             series_max=a["series_max"])
    print(f"wrote {out}: {idle.size} idle pods ({idle_p.size} with the power veto) of {P}")
    print("idle pods:", idle.tolist())
    print("idle pods (power veto on):", idle_p.tolist())


if __name__ == "__main__":
    main()
