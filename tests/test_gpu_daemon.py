"""GPU: daemon mode of the `gpu-pruner` binary with the window RESIDENT in HBM (main.rs:286-330 --daemon-mode /
--check-interval; BASELINE config #5).  First tick = the full range query parsed on the GPU into the resident ring;
later ticks ask only for what was scraped since the previous tick, parse that slice into the ring and rescan
(gpr_resident_advance + gpr_text_parse(GPR_TEXT_RESIDENT) + gpr_decide_resident).  Every tick's verdict line must
equal what the oracle decides on the window a fresh full-range query of that tick returns — series that appear,
disappear and age out, new pods, a tick without a usable slice, all included."""
import json
import os
import random
import subprocess

import numpy as np
import pytest

import hostlib as H
import ticks as TK
from test_resident_ticks import _series

pytestmark = pytest.mark.gpu


def _expected(root, k, duration_min, thr, oracle_np):
    """fresh full-range ingest of tick k on the CPU -> oracle -> exact sum-by: (n_series, unique pods)"""
    d = os.path.join(root, "tick-%04d" % k, "full")
    q = json.load(open(os.path.join(d, "query.json")))
    load = lambda n: json.load(open(os.path.join(d, n))) if os.path.exists(os.path.join(d, n)) else None
    u, w, meta = H.ingest(load("util.json"), load("prof.json"), load("power.json") if thr else None,
                          duration_min=duration_min, step=q["step"], t_end=q["end"])
    r = oracle_np.decide(u, w, power_threshold=thr)
    cb, db, counts, _ = H.resolve_groups(r["series_max"], r["candidate_bits"], r["decision_bits"],
                                         (r["n_series"], r["n_candidates"], r["n_decisions"]), veto_bits=r["veto_bits"])
    names = {(p["name"], p["namespace"]) for i, p in enumerate(meta["pods"]) if oracle_np.unpack_bits(cb, len(meta["pods"]))[i]}
    return counts[0], len(names)


def _run(root, n_ticks, duration_min, *extra):
    cmd = [H.BIN, "--prometheus-url", f"file://{root}", "-d", "-c", "0", "--max-ticks", str(n_ticks), "-t", str(duration_min),
           "-l", "json", *extra]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-3000:]
    msgs = [json.loads(l)["fields"]["message"] for l in p.stderr.splitlines() if l.startswith("{")]
    return msgs


def _scenario(tmp_path, seed, power):
    rng = random.Random(seed)
    N, step, interval, dur = 120, 2, 30, 2
    t0 = 1_700_000_000
    times = [t0 + N + k * interval for k in range(9)]
    horizon = times[-1] + 5
    store = [_series(rng, f"pod-{p}", g, t0, horizon, step, rng.choice(["idle", "idle", "busy"])) for p in range(40) for g in range(2)]
    store.append(_series(rng, "leaver", 0, t0, times[2] - 3, step, "idle"))
    store.append(_series(rng, "joiner", 0, times[3] + 1, horizon, step, "idle"))
    store.append(_series(rng, "pod-0", 1, times[5] + 1, horizon, step, "idle", UUID="GPU-late"))   # third slot: rebuild
    if power:
        store += [_series(rng, f"pod-{p}", 0, t0, horizon, step, "x", metric="DCGM_FI_DEV_POWER_USAGE") for p in range(40)]
    TK.write_ticks(str(tmp_path), lambda k: store, times, N, step, with_power=power, skip_delta={7})
    return str(tmp_path), len(times), dur


@pytest.mark.parametrize("power", [False, True], ids=["util", "util+power"])
def test_resident_daemon_ticks_equal_fresh_queries(tmp_path, power, oracle_np):
    root, n, dur = _scenario(tmp_path, 4 + power, power)
    extra = ("--power-threshold", "150") if power else ()
    msgs = _run(root, n, dur, *extra)
    verdicts = [m for m in msgs if m.startswith("Query returned")]
    assert len(verdicts) == n
    for k, v in enumerate(verdicts):
        n_series, n_pods = _expected(root, k, dur, 150.0 if power else None, oracle_np)
        assert v == f"Query returned {n_series} series across {n_pods} unique pods", (k, v)
    ingests = [m for m in msgs if m.startswith("Device ingest")]
    appended = ["appended to the resident" in m for m in ingests]
    # tick 0 full; the late duplicate of pod-0 needs a third GPU slot at tick 6 and tick 7 has no slice: full again
    assert appended == [False, True, True, True, True, True, False, False, True], ingests
    assert any(m.startswith("Resident window rebuilt from the full range") for m in msgs)
    assert sum("into a resident" in m for m in ingests) == 3


def test_cpu_ingest_in_daemon_mode_takes_the_full_range_every_tick(tmp_path, oracle_np):
    root, n, dur = _scenario(tmp_path, 9, False)
    env_msgs = None
    cmd = [H.BIN, "--prometheus-url", f"file://{root}", "-d", "-c", "0", "--max-ticks", str(n), "-t", str(dur), "-l", "json"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, GPR_INGEST="cpu"))
    assert p.returncode == 0
    msgs = [json.loads(l)["fields"]["message"] for l in p.stderr.splitlines() if l.startswith("{")]
    verdicts = [m for m in msgs if m.startswith("Query returned")]
    for k, v in enumerate(verdicts):
        n_series, n_pods = _expected(root, k, dur, None, oracle_np)
        assert v == f"Query returned {n_series} series across {n_pods} unique pods", (k, v)
    assert len(verdicts) == n and not any("resident" in m for m in msgs)


def test_prof_metric_through_the_device_ingest(tmp_path, oracle_np):
    """a4 / VERDICT r1 #6: DCGM_FI_PROF_GR_ENGINE_ACTIVE ratios (shortest-round-trip doubles, up to 17 digits) go
    through the binary's device ingest without a single span falling back to the CPU, and PROF shadows UTIL on
    identical label sets (query.promql.j2:10-20): the verdict equals the oracle's on the CPU-ingested window"""
    rng = random.Random(77)
    N, step, dur = 120, 1, 2
    t0 = 1_700_000_000
    t_end = t0 + N
    store = []
    for p in range(30):
        for g in range(2):
            lab = TK.labels(f"pod-{p}", g)
            busy_util = rng.random() < 0.5
            store.append(("DCGM_FI_DEV_GPU_UTIL", lab, [(t, rng.choice([0, 0, 40]) if busy_util else 0) for t in range(t0, t_end + 1)]))
            r = rng.random()
            if r < 0.4:        # PROF with the identical label set: it decides, whatever UTIL says
                idle = rng.random() < 0.5
                store.append(("DCGM_FI_PROF_GR_ENGINE_ACTIVE", lab,
                              [(t, 0.0 if idle else rng.random()) for t in range(t0, t_end + 1)]))
            elif r < 0.5:      # PROF with another label: both survive the `or`, `sum by` adds them
                store.append(("DCGM_FI_PROF_GR_ENGINE_ACTIVE", dict(lab, profiled="yes"),
                              [(t, rng.random() * 1e-7) for t in range(t0, t_end + 1)]))   # Go prints these as 5.1e-08
    TK.write_ticks(str(tmp_path), lambda k: store, [t_end], N, step)
    msgs = _run(str(tmp_path), 1, dur)
    note = [m for m in msgs if m.startswith("Device ingest")][0]
    assert "parsed on the GPU" in note and "(0 re-parsed on the CPU, 0 rows patched" in note, note
    n_series, n_pods = _expected(str(tmp_path), 0, dur, None, oracle_np)
    assert f"Query returned {n_series} series across {n_pods} unique pods" in msgs
    assert 0 < n_pods < 30
