"""CPU: PromQL semantics -> (C++ ingest of the wire format) -> dense tensor -> oracle, end to end.

tests/promql_mini.py evaluates the reference's expression (/root/reference/gpu-pruner/src/
query.promql.j2:1-44) the way Prometheus does, on labelled instant vectors.  For random clusters —
multi-GPU pods, PROF + UTIL series, `sum by` duplicates, hosts with and without node_dmi_info,
scrape gaps, series that start late, unconvertible series, power draw around the threshold — the set
of pods it returns (after the Rust-side dedup, main.rs:416-437) must equal the candidates the dense
path produces: range-query wire format -> gpu-pruner_b200/host ingest -> oracle decision.
"""
import math
import random

import numpy as np
import pytest

import hostlib as H
import promql_mini as Q


def make_cluster(rng, honor_labels=False, t_eval=100_000, duration_min=2):
    T = duration_min * 60
    pl, nl, cl = ("pod", "namespace", "container") if honor_labels else (
        "exported_pod", "exported_namespace", "exported_container")
    db = []
    ts = list(range(t_eval - T - 30, t_eval + 1))           # a little history before the window too

    def pattern(kind):
        if kind == "idle":
            return [0.0] * len(ts)
        if kind == "busy":
            return [float(rng.choice([0, 0, 0, 37, 100])) for _ in ts]
        if kind == "burst":                                   # one sample somewhere in the window
            v = [0.0] * len(ts)
            v[rng.randrange(31, len(ts))] = 5.0
            return v
        if kind == "old_burst":                               # activity only BEFORE the window
            v = [0.0] * len(ts)
            v[rng.randrange(0, 30)] = 80.0
            return v
        raise AssertionError(kind)

    def emit(name, labels, vals, gap=0.0, start=0):
        smp = [(t, v) for i, (t, v) in enumerate(zip(ts, vals)) if i >= start and rng.random() >= gap]
        if smp:
            db.append(Q.series(name, labels, smp))

    hosts = [f"node-{i}" for i in range(4)]
    for h in hosts[:3]:                                       # node-3 has no DMI series
        db.append(Q.series("node_dmi_info", {"instance": h, "product_name": "DGX-B200"}, [(t_eval - 5, 1.0)]))
    n_pods = rng.randrange(6, 14)
    for p in range(n_pods):
        pod, ns = f"pod-{p}", rng.choice(["ml-team", "ml-team", "infra"])
        host = rng.choice(hosts)
        for g in range(rng.randrange(1, 5)):
            base = {"Hostname": host, "gpu": str(g), "modelName": rng.choice(["NVIDIA B200", "NVIDIA A100"]),
                    "UUID": f"GPU-{p}-{g}", pl: pod, nl: ns, cl: "main", "instance": host + ":9400", "job": "dcgm"}
            if not honor_labels:                              # Prometheus' own target labels ride along
                base.update(pod="dcgm-exporter-xyz", namespace="monitoring", container="exporter")
            kind = rng.choice(["idle", "idle", "busy", "burst", "old_burst"])
            start = rng.choice([0, 0, 0, rng.randrange(30, len(ts))])      # young series
            vals = pattern(kind)
            emit("DCGM_FI_DEV_GPU_UTIL", base, vals, gap=rng.choice([0, 0.05]), start=start)
            r = rng.random()
            if r < 0.3:        # PROF with the identical label set: wins the `or`
                pkind = rng.choice(["idle", "busy"])
                emit("DCGM_FI_PROF_GR_ENGINE_ACTIVE", base, [v / 100 for v in pattern(pkind)], start=start)
            elif r < 0.4:      # PROF with an extra label: both survive `or`, `sum by` adds them
                emit("DCGM_FI_PROF_GR_ENGINE_ACTIVE", dict(base, profiled="yes"),
                     [v / 100 for v in pattern(rng.choice(["idle", "busy"]))])
            if rng.random() < 0.15:   # `sum by` duplicate: same group, another UUID
                emit("DCGM_FI_DEV_GPU_UTIL", dict(base, UUID=f"GPU-{p}-{g}-b"), pattern(rng.choice(["idle", "busy"])))
            watts = rng.choice([60.0, 60.0, 149.0, 150.0, 151.0, 400.0])
            pw = dict(base)
            pw.pop("modelName") if rng.random() < 0.2 else None
            emit("DCGM_FI_DEV_POWER_USAGE", pw, [watts if i % 17 == 0 else 55.0 for i in range(len(ts))])
    # series the selector must ignore: empty pod label; and one that cannot become PodMetricData
    db.append(Q.series("DCGM_FI_DEV_GPU_UTIL", {"Hostname": "node-0", "gpu": "7", "modelName": "x", pl: "",
                                                 nl: "ml-team", cl: "main"}, [(t_eval, 0.0)]))
    return db, t_eval, duration_min


def wire(db, name, matchers, t_eval, range_s):
    """what a range query for `name{matchers}[range]` returns: the matrix wire format"""
    res = []
    for s in Q.select(db, name, matchers):
        vals = [[t, repr(v)] for (t, v) in s.samples if t_eval - range_s < t <= t_eval]
        if vals:
            res.append({"metric": dict(s.labels), "values": vals})
    return {"status": "success", "data": {"resultType": "matrix", "result": res}}


@pytest.mark.parametrize("seed", range(40))
def test_dense_path_equals_promql_semantics(seed, oracle_np, oracle_c):
    rng = random.Random(seed)
    honor = bool(seed % 2)
    ns_filter = rng.choice([None, None, "ml-.*"])
    model_filter = rng.choice([None, None, "NVIDIA B200"])
    thr = rng.choice([None, 0.0, 150.0])
    db, t_eval, dur = make_cluster(rng, honor)
    # (A) Prometheus-style evaluation of the template + the Rust dedup
    vec = Q.evaluate_template(db, t_eval, dur, ns_filter, model_filter, thr, honor)
    n_series, pods_a = Q.unique_pods(vec, honor)
    # (B) selectors -> wire format -> C++ ingest -> dense tensor -> oracle
    pl, nl = ("pod", "namespace") if honor else ("exported_pod", "exported_namespace")
    m_compute = [(pl, "!=", "")] + ([(nl, "=~", ns_filter)] if ns_filter else [])
    m_power = list(m_compute)
    if model_filter:
        m_compute.append(("modelName", "=~", model_filter))
    rng_s = dur * 60
    util = wire(db, "DCGM_FI_DEV_GPU_UTIL", m_compute, t_eval, rng_s)
    prof = wire(db, "DCGM_FI_PROF_GR_ENGINE_ACTIVE", m_compute, t_eval, rng_s)
    power = wire(db, "DCGM_FI_DEV_POWER_USAGE", m_power, t_eval, rng_s) if thr else None
    if not util["data"]["result"] and not prof["data"]["result"]:
        assert pods_a == []
        return
    H.ingest_mode(-1 if seed % 3 == 0 else seed % 3)      # DOM reference path / threaded text path
    try:
        u, w, meta = H.ingest(util, prof, power, duration_min=dur, step=1, t_end=t_eval)
    finally:
        H.ingest_mode(-1)
    names = [(p["name"], p["namespace"]) for p in meta["pods"]]
    veto_bits = oracle_np.decide(u, w, power_threshold=thr)["veto_bits"]
    for orc in (oracle_np, oracle_c):
        r = orc.decide(u, w, power_threshold=thr)      # every tensor ROW an element: what the kernels compute
        # duplicate series of a `sum by` group: element = sum of the members' maxima (host-side, ingest.cpp)
        cb, db, counts, _ = H.resolve_groups(r["series_max"], r["candidate_bits"], r["decision_bits"],
                                             (r["n_series"], r["n_candidates"], r["n_decisions"]), veto_bits=veto_bits)
        cand = oracle_np.unpack_bits(cb, len(names))
        pods_b = {names[i] for i in np.flatnonzero(cand)}
        assert pods_b == set(pods_a), (seed, sorted(pods_b ^ set(pods_a)))
        assert counts[0] == n_series
        # the value PodMetricData would carry (lib.rs:184): 0 for every element that survived `== 0`, and the
        # same elements as the PromQL-style evaluation (one per idle group of a surviving pod)
        vals = H.group_values(r["series_max"])
        idle_groups = sum(int((vals[i] == 0.0).sum()) for i in np.flatnonzero(cand))
        assert idle_groups == n_series and all(v == 0.0 for v in vec.values())


def test_duplicate_dmi_series_is_a_query_error():
    """two node_dmi_info series for one Hostname make Prometheus fail the whole query (many-to-many
    matching) — the error path noted in SURVEY.md §8(a6), not a value path"""
    rng = random.Random(1)
    db, t_eval, dur = make_cluster(rng)
    db.append(Q.series("node_dmi_info", {"instance": "node-0", "product_name": "other"}, [(t_eval - 3, 1.0)]))
    with pytest.raises(ValueError, match="many-to-many"):
        Q.evaluate_template(db, t_eval, dur)
