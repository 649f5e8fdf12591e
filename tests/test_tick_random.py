"""CPU: the whole tick after the decision (Controller::run_query_and_scale) against a model of the reference's
main.rs:416-570 + lib.rs:437-513 on random clusters: candidate pods -> GET pod (vanished / Pending / no creation
timestamp / younger than window + grace are skipped) -> owner walk -> roots deduplicated by (kind, uid) -> the
requests of scale-down mode filtered by --enabled-resources.  Verdicts are recorded (computed by the oracle and
injected through the test-only C API, as in test_controller_cpu.py); the scenario tests there pin the named cases."""
import json
import random

import numpy as np

import hostlib as H
from hostworld import NOW, NS, _obj, _pod, _series, _write

KIND_LETTER = {"Deployment": "d", "ReplicaSet": "r", "StatefulSet": "s", "InferenceService": "i", "Notebook": "n"}
PLURAL = {"Deployment": "deployments", "ReplicaSet": "replicasets", "StatefulSet": "statefulsets",
          "Notebook": "notebooks", "InferenceService": "inferenceservices"}


def _root_of(objects, pod_meta):
    """lib.rs:437-513 (see tests/test_owner_walk_random.py)"""
    ks = (pod_meta.get("labels") or {}).get("serving.kserve.io/inferenceservice")
    if ks is not None:
        return ("InferenceService", ks) if ("InferenceService", ks) in objects else None
    for ref in pod_meta.get("ownerReferences") or []:
        for owner_kind, parent_kind in (("ReplicaSet", "Deployment"), ("StatefulSet", "Notebook")):
            if ref["kind"] != owner_kind:
                continue
            obj = objects.get((owner_kind, ref["name"]))
            if obj is None:
                break
            for up in obj["metadata"].get("ownerReferences") or []:
                if up["kind"] == parent_kind:
                    return (parent_kind, up["name"]) if (parent_kind, up["name"]) in objects else None
            return (owner_kind, ref["name"])
    return None


def test_random_clusters_through_the_tick(tmp_path, oracle_np):
    rng = random.Random(20260921)
    T, grace = 120, 300
    for round_ in range(10):
        root = tmp_path / f"w{round_}"
        prom, kube = root / "prom", root / "kube"
        prom.mkdir(parents=True)
        names = [f"o{i}" for i in range(6)]
        objects = {}
        for kind in PLURAL:
            for name in rng.sample(names, rng.randrange(2, 6)):
                owners = None
                if kind in ("ReplicaSet", "StatefulSet") and rng.random() < 0.7:
                    owners = [(rng.choice(["Deployment", "Notebook", "Job"]), rng.choice(names)) for _ in range(rng.randrange(1, 3))]
                obj = _obj(name, f"uid-{kind}-{name}", owners)
                objects[(kind, name)] = obj
                _write(kube, PLURAL[kind], NS, obj)
        util, pods = [], {}
        for i in range(rng.randrange(8, 25)):
            name = f"pod-{i}"
            idle = rng.random() < 0.7
            gpus = rng.randrange(1, 4)
            for g in range(gpus):
                vals = [0] * T if (idle and (g == 0 or rng.random() < 0.5)) else [0, 40] * (T // 2)
                util.append(_series(name, g, vals, NOW))
            fate = rng.choice(["ok", "ok", "ok", "ok", "gone", "pending", "young", "no-ts", "edge"])
            labels = {"serving.kserve.io/inferenceservice": rng.choice(names)} if rng.random() < 0.15 else None
            owners = [(rng.choice(["ReplicaSet", "StatefulSet", "Job"]), rng.choice(names)) for _ in range(rng.randrange(0, 3))] or None
            pods[name] = (fate, labels, owners)
            if fate == "gone":
                continue
            age = {"young": 60, "edge": T + grace}.get(fate, 7200)       # edge: created == cutoff -> skipped (`>=`)
            p = _pod(name, owners, labels, age_s=age, phase="Pending" if fate == "pending" else "Running")
            if fate == "no-ts":
                del p["metadata"]["creationTimestamp"]
            _write(kube, "pods", NS, p)
        rng.shuffle(util)
        body = {"status": "success", "data": {"resultType": "matrix", "result": util}}
        (prom / "util.json").write_text(json.dumps(body, separators=(",", ":")))
        (prom / "query.json").write_text(json.dumps({"end": NOW, "step": 1}))
        u, _, meta = H.ingest(body, None, None, duration_min=2, step=1, t_end=NOW)
        r = oracle_np.decide(u)
        enabled = "".join(rng.sample("drsin", rng.randrange(1, 6)))
        argv = ["--prometheus-url", f"file://{prom}", "--kube-fixture", str(kube), "-t", "2", "-g", str(grace), "--now", str(NOW),
                "-l", "json", "-r", "scale-down", "-e", enabled]
        out = H.run_tick(argv, r["candidate_bits"], r["series_max"], r["n_series"], log_path=str(root / "log.jsonl"))
        assert out["ok"], out
        # ---- the model -------------------------------------------------------------------------------------------
        order = [p["name"] for p in meta["pods"]]
        cand = [n for n, c in zip(order, r["candidate"]) if c]
        assert [p["name"] for p in out["unique_pods"]] == cand
        assert out["num_pods"] == r["n_series"]
        roots = []
        for n in cand:
            fate, labels, owners = pods[n]
            if fate in ("gone", "pending", "no-ts", "young", "edge"):
                continue
            pod_meta = {"labels": labels, "ownerReferences": [{"kind": k, "name": o} for k, o in owners or []]}
            root_ = _root_of(objects, pod_meta)
            if root_ is not None and root_ not in roots:
                roots.append(root_)
        assert sorted((x["kind"], x["name"]) for x in out["roots"]) == sorted(roots), (round_, out["roots"], roots)
        assert out["shutdown_events"] == len(roots)
        want_patched = sorted(f"{PLURAL[k]}/{n}" for k, n in roots if KIND_LETTER[k] in enabled)
        got_patched = sorted("/".join(q["path"].replace("/scale", "").split("/")[-2:]) for q in out["requests"] if q["method"] == "PATCH")
        assert got_patched == want_patched, (round_, enabled, got_patched, want_patched)
        assert sum(q["method"] == "POST" for q in out["requests"]) == len(want_patched)
