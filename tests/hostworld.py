"""Shared fixture for the controller tests: a Prometheus range-query result (util + power) and a
Kubernetes object tree covering every branch of the reference's tick (main.rs:416-570, lib.rs:437-513):
multi-GPU pods, dedup of two pods under one Deployment, StatefulSet, Notebook via StatefulSet, KServe
label shortcut, young pod, Pending pod, vanished pod, orphan pod, power-hungry idle pod."""
import json

import pytest

import hostlib as H

NOW = 1_700_000_000          # --now (unix seconds)
NS = "team-a"


def _write(root, plural, ns, obj):
    d = root / plural / ns
    d.mkdir(parents=True, exist_ok=True)
    (d / (obj["metadata"]["name"] + ".json")).write_text(json.dumps(obj))


def _ts(secs):
    return H.rfc3339(secs * 1_000_000_000)


def _pod(name, owners=None, labels=None, age_s=7200, phase="Running"):
    meta = {"name": name, "namespace": NS, "uid": "pod-" + name, "creationTimestamp": _ts(NOW - age_s)}
    if owners:
        meta["ownerReferences"] = [{"kind": k, "name": n, "apiVersion": "apps/v1", "uid": "o"} for k, n in owners]
    if labels:
        meta["labels"] = labels
    return {"metadata": meta, "status": {"phase": phase}}


def _obj(name, uid, owners=None):
    meta = {"name": name, "namespace": NS, "uid": uid, "resourceVersion": "7"}
    if owners:
        meta["ownerReferences"] = [{"kind": k, "name": n} for k, n in owners]
    return {"metadata": meta}


def _series(pod, gpu, vals, t_end):
    lab = {"Hostname": "node-1", "gpu": str(gpu), "modelName": "NVIDIA B200", "UUID": f"GPU-{pod}-{gpu}",
           "exported_pod": pod, "exported_namespace": NS, "exported_container": "main"}
    return {"metric": lab, "values": [[t_end - (len(vals) - 1 - i), str(v)] for i, v in enumerate(vals)]}


def build_world(tmp_path, compact=True):
    """compact=True: Prometheus' own encoding (no whitespace), which the device ingest parses on the GPU"""
    prom, kube = tmp_path / "prom", tmp_path / "kube"
    prom.mkdir()
    T = 120                                     # -t 2 minutes @ 1 s
    idle, busy = [0] * T, [0, 0, 35] * (T // 3)
    util, power = [], []
    pods = {
        "idle-dep-0": ([idle, idle], [("ReplicaSet", "web-rs")], None),
        "idle-dep-1": ([busy, idle], [("ReplicaSet", "web-rs")], None),    # ANY GPU idle is enough
        "busy-dep-0": ([busy, busy], [("ReplicaSet", "api-rs")], None),
        "idle-ss-0": ([idle], [("StatefulSet", "db")], None),
        "nb-0": ([idle], [("StatefulSet", "nb-ss")], None),
        "llm-0": ([idle], [("ReplicaSet", "web-rs")], {"serving.kserve.io/inferenceservice": "llm"}),
        "young-0": ([idle], [("StatefulSet", "db")], None),
        "pending-0": ([idle], [("StatefulSet", "db")], None),
        "gone-0": ([idle], None, None),
        "orphan-0": ([idle], None, None),
        "hot-0": ([idle], [("StatefulSet", "db")], None),
    }
    for name, (gpus, owners, labels) in pods.items():
        for g, vals in enumerate(gpus):
            util.append(_series(name, g, vals, NOW))
            power.append(_series(name, g, [300 if name == "hot-0" else 60] * T, NOW))
        if name == "gone-0":
            continue
        _write(kube, "pods", NS, _pod(name, owners, labels,
                                      age_s=60 if name == "young-0" else 7200,
                                      phase="Pending" if name == "pending-0" else "Running"))
    sep = (",", ":") if compact else (", ", ": ")
    (prom / "util.json").write_text(json.dumps({"status": "success", "data": {"resultType": "matrix", "result": util}},
                                               separators=sep))
    (prom / "power.json").write_text(json.dumps({"status": "success", "data": {"resultType": "matrix", "result": power}},
                                                separators=sep))
    (prom / "query.json").write_text(json.dumps({"end": NOW, "step": 1}))
    _write(kube, "deployments", NS, _obj("web", "dep-web"))
    _write(kube, "deployments", NS, _obj("api", "dep-api"))
    _write(kube, "replicasets", NS, _obj("web-rs", "rs-web", [("Deployment", "web")]))
    _write(kube, "replicasets", NS, _obj("api-rs", "rs-api", [("Deployment", "api")]))
    _write(kube, "statefulsets", NS, _obj("db", "ss-db"))
    _write(kube, "statefulsets", NS, _obj("nb-ss", "ss-nb", [("Notebook", "my-nb")]))
    _write(kube, "notebooks", NS, _obj("my-nb", "nb-1"))
    _write(kube, "inferenceservices", NS, _obj("llm", "is-1"))
    return tmp_path, prom, kube


