"""CPU: the controller tick (Controller::run_query_and_scale, mirror of
/root/reference/gpu-pruner/src/main.rs:390-570) with RECORDED verdicts.

The product decides on the GPU (gpr_engine.cpp -> libgpr.so; covered by tests/test_gpu_host_e2e.py).
Here the same window is ingested by the C++ host, the verdict arrays are computed by the CPU oracle
(test infrastructure) and injected through the test-only C API, and everything after the decision —
fused Pending/age gate inputs, skip reasons, owner walk, dedup by UID, dry-run vs scale-down,
--enabled-resources filter, request bodies, failure path — runs exactly as in the binary."""
import json

import numpy as np
import pytest

import hostlib as H
from hostworld import NOW, NS, build_world


def _tick(tmp_path, oracle_np, *extra, thr=None, fail=False):
    tmp, prom, kube = build_world(tmp_path)
    util = json.load(open(prom / "util.json"))
    power = json.load(open(prom / "power.json")) if thr else None
    u, w, meta = H.ingest(util, None, power, duration_min=2, step=1, t_end=NOW)
    r = oracle_np.decide(u, w, power_threshold=thr)
    argv = ["--prometheus-url", f"file://{prom}", "--kube-fixture", str(kube), "-t", "2", "-g", "300",
            "--now", str(NOW), "-l", "json", *extra]
    if thr:
        argv += ["--power-threshold", str(thr)]
    log = str(tmp_path / "log.jsonl")
    out = H.run_tick(argv, r["candidate_bits"], r["series_max"], r["n_series"], fail=fail, log_path=log)
    msgs = [json.loads(l)["fields"]["message"] for l in open(log) if l.startswith("{")]
    return out, msgs, [p["name"] for p in meta["pods"]]


def test_dry_run_tick(tmp_path, oracle_np):
    out, msgs, pods = _tick(tmp_path, oracle_np)
    assert out["ok"] and out["requests"] == []
    assert out["num_pods"] == 11 and out["shutdown_events"] == 4          # series pre-dedup / distinct roots
    assert sorted((r["kind"], r["name"]) for r in out["roots"]) == [
        ("Deployment", "web"), ("InferenceService", "llm"), ("Notebook", "my-nb"), ("StatefulSet", "db")]
    assert [p["name"] for p in out["unique_pods"]] == [p for p in pods if p != "busy-dep-0"]
    assert all(p["node_type"] == "unknown" and p["gpu_model"] == "NVIDIA B200" and p["value"] == 0
               for p in out["unique_pods"])
    # what the engine was handed for the fused gate (main.rs:473-510)
    elig = dict(zip(pods, out["eligible"]))
    assert elig["pending-0"] == 0 and elig["gone-0"] == 0 and elig["idle-ss-0"] == 1 and elig["young-0"] == 1
    assert out["cutoff"] == (NOW - (2 * 60 + 300)) * 10**9 and out["power_on"] is False
    assert "Query returned 11 series across 10 unique pods" in msgs
    for needle in ("Skipping team-a:young-0, created after the lookback window",
                   "Skipping team-a:pending-0, it's still pending",
                   "Skipping team-a:gone-0, pod no longer exists",
                   "Skipping team-a:orphan-0, no scalable root object"):
        assert any(needle in m for m in msgs), needle
    assert sum(m.startswith("Dry-run: Would have sent") for m in msgs) == 4


def test_scale_down_tick_requests(tmp_path, oracle_np):
    out, msgs, _ = _tick(tmp_path, oracle_np, "-r", "scale-down")
    patches = {r["path"]: r["body"] for r in out["requests"] if r["method"] == "PATCH"}
    assert set(patches) == {
        "/apis/apps/v1/namespaces/team-a/deployments/web/scale",
        "/apis/apps/v1/namespaces/team-a/statefulsets/db/scale",
        "/apis/kubeflow.org/v1/namespaces/team-a/notebooks/my-nb",
        "/apis/serving.kserve.io/v1beta1/namespaces/team-a/inferenceservices/llm"}
    assert patches["/apis/apps/v1/namespaces/team-a/deployments/web/scale"] == {"spec": {"replicas": 0}}
    assert patches["/apis/serving.kserve.io/v1beta1/namespaces/team-a/inferenceservices/llm"] == {
        "spec": {"predictor": {"minReplicas": 0}}}
    events = [r for r in out["requests"] if r["method"] == "POST"]
    assert len(events) == 4
    assert {e["body"]["metadata"]["name"] for e in events} == {"gpuscaler-00000000000040008000000000000000"}
    # every PATCH is preceded by the Event for the same object (lib.rs:340-349)
    for i, r in enumerate(out["requests"]):
        if r["method"] == "PATCH":
            ev = out["requests"][i - 1]
            assert ev["method"] == "POST" and ev["body"]["involvedObject"]["name"] in r["path"]


def test_enabled_resources_and_power_veto(tmp_path, oracle_np):
    out, msgs, _ = _tick(tmp_path, oracle_np, "-r", "scale-down", "-e", "dn", thr=150.0)
    assert out["power_on"] is True and out["power_threshold"] == 150.0
    assert out["num_pods"] == 10 and len(out["unique_pods"]) == 9          # hot-0 vetoed pod-wide
    kinds = sorted(r["body"]["involvedObject"]["kind"] for r in out["requests"] if r["method"] == "POST")
    assert kinds == ["Deployment", "Notebook"]
    assert any('Skipping resource type "StatefulSet" because it is not enabled' in m for m in msgs)
    assert out["shutdown_events"] == 4      # the filter acts on the consumer side (main.rs:337-345)


def test_engine_failure_is_a_failed_query(tmp_path, oracle_np):
    out, msgs, _ = _tick(tmp_path, oracle_np, fail=True)
    assert not out["ok"] and out["error"].startswith("Failed to run query!") and out["requests"] == []


def test_missing_window_is_a_failed_query(tmp_path):
    out = H.run_tick(["--prometheus-url", f"file://{tmp_path}/nope", "-t", "2"], np.zeros(1, np.uint32),
                     np.zeros(1, np.float32), 0)
    assert not out["ok"] and "Failed to run query!" in out["error"]
    out = H.run_tick(["--prometheus-url", "http://thanos-querier:9091"], np.zeros(1, np.uint32),
                     np.zeros(1, np.float32), 0)
    assert not out["ok"] and "HTTP transport" in out["error"]
