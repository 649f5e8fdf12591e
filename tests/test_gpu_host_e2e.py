"""GPU: the `gpu-pruner` binary end to end — Prometheus matrix fixtures -> ingest -> libgpr decision on
the GPU -> Pending/age gates -> owner walk -> dedup -> scale requests — against the behaviour of
the reference controller (/root/reference/gpu-pruner/src/main.rs:390-570, lib.rs:337-576)."""
import json
import os
import subprocess

import numpy as np
import pytest

import hostlib as H

pytestmark = pytest.mark.gpu
from hostworld import NOW, NS, build_world


@pytest.fixture()
def world(tmp_path):
    return build_world(tmp_path)


def _run(world, *extra):
    tmp, prom, kube = world
    out = tmp / "patches.jsonl"
    if out.exists():
        out.unlink()
    cmd = [H.BIN, "--prometheus-url", f"file://{prom}", "--kube-fixture", str(kube), "-t", "2", "-g", "300",
           "--now", str(NOW), "--patch-out", str(out), "-l", "json", *extra]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    logs = [json.loads(l) for l in p.stderr.splitlines() if l.startswith("{")]
    reqs = [json.loads(l) for l in out.read_text().splitlines()] if out.exists() else []
    return p, logs, reqs


def _msgs(logs):
    return [l["fields"]["message"] for l in logs]


def test_dry_run_reports_the_same_roots_and_sends_nothing(world):
    p, logs, reqs = _run(world)
    assert p.returncode == 0, p.stderr
    msgs = _msgs(logs)
    assert reqs == []
    would = sorted(m for m in msgs if m.startswith("Dry-run: Would have sent"))
    assert would == sorted([
        "Dry-run: Would have sent [Deployment] team-a:web for scaledown",
        "Dry-run: Would have sent [StatefulSet] team-a:db for scaledown",
        "Dry-run: Would have sent [Notebook] team-a:my-nb for scaledown",
        "Dry-run: Would have sent [InferenceService] team-a:llm for scaledown"])
    # 11 idle series (hot-0 is idle too: no power clause without --power-threshold) across 10 unique pods
    assert "Query returned 11 series across 10 unique pods" in msgs
    assert any(m.startswith("Device ingest: ") and "parsed on the GPU" in m for m in msgs)   # the default ingest
    assert any("Skipping team-a:young-0, created after the lookback window" in m for m in msgs)
    assert any("Skipping team-a:pending-0, it's still pending" in m for m in msgs)
    assert any("Skipping team-a:gone-0, pod no longer exists" in m for m in msgs)
    assert any("Skipping team-a:orphan-0, no scalable root object" in m for m in msgs)
    cnt = {k: v for l in logs for k, v in l["fields"].items() if k.startswith(("counter.", "monotonic_counter."))}
    assert cnt["counter.query_returned_candidates"] == "11"
    assert cnt["counter.query_returned_shutdown_events"] == "4"
    assert cnt["monotonic_counter.query_successes"] == "1"


def test_scale_down_emits_event_plus_patch_per_root(world):
    p, logs, reqs = _run(world, "-r", "scale-down")
    assert p.returncode == 0, p.stderr
    patches = {r["path"]: r["body"] for r in reqs if r["method"] == "PATCH"}
    assert patches == {
        "/apis/apps/v1/namespaces/team-a/deployments/web/scale": {"spec": {"replicas": 0}},
        "/apis/apps/v1/namespaces/team-a/statefulsets/db/scale": {"spec": {"replicas": 0}},
        "/apis/kubeflow.org/v1/namespaces/team-a/notebooks/my-nb":
            patches["/apis/kubeflow.org/v1/namespaces/team-a/notebooks/my-nb"],
        "/apis/serving.kserve.io/v1beta1/namespaces/team-a/inferenceservices/llm":
            {"spec": {"predictor": {"minReplicas": 0}}}}
    assert "kubeflow-resource-stopped" in patches[
        "/apis/kubeflow.org/v1/namespaces/team-a/notebooks/my-nb"]["metadata"]["annotations"]
    events = [r for r in reqs if r["method"] == "POST"]
    assert len(events) == 4 and all(e["path"] == "/api/v1/namespaces/team-a/events" for e in events)
    assert sorted(e["body"]["involvedObject"]["kind"] for e in events) == [
        "Deployment", "InferenceService", "Notebook", "StatefulSet"]
    assert any(e["body"]["reason"] == "Pod team-a::web was not using GPU" for e in events)


def test_enabled_resources_filter_and_power_veto(world):
    p, logs, reqs = _run(world, "-r", "scale-down", "-e", "dn", "--power-threshold", "150")
    assert p.returncode == 0, p.stderr
    kinds = sorted(r["body"]["involvedObject"]["kind"] for r in reqs if r["method"] == "POST")
    assert kinds == ["Deployment", "Notebook"]
    msgs = _msgs(logs)
    assert any('Skipping resource type "StatefulSet" because it is not enabled' in m for m in msgs)
    # hot-0 draws 300 W >= 150 W: vetoed pod-wide, so 10 series / 9 pods survive the query
    assert "Query returned 10 series across 9 unique pods" in msgs


def test_query_failure_is_counted_not_fatal(world, tmp_path):
    tmp, prom, kube = world
    p = subprocess.run([H.BIN, "--prometheus-url", "http://thanos-querier:9091", "-l", "json"],
                       capture_output=True, text=True, timeout=60)
    assert p.returncode == 0                      # like the reference: logged + counted (main.rs:310-321)
    logs = [json.loads(l) for l in p.stderr.splitlines() if l.startswith("{")]
    assert any(l["level"] == "ERROR" and "monotonic_counter.query_failures" in l["fields"] for l in logs)
    # daemon mode gives up after the 7th consecutive failure
    p = subprocess.run([H.BIN, "--prometheus-url", f"file://{tmp_path}/nope", "-d", "-c", "0", "-l", "json",
                        "--max-ticks", "20"], capture_output=True, text=True, timeout=60)
    logs = [json.loads(l) for l in p.stderr.splitlines() if l.startswith("{")]
    fails = [l for l in logs if "monotonic_counter.query_failures" in l["fields"]]
    assert len(fails) == 7 and _msgs(logs)[-1] == "Too many consecutive failures, exiting"


def test_engine_verdicts_equal_oracle_on_the_ingested_window(world, oracle_c):
    """the tensor the host builds from the wire format, decided by the GPU == decided by the oracle"""
    tmp, prom, kube = world
    util, power, meta = H.ingest(json.load(open(prom / "util.json")), None, json.load(open(prom / "power.json")),
                                 duration_min=2, step=1, t_end=NOW)
    import gpu_pruner_b200 as g
    with g.IdleEngine(device=0, max_pods=64, max_gpus=4, max_samples=256, power_plane=True) as eng:
        for thr in (0.0, 150.0):
            d = eng.decide(util, power, power_threshold=thr, want_series_max=True)
            exp = oracle_c.decide(util, power, power_threshold=thr)
            assert np.array_equal(d.candidate_bits, exp["candidate_bits"]) and d.n_series == exp["n_series"]
