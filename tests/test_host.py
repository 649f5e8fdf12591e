"""CPU: the C++ host side (gpu-pruner_b200/host) against the reference's own tests as specification.

* CLI: flags / shorts / defaults of /root/reference/gpu-pruner/src/main.rs:46-134
* K15: the 11 template-text assertions of main.rs:584-739, plus byte-exact equality with the
  reference template rendered by jinja2 (tests/golden/query_render.json)
* the 36 unit tests of /root/reference/gpu-pruner/src/lib.rs:578-998 (enabled resources, bitflags,
  ScaleKind Eq/Hash/Meta, Event fields) re-stated against the C++ port
* the owner-walk scenarios of /root/reference/gpu-pruner/tests/e2e.rs:168-252 on JSON fixtures
* ingest of the Prometheus matrix wire format (querytest.rs:41-53)
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

import hostlib as H

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "query_render.json")
URL = ["--prometheus-url", "http://prom:9090"]


# ---------------------------------------------------------------------------------------------
# CLI (main.rs:46-134)
# ---------------------------------------------------------------------------------------------
def test_cli_defaults():
    r = H.parse_cli(URL)
    assert r["ok"]
    c = r["cli"]
    assert c == {"duration": 30, "daemon_mode": False, "enabled_resources": "drsin", "check_interval": 180,
                 "namespace": None, "grace_period": 300, "model_name": None, "power_threshold": None,
                 "honor_labels": False, "run_mode": "dry-run", "prometheus_url": "http://prom:9090",
                 "prometheus_token": None, "prometheus_tls_mode": "verify", "prometheus_tls_cert": None,
                 "log_format": "default"}


def test_cli_short_and_long_forms():
    r = H.parse_cli(["-t", "45", "-d", "-e", "dn", "-c", "60", "-n", "ml-.*", "-g", "120", "-m", "NVIDIA A100",
                     "--power-threshold", "150", "--honor-labels", "-r", "scale-down", "-l", "json",
                     "--prometheus-tls-mode", "skip", "--prometheus-tls-cert", "/x.crt",
                     "--prometheus-token", "tok"] + URL)
    assert r["ok"], r["message"]
    c = r["cli"]
    assert (c["duration"], c["daemon_mode"], c["enabled_resources"], c["check_interval"]) == (45, True, "dn", 60)
    assert (c["namespace"], c["grace_period"], c["model_name"]) == ("ml-.*", 120, "NVIDIA A100")
    assert c["power_threshold"] == 150.0 and c["honor_labels"] is True
    assert (c["run_mode"], c["log_format"], c["prometheus_tls_mode"]) == ("scale-down", "json", "skip")
    assert c["prometheus_tls_cert"] == "/x.crt" and c["prometheus_token"] == "tok"
    r2 = H.parse_cli(["--duration=45", "-c60", "--run-mode=scale-down", "--namespace=a b"] + URL)
    assert r2["ok"] and r2["cli"]["duration"] == 45 and r2["cli"]["check_interval"] == 60
    assert r2["cli"]["run_mode"] == "scale-down" and r2["cli"]["namespace"] == "a b"


@pytest.mark.parametrize("argv,needle", [
    ([], "--prometheus-url"),
    (URL + ["--run-mode", "nuke"], "possible values: scale-down, dry-run"),
    (URL + ["--log-format", "xml"], "possible values: json, default, pretty"),
    (URL + ["--prometheus-tls-mode", "maybe"], "possible values: skip, verify"),
    (URL + ["-t", "abc"], "invalid value 'abc' for '--duration'"),
    (URL + ["-c", "-5"], "invalid value '-5' for '--check-interval'"),
    (URL + ["--bogus"], "unexpected argument '--bogus'"),
    (URL + ["-z"], "unexpected argument '-z'"),
    (URL + ["--duration"], "a value is required for '--duration'"),
    (URL + ["--daemon-mode=yes"], "unexpected value 'yes'"),
    (URL + ["stray"], "unexpected argument 'stray'"),
])
def test_cli_errors_exit_2(argv, needle):
    r = H.parse_cli(argv)
    assert not r["ok"] and r["exit_code"] == 2
    assert needle in r["message"] and r["message"].startswith("error: ")


def test_cli_help_exits_0_and_lists_every_flag():
    r = H.parse_cli(["--help"])
    assert not r["ok"] and r["exit_code"] == 0
    for flag in ("-t, --duration", "-d, --daemon-mode", "-e, --enabled-resources", "-c, --check-interval",
                 "-n, --namespace", "-g, --grace-period", "-m, --model-name", "--power-threshold",
                 "--honor-labels", "-r, --run-mode", "--prometheus-url", "--prometheus-token",
                 "--prometheus-tls-mode", "--prometheus-tls-cert", "-l, --log-format"):
        assert flag in r["message"], flag


def test_binary_cli_surface():
    if not os.path.exists(H.BIN):
        pytest.skip("gpu-pruner binary not built")
    p = subprocess.run([H.BIN, "--help"], capture_output=True, text=True)
    assert p.returncode == 0 and "Usage: gpu-pruner [OPTIONS] --prometheus-url <PROMETHEUS_URL>" in p.stdout
    p = subprocess.run([H.BIN], capture_output=True, text=True)
    assert p.returncode == 2 and "--prometheus-url" in p.stderr
    p = subprocess.run([H.BIN, "--print-query", "-t", "45"], capture_output=True, text=True)
    assert p.returncode == 0 and "[45m]" in p.stdout


# ---------------------------------------------------------------------------------------------
# K15: template-text parity
# ---------------------------------------------------------------------------------------------
def _argv(a):
    v = ["--print-query", "-t", str(a["duration"])]
    if a.get("namespace"):
        v += ["-n", a["namespace"]]
    if a.get("model_name"):
        v += ["-m", a["model_name"]]
    if a.get("power_threshold") is not None:
        v += ["--power-threshold", repr(a["power_threshold"])]
    if a.get("honor_labels"):
        v += ["--honor-labels"]
    return v


def test_render_is_byte_exact_with_the_reference_template():
    cases = json.load(open(GOLD))
    assert len(cases) == 140   # 120: the full grid of ordinary arguments; 20: odd strings, floats, look-backs
    for c in cases:
        assert H.render_query(_argv(c["args"])) == c["text"], c["args"]


def render(**a):
    a.setdefault("duration", 30)
    return H.render_query(_argv(a))


def test_query_uses_max_over_time():                      # main.rs:584-595
    q = render()
    assert "max_over_time(" in q and "avg_over_time(" not in q


def test_query_includes_gpu_util_fallback():              # main.rs:597-612
    q = render()
    assert "DCGM_FI_PROF_GR_ENGINE_ACTIVE" in q and "DCGM_FI_DEV_GPU_UTIL" in q and "/ 100" in q


def test_query_without_power_threshold_has_no_unless():   # main.rs:614-625
    q = render()
    assert "unless" not in q and "DCGM_FI_DEV_POWER_USAGE" not in q


def test_query_with_power_threshold_adds_unless():        # main.rs:627-642
    q = render(power_threshold=150.0)
    assert "unless on (exported_pod, exported_namespace)" in q
    assert "DCGM_FI_DEV_POWER_USAGE" in q and ">= 150" in q


def test_query_with_namespace_filter():                   # main.rs:644-654
    assert render(duration=15, namespace="ml-team").count('exported_namespace =~ "ml-team"') == 4


def test_query_with_namespace_and_power_threshold():      # main.rs:656-668
    q = render(duration=15, namespace="ml-team", power_threshold=100.0)
    assert q.count('exported_namespace =~ "ml-team"') == 5


def test_query_with_model_name_filter():                  # main.rs:670-679
    assert render(model_name="NVIDIA A100").count('modelName =~ "NVIDIA A100"') == 4


def test_query_duration_is_interpolated():                # main.rs:681-688
    assert "[45m]" in render(duration=45)


def test_query_default_uses_exported_labels():            # main.rs:690-705
    q = render()
    assert "exported_pod" in q and "exported_namespace" in q and "exported_container" in q


def test_query_honor_labels_uses_native_labels():         # main.rs:707-726
    q = render(honor_labels=True)
    assert "exported_pod" not in q and "exported_namespace" not in q
    assert "pod !=" in q and "sum by (Hostname, container, pod, namespace" in q


def test_query_honor_labels_with_power_threshold():       # main.rs:728-739
    assert "unless on (pod, namespace)" in render(honor_labels=True, power_threshold=120.0)


def test_zero_power_threshold_is_falsy():                 # query.promql.j2:36
    assert "unless" not in render(power_threshold=0.0)


def test_selector_forms_share_the_filters():
    s = H.render_selectors(["-t", "15", "-n", "ml-team", "-m", "A100", "--power-threshold", "150"] + URL)
    assert s["util"] == 'DCGM_FI_DEV_GPU_UTIL{exported_pod != "", exported_namespace =~ "ml-team", modelName =~ "A100"}[15m]'
    assert s["prof"].startswith("DCGM_FI_PROF_GR_ENGINE_ACTIVE{") and s["prof"].endswith("}[15m]")
    assert s["power"] == 'DCGM_FI_DEV_POWER_USAGE{exported_pod != "", exported_namespace =~ "ml-team"}[15m]'
    assert H.render_selectors(URL)["power"] == ""


@pytest.mark.parametrize("v,s", [(150.0, "150.0"), (120.5, "120.5"), (100.0, "100.0"), (0.1, "0.1"),
                                 (1e-7, "1e-07"), (1e21, "1e+21"), (-3.25, "-3.25")])
def test_float_formatting_matches_jinja(v, s):
    assert H.format_float(v) == s


# ---------------------------------------------------------------------------------------------
# lib.rs unit tests, re-stated (lib.rs:656-997)
# ---------------------------------------------------------------------------------------------
D, R, S, I, N = 1, 2, 4, 8, 16


def test_enabled_resources():
    assert H.enabled_resources("drsin") == D | R | S | I | N          # all flags
    assert H.enabled_resources("n") == N                               # single flag
    assert H.enabled_resources("di") == D | I                          # subset
    assert H.enabled_resources("") == 0                                # empty
    assert H.enabled_resources("xdqz") == D                            # unknown chars ignored
    assert H.enabled_resources("dddd") == H.enabled_resources("d")     # idempotent


def mk(name, ns, uid=None, **extra):
    meta = {"name": name}
    if ns is not None:
        meta["namespace"] = ns
    if uid is not None:
        meta["uid"] = uid
    o = {"metadata": meta}
    o.update(extra)
    return o


KINDS = {"Deployment": D, "ReplicaSet": R, "StatefulSet": S, "InferenceService": I, "Notebook": N}


def test_scale_kind_equality_and_hash():
    eq = lambda ka, a, kb, b: H.scalekind_eq(ka, a, kb, b)
    assert eq("Deployment", mk("d", "ns", "uid-1"), "Deployment", mk("d", "ns", "uid-1"))[0]      # same
    assert not eq("Deployment", mk("d", "ns", "uid-1"), "Deployment", mk("d", "ns", "uid-2"))[0]  # uid differs
    assert not eq("Deployment", mk("x", "ns", "uid-1"), "ReplicaSet", mk("x", "ns", "uid-1"))[0]  # variants
    assert eq("Notebook", mk("nb-a", "ns", "same-uid"), "Notebook", mk("nb-b", "ns", "same-uid"))[0]
    assert eq("InferenceService", mk("is-a", "ns", "uid-x"), "InferenceService", mk("is-b", "ns", "uid-x"))[0]
    # hash = variant + uid: equal objects hash equal; same uid in another variant hashes differently
    _, h1, h2 = eq("Deployment", mk("d", "ns", "uid-1"), "Deployment", mk("d", "ns", "uid-1"))
    assert h1 == h2
    _, h1, h2 = eq("Deployment", mk("x", "ns", "uid-1"), "ReplicaSet", mk("x", "ns", "uid-1"))
    assert h1 != h2
    _, h1, h2 = eq("Notebook", mk("nb-a", "ns", "uid-nb"), "Notebook", mk("nb-b", "ns", "uid-nb"))
    assert h1 == h2


@pytest.mark.parametrize("kind,api", [("Deployment", "apps/v1"), ("ReplicaSet", "apps/v1"),
                                      ("StatefulSet", "apps/v1"), ("Notebook", "v1"),
                                      ("InferenceService", "v1beta1")])
def test_meta_and_event_fields(kind, api):
    ev = H.generate_event(kind, mk("my-obj", "prod", "the-uid", **{"metadata": {
        "name": "my-obj", "namespace": "prod", "uid": "the-uid", "resourceVersion": "42"}}))
    io = ev["involvedObject"]
    assert (io["name"], io["namespace"], io["kind"], io["uid"], io["apiVersion"]) == (
        "my-obj", "prod", kind, "the-uid", api)
    assert io["resourceVersion"] == "42"
    assert ev["action"] == "scale_down" and ev["type"] == "Normal"
    assert ev["reason"] == "Pod prod::my-obj was not using GPU"
    assert ev["reportingComponent"] == "gpu-pruner" and ev["reportingInstance"] == "gpu_pruner"
    assert ev["metadata"]["name"].startswith("gpuscaler-") and len(ev["metadata"]["name"]) == 10 + 32
    assert ev["metadata"]["namespace"] == "prod"
    assert ev["firstTimestamp"] and ev["lastTimestamp"] and ev["eventTime"]


def test_event_names_are_unique_and_pod_name_env():
    a = H.generate_event("Notebook", mk("nb", "ns"))
    b = H.generate_event("Notebook", mk("nb", "ns"), pod_name="gpu-pruner-7d9f")
    assert a["metadata"]["name"] != b["metadata"]["name"]
    assert b["reportingInstance"] == "gpu-pruner-7d9f"
    assert "uid" not in a["involvedObject"]                       # event_for_replica_set: uid None


def test_event_with_no_namespace():
    ev = H.generate_event("Deployment", mk("orphan", None))
    assert "namespace" not in ev["involvedObject"] and ev["reason"] == "Pod ::orphan was not using GPU"
    # and scale() emits no Event POST without a namespace (lib.rs:340)
    rq = H.scale_requests("Deployment", mk("orphan", None))
    assert [r["method"] for r in rq] == ["PATCH"]


def test_filter_integration():
    enabled = H.enabled_resources("dn")
    assert enabled & KINDS["Deployment"] and enabled & KINDS["Notebook"] and not enabled & KINDS["StatefulSet"]


# ---------------------------------------------------------------------------------------------
# scale-to-zero request bodies (lib.rs:517-576)
# ---------------------------------------------------------------------------------------------
def test_scale_requests():
    now = 1_700_000_000_123_456_789
    rq = H.scale_requests("Deployment", mk("web", "prod", "u1"), now)
    assert rq[0]["method"] == "POST" and rq[0]["path"] == "/api/v1/namespaces/prod/events"
    assert rq[0]["body"]["metadata"]["name"] == "gpuscaler-0123456789abcdef0123456789abcdef"
    assert rq[0]["body"]["eventTime"] == "2023-11-14T22:13:20.123456Z"
    assert rq[1] == {"method": "PATCH", "path": "/apis/apps/v1/namespaces/prod/deployments/web/scale",
                     "contentType": "application/merge-patch+json", "body": {"spec": {"replicas": 0}}}
    assert H.scale_requests("ReplicaSet", mk("rs", "a"))[1]["path"] == "/apis/apps/v1/namespaces/a/replicasets/rs/scale"
    assert H.scale_requests("StatefulSet", mk("ss", "a"))[1]["path"] == "/apis/apps/v1/namespaces/a/statefulsets/ss/scale"
    nb = H.scale_requests("Notebook", mk("nb", "ml"), now)[1]
    assert nb["path"] == "/apis/kubeflow.org/v1/namespaces/ml/notebooks/nb"
    assert nb["body"] == {"metadata": {"annotations": {"kubeflow-resource-stopped": "2023-11-14T22:13:20.123456789Z"}}}
    isv = H.scale_requests("InferenceService", mk("llm", "serving"), now)[1]
    assert isv["path"] == "/apis/serving.kserve.io/v1beta1/namespaces/serving/inferenceservices/llm"
    assert isv["body"] == {"spec": {"predictor": {"minReplicas": 0}}}


def test_rfc3339_round_trip():
    assert H.rfc3339(1_700_000_000_000_000_000) == "2023-11-14T22:13:20Z"
    assert H.rfc3339(1_700_000_000_500_000_000) == "2023-11-14T22:13:20.5Z"
    for s in ("2023-11-14T22:13:20Z", "2024-02-29T00:00:00Z", "1999-12-31T23:59:59Z"):
        assert H.rfc3339(H.parse_rfc3339(s)) == s
    assert H.parse_rfc3339("2023-11-14T23:13:20+01:00") == 1_700_000_000_000_000_000


# ---------------------------------------------------------------------------------------------
# owner walk on fixtures (lib.rs:437-513; scenarios of tests/e2e.rs:168-252)
# ---------------------------------------------------------------------------------------------
def _write(root, plural, ns, obj):
    d = root / plural / ns
    d.mkdir(parents=True, exist_ok=True)
    (d / (obj["metadata"]["name"] + ".json")).write_text(json.dumps(obj))


def _owner(kind, name):
    return {"apiVersion": "apps/v1", "kind": kind, "name": name, "uid": "o-" + name}


@pytest.fixture()
def cluster(tmp_path):
    ns = "team-a"
    _write(tmp_path, "deployments", ns, mk("web", ns, "dep-uid"))
    _write(tmp_path, "replicasets", ns, {"metadata": {"name": "web-5d9", "namespace": ns, "uid": "rs-uid",
                                                      "ownerReferences": [_owner("Deployment", "web")]}})
    _write(tmp_path, "replicasets", ns, mk("bare-rs", ns, "rs2-uid"))
    _write(tmp_path, "replicasets", ns, {"metadata": {"name": "dangling-rs", "namespace": ns, "uid": "rs3-uid",
                                                      "ownerReferences": [_owner("Deployment", "gone")]}})
    _write(tmp_path, "statefulsets", ns, mk("db", ns, "ss-uid"))
    _write(tmp_path, "statefulsets", ns, {"metadata": {"name": "nb-ss", "namespace": ns, "uid": "ss2-uid",
                                                       "ownerReferences": [{"kind": "Notebook", "name": "my-nb"}]}})
    _write(tmp_path, "notebooks", ns, mk("my-nb", ns, "nb-uid"))
    _write(tmp_path, "inferenceservices", ns, mk("llm", ns, "is-uid"))
    return str(tmp_path), ns


def test_deployment_chain_resolves_to_deployment_not_replicaset(cluster):   # e2e.rs:168-197
    d, ns = cluster
    r = H.find_root(d, {"name": "web-5d9-abc", "namespace": ns, "ownerReferences": [_owner("ReplicaSet", "web-5d9")]})
    assert (r["kind"], r["name"], r["uid"], r["apiVersion"]) == ("Deployment", "web", "dep-uid", "apps/v1")
    assert r["calls"] == 2


def test_bare_statefulset_and_bare_replicaset(cluster):                      # e2e.rs:199-236
    d, ns = cluster
    r = H.find_root(d, {"name": "db-0", "namespace": ns, "ownerReferences": [_owner("StatefulSet", "db")]})
    assert (r["kind"], r["name"]) == ("StatefulSet", "db")
    r = H.find_root(d, {"name": "p", "namespace": ns, "ownerReferences": [_owner("ReplicaSet", "bare-rs")]})
    assert (r["kind"], r["name"]) == ("ReplicaSet", "bare-rs")


def test_orphan_pod_errors(cluster):                                         # e2e.rs:238-252
    d, ns = cluster
    r = H.find_root(d, {"name": "orphan", "namespace": ns})
    assert "no scalable root object found" in r["error"] and "orphan" in r["error"]
    r = H.find_root(d, {"name": "job-pod", "namespace": ns, "ownerReferences": [_owner("Job", "batch")]})
    assert "no scalable root object found" in r["error"]


def test_statefulset_owned_by_notebook(cluster):
    d, ns = cluster
    r = H.find_root(d, {"name": "nb-ss-0", "namespace": ns, "ownerReferences": [_owner("StatefulSet", "nb-ss")]})
    assert (r["kind"], r["name"], r["apiVersion"], r["resource_kind"]) == ("Notebook", "my-nb", "v1", N)


def test_kserve_label_shortcut_wins_over_owner_refs(cluster):
    d, ns = cluster
    r = H.find_root(d, {"name": "llm-predictor-0", "namespace": ns,
                        "labels": {"serving.kserve.io/inferenceservice": "llm"},
                        "ownerReferences": [_owner("ReplicaSet", "web-5d9")]})
    assert (r["kind"], r["name"], r["calls"]) == ("InferenceService", "llm", 1)
    r = H.find_root(d, {"name": "x", "namespace": ns, "labels": {"serving.kserve.io/inferenceservice": "nope"}})
    assert "not found" in r["error"]          # the GET error propagates (`?`, lib.rs:454)


def test_missing_replicaset_is_swallowed_missing_deployment_propagates(cluster):
    d, ns = cluster
    # RS lookup fails -> `if let Ok(rs)` falls through to the next owner ref (lib.rs:465)
    r = H.find_root(d, {"name": "p", "namespace": ns,
                        "ownerReferences": [_owner("ReplicaSet", "nope"), _owner("StatefulSet", "db")]})
    assert (r["kind"], r["name"]) == ("StatefulSet", "db")
    # RS found, its Deployment is gone -> error propagates (lib.rs:472)
    r = H.find_root(d, {"name": "p", "namespace": ns, "ownerReferences": [_owner("ReplicaSet", "dangling-rs")]})
    assert "error" in r and "gone" in r["error"]


# ---------------------------------------------------------------------------------------------
# ingest of the matrix wire format: DOM reference path and the threaded text path
# ---------------------------------------------------------------------------------------------
@pytest.fixture(params=[-1, 1, 4], ids=["dom", "text-1thread", "text-4threads"], autouse=True)
def ingest_path(request):
    H.ingest_mode(request.param)
    yield request.param
    H.ingest_mode(-1)


def series(labels, samples):
    return {"metric": labels, "values": [[t, str(v)] for t, v in samples]}


def resp(*ss):
    return {"status": "success", "data": {"resultType": "matrix", "result": list(ss)}}


BASE = {"Hostname": "node1", "modelName": "NVIDIA A100", "UUID": "GPU-x"}


def lab(pod, gpu, ns="ml", ctr="main", exported=True, **kw):
    d = dict(BASE, gpu=str(gpu), **kw)
    if exported:
        d.update(exported_pod=pod, exported_namespace=ns, exported_container=ctr)
    else:
        d.update(pod=pod, namespace=ns, container=ctr)
    return d


def test_ingest_layout_and_label_precedence():
    t_end = 1_700_000_060
    u = resp(series(lab("a", 0), [(t_end - 50, 0), (t_end - 40, 0), (t_end, 7)]),
             series(lab("a", 1), [(t_end - 10, 0)]),
             series(lab("b", 0, exported=False, node_type="DGX"), [(t_end - 20, "NaN"), (t_end - 10, 3.5)]),
             # exported_* wins over the bare label (lib.rs:158-175)
             series(dict(lab("c", 0), pod="prometheus-exporter", namespace="monitoring"), [(t_end, 0)]),
             # no pod label at all / empty pod label: not part of the selection (j2:11)
             series({"Hostname": "n", "gpu": "0", "modelName": "m"}, [(t_end, 0)]),
             series(lab("", 0), [(t_end, 0)]))
    util, _, meta = H.ingest(u, duration_min=1, step=10, t_end=t_end)
    assert util.shape == (3, 2, 6)
    pods = meta["pods"]
    assert [(p["name"], p["namespace"]) for p in pods] == [("a", "ml"), ("b", "ml"), ("c", "ml")]
    # `sum by (Hostname, container, pod, namespace, gpu, modelName)` (j2:9) drops every other label: a node_type
    # label on the DCGM series itself never reaches PodMetricData; only the node_dmi_info join sets it (j2:23-34)
    assert pods[1]["slots"][0]["node_type"] == "unknown" and pods[0]["slots"][0]["node_type"] == "unknown"
    assert meta["series_skipped"] == 2
    nan = np.nan
    np.testing.assert_array_equal(util[0, 0], np.array([0, 0, nan, nan, nan, 7], np.float32))
    np.testing.assert_array_equal(util[0, 1], np.array([nan, nan, nan, nan, 0, nan], np.float32))
    np.testing.assert_array_equal(util[1, 0], np.array([nan, nan, nan, nan, 3.5, nan], np.float32))
    assert np.isnan(util[1, 1]).all()                        # pod b has one GPU: second slot absent


def test_ingest_window_edges_duplicates_and_specials():
    t_end = 1000
    u = resp(series(lab("a", 0), [(t_end - 60, 9), (t_end - 59, 5), (t_end, "+Inf")]),   # -60 is outside (t-N, t]
             series(dict(lab("a", 0), UUID="GPU-dup"), [(t_end - 59, 8), (t_end - 1, 1e-60)]),  # same group
             series(lab("a", 0, ctr="sidecar"), [(t_end, 0)]))                           # other group
    util, _, meta = H.ingest(u, duration_min=1, step=1, t_end=t_end)
    # every series keeps its own row; the two members of the `sum by` group are tied by `group` (j2:9)
    assert util.shape == (1, 3, 60)
    assert meta["samples_out_of_window"] == 1 and meta["duplicates_merged"] == 1
    slots = meta["pods"][0]["slots"]
    assert [s["group"] for s in slots] == [0, 0, 2] and meta["pods"][0]["has_groups"]
    assert util[0, 0, 0] == 5 and np.isinf(util[0, 0, 59]) and util[0, 1, 0] == 8
    assert util[0, 1, 58] > 0 and meta["tiny_values_clamped"] == 1  # 1e-60 stays non-zero in f32
    assert util[0, 2, 59] == 0


def test_window_edges_every_second_of_the_range_has_a_bucket():
    """with step 30 a sample 1790 s before t_end is inside max_over_time(X[30m]) evaluated at t_end
    and must land in the tensor (it used to fall off the oldest column)"""
    t_end = 1_700_001_800
    for step, dur in ((30, 30), (7, 1), (60, 30)):
        span = dur * 60
        ts = [t_end - span + 1, t_end - span + step // 2 + 1, t_end - 1, t_end]
        u = resp(series(lab("edge", 0), [(t_end - span, 77)] + [(t, 50 if i == 0 else 0) for i, t in enumerate(sorted(set(ts)))]))
        for mode in (-1, 2):
            H.ingest_mode(mode)
            try:
                util, _, meta = H.ingest(u, duration_min=dur, step=step, t_end=t_end)
            finally:
                H.ingest_mode(-1)
            assert util.shape[2] == -(-span // step) and meta["span"] == span
            assert meta["samples_out_of_window"] == 1                 # only the sample AT t_end - N is outside
            assert np.nanmax(util[0, 0]) == 50                        # the oldest in-window second is there


def test_exact_sum_by_of_duplicate_series(oracle_np):
    """`sum by` (j2:9,21): the element of a group is the SUM of its members' maxima.  A +5 / -5 pair sums to
    0 (idle), a 0 / 7 pair does not, and an idle single series beside them is unaffected"""
    t_end = 2000
    u = resp(series(dict(lab("mixed", 0), UUID="a"), [(t_end - 1, 500), (t_end, 100)]),
             series(dict(lab("mixed", 0), UUID="b"), [(t_end, -500)]),
             series(dict(lab("halfbusy", 0), UUID="a"), [(t_end, 0)]),
             series(dict(lab("halfbusy", 0), UUID="b"), [(t_end, 7)]),
             series(dict(lab("allidle", 0), UUID="a"), [(t_end, 0)]),
             series(dict(lab("allidle", 0), UUID="b"), [(t_end - 5, 0)]),
             series(lab("single", 0), [(t_end, 0)]))
    util, _, meta = H.ingest(u, duration_min=1, step=1, t_end=t_end)
    names = [p["name"] for p in meta["pods"]]
    raw = oracle_np.decide(util)                                   # every ROW an element: the engine's raw verdict
    raw_set = {names[i] for i in np.flatnonzero(raw["candidate"])}
    assert raw_set == {"halfbusy", "allidle", "single"}            # halfbusy's idle member fools the raw rule
    counts = (raw["n_series"], raw["n_candidates"], raw["n_decisions"])
    cb, db, counts2, changed = H.resolve_groups(raw["series_max"], raw["candidate_bits"], raw["decision_bits"], counts)
    cand = oracle_np.unpack_bits(cb, len(names))
    assert {names[i] for i in np.flatnonzero(cand)} == {"mixed", "allidle", "single"} and changed == 2
    assert counts2 == (3, 3, 3)                                    # one element per idle group
    vals = H.group_values(raw["series_max"])
    assert vals[names.index("mixed"), 0] == 0.0 and vals[names.index("halfbusy"), 0] == 0.07
    assert np.isnan(vals[names.index("mixed"), 1])                 # the second row starts no group


def test_node_type_from_node_dmi_info():
    """query.promql.j2:23-34: Hostname := instance, node_type := product_name, joined on Hostname; hosts without a
    DMI series read back as "unknown" (lib.rs:176-179); a duplicate Hostname on the DMI side fails the query"""
    t_end = 3000
    u = resp(series(dict(lab("a", 0), Hostname="node-1"), [(t_end, 0)]),
             series(dict(lab("b", 0), Hostname="node-2"), [(t_end, 0)]),
             series(dict(lab("c", 0), Hostname="node-3"), [(t_end, 0)]))
    dmi = {"status": "success", "data": {"resultType": "vector", "result": [
        {"metric": {"__name__": "node_dmi_info", "instance": "node-1", "product_name": "DGX B200"}, "value": [t_end, "1"]},
        {"metric": {"__name__": "node_dmi_info", "instance": "node-2"}, "value": [t_end, "1"]}]}}
    H.ingest_dmi(dmi)
    try:
        _, _, meta = H.ingest(u, duration_min=1, step=1, t_end=t_end)
        assert [p["slots"][0]["node_type"] for p in meta["pods"]] == ["DGX B200", "unknown", "unknown"]
        dmi["data"]["result"].append({"metric": {"instance": "node-1", "product_name": "other"}, "value": [t_end, "1"]})
        H.ingest_dmi(dmi)
        with pytest.raises(RuntimeError, match="duplicate series"):
            H.ingest(u, duration_min=1, step=1, t_end=t_end)
    finally:
        H.ingest_dmi(None)


def test_ingest_prof_shadows_util_and_power_plane():
    t_end = 500
    prof = resp(series(lab("a", 0), [(t_end, 0.25)]))
    u = resp(series(lab("a", 0), [(t_end, 99)]), series(lab("a", 1), [(t_end, 0)]))
    pw = resp(series({k: v for k, v in lab("a", 1).items() if k != "modelName"}, [(t_end - 1, 180), (t_end, 60)]))
    util, power, meta = H.ingest(u, prof, pw, duration_min=1, step=1, t_end=t_end)
    assert util[0, 0, 59] == np.float32(0.25) and meta["pods"][0]["slots"][0]["from_prof"]   # `A or B` (j2:10-20)
    assert util[0, 1, 59] == 0 and not meta["pods"][0]["slots"][1]["from_prof"]
    assert power.shape == util.shape and np.nanmax(power[0]) == 180


def test_ingest_rejects_non_matrix():
    with pytest.raises(RuntimeError, match="expected matrix"):
        H.ingest({"status": "success", "data": {"resultType": "vector", "result": []}})
    with pytest.raises(RuntimeError, match="status"):
        H.ingest({"status": "error", "error": "query timed out"})


def test_ingested_window_feeds_the_oracle(oracle_np):
    """wire format -> tensor -> decision: the verdicts the PromQL would have produced"""
    t_end = 10_000
    ts = list(range(t_end - 59, t_end + 1))
    u = resp(series(lab("idle-pod", 0), [(t, 0) for t in ts]),
             series(lab("busy-pod", 0), [(t, 40 if t % 7 == 0 else 0) for t in ts]),
             series(lab("two-gpu", 0), [(t, 90) for t in ts]),
             series(lab("two-gpu", 1), [(t, 0) for t in ts[30:]]),     # young series, leading gap
             series(lab("hot-idle", 0), [(t, 0) for t in ts]))
    pw = resp(series(lab("hot-idle", 0), [(t, 200) for t in ts]), series(lab("idle-pod", 0), [(t, 55) for t in ts]))
    util, power, meta = H.ingest(u, None, pw, duration_min=1, step=1, t_end=t_end)
    names = [p["name"] for p in meta["pods"]]
    r = oracle_np.decide(util, power, power_threshold=150.0)
    verdict = dict(zip(names, r["candidate"]))
    assert verdict == {"idle-pod": True, "busy-pod": False, "two-gpu": True, "hot-idle": False}


def test_text_ingest_equals_dom_ingest_on_random_responses():
    """the fast path must reproduce the reference path bit for bit: odd spacing, float timestamps,
    exponent / NaN / Inf value strings, empty series, duplicates, out-of-window samples"""
    import random
    for seed in range(25):
        rng = random.Random(seed)
        t_end = 1_700_000_000 + rng.randrange(1000)
        step = rng.choice([1, 5, 15])
        dur = rng.choice([1, 2, 5])
        sers = []
        for p in range(rng.randrange(1, 9)):
            for g in range(rng.randrange(1, 4)):
                n = rng.randrange(0, 40)
                tss = sorted({t_end - step * rng.randrange(0, dur * 60 // step + 10) for _ in range(n)})
                smp = []
                for t in tss:
                    tv = t if rng.random() < 0.8 else t + rng.choice([0.123, 0.4, 0.499])
                    v = rng.choice(["0", "0", "37", "100", "0.25", "1e-3", "2.5E+1", "NaN", "+Inf", "-Inf",
                                    "99.99999999999999", "0.30000000000000004", "1e-60", "12345678901234567890"])
                    smp.append((tv, v))
                lab = lab_(rng, p, g)
                sers.append({"metric": lab, "values": [[t, v] for t, v in smp]})
                if rng.random() < 0.2:
                    sers.append({"metric": dict(lab, UUID="dup"), "values": [[t_end, "3"]]})
        text = {"status": "success", "data": {"resultType": "matrix", "result": sers}}
        pw = {"status": "success", "data": {"result": [s for s in sers if rng.random() < 0.5], "resultType": "matrix"}}
        outs = []
        for mode in (-1, 1, 3):
            H.ingest_mode(mode)
            for kw in (dict(step=step, t_end=t_end), dict()):      # given, or inferred from the data
                outs.append((mode, kw, H.ingest(text, None, pw, duration_min=dur, **kw)))
        ref = {repr(kw): o for m, kw, o in outs if m == -1}
        for m, kw, (u, w, meta) in outs:
            ru, rw, rmeta = ref[repr(kw)]
            assert u.shape == ru.shape and np.array_equal(u.view(np.uint32), ru.view(np.uint32)), (seed, m, kw)
            assert np.array_equal(w.view(np.uint32), rw.view(np.uint32))
            meta.pop("ingest_ms"), rmeta.pop("ingest_ms", None)
            assert meta == rmeta
    H.ingest_mode(-1)


def lab_(rng, p, g):
    d = {"Hostname": f"n{p % 3}", "gpu": str(g), "modelName": "NVIDIA B200", "UUID": f"GPU-{p}-{g}",
         "exported_pod": f"pod-{p}", "exported_namespace": "ml", "exported_container": "main"}
    if rng.random() < 0.1:
        d["weird \"label\\name"] = "va\"lue, with ]] brackets [[ and \\ slashes"
    return d


def test_text_ingest_throughput_smoke():
    """not a benchmark: the text path must be far quicker than the DOM path on a mid-size response"""
    import time
    t_end, T = 1_700_000_000, 600
    sers = [{"metric": {"Hostname": "n", "gpu": str(g), "modelName": "m", "exported_pod": f"p{p}",
                        "exported_namespace": "ns", "exported_container": "c"},
             "values": [[t_end - T + 1 + i, "0" if (i + p) % 7 else "42"] for i in range(T)]}
            for p in range(300) for g in range(2)]
    text = {"status": "success", "data": {"resultType": "matrix", "result": sers}}
    times = {}
    for mode in (-1, 0):
        H.ingest_mode(mode)
        t0 = time.perf_counter()
        u, _, _ = H.ingest(text, duration_min=10, step=1, t_end=t_end)
        times[mode] = time.perf_counter() - t0
        assert u.shape == (300, 2, 600)
    H.ingest_mode(-1)
    assert times[0] < times[-1]


def test_text_ingest_accepts_every_json_layout():
    """compact (what Prometheus emits), spaced, and indented responses give the same tensor through the
    text path as through the DOM parser — the byte-level shortcuts (SWAR search for the list end, in-place
    label maps) must not make the parser pickier than JSON"""
    import ctypes as C
    import random
    rng = random.Random(5)
    t_end = 1_700_000_000
    sers = []
    for p in range(6):
        for g in range(rng.randrange(1, 3)):
            n = rng.choice([0, 1, 7, 60])
            sers.append({"metric": lab_(rng, p, g),
                         "values": [[t_end - n + 1 + i, rng.choice(["0", "5", "NaN", "0.5"])] for i in range(n)]})
    resp = {"status": "success", "data": {"resultType": "matrix", "result": sers}}
    layouts = {"compact": json.dumps(resp, separators=(",", ":")), "spaced": json.dumps(resp),
               "indent1": json.dumps(resp, indent=1), "indent4": json.dumps(resp, indent=4),
               "tabs": json.dumps(resp, indent="\t"), "crlf": json.dumps(resp, indent=2).replace("\n", "\r\n")}

    def run(text, mode):
        H.ingest_mode(mode)
        dims = (C.c_uint * 3)()
        meta = C.create_string_buffer(1 << 20)
        u = np.zeros((6, 2, 60), np.float32)
        rc = H.lib().gph_ingest(text.encode(), None, None, C.c_longlong(1), C.c_longlong(1), C.c_longlong(t_end), dims,
                                u.ctypes.data_as(C.c_void_p), None, meta, 1 << 20)
        assert rc == 0, (mode, meta.value[:200])
        m = json.loads(meta.value.decode())
        m.pop("ingest_ms", None)
        return list(dims), u, m

    ref = run(layouts["compact"], -1)
    for name, text in layouts.items():
        for mode in (-1, 1, 2):
            dims, u, m = run(text, mode)
            assert dims == ref[0] and np.array_equal(u.view(np.uint32), ref[1].view(np.uint32)), (name, mode)
            assert m == ref[2], (name, mode)
    H.ingest_mode(-1)


# ---------------------------------------------------------------------------------------------
# a sample value that is not a number fails the query (the reference decodes it with f64::from_str through
# prometheus-http-query, main.rs:405-409) — it must never be read as 0.0, which is an idle GPU
# ---------------------------------------------------------------------------------------------
def _one_series(value):
    lab = {"__name__": "DCGM_FI_DEV_GPU_UTIL", "Hostname": "n", "gpu": "0", "modelName": "m", "exported_pod": "p",
           "exported_namespace": "ns", "exported_container": "c"}
    return {"status": "success", "data": {"resultType": "matrix", "result": [
        {"metric": lab, "values": [[1_700_000_000 - 1, "7"], [1_700_000_000, value]]}]}}


@pytest.mark.parametrize("value", ["abc", "", "0x10", "1e", "0.5x", " 1", "1 ", "1,5", "nan(1)", "1.2.3", "--1", "0.5e+", "٣"])
def test_a_sample_value_that_is_not_a_number_fails_the_query(value):
    with pytest.raises(RuntimeError) as ei:
        H.ingest(_one_series(value), duration_min=1, step=1, t_end=1_700_000_000)
    assert "not a number" in str(ei.value)


@pytest.mark.parametrize("value, want", [("0", 0.0), ("-0", -0.0), ("+3", 3.0), ("12.25", 12.25), ("1e2", 100.0), ("1E1", 10.0),
                                          ("5e-07", 5e-07), ("NaN", float("nan")), ("+Inf", float("inf")), ("-Inf", float("-inf")),
                                          ("Inf", float("inf")), ("0.30000000000000004", 0.30000000000000004), (".5", 0.5), ("5.", 5.0)])
def test_every_spelling_prometheus_prints_is_a_number(value, want):
    u, _, _ = H.ingest(_one_series(value), duration_min=1, step=1, t_end=1_700_000_000)
    got = float(u[0, 0, -1])
    assert (got != got and want != want) or got == np.float32(want), (value, got)


def test_garbage_is_garbage_wherever_it_sits():
    """out of the window, or in a series that gets no row: the reference's decode reads every sample of the response"""
    T = 1_700_000_000
    good = _one_series("7")["data"]["result"][0]
    orphan = {"metric": {k: v for k, v in good["metric"].items() if k != "exported_pod"}, "values": [[T, "abc"]]}
    for series in ([{"metric": good["metric"], "values": [[T - 5000, "abc"], [T, "7"]]}], [good, orphan],
                   [good, {"metric": orphan["metric"], "values": [[T, "7"], "x"]}]):
        with pytest.raises(RuntimeError):
            H.ingest({"status": "success", "data": {"resultType": "matrix", "result": series}}, duration_min=1, step=1, t_end=T)
    orphan["values"] = [[T, "NaN"]]
    u, _, meta = H.ingest({"status": "success", "data": {"resultType": "matrix", "result": [good, orphan]}},
                          duration_min=1, step=1, t_end=T)
    assert u.shape[0] == 1


def test_repeated_envelope_field_fails_the_decode():
    """`{"status":…,"data":…,"data":{…}}`: the envelope is a struct in the reference's decoder (a repeated field is
    an error); found by a long run of the parser fuzz, where the DOM path took the last `data` and the text path the
    first.  Repeated LABELS are a map there: the last value wins, in every path."""
    T = 1_700_000_000
    good = _one_series("7")
    text = json.dumps(good, separators=(",", ":")).replace('"data":{', '"data":"success","data":{', 1)
    for mode in (-1, 1, 4):
        H.ingest_mode(mode)
        rc = H.lib().gph_ingest(text.encode(), None, None, C.c_longlong(1), C.c_longlong(1), C.c_longlong(T),
                                (C.c_uint * 3)(), None, None, C.create_string_buffer(1 << 16), 1 << 16)
        assert rc != 0, mode
    dup_label = json.dumps(good, separators=(",", ":")).replace('"gpu":"0"', '"gpu":"9","gpu":"0"', 1)
    for mode in (-1, 1):
        H.ingest_mode(mode)
        dims = (C.c_uint * 3)()
        assert H.lib().gph_ingest(dup_label.encode(), None, None, C.c_longlong(1), C.c_longlong(1), C.c_longlong(T),
                                  dims, None, None, C.create_string_buffer(1 << 16), 1 << 16) == 0
        assert tuple(dims) == (1, 1, 60)
