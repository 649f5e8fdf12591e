"""CPU: the owner walk (gpu-pruner_b200/host/kube.cpp find_root_object) against a model of the reference's
decision tree (/root/reference/gpu-pruner/src/lib.rs:437-513), on random clusters.

The model, straight from the reference:
  1. label serving.kserve.io/inferenceservice present  -> GET that InferenceService; a failed GET is the error (`?`)
  2. for every owner reference, in order:
       ReplicaSet : GET it; failure -> next owner reference (`if let Ok`); found: first owner of kind Deployment ->
                    GET the Deployment (`?`: failure is the error), else the ReplicaSet itself
       StatefulSet: the same with Notebook
       any other kind: ignored
  3. nothing returned -> error "no scalable root object found"
The scenario tests in test_host.py pin the named cases of tests/e2e.rs; this one walks the whole tree."""
import json
import random

import hostlib as H

KINDS = {"Deployment": "deployments", "ReplicaSet": "replicasets", "StatefulSet": "statefulsets",
         "Notebook": "notebooks", "InferenceService": "inferenceservices"}


def _model(objects, pod):
    """objects: {(kind, ns, name): object}; returns (kind, name) or ('error', substring)"""
    ns = pod.get("namespace", "")
    ks = (pod.get("labels") or {}).get("serving.kserve.io/inferenceservice")
    if ks is not None:
        return ("InferenceService", ks) if ("InferenceService", ns, ks) in objects else ("error", ks)
    for ref in pod.get("ownerReferences") or []:
        for owner_kind, parent_kind in (("ReplicaSet", "Deployment"), ("StatefulSet", "Notebook")):
            if ref["kind"] != owner_kind:
                continue
            obj = objects.get((owner_kind, ns, ref["name"]))
            if obj is None:
                break                                   # GET failed: swallowed, next owner reference
            for up in obj["metadata"].get("ownerReferences") or []:
                if up["kind"] == parent_kind:
                    if (parent_kind, ns, up["name"]) in objects:
                        return (parent_kind, up["name"])
                    return ("error", up["name"])       # `?` on the parent's GET
            return (owner_kind, ref["name"])
    return ("error", "no scalable root object found")


def _ref(kind, name):
    return {"apiVersion": "apps/v1", "kind": kind, "name": name, "uid": f"u-{kind}-{name}"}


def test_owner_walk_equals_the_reference_decision_tree(tmp_path):
    rng = random.Random(20260921)
    namespaces = ["team-a", "team-b"]
    for round_ in range(12):
        root = tmp_path / f"c{round_}"
        objects = {}
        names = [f"o{i}" for i in range(8)]
        for ns in namespaces:
            for kind in KINDS:
                for name in rng.sample(names, rng.randrange(2, 6)):      # the others do not exist (GET fails)
                    meta = {"name": name, "namespace": ns, "uid": f"uid-{kind}-{ns}-{name}"}
                    if kind in ("ReplicaSet", "StatefulSet") and rng.random() < 0.7:
                        ups = []
                        for _ in range(rng.randrange(1, 4)):
                            ups.append(_ref(rng.choice(["Deployment", "Notebook", "Job", "Rollout"]), rng.choice(names)))
                        meta["ownerReferences"] = ups
                    obj = {"metadata": meta}
                    objects[(kind, ns, name)] = obj
                    d = root / KINDS[kind] / ns
                    d.mkdir(parents=True, exist_ok=True)
                    (d / f"{name}.json").write_text(json.dumps(obj))
        for i in range(120):
            ns = rng.choice(namespaces)
            pod = {"name": f"pod-{i}", "namespace": ns}
            if rng.random() < 0.15:
                pod["labels"] = {"app": "x", "serving.kserve.io/inferenceservice": rng.choice(names)}
            elif rng.random() < 0.3:
                pod["labels"] = {"app": "x"}
            if rng.random() < 0.85:
                pod["ownerReferences"] = [_ref(rng.choice(["ReplicaSet", "StatefulSet", "Job", "DaemonSet"]), rng.choice(names))
                                          for _ in range(rng.randrange(1, 4))]
            want = _model(objects, pod)
            got = H.find_root(str(root), pod)
            if want[0] == "error":
                assert "error" in got and want[1] in got["error"], (pod, want, got)
            else:
                assert "error" not in got, (pod, want, got)
                assert (got["kind"], got["name"]) == want, (pod, want, got)
                assert got["uid"] == f"uid-{want[0]}-{ns}-{want[1]}"
