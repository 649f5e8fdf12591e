"""CPU: label maps read in place (FlatLabels) vs through the DOM parser.

The ingest paths read Prometheus' own label-map shape — {"k":"v",...}, string values, no escapes, no
whitespace, unique keys — as string_views into the response text and fall back to the DOM parser for
anything else (gpu-pruner_b200/host/ingest_internal.hpp).  Both views must drive the row assignment
(label precedence lib.rs:153-187, `sum by` groups query.promql.j2:9) identically, and the in-place reader
must accept exactly the shape it claims to."""
import json
import os
import random
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "gpu-pruner_b200", "host")


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    out = tmp_path_factory.mktemp("labels") / "labels_check"
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
                           "-I", HOST, os.path.join(ROOT, "tests", "cpp", "labels_check.cpp"),
                           os.path.join(HOST, "json.cpp"), "-o", str(out)])
    return str(out)


def _maps(rng, n):
    keys = ["exported_pod", "pod", "exported_namespace", "namespace", "exported_container", "container",
            "modelName", "Hostname", "gpu", "UUID", "node_type", "__name__", "instance"]
    vals = ["", "a", "pod-1", "pod-2", "ns", "ns2", "main", "B200", "node-1", "0", "1", "x y", "ü中", "q\"uote", "back\\slash",
            "tab\t", "]]}", '},"values":[']
    out = []
    for _ in range(n):
        m = {}
        for k in rng.sample(keys, rng.randrange(0, len(keys))):
            v = rng.choice(vals)
            if rng.random() < 0.03:
                v = rng.choice([3, None, True, ["l"], {"o": 1}])      # non-string values exist in JSON
            m[k] = v
        style = rng.random()
        if style < 0.8:
            s = json.dumps(m, separators=(",", ":"), ensure_ascii=False)
        elif style < 0.9:
            s = json.dumps(m, ensure_ascii=False)                     # spaces after , and :
        else:
            s = json.dumps(m, separators=(",", ":"))                  # \\uXXXX escapes for non-ASCII
        if rng.random() < 0.03 and m:                                  # duplicate key: the last one wins in the DOM
            k = next(iter(m))
            s = s[:-1] + "," + json.dumps(k) + ":" + json.dumps("dup") + "}"
        out.append((s, m))
    return out


def _flat_shape(s):
    """the shape FlatLabels claims: no backslash, no control byte, no whitespace outside strings, all values
    strings, unique keys"""
    if "\\" in s or any(ord(c) < 0x20 for c in s):
        return False
    try:
        pairs = json.loads(s, object_pairs_hook=list)
    except ValueError:
        return False
    if not isinstance(pairs, list) or not all(isinstance(v, str) for _, v in pairs):
        return False
    if len({k for k, _ in pairs}) != len(pairs):
        return False
    return s == "{" + ",".join(json.dumps(k, ensure_ascii=False) + ":" + json.dumps(v, ensure_ascii=False)
                               for k, v in pairs) + "}"


@pytest.mark.parametrize("is_power", [0, 1])
def test_flat_view_and_dom_assign_identically(is_power, driver, tmp_path):
    rng = random.Random(99 + is_power)
    maps = _maps(rng, 3000)
    f = tmp_path / "maps.txt"
    f.write_text("\n".join(s for s, _ in maps) + "\n", encoding="utf-8")
    r = subprocess.run([driver, str(f), str(is_power)], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1"))
    lines = r.stdout.splitlines()
    assert r.returncode == 0 and lines[-1] == "WINDOWS_EQUAL", (r.stderr[-2000:], lines[-3:])
    assert len(lines) == len(maps) + 1
    n_flat = n_placed = 0
    for (s, m), line in zip(maps, lines):
        parts = line.split()
        accepted = parts[0] == "1"
        assert accepted == _flat_shape(s), (s, line)
        if accepted:
            n_flat += 1
            assert parts[1:4] == parts[4:7], (s, line)       # same result, pod and slot through both views
        n_placed += parts[1] == "P"
    assert n_flat > 500 and n_placed > 100
