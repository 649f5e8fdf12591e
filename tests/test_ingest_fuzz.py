"""CPU: mutation fuzzing of the hand-written matrix parser under AddressSanitizer + UBSan.

The text ingest path (gpu-pruner_b200/host/ingest.cpp) scans raw bytes with memmem / SWAR and its own
number parser; truncated or corrupted responses must be rejected with an exception — never a crash, an
out-of-bounds read or a hang — and every input the DOM reference path accepts must give the same tensor."""
import json
import os
import random
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "gpu-pruner_b200", "host")


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    out = tmp_path_factory.mktemp("fuzz") / "ingest_fuzz"
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
           "-I", HOST, os.path.join(ROOT, "tests", "cpp", "ingest_fuzz_driver.cpp"),
           os.path.join(HOST, "ingest.cpp"), os.path.join(HOST, "json.cpp"), "-o", str(out), "-lpthread"]
    subprocess.check_call(cmd)
    return str(out)


def _valid(rng):
    t_end = 1_700_000_000
    sers = []
    for p in range(rng.randrange(1, 5)):
        for g in range(rng.randrange(1, 3)):
            lab = {"Hostname": "n", "gpu": str(g), "modelName": "m ]] \"q\" \\", "exported_pod": f"p{p}",
                   "exported_namespace": "ns", "exported_container": "c"}
            vals = [[t_end - 59 + i, rng.choice(["0", "7", "0.5", "NaN", "1e2", "+Inf"])] for i in range(rng.randrange(0, 60))]
            ser = {"metric": lab, "values": vals}
            if rng.random() < 0.3:
                ser = {"values": vals, "metric": lab}          # member order is not guaranteed
            sers.append(ser)
    return json.dumps({"status": "success", "data": {"resultType": "matrix", "result": sers}},
                      separators=rng.choice([(",", ":"), (", ", ": ")]))


def _mutate(rng, s):
    b = bytearray(s.encode())
    kind = rng.randrange(6)
    if kind == 0 and b:                      # truncate
        del b[rng.randrange(len(b)):]
    elif kind == 1 and b:                    # flip bytes
        for _ in range(rng.randrange(1, 6)):
            b[rng.randrange(len(b))] = rng.choice(b'[]{}",:\\0 e-+.x')
    elif kind == 2 and b:                    # delete a slice
        i = rng.randrange(len(b))
        del b[i:i + rng.randrange(1, 12)]
    elif kind == 3:                          # duplicate a slice
        i = rng.randrange(len(b) + 1)
        b[i:i] = b[max(0, i - rng.randrange(1, 30)):i]
    elif kind == 4:                          # garbage tail
        b += bytes(rng.choice(b']}"[,x') for _ in range(rng.randrange(1, 8)))
    return bytes(b)


def test_mutated_responses_never_crash(driver, tmp_path):
    rng = random.Random(20260921)
    files = []
    for i in range(400):
        s = _valid(rng)
        data = s.encode() if i % 5 == 0 else _mutate(rng, s)
        p = tmp_path / f"in_{i}.json"
        p.write_bytes(data)
        files.append(str(p))
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1")
    r = subprocess.run([driver] + files, capture_output=True, text=True, timeout=600, env=env)
    lines = r.stdout.splitlines()
    assert r.returncode == 0, (r.stderr[-3000:], [l for l in lines if l.startswith("MISMATCH")][:5])
    assert len(lines) == len(files)                       # no crash half-way
    verdicts = [l.split()[0] for l in lines]
    assert verdicts.count("OK") >= 80 and verdicts.count("REJECT") >= 50
