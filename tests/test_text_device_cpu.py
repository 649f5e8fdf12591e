"""CPU: the device ingest (matrix JSON parsed by the GPU straight into the tensor) on an EMULATED device.

gpu-pruner_b200/csrc/gpr_text.cuh keeps every byte-level decision in functions that compile for host and
device; tests/cpp/text_emul.cpp runs them slice by slice on the CPU behind the same TextDevice interface
the CUDA implementation has, so the whole orchestration (marker scan -> label maps -> spans -> parse ->
hard-row patching) is checked here, without a GPU, against the CPU text ingest: identical shape, pods,
statistics and tensor cells for every accepted input, rejection for every rejected one, under ASan/UBSan.
The GPU run of the same kernels is tests/test_gpu_text.py."""
import json
import os
import random
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "gpu-pruner_b200", "host")
T_END = 1_700_000_000


@pytest.fixture(scope="module", params=["tiles", "kernel"])
def driver(request, tmp_path_factory):
    """both flavours of the emulated device (tests/emul_build.py): parser core tile by tile / k_text_parse's source"""
    import emul_build
    return emul_build.build(tmp_path_factory.mktemp("emul_" + request.param), request.param)


def _run(driver, dirs, step=1, duration_min=1, t_end=T_END):
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1")
    r = subprocess.run([driver, str(t_end), str(step), str(duration_min)] + [str(d) for d in dirs],
                       capture_output=True, text=True, timeout=600, env=env)
    lines = r.stdout.splitlines()
    assert len(lines) == len(dirs), (r.stderr[-3000:], lines[-3:])
    return r.returncode, lines


def _dump(obj):
    return json.dumps(obj, separators=(",", ":"))


def _resp(series):
    return _dump({"status": "success", "data": {"resultType": "matrix", "result": series}})


def _labels(p, g, **extra):
    lab = {"Hostname": f"node-{p % 3}", "gpu": str(g), "modelName": "NVIDIA B200", "exported_pod": f"pod-{p}",
           "exported_namespace": f"ns-{p % 2}", "exported_container": "main", "UUID": f"GPU-{p}-{g}"}
    lab.update(extra)
    return lab


VALUES = ["0", "0", "0", "7", "100", "0.5", "NaN", "1e2", "+Inf", "-Inf", "Inf", "-0", "+3", "12.25", "1E1",
          "0.30000000000000004", "1e-50", "123456789012345678", "1e23", "0.1e-3"]


def _case(rng, tmp, name, *, prof=False, power=False, pretty=False, n_pods=None, T=60, frac_ts=False,
          collide=False, backwards=False, dup=False, values=VALUES):
    d = tmp / name
    d.mkdir()
    n_pods = n_pods or rng.randrange(1, 6)

    def samples(n):
        out, t = [], T_END - T + 1
        for i in range(n):
            ts = t + i
            if frac_ts and rng.random() < 0.3:
                ts = ts + rng.choice([0.123, 0.5, 0.499, 0.9])
            out.append([ts, rng.choice(values)])
            if collide and rng.random() < 0.05:
                out.append([ts + 0.25, rng.choice(values)])
        if backwards and len(out) > 3:
            i = rng.randrange(1, len(out) - 1)
            out[i], out[i + 1] = out[i + 1], out[i]
        return out

    util, pr, pw = [], [], []
    for p in range(n_pods):
        for g in range(rng.randrange(1, 4)):
            n = rng.choice([0, 1, 2, T // 2, T, T, T + 5])
            util.append({"metric": dict(_labels(p, g), __name__="DCGM_FI_DEV_GPU_UTIL"), "values": samples(n)})
            if dup and rng.random() < 0.4:   # same `sum by` group, different UUID: merged into one row
                util.append({"metric": dict(_labels(p, g, UUID="other"), __name__="DCGM_FI_DEV_GPU_UTIL"),
                             "values": samples(rng.choice([1, T]))})
            if prof and rng.random() < 0.5:
                same = rng.random() < 0.5     # identical label set shadows the UTIL series, else both merge
                pr.append({"metric": dict(_labels(p, g) if same else _labels(p, g, UUID="prof"),
                                          __name__="DCGM_FI_PROF_GR_ENGINE_ACTIVE"),
                           "values": samples(T)})
            if power:
                pw.append({"metric": dict(_labels(p, g), __name__="DCGM_FI_DEV_POWER_USAGE"),
                           "values": [[T_END - T + 1 + i, rng.choice(["55.5", "149.99", "150", "420.125"])]
                                      for i in range(rng.choice([0, T]))]})
    if rng.random() < 0.3:   # a series that cannot be turned into a pod (no modelName): skipped
        util.append({"metric": {"Hostname": "x", "gpu": "0", "exported_pod": "nomodel", "exported_namespace": "n",
                                "exported_container": "c"}, "values": samples(3)})
    enc = (lambda s: json.dumps({"status": "success", "data": {"resultType": "matrix", "result": s}},
                                separators=(", ", ": "))) if pretty else _resp
    (d / "util.json").write_text(enc(util))
    if prof:
        (d / "prof.json").write_text(enc(pr))
    if power:
        (d / "power.json").write_text(enc(pw))
    return d


def test_device_path_matches_cpu_text_path(driver, tmp_path):
    rng = random.Random(7)
    dirs = []
    for i in range(40):
        dirs.append(_case(rng, tmp_path, f"plain{i}", values=["0", "0", "7", "100", "NaN", "12.5"]))
    for i in range(30):
        dirs.append(_case(rng, tmp_path, f"mixed{i}", prof=i % 2 == 0, power=i % 3 == 0, dup=i % 4 == 0,
                          frac_ts=i % 5 == 0))
    for i in range(15):
        dirs.append(_case(rng, tmp_path, f"collide{i}", collide=True, frac_ts=True))
    for i in range(10):
        dirs.append(_case(rng, tmp_path, f"back{i}", backwards=True))
    rc, lines = _run(driver, dirs)
    assert rc == 0, [l for l in lines if not l.startswith("OK")][:5]
    assert all(l.startswith("OK") and " device=1 " in l for l in lines), [l for l in lines if " device=1 " not in l][:3]
    plain = lines[:40]
    # integers / short decimals / NaN: nothing for the CPU to redo
    assert all(" hard=0 patched=0" in l for l in plain), [l for l in plain if " hard=0 " not in l][:3]
    # the awkward values (17 digits, exponents beyond the exact range, collisions, time going backwards)
    # went through the hard-span path and still agree
    assert sum(int(l.split("hard=")[1].split()[0]) for l in lines[40:]) > 20


def test_long_window_and_steps(driver, tmp_path):
    rng = random.Random(11)
    dirs = [_case(rng, tmp_path, f"long{i}", n_pods=3, T=1800, values=["0", "0", "0", "37", "100"]) for i in range(3)]
    rc, lines = _run(driver, dirs, step=1, duration_min=30)
    assert rc == 0 and all(" device=1 " in l and " hard=0 " in l for l in lines), lines
    # 15 s scrape interval: columns are (t_end - ts + 7) / 15
    d = tmp_path / "step15"
    d.mkdir()
    vals = [[T_END - 15 * i, str(i % 3)] for i in range(119, -1, -1)]
    (d / "util.json").write_text(_resp([{"metric": _labels(0, 0), "values": vals}]))
    rc, lines = _run(driver, [d], step=15, duration_min=30)
    assert rc == 0 and " device=1 " in lines[0] and " hard=0 " in lines[0], lines


def test_tricky_label_values_do_not_confuse_the_scan(driver, tmp_path):
    """byte patterns of the markers inside label strings (where they can only appear escaped)"""
    d = tmp_path / "labels"
    d.mkdir()
    sers = []
    nasty = ['x"]]y', 'a},"values":[b', '\\"]]', 'tab\there', 'unié中', ']]}', '{"metric":{', 'back\\slash"]]']
    for i, s in enumerate(nasty):
        sers.append({"metric": dict(_labels(i, 0), note=s, modelName="m " + s),
                     "values": [[T_END - 5 + k, "0"] for k in range(5)]})
    (d / "util.json").write_text(_resp(sers))
    rc, lines = _run(driver, [d])
    assert rc == 0 and lines[0].startswith("OK") and " device=1 " in lines[0] and " hard=0 " in lines[0], lines


def test_anything_but_the_compact_encoding_falls_back_to_the_cpu_parser(driver, tmp_path):
    rng = random.Random(3)
    pretty = _case(rng, tmp_path, "pretty", pretty=True)
    bare = tmp_path / "bare"
    bare.mkdir()
    (bare / "util.json").write_text(_dump([{"metric": _labels(0, 0), "values": [[T_END, "0"]]}]))
    extra = tmp_path / "extra"
    extra.mkdir()
    (extra / "util.json").write_text(_dump({"status": "success", "data": {"resultType": "matrix", "result": [
        {"metric": _labels(0, 0), "values": [[T_END, "0"]]}]}, "warnings": ["w"]}))
    swapped = tmp_path / "swapped"
    swapped.mkdir()
    (swapped / "util.json").write_text(_resp([{"values": [[T_END, "0"]], "metric": _labels(0, 0)}]))
    hist = tmp_path / "hist"
    hist.mkdir()
    (hist / "util.json").write_text(_resp([{"metric": _labels(0, 0), "values": [[T_END, "0"]], "histograms": []}]))
    empty = tmp_path / "empty"
    empty.mkdir()
    (empty / "util.json").write_text(_resp([]))
    indented = tmp_path / "indented"
    indented.mkdir()
    (indented / "util.json").write_text(json.dumps({"status": "success", "data": {"resultType": "matrix", "result": [
        {"metric": _labels(0, 0), "values": [[T_END - 1, "0"], [T_END, "4"]]}]}}, indent=2))
    rc, lines = _run(driver, [pretty, bare, extra, swapped, hist, empty, indented])
    assert rc == 0, lines
    assert lines[6].startswith("OK") and " device=0 " in lines[6]
    assert " device=0 " in lines[0] and " device=1 " in lines[1] and " device=0 " in lines[2]
    assert " device=0 " in lines[3] and " device=0 " in lines[4] and " device=1 " in lines[5]
    # and without end / step the device path is not attempted at all
    rc, lines = _run(driver, [bare], step=0, t_end=0)
    assert rc == 0 and " device=0 " in lines[0]


def test_error_status_is_rejected_by_both(driver, tmp_path):
    d = tmp_path / "err"
    d.mkdir()
    (d / "util.json").write_text(_dump({"status": "error", "errorType": "bad_data", "error": "boom"}))
    v = tmp_path / "vector"
    v.mkdir()
    (v / "util.json").write_text(_dump({"status": "success", "data": {"resultType": "vector", "result": []}}))
    rc, lines = _run(driver, [d, v])
    assert rc == 0 and all(l.startswith("REJECT") for l in lines), lines


def _mutate(rng, s):
    b = bytearray(s.encode())
    kind = rng.randrange(5)
    if kind == 0 and b:
        del b[rng.randrange(len(b)):]
    elif kind == 1 and b:
        for _ in range(rng.randrange(1, 6)):
            b[rng.randrange(len(b))] = rng.choice(b'[]{}",:\\0 e-+.x')
    elif kind == 2 and b:
        i = rng.randrange(len(b))
        del b[i:i + rng.randrange(1, 12)]
    elif kind == 3:
        i = rng.randrange(len(b) + 1)
        b[i:i] = b[max(0, i - rng.randrange(1, 30)):i]
    else:
        b += bytes(rng.choice(b']}"[,x') for _ in range(rng.randrange(1, 8)))
    return bytes(b)


def test_mutated_responses_agree_or_are_rejected(driver, tmp_path):
    """corrupted responses: the device path (emulated) must never crash, read out of bounds, or produce a
    tensor the CPU path would not — whatever the corruption hits (markers, samples, labels, structure)"""
    rng = random.Random(20260921)
    dirs = []
    for i in range(300):
        d = _case(rng, tmp_path, f"m{i}", values=["0", "7", "0.5", "NaN", "1e2"], T=20)
        p = d / "util.json"
        p.write_bytes(_mutate(rng, p.read_text()))
        dirs.append(d)
    rc, lines = _run(driver, dirs)
    assert rc == 0, [l for l in lines if l.startswith("MISMATCH")][:5]
    verdicts = [l.split()[0] for l in lines]
    assert verdicts.count("OK") >= 30 and verdicts.count("REJECT") >= 30


def test_samples_of_series_without_a_row_are_still_checked(driver, tmp_path):
    """A series that gets no tensor row (no workload-pod label) never reaches the device parser; garbage inside its
    values list must not slip through — the CPU parser reads every sample and rejects such a response.  (Found by an
    extended run of the mutation fuzz: one mutation hit a label key, a second one a sample of the same series.)"""
    ok_series = {"metric": dict(_labels(1, 0), __name__="DCGM_FI_DEV_GPU_UTIL"), "values": [[T_END - 1, "0"], [T_END, "7"]]}
    orphan = {"metric": {"Hostname": "node-9", "gpu": "0", "modelName": "NVIDIA B200", "__name__": "DCGM_FI_DEV_GPU_UTIL"},
              "values": [[T_END - 2, "0.5"], [T_END - 1, "NaN"], [T_END, "1e2"]]}
    good = _resp([ok_series, orphan])
    cases = {"fine": good,
             "quote": good.replace('"NaN"]', '"NaN]]'),
             "bracket": good.replace(f'[{T_END - 2},"0.5"]', f'{T_END - 2},"0.5"]'),
             "junk": good.replace('"1e2"]]', '"1e2"]x]')}
    dirs = []
    for name, text in cases.items():
        d = tmp_path / name
        d.mkdir()
        (d / "util.json").write_text(text)
        dirs.append(d)
    rc, lines = _run(driver, dirs)
    assert rc == 0, lines
    assert lines[0].startswith("OK") and " device=1 " in lines[0], lines[0]
    # the device path never accepts them: the CPU parser is the judge (it rejects two and is lenient about one)
    assert all(" device=1 " not in l for l in lines[1:]), lines
    assert sum(l.startswith("REJECT") for l in lines[1:]) >= 2, lines


def test_a_device_that_declines_hands_the_response_to_the_cpu_parser(driver, tmp_path):
    """The scan has room for a fixed number of series markers per upload chunk; a response that exceeds it (label sets
    a tenth of DCGM's size) is not malformed, so the tick must not fail: the device declines, the CPU parser produces
    the window.  The emulated device declines in the middle of the upload pipeline here."""
    rng = random.Random(3)
    dirs = [_case(rng, tmp_path, f"d{i}", n_pods=4, T=60, prof=i == 1, power=i == 2) for i in range(3)]
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1", EMUL_DECLINE_SCAN="1")
    r = subprocess.run([driver, str(T_END), "1", "1"] + [str(d) for d in dirs], capture_output=True, text=True,
                       timeout=600, env=env)
    lines = r.stdout.splitlines()
    assert r.returncode == 0 and len(lines) == 3, (r.stdout, r.stderr[-2000:])
    assert all(l.startswith("OK") and " device=0 " in l and "too many markers" in l for l in lines), lines
