"""CPU: daemon mode's resident window (gpu-pruner_b200/host/ingest_device.cpp DeviceIngestSession) on the EMULATED
device (tests/cpp/text_emul.cpp --ticks): first tick = the full range into the ring, later ticks = only what was
scraped since, appended.  After every tick the ring must hold exactly what a fresh full-range ingest of that tick
yields — the reference re-runs the whole [Nm] query every --check-interval (main.rs:286-330); keeping the window in HBM
must never change what is decided.  The GPU run of the same scenarios through the `gpu-pruner` binary is
tests/test_gpu_daemon.py."""
import os
import random
import subprocess

import pytest

import ticks as TK

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "gpu-pruner_b200", "host")


@pytest.fixture(scope="module", params=["tiles", "kernel"])
def driver(request, tmp_path_factory):
    """both flavours of the emulated device (tests/emul_build.py): parser core tile by tile / k_text_parse's source"""
    import emul_build
    return emul_build.build(tmp_path_factory.mktemp("emul_ticks_" + request.param), request.param)


def _run(driver, root, duration_min):
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1")
    r = subprocess.run([driver, "--ticks", str(duration_min), str(root)], capture_output=True, text=True, timeout=600, env=env)
    lines = r.stdout.splitlines()
    assert r.returncode == 0 and lines and all(l.startswith("OK ") for l in lines), (r.stdout[-3000:], r.stderr[-2000:])
    return [dict(kv.split("=", 1) for kv in l.split()[1:3]) | {"why": " ".join(l.split()[3:])} for l in lines]


def _series(rng, pod, gpu, t0, t1, step, kind, metric="DCGM_FI_DEV_GPU_UTIL", jitter=False, **extra):
    smp = []
    t = t0 + rng.randrange(step)
    while t <= t1:
        ts = t + (rng.choice([-0.4, 0, 0.25, 0.499]) if jitter else 0)
        v = 0 if kind == "idle" else (rng.choice([0, 0, 37, 100]) if kind == "busy" else 0.0)
        if metric == "DCGM_FI_PROF_GR_ENGINE_ACTIVE":
            v = 0.0 if kind == "idle" else rng.random()
        if metric == "DCGM_FI_DEV_POWER_USAGE":
            v = rng.choice([55, 60, 149.5, 310.25])
        if rng.random() < 0.02:
            v = "NaN"
        if rng.random() > 0.03:                       # scrape gaps
            smp.append((ts, v))
        t += step + (rng.choice([-1, 0, 0, 1]) if jitter and step > 2 else 0)
    return (metric, TK.labels(pod, gpu, **extra), smp)


@pytest.mark.parametrize("step,interval,duration_min", [(1, 20, 1), (5, 45, 2), (15, 180, 30)])
def test_steady_state_appends_equal_fresh_queries(driver, tmp_path, step, interval, duration_min):
    rng = random.Random(step * 1000 + interval)
    N = duration_min * 60
    t0 = 1_700_000_000
    times = [t0 + N + k * interval for k in range(7)]
    store = [_series(rng, f"pod-{p}", g, t0 - 50, times[-1] + 10, step, rng.choice(["idle", "busy"]), jitter=step > 1)
             for p in range(6) for g in range(rng.randrange(1, 4))]
    TK.write_ticks(str(tmp_path), lambda k: store, times, N, step)
    modes = _run(driver, tmp_path, duration_min)
    assert [m["mode"] for m in modes] == ["full"] + ["delta"] * 6


def test_series_come_and_go_and_age_out(driver, tmp_path):
    rng = random.Random(5)
    N, step, interval = 120, 2, 30
    t0 = 1_700_000_000
    times = [t0 + N + k * interval for k in range(10)]
    horizon = times[-1] + 5
    base = [_series(rng, f"pod-{p}", g, t0, horizon, step, "busy") for p in range(4) for g in range(2)]
    leaves = _series(rng, "leaver", 0, t0, times[2] - 3, step, "idle")           # stops reporting, ages out by tick 7
    joins = _series(rng, "joiner", 0, times[3] + 1, horizon, step, "idle")        # a NEW pod: gets a spare row
    second = _series(rng, "pod-0", 1, times[4] + 1, horizon, step, "idle", UUID="GPU-late")   # duplicate of an existing group
    store = base + [leaves, joins, second]
    TK.write_ticks(str(tmp_path), lambda k: store, times, N, step)
    modes = _run(driver, tmp_path, 2)
    # pod-0 already has two slots (G = 2): the late duplicate needs a third -> the ring's shape no longer fits
    assert [m["mode"] for m in modes] == ["full"] * 1 + ["delta"] * 4 + ["full"] + ["delta"] * 4
    assert "GPU slot" in modes[5]["why"]


def test_gap_tick_and_prof_changes_force_the_full_range(driver, tmp_path):
    rng = random.Random(9)
    N, step, interval = 60, 1, 15
    t0 = 1_700_000_000
    times = [t0 + N + k * interval for k in range(8)]
    horizon = times[-1] + 5
    util = [_series(rng, f"pod-{p}", 0, t0, horizon, step, "idle") for p in range(3)]
    prof_same = ("DCGM_FI_PROF_GR_ENGINE_ACTIVE", util[0][1], [(t, 0.25 if t % 7 == 0 else 0.0) for t in range(t0, horizon)])
    # a PROF series with the label set of pod-1's UTIL series starts late: from then on it shadows the UTIL series
    prof_late = ("DCGM_FI_PROF_GR_ENGINE_ACTIVE", util[1][1], [(t, 0.5) for t in range(times[4] + 2, horizon)])
    power = [_series(rng, f"pod-{p}", 0, t0, horizon, step, "x", metric="DCGM_FI_DEV_POWER_USAGE") for p in range(3)]
    store = util + [prof_same, prof_late] + power
    TK.write_ticks(str(tmp_path), lambda k: store, times, N, step, with_power=True, skip_delta={2})
    modes = _run(driver, tmp_path, 1)
    assert [m["mode"] for m in modes] == ["full", "delta", "full", "delta", "delta", "full", "delta", "delta"]
    assert "PROF" in modes[5]["why"] or "GPU slot" in modes[5]["why"]   # the new PROF series needs a row of its own


def test_fuzz_timelines(driver, tmp_path):
    """random clusters over random tick schedules: whatever path the session takes, the ring equals a fresh ingest"""
    for seed in range(12):
        rng = random.Random(1000 + seed)
        step = rng.choice([1, 2, 10])
        duration_min = rng.choice([1, 2])
        N = duration_min * 60
        interval = step * rng.randrange(2, 12)
        t0 = 1_700_000_000 + rng.randrange(1000)
        times = [t0 + N + k * interval for k in range(rng.randrange(4, 9))]
        horizon = times[-1] + 5
        store = []
        for p in range(rng.randrange(2, 7)):
            for g in range(rng.randrange(1, 4)):
                a = rng.choice([t0, t0, rng.randrange(t0, horizon)])
                b = rng.choice([horizon, horizon, rng.randrange(a, horizon + 1)])
                store.append(_series(rng, f"p{p}", g, a, b, step, rng.choice(["idle", "busy"]), jitter=rng.random() < 0.5))
                if rng.random() < 0.2:
                    store.append(_series(rng, f"p{p}", g, a, b, step, "busy", metric="DCGM_FI_PROF_GR_ENGINE_ACTIVE"))
                if rng.random() < 0.5:
                    store.append(_series(rng, f"p{p}", g, a, b, step, "x", metric="DCGM_FI_DEV_POWER_USAGE"))
        d = tmp_path / f"s{seed}"
        TK.write_ticks(str(d), lambda k: store, times, N, step, with_power=True,
                       skip_delta={rng.randrange(1, len(times))} if rng.random() < 0.3 else ())
        _run(driver, d, duration_min)


def test_many_series_worker_pool_under_thread_sanitizer(tmp_path):
    """More series than the worker pool's threshold (256): the label maps of a tick are parsed on several threads
    (ingest_device.cpp Workers / plan_text).  Same invariant as everywhere in this file — the ring equals a fresh
    full-range ingest at every tick — with the emulator built under ThreadSanitizer."""
    import emul_build
    exe = emul_build.build(tmp_path, "kernel", sanitize="thread")   # worker pool AND the parse kernel's threads
    rng = random.Random(77)
    N, step, interval = 60, 5, 15
    t0 = 1_700_000_000
    times = [t0 + N + k * interval for k in range(4)]
    horizon = times[-1] + 5
    store = [_series(rng, f"pod-{p}", g, t0, horizon, step, rng.choice(["idle", "busy"])) for p in range(150) for g in range(4)]
    store.append(_series(rng, "joiner", 0, times[1] + 1, horizon, step, "idle"))
    root = tmp_path / "ticks"
    TK.write_ticks(str(root), lambda k: store, times, N, step)
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1", GPR_LABEL_THREADS="6")
    r = subprocess.run([str(exe), "--ticks", "1", str(root)], capture_output=True, text=True, timeout=900, env=env)
    lines = r.stdout.splitlines()
    assert r.returncode == 0 and len(lines) == 4 and all(l.startswith("OK ") for l in lines), (r.stdout[-2000:], r.stderr[-3000:])
    assert "ThreadSanitizer" not in r.stderr


def test_failed_tick_invalidates_the_resident_window(driver, tmp_path):
    """A tick whose slice dies half-way (device error in the parse) has already advanced the ring; the next tick must
    not append to it — the slice that never arrived would read as "no samples" and a busy GPU as idle.  The session
    forgets the ring and the next tick rebuilds it from the full range."""
    rng = random.Random(21)
    N, step, interval = 60, 1, 15
    t0 = 1_700_000_000
    times = [t0 + N + k * interval for k in range(6)]
    horizon = times[-1] + 5
    # busy only inside the slice of tick 2: lose that slice and these pods look idle
    store = [(m, l, [(t, (50 if times[1] < t <= times[2] else 0)) for t, _ in smp])
             for m, l, smp in (_series(rng, f"pod-{p}", g, t0, horizon, step, "idle") for p in range(5) for g in range(2))]
    TK.write_ticks(str(tmp_path), lambda k: store, times, N, step)
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1", EMUL_FAIL_PARSE_TICK="2")
    r = subprocess.run([driver, "--ticks", "1", str(tmp_path)], capture_output=True, text=True, timeout=600, env=env)
    lines = r.stdout.splitlines()
    assert r.returncode == 0 and all(l.startswith("OK ") for l in lines), (r.stdout[-3000:], r.stderr[-2000:])
    modes = [l.split()[2].split("=", 1)[1] for l in lines]
    assert modes == ["full", "delta", "failed", "full", "delta", "delta"], lines
