"""Long run of random daemon-mode timelines through DeviceIngestSession on the emulated ring: at every tick the ring
must equal a fresh full-range ingest:  daemon_ticks.py DRIVER FIRST_SEED N   (DRIVER as for device_ingest.py)"""
import sys, os, random, subprocess, tempfile, pathlib, shutil
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'tests'))
import test_resident_ticks as R
import ticks as TK
drv=sys.argv[1]; s0=int(sys.argv[2]); n=int(sys.argv[3])
bad=0
for seed in range(s0, s0+n):
    rng = random.Random(50000 + seed)
    step = rng.choice([1, 2, 5, 10, 15])
    duration_min = rng.choice([1, 2, 3])
    N = duration_min * 60
    interval = step * rng.randrange(1, 14)
    t0 = 1_700_000_000 + rng.randrange(1000)
    times = [t0 + N + k * interval for k in range(rng.randrange(3, 10))]
    horizon = times[-1] + 5
    store = []
    for p in range(rng.randrange(1, 9)):
        for g in range(rng.randrange(1, 5)):
            a = rng.choice([t0, t0, rng.randrange(t0, horizon)])
            b = rng.choice([horizon, horizon, rng.randrange(a, horizon + 1)])
            store.append(R._series(rng, f"p{p}", g, a, b, step, rng.choice(["idle", "busy"]), jitter=rng.random() < 0.5))
            if rng.random() < 0.25:
                store.append(R._series(rng, f"p{p}", g, a, b, step, "busy", metric="DCGM_FI_PROF_GR_ENGINE_ACTIVE"))
            if rng.random() < 0.15:
                store.append(R._series(rng, f"p{p}", g, a, b, step, "busy", UUID="dup-%d" % rng.randrange(3)))
            if rng.random() < 0.5:
                store.append(R._series(rng, f"p{p}", g, a, b, step, "x", metric="DCGM_FI_DEV_POWER_USAGE"))
    d = pathlib.Path(tempfile.mkdtemp(prefix='tickfuzz'))
    skip = {rng.randrange(1, len(times))} if rng.random() < 0.3 else ()
    TK.write_ticks(str(d), lambda k: store, times, N, step, with_power=rng.random()<0.7, skip_delta=skip)
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1")
    r = subprocess.run([drv, "--ticks", str(duration_min), str(d)], capture_output=True, text=True, timeout=600, env=env)
    lines = r.stdout.splitlines()
    if r.returncode != 0 or not lines or not all(l.startswith("OK ") for l in lines):
        bad += 1; print("SEED", seed, r.returncode, [l for l in lines if not l.startswith("OK ")][:3], r.stderr[-800:]); print("kept", d)
    else:
        shutil.rmtree(d)
print("from", s0, "n", n, "bad", bad)
