"""Long mutation-fuzz run of the CPU matrix parser (gpu-pruner_b200/host/ingest.cpp) against the DOM reference path,
under ASan + UBSan:  cpu_parser.py DRIVER SEED ROUNDS   (400 inputs per round)
DRIVER = tests/cpp/ingest_fuzz_driver.cpp built as in tests/test_ingest_fuzz.py."""
import os, random, shutil, subprocess, sys, tempfile
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'tests'))
import test_ingest_fuzz as F
drv = sys.argv[1]; seed0 = int(sys.argv[2]); rounds = int(sys.argv[3])
bad = 0
for r in range(rounds):
    rng = random.Random(seed0 * 100003 + r)
    tmp = tempfile.mkdtemp(prefix='cpufuzz')
    files = []
    for i in range(400):
        s = F._valid(rng)
        data = s.encode() if i % 5 == 0 else F._mutate(rng, s)
        if rng.random() < 0.2:
            data = F._mutate(rng, data.decode('latin1'))   # a second mutation
        p = os.path.join(tmp, f"in_{i}.json")
        open(p, 'wb').write(data if isinstance(data, bytes) else data.encode('latin1'))
        files.append(p)
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1")
    res = subprocess.run([drv] + files, capture_output=True, text=True, timeout=900, env=env)
    lines = res.stdout.splitlines()
    mm = [l for l in lines if l.startswith("MISMATCH")]
    if res.returncode != 0 or len(lines) != len(files) or mm:
        bad += 1
        print("ROUND", seed0, r, res.returncode, len(lines), mm[:3], res.stderr[-1200:]); print("kept", tmp)
    else:
        shutil.rmtree(tmp)
print("seed", seed0, "rounds", rounds, "bad", bad)
