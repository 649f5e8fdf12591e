"""Long differential run of the device ingest on the emulated device (tests/cpp/text_emul.cpp) against the CPU text
parser, well-formed and mutated responses:  device_ingest.py DRIVER SEED ROUNDS   (120 responses per round)
DRIVER = text_emul in either flavour, built by tests/emul_build.py (ASan + UBSan): build(dir, "tiles" | "kernel")."""
import sys, os, random, subprocess, tempfile, pathlib, shutil
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'tests'))
import test_text_device_cpu as T
drv=sys.argv[1]; seed0=int(sys.argv[2]); rounds=int(sys.argv[3])
bad=0
for r in range(rounds):
    rng=random.Random(seed0*100000+r)
    tmp=pathlib.Path(tempfile.mkdtemp(prefix='devfuzz'))
    dirs=[]
    for i in range(60):
        kw=dict(prof=rng.random()<0.4, power=rng.random()<0.3, dup=rng.random()<0.3, frac_ts=rng.random()<0.4,
                collide=rng.random()<0.3, backwards=rng.random()<0.2, T=rng.choice([5,20,60,61,127]))
        dirs.append(T._case(rng, tmp, f"c{i}", **kw))
    for i in range(60):
        d=T._case(rng, tmp, f"m{i}", values=["0","7","0.5","NaN","1e2","0.30000000000000004"], T=rng.choice([7,20,33]))
        p=d/"util.json"; p.write_bytes(T._mutate(rng, p.read_text())); dirs.append(d)
    step=rng.choice([1,1,1,2,5]); dur=rng.choice([1,1,2,3])
    env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1")
    res=subprocess.run([drv,str(T.T_END),str(step),str(dur)]+[str(d) for d in dirs],capture_output=True,text=True,timeout=900,env=env)
    lines=res.stdout.splitlines()
    mm=[l for l in lines if l.startswith("MISMATCH")]
    if res.returncode!=0 or len(lines)!=len(dirs) or mm:
        bad+=1; print("ROUND",seed0,r,"rc",res.returncode,len(lines),mm[:3],res.stderr[-1500:])
        print("kept",tmp)
    else:
        shutil.rmtree(tmp)
print("seed",seed0,"rounds",rounds,"bad",bad)
