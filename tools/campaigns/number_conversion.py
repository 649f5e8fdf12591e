"""Long differential run of the device text parser's number conversion (gpr_text.cuh compiled for the host by
tests/cpp/number_check.cpp) against Python's correctly rounded float():  number_conversion.py SEED N_CASES
Build first:  g++ -O2 -std=c++17 tests/cpp/number_check.cpp -o /tmp/numf/number_check"""
import os
import random, struct, subprocess, sys, numpy as np
def bits64(x): return struct.unpack("<Q", struct.pack("<d", x))[0]
def bits32(x): return struct.unpack("<I", struct.pack("<f", x))[0]
seed=int(sys.argv[1]); rng=random.Random(seed)
N=int(sys.argv[2])
bad=0; declined_real=0
for chunk in range(N//200000):
    E=[]; V=[]
    for _ in range(100000):
        r=rng.random()
        if r<0.5:
            nd=rng.choice([15,16,17,17,17,18,19]); man=rng.randrange(10**(nd-1),10**nd)
            if man>=1<<64: continue
            E.append((man, rng.randrange(-40,25)))
        else:
            x=struct.unpack("<d", struct.pack("<Q", rng.getrandbits(64)))[0]
            if x!=x or x in (float('inf'),float('-inf')): continue
            s=repr(abs(x))
            if 'e' in s:
                m,e=s.split('e'); e=int(e)
            else: m,e=s,0
            if '.' in m:
                ip,fp=m.split('.'); man=int(ip+fp); e-=len(fp)
            else: man=int(m)
            if man<1<<64: E.append((man,e))
    for _ in range(100000):
        r=rng.random()
        if r<0.4: x=rng.random()
        elif r<0.6: x=rng.uniform(0,1000)
        elif r<0.8: x=rng.random()*10.0**rng.randrange(-12,12)
        else: x=float(rng.randrange(0,101))+rng.choice([0,0.5,0.25,0.1])
        V.append(repr(x))
    inp="\n".join([f"E {m} {e}" for m,e in E]+["V "+t for t in V])+"\n"
    out=subprocess.run(["/tmp/numf/number_check"],input=inp,capture_output=True,text=True).stdout.splitlines()
    assert len(out)==len(E)+len(V)
    for (m,e),line in zip(E,out[:len(E)]):
        ok,b=line.split()
        if ok=="0": continue
        want=float(f"{m}e{e}")
        if int(b,16)!=bits64(want): bad+=1; print("E MISMATCH",m,e,b,hex(bits64(want)))
    for t,line in zip(V,out[len(E):]):
        q,b,tiny=line.split()
        if q=="0":
            if 'e' not in t: declined_real+=1; print("declined",t)
            continue
        want=np.float32(float(t))
        if float(t)!=0.0 and want==0.0: continue
        if int(b,16)!=bits32(want): bad+=1; print("V MISMATCH",t,b,hex(bits32(want)))
print("seed",seed,"bad",bad,"declined_realistic",declined_real)
