"""dev: end-to-end host-window step time vs staging chunk size (GPR_CHUNK_MB)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpu_pruner_b200 as g
P, G, T = 10000, 4, 1800
eng = g.IdleEngine(device=0, max_pods=P, max_gpus=G, max_samples=T)
d = torch.empty((P, G, T), dtype=torch.float32, device="cuda:0")
eng.synth_fill(2, 0, d, 0, P, G, T)
h = eng.host_array((P, G, T), np.float32)
eng.memcpy(h, d, h.nbytes, 0, 1)
bits = eng.host_array((313,), np.uint32)
for _ in range(3):
    eng.decide_ptr(h, P, G, T, bits, in_kind=0, out_kind=0)
t0 = time.perf_counter()
n = 30
for _ in range(n):
    eng.decide_ptr(h, P, G, T, bits, in_kind=0, out_kind=0)
dt = (time.perf_counter() - t0) / n
print(f"GPR_CHUNK_MB={os.environ.get('GPR_CHUNK_MB','8')}: {dt*1e3:.3f} ms/step  {h.nbytes/dt/1e9:.1f} GB/s")
