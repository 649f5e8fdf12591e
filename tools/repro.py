"""dev: run one decide at a given shape/variant and compare with the oracle"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpu_pruner_b200 as g
from oracle import oracle_c
variant, P, G, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
eng = g.IdleEngine(device=0, kernel=variant)
u = torch.empty((P, G, T), dtype=torch.float32, device="cuda:0")
eng.synth_fill(5, 0, u, 0, P, G, T)
db = torch.zeros((P + 31) // 32, dtype=torch.int32, device="cuda:0")
try:
    r = eng.decide_ptr(u, P, G, T, db)
    exp = oracle_c.decide_synth(5, 0, P, G, T)
    ok = np.array_equal(db.cpu().numpy().view(np.uint32), exp["decision_bits"])
    print(variant, P, G, T, "kernel_ms", round(r.kernel_ms, 4), "parity", ok, flush=True)
except Exception as e:
    print(variant, P, G, T, "FAILED", e, flush=True)
