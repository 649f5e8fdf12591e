"""dev: per-CTA timeline of k_reduce_ldg on C2 using the instrumented build (make -C csrc dbg)"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpu_pruner_b200 import ffi
ffi.lib_path = lambda: os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpu-pruner_b200", "libgpr_dbg.so")
import gpu_pruner_b200 as g
P, G, T = 10000, 4, 1800
eng = g.IdleEngine(device=0, kernel="ldg")
wins = []
for i in range(4):
    u = torch.empty((P, G, T), dtype=torch.float32, device="cuda:0"); eng.synth_fill(2 + i, 0, u, 0, P, G, T); wins.append(u)
db = torch.zeros((P + 31) // 32, dtype=torch.int32, device="cuda:0")
for i in range(6):
    eng.decide_ptr(wins[i % 4], P, G, T, db, blocking=False)
eng.sync()
lib = ffi.load()
n = 296
buf = np.zeros(4 * n, np.uint64)
lib.gpr_debug_timeline.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
assert lib.gpr_debug_timeline(eng.handle, buf.ctypes.data, n) == 0
tl = buf.reshape(n, 4).astype(np.int64)
t0 = tl[:, 1].min()
start, stream_end, exit_ = tl[:, 1] - t0, tl[:, 2] - t0, tl[:, 3] - t0
print("CTAs per SM:", np.bincount(np.bincount(tl[:, 0].astype(int), minlength=148)))
print("start  ns: min %d  p50 %d  p99 %d  max %d" % (start.min(), np.median(start), np.percentile(start, 99), start.max()))
print("stream ns: min %d  p50 %d  p99 %d  max %d" % (stream_end.min(), np.median(stream_end), np.percentile(stream_end, 99), stream_end.max()))
print("exit   ns: min %d  p50 %d  p99 %d  max %d" % (exit_.min(), np.median(exit_), np.percentile(exit_, 99), exit_.max()))
late = np.argsort(start)[-10:]
print("latest starters (cta, smid, start, stream_end, exit):")
for c in late:
    print("  ", c, tl[c, 0], start[c], stream_end[c], exit_[c])
last = np.argmax(exit_)
print("last exit: cta", last, "sm", tl[last, 0], "start", start[last], "stream_end", stream_end[last], "exit", exit_[last])
