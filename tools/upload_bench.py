"""Developer benchmark of gpr_text_scan (upload + scan pipeline) alone: pageable vs pinned text, GB/s.
    GPR_TEXT_UPLOAD_THREADS=8 GPR_TEXT_CHUNK_MB=4 python tools/upload_bench.py [--pods 10000]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pods", type=int, default=10000)
    ap.add_argument("--samples", type=int, default=1800)
    ap.add_argument("--reps", type=int, default=4)
    a = ap.parse_args()
    import hostlib as H
    import gpu_pruner_b200 as g
    lib = H.lib()
    lib.gph_synth_response.restype = C.c_longlong
    need = -lib.gph_synth_response(a.pods, 4, a.samples, C.c_longlong(1_700_000_000), C.c_ulonglong(7), None, C.c_longlong(0))
    eng = g.IdleEngine(device=0)
    pinned = eng.host_array((need,), np.uint8)
    n = lib.gph_synth_response(a.pods, 4, a.samples, C.c_longlong(1_700_000_000), C.c_ulonglong(7),
                               pinned.ctypes.data_as(C.c_char_p), C.c_longlong(need))
    pageable = np.empty(n, np.uint8)
    pageable[:] = pinned[:n]
    out = {"bytes": int(n), "threads": os.environ.get("GPR_TEXT_UPLOAD_THREADS", "8"), "chunk_mb": os.environ.get("GPR_TEXT_CHUNK_MB", "4")}
    for name, buf in (("pinned", pinned), ("pageable", pageable)):
        best = 1e9
        for _ in range(a.reps):
            t0 = time.perf_counter()
            o, c = eng.text_scan(buf, n_bytes=n)
            best = min(best, time.perf_counter() - t0)
        out[name + "_ms"] = round(best * 1e3, 2)
        out[name + "_GBps"] = round(n / best / 1e9, 2)
        out[name + "_series"] = len(o)
    eng.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
