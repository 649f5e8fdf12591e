"""Developer benchmark of the ingest step: Prometheus matrix JSON -> dense tensor.

Generates a synthetic compact response (P pods x G GPUs x n one-second samples), writes it as a
file:// fixture and runs the `gpu-pruner` binary on it twice — GPR_INGEST=cpu (threaded CPU text parser,
window uploaded by gpr_decide) and GPR_INGEST=gpu (text parsed on the GPU into HBM) — reporting the ingest
times the binary logs, the wall time of the whole tick and that both runs reach the same verdict counts.
Also times the two device passes alone through the C ABI.  Not the judged benchmark (bench.py)."""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

T_END = 1_700_000_000


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pods", type=int, default=2000)
    ap.add_argument("--gpus", type=int, default=4)
    ap.add_argument("--samples", type=int, default=1800)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import hostlib as H
    import gpu_pruner_b200 as g
    lib = H.lib()
    lib.gph_synth_response.restype = C.c_longlong
    need = -lib.gph_synth_response(a.pods, a.gpus, a.samples, C.c_longlong(T_END), C.c_ulonglong(7), None, C.c_longlong(0))
    eng = g.IdleEngine(device=0)
    h_text = eng.host_array((need,), np.uint8)          # pinned
    n = lib.gph_synth_response(a.pods, a.gpus, a.samples, C.c_longlong(T_END), C.c_ulonglong(7),
                               h_text.ctypes.data_as(C.c_char_p), C.c_longlong(need))
    assert n > 0
    out = {"config": f"{a.pods} pods x {a.gpus} GPUs x {a.samples} samples", "text_bytes": int(n),
           "samples": a.pods * a.gpus * a.samples, "bytes_per_sample": n / (a.pods * a.gpus * a.samples)}

    # ---- the two device passes alone (C ABI, pinned text) ---------------------------------------------
    best_scan = best_parse = 1e9
    W = a.samples
    for _ in range(a.reps):
        t0 = time.perf_counter()
        opens, closes = eng.text_scan(h_text, n_bytes=n)
        best_scan = min(best_scan, time.perf_counter() - t0)
        spans = np.zeros(len(opens), eng.SPAN_DTYPE)
        spans["begin"] = opens + 12
        spans["end"] = closes[np.searchsorted(closes, opens + 12)] + 2
        spans["row"] = np.arange(len(opens))
        t0 = time.perf_counter()
        res = eng.text_parse(spans, T_END, 1, W, len(opens))
        best_parse = min(best_parse, time.perf_counter() - t0)
    out["abi"] = {"scan_ms_incl_h2d": best_scan * 1e3, "scan_GBps": n / best_scan / 1e9,
                  "parse_ms_incl_nan_fill": best_parse * 1e3, "parse_GBps": n / best_parse / 1e9,
                  "parse_Msamples_per_s": out["samples"] / best_parse / 1e6,
                  "hard_spans": int(np.count_nonzero(res["flags"] & 2)), "n_in": int(res["n_in"].sum())}

    # ---- the product path: the binary on a file:// fixture, both ingestors ------------------------------
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "util.json"), "wb") as f:
            f.write(h_text[:n].tobytes())
        with open(os.path.join(d, "query.json"), "w") as f:
            json.dump({"end": T_END, "step": 1}, f)
        for mode in ("cpu", "gpu"):
            best = None
            for _ in range(a.reps):
                t0 = time.perf_counter()
                p = subprocess.run([H.BIN, "--prometheus-url", f"file://{d}", "-t", str(a.samples // 60), "-l", "json",
                                    "--now", str(T_END)], capture_output=True, text=True, timeout=900,
                                   env=dict(os.environ, GPR_INGEST=mode))
                wall = time.perf_counter() - t0
                logs = [json.loads(l) for l in p.stderr.splitlines() if l.startswith("{")]
                msgs = [l["fields"]["message"] for l in logs]
                note = next(m for m in msgs if m.startswith("Device ingest"))
                ms = float(note.rsplit(" ms", 1)[0].rsplit(" ", 1)[1]) if mode == "cpu" else float(
                    note.split(" window in ")[1].split(" ms")[0])
                verdict = next(m for m in msgs if m.startswith("Query returned"))
                if best is None or ms < best["ingest_ms"]:
                    best = {"ingest_ms": ms, "process_wall_s": wall, "note": note, "verdict": verdict}
            out[mode] = best
        out["same_verdict"] = out["cpu"]["verdict"] == out["gpu"]["verdict"]
        out["ingest_speedup"] = out["cpu"]["ingest_ms"] / out["gpu"]["ingest_ms"]
        out["host_threads"] = os.cpu_count()
    eng.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
