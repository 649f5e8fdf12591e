"""Daemon-mode steady state (BASELINE config #5, per-GPU share): the window is resident in HBM as a
ring; every tick appends the columns that arrived during --check-interval (180 s @ 1 s = 180 columns
per series) and rescans.  Reports tick latency split into ingest (H2D + scatter) and decision.

    python tools/daemon_bench.py [--pods 312500 --gpus-per-pod 4 --samples 7200 --new 180 --ticks 5]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import gpu_pruner_b200 as g  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pods", type=int, default=312500)
    ap.add_argument("--gpus-per-pod", type=int, default=4)
    ap.add_argument("--samples", type=int, default=7200)
    ap.add_argument("--new", type=int, default=180)
    ap.add_argument("--ticks", type=int, default=5)
    ap.add_argument("--index", action="store_true", help="GPR_F_BLOCK_INDEX: decide on 64-sample block maxima")
    a = ap.parse_args()
    P, G, T, n_new = a.pods, a.gpus_per_pod, a.samples, a.new
    seed = 0x5EED0005
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        P = (P + 31) // 32 * 32      # shards are whole bitmap words
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    eng = g.IdleEngine(device=local)
    if world > 1:      # BASELINE config #5 proper: every rank holds its shard resident, fused bitmap exchange
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(dev))
        handles = [None] * world
        dist.all_gather_object(handles, eng.p2p_init(rank, world, P))
        eng.p2p_attach(handles)
    t0 = time.perf_counter()
    eng.resident_init(P, G, T, block_index=a.index)
    u_ptr, _, ld = eng.resident_planes()
    eng.synth_fill(seed, 0, u_ptr, rank * P, P, G, T)   # the history a first full range query would load
    eng.resident_reindex()
    torch.cuda.synchronize()
    init_s = time.perf_counter() - t0
    # the columns of the next ticks, staged in pinned host memory like an ingest thread would
    cols = eng.host_array((P, G, n_new), np.float32)
    stage = torch.empty((P, G, n_new), dtype=torch.float32, device=dev)
    W = (P + 31) // 32 * world
    dbits = np.zeros(W, np.uint32)
    res = []
    for tick in range(a.ticks):
        eng.synth_fill(seed + 1 + tick, 0, stage, rank * P, P, G, n_new)
        eng.memcpy(cols, stage, cols.nbytes, 0, 1)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        eng.append(cols, None, n_new)                    # H2D 4*P*G*n_new bytes + ring scatter
        t1 = time.perf_counter()
        r = eng.decide_ptr(None, 0, 0, 0, dbits, in_kind=0, out_kind=0, resident=True)
        t2 = time.perf_counter()
        res.append((t1 - t0, t2 - t1, r.kernel_ms, r.n_decisions))
    ing = np.median([x[0] for x in res]) * 1e3
    dec = np.median([x[1] for x in res]) * 1e3
    ker = np.median([x[2] for x in res])
    bytes_scan = 4.0 * P * G * T
    if world > 1:       # the slowest rank defines the tick
        t = torch.tensor([ing, dec, ker], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ing, dec, ker = (float(x) for x in t.tolist())
    if rank != 0:
        eng.close()
        return
    out = {"config": f"{world} GPU(s) x [{P} pods x {G} x {T} resident ({bytes_scan / 1e9:.1f} GB)], {n_new} new columns/tick",
           "n_gpus": world, "block_index": bool(a.index),
           "init_fill_s": init_s, "ingest_ms": ing, "ingest_h2d_bytes": int(cols.nbytes),
           "ingest_GBps": cols.nbytes / ing / 1e6, "decide_ms": dec, "kernel_ms": ker,
           "rescan_GBps": bytes_scan / ker / 1e6, "tick_ms": ing + dec,
           "duty_cycle_at_180s": (ing + dec) / 180e3, "samples_per_s": world * P * G * T / ((ing + dec) * 1e-3),
           "pod_decisions_per_s": world * P / ((ing + dec) * 1e-3), "n_decisions_last_rank0": int(res[-1][3])}
    print(json.dumps(out))
    eng.close()


if __name__ == "__main__":
    main()
