"""GPU, >= 2 devices: pod-sharded decision with the NCCL allgather of the packed bitmap, one process
per GPU, through the C ABI (gpr_comm_unique_id / gpr_comm_init / gpr_decide).  Skips on a 1-GPU box;
run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`."""
import os
import socket
import sys
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank(rank, world, port, total, G, T, seed, out_dir, mode):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import gpu_pruner_b200 as g
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    # the fused exchange has two wire protocols: tagged 64-bit slots (default) and data + fence + flag
    os.environ["GPR_EXCHANGE"] = {"p2p-flags": "flags", "p2p-pipelined": "pipelined"}.get(mode, "ll")
    eng = g.IdleEngine(device=rank, max_pods=20000, max_gpus=G, max_samples=T)
    sh = g.shard_pods(total, rank, world)
    P = sh.pods_per_rank
    if mode == "nccl":      # one ncclAllGather of the packed words per decision
        uid = [eng.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.comm_init(uid[0], rank, world)
    else:                   # fused: the folding CTA writes the words into every peer over NVLink
        handles = [None] * world
        dist.all_gather_object(handles, eng.p2p_init(rank, world, P))
        eng.p2p_attach(handles)
    dev = f"cuda:{rank}"
    u = torch.full((P, G, T), float("nan"), dtype=torch.float32, device=dev)
    e = torch.zeros(P, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()   # torch's fills before the engine's stream touches the buffers
    eng.synth_fill(seed, 0, u, sh.pod_begin, sh.pods_real, G, T)
    eng.synth_eligible(seed, e, sh.pod_begin, sh.pods_real)
    db = torch.zeros(world * P // 32, dtype=torch.int32, device=dev)
    cb = torch.zeros(world * P // 32, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    for it in range(3):   # repeated calls reuse the communicator / exchange block
        r = eng.decide_ptr(u, P, G, T, db, eligible=e, candidate_bits=cb)
    # back-to-back async decisions (overlapped launches) must still exchange step by step
    db2 = torch.zeros_like(db)
    torch.cuda.synchronize()
    for it in range(8):
        eng.decide_ptr(u, P, G, T, db2, eligible=e, blocking=False)
    eng.sync()
    assert torch.equal(db, db2)
    # ... and in order: two different windows alternate into ONE output buffer, the last one launched must be the
    # one that stays (a fold may produce, send and collect its words while its predecessor is still exchanging, but
    # it writes the caller's buffers after it); one rank launches late so that the ranks run steps apart
    u2 = torch.full((P, G, T), float("nan"), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    eng.synth_fill(seed + 1, 0, u2, sh.pod_begin, sh.pods_real, G, T)
    db_b = torch.zeros_like(db)
    eng.decide_ptr(u2, P, G, T, db_b, eligible=e)
    assert not torch.equal(db, db_b)
    for n_calls, want in ((9, db), (12, db_b)):
        db3 = torch.zeros_like(db)
        torch.cuda.synchronize()
        dist.barrier()
        for it in range(n_calls):
            if rank == world - 1 and it in (3, 4, 7):
                time.sleep(0.003)
            eng.decide_ptr(u if it % 2 == 0 else u2, P, G, T, db3, eligible=e, blocking=False)
        eng.sync()
        assert torch.equal(db3, want), (mode, n_calls)
    del u2
    # the collective timer: device-side rendezvous, then per-decision completion stamps
    eng.timer_begin()
    for it in range(5):
        eng.decide_ptr(u, P, G, T, db2, eligible=e, blocking=False)
    ms = eng.timer_end()
    eng.sync()
    t0, st = eng.step_stamps()
    assert len(st) == 5 and t0 > 0 and int(st[0]) > t0 and np.all(np.diff(st.astype(np.int64)) > 0)
    assert 0 < (int(st[-1]) - t0) / 1e6 <= ms * 1.5 + 0.1
    assert torch.equal(db, db2)
    np.save(os.path.join(out_dir, f"d_{rank}.npy"), db.cpu().numpy().view(np.uint32))
    np.save(os.path.join(out_dir, f"c_{rank}.npy"), cb.cpu().numpy().view(np.uint32))
    np.save(os.path.join(out_dir, f"n_{rank}.npy"), np.array([r.n_series, r.n_candidates, r.n_decisions]))
    # host outputs through the same path
    hd = np.zeros(world * P // 32, np.uint32)
    eng.decide_ptr(u, P, G, T, hd, eligible=e, out_kind=0)
    np.save(os.path.join(out_dir, f"h_{rank}.npy"), hd)
    # host window (chunked H2D path) through the same exchange
    if P <= 20000:
        hw = np.zeros(world * P // 32, np.uint32)
        eng.decide_ptr(u.cpu().numpy(), P, G, T, hw, eligible=e.cpu().numpy(), in_kind=0, out_kind=0)
        assert np.array_equal(hw, hd)
    dist.barrier()
    if mode == "nccl":
        eng.comm_destroy()
    eng.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["nccl", "p2p", "p2p-flags", "p2p-pipelined"])
@pytest.mark.parametrize("total", [5000, 64 * 1000 + 7])
def test_sharded_decision_allgather(total, mode, tmp_path, oracle_c):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    G, T, seed = 4, 600, 0x5EED0004
    mp.spawn(_rank, args=(world, _free_port(), total, G, T, seed, str(tmp_path), mode), nprocs=world, join=True)
    ref = oracle_c.decide_synth(seed, 0, total, G, T, use_elig=True)
    W = (total + 31) // 32
    n = np.zeros(3, np.int64)
    for rank in range(world):
        for tag, key in (("d", "decision_bits"), ("h", "decision_bits"), ("c", "candidate_bits")):
            full = np.load(tmp_path / f"{tag}_{rank}.npy")
            assert np.array_equal(full[:W], ref[key]), (tag, rank)
            assert not full[W:].any()
        n += np.load(tmp_path / f"n_{rank}.npy")
    assert tuple(n) == (ref["n_series"], ref["n_candidates"], ref["n_decisions"])
