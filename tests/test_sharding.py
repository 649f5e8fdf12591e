"""CPU: pod-axis sharding logic, single process and world_size-2 over gloo.

The N>1 data path has no exchange except the allgather of the packed bitmap.
Here each rank computes its shard's bitmap with the CPU oracle (test infrastructure standing in
for the GPU), the words are allgathered over gloo exactly as bench.py / libgpr do over NCCL, and
the result must equal the single-process decision over the whole window."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_arithmetic():
    import gpu_pruner_b200 as g
    for total in (0, 1, 31, 32, 33, 100, 10000, 80000, 1_000_000, 2_500_000):
        for world in (1, 2, 3, 4, 8):
            shards = [g.shard_pods(total, r, world) for r in range(world)]
            per = shards[0].pods_per_rank
            assert per % 32 == 0 and all(s.pods_per_rank == per for s in shards)
            assert per * world >= total and (total == 0 or per * world - total < 32 * world)
            assert sum(s.pods_real for s in shards) == total
            assert [s.pod_begin for s in shards] == [r * per for r in range(world)]
            # real pods are a prefix of every shard, padding only at the tail of the job
            covered = []
            for s in shards:
                covered += list(range(s.pod_begin, s.pod_begin + s.pods_real))
            assert covered == list(range(total))
    with pytest.raises(ValueError):
        g.shard_pods(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, total, G, T, seed, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import gpu_pruner_b200 as g
    from oracle import oracle_c
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    sh = g.shard_pods(total, rank, world)
    # this rank's window: real pods from the shared synthetic universe, padding = no samples, ineligible
    u = np.full((sh.pods_per_rank, G, T), np.nan, np.float32)
    e = np.zeros(sh.pods_per_rank, np.uint8)
    if sh.pods_real:
        u[: sh.pods_real] = oracle_c.synth_fill(seed, 0, sh.pod_begin, sh.pods_real, G, T, n_threads=2)
        e[: sh.pods_real] = oracle_c.synth_eligible(seed, sh.pod_begin, sh.pods_real)
    r = oracle_c.decide(u, None, e)
    local = torch.from_numpy(r["decision_bits"].astype(np.int32))
    assert local.numel() == sh.words_per_rank
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)          # the one collective of the path
    full = torch.cat(gathered).numpy().astype(np.uint32)
    cnt = torch.tensor([r["n_decisions"]], dtype=torch.int64)
    dist.all_reduce(cnt)
    np.save(os.path.join(out_dir, f"bits_{rank}.npy"), full)
    np.save(os.path.join(out_dir, f"cnt_{rank}.npy"), cnt.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [1000, 4097])
def test_world2_gloo_allgather_matches_single_process(total, tmp_path, oracle_c):
    import torch.multiprocessing as mp
    import gpu_pruner_b200 as g
    world, G, T, seed = 2, 4, 120, 0x5EED0004
    port = _free_port()
    mp.spawn(_rank_main, args=(world, port, total, G, T, seed, str(tmp_path)), nprocs=world, join=True)
    ref = oracle_c.decide_synth(seed, 0, total, G, T, use_elig=True, n_threads=2)
    sh = g.shard_pods(total, 0, world)
    for rank in range(world):
        full = np.load(tmp_path / f"bits_{rank}.npy")
        assert full.size == world * sh.words_per_rank
        got = g.sharding.global_pod(full, sh.pods_per_rank, total)
        want = np.flatnonzero(np.unpackbits(ref["decision_bits"].view(np.uint8), bitorder="little"))
        assert np.array_equal(got, want)
        # rank-major words ARE the global bitmap (shards are whole words), padding bits zero
        W = (total + 31) // 32
        assert np.array_equal(full[:W], ref["decision_bits"]) and not full[W:].any()
        assert int(np.load(tmp_path / f"cnt_{rank}.npy")[0]) == ref["n_decisions"]
