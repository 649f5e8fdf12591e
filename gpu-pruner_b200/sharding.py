"""Pod-axis sharding for multi-GPU runs.

Every pod's verdict depends only on its own rows (no cross-pod term anywhere in
/root/reference/gpu-pruner/src/query.promql.j2: `sum by` and `unless on (pod, namespace)` are
both scoped inside one pod), so ranks own contiguous pod ranges and exchange nothing but the
packed decision bitmap.  Shards are whole bitmap words so the allgather output is the global
bitmap with no re-packing: P_s = 32 * ceil(P / (32 * world)); pods past the end are padding
(no samples, ineligible) and their bits are zero.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class Shard:
    rank: int
    world: int
    pods_total: int
    pods_per_rank: int   # P_s, multiple of 32, identical on every rank
    pod_begin: int       # first global pod of this rank
    pods_real: int       # pods of this rank that exist (the rest is padding)

    @property
    def words_per_rank(self) -> int:
        return self.pods_per_rank // 32

    @property
    def pods_padding(self) -> int:
        return self.pods_per_rank - self.pods_real


def shard_pods(pods_total: int, rank: int, world: int) -> Shard:
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    if pods_total < 0:
        raise ValueError("pods_total < 0")
    per = 32 * ((pods_total + 32 * world - 1) // (32 * world)) if pods_total else 0
    begin = rank * per
    real = max(0, min(per, pods_total - begin))
    return Shard(rank, world, pods_total, per, begin, real)


def global_pod(shard_words_rank_major, pods_per_rank: int, pods_total: int):
    """Indices of set bits of the gathered rank-major bitmap, as global pod numbers."""
    import numpy as np
    w = np.ascontiguousarray(shard_words_rank_major, dtype="<u4")
    flat = np.unpackbits(w.view(np.uint8), bitorder="little")
    idx = np.flatnonzero(flat)
    return idx[idx < pods_total] if pods_per_rank else idx
